// mas_consumers.cu -- the callers either side of the alignment path:
//
//   mas_path_durations   w = attn.sum(2)                               SynthesizerTrn.py:237
//   mas_expand_prior     einsum('bctn,bdn->bdt', attn, m_p / logs_p)   SynthesizerTrn.py:247-248, 308-310
//   mas_generate_path    commons.generate_path(duration, mask)         commons.py:101-117
//
// The first two consume the compact per-frame index [B][T_y] (text position of every frame, -1 on padded
// frames) that mas_maximum_path can emit instead of / beside the dense 0/1 path: the sum over frames becomes a
// histogram and the one-hot GEMM becomes a gather, so the dense 50 MB path is never read back (the reference
// reads it three times and runs two [T_y x T_x].[T_x x C] GEMMs on it).  All three are HBM-bound.
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"

namespace mas {

// ---- durations: one CTA per utterance, shared-memory histogram ----------------------------------
__global__ void __launch_bounds__(256) mas_durations_kernel(const int32_t* __restrict__ index, float* __restrict__ w,
                                                            int T_y, int T_x) {
  extern __shared__ int hist[];
  const int b = blockIdx.x;
  for (int x = threadIdx.x; x < T_x; x += blockDim.x) hist[x] = 0;
  __syncthreads();
  const int32_t* idx = index + static_cast<size_t>(b) * T_y;
  for (int y = threadIdx.x; y < T_y; y += blockDim.x) {
    const int x = idx[y];
    if (x >= 0 && x < T_x) atomicAdd(&hist[x], 1);
  }
  __syncthreads();
  // counts are exact integers, as is the reference's float sum of 0/1 values
  for (int x = threadIdx.x; x < T_x; x += blockDim.x) w[static_cast<size_t>(b) * T_x + x] = static_cast<float>(hist[x]);
}

// ---- prior expansion: out[b,c,y] = src[b,c,index[b,y]] (0 on padded frames) ------------------------
// grid (B, ceil(C / CPB)); a CTA stages CPB channel rows of the source(s) in shared memory and streams
// CPB x T_y outputs with coalesced stores.  Exactly the einsum's result: one 1 per frame, every other
// product is an exact zero.
constexpr int kCPB = 8;
__global__ void __launch_bounds__(256) mas_expand_kernel(const int32_t* __restrict__ index, const float* __restrict__ a,
                                                         const float* __restrict__ bsrc, float* __restrict__ a_out,
                                                         float* __restrict__ b_out, int C, int T_y, int T_x) {
  extern __shared__ float rows[];  // [nsrc][kCPB][T_x]
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * kCPB;
  const int nc = min(kCPB, C - c0);
  const int nsrc = bsrc ? 2 : 1;
  for (int s = 0; s < nsrc; ++s) {
    const float* src = (s ? bsrc : a) + (static_cast<size_t>(b) * C + c0) * T_x;
    for (int i = threadIdx.x; i < nc * T_x; i += blockDim.x) rows[s * kCPB * T_x + i] = src[i];
  }
  __syncthreads();
  const int32_t* idx = index + static_cast<size_t>(b) * T_y;
  for (int y = threadIdx.x; y < T_y; y += blockDim.x) {
    const int x = idx[y];
    const bool ok = x >= 0 && x < T_x;
    for (int s = 0; s < nsrc; ++s) {
      float* out = (s ? b_out : a_out) + (static_cast<size_t>(b) * C + c0) * T_y + y;
      const float* r = rows + s * kCPB * T_x;
#pragma unroll
      for (int c = 0; c < kCPB; ++c)
        if (c < nc) out[static_cast<size_t>(c) * T_y] = ok ? r[c * T_x + x] : 0.0f;
    }
  }
}

// ---- generate_path (commons.py:101-117) ------------------------------------------------------------
// path[b,y,x] = ((y < cum[x]) - (y < cum[x-1])) * mask[b,y,x], cum = cumsum(duration[b,:]) accumulated
// sequentially in fp32 (torch's CPU order; durations are integers after ceil(), so any order is exact).
constexpr int kGpRows = 32;
__global__ void __launch_bounds__(256) mas_generate_path_kernel(const float* __restrict__ duration,
                                                                const float* __restrict__ mask, int64_t msb, int64_t msy,
                                                                int64_t msx, float* __restrict__ path, int T_y, int T_x) {
  extern __shared__ float cum[];  // [T_x + 1], cum[0] = "nothing before the first token"
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    float acc = 0.0f;
    const float* d = duration + static_cast<size_t>(b) * T_x;
    for (int x = 0; x < T_x; ++x) {
      acc += d[x];
      cum[x + 1] = acc;
    }
  }
  __syncthreads();
  const int y0 = blockIdx.y * kGpRows;
  const int y1 = min(T_y, y0 + kGpRows);
  for (int i = threadIdx.x; i < (y1 - y0) * T_x; i += blockDim.x) {
    const int y = y0 + i / T_x, x = i - (i / T_x) * T_x;
    const float fy = static_cast<float>(y);
    const float cur = fy < cum[x + 1] ? 1.0f : 0.0f;
    const float prev = (x > 0 && fy < cum[x]) ? 1.0f : 0.0f;  // F.pad: nothing before token 0
    const float m = mask[b * msb + y * msy + x * msx];
    path[(static_cast<size_t>(b) * T_y + y) * T_x + x] = (cur - prev) * m;
  }
}

int path_durations(const int32_t* index, float* w, int B, int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || T_y <= 0 || T_x <= 0 || T_x > 12000) return MAS_E_BAD_SHAPE;
  if (!index || !w) return MAS_E_NULL;
  mas_durations_kernel<<<B, 256, static_cast<size_t>(T_x) * sizeof(int), st>>>(index, w, T_y, T_x);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int expand_prior(const int32_t* index, const float* m_p, const float* logs_p, float* m_out, float* logs_out, int B, int C,
                 int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0 || T_x > 2048 || B > 65535) return MAS_E_BAD_SHAPE;
  if (!index || !m_p || !m_out || ((logs_p == nullptr) != (logs_out == nullptr))) return MAS_E_NULL;
  const size_t smem = static_cast<size_t>(logs_p ? 2 : 1) * kCPB * T_x * sizeof(float);
  static std::atomic<uint64_t> attr{0};
  if (cudaError_t e = ensure_dyn_smem(mas_expand_kernel, 160 * 1024, attr); e != cudaSuccess) return static_cast<int>(e);
  mas_expand_kernel<<<dim3(B, (C + kCPB - 1) / kCPB), 256, smem, st>>>(index, m_p, logs_p, m_out, logs_out, C, T_y, T_x);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int generate_path(const float* duration, const float* mask, int64_t msb, int64_t msy, int64_t msx, float* path, int B,
                  int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || T_y <= 0 || T_x <= 0 || T_x > 12000 || (T_y + kGpRows - 1) / kGpRows > 65535) return MAS_E_BAD_SHAPE;
  if (!duration || !mask || !path) return MAS_E_NULL;
  mas_generate_path_kernel<<<dim3(B, (T_y + kGpRows - 1) / kGpRows), 256, static_cast<size_t>(T_x + 1) * sizeof(float), st>>>(
      duration, mask, msb, msy, msx, path, T_y, T_x);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace mas

// ---- kl_loss on the compact index (losses.py:43-60 fused with the prior expansion) -------------------
// sum over (b,c,y) of (logs_p - logs_q - 0.5 + 0.5 (z_p - m_p)^2 exp(-2 logs_p)) * z_mask[b,y], with
// m_p/logs_p gathered along the alignment on the fly (m_p[b,c,index[b,y]]) instead of read from expanded
// [B,C,T_y] tensors; also sum of z_mask.  Accumulates in double: out[0] += kl sum, out[1] += mask sum
// (out[1] is added once, by the c0 == 0 CTAs).
namespace mas {

__global__ void __launch_bounds__(256) mas_kl_index_kernel(const int32_t* __restrict__ index, const float* __restrict__ z_p,
                                                           const float* __restrict__ logs_q, const float* __restrict__ m_p,
                                                           const float* __restrict__ logs_p, const float* __restrict__ z_mask,
                                                           double* __restrict__ out, int C, int T_y, int T_x) {
  extern __shared__ float rows[];  // [2][kCPB][T_x]
  __shared__ double red[2][8];
  const int b = blockIdx.x;
  const int c0 = blockIdx.y * kCPB;
  const int nc = min(kCPB, C - c0);
  for (int i = threadIdx.x; i < nc * T_x; i += blockDim.x) {
    rows[i] = m_p[(static_cast<size_t>(b) * C + c0) * T_x + i];
    rows[kCPB * T_x + i] = logs_p[(static_cast<size_t>(b) * C + c0) * T_x + i];
  }
  __syncthreads();
  double acc = 0.0, macc = 0.0;
  for (int y = threadIdx.x; y < T_y; y += blockDim.x) {
    const int x = index[static_cast<size_t>(b) * T_y + y];
    const float mk = z_mask[static_cast<size_t>(b) * T_y + y];
    const bool ok = x >= 0 && x < T_x;
    macc += mk;
    float part = 0.0f;
#pragma unroll
    for (int c = 0; c < kCPB; ++c) {
      if (c < nc) {
        const size_t o = (static_cast<size_t>(b) * C + c0 + c) * T_y + y;
        const float m = ok ? rows[c * T_x + x] : 0.0f;        // the einsum yields 0 on frames without a 1
        const float lp = ok ? rows[(kCPB + c) * T_x + x] : 0.0f;
        const float d = z_p[o] - m;
        part += (lp - logs_q[o] - 0.5f) + 0.5f * (d * d) * __expf(-2.0f * lp);
      }
    }
    acc += static_cast<double>(part * mk);
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    macc += __shfl_xor_sync(0xffffffffu, macc, o);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = acc;
    red[1][threadIdx.x >> 5] = macc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, m = 0.0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) {
      a += red[0][w];
      m += red[1][w];
    }
    atomicAdd(out, a);
    if (c0 == 0) atomicAdd(out + 1, m);
  }
}

int kl_from_index(const int32_t* index, const float* z_p, const float* logs_q, const float* m_p, const float* logs_p,
                  const float* z_mask, double* out, int B, int C, int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0 || T_x > 2048 || B > 65535) return MAS_E_BAD_SHAPE;
  if (!index || !z_p || !logs_q || !m_p || !logs_p || !z_mask || !out) return MAS_E_NULL;
  const size_t smem = static_cast<size_t>(2) * kCPB * T_x * sizeof(float);
  static std::atomic<uint64_t> attr{0};
  if (cudaError_t e0 = ensure_dyn_smem(mas_kl_index_kernel, 160 * 1024, attr); e0 != cudaSuccess) return static_cast<int>(e0);
  cudaError_t e = cudaMemsetAsync(out, 0, 2 * sizeof(double), st);
  if (e != cudaSuccess) return static_cast<int>(e);
  mas_kl_index_kernel<<<dim3(B, (C + kCPB - 1) / kCPB), 256, smem, st>>>(index, z_p, logs_q, m_p, logs_p, z_mask, out, C, T_y, T_x);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace mas
