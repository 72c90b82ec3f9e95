// mas_neg_cent.cu -- the Gaussian log-likelihood contraction of SynthesizerTrn.forward
// (reference SynthesizerTrn.py:223-232), fused into one pass:
//
//   neg_cent[b,t,s] = bias[b,s] + sum_d ( a2[b,d,t] * iv[b,d,s] + z[b,d,t] * mv[b,d,s] )
//     a2 = -0.5 z_p^2          (:227)      iv = exp(-2 logs_p)        (:223)
//     mv = m_p * iv            (:229)      bias = sum_d(-0.5 log 2pi - logs_p)   (:225)
//                                                + sum_d(-0.5 m_p^2 iv)          (:231)
//
// i.e. ONE GEMM with K = 2C (the two einsums concatenated along the channel axis) whose
// operands are produced on the fly from z_p / m_p / logs_p and whose epilogue adds the
// per-column bias -- instead of the reference's 2 bmm + ~10 elementwise/reduce launches.
//
// This file holds the fp32 CUDA-core implementation (impl 0): exact fp32 products, fp32
// accumulation.  It serves odd shapes and is the on-device cross-check of the tcgen05 version.
//
// The same kernel, instantiated with RND = 1 (fp16) or 2 (bf16), is the AUTOCAST-PARITY mode (SURVEY.md
// 8f rank 5): what the reference computes as trained, inside autocast(fp16_run)
// (train_and_evaluate.py:55, config_cje.yaml:11).  There the two einsums (:227, :229) run on operands
// cast to the low-precision type with fp32 accumulation and each einsum's OUTPUT is rounded to that type,
// while exp, pow and the two sums over channels (:223, :225, :231) stay fp32 and the four terms are added
// in fp32 in the order written (:232).  Here: operands rounded when staged, one accumulator per einsum,
// each rounded once in the epilogue, ((t1 + r(t2)) + r(t3)) + t4.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"

namespace mas {

constexpr int TM = 64;   // frames per CTA tile
constexpr int TN = 64;   // text positions per CTA tile
constexpr int TK = 16;   // channels per smem step

template <int RND>
__device__ __forceinline__ float rnd(float x) {
  if constexpr (RND == 1) return __half2float(__float2half_rn(x));
  if constexpr (RND == 2) return __bfloat162float(__float2bfloat16_rn(x));
  return x;
}

template <int RND>
__global__ void __launch_bounds__(256) neg_cent_simt_kernel(const float* __restrict__ z_p, const float* __restrict__ m_p,
                                                            const float* __restrict__ logs_p, float* __restrict__ out,
                                                            int C, int T_y, int T_x, int stats_lowp) {
  __shared__ float sA2[TK][TM];   // -0.5 z^2
  __shared__ float sZ[TK][TM];    // z
  __shared__ float sIv[TK][TN];   // exp(-2 logs)
  __shared__ float sMv[TK][TN];   // m * exp(-2 logs)
  __shared__ float sBias[TN];
  __shared__ float sBias4[TN];  // RND != 0: term 4 (:231) apart from term 1 (:225)

  const int b = blockIdx.z;
  const int t0 = blockIdx.y * TM;
  const int s0 = blockIdx.x * TN;
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // 4 text positions each
  const int ty = tid >> 4;   // 4 frames each

  const float* zb = z_p + static_cast<size_t>(b) * C * T_y;
  const float* mb = m_p + static_cast<size_t>(b) * C * T_x;
  const float* lb = logs_p + static_cast<size_t>(b) * C * T_x;

  float acc[4][4];
  float acc3[RND ? 4 : 1][4];  // RND != 0: the second einsum (:229) accumulates apart from the first (:227)
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = 0.0f;
      if constexpr (RND != 0) acc3[i][j] = 0.0f;
    }
  float bias = 0.0f;   // threads 0..TN-1 own one text position each
  float bias4 = 0.0f;
  const float kHalfLog2Pi = 0.91893853320467274178f;

  for (int d0 = 0; d0 < C; d0 += TK) {
    // stage operands: 16 channels x 64 positions each, 4 elements per thread
#pragma unroll
    for (int i = 0; i < (TK * TM) / 256; ++i) {
      const int e = tid + i * 256;
      const int d = e / TM, t = e % TM;
      const bool ok = (d0 + d < C) && (t0 + t < T_y);
      const float z = ok ? zb[static_cast<size_t>(d0 + d) * T_y + t0 + t] : 0.0f;
      sZ[d][t] = rnd<RND>(z);
      sA2[d][t] = rnd<RND>(-0.5f * (z * z));
    }
#pragma unroll
    for (int i = 0; i < (TK * TN) / 256; ++i) {
      const int e = tid + i * 256;
      const int d = e / TN, s = e % TN;
      const bool ok = (d0 + d < C) && (s0 + s < T_x);
      const float l = ok ? lb[static_cast<size_t>(d0 + d) * T_x + s0 + s] : 0.0f;
      const float m = ok ? mb[static_cast<size_t>(d0 + d) * T_x + s0 + s] : 0.0f;
      const float iv = ok ? expf(-2.0f * l) : 0.0f;
      sIv[d][s] = rnd<RND>(iv);
      sMv[d][s] = rnd<RND>(m * iv);
    }
    __syncthreads();
    if (tid < TN) {
      // bias terms of this channel block for text position s0+tid (:225, :231)
      if (s0 + tid < T_x) {
#pragma unroll
        for (int d = 0; d < TK; ++d) {
          if (d0 + d < C) {
            const float l = lb[static_cast<size_t>(d0 + d) * T_x + s0 + tid];
            const float m = mb[static_cast<size_t>(d0 + d) * T_x + s0 + tid];
            if constexpr (RND == 0) {
              bias += (-kHalfLog2Pi - l) + (-0.5f * (m * m)) * sIv[d][tid];
            } else {
              // :225 is plain arithmetic, so it runs in the dtype of logs_p: TextEncoder.proj returns the
              // low-precision type under autocast (TextEncoder.py:101-104) and each element is rounded to it
              // before sum() -- on autocast's fp32 list -- adds them up in fp32
              bias += stats_lowp ? rnd<RND>(-kHalfLog2Pi - l) : -kHalfLog2Pi - l;
              bias4 += (-0.5f * (m * m)) * expf(-2.0f * l);  // fp32 inverse variance, not the rounded operand
            }
          }
        }
      }
    }
#pragma unroll
    for (int d = 0; d < TK; ++d) {
      const float4 a2 = *reinterpret_cast<const float4*>(&sA2[d][ty * 4]);
      const float4 zz = *reinterpret_cast<const float4*>(&sZ[d][ty * 4]);
      const float4 iv = *reinterpret_cast<const float4*>(&sIv[d][tx * 4]);
      const float4 mv = *reinterpret_cast<const float4*>(&sMv[d][tx * 4]);
      const float a2v[4] = {a2.x, a2.y, a2.z, a2.w};
      const float zv[4] = {zz.x, zz.y, zz.z, zz.w};
      const float ivv[4] = {iv.x, iv.y, iv.z, iv.w};
      const float mvv[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j] = fmaf(a2v[i], ivv[j], acc[i][j]);
          if constexpr (RND == 0) acc[i][j] = fmaf(zv[i], mvv[j], acc[i][j]);
          else acc3[i][j] = fmaf(zv[i], mvv[j], acc3[i][j]);
        }
    }
    __syncthreads();
  }
  if (tid < TN) {
    sBias[tid] = bias;
    sBias4[tid] = bias4;
  }
  __syncthreads();

  float* ob = out + static_cast<size_t>(b) * T_y * T_x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = t0 + ty * 4 + i;
    if (t >= T_y) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = s0 + tx * 4 + j;
      if (s < T_x) {
        if constexpr (RND == 0)
          ob[static_cast<size_t>(t) * T_x + s] = acc[i][j] + sBias[tx * 4 + j];
        else  // :232, left to right
          ob[static_cast<size_t>(t) * T_x + s] = ((sBias[tx * 4 + j] + rnd<RND>(acc[i][j])) + rnd<RND>(acc3[i][j])) + sBias4[tx * 4 + j];
      }
    }
  }
}

static int g_impl = -1;  // -1 auto
void set_neg_cent_impl(int impl) { g_impl = impl; }

size_t neg_cent_scratch_bytes(int B, int C, int T_y, int T_x) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0) return 0;
  return neg_cent_tc_scratch_bytes(B, C, T_y, T_x) + 256;
}

int neg_cent(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
             int B, int C, int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0 || B > 65535) return MAS_E_BAD_SHAPE;
  if (!z_p || !m_p || !logs_p || !out) return MAS_E_NULL;
  if ((reinterpret_cast<uintptr_t>(z_p) | reinterpret_cast<uintptr_t>(m_p) | reinterpret_cast<uintptr_t>(logs_p) |
       reinterpret_cast<uintptr_t>(out)) & 3u)
    return MAS_E_ALIGN;
  // impl 1 (default): tcgen05 tensor-core GEMM with split-bf16 operands; impl 0: fp32 CUDA cores
  if (g_impl != 0) return neg_cent_tc(z_p, m_p, logs_p, out, scratch, scratch_bytes, B, C, T_y, T_x, st);
  dim3 grid((T_x + TN - 1) / TN, (T_y + TM - 1) / TM, B);
  if (grid.y > 65535) return MAS_E_BAD_SHAPE;
  neg_cent_simt_kernel<0><<<grid, 256, 0, st>>>(z_p, m_p, logs_p, out, C, T_y, T_x, 0);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return MAS_OK;
}

int neg_cent_autocast(const float* z_p, const float* m_p, const float* logs_p, float* out, int gemm_dtype, int stats_lowp,
                      int B, int C, int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0 || B > 65535) return MAS_E_BAD_SHAPE;
  if (!z_p || !m_p || !logs_p || !out) return MAS_E_NULL;
  if (gemm_dtype != MAS_F16 && gemm_dtype != MAS_BF16) return MAS_E_BAD_DTYPE;
  dim3 grid((T_x + TN - 1) / TN, (T_y + TM - 1) / TM, B);
  if (grid.y > 65535) return MAS_E_BAD_SHAPE;
  if (gemm_dtype == MAS_F16) neg_cent_simt_kernel<1><<<grid, 256, 0, st>>>(z_p, m_p, logs_p, out, C, T_y, T_x, stats_lowp);
  else neg_cent_simt_kernel<2><<<grid, 256, 0, st>>>(z_p, m_p, logs_p, out, C, T_y, T_x, stats_lowp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return MAS_OK;
}

}  // namespace mas
