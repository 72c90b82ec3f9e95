// mas_path.cu -- Monotonic Alignment Search on B200 (sm_100a): forward dynamic program,
// backtrack and dense path write-out.  Replaces monotonic_align/core.pyx:7-42 and the host
// marshalling of monotonic_align/__init__.py:14-20 of the reference.
//
// Three kernels, chained with programmatic dependent launch (PDL):
//
//   K1 mas_forward    one CTA per utterance.  A producer warp streams the utterance's neg_cent
//                     rows into a shared-memory ring with 1-D bulk async copies (TMA engine,
//                     mbarrier completion).  W "DP" warps own 32*K text columns each and sweep
//                     the mel frames in order; the x-1 neighbour comes from __shfl_up_sync inside
//                     a warp and from a small shared-memory hand-off ring between warps, so the
//                     warps run skewed (systolic) and never meet at a block barrier.  Only the
//                     previous row is kept (registers).  The decision bit of every cell
//                     (value[y-1][x] < value[y-1][x-1], core.pyx:32) is the sign of one fp32
//                     subtraction, funnel-shifted into a per-column register and written to HBM
//                     as one word per (32 frames, column): 1 bit per cell.
//   K2 mas_backtrack  one CTA per utterance.  The backtrack (core.pyx:30-33) is a composition of
//                     per-row maps index -> index - dec; it is evaluated group-parallel: for every
//                     group of 32 frames and every entry column the exit column is tabulated
//                     (all in parallel), a short serial chain over groups picks the real entries,
//                     and the groups are re-walked in parallel to emit the per-frame text index.
//   K3 mas_writeout   fills the dense [B,T_y,T_x] path.  Its zero-fill phase does not depend on
//                     K1/K2 and, thanks to PDL, overlaps them on idle SMs; after the dependency
//                     wait it drops the T_y ones per utterance.
//
// Bit-exactness notes (vs core.pyx): the accumulation is the same single-rounded fp32 add per
// cell (no FMA, no reassociation); max is FMNMX (identical to `(a > b) ? a : b` for non-NaN
// inputs; signed zeros cannot occur in-band); the -1e9 sentinel is applied on the diagonal x==y
// and at x==0 exactly as core.pyx:17-27; ties stay (strict <); index==y forces a step.
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"
#include "ptx_sm100.cuh"

namespace mas {

constexpr float kNeg = -1e9f;  // core.pyx:7 max_neg_val

// ------------------------------------------------------------------------------------------------
// lengths from the mask, as monotonic_align/__init__.py:17-18: t_y = sum_y mask[b,y,0],
// t_x = sum_x mask[b,0,x]; float sums are truncated like numpy's astype(int32).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double mask_at(const void* p, int dtype, int64_t off) {
  switch (dtype) {
    case MAS_F32: return static_cast<const float*>(p)[off];
    case MAS_F16: return __half2float(static_cast<const __half*>(p)[off]);
    case MAS_BF16: return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[off]);
    case MAS_F64: return static_cast<const double*>(p)[off];
    case MAS_U8: return static_cast<const uint8_t*>(p)[off];
    case MAS_I8: return static_cast<const int8_t*>(p)[off];
    case MAS_I16: return static_cast<const int16_t*>(p)[off];
    case MAS_I32: return static_cast<const int32_t*>(p)[off];
    default: return static_cast<double>(static_cast<const int64_t*>(p)[off]);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// K1: forward DP
// ------------------------------------------------------------------------------------------------
struct FwdParams {
  const float* nc;
  const int32_t* t_ys;
  const int32_t* t_xs;
  const void* mask;
  int mask_dtype;
  int64_t msb, msy, msx;
  int32_t* lens;    // [B][2] = (t_y, t_x), (0,0) when invalid
  int32_t* status;  // sticky MAS_STATUS_* bits
  uint32_t* bits;   // [B][G][TXP]
  int B, T_y, T_x;
  int R;            // frames per ring stage (multiple of 8)
  int S;            // ring stages
  int W;            // DP warps covering the padded T_x
  int TXP;          // W*32*K
  int G;            // ceil(T_y/32)
  int BR;           // hand-off ring length in frames (power of two >= (S+1)*R)
  uint32_t slot_bytes;
};

template <int K, bool HEAD>
__device__ __forceinline__ void row_step(float (&v)[K], uint32_t (&acc)[K], const float (&c)[K], float edge,
                                         int y, int x0, int lane) {
  float left = __shfl_up_sync(0xffffffffu, v[K - 1], 1);
  if (lane == 0) left = edge;
  if (HEAD) {
    // diagonal: the "stay" candidate value[y-1][y] is the sentinel (core.pyx:17-18).  That cell is
    // outside the band, so overwriting the register copy is harmless.
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (x0 + j == y) v[j] = kNeg;
  }
#pragma unroll
  for (int j = K - 1; j >= 1; --j) {
    const float d = v[j] - v[j - 1];                          // sign bit == (stay < step)
    acc[j] = __funnelshift_l(__float_as_uint(d), acc[j], 1);  // acc = (acc << 1) | sign
    v[j] = c[j] + fmaxf(v[j - 1], v[j]);                      // core.pyx:28
  }
  const float d = v[0] - left;
  acc[0] = __funnelshift_l(__float_as_uint(d), acc[0], 1);
  v[0] = c[0] + fmaxf(left, v[0]);
  if (HEAD) {
    // index == y forces a step in the backtrack (core.pyx:32): fold it into the stored bit.
#pragma unroll
    for (int j = 0; j < K; ++j)
      if (x0 + j == y) acc[j] |= 1u;
  }
}

template <int K, bool VEC>
__device__ __forceinline__ void load_row(float (&c)[K], const float* __restrict__ row, int xl, int T_x) {
  if (VEC) {
    if (K == 1) {
      c[0] = row[xl];
    } else if (K == 2) {
      const float2 t = *reinterpret_cast<const float2*>(row + xl);
      c[0] = t.x;
      c[1] = t.y;
    } else {
#pragma unroll
      for (int q = 0; q < K / 4; ++q) {
        // T_x % 4 == 0 on this path: a quad is entirely inside or entirely outside the row
        const float4 t = *reinterpret_cast<const float4*>(row + min(xl + 4 * q, T_x - 4));
        c[4 * q + 0] = t.x;
        c[4 * q + 1] = t.y;
        c[4 * q + 2] = t.z;
        c[4 * q + 3] = t.w;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < K; ++j) c[j] = row[min(xl + j, T_x - 1)];
  }
}

// BIG: more than 7 DP warps (block of up to 1024 threads, 64 registers each); otherwise the
// block has at most 256 threads and the compiler may use the full register file per thread.
template <int K, bool VEC, bool BIG>
__global__ void __launch_bounds__(BIG ? 1024 : 256, 1) mas_forward_kernel(const FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int S = p.S, R = p.R, W = p.W, BR = p.BR;

  // smem carve-up
  float* ring = reinterpret_cast<float*>(smem);
  float* bnd = reinterpret_cast<float*>(smem + static_cast<size_t>(S) * p.slot_bytes);  // [(W-1)][BR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(bnd + static_cast<size_t>(max(W - 1, 1)) * BR);
  uint64_t* full = bars;             // [S]
  uint64_t* empty = bars + S;        // [S]
  uint64_t* bfull = bars + 2 * S;    // [(W-1)][S]
  double* red = reinterpret_cast<double*>(bfull + static_cast<size_t>(max(W - 1, 1)) * S);  // [2][32]
  int* lens_s = reinterpret_cast<int*>(red + 64);                                            // [2]

  // Let the dependent kernels (backtrack, write-out) get scheduled right away: the write-out's
  // zero-fill does not depend on us.
  ptx::pdl_launch_dependents();

  // ---- lengths -------------------------------------------------------------------------------
  if (p.t_ys != nullptr) {
    if (tid == 0) {
      lens_s[0] = p.t_ys[b];
      lens_s[1] = p.t_xs[b];
    }
  } else {
    double sy = 0.0, sx = 0.0;
    const int64_t base = static_cast<int64_t>(b) * p.msb;
    for (int y = tid; y < p.T_y; y += blockDim.x) sy += mask_at(p.mask, p.mask_dtype, base + y * p.msy);
    for (int x = tid; x < p.T_x; x += blockDim.x) sx += mask_at(p.mask, p.mask_dtype, base + x * p.msx);
    sy = warp_sum(sy);
    sx = warp_sum(sx);
    if (lane == 0) {
      red[warp] = sy;
      red[32 + warp] = sx;
    }
    __syncthreads();
    if (warp == 0) {
      const int nw = blockDim.x >> 5;
      sy = warp_sum(lane < nw ? red[lane] : 0.0);
      sx = warp_sum(lane < nw ? red[32 + lane] : 0.0);
      if (lane == 0) {
        lens_s[0] = static_cast<int>(sy);
        lens_s[1] = static_cast<int>(sx);
      }
    }
  }
  __syncthreads();
  int t_y = lens_s[0], t_x = lens_s[1];
  {
    int st = 0;
    if (t_y < 1 || t_x < 1) st |= MAS_STATUS_EMPTY;
    if (t_y > p.T_y || t_x > p.T_x) st |= MAS_STATUS_TOO_LONG;
    if (t_x > t_y) st |= MAS_STATUS_TX_GT_TY;
    if (st) {
      if (tid == 0) {
        atomicOr(p.status, st);
        p.lens[2 * b] = 0;
        p.lens[2 * b + 1] = 0;
      }
      return;  // whole CTA: the path of this utterance stays all-zero
    }
  }
  const int W_act = (t_x + 32 * K - 1) / (32 * K);  // DP warps that own a column < t_x
  if (tid == 0) {
    p.lens[2 * b] = t_y;
    p.lens[2 * b + 1] = t_x;
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], W_act);
    }
    for (int i = 0; i < (W - 1) * S; ++i) ptx::mbar_init(&bfull[i], 1);
    ptx::mbar_fence_init();
  }
  // hand-off slot of frame 0: the "step" candidate of frame 0 at a warp's first column is the
  // virtual value[-1][x-1] = sentinel.
  if (tid < W - 1) bnd[static_cast<size_t>(tid) * BR] = kNeg;
  __syncthreads();

  const float* nc_b = p.nc + static_cast<size_t>(b) * p.T_y * p.T_x;
  const uint32_t lead_bytes = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(nc_b) & 15u);
  const int nchunks = (t_y + R - 1) / R;
  const size_t slot_floats = p.slot_bytes / 4;

  if (warp == 0) {
    // ---- producer: stream the utterance's frames into the ring ------------------------------
    if (lane == 0) {
      const unsigned char* src0 = reinterpret_cast<const unsigned char*>(nc_b) - lead_bytes;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % S;
        if (c >= S) ptx::mbar_wait(&empty[s], ((c / S) - 1) & 1);
        const int rows = min(R, t_y - c * R);
        const uint32_t bytes = (lead_bytes + static_cast<uint32_t>(rows) * p.T_x * 4u + 15u) & ~15u;
        ptx::mbar_arrive_expect_tx(&full[s], bytes);
        ptx::bulk_g2s(reinterpret_cast<unsigned char*>(ring) + static_cast<size_t>(s) * p.slot_bytes,
                      src0 + static_cast<size_t>(c) * R * p.T_x * 4u, bytes, &full[s]);
      }
    }
    return;
  }

  const int dw = warp - 1;
  if (dw >= W_act) return;

  // ---- DP warp: columns [x0, x0+K) per lane ----------------------------------------------------
  const int x0 = (dw * 32 + lane) * K;
  const int xl = (VEC && K < 4) ? min(x0, p.T_x - K) : x0;  // load column (padding lanes are clamped)
  const bool has_left = dw > 0;
  const bool has_right = dw < W_act - 1;
  const float* bnd_in = bnd + static_cast<size_t>(has_left ? dw - 1 : 0) * BR;
  float* bnd_out = bnd + static_cast<size_t>(has_right ? dw : 0) * BR;
  const bool st_lane = has_right && lane == 31;
  uint32_t* bits_b = p.bits + static_cast<size_t>(b) * p.G * p.TXP + x0;

  float v[K];
  uint32_t acc[K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    v[j] = kNeg;
    acc[j] = 0u;
  }

  auto flush_bits = [&](int g, int nrows) {
    uint32_t* dst = bits_b + static_cast<size_t>(g) * p.TXP;
    const int sh = 32 - nrows;
    if (x0 == 0) acc[0] = 0u;  // the backtrack never leaves column 0 (core.pyx:32 `index != 0`)
    if (K % 4 == 0) {
#pragma unroll
      for (int q = 0; q < K / 4; ++q)
        *reinterpret_cast<uint4*>(dst + 4 * q) =
            make_uint4(acc[4 * q] << sh, acc[4 * q + 1] << sh, acc[4 * q + 2] << sh, acc[4 * q + 3] << sh);
    } else if (K == 2) {
      *reinterpret_cast<uint2*>(dst) = make_uint2(acc[0] << sh, acc[1] << sh);
    } else {
#pragma unroll
      for (int j = 0; j < K; ++j) dst[j] = acc[j] << sh;
    }
  };

  int y = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int s = c % S;
    const uint32_t par = (c / S) & 1;
    ptx::mbar_wait(&full[s], par);
    if (has_left) ptx::mbar_wait(&bfull[(dw - 1) * S + s], par);
    const float* slot = ring + static_cast<size_t>(s) * slot_floats + (lead_bytes >> 2);
    const int rows = min(R, t_y - c * R);
    int r = 0;
    for (; r + 8 <= rows; r += 8, y += 8) {
      float cc[8][K];
      float e[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) load_row<K, VEC>(cc[i], slot + static_cast<size_t>(r + i) * p.T_x, xl, p.T_x);
      const int sl = y & (BR - 1);
      if (has_left) {
        const float4 e0 = *reinterpret_cast<const float4*>(bnd_in + sl);
        const float4 e1 = *reinterpret_cast<const float4*>(bnd_in + sl + 4);
        e[0] = e0.x; e[1] = e0.y; e[2] = e0.z; e[3] = e0.w;
        e[4] = e1.x; e[5] = e1.y; e[6] = e1.z; e[7] = e1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = kNeg;
        if (y == 0) e[0] = 0.0f;  // core.pyx:22-23: the step candidate at (0,0) is 0
      }
      if (y < t_x) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          row_step<K, true>(v, acc, cc[i], e[i], y + i, x0, lane);
          if (st_lane) bnd_out[(sl + i + 1) & (BR - 1)] = v[K - 1];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          row_step<K, false>(v, acc, cc[i], e[i], y + i, x0, lane);
          if (st_lane) bnd_out[(sl + i + 1) & (BR - 1)] = v[K - 1];
        }
      }
      if (((y + 8) & 31) == 0) flush_bits(y >> 5, 32);
    }
    for (; r < rows; ++r, ++y) {  // < 8 leftover frames of the last chunk
      float c1[K];
      load_row<K, VEC>(c1, slot + static_cast<size_t>(r) * p.T_x, xl, p.T_x);
      const int sl = y & (BR - 1);
      const float e1 = has_left ? bnd_in[sl] : (y == 0 ? 0.0f : kNeg);
      row_step<K, true>(v, acc, c1, e1, y, x0, lane);
      if (st_lane) bnd_out[(sl + 1) & (BR - 1)] = v[K - 1];
      if (((y + 1) & 31) == 0) flush_bits(y >> 5, 32);
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(&empty[s]);
    if (st_lane) ptx::mbar_arrive(&bfull[dw * S + s]);
  }
  if (t_y & 31) flush_bits(t_y >> 5, t_y & 31);
}

// ------------------------------------------------------------------------------------------------
// K2: backtrack (core.pyx:30-33) from the decision bits
// ------------------------------------------------------------------------------------------------
struct BtParams {
  const uint32_t* bits;  // [B][G][TXP]
  const int32_t* lens;   // [B][2]
  int32_t* index;        // [B][T_y]
  int T_y, TXP, G;
  int GS;   // groups per shared-memory segment
  int TXS;  // shared-memory row stride (words) of a segment
};

// One backtrack step (core.pyx:32-33) given the decision word of the current column.  The
// forward kernel already folded `index == y` (bit forced to 1) and `index != 0` (column 0
// forced to 0) into the stored bits.
__device__ __forceinline__ int bt_step(int cur, int r, uint32_t word) {
  return cur - static_cast<int>((word >> (31 - r)) & 1u);
}

__global__ void __launch_bounds__(1024, 1) mas_backtrack_kernel(const BtParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;

  uint32_t* sb = reinterpret_cast<uint32_t*>(smem);                        // [GS][TXS]
  int* sentry = reinterpret_cast<int*>(sb + static_cast<size_t>(p.GS) * p.TXS);  // [GS+1]
  int* sidx = sentry + p.GS + 1;                                           // [GS*32]
  uint16_t* sexit = reinterpret_cast<uint16_t*>(sidx + p.GS * 32);         // [GS][TXS]

  ptx::pdl_launch_dependents();
  ptx::pdl_wait();  // forward kernel finished, its bits and lengths are visible

  const int t_y = p.lens[2 * b], t_x = p.lens[2 * b + 1];
  int32_t* idx_b = p.index + static_cast<size_t>(b) * p.T_y;
  for (int y = max(t_y, 0) + tid; y < p.T_y; y += nthr) idx_b[y] = -1;
  if (t_y <= 0) return;

  const uint32_t* bits_b = p.bits + static_cast<size_t>(b) * p.G * p.TXP;
  const int g_top = (t_y - 1) >> 5;
  const int r_top_last = (t_y - 1) & 31;
  if (tid == 0) sentry[p.GS] = t_x - 1;  // entry column of the topmost group (core.pyx:13)
  __syncthreads();

  for (int seg_hi = g_top; seg_hi >= 0; seg_hi -= p.GS) {
    const int seg_lo = max(0, seg_hi - p.GS + 1);
    const int ng = seg_hi - seg_lo + 1;
    // stage the segment's decision words
    for (int i = tid; i < ng * t_x; i += nthr) {
      const int gi = i / t_x, x = i - gi * t_x;
      sb[gi * p.TXS + x] = bits_b[static_cast<size_t>(seg_lo + gi) * p.TXP + x];
    }
    __syncthreads();
    // phase 1: exit column of every (group, entry column)
    for (int i = tid; i < ng * t_x; i += nthr) {
      const int gi = i / t_x, e = i - gi * t_x;
      const int g = seg_lo + gi;
      const uint32_t* row = sb + gi * p.TXS;
      int cur = e;
      const int rt = (g == g_top) ? r_top_last : 31;
      for (int r = rt; r >= 0; --r) cur = bt_step(cur, r, row[cur]);
      sexit[gi * p.TXS + e] = static_cast<uint16_t>(cur);
    }
    __syncthreads();
    // phase 2: serial chain over the segment's groups, top to bottom
    if (tid == 0) {
      int entry = sentry[p.GS];
      for (int gi = ng - 1; gi >= 0; --gi) {
        sentry[gi] = entry;
        entry = sexit[gi * p.TXS + entry];
      }
      sentry[p.GS] = entry;  // entry of the next (lower) segment
    }
    __syncthreads();
    // phase 3: re-walk every group from its real entry, emit the per-frame index
    for (int gi = tid; gi < ng; gi += nthr) {
      const int g = seg_lo + gi;
      const uint32_t* row = sb + gi * p.TXS;
      int cur = sentry[gi];
      const int rt = (g == g_top) ? r_top_last : 31;
      for (int r = rt; r >= 0; --r) {
        sidx[(gi << 5) + r] = cur;
        cur = bt_step(cur, r, row[cur]);
      }
    }
    __syncthreads();
    const int y_lo = seg_lo << 5;
    const int y_hi = min(t_y, (seg_hi + 1) << 5);
    for (int y = y_lo + tid; y < y_hi; y += nthr) idx_b[y] = sidx[y - y_lo];
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// K3: dense path write-out
// ------------------------------------------------------------------------------------------------
struct WoParams {
  unsigned char* out;    // [B*T_y][T_x] elements of es bytes
  const int32_t* index;  // [B*T_y]
  long long rows;        // B*T_y
  int T_x;
  int es;                // element size in bytes
  unsigned long long one;  // bit pattern of 1 in the element type
};

__device__ __forceinline__ void store_elem(unsigned char* p, int es, unsigned long long bits) {
  switch (es) {
    case 1: *p = static_cast<unsigned char>(bits); break;
    case 2: *reinterpret_cast<uint16_t*>(p) = static_cast<uint16_t>(bits); break;
    case 4: *reinterpret_cast<uint32_t*>(p) = static_cast<uint32_t>(bits); break;
    default: *reinterpret_cast<unsigned long long*>(p) = bits; break;
  }
}

__global__ void __launch_bounds__(256) mas_writeout_kernel(const WoParams p) {
  const int tid = threadIdx.x;
  const long long per = (p.rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = min(p.rows, per * blockIdx.x);
  const long long r1 = min(p.rows, r0 + per);
  if (r0 < r1) {
    // phase A: zero-fill this CTA's rows (independent of the forward/backtrack kernels)
    unsigned char* beg = p.out + static_cast<size_t>(r0) * p.T_x * p.es;
    unsigned char* end = p.out + static_cast<size_t>(r1) * p.T_x * p.es;
    unsigned char* abeg = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(beg) + 15u) & ~uintptr_t(15));
    if (abeg > end) abeg = end;
    unsigned char* aend = abeg + ((end - abeg) & ~ptrdiff_t(15));
    for (unsigned char* q = beg + static_cast<size_t>(tid) * p.es; q < abeg; q += static_cast<size_t>(blockDim.x) * p.es)
      store_elem(q, p.es, 0ull);
    for (unsigned char* q = aend + static_cast<size_t>(tid) * p.es; q < end; q += static_cast<size_t>(blockDim.x) * p.es)
      store_elem(q, p.es, 0ull);
    uint4* a4 = reinterpret_cast<uint4*>(abeg);
    const size_t n4 = static_cast<size_t>(aend - abeg) >> 4;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    size_t i = tid;
    for (; i + 3 * blockDim.x < n4; i += 4 * blockDim.x) {
      a4[i] = z;
      a4[i + blockDim.x] = z;
      a4[i + 2 * blockDim.x] = z;
      a4[i + 3 * blockDim.x] = z;
    }
    for (; i < n4; i += blockDim.x) a4[i] = z;
  }
  // phase B: the ones.  Same CTA wrote the zeros of these rows; the barrier orders them.
  ptx::pdl_wait();
  __syncthreads();
  for (long long r = r0 + tid; r < r1; r += blockDim.x) {
    const int x = p.index[r];
    if (x >= 0) store_elem(p.out + (static_cast<size_t>(r) * p.T_x + x) * p.es, p.es, p.one);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_tune_K = 0, g_tune_R = 0, g_tune_S = 0, g_tune_pdl = 1;

struct Layout {
  int G, TXP_max;
  size_t off_status, off_lens, off_index, off_bits, total;
};

// Scratch layout is independent of the tuning (bits rows are padded for the widest K).
static Layout scratch_layout(int B, int T_y, int T_x) {
  Layout L;
  L.G = (T_y + 31) / 32;
  L.TXP_max = ((T_x + 255) / 256) * 256;  // multiple of 32*K for every K <= 8
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  L.off_status = 0;
  L.off_lens = 256;
  L.off_index = up(L.off_lens + static_cast<size_t>(B) * 2 * 4);
  L.off_bits = up(L.off_index + static_cast<size_t>(B) * T_y * 4);
  L.total = up(L.off_bits + static_cast<size_t>(B) * L.G * L.TXP_max * 4);
  return L;
}

struct FwdConfig {
  int K, W, R, S, BR;
  uint32_t slot_bytes;
  size_t smem;
};

static size_t fwd_smem_bytes(int W, int R, int S, int BR, uint32_t slot_bytes) {
  const int wb = W - 1 > 1 ? W - 1 : 1;
  return static_cast<size_t>(S) * slot_bytes + static_cast<size_t>(wb) * BR * 4 + (2 * S + static_cast<size_t>(wb) * S) * 8 +
         64 * 8 + 16;
}

static bool pick_fwd_config(int T_x, FwdConfig* cfg) {
  const size_t budget = 200 * 1024;
  int K = g_tune_K;
  // keep the block at <= 7 DP warps (<= 256 threads) so the DP warps get a full register budget
  if (K == 0) K = T_x <= 32 ? 1 : (T_x <= 448 ? 2 : (T_x <= 896 ? 4 : 8));
  int W = (T_x + 32 * K - 1) / (32 * K);
  while (W > 31 && K < 8) {
    K *= 2;
    W = (T_x + 32 * K - 1) / (32 * K);
  }
  if (W > 31) return false;
  for (int R = g_tune_R ? g_tune_R : 32; R >= 8; R >>= 1) {
    const uint32_t slot = (static_cast<uint32_t>(R) * T_x * 4u + 16u + 15u) & ~15u;
    int S = g_tune_S ? g_tune_S : 8;
    for (; S >= 2; --S) {
      int BR = 8;
      while (BR < (S + 1) * R) BR <<= 1;
      const size_t smem = fwd_smem_bytes(W, R, S, BR, slot);
      if (smem <= budget) {
        if (S < 3 && R > 8 && !g_tune_R) break;  // prefer more, smaller stages
        *cfg = FwdConfig{K, W, R, S, BR, slot, smem};
        return true;
      }
    }
    if (g_tune_R) break;
  }
  return false;
}

template <int K, bool VEC, bool BIG>
static cudaError_t launch_fwd_t(const FwdParams& p, size_t smem, cudaStream_t st) {
  auto kern = mas_forward_kernel<K, VEC, BIG>;
  static size_t smem_set = 0;  // per instantiation; raised once, never during a later stream capture
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    smem_set = 200 * 1024;
  }
  kern<<<p.B, 32 * (p.W + 1), smem, st>>>(p);
  return cudaGetLastError();
}

template <int K, bool VEC>
static cudaError_t launch_fwd(const FwdParams& p, size_t smem, cudaStream_t st) {
  return p.W > 7 ? launch_fwd_t<K, VEC, true>(p, smem, st) : launch_fwd_t<K, VEC, false>(p, smem, st);
}

static cudaError_t launch_fwd_dispatch(int K, bool vec, const FwdParams& p, size_t smem, cudaStream_t st) {
  switch (K) {
    case 1: return launch_fwd<1, true>(p, smem, st);  // K==1 loads are scalar either way
    case 2: return vec ? launch_fwd<2, true>(p, smem, st) : launch_fwd<2, false>(p, smem, st);
    case 4: return vec ? launch_fwd<4, true>(p, smem, st) : launch_fwd<4, false>(p, smem, st);
    case 8: return vec ? launch_fwd<8, true>(p, smem, st) : launch_fwd<8, false>(p, smem, st);
    default: return cudaErrorInvalidValue;
  }
}

template <typename Kern, typename Params>
static cudaError_t launch_pdl(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Params& p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_tune_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

static int elem_size(int dtype) {
  switch (dtype) {
    case MAS_F32: case MAS_I32: return 4;
    case MAS_F16: case MAS_BF16: case MAS_I16: return 2;
    case MAS_F64: case MAS_I64: return 8;
    case MAS_U8: case MAS_I8: return 1;
    default: return 0;
  }
}

static unsigned long long one_bits(int dtype) {
  switch (dtype) {
    case MAS_F32: return 0x3F800000ull;
    case MAS_F16: return 0x3C00ull;
    case MAS_BF16: return 0x3F80ull;
    case MAS_F64: return 0x3FF0000000000000ull;
    default: return 1ull;
  }
}

static int g_num_sms = 0;

int maximum_path(const float* neg_cent, const int32_t* t_ys, const int32_t* t_xs, const void* mask, int mask_dtype,
                 int64_t msb, int64_t msy, int64_t msx, void* path_out, int path_dtype, int32_t* index_out,
                 void* scratch, size_t scratch_bytes, int B, int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || T_y <= 0 || T_x <= 0 || T_x > 2048 || T_y > (1 << 20) || T_x > 65535) return MAS_E_BAD_SHAPE;
  if (!neg_cent || !scratch) return MAS_E_NULL;
  if ((t_ys == nullptr) != (t_xs == nullptr)) return MAS_E_NULL;
  if (!t_ys && !mask) return MAS_E_NULL;
  if (!t_ys && elem_size(mask_dtype) == 0) return MAS_E_BAD_DTYPE;
  if (!path_out && !index_out) return MAS_E_NULL;
  const int es = path_out ? elem_size(path_dtype) : 4;
  if (es == 0) return MAS_E_BAD_DTYPE;
  if ((reinterpret_cast<uintptr_t>(neg_cent) & 3u) || (reinterpret_cast<uintptr_t>(scratch) & 15u) ||
      (path_out && (reinterpret_cast<uintptr_t>(path_out) & (es - 1))))
    return MAS_E_ALIGN;
  const Layout L = scratch_layout(B, T_y, T_x);
  if (scratch_bytes < L.total) return MAS_E_SCRATCH;
  FwdConfig fc;
  if (!pick_fwd_config(T_x, &fc)) return MAS_E_UNSUPPORTED;

  unsigned char* sc = static_cast<unsigned char*>(scratch);
  int32_t* status = reinterpret_cast<int32_t*>(sc + L.off_status);
  int32_t* lens = reinterpret_cast<int32_t*>(sc + L.off_lens);
  int32_t* index = index_out ? index_out : reinterpret_cast<int32_t*>(sc + L.off_index);
  uint32_t* bits = reinterpret_cast<uint32_t*>(sc + L.off_bits);

  // K1
  FwdParams fp{};
  fp.nc = neg_cent; fp.t_ys = t_ys; fp.t_xs = t_xs;
  fp.mask = mask; fp.mask_dtype = mask_dtype; fp.msb = msb; fp.msy = msy; fp.msx = msx;
  fp.lens = lens; fp.status = status; fp.bits = bits;
  fp.B = B; fp.T_y = T_y; fp.T_x = T_x;
  fp.R = fc.R; fp.S = fc.S; fp.W = fc.W; fp.TXP = fc.W * 32 * fc.K; fp.G = L.G; fp.BR = fc.BR;
  fp.slot_bytes = fc.slot_bytes;
  const bool vec = (reinterpret_cast<uintptr_t>(neg_cent) & 15u) == 0 && (T_x % 4) == 0 && T_x >= fc.K;
  cudaError_t e = launch_fwd_dispatch(fc.K, vec, fp, fc.smem, st);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();

  // K2
  BtParams bp{};
  bp.bits = bits; bp.lens = lens; bp.index = index;
  bp.T_y = T_y; bp.TXP = fp.TXP; bp.G = L.G;
  bp.TXS = T_x | 1;  // odd stride: neighbouring groups hit different banks in phase 3
  {
    const size_t per_group = static_cast<size_t>(bp.TXS) * 6 + 32 * 4 + 4;
    int GS = static_cast<int>((160 * 1024) / per_group);
    if (GS > L.G) GS = L.G;
    if (GS < 1) return MAS_E_UNSUPPORTED;
    bp.GS = GS;
  }
  const size_t bt_smem = static_cast<size_t>(bp.GS) * bp.TXS * 6 + (bp.GS + 1) * 4 + static_cast<size_t>(bp.GS) * 32 * 4 + 64;
  static bool bt_attr = false;
  if (!bt_attr) {
    e = cudaFuncSetAttribute(mas_backtrack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    bt_attr = true;
  }
  {
    long long tasks = static_cast<long long>(bp.GS) * T_x;
    int threads = tasks >= 1024 ? 1024 : static_cast<int>((tasks + 31) / 32 * 32);
    if (threads < 64) threads = 64;
    e = launch_pdl(mas_backtrack_kernel, dim3(B), dim3(threads), bt_smem, st, bp);
    if (e != cudaSuccess) return static_cast<int>(e);
    count_launch();
  }

  // K3
  if (path_out) {
    if (g_num_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
      if (g_num_sms <= 0) g_num_sms = 148;
    }
    WoParams wp{};
    wp.out = static_cast<unsigned char*>(path_out);
    wp.index = index;
    wp.rows = static_cast<long long>(B) * T_y;
    wp.T_x = T_x;
    wp.es = es;
    wp.one = one_bits(path_dtype);
    long long want = (wp.rows * T_x * es + (64 * 1024 - 1)) / (64 * 1024);  // >= 64 KiB per CTA
    int grid = static_cast<int>(want < 1 ? 1 : (want > 8LL * g_num_sms ? 8LL * g_num_sms : want));
    e = launch_pdl(mas_writeout_kernel, dim3(grid), dim3(256), 0, st, wp);
    if (e != cudaSuccess) return static_cast<int>(e);
    count_launch();
  }
  return MAS_OK;
}

size_t maximum_path_scratch_bytes(int B, int T_y, int T_x) {
  if (B <= 0 || T_y <= 0 || T_x <= 0) return 0;
  return scratch_layout(B, T_y, T_x).total;
}

void set_tuning(int K, int R, int S, int pdl) {
  g_tune_K = K;
  g_tune_R = R;
  g_tune_S = S;
  g_tune_pdl = pdl;
}

}  // namespace mas
