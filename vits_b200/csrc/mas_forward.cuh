// mas_forward.cuh -- the forward dynamic-program kernel of the Monotonic Alignment Search
// (template; instantiated per columns-per-lane K in mas_fwd_k*.cu so the instantiations compile
// in parallel).  See mas_path.cu for the overall design.
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <type_traits>

#include "../../include/vits_mas.h"
#include "mas_internal.h"
#include "ptx_sm100.cuh"
#include "mas_common.cuh"

namespace mas {

// ------------------------------------------------------------------------------------------------
// K1: forward DP
// ------------------------------------------------------------------------------------------------
// Shared-memory layout of the forward kernel, computed on the host (byte offsets).
struct FwdSmem {
  uint32_t ring, bnd, bars, red, sbits, sexit, sentry, sidx, total;
};

struct FwdParams {
  const float* nc;
  const int32_t* t_ys;
  const int32_t* t_xs;
  const void* mask;
  int mask_dtype;
  int64_t msb, msy, msx;
  int32_t* lens;    // [B][2] = (t_y, t_x), (0,0) when invalid
  int32_t* status;  // sticky MAS_STATUS_* bits
  int32_t* mirror;  // host-mapped copy (one word per bit) or nullptr
  int32_t* wo_counters;  // the write-out kernel's two work counters; zeroed here before it may start
  uint32_t* bits;   // [B][G][TXP]   (unfused mode only); streaming mode: pairs {word, tag = 1}
  uint2* lenstag;   // streaming mode: [B] {t_y << 12 | t_x, tag = 1}; else nullptr
  int32_t* index;   // [B][T_y]      (fused mode: written by this kernel)
  unsigned long long* tl;  // optional timeline stamps (debug), or nullptr
  unsigned long long* trace;  // optional per-warp event trace of CTA 0 (debug): [8 warps][512][2]
  int B, T_y, T_x;
  int S;            // ring stages
  int W;            // DP warps covering the padded T_x
  int H;            // helper warps (fused mode), 0 otherwise
  int TXP;          // W*32*K: row stride of the decision words
  int G;            // ceil(T_y/32)
  int BR;           // hand-off ring length in frames (power of two >= (S+1)*R)
  int fused;        // 1: decision bits stay in shared memory and this kernel also backtracks
  int pdl;          // launch with the programmatic-serialization attribute
  uint32_t slot_bytes;
  FwdSmem sm;
};

// One frame of the recurrence for the K columns of a lane.
//   DIAG: this frame may hold the diagonal cell x == y in column `jd` of the lane for which `diag`
//   is true.  There the "stay" candidate value[y-1][y] is the sentinel (core.pyx:17-18; that cell is
//   outside the band, so overwriting the register copy is harmless) and the backtrack is forced to
//   step (core.pyx:32 `index == y`), which is folded into the stored decision bit.
template <int K, bool DIAG>
__device__ __forceinline__ void row_step(float (&v)[K], uint32_t (&acc)[K], const float (&c)[K], float edge,
                                         bool lane0, int jd, bool diag) {
  float left = __shfl_up_sync(0xffffffffu, v[K - 1], 1);
  if (lane0) left = edge;
  if (DIAG) v[jd] = diag ? kNeg : v[jd];
#pragma unroll
  for (int j = K - 1; j >= 1; --j) {
    const float d = v[j] - v[j - 1];                          // sign bit == (stay < step)
    acc[j] = __funnelshift_l(__float_as_uint(d), acc[j], 1);  // acc = (acc << 1) | sign
    v[j] = c[j] + fmaxf(v[j - 1], v[j]);                      // core.pyx:28
  }
  const float d = v[0] - left;
  acc[0] = __funnelshift_l(__float_as_uint(d), acc[0], 1);
  v[0] = c[0] + fmaxf(left, v[0]);
  if (DIAG) acc[jd] |= diag ? 1u : 0u;
}

// Same, any column may be the diagonal one (used only for the < R leftover frames).
template <int K>
__device__ __forceinline__ void row_step_any(float (&v)[K], uint32_t (&acc)[K], const float (&c)[K], float edge,
                                             bool lane0, int y, int x0) {
  float left = __shfl_up_sync(0xffffffffu, v[K - 1], 1);
  if (lane0) left = edge;
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (x0 + j == y) v[j] = kNeg;
#pragma unroll
  for (int j = K - 1; j >= 1; --j) {
    const float d = v[j] - v[j - 1];
    acc[j] = __funnelshift_l(__float_as_uint(d), acc[j], 1);
    v[j] = c[j] + fmaxf(v[j - 1], v[j]);
  }
  const float d = v[0] - left;
  acc[0] = __funnelshift_l(__float_as_uint(d), acc[0], 1);
  v[0] = c[0] + fmaxf(left, v[0]);
#pragma unroll
  for (int j = 0; j < K; ++j)
    if (x0 + j == y) acc[j] |= 1u;
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
    } else if (K == 3) {
#pragma unroll
      for (int j = 0; j < K; ++j) c[j] = row[min(xl + j, T_x - 1)];
    } else if (K == 6) {
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        // T_x is even on this path: a pair is entirely inside or entirely outside the row
        const float2 t = *reinterpret_cast<const float2*>(row + min(xl + 2 * q, T_x - 2));
        c[2 * q + 0] = t.x;
        c[2 * q + 1] = t.y;
      }
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

// BIG: more than 7 warps besides the producer (block of up to 1024 threads, 64 registers each);
// otherwise the block has at most 256 threads and the compiler may use the full register file.
// BPC: 8-frame blocks per ring stage (R = 8 * BPC frames); the stage is the unit of the main loop.
//
// Warp roles: warp 0 = producer (bulk async copies into the ring), warps 1..W = DP warps,
// warps W+1..W+H = helpers (fused mode): while the DP warps sweep forward, the helpers tabulate,
// for every finished group of 32 frames, the exit column of every entry column ("phase 1" of
// the backtrack), so that after the last frame only a short chain over groups and one parallel
// re-walk remain.
template <int K, bool VEC, bool BIG, int BPC>
__global__ void __launch_bounds__(BIG ? 1024 : 256, 1) mas_forward_kernel(const FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int S = p.S, W = p.W, BR = p.BR;
  constexpr int R = 8 * BPC;
  const bool fused = p.fused != 0;

  // smem carve-up (offsets from the host)
  float* ring = reinterpret_cast<float*>(smem + p.sm.ring);
  float* bnd = reinterpret_cast<float*>(smem + p.sm.bnd);            // [W+1][BR]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.sm.bars);
  uint64_t* full = bars;                                             // [S]
  uint64_t* empty = bars + S;                                        // [S]
  uint64_t* bfull = bars + 2 * S;                                    // [max(W-1,1)][S]
  uint64_t* gbar = bfull + static_cast<size_t>(max(W - 1, 1)) * S;   // [G]   (fused)
  double* red = reinterpret_cast<double*>(smem + p.sm.red);          // [2][32]
  int* lens_s = reinterpret_cast<int*>(red + 64);                    // [2]
  uint32_t* sbits = reinterpret_cast<uint32_t*>(smem + p.sm.sbits);  // [G][TXP]  (fused)
  uint16_t* sexit = reinterpret_cast<uint16_t*>(smem + p.sm.sexit);  // [G][TXP]  (fused)
  int* sentry = reinterpret_cast<int*>(smem + p.sm.sentry);          // [G]       (fused)
  int16_t* sidx = reinterpret_cast<int16_t*>(smem + p.sm.sidx);      // [G*32]    (fused)

  // Let the dependent kernels (backtrack, write-out) get scheduled right away: the write-out's
  // zero-fill does not depend on us.
  // This kernel may itself have been launched programmatically (behind the previous call's write-out):
  // everything above touched only shared memory; global memory is first used below.
  ptx::pdl_wait();
  // The first ring stages are requested before the lengths are known (they only need T_y as a bound;
  // frames beyond t_y are padding that exists in memory), so the HBM latency of the first frames
  // overlaps the mask reduction.
  const float* nc_b = p.nc + static_cast<size_t>(b) * p.T_y * p.T_x;
  const uint32_t lead_bytes = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(nc_b) & 15u);
  const int nspec = min(S, (p.T_y + R - 1) / R);
  if (tid == 0) {
    for (int s = 0; s < S; ++s) ptx::mbar_init(&full[s], 1);
    ptx::mbar_fence_init();
    const unsigned char* src0 = reinterpret_cast<const unsigned char*>(nc_b) - lead_bytes;
    for (int c = 0; c < nspec; ++c) {
      const int rows = min(R, p.T_y - c * R);
      const uint32_t bytes = (lead_bytes + static_cast<uint32_t>(rows) * p.T_x * 4u + 15u) & ~15u;
      ptx::mbar_arrive_expect_tx(&full[c], bytes);
      ptx::bulk_g2s(reinterpret_cast<unsigned char*>(ring) + static_cast<size_t>(c) * p.slot_bytes,
                    src0 + static_cast<size_t>(c) * R * p.T_x * 4u, bytes, &full[c]);
    }
  }

  if (b == 0 && tid == 0) {
    p.wo_counters[0] = 0;
    p.wo_counters[1] = 0;
  }
  const bool ll = p.lenstag != nullptr;
  if (ll) {
    // Streaming mode hands the decision words to the concurrently running backtrack kernel without any
    // fence: every 32-bit word travels in one 8-byte store together with a tag, and a reader accepts an
    // element only when the tag is set -- the flag-in-data scheme of NCCL's LL protocol (a GPU-scope
    // release per group would cost a DP warp ~0.8 us each).  Clear this utterance's tags (whatever the
    // scratch held: its layout depends on the shape) before the backtrack kernel can start.
    uint4* z = reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(p.bits) + static_cast<size_t>(b) * p.G * p.TXP);
    const int n16 = p.G * p.TXP / 2;  // TXP is a multiple of 32
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) *reinterpret_cast<unsigned long long*>(p.lenstag + b) = 0ull;
  }
  if (ll) __syncthreads();
  if (tid == 0 && (ll || b == 0)) __threadfence();  // (cumulative: covers the other threads' stores ordered by the barrier)
  __syncthreads();  // counters and tags are cleared before any thread of this CTA lets the dependent kernels start
  ptx::pdl_launch_dependents();
  if (tid == 0) tl_min(p.tl, 0);

  // ---- lengths -------------------------------------------------------------------------------
  if (p.t_ys != nullptr) {
    if (tid == 0) {
      lens_s[0] = p.t_ys[b];
      lens_s[1] = p.t_xs[b];
    }
  } else {
    double sy = 0.0, sx = 0.0;
    const int64_t base = static_cast<int64_t>(b) * p.msb;
    mask_sums(p.mask, p.mask_dtype, base, p.msy, p.T_y, p.msx, p.T_x, tid, blockDim.x, sy, sx);
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
  const int t_y = lens_s[0], t_x = lens_s[1];
  {
    int st = 0;
    if (t_y < 1 || t_x < 1) st |= MAS_STATUS_EMPTY;
    if (t_y > p.T_y || t_x > p.T_x) st |= MAS_STATUS_TOO_LONG;
    if (t_x > t_y) st |= MAS_STATUS_TX_GT_TY;
    if (st) {  // whole CTA: the path of this utterance stays all-zero
      if (tid == 0) {
        raise_status(p.status, p.mirror, st);
        p.lens[2 * b] = 0;
        p.lens[2 * b + 1] = 0;
        if (ll) *reinterpret_cast<unsigned long long*>(p.lenstag + b) = pack_tagged(0u, 1u);
        for (int c = 0; c < nspec; ++c) ptx::mbar_wait(&full[c], 0);  // no copy may outlive the CTA
      }
      if (fused)
        for (int y = tid; y < p.T_y; y += blockDim.x) p.index[static_cast<size_t>(b) * p.T_y + y] = -1;
      return;
    }
  }
  const int W_act = (t_x + 32 * K - 1) / (32 * K);  // DP warps that own a column < t_x
  const int g_top = (t_y - 1) >> 5;                  // group of the last frame
  if (tid == 0) {
    p.lens[2 * b] = t_y;
    p.lens[2 * b + 1] = t_x;
    if (ll) *reinterpret_cast<unsigned long long*>(p.lenstag + b) = pack_tagged((static_cast<uint32_t>(t_y) << 12) | static_cast<uint32_t>(t_x), 1u);
    for (int s = 0; s < S; ++s) ptx::mbar_init(&empty[s], W_act);
    for (int i = 0; i < (W - 1) * S; ++i) ptx::mbar_init(&bfull[i], 1);
    if (fused)
      for (int g = 0; g <= g_top; ++g) ptx::mbar_init(&gbar[g], W_act);
    ptx::mbar_fence_init();
  }
  // bnd[0][*]: the "step" candidate at x == 0 is the sentinel, except 0 at frame 0 (core.pyx:21-25).
  // bnd[i][0], i >= 1: frame 0's step candidate at a warp's first column is the virtual
  // value[-1][x-1] = sentinel.
  for (int i = tid; i < BR; i += blockDim.x) bnd[i] = i == 0 ? 0.0f : kNeg;
  if (tid >= 1 && tid <= W) bnd[static_cast<size_t>(tid) * BR] = kNeg;
  __syncthreads();

  const size_t slot_floats = p.slot_bytes / 4;
  const int dw = warp - 1;

  if (warp == 0) {
    // ---- producer: stream the utterance's frames into the ring ------------------------------
    if (lane == 0) {
      const int nchunks = (t_y + R - 1) / R;
      const unsigned char* src0 = reinterpret_cast<const unsigned char*>(nc_b) - lead_bytes;
      // stages 0..nspec-1 were requested in the prologue; speculative ones beyond the utterance's
      // frames are never consumed, so drain them here (no copy may outlive the CTA)
      for (int c = nchunks; c < nspec; ++c) ptx::mbar_wait(&full[c], 0);
      int s = nspec % S;
      uint32_t par = (nspec == S) ? 0u : 1u;  // parity of the previous use of the stage
      for (int c = nspec; c < nchunks; ++c) {
        ptx::mbar_wait(&empty[s], par);
        const int rows = min(R, t_y - c * R);
        const uint32_t bytes = (lead_bytes + static_cast<uint32_t>(rows) * p.T_x * 4u + 15u) & ~15u;
        ptx::mbar_arrive_expect_tx(&full[s], bytes);
        ptx::bulk_g2s(reinterpret_cast<unsigned char*>(ring) + static_cast<size_t>(s) * p.slot_bytes,
                      src0 + static_cast<size_t>(c) * R * p.T_x * 4u, bytes, &full[s]);
        if (++s == S) {
          s = 0;
          par ^= 1u;
        }
      }
    }
    if (!fused) return;
  } else if (dw < W_act) {
    // ---- DP warp: columns [x0, x0+K) per lane --------------------------------------------------
    // Hand-off arrays: warp dw reads its left edge from bnd[dw] and publishes its last column to
    // bnd[dw+1].  bnd[0] is constant (the x==0 sentinel), bnd[W_act] is a sink nobody reads, so the
    // inner loop is identical (and branch-free) for every warp.
    const int x0 = (dw * 32 + lane) * K;
    const int xl = (VEC && K < 3) ? min(x0, p.T_x - K) : x0;  // load column (padding lanes are clamped)
    const bool has_left = dw > 0;
    const bool has_right = dw < W_act - 1;
    const float* bnd_in = bnd + static_cast<size_t>(dw) * BR;
    float* bnd_out = bnd + static_cast<size_t>(dw + 1) * BR;
    const bool lane0 = lane == 0;
    const bool lane31 = lane == 31;
    // decision words: shared memory (fused) or the HBM scratch; generic stores, once per 32 frames
    uint32_t* bits_b = (fused ? sbits : p.bits + static_cast<size_t>(b) * p.G * p.TXP) + x0;
    const int T_x = p.T_x;
    const int xw0 = dw * 32 * K, xw1 = xw0 + 32 * K;  // the diagonal crosses this warp's columns
                                                      // only in frames [xw0, xw1)
    float v[K];
    uint32_t acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
      v[j] = kNeg;
      acc[j] = 0u;
    }
#ifdef MAS_TRACE  // compile-time only: even a never-taken runtime check costs the DP warps ~3 us per call
    int tr_n = 0;  // debug trace cursor (CTA 0, lane 0 of each DP warp; plain stores, no atomics)
    auto tr = [&](int tag, int idx) {
      if (p.trace && b == 0 && lane0 && tr_n < 512) {
        unsigned long long* q = p.trace + (static_cast<size_t>(dw) * 512 + tr_n) * 2;
        q[0] = (static_cast<unsigned long long>(tag) << 32) | static_cast<unsigned>(idx);
        q[1] = clock64();
        ++tr_n;
      }
    };
#else
    auto tr = [](int, int) {};
#endif

    auto flush_bits = [&](int g, int nrows) {
      uint32_t* dst = bits_b + static_cast<size_t>(g) * p.TXP;
      const int sh = 32 - nrows;
      if (x0 == 0) acc[0] = 0u;  // the backtrack never leaves column 0 (core.pyx:32 `index != 0`)
      if (ll) {
        uint2* d2 = reinterpret_cast<uint2*>(p.bits) + (static_cast<size_t>(b) * p.G + g) * p.TXP + x0;
        if (K % 2 == 0) {
#pragma unroll
          for (int q = 0; q < K / 2; ++q)
            ptx::st_global_v2_u64(d2 + 2 * q, pack_tagged(acc[2 * q] << sh, 1u), pack_tagged(acc[2 * q + 1] << sh, 1u));
        } else {
#pragma unroll
          for (int j = 0; j < K; ++j) *reinterpret_cast<unsigned long long*>(d2 + j) = pack_tagged(acc[j] << sh, 1u);
        }
      } else if (K % 4 == 0) {
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
      if (fused) {  // tell the helpers that this warp's words of group g are in place
        __syncwarp();
        if (lane0) ptx::mbar_arrive(&gbar[g]);
      }
    };
    // load the 8 frames of one block (and their left-edge values) into registers
    auto load_block = [&](const float* rp, const float* ep, float (&cc)[8][K], float (&e)[8]) {
#pragma unroll
      for (int i = 0; i < 8; ++i) load_row<K, VEC>(cc[i], rp + static_cast<size_t>(i) * T_x, xl, T_x);
      const float4 e0 = *reinterpret_cast<const float4*>(ep);
      const float4 e1 = *reinterpret_cast<const float4*>(ep + 4);
      e[0] = e0.x; e[1] = e0.y; e[2] = e0.z; e[3] = e0.w;
      e[4] = e1.x; e[5] = e1.y; e[6] = e1.z; e[7] = e1.w;
    };

    // ---- main loop: one ring stage (R frames) per iteration, straight-line inside ------------
    // The first block of the NEXT stage is prefetched during the last block of the current one when
    // that stage happens to be ready (non-blocking test only: a warp must never block on a stage
    // while it still holds one, that dead-locks once S <= W); its shared-memory latency and the
    // barrier checks then hide behind the recurrence instead of sitting between two stages.
    int stage = 0;
    uint32_t par = 0u;
    const int nfull = t_y / R;  // R is a compile-time power of two
    int y0 = 0;
    constexpr bool XPF = (BPC % 2) == 0;  // buffer 0 is free while the last (odd) block runs
    float cc[2][8][K], e[2][8];
    bool have0 = false;  // block 0 of the current stage is already in cc[0]/e[0]
    for (int c = 0; c < nfull; ++c, y0 += R) {
      const float* rp = ring + static_cast<size_t>(stage) * slot_floats + (lead_bytes >> 2);
      const int sl = y0 & (BR - 1);  // BR is a multiple of R: no wrap inside the stage
      const float* ep = bnd_in + sl;
      tr(1, c);
      if (!have0) {
        // test_wait first: it is ~3x cheaper than try_wait when the phase has already completed
        if (!ptx::mbar_test(&full[stage], par)) ptx::mbar_wait(&full[stage], par);            // frames landed
        tr(2, c);
        if (has_left && !ptx::mbar_test(&bfull[(dw - 1) * S + stage], par))
          ptx::mbar_wait(&bfull[(dw - 1) * S + stage], par);                                  // left neighbour done
        tr(3, c);
        load_block(rp, ep, cc[0], e[0]);
      }
      float* bo = bnd_out + sl;
      float* bo_last = bnd_out + ((y0 + R) & (BR - 1));
      const bool diag_stage = (y0 + R > xw0) && (y0 < xw1);  // warp-uniform
      const int dx = x0 - y0;                                 // column - frame at the stage's first frame
      // position of the next stage, for the cross-stage prefetch
      int nstage = stage + 1;
      uint32_t npar = par;
      if (nstage == S) {
        nstage = 0;
        npar ^= 1u;
      }
      bool next_ready = false;

      // the whole stage is straight-line code; the diagonal variant is chosen once per stage
      auto stage_body = [&](auto diag_tag, auto mod_tag) {
        constexpr bool DIAG = decltype(diag_tag)::value;
        constexpr int Y0MOD = decltype(mod_tag)::value;  // y0 % K (0 whenever K divides the stage length)
#pragma unroll
        for (int kb = 0; kb < BPC; ++kb) {
          if (kb + 1 < BPC) {
            load_block(rp + static_cast<size_t>(kb + 1) * 8 * T_x, ep + (kb + 1) * 8, cc[(kb + 1) & 1],
                       e[(kb + 1) & 1]);
          } else if (XPF && c + 1 < nfull) {
            bool ok = ptx::mbar_test(&full[nstage], npar);
            if (has_left) ok = ok && ptx::mbar_test(&bfull[(dw - 1) * S + nstage], npar);
            next_ready = __all_sync(0xffffffffu, ok);
            if (next_ready)
              load_block(ring + static_cast<size_t>(nstage) * slot_floats + (lead_bytes >> 2),
                         bnd_in + ((y0 + R) & (BR - 1)), cc[0], e[0]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = kb * 8 + i, jd = (Y0MOD + r) % K;  // only column (y0 + r) % K of a lane can be diagonal
            row_step<K, DIAG>(v, acc, cc[kb & 1][i], e[kb & 1][i], lane0, jd, DIAG && (dx + jd == r));
            if (lane31) *(r < R - 1 ? bo + r + 1 : bo_last) = v[K - 1];
          }
        }
      };
      if (!diag_stage) {
        stage_body(std::false_type{}, std::integral_constant<int, 0>{});
      } else if (R % K == 0) {
        stage_body(std::true_type{}, std::integral_constant<int, 0>{});
      } else {  // K in {3, 6}: y0 % K cycles through the multiples of gcd(R, K)
        switch (y0 % K) {
          case 0: stage_body(std::true_type{}, std::integral_constant<int, 0>{}); break;
          case 1: stage_body(std::true_type{}, std::integral_constant<int, 1 % K>{}); break;
          case 2: stage_body(std::true_type{}, std::integral_constant<int, 2 % K>{}); break;
          case 3: stage_body(std::true_type{}, std::integral_constant<int, 3 % K>{}); break;
          case 4: stage_body(std::true_type{}, std::integral_constant<int, 4 % K>{}); break;
          default: stage_body(std::true_type{}, std::integral_constant<int, 5 % K>{}); break;
        }
      }
      have0 = next_ready;
      tr(4, c);

      if (c == 0 && dw == 0 && lane0) bnd[0] = kNeg;  // the (0,0) special case is consumed
      if (((y0 + R) & 31) == 0) flush_bits(y0 >> 5, 32);
      tr(5, c);
      __syncwarp();
      if (lane0) ptx::mbar_arrive(&empty[stage]);                               // stage may be refilled
      if (lane31 && has_right) ptx::mbar_arrive(&bfull[dw * S + stage]);       // our last column is published
      tr(6, c);
      stage = nstage;
      par = npar;
    }
    // ---- < R leftover frames, one at a time ------------------------------------------------------
    if (y0 < t_y) {
      ptx::mbar_wait(&full[stage], par);
      if (has_left) ptx::mbar_wait(&bfull[(dw - 1) * S + stage], par);
      const float* rp = ring + static_cast<size_t>(stage) * slot_floats + (lead_bytes >> 2);
      for (int y = y0; y < t_y; ++y, rp += T_x) {
        float c1[K];
        load_row<K, VEC>(c1, rp, xl, T_x);
        const int sl = y & (BR - 1);
        const float e1 = bnd_in[sl];
        row_step_any<K>(v, acc, c1, e1, lane0, y, x0);
        if (lane31) bnd_out[(sl + 1) & (BR - 1)] = v[K - 1];
        if (dw == 0 && lane0 && y == 0) bnd[0] = kNeg;
        if (((y + 1) & 31) == 0) flush_bits(y >> 5, 32);
      }
      __syncwarp();
      if (lane0) ptx::mbar_arrive(&empty[stage]);
      if (lane31 && has_right) ptx::mbar_arrive(&bfull[dw * S + stage]);
    }
    if (t_y & 31) flush_bits(t_y >> 5, t_y & 31);
    if (lane0) tl_max(p.tl, 1);
    if (!fused) return;
  } else if (fused && dw >= W && dw < W + p.H) {
    // ---- helper warp: exit column of every entry column, group by group (the top group needs no
    // table: its entry is known) ------------------------------------------------------------------
    constexpr int QP = 4;  // independent walks per lane, to hide the shared-memory latency
    for (int g = dw - W; g < g_top; g += p.H) {
      ptx::mbar_wait(&gbar[g], 0);
      const uint32_t* row = sbits + static_cast<size_t>(g) * p.TXP;
      uint16_t* ex = sexit + static_cast<size_t>(g) * p.TXP;
      for (int e0 = lane; e0 < t_x; e0 += 32 * QP) {
        int cur[QP];
#pragma unroll
        for (int q = 0; q < QP; ++q) cur[q] = min(e0 + 32 * q, t_x - 1);
#pragma unroll
        for (int r = 31; r >= 0; --r) {
#pragma unroll
          for (int q = 0; q < QP; ++q) cur[q] = bt_step(cur[q], r, row[cur[q]]);
        }
#pragma unroll
        for (int q = 0; q < QP; ++q)
          if (e0 + 32 * q < t_x) ex[e0 + 32 * q] = static_cast<uint16_t>(cur[q]);
      }
    }
  } else if (!fused) {
    return;
  }

  // ---- fused tail: the backtrack proper (core.pyx:30-33) ------------------------------------------
  __syncthreads();  // all decision words and exit tables are in shared memory
  if (warp == 0) {
    // (1) top group, whose entry (t_y-1, t_x-1) is known.  The walk can visit at most 32 columns; lane l
    // holds the decision word of column t_x-1-l, 32 ballots transpose them into one mask per frame, and
    // the walk itself is then pure register arithmetic (no dependent shared-memory loads).
    const int rt = (t_y - 1) & 31;
    const int col = t_x - 1 - lane;
    const uint32_t w = col >= 0 ? sbits[static_cast<size_t>(g_top) * p.TXP + col] : 0u;
    uint32_t pos = 0;
#pragma unroll
    for (int r = 31; r >= 0; --r) {
      const uint32_t m = __ballot_sync(0xffffffffu, (w >> (31 - r)) & 1u);  // bit l: decision of column t_x-1-l
      if (r <= rt) pos += (m >> pos) & 1u;
    }
    // (2) chain over the groups below: one table look-up each
    if (lane == 0) {
      int cur = t_x - 1;
      sentry[g_top] = cur;
      cur -= static_cast<int>(pos);
      for (int g = g_top - 1; g >= 0; --g) {
        sentry[g] = cur;
        cur = sexit[static_cast<size_t>(g) * p.TXP + cur];
      }
    }
  }
  __syncthreads();
  // re-walk every group from its real entry, emit the per-frame index
  for (int g = tid; g <= g_top; g += blockDim.x) {
    const uint32_t* row = sbits + static_cast<size_t>(g) * p.TXP;
    int cur = sentry[g];
    for (int r = (g == g_top) ? ((t_y - 1) & 31) : 31; r >= 0; --r) {
      sidx[(g << 5) + r] = static_cast<int16_t>(cur);
      cur = bt_step(cur, r, row[cur]);
    }
  }
  __syncthreads();
  int32_t* idx_b = p.index + static_cast<size_t>(b) * p.T_y;
  for (int y = tid; y < p.T_y; y += blockDim.x) idx_b[y] = y < t_y ? static_cast<int>(sidx[y]) : -1;
  if (tid == 0) tl_max(p.tl, 2);
}

// ------------------------------------------------------------------------------------------------
// launch helpers (one static "attribute set" flag per instantiation)
// ------------------------------------------------------------------------------------------------
template <int K, bool VEC, bool BIG, int BPC>
inline cudaError_t launch_fwd_t(const FwdParams& p, cudaStream_t st) {
  auto kern = mas_forward_kernel<K, VEC, BIG, BPC>;
  static std::atomic<uint64_t> attr_set{0};  // per instantiation and device
  if (cudaError_t e = ensure_dyn_smem(kern, 200 * 1024, attr_set); e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.B);
  cfg.blockDim = dim3(32 * (1 + p.W + p.H));
  cfg.dynamicSmemBytes = p.sm.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <int K, bool VEC, bool BIG>
inline cudaError_t launch_fwd_r(const FwdParams& p, int R, cudaStream_t st) {
  switch (R) {
    case 8: return launch_fwd_t<K, VEC, BIG, 1>(p, st);
    case 16: return launch_fwd_t<K, VEC, BIG, 2>(p, st);
    case 32: return launch_fwd_t<K, VEC, BIG, 4>(p, st);
    default: return cudaErrorInvalidValue;
  }
}

template <int K, bool VEC>
inline cudaError_t launch_fwd(const FwdParams& p, int R, cudaStream_t st) {
  return (p.W + p.H) > 7 ? launch_fwd_r<K, VEC, true>(p, R, st) : launch_fwd_r<K, VEC, false>(p, R, st);
}


// one entry per K, defined in mas_fwd_k{1,2,4,8}.cu
cudaError_t launch_fwd_k1(bool vec, const FwdParams& p, int R, cudaStream_t st);
cudaError_t launch_fwd_k2(bool vec, const FwdParams& p, int R, cudaStream_t st);
cudaError_t launch_fwd_k3(bool vec, const FwdParams& p, int R, cudaStream_t st);
cudaError_t launch_fwd_k4(bool vec, const FwdParams& p, int R, cudaStream_t st);
cudaError_t launch_fwd_k6(bool vec, const FwdParams& p, int R, cudaStream_t st);
cudaError_t launch_fwd_k8(bool vec, const FwdParams& p, int R, cudaStream_t st);

}  // namespace mas
