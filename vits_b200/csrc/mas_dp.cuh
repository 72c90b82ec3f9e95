// mas_dp.cuh -- the forward dynamic-program kernel of the Monotonic Alignment Search
// (monotonic_align/core.pyx:13-28 of the reference), template instantiated per columns-per-lane K in
// mas_dp_k*.cu.  See mas_path.cu for the overall design.
//
// Wavefront inside a warp: lane l owns K adjacent text columns and, at step t, computes frame
// y = t - D*l.  The x-1 neighbour of frame y lives one lane to the left and was produced D+1 steps
// earlier, so its __shfl_up is issued D steps ahead of its use: a warp issues in order, and with
// D = 3 (~50 cycles of other work) the ~50-cycle SHFL latency is no longer on the per-frame chain
// (FMNMX -> FADD).  Every warp streams only its own 32*K columns:
// 2-D tiled TMA boxes [R frames x 32K columns] land in a private shared-memory ring (the lane that
// frees a slot re-arms it: no producer warp, no "empty" barriers), so the lag between neighbouring
// warps costs no shared memory.  Warps hand their last column to the right neighbour through a small
// shared-memory ring published with per-superstep progress counters.  The unit of work is a
// "superstep" of 32 steps: straight-line code, every shared-memory address is a per-lane base plus a
// compile-time offset, all waiting happens between supersteps.
#pragma once
#include <cstdint>
#include <cuda.h>
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
struct DpSmem {
  uint32_t ring, bnd, bars, prog, red, total;  // byte offsets, computed on the host
};

struct DpParams {
  const float* nc;
  const int32_t* t_ys;
  const int32_t* t_xs;
  const void* mask;
  int mask_dtype;
  int64_t msb, msy, msx;
  int32_t* lens;         // [B][2] = (t_y, t_x), (0,0) when invalid
  int32_t* status;       // sticky MAS_STATUS_* bits
  int32_t* mirror;       // host-mapped copy (one word per bit) or nullptr
  int32_t* wo_counters;  // the fill kernel's two work counters; zeroed here before it may start
  uint32_t* bits;        // [B][G][TXP] decision words: bit (31-r) of word [g][x] = "step left when leaving frame 32g+r";
                         // streaming mode: [B][G][TXP] pairs {word, tag}
  uint2* lenstag;        // streaming mode: [B] {t_y << 12 | t_x, tag = 1}; else nullptr
  unsigned long long* tl;
  unsigned long long* trace;  // optional debug trace of CTA 0: [8 warps][256 supersteps][8] clock64 stamps
  int B, T_y, T_x;
  int S;        // ring slots per warp (chunks of 32 frames in flight); LINEAR adds one mirror slot
  int W;        // DP warps covering the padded T_x
  int TXP;      // W*32*K: row stride of the decision words
  int G;        // ceil(T_y/32)
  int BR;       // hand-off ring length in frames (power of two)
  int use_tma;  // 1: 2-D TMA boxes (16-byte aligned base, T_x % 4 == 0); 0: 4-byte cp.async
  int pdl;
  // Streamed mode (mas_fused.cu, SynthesizerTrn.py:223-235 in one pass): neg_cent is not a dense tensor but a
  // per-utterance ring of RT tiles of 128 frames that the tensor-core contraction kernel (mas_neg_cent_tc.cu), running
  // at the same time on other SMs, fills in frame order; tile_flags[b][mt] reaches tile_need once frame block mt has
  // landed.  The tensor map then describes the ring ([B*RT*128][pitch]); chunks are requested as their tile arrives.
  // The CTA also zero-fills its own utterance's dense path (one paced warp) and raises fill_done[b].
  const uint32_t* tile_flags;  // nullptr = ordinary mode
  int tile_need, RT, MT;
  unsigned char* fill_out;     // this call's dense path [B][T_y][T_x] elements of fill_es bytes, or nullptr
  long long fill_bytes;        // bytes per utterance
  uint32_t* fill_done;         // [B] (stride fill_stride words), set to 1 after the fence
  int fill_stride;
  int fill_sleep;              // ns between bursts of four 512-byte store instructions
  int warm;                    // mas_dp2_kernel: 1 = an idle warp pre-executes the plain superstep variant (instruction cache)
  DpSmem sm;
};

template <int K>
__device__ __forceinline__ void lds_cols(float (&c)[K], uint32_t a) {
  if (K == 1) {
    c[0] = ptx::lds_f32(a);
  } else if (K == 2) {
    const float2 t = ptx::lds_f32x2(a);
    c[0] = t.x;
    c[1] = t.y;
  } else {
#pragma unroll
    for (int q = 0; q < K / 4; ++q) {
      const float4 t = ptx::lds_f32x4(a + 16u * q);
      c[4 * q + 0] = t.x;
      c[4 * q + 1] = t.y;
      c[4 * q + 2] = t.z;
      c[4 * q + 3] = t.w;
    }
  }
}

constexpr int kRows = 32;  // frames per chunk = steps per superstep = frames per decision word

// D: frames of skew between neighbouring lanes.  A superstep of lane l covers frames 32s-D*l .. 32s-D*l+31,
// so Q+1 = ceil(31*D/32)+1 chunks are live.
// LINEAR: the ring has S slots plus a mirror: chunk c lives in slot c % S and chunks with c % S == 0 are ALSO
// copied into slot S, so any run of 32 frames that starts inside the ring is one linear address range (a
// per-lane base plus compile-time offsets).  Otherwise (D == 1 only; not enough shared memory for the mirror)
// the address is selected per step between two per-lane bases.
// STREAMED: neg_cent arrives tile by tile from the contraction kernel (DpParams::tile_flags).  A template parameter, not a
// run-time flag, so that the ordinary kernel's code -- whose speed depends on its very layout in the instruction cache --
// is exactly what it is without the streamed mode.
template <int K, int D, bool LINEAR, bool STREAMED = false>
__global__ void __launch_bounds__(416, 1) mas_dp_kernel(const __grid_constant__ CUtensorMap tmap, const DpParams p) {
  static_assert(LINEAR || D == 1, "the select ring handles one chunk boundary per superstep");
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int R = kRows;
  constexpr int Q = (31 * D + 31) / 32;    // chunks (and decision-word groups) a superstep reaches back
  constexpr int LAG = (30 + 31 * D) / 32 + 1;  // supersteps the left neighbour must be ahead
  constexpr uint32_t ROWB = 32u * K * 4u;  // bytes of one ring row: this warp's 32*K columns of one frame
  constexpr uint32_t SLOTB = R * ROWB;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  // Warp roles.  Up to three DP warps sit on schedulers 0..2 and every helper warp (two producers, the lengths
  // warp) on scheduler 3 (warp ids 3, 7, 11; the warps in between exit at once): a helper that shares a
  // scheduler with a DP warp steals its issue slots while it polls (measured: 3-4 us on the DP).  With more DP
  // warps the helpers simply follow them.
  const int wid = tid >> 5;
  const bool spread = p.W <= 3;
  const int NP = p.W <= 3 ? 2 : 4;                                   // producer warps
  const int dw = spread ? (wid < 3 ? (wid < p.W ? wid : -1) : (wid & 3) == 3 ? p.W + (wid >> 2) : -1)
                        : wid;                                       // 0..W-1 DP, W..W+NP-1 producers, W+NP lengths
  const int lane = tid & 31;
  const int S = p.S, W = p.W, BR = p.BR;
  const int nphys = LINEAR ? S + 1 : S;

  // warps 0..W-1: DP warps; warp W: producer (requests the chunks of every DP warp's ring)
  unsigned char* ring_all = smem + p.sm.ring;
  unsigned char* ringw = ring_all + static_cast<size_t>(dw) * nphys * SLOTB;
  float* bnd = reinterpret_cast<float*>(smem + p.sm.bnd);                                  // [W+1][BR]
  uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + p.sm.bars);                      // [W][S]
  uint64_t* full = full_all + static_cast<size_t>(dw) * S;
  int* prog = reinterpret_cast<int*>(smem + p.sm.prog);                                    // [W] supersteps completed
  double* red = reinterpret_cast<double*>(smem + p.sm.red);                                // [2][32]
  int* lens_s = reinterpret_cast<int*>(red + 64);                                          // [2]

  // This kernel may have been launched programmatically behind the previous call's kernels: global
  // memory is first touched after the wait.
  constexpr bool streamed = STREAMED;
  if (!streamed) ptx::pdl_wait();  // (streamed: everything read before the tile flags is host-written or ours; see the end)
  // ---- chunk loads -----------------------------------------------------------------------------
  const long long rows_total = static_cast<long long>(p.B) * p.T_y;
  // (called by the producer warp only; w = the DP warp whose ring is filled)
  auto copy_chunk = [&](int w, int c, unsigned char* dst, uint64_t* bar) {
    if (p.use_tma) {
      if (lane == 0) {
        const int row = streamed ? b * p.RT * 128 + (c * R) % (p.RT * 128) : b * p.T_y + c * R;
        ptx::tma_load_2d(dst, &tmap, w * 32 * K, row, bar);
      }
    } else {
      // generic path (unaligned base or T_x % 4 != 0): every lane fetches K columns of the R frames
      // with 4-byte async copies; out-of-range cells are zero-filled like the TMA does
      const uint32_t d0 = ptx::smem_u32(dst) + static_cast<uint32_t>(lane) * K * 4u;
      const int xb = (w * 32 + lane) * K;
#pragma unroll 2
      for (int r = 0; r < R; ++r) {
        const long long grow = static_cast<long long>(b) * p.T_y + c * R + r;
        const bool rok = grow < rows_total;
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const bool ok = rok && (xb + j) < p.T_x;
          const float* src = ok ? p.nc + static_cast<size_t>(grow) * p.T_x + xb + j : p.nc;
          ptx::cp_async4_zfill(d0 + static_cast<uint32_t>(r) * ROWB + 4u * j, src, ok ? 4u : 0u);
        }
      }
    }
  };
  // chunk c of warp w -> slot ls = c % S (+ the mirror slot S when ls == 0)
  auto issue_chunk = [&](int w, int c, int ls) {
    uint64_t* bar = full_all + static_cast<size_t>(w) * S + ls;
    unsigned char* rw = ring_all + static_cast<size_t>(w) * nphys * SLOTB;
    const bool mirror = LINEAR && ls == 0;
    if (p.use_tma && lane == 0) ptx::mbar_arrive_expect_tx(bar, mirror ? 2u * SLOTB : SLOTB);
    copy_chunk(w, c, rw + static_cast<size_t>(ls) * SLOTB, bar);
    if (mirror) copy_chunk(w, c, rw + static_cast<size_t>(S) * SLOTB, bar);
    if (!p.use_tma) ptx::cp_async_mbar_arrive_noinc(bar);
  };

  // The first two chunks of every ring are requested right away (they only need T_y as a bound; frames beyond
  // t_y are padding that exists in memory); the producer loop requests the rest -- a TMA issue costs its warp
  // up to 350 cycles, and everybody waits for the barrier below.
  const int nspec = streamed ? 0 : min(2, (p.T_y + R - 1) / R);  // streamed: nothing exists before its tile flag says so
  if (dw == W) {
    if (lane == 0) {
      for (int i = 0; i < W * S; ++i) ptx::mbar_init(&full_all[i], p.use_tma ? 1 : 32);
      ptx::mbar_fence_init();
    }
    __syncwarp();
    for (int c = 0; c < nspec; ++c)
      for (int w = 0; w < W; ++w) issue_chunk(w, c, c);
  }

  if (!streamed && b == 0 && tid == 0) {
    p.wo_counters[0] = 0;
    p.wo_counters[1] = 0;
  }
  // Streaming mode hands the decision words to the concurrently running backtrack kernel without any
  // fence: every 32-bit word travels in one 8-byte store together with a tag, and a reader accepts an
  // element only when the tag is set -- the flag-in-data scheme of NCCL's LL protocol (a GPU-scope release
  // per group costs the DP warp ~0.8 us each).  Clear this utterance's tags (whatever the scratch held:
  // its layout depends on the shape) before the backtrack kernel can start.
  constexpr uint32_t tag = 1u;
  if (p.lenstag) {
    uint4* z = reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(p.bits) + static_cast<size_t>(b) * p.G * p.TXP);
    const int n16 = p.G * p.TXP / 2;  // TXP is a multiple of 32
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) *reinterpret_cast<unsigned long long*>(p.lenstag + b) = 0ull;
  }
  if (tid == 0) tl_min(p.tl, 0);
#ifdef MAS_TRACE
  if (p.trace && tid == 0) p.trace[8 * 256 * 8 + 2 * b] = globaltimer_ns();
#endif

  // hand-off rings and progress counters (independent of the lengths)
  for (int i = tid; i < (W + 1) * BR; i += blockDim.x) bnd[i] = (i == 0) ? 0.0f : kNeg;  // (0,0): v_prev = 0 (core.pyx:22-23)
  if (tid < W) prog[tid] = 0;  // supersteps completed

  volatile int* lens_v = lens_s;  // [0] t_y, [1] t_x, [2] 1 once they are known, [3] 1 = give up (watchdog, streamed mode)
  if (tid == 0) {
    lens_s[2] = 0;
    if (streamed) lens_s[3] = 0;
  }
  __syncthreads();  // rings, barriers, hand-off arrays and flags are initialised

  // ---- lengths: computed by a warp of their own WHILE the DP already runs -----------------------------
  // Nothing in the recurrence needs them: every DP warp runs (columns >= t_x compute garbage nobody reads) and
  // t_y only says when to stop.  With a cold mask the column walk is DRAM-row-activation bound (64 utterances x
  // 1024 strided elements: ~8 us measured), which this takes off the critical path.
  if (dw < 0) return;  // filler warps
  if (dw == W + NP) {
    // Counters and tags were cleared before the barrier above: make that visible GPU-wide, then let the
    // dependent kernels start.  Only this warp pays for the fence; the DP warps are already running.
    if (lane == 0) {
      __threadfence();  // (cumulative: covers the other threads' stores ordered by the barrier)
      ptx::pdl_launch_dependents();
    }
    int t_y, t_x;
    if (p.t_ys != nullptr) {
      t_y = p.t_ys[b];
      t_x = p.t_xs[b];
    } else {
      double sy, sx;
      mask_sums(p.mask, p.mask_dtype, static_cast<int64_t>(b) * p.msb, p.msy, p.T_y, p.msx, p.T_x, lane, 32, sy, sx);
      t_y = static_cast<int>(warp_sum(sy));
      t_x = static_cast<int>(warp_sum(sx));
    }
    int st = 0;
    if (t_y < 1 || t_x < 1) st |= MAS_STATUS_EMPTY;
    if (t_y > p.T_y || t_x > p.T_x) st |= MAS_STATUS_TOO_LONG;
    if (t_x > t_y) st |= MAS_STATUS_TX_GT_TY;
    if (st) t_y = t_x = 0;  // the path of this utterance stays all-zero
    if (lane == 0) {
      if (st) raise_status(p.status, p.mirror, st);
      p.lens[2 * b] = t_y;
      p.lens[2 * b + 1] = t_x;
      if (p.lenstag)  // streaming backtrack: the lengths are published
        *reinterpret_cast<unsigned long long*>(p.lenstag + b) = pack_tagged((static_cast<uint32_t>(t_y) << 12) | static_cast<uint32_t>(t_x), tag);
      lens_v[0] = t_y;
      lens_v[1] = t_x;
      __threadfence_block();
      lens_v[2] = 1;
      tl_max(p.tl, 7);  // (debug timeline) lengths known
    }
    if (streamed) {
      // This warp has nothing else to do: it zero-fills the utterance's dense path (np.zeros, __init__.py:15) with
      // paced 16-byte stores -- 0.8 MB over the tens of microseconds the DP takes is one store instruction per ~40
      // cycles, on the scheduler no DP warp uses -- and then publishes fill_done[b] for the backtrack kernel.
      if (p.fill_out) {
        unsigned char* beg = p.fill_out + static_cast<size_t>(b) * p.fill_bytes;
        unsigned char* end = beg + p.fill_bytes;
        unsigned char* abeg = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(beg) + 15u) & ~uintptr_t(15));
        if (abeg > end) abeg = end;
        unsigned char* aend = abeg + ((end - abeg) & ~ptrdiff_t(15));
        for (unsigned char* q = beg + lane; q < abeg; q += 32) *q = 0;
        for (unsigned char* q = aend + lane; q < end; q += 32) *q = 0;
        uint4* a4 = reinterpret_cast<uint4*>(abeg);
        const size_t n4 = static_cast<size_t>(aend - abeg) >> 4;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (size_t i = lane; i < n4; i += 128) {
          a4[i] = z4;
          if (i + 32 < n4) a4[i + 32] = z4;
          if (i + 64 < n4) a4[i + 64] = z4;
          if (i + 96 < n4) a4[i + 96] = z4;
          if (p.fill_sleep) __nanosleep(p.fill_sleep);
        }
        __syncwarp();
        if (lane == 0) {
          __threadfence();
          *reinterpret_cast<volatile uint32_t*>(p.fill_done + static_cast<size_t>(b) * p.fill_stride) = 1u;
        }
      }
      // "this grid complete" must imply "the contraction grid complete" for whatever follows in the stream (we were
      // launched programmatically behind it and never waited for it): wait here, where it costs nothing.
      ptx::pdl_wait();
    }
    return;
  }
  // the DP and producer warps learn the lengths when they are there
  bool known = false;
  int NS = ((p.T_y - 1) >> 5) + Q + 1;       // supersteps: the words of group g are complete after superstep g+Q
  int nchunks = (p.T_y + R - 1) / R;          // chunks to stream (until t_y is known: everything that exists)
  auto check_lens = [&]() {
    if (!known && lens_v[2] != 0) {
      known = true;
      const int t_y = lens_v[0];
      NS = t_y > 0 ? ((t_y - 1) >> 5) + Q + 1 : 0;
      nchunks = (t_y + R - 1) / R;
    }
    if (streamed && lens_v[3] != 0) {  // a producer gave up waiting for the contraction kernel: wind down at once
      known = true;
      NS = 0;
      nchunks = 0;
    }
  };
  if (dw >= W) {
    // ---- producer warps (two: issuing one TMA box costs its warp 100-350 cycles, and three rings need three
    // boxes per ~1000-cycle superstep): keep every active ring full.  Chunk c of warp w may be requested once the
    // warp has finished superstep c-S+Q (all its lanes are past chunk c-S): they watch the progress counters.
    // Producer q of NP serves the rings w with w % NP == q; lane w tracks ring w. ----
    const int q = dw - W;
    int ci = nspec;        // lane w: next chunk of warp w's ring ...
    int cs = nspec % S;    // ... and its slot
    const bool mine = lane < W && (lane % NP) == q;
    int tiles_ready = 0;  // streamed: frame blocks [0, tiles_ready) of this utterance have landed in the ring
    unsigned long long t_wd = streamed ? globaltimer_ns() : 0ull;
    for (;;) {
      check_lens();
      bool want = mine && ci < nchunks;
      if (streamed) {
        // no chunk before the lengths are known (tiles past t_y are never produced) nor before its tile's flag
        want = want && known;
        // Tile flags are polled AHEAD of need and 32 at a time (lane l looks at frame block tiles_ready + l): one L2
        // round trip of this warp finds every tile that has landed.  (Polling one flag when a ring ran into it cost the
        // producer an acquire round trip + proxy fence per tile, right when its rings needed chunks: DP 47 us vs 34.)
        const int tiles_total = known ? (lens_v[0] + 127) >> 7 : 0;
        if (tiles_ready < tiles_total) {
          const int mt = tiles_ready + lane;
          const bool landed = mt < tiles_total &&
                              ptx::ld_acquire_gpu_u32(p.tile_flags + static_cast<size_t>(b) * p.MT + mt) >= static_cast<uint32_t>(p.tile_need);
          const unsigned lb = __ballot_sync(0xffffffffu, landed);
          const int cnt = lb == 0xffffffffu ? 32 : __ffs(~lb) - 1;  // consecutive landed tiles from tiles_ready on
          __syncwarp();  // the acquiring lanes' observations are ordered before lane 0's TMA requests
          if (cnt > 0) {
            tiles_ready += cnt;
            asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of the contraction kernel -> TMA reads
            t_wd = globaltimer_ns();
          } else if (globaltimer_ns() - t_wd > 2000000000ull) {
            // Watchdog: the contraction kernel is an EARLIER kernel of the stream; if no tile arrives for 2 s it
            // died.  Flag it (the host raises on its next call), wind the CTA down, leave an all-zero path.
            if (lane == 0) {
              raise_status(p.status, p.mirror, MAS_STATUS_TIMEOUT);
              lens_v[3] = 1;
            }
            __syncwarp();
          }
        }
        want = want && (ci >> 2) < tiles_ready;
      }
      const bool ready = want && ptx::ld_volatile_s32(&prog[lane]) >= ci - S + Q + 1;
      unsigned m = __ballot_sync(0xffffffffu, ready);
      if (known && !__any_sync(0xffffffffu, mine && ci < nchunks)) break;
      if (m == 0u) __nanosleep(64);
      while (m) {
        const int w = __ffs(m) - 1;
        m &= m - 1;
        const int c = __shfl_sync(0xffffffffu, ci, w);
        const int sl = __shfl_sync(0xffffffffu, cs, w);
        issue_chunk(w, c, sl);
        if (lane == w) {
          ++ci;
          if (++cs == S) cs = 0;
        }
      }
    }
    // chunks that were requested before t_y was known and will never be consumed: no copy may outlive the CTA
    // (at most one per slot can be outstanding: chunk c+S is only requested once chunk c has been consumed)
    if (mine)
      for (int c = max(nchunks, ci - S); c < ci; ++c) ptx::mbar_wait(full_all + static_cast<size_t>(lane) * S + (c % S), (c / S) & 1);
    return;
  }

  // ---- DP warp: columns [x0, x0+K) per lane ----------------------------------------------------
  // Hand-off: warp dw reads its left edge from bnd[dw] and publishes its last column to bnd[dw+1]
  // (slot (y+1) & (BR-1) holds frame y's value).  bnd[0] is constant (the x == 0 sentinel).
  const int x0 = (dw * 32 + lane) * K;
  const bool has_left = dw > 0;
  const bool has_right = dw < W - 1;
  const bool lane0 = lane == 0;
  const bool lane31 = lane == 31;
  const uint32_t bnd_in = ptx::smem_u32(bnd + static_cast<size_t>(dw) * BR);
  float* bnd_out = bnd + static_cast<size_t>(dw + 1) * BR;
  const bool ll = p.lenstag != nullptr;
  uint32_t* bits_b = p.bits + (static_cast<size_t>(b) * p.G * p.TXP + x0) * (ll ? 2 : 1);
  // lane-private diagonal: column x0+j is the diagonal cell x == y of frame y = t - D*lane when t == dt + j
  const int dt = x0 + D * lane;
  const int diag_lo = dw * 32 * K, diag_hi = dw * 32 * K + 31 * (K + D) + K - 1;  // steps at which some lane is diagonal
  // decision words: lane l finishes group g's 32 frames partly in superstep g + a and partly in g + a + 1
  const int hsel = (D * lane) >> 5;    // a
  const int hsh = (D * lane) & 31;

  float v[K];
  uint32_t hist[Q + 1][K];  // hist[k] = the 32 decision bits of superstep s-k (hist[0] is being filled)
#pragma unroll
  for (int j = 0; j < K; ++j) {
    v[j] = kNeg;
#pragma unroll
    for (int k = 0; k <= Q; ++k) hist[k][j] = 0u;
  }
  float left[D];  // value[y-1][x0-1] for the next D steps, fetched from the lane to the left D steps ahead
#pragma unroll
  for (int k = 0; k < D; ++k) left[k] = kNeg;

  const uint32_t ring_lane = ptx::smem_u32(ringw) + static_cast<uint32_t>(lane) * K * 4u;
  const int ring_frames = S * R;
  int foff = (ring_frames * 4 - D * lane) % ring_frames;  // ring position of the lane's first frame of superstep s
  int ls = 0;          // slot of chunk s
  uint32_t par = 0u;   // parity of its current use

  // State that crosses superstep boundaries: the first three frames and the first eight left-edge values of
  // the NEXT superstep are loaded during the last steps of the current one when its inputs were found
  // ready by non-blocking probes in mid-superstep (`pre`); only otherwise does a superstep start by waiting.
  float cb[4][K], e[2][8];
  bool pre = false;

  auto superstep = [&](int s, auto first_tag, auto diag_tag) {
    constexpr bool FIRST = decltype(first_tag)::value;
    constexpr bool DIAG = decltype(diag_tag)::value;
#ifdef MAS_TRACE  // compile-time only: the stamps cost ~30 cycles each
    unsigned long long* tr = (p.trace && b == 0 && lane0 && s < 256) ? p.trace + (static_cast<size_t>(dw) * 256 + s) * 8 : nullptr;
#else
    constexpr unsigned long long* tr = nullptr;
#endif
    if (tr) tr[0] = clock64();
    // per-lane base address of frame 32s - D*lane, now and one superstep later
    auto bases = [&](int slot, int fo, uint32_t& pc, uint32_t& pp) {
      if (LINEAR) {
        pc = ring_lane + static_cast<uint32_t>(fo) * ROWB;
        pp = 0;
      } else {
        const int ps = slot == 0 ? S - 1 : slot - 1;
        pc = ring_lane + static_cast<uint32_t>(slot) * SLOTB - static_cast<uint32_t>(lane) * ROWB;
        pp = ring_lane + static_cast<uint32_t>(ps) * SLOTB + static_cast<uint32_t>(32 - lane) * ROWB;
      }
    };
    const int ls_n = ls + 1 == S ? 0 : ls + 1;
    const uint32_t par_n = ls + 1 == S ? par ^ 1u : par;
    const int foff_n = foff + R >= ring_frames ? foff + R - ring_frames : foff + R;
    uint32_t pcur, pprev, pcur_n, pprev_n;
    bases(ls, foff, pcur, pprev);
    bases(ls_n, foff_n, pcur_n, pprev_n);
    // FIRST: step i of superstep s (+1) works on a frame < 0 while i < fneg (fneg_n)
    const int fneg = D * lane - 32 * s, fneg_n = fneg - 32;
    auto load_row = [&](uint32_t pc, uint32_t pp, int fn, int i, float (&c)[K]) {
      uint32_t a;
      if (LINEAR) a = pc + static_cast<uint32_t>(i) * ROWB;
      else a = (i >= lane ? pc : pp) + static_cast<uint32_t>(i) * ROWB;
      lds_cols<K>(c, a);
      if (FIRST) {  // frame < 0: contributes nothing, the row stays at the sentinel
#pragma unroll
        for (int j = 0; j < K; ++j) c[j] = (i < fn) ? 0.0f : c[j];
      }
    };
    auto load_e = [&](uint32_t ea, int blk, float (&e8)[8]) {
      const float4 e0 = ptx::lds_f32x4(ea + 32u * blk);
      const float4 e1 = ptx::lds_f32x4(ea + 32u * blk + 16u);
      e8[0] = e0.x; e8[1] = e0.y; e8[2] = e0.z; e8[3] = e0.w;
      e8[4] = e1.x; e8[5] = e1.y; e8[6] = e1.z; e8[7] = e1.w;
    };
    uint32_t ea = bnd_in + 4u * static_cast<uint32_t>((32 * s) & (BR - 1));
    uint32_t ea_n = bnd_in + 4u * static_cast<uint32_t>((32 * s + 32) & (BR - 1));
    // the left neighbour's last column for frames <= 32s+30 is finished by its lane 31 in superstep s+LAG-1;
    // the right neighbour must have consumed the hand-off slots this superstep overwrites
    const int need_r = s - BR / 32 + 1;
    if (!pre) {
      // ---- blocking start (first superstep, or an input was late) ----
      // frames 32s..32s+31 landed (or turn out not to exist: t_y became known and is smaller)
      while (s < nchunks && !ptx::mbar_test(&full[ls], par)) check_lens();
      load_row(pcur, pprev, fneg, 0, cb[0]);
      load_row(pcur, pprev, fneg, 1, cb[1]);
      load_row(pcur, pprev, fneg, 2, cb[2]);
      int fl = 0;
      if (has_left) {
        for (;;) {
          fl = ptx::ld_volatile_s32(&prog[dw - 1]);
          if (fl >= min(s + LAG, NS)) break;
          check_lens();  // the neighbour may have stopped at a smaller NS than the one assumed so far
        }
      }
      if (has_right && need_r > 0)
        while (ptx::ld_volatile_s32(&prog[dw + 1]) < need_r) {
          if (streamed) {
            check_lens();
            if (NS == 0) break;
          }
        }
      // The hand-off values are plain shared-memory loads: give their address a (null) data dependency on
      // the flag just read, or ptxas is free to hoist them above the polling loop.
      ea += static_cast<uint32_t>(fl) >> 31;
      load_e(ea, 0, e[0]);
    }
    if (tr) tr[1] = clock64();
    // lane 31 publishes frame 32s+i-31D into slot (frame+1) = 32(s-Q) + R0 + i
    constexpr int R0 = 32 * Q - (31 * D - 1);  // 2, 3, 4 for D = 1, 2, 3
    float* bo1 = bnd_out + ((32 * (s - Q)) & (BR - 1)) + R0;
    float* bo2 = bnd_out + ((32 * (s - Q + 1)) & (BR - 1)) - (32 - R0);
    const int dd = dt - 32 * s;  // DIAG: column j of this lane is diagonal at step i == dd + j
    bool chunk_n = true;
    int pl = 0x7fffffff, pr = 0x7fffffff;

#pragma unroll
    for (int i = 0; i < 32; ++i) {
      if (i + 3 < 32) load_row(pcur, pprev, fneg, i + 3, cb[(i + 3) & 3]);
      if ((i & 7) == 0 && i + 8 < 32) load_e(ea, i / 8 + 1, e[(i / 8 + 1) & 1]);
      if (i == 12) {  // non-blocking probes of the next superstep's inputs
        if (s + 1 < nchunks) chunk_n = ptx::mbar_test(&full[ls_n], par_n);
        if (has_left) pl = ptx::ld_volatile_s32(&prog[dw - 1]);
        if (has_right) pr = ptx::ld_volatile_s32(&prog[dw + 1]);
        ea_n += static_cast<uint32_t>(pl) >> 31;  // (null) dependency: the prefetch below stays behind this probe
      }
      // ... and, ready or not, its first loads (discarded when it was not ready)
      if (i == 28) load_e(ea_n, 0, e[0]);
      if (i >= 29) load_row(pcur_n, pprev_n, fneg_n, i - 29, cb[i - 29]);
      const float (&c)[K] = cb[i & 3];
      const float nxt = __shfl_up_sync(0xffffffffu, v[K - 1], 1);  // for step i+D
      const float le = lane0 ? e[(i / 8) & 1][i & 7] : left[0];
      if (DIAG) {
        // x == y: the "stay" candidate value[y-1][y] is the sentinel (core.pyx:17-18; that cell is outside
        // the band, so overwriting the register copy is harmless) ...
#pragma unroll
        for (int jj = 0; jj < K; ++jj) v[jj] = (i == dd + jj) ? kNeg : v[jj];
      }
#pragma unroll
      for (int jj = K - 1; jj >= 1; --jj) {
        const float d = v[jj] - v[jj - 1];                                    // sign bit == (stay < step), core.pyx:32
        hist[0][jj] = __funnelshift_l(__float_as_uint(d), hist[0][jj], 1);   // (bits << 1) | sign
        v[jj] = c[jj] + fmaxf(v[jj - 1], v[jj]);                              // core.pyx:28
      }
      const float d = v[0] - le;
      hist[0][0] = __funnelshift_l(__float_as_uint(d), hist[0][0], 1);
      v[0] = c[0] + fmaxf(le, v[0]);
      if (DIAG) {
        // ... and the backtrack is forced to step there (core.pyx:32 `index == y`)
#pragma unroll
        for (int jj = 0; jj < K; ++jj) hist[0][jj] |= (i == dd + jj) ? 1u : 0u;
      }
      if (lane31) ptx::st_volatile_f32((i < 32 - R0 ? bo1 : bo2) + i, v[K - 1]);
#pragma unroll
      for (int k = 0; k + 1 < D; ++k) left[k] = left[k + 1];
      left[D - 1] = nxt;
    }
    if (tr) tr[3] = clock64();
    __syncwarp();
    // frames <= 32(s+1)-31D-1 of our last column are published; every lane is past chunk s-Q (the producer may refill it)
    if (lane31) ptx::st_volatile_s32(&prog[dw], s + 1);
    if (FIRST && s == 0 && dw == 0 && lane0) bnd[0] = kNeg;  // the (0,0) special case is consumed
    pre = chunk_n && pl >= min(s + 1 + LAG, NS) && pr >= need_r + 1;
    ls = ls_n;
    par = par_n;
    foff = foff_n;
    // decision words of group s-Q: its frames sit in hist[Q-a] (older part) and hist[Q-a-1], a = D*lane/32
    if (s >= Q) {
      uint32_t w[K];
#pragma unroll
      for (int jj = 0; jj < K; ++jj) {
        uint32_t hi = hist[Q][jj], lo = hist[Q - 1][jj];
#pragma unroll
        for (int a = 1; a < Q; ++a) {
          hi = (hsel == a) ? hist[Q - a][jj] : hi;
          lo = (hsel == a) ? hist[Q - a - 1][jj] : lo;
        }
        w[jj] = __funnelshift_l(lo, hi, hsh);
      }
      if (x0 == 0) w[0] = 0u;  // the backtrack never leaves column 0 (core.pyx:32 `index != 0`)
      if (ll) {
        uint2* dst = reinterpret_cast<uint2*>(bits_b) + static_cast<size_t>(s - Q) * p.TXP;
        if (K % 2 == 0) {
#pragma unroll
          for (int q = 0; q < K / 2; ++q)
            ptx::st_global_v2_u64(dst + 2 * q, pack_tagged(w[2 * q], tag), pack_tagged(w[2 * q + 1], tag));
        } else {
#pragma unroll
          for (int jj = 0; jj < K; ++jj) *reinterpret_cast<unsigned long long*>(dst + jj) = pack_tagged(w[jj], tag);
        }
      } else {
        uint32_t* dst = bits_b + static_cast<size_t>(s - Q) * p.TXP;
        if (K % 4 == 0) {
#pragma unroll
          for (int q = 0; q < K / 4; ++q)
            *reinterpret_cast<uint4*>(dst + 4 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        } else if (K == 2) {
          *reinterpret_cast<uint2*>(dst) = make_uint2(w[0], w[1]);
        } else {
#pragma unroll
          for (int jj = 0; jj < K; ++jj) dst[jj] = w[jj];
        }
      }
    }
#pragma unroll
    for (int k = Q; k >= 1; --k)
#pragma unroll
      for (int jj = 0; jj < K; ++jj) hist[k][jj] = hist[k - 1][jj];
    if (tr) tr[7] = clock64();
  };

  // supersteps [0, Q) work on frames < 0 in some lanes (FIRST); the diagonal crosses this warp's columns in
  // supersteps [sd0, sd1]; everything else runs the plain variant
  const int sd0 = diag_lo >> 5, sd1 = diag_hi >> 5;
  for (int s = 0;; ++s) {
    check_lens();
    if (s >= NS) {
      if (known) break;
      while (!known) check_lens();  // ran through every frame that exists before the lengths arrived
      if (s >= NS) break;
    }
    const bool diag = s >= sd0 && s <= sd1;
    if (s < Q) {
      if (diag) superstep(s, std::true_type{}, std::true_type{});
      else superstep(s, std::true_type{}, std::false_type{});
    } else if (diag) {
      superstep(s, std::false_type{}, std::true_type{});
    } else {
      superstep(s, std::false_type{}, std::false_type{});
    }
  }
  if (lane0) tl_max(p.tl, 1);
  if (lane0) tl_max(p.tl, 2);
#ifdef MAS_TRACE
  if (p.trace && lane0) atomicMax(p.trace + 8 * 256 * 8 + 2 * b + 1, globaltimer_ns());
#endif
}

template <int K, int D, bool LINEAR, bool STREAMED = false>
inline cudaError_t launch_dp_t(const CUtensorMap& tmap, const DpParams& p, cudaStream_t st) {
  auto kern = mas_dp_kernel<K, D, LINEAR, STREAMED>;
  static std::atomic<uint64_t> attr_set{0};  // per instantiation and device
  if (cudaError_t e = ensure_dyn_smem(kern, 227 * 1024, attr_set); e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.B);
  cfg.blockDim = dim3(p.W <= 3 ? 32 * 12 : 32 * (p.W + 4 + 1));  // see the warp roles in the kernel
  cfg.dynamicSmemBytes = p.sm.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, tmap, p);
}

// skew: frames between neighbouring lanes (1..3); linear: ring with mirror slot (required for skew > 1)
template <int K>
inline cudaError_t launch_dp(const CUtensorMap& tmap, const DpParams& p, int skew, bool linear, cudaStream_t st) {
  if (p.tile_flags != nullptr) {  // streamed source: linear ring only, skew 1 or 2
    if (!linear) return cudaErrorInvalidValue;
    switch (skew) {
      case 1: return launch_dp_t<K, 1, true, true>(tmap, p, st);
      case 2: return launch_dp_t<K, 2, true, true>(tmap, p, st);
      default: return cudaErrorInvalidValue;
    }
  }
  if (!linear) return skew == 1 ? launch_dp_t<K, 1, false>(tmap, p, st) : cudaErrorInvalidValue;
  switch (skew) {
    case 1: return launch_dp_t<K, 1, true>(tmap, p, st);
    case 2: return launch_dp_t<K, 2, true>(tmap, p, st);
    case 3: return launch_dp_t<K, 3, true>(tmap, p, st);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_dp_k1(const CUtensorMap& tmap, const DpParams& p, int skew, bool linear, cudaStream_t st);
cudaError_t launch_dp_k2(const CUtensorMap& tmap, const DpParams& p, int skew, bool linear, cudaStream_t st);
cudaError_t launch_dp_k4(const CUtensorMap& tmap, const DpParams& p, int skew, bool linear, cudaStream_t st);

}  // namespace mas
