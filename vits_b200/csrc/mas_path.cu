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
#include <cstdlib>
#include <cuda_runtime.h>

#include <cuda.h>

#include "mas_forward.cuh"
#include "mas_dp.cuh"
#include "mas_dp2.cuh"

namespace mas {

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
  unsigned long long* tl;
};

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
  if (tid == 0) tl_min(p.tl, 3);

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
  if (tid == 0) tl_max(p.tl, 4);
}

// ------------------------------------------------------------------------------------------------
// K3: dense path write-out
// ------------------------------------------------------------------------------------------------
struct WoParams {
  unsigned char* out;    // [B*T_y][T_x] elements of es bytes
  const int32_t* index;  // [B*T_y]
  int32_t* counters;     // [0] next zero-fill chunk to claim, [1] chunks finished (zeroed by the forward kernel)
  long long rows;        // B*T_y
  long long bytes;       // rows * T_x * es
  int nchunks;           // zero-fill chunks of kWoChunk bytes
  int T_x;
  int es;                // element size in bytes
  unsigned long long one;  // bit pattern of 1 in the element type
  int ones;                // 1: phase B (the ones) runs here; 0: the streaming backtrack kernel drops them
  unsigned long long* tl;
};
constexpr long long kWoChunk = 32 * 1024;

__device__ __forceinline__ void store_elem(unsigned char* p, int es, unsigned long long bits) {
  switch (es) {
    case 1: *p = static_cast<unsigned char>(bits); break;
    case 2: *reinterpret_cast<uint16_t*>(p) = static_cast<uint16_t>(bits); break;
    case 4: *reinterpret_cast<uint32_t*>(p) = static_cast<uint32_t>(bits); break;
    default: *reinterpret_cast<unsigned long long*>(p) = bits; break;
  }
}

// Phase A (zero-fill) is work-stealing: whichever CTAs are resident claim 32 KiB chunks from a counter
// until none are left, so the fill is done by the CTAs that land on SMs the forward kernel leaves idle
// (the launch requests enough shared memory that a write-out CTA cannot share an SM with a forward
// CTA -- its stores would slow the latency-bound DP warps) and CTAs that only become resident after
// the forward kernel has finished find nothing left to do.  Phase B (the ones) waits for the forward /
// backtrack kernels (griddepcontrol.wait) and for every chunk to be finished.
__global__ void __launch_bounds__(256) mas_writeout_kernel(const WoParams p) {
  __shared__ int s_chunk;
  const int tid = threadIdx.x;
  ptx::pdl_launch_dependents();  // the next call's forward kernel may set up while we run
  if (tid == 0) tl_min(p.tl, 5);
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  for (;;) {
    if (tid == 0) s_chunk = atomicAdd(&p.counters[0], 1);
    __syncthreads();
    const int c = s_chunk;
    __syncthreads();
    if (c >= p.nchunks) break;
    const long long b0 = static_cast<long long>(c) * kWoChunk;
    const long long b1 = min(p.bytes, b0 + kWoChunk);
    unsigned char* beg = p.out + b0;
    unsigned char* end = p.out + b1;
    // out is aligned to its element size only: unaligned head/tail bytes go element-wise
    unsigned char* abeg = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(beg) + 15u) & ~uintptr_t(15));
    if (abeg > end) abeg = end;
    unsigned char* aend = abeg + ((end - abeg) & ~ptrdiff_t(15));
    for (unsigned char* q = beg + tid; q < abeg; q += blockDim.x) *q = 0;
    for (unsigned char* q = aend + tid; q < end; q += blockDim.x) *q = 0;
    uint4* a4 = reinterpret_cast<uint4*>(abeg);
    const size_t n4 = static_cast<size_t>(aend - abeg) >> 4;
    size_t i = tid;
    for (; i + 3 * blockDim.x < n4; i += 4 * blockDim.x) {
      a4[i] = z;
      a4[i + blockDim.x] = z;
      a4[i + 2 * blockDim.x] = z;
      a4[i + 3 * blockDim.x] = z;
    }
    for (; i < n4; i += blockDim.x) a4[i] = z;
    __syncthreads();
    if (tid == 0) {
      __threadfence();  // release: the chunk's zeros are visible before the count
      atomicAdd(&p.counters[1], 1);
    }
  }
  if (tid == 0) tl_max(p.tl, 6);
  if (!p.ones) return;
  // phase B: the ones
  ptx::pdl_wait();
  if (tid == 0) {
    while (atomicAdd(&p.counters[1], 0) < p.nchunks) __nanosleep(64);
    __threadfence();  // acquire
  }
  __syncthreads();
  const long long per = (p.rows + gridDim.x - 1) / gridDim.x;
  const long long r0 = min(p.rows, per * blockIdx.x);
  const long long r1 = min(p.rows, r0 + per);
  for (long long r = r0 + tid; r < r1; r += blockDim.x) {
    const int x = p.index[r];
    if (x >= 0) store_elem(p.out + (static_cast<size_t>(r) * p.T_x + x) * p.es, p.es, p.one);
  }
  if (tid == 0) tl_max(p.tl, 7);
}

// ------------------------------------------------------------------------------------------------
// K2 (streaming): backtrack concurrent with the forward kernel
// ------------------------------------------------------------------------------------------------
struct BsParams {
  const uint2* bits;        // [B][G][TXP] {decision word, tag} written by K1
  const uint2* lenstag;     // [B] {t_y << 12 | t_x}
  int32_t* index;           // [B][T_y] or nullptr
  unsigned char* path;      // [B][T_y][T_x] elements of es bytes, or nullptr
  const int32_t* fill_counters;  // [1] = zero-fill chunks finished
  int nchunks;
  const uint32_t* fill_flag;     // streamed source: fill_flag[b * fill_stride] = 1 once utterance b's plane is zero-filled
  int fill_stride;
  int T_y, T_x, TXP, G;
  int TXS;                  // shared-memory row stride in words (odd)
  int cols_per_warp;        // 32*K of the forward kernel: ceil(t_x / cols_per_warp) warps arrive per group
  int32_t* status;          // sticky MAS_STATUS_* bits (MAS_STATUS_TIMEOUT: a poll below gave up)
  int32_t* mirror;          // host-mapped copy of the status bits (one word per bit) or nullptr
  int lazy;                 // 1: the group walks do not wait for the lengths (wavefront forward kernels: every column of every
                            // group gets its tagged word whatever the lengths are)
  int dec16;                // 1: tables hold 16-bit exit columns (long utterances), the per-frame index is re-walked from the words
  int es;
  unsigned long long one;
  unsigned long long* tl;
};

__device__ __forceinline__ void store_one(unsigned char* p, int es, unsigned long long bits) {
  switch (es) {
    case 1: *p = static_cast<unsigned char>(bits); break;
    case 2: *reinterpret_cast<uint16_t*>(p) = static_cast<uint16_t>(bits); break;
    case 4: *reinterpret_cast<uint32_t*>(p) = static_cast<uint32_t>(bits); break;
    default: *reinterpret_cast<unsigned long long*>(p) = bits; break;
  }
}

// Walk one 32-frame group from QP*32 entry columns per pass (QP independent chains per lane hide the
// shared-memory latency) and record, per entry column, the 32 decisions taken: bit r = "stepped left
// when leaving frame r".  The exit column is entry - popc(word); the column at frame r is
// entry - popc(word >> (r+1)).
template <int QP>
__device__ __forceinline__ void walk_group(const uint32_t* row, uint32_t* dec, uint16_t* ex, int t_x, int lane, int e_lo = 0) {
  const uint32_t row_addr = ptx::smem_u32(row);
  for (int e0 = e_lo + lane; e0 < t_x; e0 += 32 * QP) {  // entry columns [e_lo, t_x)
    uint32_t addr[QP], dw[QP];
#pragma unroll
    for (int q = 0; q < QP; ++q) {
      addr[q] = row_addr + 4u * static_cast<uint32_t>(min(e0 + 32 * q, t_x - 1));
      dw[q] = 0u;
    }
#pragma unroll
    for (int r = 31; r >= 0; --r) {
#pragma unroll
      for (int q = 0; q < QP; ++q) {
        uint32_t w;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr[q]));
        // one test and two predicated updates per step (the arithmetic form -- shift, mask, scale, add, twice -- was ~9
        // instructions per step and entry column, and this walk is issue-bound: 1700 instructions = 2.7 us per group at c2)
        if (w & (1u << (31 - r))) {
          dw[q] |= 1u << r;  // the decision of frame r in bit r
          addr[q] -= 4u;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < QP; ++q)
      if (e0 + 32 * q < t_x) {
        if (dec) dec[e0 + 32 * q] = dw[q];
        else ex[e0 + 32 * q] = static_cast<uint16_t>(e0 + 32 * q - __popc(dw[q]));
      }
  }
}

constexpr int kBallotGroups = 1;  // top groups walked from their single real entry instead of tabulated (see the kernel)

__global__ void __launch_bounds__(512, 2) mas_backtrack_stream_kernel(const BsParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int nw = blockDim.x >> 5;
  uint32_t* sdec = reinterpret_cast<uint32_t*>(smem);                       // [G][TXS] walk decisions, or ...
  uint16_t* sexit = reinterpret_cast<uint16_t*>(smem);                      // ... [G][TXS] exit columns (dec16)
  uint32_t* sstage = sdec + ((static_cast<size_t>(p.G) * p.TXS) >> (p.dec16 ? 1 : 0)) + 1;  // [nw][TXS] a group's words
  int* sentry = reinterpret_cast<int*>(sstage + static_cast<size_t>(nw) * p.TXS);  // [G]
  int* smisc = sentry + p.G;                                                // [5]: t_y, t_x, entry below the top group, top decisions, timed out

  if (tid == 0) tl_min(p.tl, 3);
  // Watchdog: everything this kernel polls is produced by EARLIER kernels of the stream, so a poll can only
  // fail to terminate if one of them died (or the context was descheduled for seconds).  Give up after 2 s rather
  // than hang -- and never emit a path built from words that did not arrive: the utterance keeps an all-zero path
  // and index -1, MAS_STATUS_TIMEOUT is raised in the scratch word AND in the host-mapped mirror, where the host
  // side sees it on its next call without synchronising (vits_b200 raises MasError then).
  const unsigned long long t_start = globaltimer_ns();
  volatile int* dead = smisc + 4;
  if (tid == 0) *dead = 0;
  auto expired = [&]() {
    if (globaltimer_ns() - t_start < 2000000000ull) return false;
    raise_status(p.status, p.mirror, MAS_STATUS_TIMEOUT);
    *dead = 1;
    return true;
  };
  // The lengths.  Eager form: wait for them, then start on the groups.  Lazy form (wavefront forward kernels): the group
  // walks need neither -- every column of every group arrives tagged, and an entry column's walk only ever moves left, so
  // walking all T_x entry columns gives the same exits for the real ones -- and t_y only says which group is the top one.
  // With the lengths taken from the mask they arrive ~10 us into a c2 call: the eager form then started 5 us late and did
  // not catch up before the end of the call (tools/mask_traffic.py: 1.4 us per call).
  int t_y = -1, t_x = p.T_x;
  bool known = false;
  if (!p.lazy) {
    if (tid == 0) {
      uint2 lt;
      while ((lt = load_tagged(p.lenstag + b)).y != 1u) {
        if (expired()) {
          lt = make_uint2(0u, 1u);
          break;
        }
        __nanosleep(100);
      }
      smisc[0] = static_cast<int>(lt.x >> 12);
      smisc[1] = static_cast<int>(lt.x & 4095u);
    }
    __syncthreads();
    t_y = smisc[0];
    t_x = smisc[1];
    known = true;
  }
  auto poll_lens = [&]() {  // (warp-uniform)
    if (known) return;
    uint2 lt = make_uint2(0u, 0u);
    if (lane == 0) lt = load_tagged(p.lenstag + b);
    lt.x = __shfl_sync(0xffffffffu, lt.x, 0);
    lt.y = __shfl_sync(0xffffffffu, lt.y, 0);
    if (lt.y == 1u) {
      known = true;
      t_y = static_cast<int>(lt.x >> 12);
      t_x = static_cast<int>(lt.x & 4095u);
    }
  };
  constexpr uint32_t tag = 1u;  // cleared by the forward kernel before this kernel can start
  auto finish = [&]() {
    // This kernel is the last of the call's chain and the next call's forward kernel waits only for it:
    // completing after the fill kernel (our programmatic predecessor, which never blocks) makes "K2 done"
    // imply "K3 done", so a late fill CTA can never see the next call's re-armed chunk counter.
    ptx::pdl_wait();
  };
  int32_t* idx_b = p.index ? p.index + static_cast<size_t>(b) * p.T_y : nullptr;
  const uint2* bits_b = p.bits + static_cast<size_t>(b) * p.G * p.TXP;
  uint32_t* stage = sstage + static_cast<size_t>(warp) * p.TXS;

  // ---- groups below the top one: as soon as a group's words have arrived, walk it from every entry column ----
  for (int g = warp; g < p.G; g += nw) {
    const uint2* row = bits_b + static_cast<size_t>(g) * p.TXP;
    bool have = false;
    for (;;) {
      poll_lens();
      if (known && (t_y <= 0 || g > ((t_y - 1) >> 5) - kBallotGroups)) break;  // one of the top groups (handled below), or beyond
      // wait for the group's words: an element is valid once it carries this call's tag.  Poll ONE element per
      // forward warp (the last column it owns, or the last valid column) with back-off -- 1000 warps re-reading
      // whole rows would saturate the L2 -- then read the row and check every tag.
      const int nfw = (t_x + p.cols_per_warp - 1) / p.cols_per_warp;  // forward warps that write the words walked here
      // columns needed: all of them -- or, with the lengths known, the reachable entry columns and the 32 to their left
      // (see e_lo below).  A window of at most 256 columns (the groups next to the top: the ones the end of the call
      // waits for) is read at once, without the one-element poll in front: one L2 round trip instead of two.
      const int x_lo = known ? max(0, t_x - 1 - (t_y - 32 * (g + 1)) - 32) & ~31 : 0;
      const bool direct = known && t_x - x_lo <= 256;
      bool ok = true;
      if (!direct && lane < nfw) ok = load_tagged(row + min((lane + 1) * p.cols_per_warp, t_x) - 1).y == tag;
      if (__all_sync(0xffffffffu, ok)) {
        // eight loads per lane in flight at once (one L2 round trip per 256 columns: as a plain loop the loads went out
        // one after the other, six round trips at c2 -- ~1 us of every call's tail after the last frame)
        for (int xb = x_lo; xb < t_x; xb += 256) {
          uint2 el[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int x = xb + 32 * k + lane;
            el[k] = x < t_x ? load_tagged(row + x) : make_uint2(0u, tag);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int x = xb + 32 * k + lane;
            if (x < t_x) stage[x] = el[k].x;
            ok = ok && el[k].y == tag;
          }
        }
        if (__all_sync(0xffffffffu, ok)) {
          have = true;
          break;
        }
      }
      if (__any_sync(0xffffffffu, expired())) break;
      __nanosleep(direct ? 40 : 100);
    }
    if (!have) break;
    if (lane == 0) tl_max(p.tl, 13);  // (debug timeline) last time a tabulated group's words were complete
    __syncwarp();
    // Once the lengths are known, only the entry columns the path can reach matter: it enters group g at most
    // t_y - 32(g+1) columns to the left of t_x-1.  For the groups next to the top that is a handful of columns -- one or
    // two chains per lane instead of six at c2 -- and it is their tables the end of the call waits for.
    const int e_lo = known ? max(0, t_x - 1 - (t_y - 32 * (g + 1))) : 0;
    const int qp = (t_x - e_lo + 31) / 32;
    uint32_t* dec = p.dec16 ? nullptr : sdec + static_cast<size_t>(g) * p.TXS;
    uint16_t* ex = sexit + static_cast<size_t>(g) * p.TXS;
    if (qp <= 1) walk_group<1>(stage, dec, ex, t_x, lane, e_lo);
    else if (qp <= 2) walk_group<2>(stage, dec, ex, t_x, lane, e_lo);
    else if (qp <= 4) walk_group<4>(stage, dec, ex, t_x, lane, e_lo);
    else if (qp <= 6) walk_group<6>(stage, dec, ex, t_x, lane, e_lo);
    else walk_group<8>(stage, dec, ex, t_x, lane, e_lo);
    __syncwarp();
    if (lane == 0) tl_max(p.tl, 11);  // ... and its table finished
  }
  while (!known) {  // (lazy form: every group that exists was walked before the lengths arrived)
    poll_lens();
    if (known || __any_sync(0xffffffffu, expired())) break;
    __nanosleep(100);
  }
  if (!known) {  // gave up (`dead` is set): fall through to the barrier, after which nothing leaves this kernel
    t_y = 1;
    t_x = 1;
  } else if (t_y <= 0) {  // invalid lengths: the path stays all-zero
    if (idx_b)
      for (int y = tid; y < p.T_y; y += blockDim.x) idx_b[y] = -1;
    finish();
    return;
  }
  const int g_top = (t_y - 1) >> 5;
  if (warp == g_top % nw) {
    // The top kBallotGroups groups, one after the other from the known entry (t_y-1, t_x-1): a 32-frame walk can visit
    // at most 32 columns; lane l holds the decision word of column entry-l, 32 ballots transpose them into one mask per
    // frame, and the walk itself is then pure register arithmetic (no dependent shared-memory loads).  These are the
    // groups whose words arrive last: tabulating them from every entry column, as the groups below are, ended 2.5 us after
    // the top group's words had arrived -- on every call's critical path (tools/timeline_gap.py, profiles/r02as_tail.txt).
    int entry = t_x - 1;
    auto poll_cols = [&](int g, int e) {  // lane l: the word of column e-l of group g (0 left of column 0), once it is tagged
      const uint2* row = bits_b + static_cast<size_t>(g) * p.TXP;
      const int col = e - lane;
      uint32_t w = 0u;
      for (;;) {
        bool ok = true;
        if (col >= 0) {
          const uint2 el = load_tagged(row + col);
          w = el.x;
          ok = el.y == tag;
        }
        if (__all_sync(0xffffffffu, ok)) break;
        if (__any_sync(0xffffffffu, expired())) break;
        __nanosleep(50);
      }
      return w;
    };
    uint32_t w = poll_cols(g_top, entry);
    if (lane == 0) tl_max(p.tl, 8);  // (debug timeline) the top group's words have arrived
    for (int g = g_top; g > g_top - kBallotGroups && g >= 0; --g) {
      // The next group's entry lies within 32 columns of this one's: its 64 candidate words are requested NOW, so the
      // L2 round trip overlaps this group's walk, and are shuffled into place once the exit is known.
      const bool more = g - 1 > g_top - kBallotGroups && g - 1 >= 0;
      uint2 ca = make_uint2(0u, tag), cb = make_uint2(0u, tag);
      if (more) {
        const uint2* nrow = bits_b + static_cast<size_t>(g - 1) * p.TXP;
        if (entry - lane >= 0) ca = load_tagged(nrow + entry - lane);
        if (entry - 32 - lane >= 0) cb = load_tagged(nrow + entry - 32 - lane);
      }
      const int rt = g == g_top ? (t_y - 1) & 31 : 31;
      // (four steps per trip, not 32 unrolled: this runs once or twice per CTA, and code that runs once is fetched
      // cold at ~16 cycles per instruction -- the unrolled form took ~1 us per group; and as few instructions per step
      // as possible: a lone warp's dependent scalar code issues one instruction per ~5 cycles)
      uint32_t decw = 0u, posbit = 1u;  // posbit = 1 << (columns stepped so far)
#pragma unroll 4
      for (int r = rt; r >= 0; --r) {
        const uint32_t m = __ballot_sync(0xffffffffu, (w & (0x80000000u >> r)) != 0u);  // bit l: decision of column entry-l
        if (m & posbit) {
          decw |= 1u << r;
          posbit <<= 1;
        }
      }
      const uint32_t pos = static_cast<uint32_t>(__popc(decw));
      if (lane == 0) {
        sentry[g] = entry;
        if (!p.dec16) sdec[static_cast<size_t>(g) * p.TXS + entry] = decw;
        if (g == g_top) smisc[3] = static_cast<int>(decw);
      }
      entry -= static_cast<int>(pos);
      if (more) {
        const int k = static_cast<int>(pos) + lane;  // lane l now needs candidate column number pos+l
        const uint32_t wa = __shfl_sync(0xffffffffu, ca.x, k & 31), wb = __shfl_sync(0xffffffffu, cb.x, k & 31);
        const uint32_t ta = __shfl_sync(0xffffffffu, ca.y, k & 31), tb = __shfl_sync(0xffffffffu, cb.y, k & 31);
        w = k < 32 ? wa : wb;
        if (!__all_sync(0xffffffffu, (k < 32 ? ta : tb) == tag)) w = poll_cols(g - 1, entry);  // (not there yet)
      }
    }
    if (lane == 0) smisc[2] = entry;  // entry of the first tabulated group
    if (lane == 0) tl_max(p.tl, 12);
  }
  __syncthreads();  // all tables are in shared memory
  // Whatever follows in the stream (the next call's forward kernel, when that is launched programmatically:
  // mas_set_tuning pdl = 2) may be scheduled from here on: only the chain over the groups and the ones are left, about
  // as long as a launch takes -- and the forward CTAs of THIS call are gone, so the early CTAs take nobody's SM.  (At the
  // top of the kernel the trigger let the next call's forward CTAs become resident one by one during this call.)
  ptx::pdl_launch_dependents();
  if (tid == 0) tl_max(p.tl, 9);
  if (*dead) {  // a poll gave up: nothing derived from incomplete words leaves this kernel
    if (idx_b)
      for (int y = tid; y < p.T_y; y += blockDim.x) idx_b[y] = -1;
    finish();
    return;
  }
  if (warp == 0) {
    // chain over the groups below the top one: entry - popc(decisions)
    int cur = smisc[2];
    for (int g = g_top - kBallotGroups; g >= 0; --g) {
      if (lane == 0) sentry[g] = cur;
      cur = p.dec16 ? static_cast<int>(sexit[static_cast<size_t>(g) * p.TXS + cur])
                    : cur - __popc(sdec[static_cast<size_t>(g) * p.TXS + cur]);
    }
  } else if (warp == 1 && p.path) {
    // the zero-fill must be complete before the ones are dropped
    if (lane == 0) {
      const uint32_t* word = p.fill_flag ? p.fill_flag + static_cast<size_t>(b) * p.fill_stride
                                         : reinterpret_cast<const uint32_t*>(p.fill_counters + 1);
      const uint32_t need = p.fill_flag ? 1u : static_cast<uint32_t>(p.nchunks);
      while (ptx::ld_acquire_gpu_u32(word) < need) {
        if (expired()) break;
        __nanosleep(100);
      }
    }
  }
  __syncthreads();
  if (*dead) {  // (the zero-fill never completed)
    if (idx_b)
      for (int y = tid; y < p.T_y; y += blockDim.x) idx_b[y] = -1;
    finish();
    return;
  }
  if (tid == 0) tl_max(p.tl, 10);  // chain over the groups done, zero-fill seen complete
  // per-frame index: entry of the frame's group minus the steps taken above the frame
  unsigned char* path_b = p.path ? p.path + static_cast<size_t>(b) * p.T_y * p.T_x * p.es : nullptr;
  if (!p.dec16) {
    for (int y = tid; y < p.T_y; y += blockDim.x) {
      int v = -1;
      if (y < t_y) {
        const int g = y >> 5, r = y & 31;
        const int en = sentry[g];
        const uint32_t dec = sdec[static_cast<size_t>(g) * p.TXS + en];
        v = en - __popc(static_cast<uint32_t>(static_cast<unsigned long long>(dec) >> (r + 1)));
        if (path_b) store_one(path_b + (static_cast<size_t>(y) * p.T_x + v) * p.es, p.es, p.one);
      }
      if (idx_b) idx_b[y] = v;
    }
  } else {
    // 16-bit tables keep only the exits: one thread per group re-walks it from its real entry, reading the
    // group's words from the scratch (L2-resident)
    if (idx_b)
      for (int y = t_y + tid; y < p.T_y; y += blockDim.x) idx_b[y] = -1;
    // A group's walk starts at its entry and moves left by at most one column per frame: the 32 words at and to the
    // left of the entry are all it can read.  The warps fetch them into shared memory first, eight groups' loads in
    // flight per warp -- a thread re-walking its group straight from the scratch paid 32 dependent L2 round trips
    // (c4: 9.7 us after the chain over the groups, profiles/r02_final_bench_c4.json).
    uint32_t* swin = reinterpret_cast<uint32_t*>(smisc + 16);  // [G][32] (dec16 == 2)
    if (p.dec16 == 2) {
      for (int g0 = warp; g0 < g_top; g0 += 8 * nw) {
        uint32_t w8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int g = g0 + i * nw;
          w8[i] = 0u;
          if (g < g_top) {
            const int c = sentry[g] - lane;
            if (c >= 0) w8[i] = load_tagged(bits_b + static_cast<size_t>(g) * p.TXP + c).x;
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int g = g0 + i * nw;
          if (g < g_top) swin[g * 32 + lane] = w8[i];
        }
      }
      __syncthreads();
      if (tid == 0) tl_max(p.tl, 14);
    }
    if (p.dec16 == 2) {
      // one thread per group walks it in shared memory and leaves the group's 32 decisions as one word (in the first
      // slot of its own window); the frames are then written as with 32-bit tables, by all threads
      for (int g = tid; g < g_top; g += blockDim.x) {
        const uint32_t* win = swin + g * 32;
        uint32_t m = 0u;
        int off = 0;
#pragma unroll 8
        for (int r = 31; r >= 0; --r) {
          const uint32_t bit = (win[off] >> (31 - r)) & 1u;
          m |= bit << r;
          off += static_cast<int>(bit);
        }
        swin[g * 32] = m;
      }
      __syncthreads();
      if (tid == 0) tl_max(p.tl, 15);
      const uint32_t topw = static_cast<uint32_t>(smisc[3]);
      for (int y = tid; y < t_y; y += blockDim.x) {
        const int g = y >> 5, r = y & 31;
        const uint32_t dec = g == g_top ? topw : swin[g * 32];
        const int v = sentry[g] - __popc(static_cast<uint32_t>(static_cast<unsigned long long>(dec) >> (r + 1)));
        if (path_b) store_one(path_b + (static_cast<size_t>(y) * p.T_x + v) * p.es, p.es, p.one);
        if (idx_b) idx_b[y] = v;
      }
    } else {
      for (int g = tid; g <= g_top; g += blockDim.x) {
        const uint2* row = bits_b + static_cast<size_t>(g) * p.TXP;
        int cur = sentry[g];
        const uint32_t topw = static_cast<uint32_t>(smisc[3]);
        for (int r = (g == g_top) ? ((t_y - 1) & 31) : 31; r >= 0; --r) {
          const int y = (g << 5) + r;
          if (path_b) store_one(path_b + (static_cast<size_t>(y) * p.T_x + cur) * p.es, p.es, p.one);
          if (idx_b) idx_b[y] = cur;
          const uint32_t bit = (g == g_top) ? (topw >> r) & 1u : (load_tagged(row + cur).x >> (31 - r)) & 1u;
          cur -= static_cast<int>(bit);
        }
      }
    }
  }
  if (tid == 0) tl_max(p.tl, 4);
  finish();
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int g_tune_K = 0, g_tune_R = 0, g_tune_S = 0, g_tune_pdl = -1, g_tune_fused = -1, g_tune_H = 0;
static int g_tune_wf = -1, g_tune_ring = 0, g_tune_wfS = 0, g_tune_wfK = 0;  // wavefront forward kernel (mas_set_tuning3)
static int g_fill_div = 4;  // streaming mode: fill CTAs = g_fill_div/4 x SM count (tunable through MAS_FILL_DIV)
static int g_debug_kernels = 7;  // bit0 forward, bit1 backtrack, bit2 write-out (benchmark isolation only); bit3: skip the
                                 // wavefront forward kernel but keep the streaming backtrack (exercises its watchdog in the tests)
static unsigned long long* g_timeline = nullptr;
static unsigned long long* g_trace = nullptr;

struct Layout {
  int G, TXP_max;
  size_t off_status, off_lens, off_lenstag, off_index, off_bits, total;
};

// Scratch layout is independent of the tuning (bits rows are padded for the widest K).
static Layout scratch_layout(int B, int T_y, int T_x) {
  Layout L;
  L.G = (T_y + 31) / 32;
  // row stride of the decision words is W*32*K; take the largest over the supported K (1,2,3,4,6,8)
  L.TXP_max = ((T_x + 255) / 256) * 256;
  if (((T_x + 191) / 192) * 192 > L.TXP_max) L.TXP_max = ((T_x + 191) / 192) * 192;
  if (((T_x + 95) / 96) * 96 > L.TXP_max) L.TXP_max = ((T_x + 95) / 96) * 96;
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  L.off_status = 0;
  L.off_lens = 256;
  L.off_lenstag = up(L.off_lens + static_cast<size_t>(B) * 2 * 4);
  L.off_index = up(L.off_lenstag + static_cast<size_t>(B) * 8);
  L.off_bits = up(L.off_index + static_cast<size_t>(B) * T_y * 4);
  L.total = up(L.off_bits + static_cast<size_t>(B) * L.G * L.TXP_max * 8);  // streaming mode: {word, tag} pairs
  return L;
}

struct FwdConfig {
  int K, W, H, R, S, BR, fused;
  uint32_t slot_bytes;
  FwdSmem sm;
};

static FwdSmem fwd_smem_layout(int W, int S, int BR, uint32_t slot_bytes, int G, int TXP, bool fused) {
  auto up = [](uint32_t v) { return (v + 127u) & ~127u; };
  const int wb = W - 1 > 1 ? W - 1 : 1;
  FwdSmem m{};
  m.ring = 0;
  m.bnd = up(static_cast<uint32_t>(S) * slot_bytes);
  m.bars = up(m.bnd + static_cast<uint32_t>(W + 1) * BR * 4u);
  m.red = up(m.bars + (2u * S + static_cast<uint32_t>(wb) * S + (fused ? G : 0)) * 8u);
  m.sbits = up(m.red + 64u * 8u + 16u);
  if (fused) {
    m.sexit = up(m.sbits + static_cast<uint32_t>(G) * TXP * 4u);
    m.sentry = up(m.sexit + static_cast<uint32_t>(G) * TXP * 2u);
    m.sidx = up(m.sentry + static_cast<uint32_t>(G) * 4u);
    m.total = up(m.sidx + static_cast<uint32_t>(G) * 32u * 2u);
  } else {
    m.sexit = m.sentry = m.sidx = m.sbits;
    m.total = m.sbits;
  }
  return m;
}

// want_fused: -1 fused if it fits else unfused, 0 unfused, 1 fused or fail
static bool pick_fwd_config(int T_y, int T_x, int want_fused, FwdConfig* cfg) {
  const uint32_t budget = 200 * 1024;
  int K = g_tune_K;
  // keep the block small: <= 7 warps besides the producer get the full register budget
  // measured (tools/sweep.py): three DP warps of 64 columns win at T_x = 192 (39.5 vs 41.4 us), but a fourth
  // warp costs more than wider lanes do: T_x = 256 runs 58 us with K = 4 against 78 us with K = 2
  if (K == 0) K = T_x <= 32 ? 1 : (T_x <= 192 ? 2 : (T_x <= 896 ? 4 : 8));
  int W = (T_x + 32 * K - 1) / (32 * K);
  while (W > 27 && K < 8) {
    K = (K == 3 || K == 6) ? 8 : K * 2;
    W = (T_x + 32 * K - 1) / (32 * K);
  }
  if (W > 27) return false;
  if (g_tune_R && g_tune_R != 8 && g_tune_R != 16 && g_tune_R != 32) return false;
  const int G = (T_y + 31) / 32;
  const int TXP = W * 32 * K;
  // first choice: fused (bits + exit tables in shared memory, >= 3 ring stages); else unfused
  for (int fused = (want_fused == 0 ? 0 : 1); fused >= (want_fused == 1 ? 1 : 0); --fused) {
    int H = 0;
    if (fused) {
      H = g_tune_H ? g_tune_H : 4;
      if (W <= 7 && W + H > 7) H = 7 - W;  // stay within 256 threads when possible ...
      if (H < 1) H = g_tune_H ? g_tune_H : 2;  // ... else take the 1024-thread (64-register) variant
    }
    for (int R = g_tune_R ? g_tune_R : 32; R >= 8; R >>= 1) {
      const uint32_t slot = (static_cast<uint32_t>(R) * T_x * 4u + 16u + 15u) & ~15u;
      for (int S = g_tune_S ? g_tune_S : (fused ? 6 : 8); S >= 2; --S) {
        int BR = 8;
        while (BR < (S + 1) * R) BR <<= 1;
        if (static_cast<uint64_t>(S) * slot + (fused ? static_cast<uint64_t>(G) * TXP * 6 : 0) > budget) continue;
        const FwdSmem m = fwd_smem_layout(W, S, BR, slot, G, TXP, fused != 0);
        if (m.total <= budget) {
          if (S < 3 && R > 8 && !g_tune_R) break;  // prefer more, smaller stages
          if (fused && S < 3) break;
          *cfg = FwdConfig{K, W, H, R, S, BR, fused, slot, m};
          return true;
        }
      }
      if (g_tune_R) break;
    }
  }
  return false;
}

static cudaError_t launch_fwd_dispatch(int K, bool vec, const FwdParams& p, int R, cudaStream_t st) {
  switch (K) {
    case 1: return launch_fwd_k1(vec, p, R, st);
    case 2: return launch_fwd_k2(vec, p, R, st);
    case 3: return launch_fwd_k3(vec, p, R, st);
    case 6: return launch_fwd_k6(vec, p, R, st);
    case 4: return launch_fwd_k4(vec, p, R, st);
    case 8: return launch_fwd_k8(vec, p, R, st);
    default: return cudaErrorInvalidValue;
  }
}

constexpr int kDp2ClusterFromWarps = 5;  // DP warps (64 columns each) from which an utterance is split over two CTAs
                                         // (measured: c3's four warps 42.7 us on one CTA, 49.8 us split; c4's eight 188 vs 129 us)
constexpr int kDp2Default = 33;  // mas_set_tuning3 `wavefront` value used when it is -1: mas_dp2 with the warmer
// the second-generation wavefront kernel covers K = 2 (every automatic choice) unless the first one is asked for
static bool dp2_selected(int K) { return K == 2 && g_tune_wf != 1 && ((g_tune_wf >= 16 ? g_tune_wf : kDp2Default) & ~31) == 32; }
static int dp2_flags() { return (g_tune_wf >= 16 ? g_tune_wf : kDp2Default) & 31; }  // 1 warmer, 4 dummy mask walk, 8 never a cluster, 16 always

struct DpConfig {
  int K, W, S, BR, linear, skew;  // W: DP warps per CTA
  DpSmem sm;
  int CL;  // CTAs per utterance (mas_dp2 only): the text's columns split over a cluster
};

// Backtrack warps per CTA when the source is streamed: small batches leave SMs to spare (a CTA of 16 warps per
// utterance, as in the ordinary mode); large ones pack four-warp CTAs onto a few SMs (fused_backtrack_sms).
static int fused_bt_warps(int B) { return B <= 24 ? 16 : 4; }
constexpr uint32_t kSmemMax = 227 * 1024;  // per CTA on sm_100
constexpr uint32_t kSmemSM = 228 * 1024;   // per SM

static DpSmem dp_smem_layout(int K, int W, int nphys, int S, int BR) {
  auto up = [](uint32_t v) { return (v + 127u) & ~127u; };
  DpSmem m{};
  m.ring = 0;
  m.bnd = up(static_cast<uint32_t>(W) * nphys * kRows * 32u * K * 4u);
  m.bars = up(m.bnd + static_cast<uint32_t>(W + 1) * BR * 4u);
  m.prog = up(m.bars + static_cast<uint32_t>(W) * S * 8u);
  m.red = up(m.prog + static_cast<uint32_t>(W) * 4u);
  m.total = up(m.red + 64u * 8u + 16u);
  return m;
}

// A superstep of a warp reads frames 32s-31*D .. 32s+31 (D = skew between lanes): Q+1 = ceil(31D/32)+1 chunks
// are live and at least one more must be in flight, so the linear ring needs S >= Q+2 slots plus the mirror;
// the select ring (D = 1, no mirror) needs 3.  D = 3 hides the SHFL latency completely, D = 1 not at all.
static bool pick_dp_config_cl(int T_y, int T_x, DpConfig* cfg, uint32_t budget, int CL) {
  int K = g_tune_wfK;
  if (K == 0) K = 2;  // K = 2 keeps the per-step dependency chain short and fills one scheduler per 64 columns
  if (K != 1 && K != 2 && K != 4) return false;
  int W = (T_x + 32 * K - 1) / (32 * K);
  while (W > 8 && K < 4) {
    K *= 2;
    W = (T_x + 32 * K - 1) / (32 * K);
  }
  if (W > 8) return false;
  if (CL > 1) {
    if (!dp2_selected(K) || W < CL) return false;
    W = (W + CL - 1) / CL;  // DP warps per CTA
  }
  const int BR = 256;
  if (budget < 48 * 1024) return false;
  const uint32_t slotset = static_cast<uint32_t>(W) * kRows * 32u * K * 4u;  // one chunk of every warp
  const int nphys = static_cast<int>((budget - 12 * 1024) / slotset);
  int linear = 1, skew = 0;
  if (g_tune_ring >= 1 && g_tune_ring <= 3) skew = g_tune_ring;
  else if (g_tune_ring == 4) { linear = 0; skew = 1; }
  // Measured back to back (tools/sweep_wf.py --ab, one graph of 8 calls, alternating): three DP warps (c2,
  // T_x = 192) 39.7 / 38.1 / 44.2 us per call at skew 1 / 2 / 3 with variable lengths (40.3 / 39.7 / 46.3
  // full-length); four DP warps (c3, T_x = 256) 53.8 / 55.2 / 69.7 us.  A larger skew takes more of the
  // neighbour's SHFL latency off the chain but adds a superstep of lag per warp hop.
  // (the second-generation kernel, mas_dp2.cuh: 30.4 / 31.0 us at skew 1 / 2 on c2, equal on c3 -- skew 1)
  else if (nphys >= 5 && W <= 3 && !dp2_selected(K)) skew = 2;
  else if (nphys >= 4) skew = 1;
  else { linear = 0; skew = 1; }
  if (CL > 1 && (!linear || skew > 2)) return false;
  const int Q = (31 * skew + 31) / 32;
  const int smin = linear ? Q + 2 : 3;
  int S = linear ? nphys - 1 : nphys;
  if (S > 8) S = 8;
  if (g_tune_wfS) S = g_tune_wfS;
  if (S < smin) return false;
  const DpSmem m = dp_smem_layout(K, W, linear ? S + 1 : S, S, BR);
  if (m.total > kSmemMax) return false;
  (void)T_y;
  *cfg = DpConfig{K, W, S, BR, linear, skew, m, CL};
  return true;
}

// Clusters (mas_dp2 only): automatic when the text needs four or more DP warps -- two CTAs of W/2 warps keep one DP warp
// per scheduler and the linear ring (c4: eight warps on one SM with the select ring otherwise).
static bool pick_dp_config(int T_y, int T_x, DpConfig* cfg, uint32_t budget = 208 * 1024, bool allow_cluster = true) {
  const int Wt = (T_x + 63) / 64;
  const int fl = dp2_flags();
  const bool want2 = allow_cluster && !(fl & 8) && ((fl & 16) ? Wt >= 2 : Wt >= kDp2ClusterFromWarps);
  if (want2 && pick_dp_config_cl(T_y, T_x, cfg, budget, 2)) return true;
  return pick_dp_config_cl(T_y, T_x, cfg, budget, 1);
}

static cudaError_t launch_dp_dispatch(int K, const CUtensorMap& tmap, const DpParams& p, int skew, bool linear, cudaStream_t st) {
  switch (K) {
    case 1: return launch_dp_k1(tmap, p, skew, linear, st);
    case 2: return launch_dp_k2(tmap, p, skew, linear, st);
    case 4: return launch_dp_k4(tmap, p, skew, linear, st);
    default: return cudaErrorInvalidValue;
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// [B*T_y][T_x] fp32 tensor, box = [R frames][32*K columns]; out-of-range cells read as zero
static bool make_tensor_map(CUtensorMap* tm, const float* nc, long long rows, int T_x, int R, int cols) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(T_x), static_cast<cuuint64_t>(rows)};
  const cuuint64_t gstr[1] = {static_cast<cuuint64_t>(T_x) * 4u};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(cols), static_cast<cuuint32_t>(R)};
  const cuuint32_t estr[2] = {1u, 1u};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(nc), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
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
  cfg.numAttrs = g_tune_pdl != 0 ? 1 : 0;
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

// where a CUDA call of the last failing entry point failed (diagnostics: mas_last_error_site)
static thread_local int g_fail_line = 0;
static int fail_at(cudaError_t e, int line) {
  g_fail_line = line;
  cudaGetLastError();  // a failed launch must not leak into the caller's next runtime call
  return static_cast<int>(e);
}
int last_fail_line() { return g_fail_line; }

// SMs to set aside for the streaming backtrack CTAs of a streamed (fused) call: they may share an SM neither with a
// forward CTA nor with a contraction CTA, and all B of them should be resident while the DP runs.
int fused_backtrack_sms(int B, int T_y, int T_x) {
  const int TXS = T_x | 1, G = (T_y + 31) / 32;
  const int warps = G + 1 < fused_bt_warps(B) ? (G + 1 < 2 ? 2 : G + 1) : fused_bt_warps(B);
  size_t smem = (static_cast<size_t>(G) * TXS + static_cast<size_t>(warps) * TXS + G + 16) * 4;
  if (smem > 200 * 1024) smem = (static_cast<size_t>(G) * TXS / 2 + 1 + static_cast<size_t>(warps) * TXS + G + 16) * 4;
  int per_sm = static_cast<int>((kSmemSM - 1024) / (smem + 1024));  // 1 KB of system shared memory per resident CTA
  const int by_threads = 2048 / (32 * warps), by_regs = 65536 / (64 * 32 * warps);
  per_sm = per_sm < by_threads ? per_sm : by_threads;
  per_sm = per_sm < by_regs ? per_sm : by_regs;
  if (per_sm < 1) per_sm = 1;
  return (B + per_sm - 1) / per_sm;
}

int num_sms() {
  static int sms[kMaxDevices] = {};
  const int dev = current_device();
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : 148;
  }
  return sms[dev];
}

int maximum_path(const float* neg_cent, const int32_t* t_ys, const int32_t* t_xs, const void* mask, int mask_dtype,
                 int64_t msb, int64_t msy, int64_t msx, void* path_out, int path_dtype, int32_t* index_out,
                 void* scratch, size_t scratch_bytes, int B, int T_y, int T_x, cudaStream_t st, const FusedSrc* fused,
                 bool probe_only) {
  if (B <= 0 || T_y <= 0 || T_x <= 0 || T_x > 2048 || T_y > (1 << 20)) return MAS_E_BAD_SHAPE;
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
  static const int fill_div_env = [] {
    const char* fd = getenv("MAS_FILL_DIV");  // benchmark hook
    return fd && atoi(fd) > 0 ? atoi(fd) : 0;
  }();
  if (fill_div_env) g_fill_div = fill_div_env;
  const int g_num_sms = num_sms();
  int32_t* mirror = status_mirror();
  // Backtrack mode (mas_set_tuning2): 2 = streaming kernel on the idle SMs while the forward kernel runs
  // (per-group tables of every group + one staging row per warp must fit in its shared memory), 1 = fused
  // into the forward kernel (helper warps), 0 = separate kernel after the forward kernel.
  const int TXS = T_x | 1;  // odd shared-memory row stride
  const int G_ = (T_y + 31) / 32;
  int bt_warps = G_ + 1 < 16 ? (G_ + 1 < 2 ? 2 : G_ + 1) : 16;
  // streamed source: the backtrack CTAs get a few SMs of their own (fused_backtrack_sms); small CTAs pack densely there,
  // and four warps keep up with the DP (a group's walk takes ~1.5 us, the DP finishes a group per ~1 us)
  if (fused && bt_warps > fused_bt_warps(B)) bt_warps = fused_bt_warps(B);
  size_t bs_smem = (static_cast<size_t>(G_) * TXS + static_cast<size_t>(bt_warps) * TXS + G_ + 16) * 4;
  int dec16 = 0;
  if (bs_smem > 200 * 1024) {  // long utterances: 16-bit exit columns instead of 32-bit decision words
    dec16 = 1;
    bs_smem = (static_cast<size_t>(G_) * TXS / 2 + 1 + static_cast<size_t>(bt_warps) * TXS + G_ + 16) * 4;
    if (bs_smem + static_cast<size_t>(G_) * 128 <= 200 * 1024) {  // room for every group's 32-word walk window
      dec16 = 2;
      bs_smem += static_cast<size_t>(G_) * 128;
    }
  }
  const bool stream_ok = bs_smem <= 200 * 1024 && (g_debug_kernels & 7) == 7;
  int mode = g_tune_fused;
  if ((mode == 2 || mode == 3) && !stream_ok) return MAS_E_UNSUPPORTED;
  // mode 3: streaming backtrack behind the WAVEFRONT forward kernel (mas_dp.cuh)
  DpConfig dc{};
  // (Streamed source: a backtrack CTA sharing an SM with a forward CTA was measured to stretch the DP from 70 to 100 us
  // at c2 -- its sixteen, even eight, polling and walking warps take the DP warps' issue slots -- so, exactly as in the
  // ordinary mode, the forward CTAs fill their SM's shared memory and the backtrack CTAs run elsewhere.)
  const bool wf_ok = stream_ok && T_x <= 512 && pick_dp_config(T_y, T_x, &dc, 208 * 1024, fused == nullptr);
  if (fused) {
    if (!wf_ok || !dc.linear || dc.skew > 2 || !t_ys || !path_out || B + fused->reserve_ctas > g_num_sms ||
        (fused->pitch & 3) || (reinterpret_cast<uintptr_t>(fused->ring) & 15u))
      return MAS_E_UNSUPPORTED;
    mode = 3;
  }
  if (mode == 3 && !wf_ok) return MAS_E_UNSUPPORTED;
  if (mode < 0 && wf_ok && g_tune_wf != 0) mode = 3;
  FwdConfig fc{};
  if (mode != 3 && !pick_fwd_config(T_y, T_x, mode == 2 ? 0 : mode, &fc)) return MAS_E_UNSUPPORTED;
  // Automatic choice, from the measurements in DESIGN.md section 6: the wavefront kernel with the streaming
  // backtrack when it fits (T_x <= 512); else the fused backtrack when its tables fit next to the ring; else
  // the stage-granular forward kernel with the streaming backtrack (long utterances).
  if (mode < 0 && !fc.fused && stream_ok) mode = 2;
  const bool stream = mode == 2 || mode == 3;
  const bool wavefront = mode == 3;
  const int TXP = wavefront ? dc.W * dc.CL * 32 * dc.K : fc.W * 32 * fc.K;
  const int cols_per_warp = 32 * (wavefront ? dc.K : fc.K);
  const long long fwd_smem = wavefront ? dc.sm.total : fc.sm.total;

  unsigned char* sc = static_cast<unsigned char*>(scratch);
  int32_t* status = reinterpret_cast<int32_t*>(sc + L.off_status);
  int32_t* lens = reinterpret_cast<int32_t*>(sc + L.off_lens);
  uint2* lenstag = reinterpret_cast<uint2*>(sc + L.off_lenstag);
  int32_t* index = index_out ? index_out : reinterpret_cast<int32_t*>(sc + L.off_index);
  uint32_t* bits = reinterpret_cast<uint32_t*>(sc + L.off_bits);

  if (probe_only) return MAS_OK;
  // K1: forward (+ backtrack when fused)
  cudaError_t e = cudaSuccess;
  if (wavefront) {
    DpParams dp{};
    dp.nc = neg_cent; dp.t_ys = t_ys; dp.t_xs = t_xs;
    dp.mask = mask; dp.mask_dtype = mask_dtype; dp.msb = msb; dp.msy = msy; dp.msx = msx;
    dp.lens = lens; dp.status = status; dp.mirror = mirror; dp.bits = bits; dp.lenstag = lenstag; dp.tl = g_timeline; dp.trace = g_trace;
    dp.wo_counters = fused ? nullptr : status + 4;
    // see set_tuning: behind the previous call when asked for, or by default in the measured configuration (streaming
    // backtrack kernel, whose trigger comes late)
    dp.pdl = fused ? 1 : (g_tune_pdl == 2 || (g_tune_pdl < 0 && stream));
    if (fused) {  // launched programmatically behind the contraction kernel and fed by it tile by tile
      dp.tile_flags = fused->tile_flags; dp.tile_need = fused->tile_need; dp.RT = fused->RT; dp.MT = fused->MT;
      dp.fill_out = static_cast<unsigned char*>(path_out);
      dp.fill_bytes = static_cast<long long>(T_y) * T_x * es;
      dp.fill_done = fused->fill_done; dp.fill_stride = fused->fill_stride;
      // pace the fill so that it ends within ~60 % of the expected DP time (T_y x ~30 ns)
      const double iters = static_cast<double>(dp.fill_bytes) / 2048.0;
      const double ns = 0.6 * T_y * 30.0 / (iters > 1 ? iters : 1) - 20.0;
      dp.fill_sleep = ns > 0 ? static_cast<int>(ns) : 0;
    }
    dp.B = B; dp.T_y = T_y; dp.T_x = T_x;
    dp.S = dc.S; dp.W = dc.W; dp.TXP = TXP; dp.G = L.G; dp.BR = dc.BR;
    dp.sm = dc.sm;
    CUtensorMap tmap{};
    dp.use_tma = ((reinterpret_cast<uintptr_t>(neg_cent) & 15u) == 0 && (T_x % 4) == 0) ? 1 : 0;
    if (fused) {
      dp.use_tma = 1;
      if (!make_tensor_map(&tmap, fused->ring, static_cast<long long>(B) * fused->RT * 128, fused->pitch, kRows, 32 * dc.K))
        return MAS_E_UNSUPPORTED;
    } else if (dp.use_tma && !make_tensor_map(&tmap, neg_cent, static_cast<long long>(B) * T_y, T_x, kRows, 32 * dc.K))
      dp.use_tma = 0;
    // Second-generation kernel (mas_dp2.cuh): linear ring, K = 2, skew <= 2, ordinary source.  mas_set_tuning3's
    // `wavefront`: -1 automatic, 1 the first-generation kernel, 32 = mas_dp2, 33 = mas_dp2 with the instruction-cache
    // warmer.
    int dp2_hs = 0;
    if (!fused && dc.linear && dc.skew <= 2 && dp2_selected(dc.K)) {
      dp2_hs = 32;
      dp.warm = dp2_flags() & 7;  // bit 0: instruction-cache warmer; bit 2: experiment (dummy mask walk)
    }
    if (!(g_debug_kernels & 8)) {  // (bit 3: watchdog test hook -- the backtrack kernel then never gets its words)
      e = dp2_hs ? launch_dp2_k2(tmap, dp, dc.skew, dc.CL, st)
                 : launch_dp_dispatch(dc.K, tmap, dp, dc.skew, dc.linear != 0, st);
      if (e != cudaSuccess) return fail_at(e, __LINE__);
      count_launch();
    }
  } else if (g_debug_kernels & 1) {
    FwdParams fp{};
    fp.nc = neg_cent; fp.t_ys = t_ys; fp.t_xs = t_xs;
    fp.mask = mask; fp.mask_dtype = mask_dtype; fp.msb = msb; fp.msy = msy; fp.msx = msx;
    fp.lens = lens; fp.status = status; fp.mirror = mirror; fp.bits = bits; fp.index = index; fp.tl = g_timeline;
    fp.lenstag = stream ? lenstag : nullptr;
    fp.wo_counters = status + 4;
    fp.pdl = g_tune_pdl == 2;
    fp.trace = g_trace;
    fp.B = B; fp.T_y = T_y; fp.T_x = T_x;
    fp.S = fc.S; fp.W = fc.W; fp.H = fc.H; fp.TXP = TXP; fp.G = L.G; fp.BR = fc.BR;
    fp.fused = fc.fused; fp.slot_bytes = fc.slot_bytes; fp.sm = fc.sm;
    const bool vec = (reinterpret_cast<uintptr_t>(neg_cent) & 15u) == 0 && (T_x % 4) == 0 && T_x >= fc.K;
    e = launch_fwd_dispatch(fc.K, vec, fp, fc.R, st);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    count_launch();
  }

  // K2: backtrack (only when the decision bits did not fit in shared memory)
  if (!fc.fused && !stream && (g_debug_kernels & 2)) {
    BtParams bp{};
    bp.bits = bits; bp.lens = lens; bp.index = index; bp.tl = g_timeline;
    bp.T_y = T_y; bp.TXP = TXP; bp.G = L.G;
    bp.TXS = T_x | 1;  // odd stride: neighbouring groups hit different banks in phase 3
    const size_t per_group = static_cast<size_t>(bp.TXS) * 6 + 32 * 4 + 4;
    int GS = static_cast<int>((160 * 1024) / per_group);
    if (GS > L.G) GS = L.G;
    if (GS < 1) return MAS_E_UNSUPPORTED;
    bp.GS = GS;
    const size_t bt_smem = static_cast<size_t>(bp.GS) * bp.TXS * 6 + (bp.GS + 1) * 4 + static_cast<size_t>(bp.GS) * 32 * 4 + 64;
    static std::atomic<uint64_t> bt_attr{0};
    e = ensure_dyn_smem(mas_backtrack_kernel, 200 * 1024, bt_attr);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    long long tasks = static_cast<long long>(bp.GS) * T_x;
    int threads = tasks >= 1024 ? 1024 : static_cast<int>((tasks + 31) / 32 * 32);
    if (threads < 64) threads = 64;
    e = launch_pdl(mas_backtrack_kernel, dim3(B), dim3(threads), bt_smem, st, bp);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    count_launch();
  }

  // K3: dense path
  int nchunks = 0;
  if (path_out && (g_debug_kernels & 4) && !fused) {
    WoParams wp{};
    wp.out = static_cast<unsigned char*>(path_out);
    wp.index = index;
    wp.counters = reinterpret_cast<int32_t*>(sc + L.off_status) + 4;
    wp.tl = g_timeline;
    wp.rows = static_cast<long long>(B) * T_y;
    wp.bytes = wp.rows * T_x * es;
    wp.nchunks = static_cast<int>((wp.bytes + kWoChunk - 1) / kWoChunk);
    nchunks = wp.nchunks;
    wp.T_x = T_x;
    wp.es = es;
    wp.one = one_bits(path_dtype);
    wp.ones = stream ? 0 : 1;
    // few enough CTAs that all of them are resident at once on the SMs the forward kernel leaves idle
    // (they all have to run phase B; a CTA that starts only after the forward kernel adds tail latency)
    // (streaming mode has the whole DP time for the fill: fewer CTAs = less HBM contention for the forward kernel -- but
    // the fill must not outlast the DP.  With the lengths GIVEN the DP of a c2 call ends at ~22 us, the fill of its
    // 50 MB at one CTA per SM at ~25 us: half as many CTAs again take 0.8 us off the call; with the lengths from the
    // mask the DP ends later and the extra CTAs only cost.  profiles/r02bs_fill_cta_sweep.txt)
    const int fill_div = fill_div_env ? g_fill_div : (t_ys != nullptr ? 6 : g_fill_div);
    const int fill_ctas = (stream ? fill_div : 2) * g_num_sms / (stream ? 4 : 1);
    int grid = wp.nchunks < fill_ctas ? wp.nchunks : fill_ctas;
    if (grid < 1) grid = 1;
    // Dynamic shared memory the write-out does not use: just enough that its CTAs cannot be co-resident
    // with a forward CTA; they still pack several per SM on the SMs the forward kernel leaves idle.
    long long wo_smem = 229LL * 1024 - fwd_smem;
    if (wo_smem < 0 || wo_smem > 56 * 1024) wo_smem = 0;  // forward CTA too small to exclude cheaply
    static std::atomic<uint64_t> wo_attr{0};
    e = ensure_dyn_smem(mas_writeout_kernel, 64 * 1024, wo_attr);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    e = launch_pdl(mas_writeout_kernel, dim3(grid), dim3(256), static_cast<size_t>(wo_smem), st, wp);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    count_launch();
  }

  // K2 (streaming variant): backtracks while K1 runs; launched last, so that everything it waits for (K1's
  // words, K3's zero-fill) comes from EARLIER kernels of the stream -- tools that serialise kernels (ncu)
  // then simply find everything ready.
  if (stream) {
    BsParams sp{};
    sp.bits = reinterpret_cast<const uint2*>(bits); sp.lenstag = lenstag; sp.status = status; sp.mirror = mirror;
    sp.index = index_out;  // only when the caller wants it
    sp.path = (path_out && (g_debug_kernels & 4)) ? static_cast<unsigned char*>(path_out) : nullptr;
    sp.fill_counters = reinterpret_cast<int32_t*>(sc + L.off_status) + 4;
    sp.nchunks = nchunks;
    if (fused) {
      sp.fill_flag = fused->fill_done;
      sp.fill_stride = fused->fill_stride;
    }
    sp.T_y = T_y; sp.T_x = T_x; sp.TXP = TXP; sp.G = L.G; sp.TXS = TXS;
    sp.cols_per_warp = cols_per_warp;
    sp.dec16 = dec16;
    sp.lazy = wavefront ? 1 : 0;
    sp.es = es; sp.one = one_bits(path_dtype);
    sp.tl = g_timeline;
    // not co-resident with a forward CTA either (same trick as the write-out kernel)
    long long excl = 229LL * 1024 - fwd_smem;
    if (excl < 0 || excl > 72 * 1024) excl = 0;
    const size_t smem = bs_smem > static_cast<size_t>(excl) ? bs_smem : static_cast<size_t>(excl);
    static std::atomic<uint64_t> bs_attr{0};
    e = ensure_dyn_smem(mas_backtrack_stream_kernel, 200 * 1024, bs_attr);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    e = launch_pdl(mas_backtrack_stream_kernel, dim3(B), dim3(32 * bt_warps), smem, st, sp);
    if (e != cudaSuccess) return fail_at(e, __LINE__);
    count_launch();
  }
  return MAS_OK;
}

size_t maximum_path_scratch_bytes(int B, int T_y, int T_x) {
  if (B <= 0 || T_y <= 0 || T_x <= 0) return 0;
  return scratch_layout(B, T_y, T_x).total;
}

void set_debug_kernels(int mask) { g_debug_kernels = mask; }
void set_timeline(unsigned long long* dev_ptr) { g_timeline = dev_ptr; }
unsigned long long* timeline_ptr() { return g_timeline; }
void set_trace(unsigned long long* dev_ptr) { g_trace = dev_ptr; }

// pdl: 0 = every kernel an ordinary launch; 1 = write-out and backtrack kernels launched programmatically behind the
// forward kernel of their call, the forward kernel itself ordinarily; 2 = the forward kernel programmatically too,
// behind whatever precedes it in the stream; negative (default) = 2 for the wavefront forward kernels followed by the
// streaming backtrack kernel (the configuration measured below), 1 otherwise.  History: 2 with the backtrack kernel's trigger at its TOP let the
// next call's forward CTAs become resident one by one as SMs drained during the current call, and the DP of every
// call but the first took 40-44 us instead of 33 us (tools/timeline_gap.py, four calls in one graph) -- so 1 became the
// default.  With the trigger AFTER the backtrack kernel's tables (only the chain over the groups and the ones are
// left, the call's forward CTAs are gone) and the forward kernel's shared-memory set-up in front of its wait, 2 is
// 0.15-0.3 us per call faster than 1 at c2/c3/c4 (profiles/r02bo_forward_behind_previous_call.txt).
void set_tuning(int K, int R, int S, int pdl) {
  g_tune_K = K;
  g_tune_R = R;
  g_tune_S = S;
  g_tune_pdl = pdl;  // (negative: automatic)
}
void set_tuning3(int wavefront, int ring_mode, int ring_slots, int cols_per_lane) {
  g_tune_wf = wavefront;
  g_tune_ring = ring_mode;
  g_tune_wfS = ring_slots;
  g_tune_wfK = cols_per_lane;
}
void set_tuning2(int fused, int helpers) {
  g_tune_fused = fused;
  g_tune_H = helpers;
}

}  // namespace mas
