// mas_dp2.cuh -- second generation of the wavefront forward DP kernel (monotonic_align/core.pyx:13-28).
//
// Same algorithm, data layout, hand-off protocol and producer warps as mas_dp.cuh (linear ring with mirror
// slot, ordinary -- not streamed -- source); what changed is the shape of the DP warps' code, after a
// per-superstep cycle trace (tools/trace_dp.py, profiles/r02s_trace_dp.txt) showed where a c2 call's 35 us went:
//
//   * a superstep (32 frames) took 1100 cycles of which the 32 unrolled steps were 558; 322 cycles passed
//     between the end of one superstep and the first instruction of the next -- the loop's straight-line body
//     (four variants of ~11 KB each) does not fit the instruction cache next to the scheduler, so every
//     superstep re-fetched it;
//   * the FIRST execution of each variant cost 6500-7400 cycles (instruction fetch from L2): warp 0 ran three
//     of them cold back to back (frames < 0 + diagonal, diagonal, plain) = ~10 us of the call.
//
// Here (1) the "frames < 0" variants are gone: the ring slots that precede chunk 0 are zero-filled, so a lane
// that has not reached frame 0 yet adds 0 to the sentinel row (exactly what the variant's selects did);
// (2) a superstep is ONE basic block: the previous superstep's decision words, this superstep's address
// arithmetic and its 32 steps are straight-line code (the words leave in a predicated store), so the
// bookkeeping issues in the slots the recurrence leaves empty, and everything rare -- waiting for an input
// that was late, the lengths arriving, the end -- sits behind a single test after the block; (3) an otherwise
// idle warp executes the plain variant once while warp 0 is still busy with the diagonal one, so that nobody
// meets it cold; (4) when the lengths come from the mask, every idle warp of the CTA shares the strided walk.
// (Tried and dropped: 16 unrolled steps in a rolled loop of two -- the body ran 1020 instead of 605 cycles.)
#pragma once
#include "mas_dp.cuh"

// How the diagonal x == y (core.pyx:17-18 `v_cur = max_neg_val`, :32 `index == y`) is handled.  1: cells ABOVE the
// diagonal add 0 instead of their neg_cent value (a select on the ring value, off the recurrence's dependency chain), so
// everything above the diagonal stays at exactly the sentinel -- which is then what a diagonal cell finds as v_cur -- and
// the forced decision bit is OR-ed into the group's finished word.  0: the first form -- per step, the running value is
// replaced by the sentinel and the bit forced where i == the lane's diagonal step (two selects on the chain: a
// diagonal superstep's 32 steps took 1350 instead of 780 cycles, and with one warp after the other crossing the
// diagonal the LAST warp's first seven supersteps all ran at that pace: profiles/r02bj_c2_trace.txt).
#ifndef MAS_DIAGC
#define MAS_DIAGC 1
#endif
// (Tried and dropped: lane 31 collecting four frames of the last column and publishing them as one 16-byte store --
// eight hand-off stores per superstep instead of 32: the block's 32 steps took 765 instead of 753 cycles and the code
// between two blocks 300 instead of 262, c2 +1.0 us, c4 +3.5 us; profiles/r02bk_handoff_16_byte_stores.txt.)

namespace mas {

template <int K, int D, int CL>
__global__ void __launch_bounds__(CL > 1 ? 384 : 416, 1) mas_dp2_kernel(const __grid_constant__ CUtensorMap tmap, const DpParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int R = kRows;
  constexpr int Q = (31 * D + 31) / 32;        // chunks (and decision-word groups) a superstep reaches back
  constexpr int LAG = (30 + 31 * D) / 32 + 1;  // supersteps the left neighbour must be ahead
  constexpr uint32_t ROWB = 32u * K * 4u;      // bytes of one ring row: this warp's 32*K columns of one frame
  constexpr uint32_t SLOTB = R * ROWB;
  // CL > 1: an utterance's columns are split over a cluster of CL CTAs (W = DP warps of THIS CTA, global warp index
  // gw0 + dw): the last column crosses the CTA boundary through distributed shared memory, by READS only, and not by a
  // DP warp -- a COURIER warp of the right CTA polls the progress counter of the left CTA's last warp, copies each new
  // block of 32 hand-off values from the left CTA's ring into the local one (ring 0, unused otherwise) and publishes
  // its own counter, so that the right CTA's first DP warp sees an ordinary local neighbour; the left CTA's last warp
  // reads the courier's counter for back-pressure (one remote load per superstep).  History: remote STORES first (a
  // store per step through shared::cluster ~20 cycles each, fence.acq_rel.cluster in front of the progress store
  // 0.7 us per superstep: c3 94 instead of 43 us); then remote LOADS by the DP warp itself, two blocks ahead -- but a
  // remote load takes ~215 cycles and shares its scoreboard with the step's shuffles and ring loads (six scoreboards
  // for everything in flight), so every wait for a shuffle also waited for the remote load: the boundary warp's 32
  // steps took 1237 instead of 917 cycles and the whole chain ran at its pace (profiles/r02bi_c4_trace.txt).  A text
  // of 512 tokens is then four DP warps per SM (one per scheduler, linear ring) instead of eight on one SM with the
  // select ring of the first generation.
  const int rank = CL > 1 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int b = CL > 1 ? blockIdx.x / CL : blockIdx.x;
  const int tid = threadIdx.x;
  const int wid = tid >> 5;
  const bool spread = p.W <= 3;  // warp roles exactly as in mas_dp_kernel
  const int NP = p.W <= 3 ? 2 : 4;
  int dw = spread ? (wid < 3 ? (wid < p.W ? wid : -1) : (wid & 3) == 3 ? p.W + (wid >> 2) : -1) : wid;
  if (CL > 1 && !spread && wid == p.W + NP + 1) dw = -1;  // the extra warp of a cluster launch: courier (rank > 0) or filler
  const int lane = tid & 31;
  const int S = p.S, W = p.W, BR = p.BR;
  const int gw0 = rank * W;  // global index of this CTA's first DP warp
  const int nphys = S + 1;
  // the instruction-cache warmer: an idle warp on the scheduler of the last DP warp (which starts last)
  const bool shadow = (p.warm & 1) != 0 && spread && wid == (W < 3 ? 4 + W : 6);
  const int dwa = shadow ? 0 : dw;  // whose ring / barriers a warp addresses
  const int cour_wid = (CL > 1 && rank > 0) ? (spread ? 9 : W + NP + 1) : -1;
  const bool courier = wid == cour_wid;

  unsigned char* ring_all = smem + p.sm.ring;
  unsigned char* ringw = ring_all + static_cast<size_t>(dwa < 0 ? 0 : dwa) * nphys * SLOTB;
  float* bnd = reinterpret_cast<float*>(smem + p.sm.bnd);              // [W+1][BR]
  uint64_t* full_all = reinterpret_cast<uint64_t*>(smem + p.sm.bars);  // [W][S]
  uint64_t* full = full_all + static_cast<size_t>(dwa < 0 ? 0 : dwa) * S;
  int* prog = reinterpret_cast<int*>(smem + p.sm.prog);                // [W] supersteps completed
  double* red = reinterpret_cast<double*>(smem + p.sm.red);
  int* lens_s = reinterpret_cast<int*>(red + 64);

  if (p.use_tma && tid == 0) ptx::prefetch_tensormap(&tmap);  // (the descriptor's fetch overlaps the barrier set-up)
  const long long rows_total = static_cast<long long>(p.B) * p.T_y;
  auto copy_chunk = [&](int w, int c, unsigned char* dst, uint64_t* bar) {
    if (p.use_tma) {
      if (lane == 0) ptx::tma_load_2d(dst, &tmap, (gw0 + w) * 32 * K, b * p.T_y + c * R, bar);
    } else {
      const uint32_t d0 = ptx::smem_u32(dst) + static_cast<uint32_t>(lane) * K * 4u;
      const int xb = ((gw0 + w) * 32 + lane) * K;
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
  auto issue_chunk = [&](int w, int c, int ls) {
    uint64_t* bar = full_all + static_cast<size_t>(w) * S + ls;
    unsigned char* rw = ring_all + static_cast<size_t>(w) * nphys * SLOTB;
    const bool mirror = ls == 0;
    if (p.use_tma && lane == 0) ptx::mbar_arrive_expect_tx(bar, mirror ? 2u * SLOTB : SLOTB);
    copy_chunk(w, c, rw + static_cast<size_t>(ls) * SLOTB, bar);
    if (mirror) copy_chunk(w, c, rw + static_cast<size_t>(S) * SLOTB, bar);
    if (!p.use_tma) ptx::cp_async_mbar_arrive_noinc(bar);
  };

  // Everything that touches only this CTA's shared memory comes BEFORE the wait: launched programmatically behind the
  // previous call (mas_set_tuning pdl = 2, the default) the CTA is resident while that call's backtrack kernel drops its
  // ones, and has its barriers, rings and counters ready when the wait returns.
  const int nspec = min(2, (p.T_y + R - 1) / R);
  if (dw == W && lane == 0) {
    for (int i = 0; i < W * S; ++i) ptx::mbar_init(&full_all[i], p.use_tma ? 1 : 32);
    ptx::mbar_fence_init();
  }
  // Frames < 0: the Q ring slots that precede chunk 0 (slots S-Q .. S-1 of every warp's ring) read as zero, so
  // a lane whose skew has not brought it to frame 0 yet computes 0 + max(sentinel, sentinel): its row stays at
  // the sentinel without a select per step.  (The producers refill those slots only after superstep Q-1, by
  // the ordinary rule; S >= Q+2 keeps them clear of the two speculative chunks.)
  {
    const int per_w = Q * static_cast<int>(SLOTB / 16u);
    for (int i = tid; i < W * per_w; i += blockDim.x) {
      const int w = i / per_w, o = i - w * per_w;
      reinterpret_cast<uint4*>(ring_all + (static_cast<size_t>(w) * nphys + (S - Q)) * SLOTB)[o] = make_uint4(0u, 0u, 0u, 0u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros before the TMA's later writes
  }
  for (int i = tid; i < (W + 1) * BR; i += blockDim.x) bnd[i] = (i == 0 && rank == 0) ? 0.0f : kNeg;  // (0,0): v_prev = 0 (core.pyx:22-23)
  if (tid <= W) prog[tid] = 0;  // ([W]: the courier's counter)
  volatile int* lens_v = lens_s;  // [0] t_y, [1] t_x, [2] 1 once they are known
  if (tid == 0) {
    lens_s[2] = 0;
    lens_s[3] = 0;
    red[0] = 0.0;
    red[1] = 0.0;
  }
  ptx::pdl_wait();  // global memory is first touched from here on
  if (dw == W) {
    __syncwarp();
    for (int c = 0; c < nspec; ++c)
      for (int w = 0; w < W; ++w) issue_chunk(w, c, c);
  }
  if (blockIdx.x == 0 && tid == 0) {
    p.wo_counters[0] = 0;
    p.wo_counters[1] = 0;
  }
  constexpr uint32_t tag = 1u;
  if (p.lenstag) {
    uint4* z = reinterpret_cast<uint4*>(reinterpret_cast<uint2*>(p.bits) + static_cast<size_t>(b) * p.G * p.TXP);
    const int n16 = p.G * p.TXP / 2;
    for (int i = tid + rank * blockDim.x; i < n16; i += blockDim.x * CL) z[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0 && rank == 0) *reinterpret_cast<unsigned long long*>(p.lenstag + b) = 0ull;
  }
  if (tid == 0) tl_min(p.tl, 0);
#ifdef MAS_TRACE
  if (p.trace && tid == 0) p.trace[8 * 256 * 8 + 2 * b] = globaltimer_ns();
#endif
  __syncthreads();
  // A peer reads this CTA's hand-off ring and progress counters: nobody starts before both CTAs are initialised, and
  // nobody leaves (leave() below) before both are done.
  if (CL > 1) {
    ptx::cluster_arrive();
    ptx::cluster_wait();
  }
  auto leave = [&]() {
    if (CL > 1) {
      ptx::cluster_arrive();
      ptx::cluster_wait();
    }
  };

  // Lengths: one warp publishes them; when they come from the mask (monotonic_align/__init__.py:17-18) the
  // otherwise idle warps of the CTA help with the strided walk over column 0 -- 1024 sectors per utterance, one DRAM
  // round trip when every thread issues a handful of loads instead of two batches of sixteen by one warp.
  const bool lenw = dw == W + NP;
  const bool dummy_walk = (p.warm & 4) != 0 && p.t_ys != nullptr && p.mask != nullptr;  // (experiment: the walk's traffic alone)
  const bool helper = spread && dw < 0 && !shadow && !courier && (p.t_ys == nullptr || dummy_walk);
  if (dw < 0 && !shadow && !helper && !courier) {  // filler warps
    leave();
    return;
  }
  if (lenw || helper) {
    if (lenw && lane == 0) {
      __threadfence();  // tags and counters were cleared before the barrier: visible GPU-wide before the dependents start
      ptx::pdl_launch_dependents();
    }
    int t_y, t_x;
    if (p.t_ys != nullptr && !(dummy_walk && !lenw)) {
      t_y = p.t_ys[b];
      t_x = p.t_xs[b];
    } else {
      int nh = 0, hrank = 0;  // warps that walk the mask, and this warp's place among them
      if (spread) {
        for (int u = 0; u < 12; ++u) {
          const bool sh = (p.warm & 1) != 0 && u == (W < 3 ? 4 + W : 6);
          const bool part = u != cour_wid && (u == 11 || (!sh && ((u < 3 && u >= W) || (u > 3 && (u & 3) != 3))));
          nh += part ? 1 : 0;
          hrank += part && u < wid ? 1 : 0;
        }
      } else {
        nh = 1;
      }
      double sy, sx;
      mask_sums(p.mask, p.mask_dtype, static_cast<int64_t>(b) * p.msb, p.msy, p.T_y, p.msx, p.T_x, hrank * 32 + lane, nh * 32, sy, sx);
      sy = warp_sum(sy);
      sx = warp_sum(sx);
      if (nh > 1) {
        if (lane == 0) {
          atomicAdd(&red[0], sy);
          atomicAdd(&red[1], sx);
          __threadfence_block();
          atomicAdd(&lens_s[3], 1);
        }
        if (!lenw) {
          leave();
          return;
        }
        if (lane == 0)
          while (ptx::ld_volatile_s32(&lens_s[3]) < nh) {
            __nanosleep(200);
          }
        __syncwarp();
        sy = *reinterpret_cast<volatile double*>(&red[0]);
        sx = *reinterpret_cast<volatile double*>(&red[1]);
      }
      t_y = static_cast<int>(sy);
      t_x = static_cast<int>(sx);
    }
    int st = 0;
    if (t_y < 1 || t_x < 1) st |= MAS_STATUS_EMPTY;
    if (t_y > p.T_y || t_x > p.T_x) st |= MAS_STATUS_TOO_LONG;
    if (t_x > t_y) st |= MAS_STATUS_TX_GT_TY;
    if (st) t_y = t_x = 0;
    if (lane == 0) {
      if (rank == 0) {  // (every CTA of a cluster finds the lengths for itself; the first one publishes them)
        if (st) raise_status(p.status, p.mirror, st);
        p.lens[2 * b] = t_y;
        p.lens[2 * b + 1] = t_x;
        if (p.lenstag)
          *reinterpret_cast<unsigned long long*>(p.lenstag + b) = pack_tagged((static_cast<uint32_t>(t_y) << 12) | static_cast<uint32_t>(t_x), tag);
      }
      lens_v[0] = t_y;
      lens_v[1] = t_x;
      __threadfence_block();
      lens_v[2] = 1;
      tl_max(p.tl, 7);
    }
    leave();
    return;
  }
  bool known = false;
  int NS = ((p.T_y - 1) >> 5) + Q + 1;
  int nchunks = (p.T_y + R - 1) / R;
  auto check_lens = [&]() {
    if (!known && lens_v[2] != 0) {
      known = true;
      const int t_y = lens_v[0];
      NS = t_y > 0 ? ((t_y - 1) >> 5) + Q + 1 : 0;
      nchunks = (t_y + R - 1) / R;
    }
  };
  if (CL > 1 && courier) {
    // ---- courier: the left CTA's last column, block by block, into ring 0 of this CTA ----
    // The left warp's superstep s publishes slots 32(s-Q)+R0 .. +31 of its ring and then progress s+1; `n` mirrors that
    // counter: chunk n (0-based) may be copied once the remote counter is past it and the local first warp has left the
    // slots it overwrites (the left warp's own rule for its right neighbour, need_r below).
    constexpr int R0 = 32 * Q - (31 * D - 1);
    const uint32_t lrank = static_cast<uint32_t>(rank > 0 ? rank - 1 : 0);
    const uint32_t rprog = ptx::mapa(ptx::smem_u32(&prog[W - 1]), lrank);
    const uint32_t rbnd = ptx::mapa(ptx::smem_u32(bnd + static_cast<size_t>(W) * BR), lrank);
    int n = 0;
    for (;;) {
      check_lens();
      const int lim = min(ptx::ld_volatile_cluster_s32(rprog), ptx::ld_volatile_s32(&prog[0]) + BR / 32);
      if (lim > n) {
        while (n < lim) {
          const int m = min(lim - n, 4);
          float x[4];
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < m) x[k] = ptx::ld_volatile_cluster_f32(rbnd + 4u * static_cast<uint32_t>((32 * (n + k - Q) + R0 + lane) & (BR - 1)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < m) ptx::st_volatile_f32(bnd + ((32 * (n + k - Q) + R0 + lane) & (BR - 1)), x[k]);
          n += m;
        }
        __syncwarp();
        if (lane == 0) ptx::st_volatile_s32(&prog[W], n);
        __nanosleep(300);  // the next block is most of a superstep (~600 ns) away
        continue;
      }
      if (known && n >= NS) break;
      __nanosleep(150);
    }
    leave();
    return;
  }
  if (dw >= W && !shadow) {
    // ---- producer warps: unchanged from mas_dp_kernel ----
    const int q = dw - W;
    int ci = nspec;
    int cs = nspec % S;
    const bool mine = lane < W && (lane % NP) == q;
    for (;;) {
      check_lens();
      const bool want = mine && ci < nchunks;
      const bool ready = want && ptx::ld_volatile_s32(&prog[lane]) >= ci - S + Q + 1;
      unsigned m = __ballot_sync(0xffffffffu, ready);
      if (known && !__any_sync(0xffffffffu, mine && ci < nchunks)) break;
      if (m == 0u) __nanosleep(64);  // (64..200 ns measure the same in the wide layout, where producers share the DP warps' schedulers)
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
    if (mine)
      for (int c = max(nchunks, ci - S); c < ci; ++c) ptx::mbar_wait(full_all + static_cast<size_t>(lane) * S + (c % S), (c / S) & 1);
    leave();
    return;
  }

  // ---- DP warp (and the warmer, which runs one superstep of the plain variant on whatever the ring holds and
  // touches nothing outside its registers) ----
  const int gw = gw0 + dwa;  // this warp's place among the utterance's DP warps
  const int x0 = (gw * 32 + lane) * K;
  const bool has_left = gw > 0;
  const bool has_right = gw < W * CL - 1 && !shadow;
  // progress of the neighbours and the left edge: this CTA's shared memory -- the first warp of a right CTA has the
  // courier (counter prog[W], ring 0) for a left neighbour; the last warp of a left CTA reads the right CTA's courier
  // counter through a shared::cluster address
  const bool right_remote = CL > 1 && dwa == W - 1 && rank < CL - 1;
  const int* plp = (CL > 1 && dwa == 0 && rank > 0) ? &prog[W] : &prog[dwa > 0 ? dwa - 1 : 0];
  const int* prp = &prog[dwa < W - 1 ? dwa + 1 : 0];
  const uint32_t prc = CL > 1 ? (right_remote ? ptx::mapa(ptx::smem_u32(&prog[W]), rank + 1) : ptx::mapa(ptx::smem_u32(prp), rank)) : 0u;
  auto ld_left = [&]() { return ptx::ld_volatile_s32(plp); };
  auto ld_right = [&]() { return CL > 1 ? ptx::ld_volatile_cluster_s32(prc) : ptx::ld_volatile_s32(prp); };
  const bool lane0 = lane == 0;
  const bool lane31 = lane == 31 && !shadow;
  const uint32_t bnd_in = ptx::smem_u32(bnd + static_cast<size_t>(dwa) * BR);
  float* bnd_out = bnd + static_cast<size_t>(dwa + 1) * BR;
  uint2* bits_b = reinterpret_cast<uint2*>(p.bits) + static_cast<size_t>(b) * p.G * p.TXP + x0;  // {word, tag} pairs
  const int dt = x0 + D * lane;
  const int diag_lo = gw * 32 * K, diag_hi = gw * 32 * K + 31 * (K + D) + K - 1;
  const int hsel = (D * lane) >> 5;
  const int hsh = (D * lane) & 31;

  float v[K];
  uint32_t hist[Q + 1][K];
#pragma unroll
  for (int j = 0; j < K; ++j) {
    v[j] = kNeg;
#pragma unroll
    for (int k = 0; k <= Q; ++k) hist[k][j] = 0u;
  }
  float left[D];
#pragma unroll
  for (int k = 0; k < D; ++k) left[k] = kNeg;

  const uint32_t ring_lane = ptx::smem_u32(ringw) + static_cast<uint32_t>(lane) * K * 4u;
  const int ring_frames = S * R;
  int foff = (ring_frames * 4 - D * lane) % ring_frames;
  int ls = 0;
  uint32_t par = 0u;

  constexpr int NE = 2;  // hand-off blocks of 8 frames held in registers ...
  constexpr int AH = 1;  // ... loaded this many blocks ahead of their use
  float cb[4][K], e[NE][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < K; ++j) cb[k][j] = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int q = 0; q < NE; ++q) e[q][k] = kNeg;
  bool pre = false;

  auto load_e = [&](uint32_t a, float (&e8)[8]) {
    const float4 e0 = ptx::lds_f32x4(a);
    const float4 e1 = ptx::lds_f32x4(a + 16u);
    e8[0] = e0.x; e8[1] = e0.y; e8[2] = e0.z; e8[3] = e0.w;
    e8[4] = e1.x; e8[5] = e1.y; e8[6] = e1.z; e8[7] = e1.w;
  };

  // The decision words of group g = s-Q are complete after superstep s: frames 32g+r of lane l sit in hist[Q-a] (the
  // older part) and hist[Q-a-1], a = D*l/32, shifted by D*l % 32.  One 16-byte store of two {word, tag} pairs.
  auto emit_words = [&](int g, bool on, auto force_tag) {
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
      if (MAS_DIAGC && decltype(force_tag)::value)  // core.pyx:32 `index == y`: frame x of column x steps
        w[jj] |= (g == ((x0 + jj) >> 5)) ? (0x80000000u >> ((x0 + jj) & 31)) : 0u;
    }
    if (x0 == 0) w[0] = 0u;  // core.pyx:32 `index != 0`
    uint2* dst = bits_b + static_cast<long long>(g) * p.TXP;
    static_assert(K == 2, "one 16-byte store per lane");
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %3, 0;\n\t"
        "@q st.global.v2.u64 [%0], {%1, %2};\n\t}" ::"l"(dst),
        "l"(pack_tagged(w[0], 1u)), "l"(pack_tagged(w[1], 1u)), "r"(static_cast<uint32_t>(on))
        : "memory");
  };

  // Blocking start of superstep s (the first one, or one whose inputs were not there yet when the previous
  // superstep probed for them): wait for chunk s, the left neighbour's hand-off values and the right neighbour's
  // progress, then load the first three rows and the first hand-off block.
  auto blocking_start = [&](int s) {
    const uint32_t pcur = ring_lane + static_cast<uint32_t>(foff) * ROWB;
    uint32_t ea = bnd_in + 4u * static_cast<uint32_t>((32 * s) & (BR - 1));
    const int need_r = s - BR / 32 + 1;
    while (s < nchunks && !ptx::mbar_test(&full[ls], par)) check_lens();
    lds_cols<K>(cb[0], pcur);
    lds_cols<K>(cb[1], pcur + ROWB);
    lds_cols<K>(cb[2], pcur + 2u * ROWB);
    int fl = 0;
    if (has_left) {
      for (;;) {
        fl = ld_left();
        if (fl >= min(s + LAG, NS)) break;
        check_lens();  // the neighbour may have stopped at a smaller NS than the one assumed so far
      }
    }
    if (has_right && need_r > 0)
      while (ld_right() < need_r) {
        check_lens();  // (a right neighbour that knew the lengths first stops at NS: nothing left to wait for then)
        if (known && s >= NS) break;
      }
    ea += static_cast<uint32_t>(fl) >> 31;  // (null) dependency: the hand-off loads stay behind the poll
#pragma unroll
    for (int q = 0; q < AH; ++q) load_e(ea + 32u * q, e[q]);
  };

  // One superstep = one basic block: the previous superstep's decision words, this superstep's addresses and
  // its 32 steps are straight-line code, so the bookkeeping fills the issue slots the recurrence leaves empty.
  auto superstep = [&](int s, auto diag_tag) {
    constexpr bool DIAG = decltype(diag_tag)::value;
#ifdef MAS_TRACE
    unsigned long long* tr = (p.trace && b == 0 && lane0 && s < 256 && !shadow) ? p.trace + (static_cast<size_t>(gw) * 256 + s) * 8 : nullptr;
    if (tr) tr[0] = clock64();
#endif
    emit_words(s - 1 - Q, s - 1 >= Q && !shadow, diag_tag);
#pragma unroll
    for (int k = Q; k >= 1; --k)
#pragma unroll
      for (int jj = 0; jj < K; ++jj) hist[k][jj] = hist[k - 1][jj];
    const int ls_n = ls + 1 == S ? 0 : ls + 1;
    const uint32_t par_n = ls + 1 == S ? par ^ 1u : par;
    const int foff_n = foff + R >= ring_frames ? foff + R - ring_frames : foff + R;
    const uint32_t pcur = ring_lane + static_cast<uint32_t>(foff) * ROWB;    // frame 32s - D*lane of this lane
    const uint32_t pcur_n = ring_lane + static_cast<uint32_t>(foff_n) * ROWB;  // ... one superstep later
    const uint32_t ea = bnd_in + 4u * static_cast<uint32_t>((32 * s) & (BR - 1));
    const uint32_t ea_n = bnd_in + 4u * static_cast<uint32_t>((32 * s + 32) & (BR - 1));
    const int need_r = s - BR / 32 + 1;
    constexpr int R0 = 32 * Q - (31 * D - 1);  // lane 31 publishes frame 32s+i-31D into slot 32(s-Q) + R0 + i
    const int bi1 = ((32 * (s - Q)) & (BR - 1)) + R0, bi2 = ((32 * (s - Q + 1)) & (BR - 1)) - (32 - R0);
    float* bo1 = bnd_out + bi1;
    float* bo2 = bnd_out + bi2;
    const int dd = dt - 32 * s;
    bool chunk_n = true;
    int pl = 0x7fffffff, pr = 0x7fffffff;
    uint32_t dep = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      lds_cols<K>(cb[(i + 3) & 3], (i + 3 < 32 ? pcur : pcur_n - 32u * ROWB) + static_cast<uint32_t>(i + 3) * ROWB);
      if ((i & 7) == 0) {
        const int blk = i / 8 + AH;  // (blocks 4, 5 are the next superstep's 0, 1)
        load_e(blk < 4 ? ea + 32u * blk : ea_n + 32u * (blk - 4) + dep, e[blk & (NE - 1)]);
      }
      if (i == 8) {  // non-blocking probes of the next superstep's inputs
        if (s + 1 < nchunks) chunk_n = ptx::mbar_test(&full[ls_n], par_n);
        if (has_left) pl = ld_left();
        if (has_right) pr = ld_right();
        dep = static_cast<uint32_t>(pl) >> 31;  // (null) dependency: the prefetch of the next block 0 stays behind this probe
      }
      float c[K];
#pragma unroll
      for (int jj = 0; jj < K; ++jj) c[jj] = (MAS_DIAGC && DIAG && i < dd + jj) ? 0.0f : cb[i & 3][jj];  // (above the diagonal)
      const float nxt = __shfl_up_sync(0xffffffffu, v[K - 1], 1);  // for step i+D
      const float le = lane0 ? e[(i / 8) & (NE - 1)][i & 7] : left[0];
      if (!MAS_DIAGC && DIAG) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) v[jj] = (i == dd + jj) ? kNeg : v[jj];  // core.pyx:17-18
      }
#pragma unroll
      for (int jj = K - 1; jj >= 1; --jj) {
        const float d = v[jj] - v[jj - 1];                                    // sign bit == (stay < step), core.pyx:32
        hist[0][jj] = __funnelshift_l(__float_as_uint(d), hist[0][jj], 1);
        v[jj] = c[jj] + fmaxf(v[jj - 1], v[jj]);                              // core.pyx:28
      }
      const float d = v[0] - le;
      hist[0][0] = __funnelshift_l(__float_as_uint(d), hist[0][0], 1);
      v[0] = c[0] + fmaxf(le, v[0]);
      if (!MAS_DIAGC && DIAG) {
#pragma unroll
        for (int jj = 0; jj < K; ++jj) hist[0][jj] |= (i == dd + jj) ? 1u : 0u;  // core.pyx:32 `index == y`
      }
      if (lane31) ptx::st_volatile_f32((i < 32 - R0 ? bo1 : bo2) + i, v[K - 1]);
#pragma unroll
      for (int k = 0; k + 1 < D; ++k) left[k] = left[k + 1];
      left[D - 1] = nxt;
    }
#ifdef MAS_TRACE
    if (tr) tr[3] = clock64();
#endif
    __syncwarp();
    // frames <= 32(s+1)-31D-1 of our last column are published; every lane is past chunk s-Q (the producer may refill it)
    if (lane31) ptx::st_volatile_s32(&prog[dwa], s + 1);
    pre = chunk_n && pl >= min(s + 1 + LAG, NS) && pr >= need_r + 1;
    ls = ls_n;
    par = par_n;
    foff = foff_n;
#ifdef MAS_TRACE
    if (tr) tr[7] = clock64();
#endif
  };

  // supersteps in which the diagonal x == y crosses this warp's columns.  MAS_DIAGC: the variant runs from superstep 0
  // (every cell above the diagonal must stay at the sentinel) until the words of the last group that holds a diagonal
  // cell of these columns have left (group 2gw+1, emitted at the start of superstep 2gw+2+Q).
  const int sd0 = MAS_DIAGC ? 0 : diag_lo >> 5;
  const int sd1 = MAS_DIAGC ? max(diag_hi >> 5, ((diag_lo + 32 * K - 1) >> 5) + 1 + Q) : diag_hi >> 5;
  int s = 0;
  if (shadow) {  // one superstep of the plain variant, far from any diagonal, on whatever the ring holds
    s = sd1 + 1;
    known = true;
    NS = s + 1;
    nchunks = 0;
    pre = true;
  } else {
    check_lens();
    while (NS == 0 && !known) check_lens();
  }
  if (s < NS) {
    int lim = 0;  // supersteps below this one start on `pre` alone
    if (!pre) blocking_start(s);
    for (;;) {
      if (s >= sd0 && s <= sd1) superstep(s, std::true_type{});
      else superstep(s, std::false_type{});
      ++s;
      if (s == 1 && gw == 0 && lane0 && !shadow) bnd[0] = kNeg;  // the (0,0) special case is consumed
      // The common case: everything the next superstep needs is there.  While the lengths are still unknown (taken from
      // the mask they arrive ~10 us into a c2 call, 40 us into a c4 call) the DP warps look for them every eighth
      // superstep only: the look costs ~80 cycles on the lone warp's critical path (c4: 2.6 us per call), and nothing
      // here needs them before the utterance's last frame -- a warp that learns them late runs at most seven supersteps
      // past a short utterance's end, on padding, like every warp does before they arrive.  (Probing the flag inside the
      // superstep's block instead: one more volatile load in the body, +0.5 us at c2 with the lengths given.)  `lim` folds
      // "known and s < NS" and "unknown and not yet the eighth" into the one comparison the loop had before.
      if (pre && s < lim) continue;
      if (!known) check_lens();
      if (s >= NS) {
        if (known) break;
        while (!known) check_lens();  // ran through every frame that exists before the lengths arrived
        if (s >= NS) break;
      }
      lim = known ? NS : min(NS, (s | 7) + 1);
      if (!pre) blocking_start(s);
    }
    if (!shadow) emit_words(s - 1 - Q, s - 1 >= Q, std::true_type{});  // the last group's words
  }
  if (shadow) {
    leave();
    return;
  }
  if (lane0) tl_max(p.tl, 1);
  if (lane0) tl_max(p.tl, 2);
#ifdef MAS_TRACE
  if (p.trace && lane0) atomicMax(p.trace + 8 * 256 * 8 + 2 * b + 1, globaltimer_ns());
#endif
  leave();
}

template <int K, int D, int CL>
inline cudaError_t launch_dp2_t(const CUtensorMap& tmap, const DpParams& p, cudaStream_t st) {
  auto kern = mas_dp2_kernel<K, D, CL>;
  static std::atomic<uint64_t> attr_set{0};
  if (cudaError_t e = ensure_dyn_smem(kern, 227 * 1024, attr_set); e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.B * CL);
  cfg.blockDim = dim3(p.W <= 3 ? 32 * 12 : 32 * (p.W + 4 + 1 + (CL > 1 ? 1 : 0)));  // (+ the courier warp)
  cfg.dynamicSmemBytes = p.sm.total;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (p.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (CL > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CL;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, tmap, p);
}

// K = 2 only (the automatic choice for every T_x the wavefront kernel covers); skew 1 or 2; cl = CTAs per utterance (1, 2)
cudaError_t launch_dp2_k2(const CUtensorMap& tmap, const DpParams& p, int skew, int cl, cudaStream_t st);

}  // namespace mas
