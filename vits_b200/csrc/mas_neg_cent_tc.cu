// mas_neg_cent_tc.cu -- neg_cent (reference SynthesizerTrn.py:223-232) on the 5th-generation tensor
// cores: one fused tcgen05 GEMM per [128 frames x T_x] tile with the accumulator in TMEM.
//
//   neg_cent[b,t,s] = bias[b,s] + sum_d a2[d,t]*iv[d,s] + sum_d z[d,t]*mv[d,s]        (K = 2C)
//
// fp32 parity on bf16 tensor cores: every operand is split x = hi + lo (two bf16, 16 mantissa
// bits together) and each product is evaluated as hi*hi + hi*lo + lo*hi with fp32 accumulation in
// TMEM; the dropped lo*lo term is < 2^-16 relative per product, ~1e-7 of max|neg_cent| after the
// 384-term sum (tests/test_neg_cent_gpu.py holds the 1e-5 bar).  kind::tf32 alone (10 mantissa
// bits) would miss it; 3xTF32 costs twice the tensor time of this split.
//
// Two kernels, chained with programmatic dependent launch:
//   neg_cent_prep_kernel   text side, once per utterance: iv = exp(-2 logs_p), mv = m_p*iv split
//                          into bf16 hi/lo and stored ALREADY in the shared-memory operand layout
//                          the MMA wants (canonical K-major, no swizzle), plus the per-column bias
//                          (:225 and :231).  19 MB for B=64: stays in L2 for the GEMM.
//   neg_cent_tc_kernel     persistent, warp-specialised:
//        warp 0      B loader: one 1-D bulk async copy (TMA engine) per stage brings the four
//                    pre-laid-out B operand tiles; mbarrier complete_tx
//        warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=T_x tile, K=16) --
//                    6 operand pairs x 2 k-steps per 32-channel stage; tcgen05.commit frees the
//                    stage and publishes the accumulator
//        warps 2-9   epilogue: tcgen05.ld the accumulator (TMEM -> registers), add the bias,
//                    store the fp32 rows; double-buffered TMEM so it overlaps the next tile's MMAs
//        warps 10-13 z loaders: read z_p (coalesced along frames) into an fp32 staging ring
//        warps 14-21 converters: form -0.5 z^2 and z, split to bf16 hi/lo (integer rounding) and write
//                    the four A operand tiles straight into the MMA layout, fence.proxy.async, arrive
//                    (the transform is why A cannot simply be TMA-loaded; loaders and converters are
//                    separate warps because the proxy fence waits for a thread's outstanding loads)
//
// Operand tiles use the canonical K-major SWIZZLE_NONE layout: 8-row x 16-byte core matrices,
// address(row, k) = (k/8)*LBO + (row/8)*128 + (row%8)*16 + (k%8)*2, i.e. [k/8][row][k%8]; with
// 16-byte chunks contiguous along rows both our producer stores and the tensor core reads are
// bank-conflict free, so no swizzle is needed.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"
#include "ptx_sm100.cuh"

namespace mas {

namespace tc {
constexpr int M = 128;            // frames per tile (UMMA M)
constexpr int CB = 16;            // channels per pipeline stage (one UMMA k-step per operand pair)
constexpr int STAGES = 4;         // operand ring: a stage freed by MMA(g) must be refilled before MMA(g+3) ends
constexpr int PCB = 32;           // channels per prep-kernel CTA
constexpr int A_ARR = M * CB * 2; // bytes of one A operand array per stage (8 KB)
constexpr int A_LBO = M * 16;     // bytes between 8-channel chunks of A
constexpr int W_LOAD = 0, W_MMA = 1, W_EPI0 = 2, N_EPI = 8, W_ZL0 = 10, W_CV0 = 14, N_ZL = 4, N_CV = 8, N_WARPS = 22;
constexpr int ZS_MAX = 6;             // fp32 z staging ring between the z loaders and the converters (p.ZS slots)
constexpr int Z_STAGE = CB * M * 4;   // 8 KB
constexpr int TMEM_COLS = 512;    // two accumulator buffers of up to 256 columns
}  // namespace tc

struct TcParams {
  const float* z_p;      // [B][C][T_y]
  float* out;            // [B][T_y][T_x]
  const unsigned char* bops;  // [B][NTL][NCB][4 arrays][CB/8 chunks][Nt][8] bf16
  const float* bias;     // [nsplit][B][NTL*Nt] partial sums
  int nsplit;
  int B, C, T_y, T_x;
  int Nt;                // columns per tile, multiple of 16, <= 256
  int NTL;               // column tiles
  int MT;                // frame tiles
  int NCB;               // channel blocks = pipeline stages per tile (ceil(C/CB))
  int tiles;             // B*MT*NTL
  int ZS;                // z staging slots (3, or fewer when the column tile is wide)
  // streamed mode (stats -> path without the HBM round trip of neg_cent, SynthesizerTrn.py:223-235): tiles are taken in
  // FRAME-major order (all utterances' frames 0..127, then 128..255, ...) and written into a per-utterance ring of RT
  // tiles that the concurrently running DP kernel (mas_dp.cuh) drains through L2; flags[b][mt] counts the column tiles
  // of frame block mt that have landed.  Tiles past an utterance's length are skipped altogether.
  int stream;            // 0: dense [B][T_y][T_x] output; 1: ring + flags
  float* ring;           // [B][RT*128][pitch]
  uint32_t* flags;       // [B][MT], zeroed by the prep kernel
  const int32_t* t_ys;   // [B] device lengths (stream mode)
  const int32_t* t_xs;
  int RT, pitch;
  // Two-phase schedule of the streamed mode: the grid covers EVERY SM; the first n1 * gridDim.x tiles (frame-major) are
  // dealt round-robin to all CTAs, the rest only to the first n_long CTAs.  The other CTAs exit after their n1 tiles,
  // and the DP kernel's CTAs (launched programmatically behind this kernel) take over their SMs: the contraction gets
  // the whole machine until the search has enough frame blocks to start on.
  int n_long, n1;
  unsigned long long* tl;  // debug timeline (mas_set_timeline): slot 5 = first CTA start, slot 6 = last tile published
  int dbg;               // timing bisection only (wrong results): 1 no stores, 2 no A conversion, 4 no B copy, 8 no MMA; 16 trace; 32 no z loads; 64 no wait for prep; 128 prep only; 256 GEMM only
  unsigned long long* trace;  // dbg & 16: CTA 0 appends (tag, clock) pairs
};

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM helpers
// ------------------------------------------------------------------------------------------------
// debug trace (dbg & 16): CTA 0, one designated lane per role appends (tag, clock) to a shared-memory
// log (cheap: no global traffic inside the pipeline); the log is copied out at the end of the kernel.
constexpr int kTraceCap = 200;  // events per role
constexpr int kMaxLocalTiles = 256;  // streamed mode: tiles one CTA may own (host checks tiles / grid <= this)
__device__ __forceinline__ void trace_ev(const TcParams& p, unsigned long long* tsm, int* tcnt, int role, int ev, int idx) {
  if ((p.dbg & 16) && blockIdx.x == 0) {
    const int k = tcnt[role];
    if (k < kTraceCap) {
      tsm[(role * kTraceCap + k) * 2] = (static_cast<unsigned long long>(role) << 40) | (static_cast<unsigned long long>(ev) << 32) | static_cast<unsigned>(idx);
      tsm[(role * kTraceCap + k) * 2 + 1] = clock64();
      tcnt[role] = k + 1;
    }
  }
}

__device__ __forceinline__ unsigned long long ptx_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ptx::smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 16 TMEM lanes x 16 columns: lane quad t/4 <-> frame (and frame+8), lane%4 <-> column pair; registers
// {0,1} = frame A cols 0-7 slice, {2,3} = frame A+8, {4..7} = the same for columns 8-15.
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, canonical K-major layout without swizzle (version 1 = sm_100):
// bits [0,14) address>>4, [16,30) leading byte offset>>4 (between the two 8-element K chunks),
// [32,46) stride byte offset>>4 (between 8-row groups), [46,48) version, [61,64) layout type 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFFu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// Instruction descriptor: D fp32 (bits 4-5 = 1), A and B bf16 (bits 7-9, 10-12 = 1), both K-major,
// N>>3 at bit 17, M>>4 at bit 24.
__host__ __device__ constexpr uint32_t instr_desc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// x = hi + lo with hi, lo bf16, using integer rounding (round-half-up on the dropped 16 bits) instead
// of F2F conversions, which issue at a quarter of the ALU rate and were the A producers' bottleneck.
// Returns hi in the upper and lo in the lower 16 bits of the two outputs' high halves:
//   hi_bits, lo_bits are fp32 bit patterns whose low 16 bits are zero.
__device__ __forceinline__ void split_bits(float x, uint32_t& hi_bits, uint32_t& lo_bits) {
  hi_bits = (__float_as_uint(x) + 0x8000u) & 0xFFFF0000u;
  const float lo = x - __uint_as_float(hi_bits);  // exact
  lo_bits = (__float_as_uint(lo) + 0x8000u) & 0xFFFF0000u;
}
// two bf16 (given as fp32 bit patterns with zero low halves) -> one 32-bit word, first element low
__device__ __forceinline__ uint32_t pack2(uint32_t first_bits, uint32_t second_bits) {
  return __byte_perm(first_bits, second_bits, 0x7632);
}

// ------------------------------------------------------------------------------------------------
// text-side prep: B operand tiles (already in the MMA's shared-memory layout) + bias
// ------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* m_p;     // [B][C][T_x]
  const float* logs_p;  // [B][C][T_x]
  unsigned char* bops;
  float* bias;   // [nsplit][B][NTL*Nt] partial sums over the split's channels
  int B, C, T_x, Nt, NTL, NCB;
  uint32_t* zero_base;  // streamed mode: [B][zero_stride] words (tile flags, fill flag) cleared for this call; else nullptr
  int zero_stride;
};

// One small CTA (64 threads) per (utterance, 16 columns of a column tile, 32-channel block); thread =
// (8-channel chunk q, ONE text column): 16 scalar loads up front, 8 elements of work, one 16-byte group per
// operand array.  The earlier shape -- 4 columns per thread in CTAs of Nt threads -- left ~16 warps per SM,
// each running 1500 dependent instructions behind its loads, and took 12 us, all of it in front of the GEMM.
// Two-warp CTAs fill the SMs to 64 warps and have no wave tail (Nt is a multiple of 16).
// The per-column bias partial of the CTA's 32 channels goes to bias[pb]; the GEMM epilogue sums the partials.
__global__ void __launch_bounds__(64) neg_cent_prep_kernel(const PrepParams p) {
  __shared__ float sred[4][16];
  const int ncg = p.Nt >> 4;
  const int nt = blockIdx.x / ncg, cg = blockIdx.x - nt * ncg, b = blockIdx.y, pb = blockIdx.z;
  const int tid = threadIdx.x;
  ptx::pdl_launch_dependents();  // the GEMM kernel's A producers do not depend on us
  if (p.zero_base != nullptr && blockIdx.x == 0 && blockIdx.z == 0)
    for (int i = tid; i < p.zero_stride; i += 64) p.zero_base[static_cast<size_t>(b) * p.zero_stride + i] = 0u;
  const int q = tid >> 4;              // chunk of 8 channels inside the block
  const int n = cg * 16 + (tid & 15);  // column inside the tile
  const int s0 = nt * p.Nt + n;  // text column
  const bool col_ok = s0 < p.T_x;
  const int d0 = pb * tc::PCB + q * 8;
  const int sb = d0 / tc::CB, chunk = (d0 % tc::CB) / 8;  // pipeline stage block and 8-channel chunk inside it
  const size_t arr = static_cast<size_t>(p.Nt) * (tc::CB / 8) * 16;  // bytes of one operand array of a stage block
  const float* mb = p.m_p + static_cast<size_t>(b) * p.C * p.T_x + (col_ok ? s0 : 0);
  const float* lb = p.logs_p + static_cast<size_t>(b) * p.C * p.T_x + (col_ok ? s0 : 0);
  const float kHalfLog2Pi = 0.91893853320467274178f;
  float l[8], m[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {  // unconditional loads from a clamped address, masked below
    const size_t off = static_cast<size_t>(min(d0 + i, p.C - 1)) * p.T_x;
    l[i] = __ldg(lb + off);
    m[i] = __ldg(mb + off);
  }
  uint32_t ivh[4], ivl[4], mvh[4], mvl[4];
  float bias = 0.0f;
#pragma unroll
  for (int i2 = 0; i2 < 4; ++i2) {
    uint32_t h[2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * i2 + e;
      const bool ok = col_ok && d0 + i < p.C;
      const float iv = ok ? __expf(-2.0f * l[i]) : 0.0f;  // :223 (ex2.approx: ~2 ulp, far inside the 1e-5 bar)
      const float mv = m[i] * iv;                          // :229
      if (ok) bias += (-kHalfLog2Pi - l[i]) + (-0.5f * (m[i] * m[i])) * iv;  // :225, :231
      split_bits(iv, h[e][0], h[e][1]);
      split_bits(mv, h[e][2], h[e][3]);
    }
    ivh[i2] = pack2(h[0][0], h[1][0]);
    ivl[i2] = pack2(h[0][1], h[1][1]);
    mvh[i2] = pack2(h[0][2], h[1][2]);
    mvl[i2] = pack2(h[0][3], h[1][3]);
  }
  unsigned char* dst = p.bops + ((static_cast<size_t>(b) * p.NTL + nt) * p.NCB + sb) * (4 * arr) +
                       (static_cast<size_t>(chunk) * p.Nt + n) * 16;
  *reinterpret_cast<uint4*>(dst + 0 * arr) = make_uint4(ivh[0], ivh[1], ivh[2], ivh[3]);
  *reinterpret_cast<uint4*>(dst + 1 * arr) = make_uint4(ivl[0], ivl[1], ivl[2], ivl[3]);
  *reinterpret_cast<uint4*>(dst + 2 * arr) = make_uint4(mvh[0], mvh[1], mvh[2], mvh[3]);
  *reinterpret_cast<uint4*>(dst + 3 * arr) = make_uint4(mvl[0], mvl[1], mvl[2], mvl[3]);
  sred[q][tid & 15] = bias;
  __syncthreads();
  if (tid < 16)
    p.bias[((static_cast<size_t>(pb) * p.B + b) * p.NTL + nt) * p.Nt + n] =
        (sred[0][tid] + sred[1][tid]) + (sred[2][tid] + sred[3][tid]);
}

// ------------------------------------------------------------------------------------------------
// the GEMM
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(tc::N_WARPS * 32, 1) neg_cent_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int Nt = p.Nt;
  const uint32_t b_arr = static_cast<uint32_t>(Nt) * (tc::CB / 8) * 16u;  // bytes of one B operand array per stage
  const uint32_t b_lbo = static_cast<uint32_t>(Nt) * 16u;       // bytes between 8-channel chunks of B
  const uint32_t stage_bytes = 4u * tc::A_ARR + 4u * b_arr;

  // smem: [stage][A a2_hi | a2_lo | z_hi | z_lo | B iv_hi | iv_lo | mv_hi | mv_lo] ... barriers
  float* zstage = reinterpret_cast<float*>(smem + tc::STAGES * stage_bytes);  // [ZS][CB][M] fp32
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tc::STAGES * stage_bytes + p.ZS * tc::Z_STAGE);
  uint64_t* full = bars;                         // [STAGES] count N_CV/2 (one converter group) + 1 (B loader) + B bytes
  uint64_t* empty = full + tc::STAGES;           // [STAGES] count 1 (tcgen05.commit)
  uint64_t* t_full = empty + tc::STAGES;         // [2] count 1 (tcgen05.commit)
  uint64_t* t_empty = t_full + 2;                // [2] count 4 (epilogue warps)
  uint64_t* z_full = t_empty + 2;                // [ZS_MAX] count N_ZL
  uint64_t* z_empty = z_full + tc::ZS_MAX;       // [ZS_MAX] count N_CV/2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(z_empty + tc::ZS_MAX);
  float* sbias = reinterpret_cast<float*>(tmem_slot + 4);  // [2][256]
  unsigned long long* tsm = reinterpret_cast<unsigned long long*>(sbias + 512);  // debug trace log (dbg & 16 only)
  int* tcnt = reinterpret_cast<int*>(tsm + 4 * kTraceCap * 2);
  if ((p.dbg & 16) && tid < 4) tcnt[tid] = 0;
  __shared__ int s_ntiles;
  __shared__ int s_tiles[kMaxLocalTiles];

  if (tid == 0) {
    if (p.stream) {
      int n = 0;
      auto consider = [&](int tile) {
        const int b = (tile / p.NTL) % p.B, mt = tile / (p.NTL * p.B);
        const int ty = p.t_ys[b], tx = p.t_xs[b];
        const bool valid = ty >= 1 && tx >= 1 && ty <= p.T_y && tx <= p.T_x && tx <= ty;  // else: all-zero path, no tiles
        if (valid && mt * tc::M < ty && n < kMaxLocalTiles) s_tiles[n++] = tile;
      };
      const long long p1 = static_cast<long long>(p.n1) * gridDim.x;
      const int phase1 = p1 < p.tiles ? static_cast<int>(p1) : p.tiles;
      for (int tile = blockIdx.x; tile < phase1; tile += gridDim.x) consider(tile);
      if (static_cast<int>(blockIdx.x) < p.n_long)
        for (int tile = phase1 + blockIdx.x; tile < p.tiles; tile += p.n_long) consider(tile);
      s_ntiles = n;
    }
    for (int s = 0; s < tc::STAGES; ++s) {
      ptx::mbar_init(&full[s], tc::N_CV / 2 + 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], tc::N_EPI);
    }
    for (int i = 0; i < tc::ZS_MAX; ++i) {
      ptx::mbar_init(&z_full[i], tc::N_ZL * 32);  // one asynchronous arrival per loader thread
      ptx::mbar_init(&z_empty[i], tc::N_CV / 2);
    }
    ptx::mbar_fence_init();
  }
  if (warp == tc::W_MMA) tmem_alloc(tmem_slot, tc::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.stream && p.tl && tid == 0) atomicMin(p.tl + 5, ptx_globaltimer());

  // Tile schedule of this CTA: tile = blockIdx.x + k * gridDim.x.  Streamed mode keeps only the tiles inside their
  // utterance's length (compacted list in shared memory, built once by thread 0 before the barrier above).
  const int ntile_local = p.stream ? s_ntiles : (p.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  auto tile_of = [&](int lt) -> int { return p.stream ? s_tiles[lt] : static_cast<int>(blockIdx.x) + lt * static_cast<int>(gridDim.x); };
  auto decode = [&](int tile, int& b, int& mt, int& nt) {
    nt = tile % p.NTL;
    if (p.stream) {
      b = (tile / p.NTL) % p.B;
      mt = tile / (p.NTL * p.B);
    } else {
      mt = (tile / p.NTL) % p.MT;
      b = tile / (p.NTL * p.MT);
    }
  };

  if (warp == tc::W_LOAD) {
    // ---------------- B loader ----------------
    if (lane == 0) {
      if (!(p.dbg & 64)) ptx::pdl_wait();  // the prep kernel's tiles must be complete
      // Streamed mode: the dependent DP kernel may only start once the prep kernel has zeroed this call's flags,
      // i.e. after the wait above (the ordinary mode never triggers early, see below).
      if (p.stream) ptx::pdl_launch_dependents();
      uint32_t it = 0;
      for (int lt = 0; lt < ntile_local; ++lt) {
        // (Ordinary mode never triggers its dependents early: the search's forward kernel, launched programmatically behind
        // this kernel, would become resident on whatever SMs drain first -- at kernel start: one by one, the search 2 us
        // slower; in every CTA's last tile: on the 80 SMs of the CTAs with one tile fewer, 1.3 us slower than a forward
        // kernel that is placed over the whole machine when this kernel has ended.)
        const int tile = tile_of(lt);
        int b, mt_, nt;
        decode(tile, b, mt_, nt);
        const unsigned char* src = p.bops + (static_cast<size_t>(b) * p.NTL + nt) * p.NCB * (4 * static_cast<size_t>(b_arr));
        for (int cb = 0; cb < p.NCB; ++cb, ++it) {
          const int s = it % tc::STAGES;
          if (it >= tc::STAGES) ptx::mbar_wait(&empty[s], ((it / tc::STAGES) - 1) & 1);
          trace_ev(p, tsm, tcnt, 0, 1, it);
          if (p.dbg & 4) {
            ptx::mbar_arrive(&full[s]);
            continue;
          }
          ptx::mbar_arrive_expect_tx(&full[s], 4u * b_arr);
          ptx::bulk_g2s(smem + s * stage_bytes + 4u * tc::A_ARR, src + static_cast<size_t>(cb) * (4 * static_cast<size_t>(b_arr)), 4u * b_arr,
                        &full[s]);
        }
      }
    }
  } else if (warp == tc::W_MMA) {
    // ---------------- MMA issuer ----------------
    // The whole warp runs this loop and one elected lane issues: with warp-uniform control flow the
    // descriptors, the stage offset and the barrier addresses live in uniform registers.  (Inside an
    // `if (lane == 0)` the same code kept them in vector registers and paid a chain of R2UR moves in front of
    // every tcgen05.mma -- ~100 cycles of issue per 96-cycle MMA, with the tensor pipe waiting for this thread.)
    {
      const uint32_t idesc = instr_desc(tc::M, Nt);
      // Base descriptors of the 4 A and 4 B operand arrays in stage 0; a stage only shifts the 14-bit
      // address field (smem < 256 KB, so the add never carries out of it).
      uint64_t adesc[4], bdesc[4];
      {
        const uint32_t a0 = ptx::smem_u32(smem);
        const uint32_t b0 = a0 + 4u * tc::A_ARR;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          adesc[i] = smem_desc(a0 + i * tc::A_ARR, tc::A_LBO, 128);
          bdesc[i] = smem_desc(b0 + i * b_arr, b_lbo, 128);
        }
      }
      const uint32_t stage_step = stage_bytes >> 4;  // descriptor address units
      const bool mma = !(p.dbg & 8);
      uint32_t it = 0;
      int s = 0;
      uint32_t par = 0;
      for (int lt = 0; lt < ntile_local; ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) ptx::mbar_wait(&t_empty[buf], ((lt >> 1) - 1) & 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf) * 256u;
        for (int cb = 0; cb < p.NCB; ++cb, ++it) {
          if (lane == 0) trace_ev(p, tsm, tcnt, 1, 0, it);
          // test_wait first: ~10x cheaper than try_wait when the phase already completed (the usual case)
          if (!ptx::mbar_test(&full[s], par)) ptx::mbar_wait(&full[s], par);
          if (lane == 0) trace_ev(p, tsm, tcnt, 1, 2, it);
          tc_fence_after();
          const uint64_t soff = static_cast<uint64_t>(static_cast<uint32_t>(s) * stage_step);
          if (elect_one()) {
            if (mma) {
              // operand pairs: (a2_hi,iv_hi) (a2_hi,iv_lo) (a2_lo,iv_hi) (z_hi,mv_hi) (z_hi,mv_lo) (z_lo,mv_hi)
              umma_bf16(d_tmem, adesc[0] + soff, bdesc[0] + soff, idesc, cb != 0 ? 1u : 0u);
              umma_bf16(d_tmem, adesc[0] + soff, bdesc[1] + soff, idesc, 1u);
              umma_bf16(d_tmem, adesc[1] + soff, bdesc[0] + soff, idesc, 1u);
              umma_bf16(d_tmem, adesc[2] + soff, bdesc[2] + soff, idesc, 1u);
              umma_bf16(d_tmem, adesc[2] + soff, bdesc[3] + soff, idesc, 1u);
              umma_bf16(d_tmem, adesc[3] + soff, bdesc[2] + soff, idesc, 1u);
            }
            umma_commit(&empty[s]);  // stage reusable once these MMAs have read it
            if (cb == p.NCB - 1) umma_commit(&t_full[buf]);  // accumulator complete
          }
          __syncwarp();
          if (lane == 0) trace_ev(p, tsm, tcnt, 1, 3, it);
          if (++s == tc::STAGES) {
            s = 0;
            par ^= 1u;
          }
        }
      }
    }
  } else if (warp >= tc::W_EPI0 && warp < tc::W_ZL0) {
    // ---------------- epilogue: TMEM -> registers -> (+bias) -> global ----------------
    // tcgen05.ld shape 16x256b: a quad of lanes holds 8 consecutive fp32 columns (one 32-byte sector)
    // of one frame, so every global store instruction writes whole sectors exactly once.
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int eh = (warp - tc::W_EPI0) >> 2;  // two warps share a lane quarter: even / odd column groups
    const int cq = 2 * (lane & 3);  // this lane's column pair inside each group of 8
    if (!(p.dbg & 64)) ptx::pdl_wait();  // bias comes from the prep kernel
    constexpr int kBiasRegs = 8;
    float bpart[kBiasRegs];
    auto bias_fetch = [&](int lt_) {
      int b_, mt__, nt_;
      decode(tile_of(lt_), b_, mt__, nt_);
      const int e = (warp - tc::W_EPI0) * 32 + lane;
#pragma unroll
      for (int sp = 0; sp < kBiasRegs; ++sp)
        bpart[sp] = (e < Nt && sp < p.nsplit) ? p.bias[((static_cast<size_t>(sp) * p.B + b_) * p.NTL + nt_) * Nt + e] : 0.0f;
      for (int sp = kBiasRegs; sp < p.nsplit; ++sp)  // more than 8 channel blocks (C > 256): summed right away
        if (e < Nt) bpart[0] += p.bias[((static_cast<size_t>(sp) * p.B + b_) * p.NTL + nt_) * Nt + e];
    };
    if (ntile_local > 0) bias_fetch(0);
    for (int lt = 0; lt < ntile_local; ++lt) {
      const int tile = tile_of(lt);
      int b, mt, nt;
      decode(tile, b, mt, nt);
      const int buf = lt & 1;
      const int n0 = nt * Nt;
      // bias of this tile's columns (sum of the prep kernel's per-channel-block partials): the partials were
      // loaded into registers during the previous tile's drain and are only summed here, so their latency
      // is never waited for (Nt <= 256 = one column per epilogue thread)
      {
        const int e = (warp - tc::W_EPI0) * 32 + lane;
        if (e < Nt) {
          float v = 0.0f;
#pragma unroll
          for (int sp = 0; sp < kBiasRegs; ++sp) v += bpart[sp];
          sbias[buf * 256 + e] = v;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(tc::N_EPI * 32) : "memory");  // the epilogue warps only
        if (lt + 1 < ntile_local) bias_fetch(lt + 1);
      }
      if (warp == tc::W_EPI0 && lane == 0) trace_ev(p, tsm, tcnt, 2, 0, lt);
      ptx::mbar_wait(&t_full[buf], (lt >> 1) & 1);
      if (warp == tc::W_EPI0 && lane == 0) trace_ev(p, tsm, tcnt, 2, 1, lt);
      tc_fence_after();
      const float* bias = sbias + buf * 256;
      const bool pair_ok = (p.T_x & 1) == 0;
      const int ngroups = Nt >> 4;  // 16 columns per tcgen05.ld
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int mrow = quarter * 32 + h * 16 + (lane >> 2);  // this lane's first frame; the second is +8
        const int tA = mt * tc::M + mrow, tB = tA + 8;
        // streamed mode: row (mt % RT)*128 + mrow of utterance b's ring, every row and column of the tile is stored
        const int ostride = p.stream ? p.pitch : p.T_x;
        float* rowA = p.stream ? p.ring + (static_cast<size_t>(b) * p.RT * tc::M + static_cast<size_t>(mt % p.RT) * tc::M + mrow) * p.pitch + n0
                               : p.out + (static_cast<size_t>(b) * p.T_y + tA) * p.T_x + n0;
        float* rowB = rowA + static_cast<size_t>(8) * ostride;
        const uint32_t taddr = tmem_base + static_cast<uint32_t>(buf) * 256u + (static_cast<uint32_t>(quarter * 32 + h * 16) << 16);
        // (+bias) and store one group of 16 columns held in registers
        auto emit = [&](int g, const uint32_t (&cur)[8]) {
          if (p.dbg & 1) return;
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int c = g * 16 + i * 8 + cq;
            const float2 bq = *reinterpret_cast<const float2*>(bias + c);
            const float2 oA = make_float2(__uint_as_float(cur[4 * i + 0]) + bq.x, __uint_as_float(cur[4 * i + 1]) + bq.y);
            const float2 oB = make_float2(__uint_as_float(cur[4 * i + 2]) + bq.x, __uint_as_float(cur[4 * i + 3]) + bq.y);
            const int n = n0 + c;
            if (p.stream) {
              *reinterpret_cast<float2*>(rowA + c) = oA;
              *reinterpret_cast<float2*>(rowB + c) = oB;
            } else if (pair_ok && n + 1 < p.T_x) {
              if (tA < p.T_y) *reinterpret_cast<float2*>(rowA + c) = oA;
              if (tB < p.T_y) *reinterpret_cast<float2*>(rowB + c) = oB;
            } else {
              if (tA < p.T_y && n < p.T_x) rowA[c] = oA.x;
              if (tA < p.T_y && n + 1 < p.T_x) rowA[c + 1] = oA.y;
              if (tB < p.T_y && n < p.T_x) rowB[c] = oB.x;
              if (tB < p.T_y && n + 1 < p.T_x) rowB[c + 1] = oB.y;
            }
          }
        };
        // two register sets, statically named: the load of group g+1 is in flight while group g is stored
        uint32_t ra[8], rb[8];
        if (eh < ngroups) {
          tmem_ld_16x256b_x2(taddr + eh * 16, ra);
          tmem_wait_ld();
        }
        for (int g = eh; g < ngroups; g += 4) {
          if (g + 2 < ngroups) tmem_ld_16x256b_x2(taddr + (g + 2) * 16, rb);
          emit(g, ra);
          tmem_wait_ld();
          if (g + 2 >= ngroups) break;
          if (g + 4 < ngroups) tmem_ld_16x256b_x2(taddr + (g + 4) * 16, ra);
          emit(g + 2, rb);
          tmem_wait_ld();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (warp == tc::W_EPI0 && lane == 0) trace_ev(p, tsm, tcnt, 2, 2, lt);
      if (lane == 0) ptx::mbar_arrive(&t_empty[buf]);
      if (p.stream) {
        // publish the tile: every epilogue warp's stores are ordered before the barrier, the fence makes them
        // visible GPU-wide (cumulativity) before the count that the DP kernel's producer warps acquire
        asm volatile("bar.sync 1, %0;" ::"n"(tc::N_EPI * 32) : "memory");
        if (warp == tc::W_EPI0 && lane == 0) {
          __threadfence();
          atomicAdd(p.flags + static_cast<size_t>(b) * p.MT + mt, 1u);
          if (p.tl) atomicMax(p.tl + 6, ptx_globaltimer());
        }
      }
    }
  } else if (warp >= tc::W_ZL0 && warp < tc::W_CV0) {
    // ---------------- z loaders: z_p (global, coalesced along frames) -> fp32 staging ring ----------------
    // cp.async (LDGSTS) straight into shared memory: no registers, no blocking -- a loader thread only
    // waits for a free staging slot, so ZS stages of HBM latency are in flight per SM (with register
    // staging only ~16 KB/SM were in flight and the whole pipeline paced at the HBM round trip).
    // Kept apart from the converters on purpose: the converters must execute fence.proxy.async, and that
    // fence waits for every outstanding memory operation of its thread.
    const int lt_id = tid - tc::W_ZL0 * 32;  // 0..127
    const int G = ntile_local * p.NCB;       // stages this CTA runs
    const size_t zstride = static_cast<size_t>(p.T_y);
    const bool vec16 = ((p.T_y & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.z_p) & 15u) == 0);
    // 16-byte mode: thread = (4 consecutive frames, channel lt_id/32 + 4k); 4-byte mode: thread = frame lt_id
    const int fm = vec16 ? (lt_id & 31) * 4 : lt_id;
    const int ch0 = vec16 ? (lt_id >> 5) : 0;
    int ld_lt = 0, ld_cb = 0, zs = 0, use = 0;
    const float* ld_ptr = nullptr;
    int t_rem = 0;  // frames of this thread that exist (vec16: 0..4, else 0..1)
    auto ld_new_tile = [&]() {
      int b, mt, nt_;
      decode(tile_of(ld_lt), b, mt, nt_);
      const int t = mt * tc::M + fm;
      t_rem = max(0, min(vec16 ? 4 : 1, p.T_y - t));
      ld_ptr = p.z_p + static_cast<size_t>(b) * p.C * p.T_y + (t_rem > 0 ? t : 0);
    };
    if (G > 0) ld_new_tile();
    for (int g = 0; g < G; ++g) {
      if (use >= 1) ptx::mbar_wait(&z_empty[zs], (use - 1) & 1);  // converters finished reading the slot
      float* dst = zstage + static_cast<size_t>(zs) * (tc::CB * tc::M) + fm;
      const int d0 = ld_cb * tc::CB;
      if (p.dbg & 32) {
      } else if (vec16) {
#pragma unroll
        for (int k = 0; k < tc::CB / 4; ++k) {
          const int ch = ch0 + 4 * k;
          const int nbytes = (d0 + ch < p.C) ? t_rem * 4 : 0;  // bytes beyond are zero-filled
          const float* src = ld_ptr + static_cast<size_t>(d0 + ch < p.C ? d0 + ch : 0) * zstride;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(ptx::smem_u32(dst + ch * tc::M)), "l"(src),
                       "r"(nbytes)
                       : "memory");
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < tc::CB; ++ch) {
          const int nbytes = (d0 + ch < p.C) ? t_rem * 4 : 0;
          const float* src = ld_ptr + static_cast<size_t>(d0 + ch < p.C ? d0 + ch : 0) * zstride;
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ptx::smem_u32(dst + ch * tc::M)), "l"(src),
                       "r"(nbytes)
                       : "memory");
        }
      }
      // arrive on z_full[zs] when this thread's copies have landed (does not block the thread)
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(&z_full[zs])) : "memory");
      if (++zs == p.ZS) {
        zs = 0;
        ++use;
      }
      if (++ld_cb == p.NCB) {
        ld_cb = 0;
        ++ld_lt;
        if (g + 1 < G) ld_new_tile();
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // nothing may still be writing our shared memory at exit
  } else if (warp >= tc::W_CV0) {
    // ---------------- converters: staging -> (-0.5 z^2, z) -> bf16 hi/lo operand tiles ----------------
    // Two groups of four warps work on ALTERNATE stages: one stage's chain (wait for z, 16 loads, split,
    // stores, proxy fence) is ~500 cycles of latency, and with all eight warps in lockstep on one stage that
    // latency -- not the tensor pipe (~580 cycles of MMA per stage) -- set the stage period (measured in the
    // role trace: converters 840 cycles per stage, MMA issuer waiting for operands).
    const int ct = tid - tc::W_CV0 * 32;   // 0..255
    const int m = ct & (tc::M - 1);        // frame inside the tile
    const int grp = ct >> 7;               // which of the two converter groups
    const int G = ntile_local * p.NCB;
    for (int g = grp; g < G; g += 2) {
      const uint32_t it = static_cast<uint32_t>(g);
      const int s = it % tc::STAGES;
      const int zs = g % p.ZS;
      const int zuse = g / p.ZS;
      if (warp == tc::W_CV0 && lane == 0) trace_ev(p, tsm, tcnt, 3, 0, g);
      ptx::mbar_wait(&z_full[zs], zuse & 1);
      if (warp == tc::W_CV0 && lane == 0) trace_ev(p, tsm, tcnt, 3, 3, g);
      const float* src = zstage + static_cast<size_t>(zs) * (tc::CB * tc::M) + m;
      float z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = src[i * tc::M];
      if (it >= tc::STAGES) ptx::mbar_wait(&empty[s], ((it / tc::STAGES) - 1) & 1);  // MMAs of the previous use are done
      if (warp == tc::W_CV0 && lane == 0) trace_ev(p, tsm, tcnt, 3, 1, g);
      unsigned char* a_base = smem + s * stage_bytes;
      if (!(p.dbg & 2)) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {  // the two 8-channel chunks of the stage's 16 channels
          uint32_t a2h[4], a2l[4], zh[4], zl[4];
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            uint32_t h[2][4];  // [elem][a2_hi, a2_lo, z_hi, z_lo] as fp32 bit patterns
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float zz = z[half * 8 + i2 * 2 + e];
              const float a2 = -0.5f * (zz * zz);  // :227
              split_bits(a2, h[e][0], h[e][1]);
              split_bits(zz, h[e][2], h[e][3]);
            }
            a2h[i2] = pack2(h[0][0], h[1][0]);
            a2l[i2] = pack2(h[0][1], h[1][1]);
            zh[i2] = pack2(h[0][2], h[1][2]);
            zl[i2] = pack2(h[0][3], h[1][3]);
          }
          const uint32_t off = static_cast<uint32_t>(half) * tc::A_LBO + static_cast<uint32_t>(m) * 16u;
          *reinterpret_cast<uint4*>(a_base + 0 * tc::A_ARR + off) = make_uint4(a2h[0], a2h[1], a2h[2], a2h[3]);
          *reinterpret_cast<uint4*>(a_base + 1 * tc::A_ARR + off) = make_uint4(a2l[0], a2l[1], a2l[2], a2l[3]);
          *reinterpret_cast<uint4*>(a_base + 2 * tc::A_ARR + off) = make_uint4(zh[0], zh[1], zh[2], zh[3]);
          *reinterpret_cast<uint4*>(a_base + 3 * tc::A_ARR + off) = make_uint4(zl[0], zl[1], zl[2], zl[3]);
        }
      }
      fence_proxy_async();  // make the generic-proxy stores visible to the tensor core (async proxy)
      __syncwarp();
      if (warp == tc::W_CV0 && lane == 0) trace_ev(p, tsm, tcnt, 3, 2, g);
      if (lane == 0) {
        ptx::mbar_arrive(&full[s]);
        ptx::mbar_arrive(&z_empty[zs]);  // the staging slot was fully read (values are in registers)
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == tc::W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tc::TMEM_COLS);
  }
  if ((p.dbg & 16) && blockIdx.x == 0 && tid == 0) {
    unsigned long long n = 0;
    for (int r = 0; r < 4; ++r)
      for (int k = 0; k < tcnt[r]; ++k, ++n) {
        p.trace[2 + 2 * n] = tsm[(r * kTraceCap + k) * 2];
        p.trace[3 + 2 * n] = tsm[(r * kTraceCap + k) * 2 + 1];
      }
    p.trace[0] = n;
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
struct TcShape {
  int Nt, NTL, MT, NCB, NPB;
  size_t bops_bytes, bias_bytes, total;
};

static TcShape tc_shape(int B, int C, int T_y, int T_x) {
  TcShape s;
  s.NTL = (T_x + 255) / 256;
  const int per = (T_x + s.NTL - 1) / s.NTL;
  s.Nt = ((per + 15) / 16) * 16;
  s.MT = (T_y + tc::M - 1) / tc::M;
  s.NCB = ((C + tc::PCB - 1) / tc::PCB) * (tc::PCB / tc::CB);  // padded to whole prep blocks (zeros beyond C)
  s.NPB = (C + tc::PCB - 1) / tc::PCB;
  s.bops_bytes = static_cast<size_t>(B) * s.NTL * (static_cast<size_t>(s.NPB) * tc::PCB / tc::CB) * s.Nt * (8 * tc::CB);
  s.bias_bytes = static_cast<size_t>(s.NPB) * B * s.NTL * s.Nt * 4;  // one partial per prep block
  s.total = ((s.bops_bytes + 255) & ~size_t(255)) + ((s.bias_bytes + 255) & ~size_t(255));
  return s;
}

size_t neg_cent_tc_scratch_bytes(int B, int C, int T_y, int T_x) { return tc_shape(B, C, T_y, T_x).total; }

void neg_cent_tc_dims(int C, int T_y, int T_x, int* Nt, int* NTL, int* MT) {
  const TcShape s = tc_shape(1, C, T_y, T_x);
  *Nt = s.Nt;
  *NTL = s.NTL;
  *MT = s.MT;
}

int neg_cent_tc(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
                int B, int C, int T_y, int T_x, cudaStream_t st) {
  return neg_cent_tc_impl(z_p, m_p, logs_p, out, scratch, scratch_bytes, B, C, T_y, T_x, st, nullptr);
}

int neg_cent_tc_impl(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
                     int B, int C, int T_y, int T_x, cudaStream_t st, const TcStream* so) {
  const TcShape s = tc_shape(B, C, T_y, T_x);
  if (scratch_bytes < s.total || !scratch) return MAS_E_SCRATCH;
  if (reinterpret_cast<uintptr_t>(scratch) & 15u) return MAS_E_ALIGN;
  unsigned char* bops = static_cast<unsigned char*>(scratch);
  float* bias = reinterpret_cast<float*>(bops + ((s.bops_bytes + 255) & ~size_t(255)));

  PrepParams pp{m_p, logs_p, bops, bias, B, C, T_x, s.Nt, s.NTL, s.NCB, so ? so->zero_base : nullptr, so ? so->zero_stride : 0};
  static const int g_dbg = getenv("MAS_NC_DEBUG") ? atoi(getenv("MAS_NC_DEBUG")) : 0;  // read once (benchmark bisection hook)
  if (!(g_dbg & 256)) neg_cent_prep_kernel<<<dim3(s.NTL * (s.Nt / 16), B, s.NPB), 64, 0, st>>>(pp);  // 4 chunks x 16 columns
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();

  TcParams tp{};
  tp.z_p = z_p; tp.out = out; tp.bops = bops; tp.bias = bias; tp.nsplit = s.NPB;
  tp.B = B; tp.C = C; tp.T_y = T_y; tp.T_x = T_x;
  tp.Nt = s.Nt; tp.NTL = s.NTL; tp.MT = s.MT; tp.NCB = s.NCB;
  tp.tiles = B * s.MT * s.NTL;
  if (so) {
    tp.stream = 1;
    tp.ring = so->ring; tp.flags = so->flags; tp.t_ys = so->t_ys; tp.t_xs = so->t_xs;
    tp.RT = so->RT; tp.pitch = so->pitch;
    tp.tl = timeline_ptr();
    const int grid = tp.tiles < so->grid ? tp.tiles : so->grid;
    tp.n_long = so->n_long < grid ? so->n_long : grid;
    tp.n1 = so->n1;
    if (grid < 1 || tp.n_long < 1 || tp.n1 < 0 || tp.n1 + (tp.tiles + tp.n_long - 1) / tp.n_long > kMaxLocalTiles)
      return MAS_E_UNSUPPORTED;
  }
  tp.dbg = g_dbg;
  static unsigned long long* d_trace = nullptr;
  if (tp.dbg & 16) {
    if (!d_trace) cudaMalloc(&d_trace, (2 + 2 * 4000) * 8);
    cudaMemsetAsync(d_trace, 0, (2 + 2 * 4000) * 8, st);
    tp.trace = d_trace;
  }
  const size_t fixed = static_cast<size_t>(tc::STAGES) * (4 * tc::A_ARR + 4 * static_cast<size_t>(s.Nt) * (tc::CB / 8) * 16) + 512 + 2 * 256 * 4 + 128;
  tp.ZS = tc::ZS_MAX;
  // dynamic + static (tile list, 3 KB) shared memory must stay within the 227 KB opt-in limit
  while (tp.ZS > 1 && fixed + static_cast<size_t>(tp.ZS) * tc::Z_STAGE > 221 * 1024) --tp.ZS;
  const size_t smem = fixed + static_cast<size_t>(tp.ZS) * tc::Z_STAGE + ((tp.dbg & 16) ? (4 * 200 * 16 + 64) : 0);
  static std::atomic<uint64_t> attr{0};
  e = ensure_dyn_smem(neg_cent_tc_kernel, 223 * 1024, attr);
  if (e != cudaSuccess) return static_cast<int>(e);
  const int sms = num_sms();
  cudaLaunchConfig_t cfg{};
  const int want_ctas = so ? so->grid : sms;
  cfg.gridDim = dim3(tp.tiles < want_ctas ? tp.tiles : want_ctas);
  cfg.blockDim = dim3(tc::N_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = la;
  cfg.numAttrs = 1;
  if (tp.dbg & 128) return MAS_OK;
  e = cudaLaunchKernelEx(&cfg, neg_cent_tc_kernel, tp);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  if ((tp.dbg & 16) && getenv("MAS_NC_TRACE_DUMP")) {  // debug only: synchronises and prints CTA 0's event trace
    cudaStreamSynchronize(st);
    static unsigned long long h[2 + 2 * 4000];
    cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost);
    const unsigned long long n = h[0] < 4000 ? h[0] : 4000;
    unsigned long long t0 = ~0ull;
    for (unsigned long long i = 0; i < n; ++i) t0 = h[3 + 2 * i] < t0 ? h[3 + 2 * i] : t0;
    for (unsigned long long i = 0; i < n; ++i)
      printf("TRACE role %llu ev %llu idx %llu t %llu\n", h[2 + 2 * i] >> 40, (h[2 + 2 * i] >> 32) & 0xff,
             h[2 + 2 * i] & 0xffffffffull, h[3 + 2 * i] - t0);
  }
  return MAS_OK;
}

}  // namespace mas
