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
//        warps 2-5   epilogue: tcgen05.ld the accumulator (TMEM -> registers), add the bias,
//                    store the fp32 rows; double-buffered TMEM so it overlaps the next tile's MMAs
//        warps 6-13  A producers: read z_p (coalesced along frames), form -0.5 z^2 and z, split to
//                    bf16 hi/lo and write the four A operand tiles straight into the MMA layout
//                    (the transform is why A cannot simply be TMA-loaded)
//
// Operand tiles use the canonical K-major SWIZZLE_NONE layout: 8-row x 16-byte core matrices,
// address(row, k) = (k/8)*LBO + (row/8)*128 + (row%8)*16 + (k%8)*2, i.e. [k/8][row][k%8]; with
// 16-byte chunks contiguous along rows both our producer stores and the tensor core reads are
// bank-conflict free, so no swizzle is needed.
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"
#include "ptx_sm100.cuh"

namespace mas {

namespace tc {
constexpr int M = 128;            // frames per tile (UMMA M)
constexpr int CB = 32;            // channels per pipeline stage
constexpr int STAGES = 2;
constexpr int A_ARR = M * CB * 2; // bytes of one A operand array per stage (8 KB)
constexpr int A_LBO = M * 16;     // bytes between 8-channel chunks of A
constexpr int W_LOAD = 0, W_MMA = 1, W_EPI0 = 2, W_TR0 = 6, N_TR = 8, N_WARPS = 14;
constexpr int TMEM_COLS = 512;    // two accumulator buffers of up to 256 columns
}  // namespace tc

struct TcParams {
  const float* z_p;      // [B][C][T_y]
  float* out;            // [B][T_y][T_x]
  const unsigned char* bops;  // [B][NTL][NCB][4 arrays][4 chunks][Nt][8] bf16
  const float* bias;     // [B][NTL*Nt]
  int B, C, T_y, T_x;
  int Nt;                // columns per tile, multiple of 16, <= 256
  int NTL;               // column tiles
  int MT;                // frame tiles
  int NCB;               // channel blocks (ceil(C/32))
  int tiles;             // B*MT*NTL
};

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM helpers
// ------------------------------------------------------------------------------------------------
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

__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo_elem, __nv_bfloat16 hi_elem) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(lo_elem)) | (static_cast<uint32_t>(__bfloat16_as_ushort(hi_elem)) << 16);
}
// x = hi + lo with hi, lo bf16
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ------------------------------------------------------------------------------------------------
// text-side prep: B operand tiles (already in the MMA's shared-memory layout) + bias
// ------------------------------------------------------------------------------------------------
struct PrepParams {
  const float* m_p;     // [B][C][T_x]
  const float* logs_p;  // [B][C][T_x]
  unsigned char* bops;
  float* bias;
  int C, T_x, Nt, NTL, NCB;
};

__global__ void __launch_bounds__(256) neg_cent_prep_kernel(const PrepParams p) {
  const int cb = blockIdx.x, nt = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  ptx::pdl_launch_dependents();  // the GEMM kernel's A producers do not depend on us
  const float* mb = p.m_p + static_cast<size_t>(b) * p.C * p.T_x;
  const float* lb = p.logs_p + static_cast<size_t>(b) * p.C * p.T_x;
  const int n0 = nt * p.Nt;
  unsigned char* dst = p.bops + ((static_cast<size_t>(b) * p.NTL + nt) * p.NCB + cb) * (static_cast<size_t>(p.Nt) * 256);
  const size_t arr = static_cast<size_t>(p.Nt) * 64;  // bytes of one operand array (4 chunks x Nt x 16 B)
  // item = (chunk q of 8 channels, column n): one 16-byte group of 8 bf16 per operand array
  for (int item = tid; item < 4 * p.Nt; item += blockDim.x) {
    const int q = item / p.Nt, n = item - q * p.Nt;
    const int s = n0 + n;
    uint32_t ivh[4], ivl[4], mvh[4], mvl[4];
#pragma unroll
    for (int i2 = 0; i2 < 4; ++i2) {
      __nv_bfloat16 h[2][4];  // [elem][iv_hi, iv_lo, mv_hi, mv_lo]
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int d = cb * tc::CB + q * 8 + i2 * 2 + e;
        float iv = 0.0f, mv = 0.0f;
        if (d < p.C && s < p.T_x) {
          const float l = lb[static_cast<size_t>(d) * p.T_x + s];
          const float m = mb[static_cast<size_t>(d) * p.T_x + s];
          iv = expf(-2.0f * l);  // :223
          mv = m * iv;           // :229
        }
        split_bf16(iv, h[e][0], h[e][1]);
        split_bf16(mv, h[e][2], h[e][3]);
      }
      ivh[i2] = pack_bf16(h[0][0], h[1][0]);
      ivl[i2] = pack_bf16(h[0][1], h[1][1]);
      mvh[i2] = pack_bf16(h[0][2], h[1][2]);
      mvl[i2] = pack_bf16(h[0][3], h[1][3]);
    }
    const size_t off = (static_cast<size_t>(q) * p.Nt + n) * 16;
    *reinterpret_cast<uint4*>(dst + 0 * arr + off) = make_uint4(ivh[0], ivh[1], ivh[2], ivh[3]);
    *reinterpret_cast<uint4*>(dst + 1 * arr + off) = make_uint4(ivl[0], ivl[1], ivl[2], ivl[3]);
    *reinterpret_cast<uint4*>(dst + 2 * arr + off) = make_uint4(mvh[0], mvh[1], mvh[2], mvh[3]);
    *reinterpret_cast<uint4*>(dst + 3 * arr + off) = make_uint4(mvl[0], mvl[1], mvl[2], mvl[3]);
  }
  if (cb == 0) {
    // bias[s] = sum_d(-0.5 log 2pi - logs) + sum_d(-0.5 m^2 exp(-2 logs))     (:225, :231)
    const float kHalfLog2Pi = 0.91893853320467274178f;
    for (int n = tid; n < p.Nt; n += blockDim.x) {
      const int s = n0 + n;
      float t1 = 0.0f, t4 = 0.0f;
      if (s < p.T_x) {
        for (int d = 0; d < p.C; ++d) {
          const float l = lb[static_cast<size_t>(d) * p.T_x + s];
          const float m = mb[static_cast<size_t>(d) * p.T_x + s];
          t1 += -kHalfLog2Pi - l;
          t4 += (-0.5f * (m * m)) * expf(-2.0f * l);
        }
      }
      p.bias[(static_cast<size_t>(b) * p.NTL + nt) * p.Nt + n] = t1 + t4;
    }
  }
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
  const uint32_t b_arr = static_cast<uint32_t>(Nt) * 64u;       // bytes of one B operand array per stage
  const uint32_t b_lbo = static_cast<uint32_t>(Nt) * 16u;       // bytes between 8-channel chunks of B
  const uint32_t stage_bytes = 4u * tc::A_ARR + 4u * b_arr;

  // smem: [stage][A a2_hi | a2_lo | z_hi | z_lo | B iv_hi | iv_lo | mv_hi | mv_lo] ... barriers
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tc::STAGES * stage_bytes);
  uint64_t* a_full = bars;            // [STAGES] count N_TR
  uint64_t* b_full = bars + 2;        // [STAGES] count 1 + tx
  uint64_t* empty = bars + 4;         // [STAGES] count 1 (tcgen05.commit)
  uint64_t* t_full = bars + 6;        // [2] count 1 (tcgen05.commit)
  uint64_t* t_empty = bars + 8;       // [2] count 4 (epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  if (tid == 0) {
    for (int s = 0; s < tc::STAGES; ++s) {
      ptx::mbar_init(&a_full[s], tc::N_TR);
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&t_full[i], 1);
      ptx::mbar_init(&t_empty[i], 4);
    }
    ptx::mbar_fence_init();
  }
  if (warp == tc::W_MMA) tmem_alloc(tmem_slot, tc::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_launch_dependents();

  const int ntile_local = (p.tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == tc::W_LOAD) {
    // ---------------- B loader ----------------
    if (lane == 0) {
      ptx::pdl_wait();  // the prep kernel's tiles must be complete
      uint32_t it = 0;
      for (int lt = 0; lt < ntile_local; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        const int nt = tile % p.NTL;
        const int b = tile / (p.NTL * p.MT);
        const unsigned char* src = p.bops + (static_cast<size_t>(b) * p.NTL + nt) * p.NCB * (static_cast<size_t>(Nt) * 256);
        for (int cb = 0; cb < p.NCB; ++cb, ++it) {
          const int s = it & 1;
          if (it >= 2) ptx::mbar_wait(&empty[s], ((it >> 1) - 1) & 1);
          ptx::mbar_arrive_expect_tx(&b_full[s], 4u * b_arr);
          ptx::bulk_g2s(smem + s * stage_bytes + 4u * tc::A_ARR, src + static_cast<size_t>(cb) * Nt * 256, 4u * b_arr,
                        &b_full[s]);
        }
      }
    }
  } else if (warp == tc::W_MMA) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      const uint32_t idesc = instr_desc(tc::M, Nt);
      uint32_t it = 0;
      for (int lt = 0; lt < ntile_local; ++lt) {
        const int buf = lt & 1;
        if (lt >= 2) ptx::mbar_wait(&t_empty[buf], ((lt >> 1) - 1) & 1);  // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf) * 256u;
        for (int cb = 0; cb < p.NCB; ++cb, ++it) {
          const int s = it & 1;
          const uint32_t par = (it >> 1) & 1;
          ptx::mbar_wait(&a_full[s], par);
          ptx::mbar_wait(&b_full[s], par);
          tc_fence_after();
          const uint32_t a0 = ptx::smem_u32(smem + s * stage_bytes);
          const uint32_t b0 = a0 + 4u * tc::A_ARR;
          // operand pairs: (a2_hi,iv_hi) (a2_hi,iv_lo) (a2_lo,iv_hi) (z_hi,mv_hi) (z_hi,mv_lo) (z_lo,mv_hi)
          const int ai[6] = {0, 0, 1, 2, 2, 3};
          const int bi[6] = {0, 1, 0, 2, 3, 2};
#pragma unroll
          for (int pr = 0; pr < 6; ++pr) {
#pragma unroll
            for (int ks = 0; ks < tc::CB / 16; ++ks) {
              const uint64_t ad = smem_desc(a0 + ai[pr] * tc::A_ARR + ks * 2 * tc::A_LBO, tc::A_LBO, 128);
              const uint64_t bd = smem_desc(b0 + bi[pr] * b_arr + ks * 2 * b_lbo, b_lbo, 128);
              umma_bf16(d_tmem, ad, bd, idesc, (cb | pr | ks) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[s]);  // stage reusable once these MMAs have read it
        }
        umma_commit(&t_full[buf]);  // accumulator complete
      }
    }
  } else if (warp >= tc::W_EPI0 && warp < tc::W_TR0) {
    // ---------------- epilogue: TMEM -> registers -> (+bias) -> global ----------------
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) belong to this warp
    const int m = quarter * 32 + lane;
    ptx::pdl_wait();  // bias comes from the prep kernel
    for (int lt = 0; lt < ntile_local; ++lt) {
      const int tile = blockIdx.x + lt * gridDim.x;
      const int nt = tile % p.NTL;
      const int mt = (tile / p.NTL) % p.MT;
      const int b = tile / (p.NTL * p.MT);
      const int buf = lt & 1;
      ptx::mbar_wait(&t_full[buf], (lt >> 1) & 1);
      tc_fence_after();
      const int t = mt * tc::M + m;
      const int n0 = nt * Nt;
      const float* bias = p.bias + (static_cast<size_t>(b) * p.NTL + nt) * Nt;
      float* orow = p.out + (static_cast<size_t>(b) * p.T_y + t) * p.T_x + n0;
      const bool vec_ok = (p.T_x & 3) == 0;
      const uint32_t taddr = tmem_base + static_cast<uint32_t>(buf) * 256u + (static_cast<uint32_t>(quarter * 32) << 16);
      for (int c = 0; c < Nt; c += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c, r);
        if (t < p.T_y) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bq = *reinterpret_cast<const float4*>(bias + c + 4 * q);
            float4 o;
            o.x = __uint_as_float(r[4 * q + 0]) + bq.x;
            o.y = __uint_as_float(r[4 * q + 1]) + bq.y;
            o.z = __uint_as_float(r[4 * q + 2]) + bq.z;
            o.w = __uint_as_float(r[4 * q + 3]) + bq.w;
            const int n = n0 + c + 4 * q;
            if (vec_ok && n + 3 < p.T_x) {
              *reinterpret_cast<float4*>(orow + c + 4 * q) = o;
            } else {
              if (n + 0 < p.T_x) orow[c + 4 * q + 0] = o.x;
              if (n + 1 < p.T_x) orow[c + 4 * q + 1] = o.y;
              if (n + 2 < p.T_x) orow[c + 4 * q + 2] = o.z;
              if (n + 3 < p.T_x) orow[c + 4 * q + 3] = o.w;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&t_empty[buf]);
    }
  } else if (warp >= tc::W_TR0) {
    // ---------------- A producers: z_p -> (-0.5 z^2, z) -> bf16 hi/lo operand tiles ----------------
    const int tt = tid - tc::W_TR0 * 32;  // 0..255
    const int m = tt & (tc::M - 1);
    const int half = tt >> 7;             // which 16 of the stage's 32 channels
    uint32_t it = 0;
    for (int lt = 0; lt < ntile_local; ++lt) {
      const int tile = blockIdx.x + lt * gridDim.x;
      const int mt = (tile / p.NTL) % p.MT;
      const int b = tile / (p.NTL * p.MT);
      const int t = mt * tc::M + m;
      const bool t_ok = t < p.T_y;
      const float* zb = p.z_p + static_cast<size_t>(b) * p.C * p.T_y + t;
      for (int cb = 0; cb < p.NCB; ++cb, ++it) {
        const int s = it & 1;
        float z[16];
        const int d0 = cb * tc::CB + half * 16;
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = (t_ok && d0 + i < p.C) ? zb[static_cast<size_t>(d0 + i) * p.T_y] : 0.0f;
        if (it >= 2) ptx::mbar_wait(&empty[s], ((it >> 1) - 1) & 1);  // MMAs of the previous use are done
        unsigned char* a_base = smem + s * stage_bytes;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t a2h[4], a2l[4], zh[4], zl[4];
#pragma unroll
          for (int i2 = 0; i2 < 4; ++i2) {
            __nv_bfloat16 h[2][4];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float zz = z[q * 8 + i2 * 2 + e];
              const float a2 = -0.5f * (zz * zz);  // :227
              split_bf16(a2, h[e][0], h[e][1]);
              split_bf16(zz, h[e][2], h[e][3]);
            }
            a2h[i2] = pack_bf16(h[0][0], h[1][0]);
            a2l[i2] = pack_bf16(h[0][1], h[1][1]);
            zh[i2] = pack_bf16(h[0][2], h[1][2]);
            zl[i2] = pack_bf16(h[0][3], h[1][3]);
          }
          const uint32_t off = static_cast<uint32_t>(half * 2 + q) * tc::A_LBO + static_cast<uint32_t>(m) * 16u;
          *reinterpret_cast<uint4*>(a_base + 0 * tc::A_ARR + off) = make_uint4(a2h[0], a2h[1], a2h[2], a2h[3]);
          *reinterpret_cast<uint4*>(a_base + 1 * tc::A_ARR + off) = make_uint4(a2l[0], a2l[1], a2l[2], a2l[3]);
          *reinterpret_cast<uint4*>(a_base + 2 * tc::A_ARR + off) = make_uint4(zh[0], zh[1], zh[2], zh[3]);
          *reinterpret_cast<uint4*>(a_base + 3 * tc::A_ARR + off) = make_uint4(zl[0], zl[1], zl[2], zl[3]);
        }
        fence_proxy_async();  // make the generic-proxy stores visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_full[s]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == tc::W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tc::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
struct TcShape {
  int Nt, NTL, MT, NCB;
  size_t bops_bytes, bias_bytes, total;
};

static TcShape tc_shape(int B, int C, int T_y, int T_x) {
  TcShape s;
  s.NTL = (T_x + 255) / 256;
  const int per = (T_x + s.NTL - 1) / s.NTL;
  s.Nt = ((per + 15) / 16) * 16;
  s.MT = (T_y + tc::M - 1) / tc::M;
  s.NCB = (C + tc::CB - 1) / tc::CB;
  s.bops_bytes = static_cast<size_t>(B) * s.NTL * s.NCB * s.Nt * 256;
  s.bias_bytes = static_cast<size_t>(B) * s.NTL * s.Nt * 4;
  s.total = ((s.bops_bytes + 255) & ~size_t(255)) + ((s.bias_bytes + 255) & ~size_t(255));
  return s;
}

size_t neg_cent_tc_scratch_bytes(int B, int C, int T_y, int T_x) { return tc_shape(B, C, T_y, T_x).total; }

int neg_cent_tc(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
                int B, int C, int T_y, int T_x, cudaStream_t st) {
  const TcShape s = tc_shape(B, C, T_y, T_x);
  if (scratch_bytes < s.total || !scratch) return MAS_E_SCRATCH;
  if (reinterpret_cast<uintptr_t>(scratch) & 15u) return MAS_E_ALIGN;
  unsigned char* bops = static_cast<unsigned char*>(scratch);
  float* bias = reinterpret_cast<float*>(bops + ((s.bops_bytes + 255) & ~size_t(255)));

  PrepParams pp{m_p, logs_p, bops, bias, C, T_x, s.Nt, s.NTL, s.NCB};
  neg_cent_prep_kernel<<<dim3(s.NCB, s.NTL, B), 256, 0, st>>>(pp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();

  TcParams tp{};
  tp.z_p = z_p; tp.out = out; tp.bops = bops; tp.bias = bias;
  tp.B = B; tp.C = C; tp.T_y = T_y; tp.T_x = T_x;
  tp.Nt = s.Nt; tp.NTL = s.NTL; tp.MT = s.MT; tp.NCB = s.NCB;
  tp.tiles = B * s.MT * s.NTL;
  const size_t smem = static_cast<size_t>(tc::STAGES) * (4 * tc::A_ARR + 4 * static_cast<size_t>(s.Nt) * 64) + 128;
  static bool attr = false;
  static int sms = 0;
  if (!attr) {
    e = cudaFuncSetAttribute(neg_cent_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return static_cast<int>(e);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tp.tiles < sms ? tp.tiles : sms);
  cfg.blockDim = dim3(tc::N_WARPS * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute la[1];
  la[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  la[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = la;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, neg_cent_tc_kernel, tp);
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return MAS_OK;
}

}  // namespace mas
