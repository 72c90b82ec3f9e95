// mas_common.cuh -- helpers shared by the forward kernels of the Monotonic Alignment Search
// (mas_forward.cuh: stage-granular kernel with fused backtrack; mas_dp.cuh: wavefront kernel).
#pragma once
#include <cstdint>
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

// Strided sums of the mask's column 0 (n_y elements, stride s_y) and row 0 (n_x elements, stride s_x) of one
// utterance.  The column walk touches one DRAM sector per element, so everything is latency: all loads of a
// thread (up to 16 + 8) are issued before the first one is consumed -- one round trip instead of one per batch
// of four (measured: the mask path cost 5-8 us per call before, ~1.5 us after).
template <typename T>
__device__ __forceinline__ void mask_sums_t(const T* base, int64_t s_y, int n_y, int64_t s_x, int n_x, int tid, int nthr,
                                            double& sy, double& sx) {
  constexpr int UY = 16, UX = 8;
  sy = 0.0;
  sx = 0.0;
  int iy = tid, ix = tid;
  while (iy < n_y || ix < n_x) {
    float vy[UY], vx[UX];
#pragma unroll
    for (int k = 0; k < UY; ++k) {
      const int i = iy + k * nthr;
      vy[k] = i < n_y ? static_cast<float>(base[i * s_y]) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < UX; ++k) {
      const int i = ix + k * nthr;
      vx[k] = i < n_x ? static_cast<float>(base[i * s_x]) : 0.0f;
    }
#pragma unroll
    for (int k = 0; k < UY; ++k) sy += static_cast<double>(vy[k]);
#pragma unroll
    for (int k = 0; k < UX; ++k) sx += static_cast<double>(vx[k]);
    iy += UY * nthr;
    ix += UX * nthr;
  }
}
__device__ __forceinline__ void mask_sums(const void* p, int dtype, int64_t off, int64_t s_y, int n_y, int64_t s_x, int n_x,
                                          int tid, int nthr, double& sy, double& sx) {
  switch (dtype) {
    case MAS_F32: mask_sums_t(static_cast<const float*>(p) + off, s_y, n_y, s_x, n_x, tid, nthr, sy, sx); break;
    case MAS_U8: mask_sums_t(static_cast<const uint8_t*>(p) + off, s_y, n_y, s_x, n_x, tid, nthr, sy, sx); break;
    default: {
      sy = 0.0;
      sx = 0.0;
      for (int i = tid; i < n_y; i += nthr) sy += mask_at(p, dtype, off + i * s_y);
      for (int i = tid; i < n_x; i += nthr) sx += mask_at(p, dtype, off + i * s_x);
    }
  }
}

// Sticky status bits: the word in the scratch (device) and, when there is one, the host-mapped mirror (one word
// per bit, plain stores: no PCIe atomics needed).  Error paths only.
__device__ __forceinline__ void raise_status(int32_t* word, int32_t* mirror, int bits) {
  atomicOr(word, bits);
  if (mirror) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((bits >> k) & 1) *reinterpret_cast<volatile int32_t*>(mirror + k) = 1;
    __threadfence_system();
  }
}

// Streaming hand-off of decision words between kernels: {word, tag} travels as ONE 64-bit scalar (word in the low
// half), so a reader that sees the tag also sees the word -- a 64-bit aligned scalar access is single-copy atomic in
// the PTX memory model, and a vector store/load is a sequence of such scalars (a 2 x 32-bit vector would formally be
// two independent 32-bit accesses).
__device__ __forceinline__ unsigned long long pack_tagged(uint32_t word, uint32_t tag) {
  return static_cast<unsigned long long>(word) | (static_cast<unsigned long long>(tag) << 32);
}
__device__ __forceinline__ uint2 load_tagged(const uint2* p) {
  const unsigned long long v = __ldcg(reinterpret_cast<const unsigned long long*>(p));
  return make_uint2(static_cast<uint32_t>(v), static_cast<uint32_t>(v >> 32));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// timeline slots: 0 dp first start, 1 dp last frame done, 2 dp last end, 3 backtrack first start,
//                 4 backtrack last end, 5 fill first start, 6 fill last chunk done, 7 fill last end
__device__ __forceinline__ void tl_min(unsigned long long* tl, int slot) {
  if (tl) atomicMin(tl + slot, globaltimer_ns());
}
__device__ __forceinline__ void tl_max(unsigned long long* tl, int slot) {
  if (tl) atomicMax(tl + slot, globaltimer_ns());
}

// One backtrack step (core.pyx:32-33) given the decision word of the current column.  The
// forward kernel already folded `index == y` (bit forced to 1) and `index != 0` (column 0
// forced to 0) into the stored bits.
__device__ __forceinline__ int bt_step(int cur, int r, uint32_t word) {
  return cur - static_cast<int>((word >> (31 - r)) & 1u);
}

}  // namespace mas
