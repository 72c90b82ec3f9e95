// tools/ubench_tma.cu -- streaming throughput of 2-D tiled TMA boxes vs 1-D bulk copies, as the MAS forward kernel issues them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_tma.bin tools/ubench_tma.cu && tools/ubench_tma.bin
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../vits_b200/csrc/ptx_sm100.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// W warps per CTA; warp w streams boxes [rows x cols floats] at column w*cols of utterance blockIdx.x, S slots in flight.
__global__ void k_tma2d(const __grid_constant__ CUtensorMap tm, int T_y, int rows, int cols, int S, float* out, long long* cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const uint32_t slotb = rows * cols * 4;
  unsigned char* ring = sm + (size_t)w * S * slotb;
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)W * S * slotb) + w * S;
  if (lane == 0) { for (int s = 0; s < S; ++s) ptx::mbar_init(&full[s], 1); ptx::mbar_fence_init(); }
  __syncwarp();
  const int nch = T_y / rows;
  long long t0 = clock64();
  float acc = 0.f;
  if (lane == 0)
    for (int c = 0; c < S && c < nch; ++c) { ptx::mbar_arrive_expect_tx(&full[c], slotb); ptx::tma_load_2d(ring + (size_t)c * slotb, &tm, w * cols, blockIdx.x * T_y + c * rows, &full[c]); }
  for (int c = 0; c < nch; ++c) {
    const int s = c % S;
    ptx::mbar_wait(&full[s], (c / S) & 1);
    acc += reinterpret_cast<float*>(ring + (size_t)s * slotb)[lane];
    __syncwarp();
    long long i0 = clock64();
    if (lane == 0 && c + S < nch) { ptx::mbar_arrive_expect_tx(&full[s], slotb); ptx::tma_load_2d(ring + (size_t)s * slotb, &tm, w * cols, blockIdx.x * T_y + (c + S) * rows, &full[s]); }
    __syncwarp();
    long long i1 = clock64();
    if (c == nch / 2 && blockIdx.x == 0 && lane == 0) cyc[512 + w] = i1 - i0;
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int B = 64, T_y = 1024, T_x = 192;
  float* src; float* out; long long* cyc;
  CK(cudaMalloc(&src, (size_t)B * T_y * T_x * 4 * 2));
  CK(cudaMemset(src, 0, (size_t)B * T_y * T_x * 4 * 2));
  CK(cudaMalloc(&out, 1 << 20)); CK(cudaMalloc(&cyc, 8 * 1024));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CK(cudaFuncSetAttribute(k_tma2d, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  struct Cfg { int W, rows, cols, S; };
  Cfg cfgs[] = {{3, 32, 64, 3}, {3, 32, 64, 5}, {3, 32, 64, 8}, {3, 64, 64, 4}, {3, 16, 64, 8}, {3, 8, 64, 16}, {1, 32, 192, 3}, {1, 32, 192, 6}, {1, 64, 192, 4}, {1, 16, 192, 12},
                {6, 32, 32, 8}, {2, 32, 96, 8}};
  for (int flip = 0; flip < 2; ++flip)
  for (auto& c : cfgs) {
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)T_x, (cuuint64_t)B * T_y}; cuuint64_t gstr[1] = {(cuuint64_t)T_x * 4};
    cuuint32_t box[2] = {(cuuint32_t)c.cols, (cuuint32_t)c.rows}; cuuint32_t es[2] = {1, 1};
    float* base = src + (flip ? (size_t)B * T_y * T_x : 0);   // alternate buffers so L2 does not help
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            flip ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    size_t smem = (size_t)c.W * c.S * c.rows * c.cols * 4 + c.W * c.S * 8 + 64;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_tma2d<<<B, 32 * c.W, smem>>>(tm, T_y, c.rows, c.cols, c.S, out, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hh[8]; cudaMemcpy(hh, cyc + 512, 64, cudaMemcpyDeviceToHost);
    printf("tma2d W=%d box=%2dx%3d S=%2d promo=%d: kernel %.1f us  (%.0f GB/s)  issue cost %lld cycles\n", c.W, c.rows, c.cols, c.S, flip ? 256 : 128, ms * 1e3, (double)B * T_y * T_x * 4 / (ms * 1e-3) / 1e9, hh[0]);
  }
  return 0;
}
