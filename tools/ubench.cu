// tools/ubench.cu -- tiny latency/throughput probes that the MAS forward kernel is designed against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../vits_b200/csrc/ptx_sm100.cuh"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 1;                                                                    \
    }                                                                              \
  } while (0)

__device__ __forceinline__ long long clk() { return clock64(); }

// (a) dependent SHFL.UP chain
__global__ void k_shfl(float* out, long long* cyc, int n) {
  float v = threadIdx.x * 1.5f;
  long long t0 = clk();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v = __shfl_up_sync(0xffffffffu, v, 1);
  }
  long long t1 = clk();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}

// (b) dependent FMNMX+FADD chain
__global__ void k_maxadd(float* out, long long* cyc, int n, float a, float b) {
  float v = threadIdx.x * 1.5f, w = a;
  long long t0 = clk();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) v = b + fmaxf(v, w);
  }
  long long t1 = clk();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}

// (c) the real row recurrence: shfl + fsel + K x (max, add, sub, shf); 1 warp, K columns per lane
template <int K>
__global__ void k_row(float* out, long long* cyc, int n, const float* __restrict__ c) {
  float v[K];
  uint32_t acc[K];
  for (int j = 0; j < K; ++j) { v[j] = threadIdx.x + j; acc[j] = 0; }
  const bool lane0 = threadIdx.x == 0;
  float cc[K];
  for (int j = 0; j < K; ++j) cc[j] = c[threadIdx.x * K + j];
  long long t0 = clk();
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float left = __shfl_up_sync(0xffffffffu, v[K - 1], 1);
      if (lane0) left = -1e9f;
#pragma unroll
      for (int j = K - 1; j >= 1; --j) {
        float d = v[j] - v[j - 1];
        acc[j] = __funnelshift_l(__float_as_uint(d), acc[j], 1);
        v[j] = cc[j] + fmaxf(v[j - 1], v[j]);
      }
      float d = v[0] - left;
      acc[0] = __funnelshift_l(__float_as_uint(d), acc[0], 1);
      v[0] = cc[0] + fmaxf(left, v[0]);
    }
  }
  long long t1 = clk();
  float s = 0;
  for (int j = 0; j < K; ++j) s += v[j] + acc[j];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}

// (c2) the same recurrence with the real kernel's data movement added step by step (K = 2):
//   MODE bit0: c values come from shared memory, loaded one 8-frame block ahead (float2 per row)
//   MODE bit1: lane 31 stores its last column to shared memory every frame (hand-off publish)
//   MODE bit2: lane 0's left edge comes from a shared-memory array (two float4 per block)
//   MODE bit3: decision bits flushed to shared memory every 32 frames
// nwarps warps run independent copies (SM-level contention of the MIO/LSU path).
template <int MODE>
__global__ void k_row2(float* out, long long* cyc, int nblk, const float* __restrict__ cg) {
  constexpr int K = 2;
  extern __shared__ __align__(16) float sm2[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* rows = sm2 + warp * (64 * 64 + 256 + 64);  // 64 frames x 64 floats
  float* edges = rows + 64 * 64;                    // 256 floats
  float* sink = edges + 256;
  for (int i = lane; i < 64 * 64; i += 32) rows[i] = cg[i & 63] * 0.001f;
  for (int i = lane; i < 256; i += 32) edges[i] = -1e9f;
  __syncwarp();
  float v[K];
  uint32_t acc[K];
  for (int j = 0; j < K; ++j) { v[j] = lane + j; acc[j] = 0; }
  const bool lane0 = lane == 0, lane31 = lane == 31;
  float cc[2][8][K], e[2][8];
  auto load_block = [&](int blk, float (&c)[8][K], float (&ee)[8]) {
    const float* rp = rows + (blk & 7) * 8 * 64 + lane * K;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE & 1) {
        const float2 t = *reinterpret_cast<const float2*>(rp + i * 64);
        c[i][0] = t.x; c[i][1] = t.y;
      } else {
        c[i][0] = 0.25f; c[i][1] = 0.5f;
      }
    }
    if (MODE & 4) {
      const float4 a = *reinterpret_cast<const float4*>(edges + (blk & 31) * 8);
      const float4 b4 = *reinterpret_cast<const float4*>(edges + (blk & 31) * 8 + 4);
      ee[0] = a.x; ee[1] = a.y; ee[2] = a.z; ee[3] = a.w; ee[4] = b4.x; ee[5] = b4.y; ee[6] = b4.z; ee[7] = b4.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ee[i] = -1e9f;
    }
  };
  load_block(0, cc[0], e[0]);
  long long t0 = clk();
#pragma unroll 1
  for (int blk = 0; blk < nblk; blk += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      load_block(blk + h + 1, cc[(h + 1) & 1], e[(h + 1) & 1]);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float left = __shfl_up_sync(0xffffffffu, v[K - 1], 1);
        if (lane0) left = e[h][r];
        const float d1 = v[1] - v[0];
        acc[1] = __funnelshift_l(__float_as_uint(d1), acc[1], 1);
        v[1] = cc[h][r][1] + fmaxf(v[0], v[1]);
        const float d0 = v[0] - left;
        acc[0] = __funnelshift_l(__float_as_uint(d0), acc[0], 1);
        v[0] = cc[h][r][0] + fmaxf(left, v[0]);
        if ((MODE & 2) && lane31) sink[(h * 8 + r) & 63] = v[K - 1];
      }
      if ((MODE & 8) && (((blk + h) & 3) == 3)) {
        *reinterpret_cast<uint2*>(rows + 63 * 64 + lane * 2) = make_uint2(acc[0], acc[1]);
      }
    }
  }
  long long t1 = clk();
  out[threadIdx.x] = v[0] + v[1] + acc[0] + acc[1];
  if (lane == 0) cyc[warp] = (t1 - t0);
}

// (d) mbarrier try_wait on an already completed phase; LDS latency; arrive cost
__global__ void k_mbar(long long* cyc, int n) {
  __shared__ uint64_t bar;
  __shared__ float buf[64];
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::mbar_fence_init();
  }
  buf[threadIdx.x] = threadIdx.x ^ 1;
  __syncthreads();
  if (threadIdx.x == 0) ptx::mbar_arrive(&bar);  // completes phase 0
  __syncthreads();
  long long t0 = clk();
  for (int i = 0; i < n; ++i) ptx::mbar_wait(&bar, 0);
  long long t1 = clk();
  int idx = threadIdx.x;
  for (int i = 0; i < n; ++i) idx = static_cast<int>(buf[idx & 31]);
  long long t2 = clk();
  bool ok = true;
  for (int i = 0; i < n; ++i) ok &= ptx::mbar_test(&bar, 0);
  long long t3 = clk();
  if (threadIdx.x == 0) {
    cyc[0] = (t1 - t0);
    cyc[1] = (t2 - t1);
    cyc[2] = (t3 - t2) + (ok ? 0 : 1) + (idx == 12345);
  }
}

// (e) one CTA streaming `total` bytes through an S-stage ring of `chunk`-byte bulk copies
__global__ void k_bulk(const unsigned char* src, long long* cyc, float* out, int chunk, int S, int nchunks) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + static_cast<size_t>(S) * chunk);
  uint64_t* empty = full + S;
  const unsigned char* mysrc = src + static_cast<size_t>(blockIdx.x) * chunk * nchunks;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_fence_init();
  }
  __syncthreads();
  long long t0 = clk();
  float accv = 0;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      for (int c = 0; c < nchunks; ++c) {
        int s = c % S;
        if (c >= S) ptx::mbar_wait(&empty[s], ((c / S) - 1) & 1);
        ptx::mbar_arrive_expect_tx(&full[s], chunk);
        ptx::bulk_g2s(sm + static_cast<size_t>(s) * chunk, mysrc + static_cast<size_t>(c) * chunk, chunk, &full[s]);
      }
    }
  } else if (threadIdx.x < 64) {
    for (int c = 0; c < nchunks; ++c) {
      int s = c % S;
      ptx::mbar_wait(&full[s], (c / S) & 1);
      accv += reinterpret_cast<float*>(sm + static_cast<size_t>(s) * chunk)[threadIdx.x];
      __syncwarp();
      if (threadIdx.x == 32) ptx::mbar_arrive(&empty[s]);
    }
  }
  long long t1 = clk();
  if (threadIdx.x == 32) cyc[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = accv;
}

int main() {
  float* out;
  long long* cyc;
  CK(cudaMalloc(&out, 1 << 20));
  CK(cudaMalloc(&cyc, 1024 * 8));
  long long h[1024];
  const int n = 1000;
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  printf("SM clock attr: %d kHz\n", clock_khz);

  for (int rep = 0; rep < 2; ++rep) {
    k_shfl<<<1, 32>>>(out, cyc, n);
    CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    if (rep) printf("SHFL.UP dependent chain: %.1f cycles each\n", h[0] / (16.0 * n));
    k_maxadd<<<1, 32>>>(out, cyc, n, 0.5f, 0.25f);
    CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));
    if (rep) printf("FMNMX+FADD dependent pair: %.1f cycles\n", h[0] / (16.0 * n));
  }
  float* c;
  CK(cudaMalloc(&c, 32 * 8 * 4));
  CK(cudaMemset(c, 0, 32 * 8 * 4));
#define ROW(K)                                                               \
  for (int rep = 0; rep < 2; ++rep) {                                        \
    k_row<K><<<1, 32>>>(out, cyc, n, c);                                     \
    CK(cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost));                       \
    if (rep) printf("row recurrence K=%d: %.1f cycles per row (1 warp, registers only)\n", K, h[0] / (8.0 * n)); \
  }
  ROW(1) ROW(2) ROW(3) ROW(4) ROW(6) ROW(8)
  {
    const int nblk = 1024;
    const size_t smem2 = 4 * (64 * 64 + 256 + 64) * sizeof(float);
#define ROW2(MODE, NW)                                                                         \
    CK(cudaFuncSetAttribute(k_row2<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
    for (int rep = 0; rep < 2; ++rep) {                                                        \
      k_row2<MODE><<<1, 32 * NW, smem2>>>(out, cyc, nblk, c);                                  \
      CK(cudaMemcpy(h, cyc, 8 * NW, cudaMemcpyDeviceToHost));                                  \
      if (rep) printf("row K=2 mode %2d warps %d: %.1f cycles/row\n", MODE, NW, h[0] / (8.0 * nblk)); \
    }
    ROW2(0, 1) ROW2(1, 1) ROW2(2, 1) ROW2(4, 1) ROW2(8, 1) ROW2(3, 1) ROW2(7, 1) ROW2(15, 1) ROW2(15, 3) ROW2(15, 4) ROW2(0, 3)
  }
  for (int rep = 0; rep < 2; ++rep) {
    k_mbar<<<1, 32>>>(cyc, n);
    CK(cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost));
    if (rep)
      printf("mbarrier try_wait (complete): %.1f cyc; dependent LDS: %.1f cyc; test_wait: %.1f cyc\n", h[0] / (double)n,
             h[1] / (double)n, h[2] / (double)n);
  }
  // bulk copy streaming: 1 CTA and 64 CTAs, various chunk sizes / depths
  unsigned char* src;
  const size_t per_cta = 768 * 1024;
  CK(cudaMalloc(&src, per_cta * 148));
  CK(cudaMemset(src, 1, per_cta * 148));
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  int chunks[] = {768, 3072, 6144, 12288, 24576, 49152};
  for (int ctas : {64}) {
    for (int chunk : chunks) {
      for (int S : {2, 4, 8}) {
        if (static_cast<size_t>(S) * chunk > 190 * 1024) continue;
        int nchunks = static_cast<int>(per_cta / chunk);
        size_t smem = static_cast<size_t>(S) * chunk + 2 * S * 8;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        k_bulk<<<ctas, 64, smem>>>(src, cyc, out, chunk, S, nchunks);
        cudaEventRecord(e0);
        k_bulk<<<ctas, 64, smem>>>(src, cyc, out, chunk, S, nchunks);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(h, cyc, 8 * ctas, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("bulk g2s ctas=%3d chunk=%6d S=%d: %8lld cycles/CTA  %.1f B/cyc/CTA  kernel %.1f us  (%.0f GB/s total)\n", ctas,
               chunk, S, mx, (double)per_cta / mx, ms * 1e3, per_cta * ctas / (ms * 1e-3) / 1e9);
      }
    }
  }
  return 0;
}
