// Internal C++ declarations shared by the translation units of libvits_mas.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include <atomic>

namespace mas {

void count_launch();

// ---- per-device host state (several GPUs may be driven from one process) ----
constexpr int kMaxDevices = 64;
inline int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: `done` (one static per call site) remembers
// the devices it was applied on.  Never called for the first time during a stream capture in practice (callers
// warm up eagerly), and legal there anyway.
template <typename Kern>
inline cudaError_t ensure_dyn_smem(Kern kern, int bytes, std::atomic<uint64_t>& done) {
  const uint64_t bit = 1ull << current_device();
  if (done.load(std::memory_order_relaxed) & bit) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_relaxed);
  return e;
}
int num_sms();  // of the current device (cached per device)
int fused_backtrack_sms(int B, int T_y, int T_x);  // mas_path.cu
int last_fail_line();  // mas_path.cu line of the last failing CUDA call on this thread (0 = none)
// Host-mapped status mirror of the current device: 4 int32 words, one per MAS_STATUS_* bit, that the kernels set
// to 1 besides the sticky word in the scratch, so the host can poll for errors without synchronising.  nullptr if
// pinned memory could not be allocated (the mirror is then simply not written).
int32_t* status_mirror();

// mas_path.cu
// Streamed source of neg_cent (mas_fused.cu): instead of a dense tensor, a per-utterance ring of RT tiles of 128
// frames (row pitch `pitch` floats, a multiple of 4) filled in frame order by the contraction kernel running
// concurrently; tile_flags[b * MT + mt] reaches tile_need when frame block mt is complete.  fill_done[b * fill_stride]
// is raised by the DP CTA once it has zero-filled utterance b's dense path.
struct FusedSrc {
  const float* ring;
  int pitch, RT, MT;
  const uint32_t* tile_flags;
  int tile_need;
  uint32_t* fill_done;
  int fill_stride;
  int reserve_ctas;  // SMs the contraction kernel occupies (the search must fit on the rest)
};
int maximum_path(const float* neg_cent, const int32_t* t_ys, const int32_t* t_xs, const void* mask, int mask_dtype,
                 int64_t msb, int64_t msy, int64_t msx, void* path_out, int path_dtype, int32_t* index_out,
                 void* scratch, size_t scratch_bytes, int B, int T_y, int T_x, cudaStream_t st,
                 const FusedSrc* fused = nullptr, bool probe_only = false);  // probe_only: validate + choose, launch nothing
// mas_fused.cu
size_t stats_to_path_scratch_bytes(int B, int C, int T_y, int T_x);
int stats_to_path(const float* z_p, const float* m_p, const float* logs_p, const int32_t* t_ys, const int32_t* t_xs,
                  void* path_out, int path_dtype, int32_t* index_out, void* scratch, size_t scratch_bytes, int B, int C,
                  int T_y, int T_x, cudaStream_t st);
size_t maximum_path_scratch_bytes(int B, int T_y, int T_x);
void set_tuning(int K, int R, int S, int pdl);
void set_debug_kernels(int mask);
void set_tuning2(int fused, int helpers);
void set_tuning3(int wavefront, int ring_mode, int ring_slots, int cols_per_lane);
void set_timeline(unsigned long long* dev_ptr);
unsigned long long* timeline_ptr();
void set_trace(unsigned long long* dev_ptr);

// mas_consumers.cu
int path_durations(const int32_t* index, float* w, int B, int T_y, int T_x, cudaStream_t st);
int expand_prior(const int32_t* index, const float* m_p, const float* logs_p, float* m_out, float* logs_out, int B, int C,
                 int T_y, int T_x, cudaStream_t st);
int generate_path(const float* duration, const float* mask, int64_t msb, int64_t msy, int64_t msx, float* path, int B,
                  int T_y, int T_x, cudaStream_t st);

int kl_from_index(const int32_t* index, const float* z_p, const float* logs_q, const float* m_p, const float* logs_p,
                  const float* z_mask, double* out, int B, int C, int T_y, int T_x, cudaStream_t st);

// mas_neg_cent.cu
int neg_cent(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
             int B, int C, int T_y, int T_x, cudaStream_t st);
size_t neg_cent_scratch_bytes(int B, int C, int T_y, int T_x);
void set_neg_cent_impl(int impl);
int neg_cent_autocast(const float* z_p, const float* m_p, const float* logs_p, float* out, int gemm_dtype, int stats_lowp,
                      int B, int C, int T_y, int T_x, cudaStream_t st);

// mas_neg_cent_tc.cu
int neg_cent_tc(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
                int B, int C, int T_y, int T_x, cudaStream_t st);
size_t neg_cent_tc_scratch_bytes(int B, int C, int T_y, int T_x);
// Streamed mode of the tensor-core contraction (mas_fused.cu): frame-major tile order, output into a per-utterance
// ring of RT tiles of 128 frames with row pitch `pitch` floats, flags[b][mt] += 1 per finished column tile; `grid`
// CTAs; the prep kernel clears zero_base[B][zero_stride] words first.
struct TcStream {
  float* ring;
  uint32_t* flags;
  const int32_t* t_ys;
  const int32_t* t_xs;
  int RT, pitch, grid;
  int n_long, n1;  // CTAs that stay for the whole job / tiles every CTA takes before the others leave (see TcParams)
  uint32_t* zero_base;
  int zero_stride;
};
int neg_cent_tc_impl(const float* z_p, const float* m_p, const float* logs_p, float* out, void* scratch, size_t scratch_bytes,
                     int B, int C, int T_y, int T_x, cudaStream_t st, const TcStream* so);
void neg_cent_tc_dims(int C, int T_y, int T_x, int* Nt, int* NTL, int* MT);

}  // namespace mas
