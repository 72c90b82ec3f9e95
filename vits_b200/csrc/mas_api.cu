// mas_api.cu -- the extern "C" surface declared in include/vits_mas.h, plus the host-buffer
// entry that mirrors the reference's native call maximum_path_c (monotonic_align/core.pyx:38).
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"

namespace mas {
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// One pinned, portable, device-mapped allocation for every device: [kMaxDevices][4] int32, word k of a device's
// row = "MAS_STATUS bit k was raised".  With unified addressing the host pointer is valid on every device.
static std::atomic<int32_t*> g_mirror{nullptr};
static std::atomic<int> g_mirror_state{0};  // 0 untried, 1 ready, 2 unavailable
static int32_t* mirror_base() {
  int st = g_mirror_state.load(std::memory_order_acquire);
  if (st == 0) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    st = g_mirror_state.load(std::memory_order_acquire);
    if (st == 0) {
      void* p = nullptr;
      const cudaError_t e = cudaHostAlloc(&p, kMaxDevices * 4 * sizeof(int32_t), cudaHostAllocPortable | cudaHostAllocMapped);
      if (e == cudaSuccess && p) {
        memset(p, 0, kMaxDevices * 4 * sizeof(int32_t));
        g_mirror.store(static_cast<int32_t*>(p), std::memory_order_release);
        st = 1;
      } else {
        cudaGetLastError();  // (e.g. inside a stream capture: try again on a later call)
        return nullptr;
      }
      g_mirror_state.store(st, std::memory_order_release);
    }
  }
  return st == 1 ? g_mirror.load(std::memory_order_acquire) : nullptr;
}
int32_t* status_mirror() {
  int32_t* base = mirror_base();
  return base ? base + 4 * current_device() : nullptr;
}
}  // namespace mas

namespace {

// State cached by the host entry, one per device (single caller thread per device, like the reference).
struct HostCtx {
  static constexpr int kChunks = 32;  // most utterance groups pipelined over PCIe
  static constexpr int kStreams = 4;
  cudaStream_t streams[kStreams] = {nullptr, nullptr, nullptr, nullptr};
  void* d_values = nullptr;
  void* d_paths = nullptr;
  void* d_index = nullptr;      // [B][T_y] int32 per-frame text position (index-only pipeline)
  int32_t* h_index = nullptr;   // pinned host copy of it
  size_t cap_index = 0;
  float* h_stage = nullptr;     // pinned mirror of a PAGEABLE `values` (filled by the host threads)
  size_t cap_stage = 0;
  void* d_lens = nullptr;
  void* d_scratch = nullptr;
  cudaEvent_t ev_in[kChunks] = {}, ev_k[kChunks] = {}, ev_idx[kChunks] = {};  // group c: inputs landed / kernels done / index on the host
  int32_t* h_status = nullptr;  // pinned, kChunks words
  size_t cap_cells = 0, cap_path_bytes = 0, cap_lens = 0, cap_scratch = 0;
  bool ready = false;
};
HostCtx g_hosts[mas::kMaxDevices];
#define g_host (g_hosts[mas::current_device()])

void host_release() {
  if (g_host.d_values) cudaFree(g_host.d_values);
  if (g_host.d_paths) cudaFree(g_host.d_paths);
  if (g_host.d_index) cudaFree(g_host.d_index);
  if (g_host.h_index) cudaFreeHost(g_host.h_index);
  if (g_host.h_stage) cudaFreeHost(g_host.h_stage);
  if (g_host.d_lens) cudaFree(g_host.d_lens);
  if (g_host.d_scratch) cudaFree(g_host.d_scratch);
  if (g_host.h_status) cudaFreeHost(g_host.h_status);
  for (auto& s : g_host.streams)
    if (s) cudaStreamDestroy(s);
  for (auto& e : g_host.ev_in)
    if (e) cudaEventDestroy(e);
  for (auto& e : g_host.ev_k)
    if (e) cudaEventDestroy(e);
  for (auto& e : g_host.ev_idx)
    if (e) cudaEventDestroy(e);
  g_host = HostCtx{};
}

#define MAS_CUDA(x)                                \
  do {                                             \
    cudaError_t e_ = (x);                          \
    if (e_ != cudaSuccess) return static_cast<int>(e_); \
  } while (0)

int host_prepare(size_t cells, size_t lens, size_t scratch, int path_es, size_t index_words) {
  if (!g_host.ready) {
    for (auto& s : g_host.streams) MAS_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    for (auto& e : g_host.ev_in) MAS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : g_host.ev_k) MAS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : g_host.ev_idx) MAS_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    MAS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_host.h_status), HostCtx::kChunks * sizeof(int32_t),
                           cudaHostAllocDefault));
    g_host.ready = true;
  }
  if (cells > g_host.cap_cells) {
    if (g_host.d_values) cudaFree(g_host.d_values);
    g_host.d_values = nullptr;
    g_host.cap_cells = 0;
    MAS_CUDA(cudaMalloc(&g_host.d_values, cells * 4));
    g_host.cap_cells = cells;
  }
  if (index_words > g_host.cap_index) {
    if (g_host.d_index) cudaFree(g_host.d_index);
    if (g_host.h_index) cudaFreeHost(g_host.h_index);
    g_host.d_index = nullptr;
    g_host.h_index = nullptr;
    g_host.cap_index = 0;
    MAS_CUDA(cudaMalloc(&g_host.d_index, index_words * sizeof(int32_t)));
    MAS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_host.h_index), index_words * sizeof(int32_t), cudaHostAllocDefault));
    g_host.cap_index = index_words;
  }
  if (cells * path_es > g_host.cap_path_bytes) {
    if (g_host.d_paths) cudaFree(g_host.d_paths);
    g_host.d_paths = nullptr;
    g_host.cap_path_bytes = 0;
    MAS_CUDA(cudaMalloc(&g_host.d_paths, cells * path_es));
    g_host.cap_path_bytes = cells * path_es;
  }
  if (lens > g_host.cap_lens) {
    if (g_host.d_lens) cudaFree(g_host.d_lens);
    g_host.d_lens = nullptr;
    g_host.cap_lens = 0;
    MAS_CUDA(cudaMalloc(&g_host.d_lens, lens * 2 * sizeof(int32_t)));
    g_host.cap_lens = lens;
  }
  if (scratch > g_host.cap_scratch) {
    if (g_host.d_scratch) cudaFree(g_host.d_scratch);
    g_host.d_scratch = nullptr;
    g_host.cap_scratch = 0;
    MAS_CUDA(cudaMalloc(&g_host.d_scratch, scratch));
    g_host.cap_scratch = scratch;
  }
  return MAS_OK;
}

}  // namespace

extern "C" {

int mas_abi_version(void) { return 2; }

int32_t* mas_status_mirror(void) { return mas::status_mirror(); }
int mas_last_error_site(void) { return mas::last_fail_line(); }

const char* mas_error_string(int code) {
  switch (code) {
    case MAS_OK: return "ok";
    case MAS_E_BAD_SHAPE: return "bad shape (B, T_y, T_x must be positive; T_x <= 2048)";
    case MAS_E_BAD_DTYPE: return "unknown element-type code";
    case MAS_E_NULL: return "required pointer is NULL";
    case MAS_E_SCRATCH: return "scratch buffer too small";
    case MAS_E_ALIGN: return "misaligned pointer";
    case MAS_E_UNSUPPORTED: return "no kernel for this configuration";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

size_t mas_maximum_path_scratch_bytes(int B, int T_y, int T_x) { return mas::maximum_path_scratch_bytes(B, T_y, T_x); }
size_t mas_scratch_status_offset(void) { return 0; }

int mas_maximum_path(const float* neg_cent, const int32_t* t_ys, const int32_t* t_xs, const void* mask,
                     int mask_dtype, int64_t mask_sb, int64_t mask_sy, int64_t mask_sx, void* path_out,
                     int path_dtype, int32_t* index_out, void* scratch, size_t scratch_bytes, int B, int T_y,
                     int T_x, mas_stream_t stream) {
  return mas::maximum_path(neg_cent, t_ys, t_xs, mask, mask_dtype, mask_sb, mask_sy, mask_sx, path_out, path_dtype,
                           index_out, scratch, scratch_bytes, B, T_y, T_x, static_cast<cudaStream_t>(stream));
}

// Copies the leading rows of utterances [b0, b0+nb) (4-byte cells, planes of T_y*T_x) up to the group's
// longest utterance, as ONE strided copy: the reference never touches a row at or beyond an utterance's
// length (core.pyx:13-33 loops over t_y rows of a caller-zeroed `paths`), so the padded tail is neither
// sent nor read back.  Batches arrive sorted by length (TextAudioSpeakerCollate.py:26-30), so a group's
// utterances are about equally long.  (One copy per utterance moved fewer bytes but was slower: 128 small
// copies per call cost more host time than the PCIe time they saved.)
static cudaError_t copy_leading_rows(void* dst, const void* src, const int32_t* t_ys, int b0, int nb, int T_y, int T_x,
                                     cudaMemcpyKind kind, cudaStream_t st, int es = 4) {
  const size_t plane = static_cast<size_t>(T_y) * T_x * es;
  int rows = 0;
  for (int b = b0; b < b0 + nb; ++b) rows = t_ys[b] > rows ? t_ys[b] : rows;
  rows = rows > T_y ? T_y : rows;
  if (rows <= 0) return cudaSuccess;
  char* d = static_cast<char*>(dst) + plane * b0;
  const char* s = static_cast<const char*>(src) + plane * b0;
  if (rows == T_y) return cudaMemcpyAsync(d, s, plane * nb, kind, st);
  return cudaMemcpy2DAsync(d, plane, s, plane, static_cast<size_t>(rows) * T_x * es, nb, kind, st);
}

static int path_elem_size(int dtype) {
  switch (dtype) {
    case MAS_F32: case MAS_I32: return 4;
    case MAS_F16: case MAS_BF16: case MAS_I16: return 2;
    case MAS_F64: case MAS_I64: return 8;
    case MAS_U8: case MAS_I8: return 1;
    default: return 0;
  }
}

static int host_run(void* paths, int path_dtype, int zero_tail, const float* values, const int32_t* t_ys,
                    const int32_t* t_xs, int B, int T_y, int T_x);

int mas_maximum_path_c_host(int32_t* paths, const float* values, const int32_t* t_ys, const int32_t* t_xs, int B,
                            int T_y, int T_x) {
  return host_run(paths, MAS_I32, 0, values, t_ys, t_xs, B, T_y, T_x);
}

int mas_maximum_path_host(void* paths, int path_dtype, int zero_tail, const float* values, const int32_t* t_ys,
                          const int32_t* t_xs, int B, int T_y, int T_x) {
  return host_run(paths, path_dtype, zero_tail, values, t_ys, t_xs, B, T_y, T_x);
}

// ---- host-side materialisation of the dense path from the per-frame index -------------------------------------
// The dense [B,T_y,T_x] path is one 1 per frame: 4 bytes of information per T_x*4 bytes of tensor.  The host entry
// therefore brings back only the int32 index (0.26 MB at c2 instead of 42.5 MB) and writes the rows -- zeros and the
// one -- into the caller's buffer with a small pool of host threads while the next groups' inputs are still crossing
// the link: PCIe then carries the input only (the link, not the GPU, bounds this entry).
namespace {
struct ExpandJob {
  unsigned char* paths;    // caller's buffer
  const int32_t* index;    // pinned host copy of the device index, [B][T_y]
  const int32_t* t_ys;
  const int32_t* t_xs;
  int T_y, T_x, es, zero_tail;
  unsigned long long one;
  // pageable `values`: utterance b's first stage_rows[b] rows are copied into the pinned mirror by the pool
  const float* values = nullptr;
  float* stage = nullptr;
  const int* stage_rows = nullptr;  // [B] rows to stage (the longest utterance of b's group)
  const int* group_of = nullptr;    // [B]
};

void expand_utterance(const ExpandJob& j, int b) {
  const size_t plane = static_cast<size_t>(j.T_y) * j.T_x * j.es;
  unsigned char* pb = j.paths + plane * b;
  const int ty = j.t_ys[b], tx = j.t_xs[b];
  const bool valid = ty >= 1 && tx >= 1 && ty <= j.T_y && tx <= j.T_x && tx <= ty;  // as the kernels decide (mas_dp.cuh)
  const int rows = ty < 0 ? 0 : (ty > j.T_y ? j.T_y : ty);
  // rows below t_y are written in full; rows at or beyond it stay as the caller zeroed them (core.pyx never touches
  // them), unless the caller asked for them to be zeroed here
  memset(pb, 0, j.zero_tail ? plane : static_cast<size_t>(rows) * j.T_x * j.es);
  if (!valid) return;
  const int32_t* idx = j.index + static_cast<size_t>(b) * j.T_y;
  for (int y = 0; y < rows; ++y) {
    const int x = idx[y];
    if (x < 0 || x >= j.T_x) continue;
    unsigned char* cell = pb + (static_cast<size_t>(y) * j.T_x + x) * j.es;
    switch (j.es) {
      case 1: *cell = static_cast<unsigned char>(j.one); break;
      case 2: *reinterpret_cast<uint16_t*>(cell) = static_cast<uint16_t>(j.one); break;
      case 4: *reinterpret_cast<uint32_t*>(cell) = static_cast<uint32_t>(j.one); break;
      default: *reinterpret_cast<unsigned long long*>(cell) = j.one; break;
    }
  }
}

void stage_utterance(const ExpandJob& j, int b) {
  const size_t plane = static_cast<size_t>(j.T_y) * j.T_x;
  const int rows = j.stage_rows[b];
  if (rows > 0) memcpy(j.stage + plane * b, j.values + plane * b, static_cast<size_t>(rows) * j.T_x * sizeof(float));
}

// Persistent worker pool (created on first use; the calling thread works too).  A task is an utterance: t >= 0 =
// materialise the dense path rows of utterance t from its index; t < 0 = copy the leading rows of utterance -1-t of a
// PAGEABLE `values` into the pinned mirror (cudaMemcpyAsync from pageable memory stages through one driver thread at
// ~10 GB/s; eight threads of ours reach the host's memory bandwidth, and the DMA then runs at the pinned rate).
class ExpandPool {
 public:
  static ExpandPool& get() {
    static ExpandPool pool;
    return pool;
  }
  int threads() const { return static_cast<int>(workers_.size()) + 1; }
  void begin(const ExpandJob& job, int groups) {
    std::lock_guard<std::mutex> lk(m_);
    job_ = job;
    pending_.clear();
    head_ = 0;
    outstanding_ = 0;
    stage_left_.assign(groups, 0);
  }
  void add(int b0, int nb) {  // path rows of utterances [b0, b0+nb)
    {
      std::lock_guard<std::mutex> lk(m_);
      for (int b = b0; b < b0 + nb; ++b) pending_.push_back(b);
      outstanding_ += nb;
    }
    cv_.notify_all();
  }
  void add_staging(int c, int b0, int nb) {  // leading rows of group c's utterances into the pinned mirror
    {
      std::lock_guard<std::mutex> lk(m_);
      for (int b = b0; b < b0 + nb; ++b) pending_.push_back(-1 - b);
      outstanding_ += nb;
      stage_left_[c] += nb;
    }
    cv_.notify_all();
  }
  void wait_staged(int c) {  // the caller helps until group c's rows are in the mirror
    for (;;) {
      int t;
      {
        std::unique_lock<std::mutex> lk(m_);
        if (stage_left_[c] == 0) return;
        if (head_ < pending_.size()) t = pending_[head_++];
        else {
          done_cv_.wait(lk, [&] { return stage_left_[c] == 0; });
          return;
        }
      }
      run(t);
    }
  }
  void finish() {  // the caller helps, then waits for the stragglers
    for (;;) {
      int t;
      {
        std::unique_lock<std::mutex> lk(m_);
        if (head_ < pending_.size()) t = pending_[head_++];
        else {
          done_cv_.wait(lk, [&] { return outstanding_ == 0; });
          return;
        }
      }
      run(t);
    }
  }

 private:
  ExpandPool() {
    int n = 0;
    if (const char* e = getenv("MAS_HOST_THREADS")) n = atoi(e);
    if (n <= 0) {
      const unsigned hw = std::thread::hardware_concurrency();
      n = hw >= 16 ? 8 : (hw >= 4 ? static_cast<int>(hw) / 2 : 1);
    }
    if (n > 32) n = 32;
    for (int i = 1; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~ExpandPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  void run(int t) {
    if (t >= 0) expand_utterance(job_, t);
    else stage_utterance(job_, -1 - t);
    std::lock_guard<std::mutex> lk(m_);
    bool wake = --outstanding_ == 0;
    if (t < 0 && --stage_left_[job_.group_of[-1 - t]] == 0) wake = true;
    if (wake) done_cv_.notify_all();
  }
  void loop() {
    for (;;) {
      int t;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || head_ < pending_.size(); });
        if (stop_) return;
        t = pending_[head_++];
      }
      run(t);
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  std::vector<int> pending_;
  std::vector<int> stage_left_;
  size_t head_ = 0;
  int outstanding_ = 0;
  bool stop_ = false;
  ExpandJob job_{};
};

unsigned long long one_pattern(int dtype) {
  switch (dtype) {
    case MAS_F32: return 0x3F800000ull;
    case MAS_F16: return 0x3C00ull;
    case MAS_BF16: return 0x3F80ull;
    case MAS_F64: return 0x3FF0000000000000ull;
    default: return 1ull;
  }
}
}  // namespace

static int host_run(void* paths, int path_dtype, int zero_tail, const float* values, const int32_t* t_ys,
                    const int32_t* t_xs, int B, int T_y, int T_x) {
  if (B <= 0 || T_y <= 0 || T_x <= 0) return MAS_E_BAD_SHAPE;
  if (!paths || !values || !t_ys || !t_xs) return MAS_E_NULL;
  const int es = path_elem_size(path_dtype);
  if (es == 0) return MAS_E_BAD_DTYPE;
  const size_t plane = static_cast<size_t>(T_y) * T_x;
  // Groups of about 12 MB: enough of them to overlap the copies, the kernels and the host-side materialisation, few
  // enough that the ~10 driver calls per group stay off the critical path.
  static const int forced_groups = [] {
    const char* e = getenv("MAS_HOST_GROUPS");  // tuning hook
    const int v = e ? atoi(e) : 0;
    return v > 0 && v <= HostCtx::kChunks - 2 ? v : 0;
  }();
  // MAS_HOST_DENSE=1: the round-1 pipeline (the dense path crosses the link too) -- kept for A/B measurements
  static const bool dense_d2h = getenv("MAS_HOST_DENSE") && atoi(getenv("MAS_HOST_DENSE")) != 0;
  int want_chunks = forced_groups;
  if (!want_chunks) {
    const size_t total = plane * B * 4, target = size_t(12) << 20;
    want_chunks = static_cast<int>((total + target / 2) / target);
    want_chunks = want_chunks < 2 ? 2 : (want_chunks > 16 ? 16 : want_chunks);
  }
  const int nch = B < want_chunks ? B : want_chunks;
  const int per = (B + nch - 1) / nch;
  // Tapered groups: nothing overlaps the first group's inbound copy nor the last group's kernels and
  // materialisation, so the first and last group are a quarter of the nominal size (c2: 4, 12, 16, 16, 12, 4).
  int gsize[HostCtx::kChunks];
  int ng = 0;
  {
    const int q = per >= 4 ? per / 4 : 0;
    int left = B;
    auto push = [&](int n) {
      n = n < left ? n : left;
      if (n > 0) {
        gsize[ng++] = n;
        left -= n;
      }
    };
    if (q && !getenv("MAS_HOST_FLAT")) {
      push(q);
      push(per - q);
      while (left > per) push(per);
      push(left - q);
      push(q);
    }
    while (left > 0) push(per);
  }
  const size_t sc_one = (mas::maximum_path_scratch_bytes(per, T_y, T_x) + 255) & ~size_t(255);
  if (sc_one == 0) return MAS_E_BAD_SHAPE;
  int rc = host_prepare(plane * B, static_cast<size_t>(B), sc_one * ng, dense_d2h ? es : 0,
                        dense_d2h ? 0 : static_cast<size_t>(B) * T_y);
  if (rc != MAS_OK) return rc;

  float* d_values = static_cast<float*>(g_host.d_values);
  unsigned char* d_paths = static_cast<unsigned char*>(g_host.d_paths);
  int32_t* d_index = static_cast<int32_t*>(g_host.d_index);
  int32_t* h_index = g_host.h_index;
  int32_t* d_ty = static_cast<int32_t*>(g_host.d_lens);
  int32_t* d_tx = d_ty + B;
  // One stream feeds the inputs group by group without ever waiting for anything (H2D copy engine busy from the
  // first byte to the last), two streams run the kernels of alternate groups, one stream returns the results (D2H
  // copy engine); events hand each group from stage to stage.
  cudaStream_t s_in = g_host.streams[0], s_out = g_host.streams[3];
  MAS_CUDA(cudaMemcpyAsync(d_ty, t_ys, B * sizeof(int32_t), cudaMemcpyHostToDevice, s_in));
  MAS_CUDA(cudaMemcpyAsync(d_tx, t_xs, B * sizeof(int32_t), cudaMemcpyHostToDevice, s_in));
  MAS_CUDA(cudaMemsetAsync(g_host.d_scratch, 0, sc_one * ng, s_in));
  ExpandPool* pool = dense_d2h ? nullptr : &ExpandPool::get();
  // A pageable `values` (a plain numpy array, as the reference's wrapper passes: __init__.py:14) is copied into a pinned
  // mirror by the host threads, group by group ahead of the DMA: cudaMemcpyAsync from pageable memory goes through one
  // staging thread of the driver (python API with CPU tensors: 10.4 k alignments/s); MAS_HOST_NOSTAGE=1 restores that.
  static const bool no_stage = getenv("MAS_HOST_NOSTAGE") && atoi(getenv("MAS_HOST_NOSTAGE")) != 0;
  bool staged = false;
  if (pool && !no_stage) {
    cudaPointerAttributes pa{};
    if (cudaPointerGetAttributes(&pa, values) != cudaSuccess) cudaGetLastError();
    else staged = pa.type == cudaMemoryTypeUnregistered;
  }
  std::vector<int> stage_rows, group_of;
  if (staged) {
    if (plane * B > g_host.cap_stage) {
      if (g_host.h_stage) cudaFreeHost(g_host.h_stage);
      g_host.h_stage = nullptr;
      g_host.cap_stage = 0;
      if (cudaHostAlloc(reinterpret_cast<void**>(&g_host.h_stage), plane * B * sizeof(float), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        staged = false;  // no pinned memory to spare: the driver's own staging it is
      } else {
        g_host.cap_stage = plane * B;
      }
    }
  }
  if (staged) {
    stage_rows.resize(B);
    group_of.resize(B);
    for (int c = 0, b0 = 0; c < ng; b0 += gsize[c], ++c) {
      int rows = 0;
      for (int b = b0; b < b0 + gsize[c]; ++b) rows = t_ys[b] > rows ? t_ys[b] : rows;
      rows = rows > T_y ? T_y : (rows < 0 ? 0 : rows);
      for (int b = b0; b < b0 + gsize[c]; ++b) {
        stage_rows[b] = rows;
        group_of[b] = c;
      }
    }
  }
  if (pool) {
    ExpandJob job{static_cast<unsigned char*>(paths), h_index, t_ys, t_xs, T_y, T_x, es, zero_tail, one_pattern(path_dtype)};
    if (staged) {
      job.values = values;
      job.stage = g_host.h_stage;
      job.stage_rows = stage_rows.data();
      job.group_of = group_of.data();
    }
    pool->begin(job, ng);
    if (staged)
      for (int c = 0, b0 = 0; c < ng; b0 += gsize[c], ++c) pool->add_staging(c, b0, gsize[c]);
  }
  const float* h_values = staged ? g_host.h_stage : values;
  struct Drain {  // no early return may leave workers behind that still read this call's arrays
    ExpandPool* p;
    ~Drain() {
      if (p) p->finish();
    }
  } drain{pool};
  int nused = 0;
  for (int c = 0, b0 = 0; c < ng; b0 += gsize[c], ++c) {
    const int nb = gsize[c];
    cudaStream_t s_k = g_host.streams[1 + (c & 1)];
    unsigned char* sc = static_cast<unsigned char*>(g_host.d_scratch) + sc_one * c;
    if (staged) pool->wait_staged(c);
    MAS_CUDA(copy_leading_rows(d_values, h_values, t_ys, b0, nb, T_y, T_x, cudaMemcpyHostToDevice, s_in));
    MAS_CUDA(cudaEventRecord(g_host.ev_in[c], s_in));
    MAS_CUDA(cudaStreamWaitEvent(s_k, g_host.ev_in[c], 0));
    if (dense_d2h)
      rc = mas::maximum_path(d_values + plane * b0, d_ty + b0, d_tx + b0, nullptr, 0, 0, 0, 0, d_paths + plane * b0 * es,
                             path_dtype, nullptr, sc, sc_one, nb, T_y, T_x, s_k);
    else  // index only: no dense path is written on the device at all
      rc = mas::maximum_path(d_values + plane * b0, d_ty + b0, d_tx + b0, nullptr, 0, 0, 0, 0, nullptr, MAS_I32,
                             d_index + static_cast<size_t>(b0) * T_y, sc, sc_one, nb, T_y, T_x, s_k);
    if (rc != MAS_OK) {
      for (auto& s : g_host.streams) cudaStreamSynchronize(s);
      if (pool) pool->finish();
      return rc;
    }
    MAS_CUDA(cudaEventRecord(g_host.ev_k[c], s_k));
    MAS_CUDA(cudaStreamWaitEvent(s_out, g_host.ev_k[c], 0));
    if (dense_d2h) {
      MAS_CUDA(copy_leading_rows(paths, d_paths, t_ys, b0, nb, T_y, T_x, cudaMemcpyDeviceToHost, s_out, es));
    } else {
      MAS_CUDA(cudaMemcpyAsync(h_index + static_cast<size_t>(b0) * T_y, d_index + static_cast<size_t>(b0) * T_y,
                               static_cast<size_t>(nb) * T_y * sizeof(int32_t), cudaMemcpyDeviceToHost, s_out));
      MAS_CUDA(cudaEventRecord(g_host.ev_idx[c], s_out));
    }
    nused = c + 1;
  }
  if (dense_d2h && zero_tail) {
    // the rows no copy writes (at or beyond the longest utterance of each group), zeroed on the host while the
    // copies are in flight: the caller may then hand in an uninitialised buffer (no np.zeros pass of its own)
    for (int c = 0, b0 = 0; c < ng; b0 += gsize[c], ++c) {
      int rows = 0;
      for (int b = b0; b < b0 + gsize[c]; ++b) rows = t_ys[b] > rows ? t_ys[b] : rows;
      rows = rows > T_y ? T_y : (rows < 0 ? 0 : rows);
      if (rows == T_y) continue;
      for (int b = b0; b < b0 + gsize[c]; ++b)
        memset(static_cast<unsigned char*>(paths) + (plane * b + static_cast<size_t>(rows) * T_x) * es, 0,
               static_cast<size_t>(T_y - rows) * T_x * es);
    }
  }
  // the status word of every group (first word of its scratch) in one strided copy
  MAS_CUDA(cudaMemcpy2DAsync(g_host.h_status, sizeof(int32_t),
                             static_cast<unsigned char*>(g_host.d_scratch) + mas_scratch_status_offset(), sc_one,
                             sizeof(int32_t), nused, cudaMemcpyDeviceToHost, s_out));
  if (pool) {
    // hand each group to the pool as soon as its index has landed, work on it too, and wait for the last row
    cudaError_t ee = cudaSuccess;
    for (int c = 0, b0 = 0; c < ng; b0 += gsize[c], ++c) {
      const cudaError_t e1 = cudaEventSynchronize(g_host.ev_idx[c]);
      if (e1 != cudaSuccess) ee = e1;
      pool->add(b0, gsize[c]);
    }
    pool->finish();
    if (ee != cudaSuccess) return static_cast<int>(ee);
  }
  for (auto& s : g_host.streams) MAS_CUDA(cudaStreamSynchronize(s));
  int status = 0;
  for (int c = 0; c < nused; ++c) status |= g_host.h_status[c];
  return status ? (status << 8) : MAS_OK;
}

void mas_host_release(void) { host_release(); }

size_t mas_neg_cent_scratch_bytes(int B, int C, int T_y, int T_x) { return mas::neg_cent_scratch_bytes(B, C, T_y, T_x); }

int mas_neg_cent(const float* z_p, const float* m_p, const float* logs_p, float* neg_cent, void* scratch,
                 size_t scratch_bytes, int B, int C, int T_y, int T_x, mas_stream_t stream) {
  return mas::neg_cent(z_p, m_p, logs_p, neg_cent, scratch, scratch_bytes, B, C, T_y, T_x,
                       static_cast<cudaStream_t>(stream));
}

size_t mas_stats_to_path_scratch_bytes(int B, int C, int T_y, int T_x) { return mas::stats_to_path_scratch_bytes(B, C, T_y, T_x); }

int mas_stats_to_path(const float* z_p, const float* m_p, const float* logs_p, const int32_t* t_ys, const int32_t* t_xs,
                      void* path_out, int path_dtype, int32_t* index_out, void* scratch, size_t scratch_bytes, int B, int C,
                      int T_y, int T_x, mas_stream_t stream) {
  return mas::stats_to_path(z_p, m_p, logs_p, t_ys, t_xs, path_out, path_dtype, index_out, scratch, scratch_bytes, B, C, T_y,
                            T_x, static_cast<cudaStream_t>(stream));
}

int mas_neg_cent_autocast(const float* z_p, const float* m_p, const float* logs_p, float* neg_cent, int gemm_dtype,
                          int stats_lowp, int B, int C, int T_y, int T_x, mas_stream_t stream) {
  return mas::neg_cent_autocast(z_p, m_p, logs_p, neg_cent, gemm_dtype, stats_lowp, B, C, T_y, T_x,
                                static_cast<cudaStream_t>(stream));
}

int mas_path_durations(const int32_t* index, float* w, int B, int T_y, int T_x, mas_stream_t stream) {
  return mas::path_durations(index, w, B, T_y, T_x, static_cast<cudaStream_t>(stream));
}

int mas_expand_prior(const int32_t* index, const float* m_p, const float* logs_p, float* m_out, float* logs_out, int B,
                     int C, int T_y, int T_x, mas_stream_t stream) {
  return mas::expand_prior(index, m_p, logs_p, m_out, logs_out, B, C, T_y, T_x, static_cast<cudaStream_t>(stream));
}

int mas_generate_path(const float* duration, const float* mask, int64_t mask_sb, int64_t mask_sy, int64_t mask_sx,
                      float* path, int B, int T_y, int T_x, mas_stream_t stream) {
  return mas::generate_path(duration, mask, mask_sb, mask_sy, mask_sx, path, B, T_y, T_x, static_cast<cudaStream_t>(stream));
}

int mas_kl_from_index(const int32_t* index, const float* z_p, const float* logs_q, const float* m_p, const float* logs_p,
                      const float* z_mask, double* out2, int B, int C, int T_y, int T_x, mas_stream_t stream) {
  return mas::kl_from_index(index, z_p, logs_q, m_p, logs_p, z_mask, out2, B, C, T_y, T_x, static_cast<cudaStream_t>(stream));
}

uint64_t mas_launch_count(void) { return mas::g_launches.load(std::memory_order_relaxed); }

/* Tuning hooks for benchmarks (not part of the reference-facing surface). 0 = automatic. */
void mas_set_tuning(int cols_per_lane, int rows_per_stage, int stages, int pdl) {
  mas::set_tuning(cols_per_lane, rows_per_stage, stages, pdl);
}
void mas_set_neg_cent_impl(int impl) { mas::set_neg_cent_impl(impl); }
void mas_set_debug_kernels(int mask) { mas::set_debug_kernels(mask); }
void mas_set_tuning2(int fused, int helpers) { mas::set_tuning2(fused, helpers); }
void mas_set_tuning3(int wavefront, int ring_mode, int ring_slots, int cols_per_lane) {
  mas::set_tuning3(wavefront, ring_mode, ring_slots, cols_per_lane);
}
void mas_set_timeline(void* dev_ptr) { mas::set_timeline(static_cast<unsigned long long*>(dev_ptr)); }
void mas_set_trace(void* dev_ptr) { mas::set_trace(static_cast<unsigned long long*>(dev_ptr)); }

}  // extern "C"
