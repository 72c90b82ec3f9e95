/*
 * vits_mas.h -- C ABI of libvits_mas.so: the B200 (sm_100a) replacement for the training-time
 * alignment hot path of Aloento/VITS.
 *
 * Every entry point names the reference interface it replaces (file:line are into the
 * reference tree).  Plain pointers and sizes only; no torch types.  All device entry points are
 * asynchronous on the given stream and never synchronise; they return 0 (MAS_OK) or a negative
 * MAS_E_* code for argument errors, or a positive cudaError_t when a launch fails.  There is no
 * CPU fallback: without a CUDA device every compute entry fails with a cudaError.
 *
 * Per-utterance length errors (t_x > t_y, zero lengths -- undefined behaviour in the reference,
 * core.pyx:13-33) cannot be reported synchronously because lengths live on the device: such an
 * utterance gets an all-zero path and a sticky bit in the status word at the start of the
 * scratch buffer (see mas_scratch_status_offset).
 */
#ifndef VITS_MAS_H
#define VITS_MAS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

typedef void* mas_stream_t; /* cudaStream_t */

enum {
  MAS_OK = 0,
  MAS_E_BAD_SHAPE = -1,   /* B, T_y, T_x <= 0 or above the supported maximum */
  MAS_E_BAD_DTYPE = -2,   /* unknown element-type code */
  MAS_E_NULL = -3,        /* a required pointer is NULL */
  MAS_E_SCRATCH = -4,     /* scratch buffer smaller than mas_maximum_path_scratch_bytes() */
  MAS_E_ALIGN = -5,       /* pointer not aligned to its element size */
  MAS_E_UNSUPPORTED = -6  /* configuration this build has no kernel for */
};

/* Element-type codes for the mask (input) and path (output) tensors. */
enum {
  MAS_F32 = 0, MAS_F16 = 1, MAS_BF16 = 2, MAS_F64 = 3,
  MAS_U8 = 4 /* also bool */, MAS_I8 = 5, MAS_I16 = 6, MAS_I32 = 7, MAS_I64 = 8
};

/* Bits of the device status word. */
enum {
  MAS_STATUS_TX_GT_TY = 1,   /* some utterance had t_x > t_y                     */
  MAS_STATUS_EMPTY = 2,      /* some utterance had t_x < 1 or t_y < 1            */
  MAS_STATUS_TOO_LONG = 4,   /* some utterance had t_y > T_y or t_x > T_x        */
  MAS_STATUS_TIMEOUT = 8     /* internal: a kernel gave up waiting (2 s) for an earlier kernel of the stream; the
                              * utterance's path is all-zero / index -1, never built from incomplete data */
};

int mas_abi_version(void);
const char* mas_error_string(int code);
/* Diagnostics: source line (vits_b200/csrc/mas_path.cu) of the CUDA call that made the calling thread's last
 * mas_maximum_path fail with a positive cudaError_t; 0 if none. */
int mas_last_error_site(void);

/* Bytes of device scratch one maximum_path call needs (direction bits, per-frame index,
 * lengths, status word).  The caller allocates it (e.g. with torch) and may reuse it across
 * calls on the same stream. */
size_t mas_maximum_path_scratch_bytes(int B, int T_y, int T_x);
/* Byte offset of the int32 status word inside the scratch buffer. */
size_t mas_scratch_status_offset(void);
/* Host pointer to four int32 words (pinned, device-mapped, owned by the library, one set per device; the call
 * returns the current device's): word k becomes 1 as soon as a kernel raises status bit k (1 << k).  Reading them
 * costs nothing and needs no synchronisation -- the device word in the scratch needs a copy -- so a caller can
 * poll for MAS_STATUS_TIMEOUT (results invalid: the affected utterances get an all-zero path and index -1) or the
 * length errors between calls; it clears the words itself.  NULL when pinned memory is unavailable. */
int32_t* mas_status_mirror(void);

/*
 * mas_maximum_path -- replaces maximum_path_c (monotonic_align/core.pyx:36-42) together with the
 * host marshalling around it (monotonic_align/__init__.py:14-20): the forward dynamic program over
 * the band, the backtrack, and the materialisation of the dense 0/1 path.
 *
 *   neg_cent   device, float32 [B, T_y, T_x] contiguous; NOT modified (the reference works on a
 *              host copy, __init__.py:14)
 *   Lengths, one of:
 *     t_ys,t_xs  device int32 [B]  (core.pyx:38 `t_ys`, `t_xs`), or
 *     mask       device [B, T_y, T_x] view with element strides (mask_sb, mask_sy, mask_sx) of type
 *                mask_dtype; lengths are the sums of column 0 and of row 0 exactly as
 *                __init__.py:17-18 computes them.  Pass t_ys = t_xs = NULL to use the mask.
 *   path_out   device [B, T_y, T_x] contiguous of type path_dtype; fully overwritten with 0/1
 *              (replaces np.zeros + the final cast, __init__.py:15,20).  May be NULL when only
 *              the index is wanted.
 *   index_out  optional device int32 [B, T_y]: text position of each frame, -1 on padded frames.
 *   scratch    device, >= mas_maximum_path_scratch_bytes(B,T_y,T_x) bytes, 16-byte aligned.
 */
int mas_maximum_path(const float* neg_cent,
                     const int32_t* t_ys, const int32_t* t_xs,
                     const void* mask, int mask_dtype, int64_t mask_sb, int64_t mask_sy, int64_t mask_sx,
                     void* path_out, int path_dtype,
                     int32_t* index_out,
                     void* scratch, size_t scratch_bytes,
                     int B, int T_y, int T_x, mas_stream_t stream);

/*
 * mas_maximum_path_c_host -- same argument meaning as the reference's native entry
 * `maximum_path_c(int[:,:,::1] paths, float[:,:,::1] values, int[::1] t_ys, int[::1] t_xs)`
 * (core.pyx:38) with HOST pointers: copies values to the device, runs mas_maximum_path, copies
 * the int32 paths back, and synchronises.  `values` is left untouched (the reference clobbers
 * its host copy; nothing reads it afterwards, __init__.py:19-20).  Like the reference, which loops
 * over the first t_ys[b] rows of a caller-zeroed `paths` (core.pyx:13-33, __init__.py:15), rows
 * y >= t_ys[b] need not be read from `values` nor written to `paths` (the padded tail is skipped up to
 * the longest utterance of each internal group; what is written there is zeros): pass `paths`
 * zero-filled, as np.zeros does.  Rows below t_ys[b] are written in full (zeros and ones).  Only `values` crosses the
 * link in full: the path comes back as the 4-byte-per-frame index and is written into `paths` by a small pool of host
 * threads of the library's own (MAS_HOST_THREADS, default min(8, cores/2)) while later groups are still in flight.
 * `values` may be pageable (a plain numpy array): it is then staged into a pinned mirror by the same threads, ahead of
 * the DMA (MAS_HOST_NOSTAGE=1 leaves the staging to the driver).
 * Uses internal streams and cached device/pinned buffers; not re-entrant (the reference has a single caller thread).
 * Returns 0, a MAS_E_* code, or MAS_STATUS_* bits << 8 when an utterance had invalid lengths.
 */
int mas_maximum_path_c_host(int32_t* paths, const float* values, const int32_t* t_ys, const int32_t* t_xs,
                            int B, int T_y, int T_x);
/*
 * mas_maximum_path_host -- the same host-buffer entry for callers that want the path in their own element type
 * (what the wrapper's final `.to(device=device, dtype=dtype)` produces, __init__.py:20) without a host-side cast
 * pass: `paths` is a HOST buffer [B, T_y, T_x] of `path_dtype` (MAS_* code).  zero_tail != 0: the rows no copy
 * writes are zeroed by the entry itself (memset overlapped with the copies), so `paths` may arrive uninitialised;
 * zero_tail == 0: `paths` must arrive zero-filled, as for mas_maximum_path_c_host.
 */
int mas_maximum_path_host(void* paths, int path_dtype, int zero_tail, const float* values, const int32_t* t_ys,
                          const int32_t* t_xs, int B, int T_y, int T_x);
/* Release the buffers cached by the host-buffer entries (current device). */
void mas_host_release(void);

/*
 * mas_neg_cent -- replaces the inline contraction of SynthesizerTrn.forward
 * (SynthesizerTrn.py:223-232): neg_cent[b,t,s] = sum_d( -0.5 log 2pi - logs_p - 0.5 z_p^2 e^{-2 logs_p}
 * + z_p m_p e^{-2 logs_p} - 0.5 m_p^2 e^{-2 logs_p} ).
 *   z_p     device float32 [B, C, T_y]   (T_y contiguous)
 *   m_p     device float32 [B, C, T_x]
 *   logs_p  device float32 [B, C, T_x]
 *   neg_cent device float32 [B, T_y, T_x], fully overwritten
 *   scratch  device, >= mas_neg_cent_scratch_bytes(B,C,T_y,T_x) bytes, 16-byte aligned
 */
size_t mas_neg_cent_scratch_bytes(int B, int C, int T_y, int T_x);
int mas_neg_cent(const float* z_p, const float* m_p, const float* logs_p, float* neg_cent,
                 void* scratch, size_t scratch_bytes,
                 int B, int C, int T_y, int T_x, mas_stream_t stream);

/*
 * mas_stats_to_path -- SynthesizerTrn.py:223-235 in one pass: the contraction above streamed into the alignment search
 * (mas_maximum_path) tile by tile, so neg_cent [B,T_y,T_x] is never written to and read back from HBM between two
 * calls -- it lives in an L2-resident ring inside `scratch` -- and the two stages overlap on disjoint SMs.  Same
 * numerics as mas_neg_cent followed by mas_maximum_path (bit-identical paths).  z_p/m_p/logs_p as for mas_neg_cent;
 * t_ys/t_xs device int32 [B] (x_lengths / y_lengths of SynthesizerTrn.forward -- no [B,T_y,T_x] mask is needed, :234);
 * path_out/index_out as for mas_maximum_path (path_out required).  scratch: 256-byte aligned,
 * >= mas_stats_to_path_scratch_bytes.  Returns MAS_E_UNSUPPORTED (nothing launched) for shapes the streamed form does
 * not cover (T_x > 512, or B too close to the SM count): call mas_neg_cent + mas_maximum_path instead.
 */
size_t mas_stats_to_path_scratch_bytes(int B, int C, int T_y, int T_x);
int mas_stats_to_path(const float* z_p, const float* m_p, const float* logs_p, const int32_t* t_ys, const int32_t* t_xs,
                      void* path_out, int path_dtype, int32_t* index_out, void* scratch, size_t scratch_bytes,
                      int B, int C, int T_y, int T_x, mas_stream_t stream);

/*
 * mas_neg_cent_autocast -- the same contraction with the numerics the reference has AS TRAINED, inside
 * torch.autocast (train_and_evaluate.py:55, config_cje.yaml:11 fp16_run): the two einsums
 * (SynthesizerTrn.py:227, :229) take their operands rounded to `gemm_dtype` (MAS_F16 or MAS_BF16),
 * accumulate in fp32 and round each einsum's output to `gemm_dtype`; exp, pow and the channel sums
 * (:223, :225, :231) stay fp32; the four terms are added in fp32 left to right (:232).  Inputs and the
 * result are float32 as in mas_neg_cent.  `stats_lowp` != 0 says that logs_p was a `gemm_dtype` tensor
 * before the caller widened it (TextEncoder.proj returns that type under autocast, TextEncoder.py:101-104):
 * the elementwise part of :225 then ran in that type, so every (-0.5 log 2pi - logs_p) is rounded to it
 * before the fp32 sum.  A parity mode on CUDA cores, not a fast path; mas_neg_cent is the fp32 formulation
 * and the default.
 */
int mas_neg_cent_autocast(const float* z_p, const float* m_p, const float* logs_p, float* neg_cent,
                          int gemm_dtype, int stats_lowp, int B, int C, int T_y, int T_x, mas_stream_t stream);

/*
 * Callers either side of the path.  All pointers are device pointers, fp32 contiguous
 * unless strides are given; asynchronous on `stream`.
 *
 * mas_path_durations -- replaces `w = attn.sum(2)` (SynthesizerTrn.py:237): w[b,x] = number of frames the
 *   alignment gives to text position x.  index = the int32 [B,T_y] output of mas_maximum_path (-1 = padded).
 * mas_expand_prior   -- replaces `einsum('bctn,bdn->bdt', attn, m_p)` and the same for logs_p
 *   (SynthesizerTrn.py:247-248, 308-310, 359-360, 410-411): out[b,c,y] = src[b,c,index[b,y]], 0 on padded frames.
 *   logs_p / logs_out may both be NULL to expand a single tensor.
 * mas_generate_path  -- replaces commons.generate_path (commons.py:101-117): path[b,y,x] =
 *   ((y < cum[x]) - (y < cum[x-1])) * mask[b,y,x] with cum = cumsum(duration[b,:]); mask is a [B,T_y,T_x] fp32
 *   view with element strides (the reference's [b,1,t_y,t_x] attn_mask, squeezed).
 */
int mas_path_durations(const int32_t* index, float* w, int B, int T_y, int T_x, mas_stream_t stream);
int mas_expand_prior(const int32_t* index, const float* m_p, const float* logs_p, float* m_out, float* logs_out,
                     int B, int C, int T_y, int T_x, mas_stream_t stream);
int mas_generate_path(const float* duration, const float* mask, int64_t mask_sb, int64_t mask_sy, int64_t mask_sx,
                      float* path, int B, int T_y, int T_x, mas_stream_t stream);
/*
 * mas_kl_from_index -- kl_loss (losses.py:43-60) fused with the prior expansion that feeds it
 *   (SynthesizerTrn.py:247-248): out2[0] = sum over (b,c,y) of (logs_p - logs_q - 0.5 + 0.5 (z_p - m_p)^2
 *   exp(-2 logs_p)) * z_mask[b,y] with m_p, logs_p [B,C,T_x] gathered along `index` on the fly, out2[1] =
 *   sum of z_mask; the loss is out2[0] / out2[1].  z_p, logs_q fp32 [B,C,T_y]; z_mask fp32 [B,T_y]; out2 two
 *   doubles on the device (overwritten).
 */
int mas_kl_from_index(const int32_t* index, const float* z_p, const float* logs_q, const float* m_p, const float* logs_p,
                      const float* z_mask, double* out2, int B, int C, int T_y, int T_x, mas_stream_t stream);

/* Number of kernel launches the library has issued since load (bench.py's gpu_launches). */
uint64_t mas_launch_count(void);

/* Benchmark/tuning hooks (not part of the reference-facing surface); 0 = automatic choice.
 * cols_per_lane in {1,2,4,8}; rows_per_stage in {8,16,32}; stages >= 2; pdl: 0 = ordinary launches,
 * 1 = programmatic dependent launch between the kernels of one call, 2 = the forward kernel is launched
 * programmatically too, behind the previous kernel of the stream; negative (default) = 2 where that was measured
 * to pay (wavefront forward kernel + streaming backtrack kernel), 1 elsewhere. */
void mas_set_tuning(int cols_per_lane, int rows_per_stage, int stages, int pdl);
/* neg_cent implementation: -1 automatic, 0 fp32 CUDA cores, 1 tcgen05 (split-bf16). */
void mas_set_neg_cent_impl(int impl);
/* Benchmark isolation: bit0 forward DP, bit1 backtrack, bit2 write-out; default 7 (all).  Test hook: 7|8 launches
 * everything but the wavefront forward kernel, so the streaming backtrack runs into its watchdog (MAS_STATUS_TIMEOUT). */
void mas_set_debug_kernels(int mask);
/* fused: -1 automatic, 0 separate backtrack kernel after the forward kernel, 1 backtrack fused into the
 * forward kernel, 2 streaming backtrack kernel on the idle SMs while the forward kernel runs, 3 the same
 * behind the wavefront forward kernel; helpers: helper warps of the fused kernel (0 = automatic). */
void mas_set_tuning2(int fused, int helpers);
/* Wavefront forward kernel: wavefront 0 = never choose it automatically, -1 = automatic (the second-generation
 * kernel mas_dp2_kernel where it applies: linear ring, two columns per lane, skew <= 2), 1 = the first-generation
 * kernel mas_dp_kernel, 32 = mas_dp2_kernel, 33 = mas_dp2_kernel with its instruction-cache warmer (= automatic),
 * 37 = 33 plus a dummy walk of the mask when the lengths are given (tools/mask_traffic.py); ring_mode 1..3 =
 * linear ring (mirror slot) with that many frames of skew between lanes, 4 = select ring (skew 1);
 * ring_slots = chunks of 32 frames per warp (>= skew + 2, or 3); cols_per_lane in {1,2,4}; 0 = automatic. */
void mas_set_tuning3(int wavefront, int ring_mode, int ring_slots, int cols_per_lane);
/* Debug timeline: device pointer to 16 uint64 (slots 0,3,5 preset to ~0, the others to 0) that the
 * kernels update with min start / max end %globaltimer stamps (slots 8-15: phases of the backtrack kernel's
 * tail, tools/timeline_gap.py); NULL disables. */
void mas_set_timeline(void* dev_ptr);
/* Debug event trace of the forward kernel's CTA 0 (-DMAS_TRACE builds, tools/trace_dp.py): device pointer to
 * 8*512*2 uint64 (zeroed; the wavefront kernels use 8*256*8 + 2*B of them), or NULL. */
void mas_set_trace(void* dev_ptr);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VITS_MAS_H */
