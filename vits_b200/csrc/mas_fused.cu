// mas_fused.cu -- z_p, m_p, logs_p -> path in one pass (reference SynthesizerTrn.py:223-235): the contraction is
// streamed into the alignment search instead of materialising neg_cent [B,T_y,T_x] in HBM between two calls.
//
//   prep kernel            text-side operands + bias (mas_neg_cent_tc.cu), also clears this call's flags
//   contraction kernel     the tcgen05 GEMM in STREAMED mode on (SMs - B) SMs: tiles in frame-major order, skipping
//                          frame blocks past an utterance's length, written into a per-utterance ring of RT tiles and
//                          announced through tile_flags[b][mt] (fence + count)
//   wavefront DP kernel    one CTA per utterance on the remaining B SMs, launched programmatically behind the
//                          contraction kernel: its producer warps request a chunk's TMA boxes from the ring as soon as
//                          the chunk's tile flag is up (acquire + fence.proxy.async), so the search of frames
//                          0..127 overlaps the contraction of frames 128.. and the tiles are read back out of L2; one
//                          paced warp per CTA zero-fills the utterance's dense path
//   streaming backtrack    unchanged (mas_path.cu), co-resident with the DP CTAs
//
// Every wait is on an EARLIER kernel of the stream and has a watchdog (MAS_STATUS_TIMEOUT, all-zero path).  Shapes the
// streamed form does not cover (B too large for the SM split, T_x > 512) return MAS_E_UNSUPPORTED and the caller runs
// the two kernels back to back (mas_neg_cent + mas_maximum_path) -- the same CUDA code, through HBM.
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../../include/vits_mas.h"
#include "mas_internal.h"

namespace mas {

namespace {
struct FusedLayout {
  int Nt, NTL, MT, RT, pitch, zstride;
  size_t off_tc, tc_bytes, off_flags, off_ring, off_mas, mas_bytes, total;
};

FusedLayout fused_layout(int B, int C, int T_y, int T_x) {
  FusedLayout L{};
  neg_cent_tc_dims(C, T_y, T_x, &L.Nt, &L.NTL, &L.MT);
  L.RT = L.MT;  // the ring holds every frame block: the producer never has to wait for the consumer
  L.pitch = L.NTL * L.Nt;
  L.zstride = L.MT + 1;  // words per utterance: MT tile counters + one fill flag (laid out as [B][MT] then [B])
  auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
  L.off_tc = 0;
  L.tc_bytes = up(neg_cent_tc_scratch_bytes(B, C, T_y, T_x));
  L.off_flags = L.off_tc + L.tc_bytes;
  L.off_ring = up(L.off_flags + static_cast<size_t>(B) * L.zstride * 4);
  L.off_mas = up(L.off_ring + static_cast<size_t>(B) * L.RT * 128 * L.pitch * 4);
  L.mas_bytes = up(maximum_path_scratch_bytes(B, T_y, T_x));
  L.total = L.off_mas + L.mas_bytes;
  return L;
}
}  // namespace

size_t stats_to_path_scratch_bytes(int B, int C, int T_y, int T_x) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0) return 0;
  return fused_layout(B, C, T_y, T_x).total;
}

int stats_to_path(const float* z_p, const float* m_p, const float* logs_p, const int32_t* t_ys, const int32_t* t_xs,
                  void* path_out, int path_dtype, int32_t* index_out, void* scratch, size_t scratch_bytes, int B, int C,
                  int T_y, int T_x, cudaStream_t st) {
  if (B <= 0 || C <= 0 || T_y <= 0 || T_x <= 0 || B > 65535) return MAS_E_BAD_SHAPE;
  if (!z_p || !m_p || !logs_p || !t_ys || !t_xs || !path_out || !scratch) return MAS_E_NULL;
  if ((reinterpret_cast<uintptr_t>(z_p) | reinterpret_cast<uintptr_t>(m_p) | reinterpret_cast<uintptr_t>(logs_p)) & 3u)
    return MAS_E_ALIGN;
  if (reinterpret_cast<uintptr_t>(scratch) & 255u) return MAS_E_ALIGN;
  const FusedLayout L = fused_layout(B, C, T_y, T_x);
  if (scratch_bytes < L.total) return MAS_E_SCRATCH;
  const int sms = num_sms();
  static const int n1_env = [] {
    const char* e = getenv("MAS_FUSED_N1");  // benchmark hook (tools/timeline_fused.py): tiles per CTA in phase 1
    return e ? atoi(e) : -1;
  }();
  // SM budget: one per utterance for the DP, a few for the backtrack CTAs, the rest contract until the end
  const int gemm_ctas = sms - B - fused_backtrack_sms(B, T_y, T_x);
  if (gemm_ctas < sms / 3 || T_x > 512) return MAS_E_UNSUPPORTED;
  // Two-phase schedule (see TcParams): every SM contracts the first n1 tiles of its share, then B of them hand their SM
  // to the search.  The search needs ~T_y x 30 ns from its first tile to its last; giving the contraction the whole
  // machine for that much less than its total keeps both ends busy.  (c2 full-length: 512 tiles, 2 per CTA first.)
  int n1 = n1_env;
  if (n1 < 0) {
    const double tile_us = 8.75, dp_us = 0.030 * T_y;
    const double tiles = static_cast<double>(B) * L.MT * L.NTL;
    const double t1 = (tiles - dp_us * gemm_ctas / tile_us) / (sms / tile_us);   // time with all SMs so that both finish together
    n1 = t1 > 0 ? static_cast<int>(t1 / tile_us + 0.7) : 0;
  }

  unsigned char* sc = static_cast<unsigned char*>(scratch);
  uint32_t* flags = reinterpret_cast<uint32_t*>(sc + L.off_flags);
  float* ring = reinterpret_cast<float*>(sc + L.off_ring);

  // Ask the search first whether it can take a streamed source at this shape (nothing is launched on refusal).
  // flags: [B][MT] tile counters (indexed b * MT + mt by both kernels), then [B] fill flags -- B * zstride words in all,
  // which is what the prep kernel clears.  (The fill flags used to sit at flags + MT with stride MT + 1, i.e. INSIDE the
  // tile counters for B > 1: utterance 0's fill flag was utterance 1's first tile counter, the backtrack kernel then did
  // not wait for the zero-fill, and about one call in a hundred at B = 2, T_y = 1100 lost the ones of its last frames
  // to a late zero -- tools/stress_streamed_small.py.)
  FusedSrc fs{ring, L.pitch, L.RT, L.MT, flags, L.NTL, flags + static_cast<size_t>(B) * L.MT, 1, gemm_ctas};
  TcStream so{ring, flags, t_ys, t_xs, L.RT, L.pitch, n1 > 0 ? sms : gemm_ctas, gemm_ctas, n1, flags, L.zstride};
  int rc = maximum_path(ring, t_ys, t_xs, nullptr, 0, 0, 0, 0, path_out, path_dtype, index_out, sc + L.off_mas, L.mas_bytes,
                        B, T_y, T_x, st, &fs, /*probe_only=*/true);
  if (rc != MAS_OK) return rc;
  rc = neg_cent_tc_impl(z_p, m_p, logs_p, nullptr, sc + L.off_tc, L.tc_bytes, B, C, T_y, T_x, st, &so);
  if (rc != MAS_OK) return rc;
  return maximum_path(ring, t_ys, t_xs, nullptr, 0, 0, 0, 0, path_out, path_dtype, index_out, sc + L.off_mas, L.mas_bytes, B,
                      T_y, T_x, st, &fs, false);
}

}  // namespace mas
