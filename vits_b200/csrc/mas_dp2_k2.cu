// The K = 2 instantiations of the second-generation wavefront DP kernel (mas_dp2.cuh), their own translation unit
// so that they compile in parallel with the rest.
#include "mas_dp2.cuh"

namespace mas {
cudaError_t launch_dp2_k2(const CUtensorMap& tmap, const DpParams& p, int skew, int cl, cudaStream_t st) {
  if (cl == 1) {
    if (skew == 1) return launch_dp2_t<2, 1, 1>(tmap, p, st);
    if (skew == 2) return launch_dp2_t<2, 2, 1>(tmap, p, st);
  } else if (cl == 2) {
    if (skew == 1) return launch_dp2_t<2, 1, 2>(tmap, p, st);
    if (skew == 2) return launch_dp2_t<2, 2, 2>(tmap, p, st);
  }
  return cudaErrorInvalidValue;
}
}  // namespace mas
