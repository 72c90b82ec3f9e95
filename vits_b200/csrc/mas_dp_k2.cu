// One columns-per-lane instantiation of the forward DP kernel (mas_dp.cuh), its own translation
// unit so the instantiations compile in parallel.
#include "mas_dp.cuh"

namespace mas {
cudaError_t launch_dp_k2(const CUtensorMap& tmap, const DpParams& p, int skew, bool linear, cudaStream_t st) {
  return launch_dp<2>(tmap, p, skew, linear, st);
}
}  // namespace mas
