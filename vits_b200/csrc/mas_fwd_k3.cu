// Instantiations of the forward kernel for K = 3 text columns per lane.
#include "mas_forward.cuh"

namespace mas {
cudaError_t launch_fwd_k3(bool vec, const FwdParams& p, int R, cudaStream_t st) {
  (void)vec;
  return launch_fwd<3, true>(p, R, st);  // K == 3 loads are scalar either way
}
}  // namespace mas
