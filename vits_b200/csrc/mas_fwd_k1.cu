// Instantiations of the forward kernel for K = 1 text columns per lane.
#include "mas_forward.cuh"

namespace mas {
cudaError_t launch_fwd_k1(bool vec, const FwdParams& p, int R, cudaStream_t st) {
  (void)vec;
  return launch_fwd<1, true>(p, R, st);  // K == 1 loads are scalar either way
}
}  // namespace mas
