// Instantiations of the forward kernel for K = 2 text columns per lane.
#include "mas_forward.cuh"

namespace mas {
cudaError_t launch_fwd_k2(bool vec, const FwdParams& p, int R, cudaStream_t st) {
  
  return vec ? launch_fwd<2, true>(p, R, st) : launch_fwd<2, false>(p, R, st);
}
}  // namespace mas
