// Instantiations of the forward kernel for K = 4 text columns per lane.
#include "mas_forward.cuh"

namespace mas {
cudaError_t launch_fwd_k4(bool vec, const FwdParams& p, int R, cudaStream_t st) {
  
  return vec ? launch_fwd<4, true>(p, R, st) : launch_fwd<4, false>(p, R, st);
}
}  // namespace mas
