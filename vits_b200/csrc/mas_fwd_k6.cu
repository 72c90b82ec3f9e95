// Instantiations of the forward kernel for K = 6 text columns per lane.
#include "mas_forward.cuh"

namespace mas {
cudaError_t launch_fwd_k6(bool vec, const FwdParams& p, int R, cudaStream_t st) {
  
  return vec ? launch_fwd<6, true>(p, R, st) : launch_fwd<6, false>(p, R, st);
}
}  // namespace mas
