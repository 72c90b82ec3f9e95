// Instantiations of the forward kernel for K = 8 text columns per lane.
#include "mas_forward.cuh"

namespace mas {
cudaError_t launch_fwd_k8(bool vec, const FwdParams& p, int R, cudaStream_t st) {
  
  return vec ? launch_fwd<8, true>(p, R, st) : launch_fwd<8, false>(p, R, st);
}
}  // namespace mas
