// placeholder until the specialised kernels land
#include "jb_forward.cuh"
bool jb_fwd_fast_eligible(const JbGeom&) { return false; }
cudaError_t jb_launch_fwd_fast(const JbFwdArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
