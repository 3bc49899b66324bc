// placeholder until the specialised kernels land
#include "jb_inverse.cuh"
bool jb_inv_fast_eligible(const JbGeom&) { return false; }
cudaError_t jb_launch_inv_fast(const JbInvArgs&, int, cudaStream_t) { return cudaErrorNotSupported; }
