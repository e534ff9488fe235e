// One (LP, R, direction) instantiation of the warp tier per translation unit, so the fully unrolled
// kernels compile in parallel:  nvcc -DGPKL_LP=16 -DGPKL_R=3 -DGPKL_BWD=1 -c gpkl_warp_inst.cu
#include "gpkl_warp.cuh"

#if !defined(GPKL_LP) || !defined(GPKL_R) || !defined(GPKL_BWD)
#error "compile with -DGPKL_LP=.. -DGPKL_R=.. -DGPKL_BWD=0|1 (see gp-vae_b200/build.py)"
#endif

namespace gpkl {
template <int LP, int R, bool BWD>
cudaError_t launch_warp_inst(const Params& P, cudaStream_t st) {
  return launch_kp<LP, R, BWD>(P, st);
}
template cudaError_t launch_warp_inst<GPKL_LP, GPKL_R, (GPKL_BWD != 0)>(const Params&, cudaStream_t);
}  // namespace gpkl
