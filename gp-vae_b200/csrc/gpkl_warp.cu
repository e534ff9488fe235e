// Warp tier dispatch (kernels: gpkl_warp.cuh, instantiated per (LP, R, direction) in gpkl_warp_inst.cu).
#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {

template <int LP, int R, bool BWD>
cudaError_t launch_warp_inst(const Params& P, cudaStream_t st);

bool warp_tier_supports(const GpklDesc& d, bool backward) {
  if (d.T_max > 64 || d.T_max < 1) return false;
  if (d.posterior != GPKL_POST_GP && d.posterior != GPKL_POST_DIAG) return false;
  // d/d ell_p (trainable prior, Full_GP_VAE_fixed_for_MovMnist.py:96): GP posterior; its tables bound S by shared memory
  if (backward && (d.flags & GPKL_FLAG_GRAD_ELL_P) && (d.posterior != GPKL_POST_GP || d.S > 4)) return false;
  if (d.S > 8) return false;  // S samples per pair live in registers; more than 8 are served by the block / generic tiers
  return true;
}

template <int LP, int R>
static cudaError_t both(const Params& P, bool backward, cudaStream_t st) {
  return backward ? launch_warp_inst<LP, R, true>(P, st) : launch_warp_inst<LP, R, false>(P, st);
}

cudaError_t launch_warp(const Params& P, bool backward, cudaStream_t st) {
  const int T = P.d.T_max;
  if (T <= 8) return both<8, 1>(P, backward, st);
  if (T <= 16) return both<16, 1>(P, backward, st);
  if (T <= 32) return both<32, 1>(P, backward, st);
  if (T <= 48) return both<16, 3>(P, backward, st);
  return both<32, 2>(P, backward, st);
}

}  // namespace gpkl
