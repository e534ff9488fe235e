// Host-side launch interface between the C ABI (gpkl_api.cu) and the kernel tiers.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "gpkl_common.cuh"

namespace gpkl {

constexpr size_t kMaxDynSmem = 227 * 1024;  // opt-in dynamic shared memory per CTA on sm_100a
constexpr int kNumSMs = 148;

// measurement hooks (gpkl_api.cu): every kernel launch goes through note_launch(); the dominant kernel of
// forward/backward is bracketed by prof_begin/prof_end (no-ops unless gpkl_profile_enable(1)).
void note_launch(int n = 1);
void prof_begin(bool backward, cudaStream_t st);
void prof_end(bool backward, cudaStream_t st);
bool pdl_enabled();  // programmatic dependent launch of the per-pair kernels behind the shared-prior pre-pass (GPKL_PDL=0 disables)

// Shared-prior records (Params::prior): floats per sequence, an upper bound over the tiers' layouts
// (warp tier: packed L_p^-1 rows + diag + K_p^-1, gpkl_warp.cuh PriorRec; block tier: gpkl_block.cu).
// float64 forward records (gpkl_prior64.cu): column pitch in doubles.  Register tier (T_max <= 64): T_max rounded up to 8;
// beyond: T_max rounded up to the 64 x 64 tiles.
__host__ __device__ inline int prior64_pitch(int T_max) {
  const int t = T_max < 1 ? 1 : T_max;
  return t <= 64 ? (t + 7) / 8 * 8 : (t + 63) / 64 * 64;
}
inline size_t prior_record_floats(int T_max) {
  const size_t al = T_max > 208 ? 64 : 16;  // as Lay (gpkl_block.cu): 64 x 64 tiles beyond the resident sizes
  const size_t TP = ((size_t)(T_max < 1 ? 1 : T_max) + al - 1) / al * al;
  const size_t p64 = (size_t)prior64_pitch(T_max);
  const size_t f32 = 2 * (TP + 1) * (TP + 4), f64 = 2 * (p64 * p64 + 2);
  return f32 > f64 ? f32 : f64;
}

// generic tier (gpkl_generic.cu)
size_t generic_smem_bytes(int T, int S, bool mats_in_smem);
size_t generic_slot_floats(int T);
int generic_slots(const GpklDesc& d);  // 0 when the matrices fit shared memory
cudaError_t launch_generic(const Params& P, bool backward, cudaStream_t st);

// warp tier (gpkl_warp.cu): register-resident, T <= 64
bool warp_tier_supports(const GpklDesc& d, bool backward);
cudaError_t launch_warp(const Params& P, bool backward, cudaStream_t st);

// block tier (gpkl_block.cu): one CTA per pair, loop-based; matrices resident in shared memory for
// T <= ~144, otherwise in one workspace slot per CTA (kBlockSlots CTAs, 2 per SM)
constexpr int kBlockSlots = 2 * kNumSMs;
bool block_tier_supports(const GpklDesc& d, bool backward);
bool block_tier_resident(const GpklDesc& d);
size_t block_slot_floats(const GpklDesc& d);  // 0 when resident
cudaError_t launch_block(const Params& P, bool backward, cudaStream_t st);

// tile tier (gpkl_tile.cu): the shared-prior GP posterior for 208 < T <= 512 -- one tile-packed, L2-resident triangle per
// CTA, operands streamed by bulk asynchronous copies; launched by the block tier between its pre-pass and its per-pair
// kernel (it uses the block tier's workspace slots and prior records)
bool tile_tier_supports(const GpklDesc& d, bool backward);
size_t tile_slot_floats(const GpklDesc& d);
cudaError_t launch_tile(const Params& P, bool backward, cudaStream_t st);
// float64 per-sequence prior record of the tile tier's forward pass (gpkl_prior64.cu): K_p^-1 (lower triangle, column-major,
// pitch TP = T_max rounded up to 64) followed by log|K_p|; fits the records sized by prior_record_floats
size_t prior64_record_floats(int T_max);
cudaError_t launch_prior_inv64(const Params& P, cudaStream_t st);
cudaError_t launch_prior_inv64_small(const Params& P, cudaStream_t st, int f32_off, int f32_tm);

// V3 hot tier (gpkl_bidiag.cu): bidiagonal-precision posterior, T <= 64, shared prior; launches the register tier's float64
// pre-pass itself.  The caller launches the generic tier behind it with skip_if_shared (non-uniform ell_p).
bool bidiag_tier_supports(const GpklDesc& d);
cudaError_t launch_bidiag(const Params& P, bool backward, cudaStream_t st);

// reconstruction term (gpkl_recon.cu), SURVEY.md S8(f) row 1
int recon_grid(long long rows);
cudaError_t launch_recon_fwd(const float* x, const float* xd, const long long* off, int B, int F, int S, long long rows,
                             double* partials, double* out, cudaStream_t st);
cudaError_t launch_recon_bwd(const float* x, const float* xd, const long long* off, int B, int F, int S, long long rows,
                             const double* g_recon, float* g_xd, cudaStream_t st);


// rows either side of the path (gpkl_adjacent.cu), SURVEY.md S8(f) rows 3 and 4
cudaError_t launch_recog_fwd(const float* mean, const float* logvar, const float* eps, const long long* off, int B, int D,
                             int S, int T_max, long long rows, float* z, float* kl_rows, cudaStream_t st);
cudaError_t launch_recog_bwd(const float* mean, const float* logvar, const float* eps, const float* g_z,
                             const double* g_kl_sum, const float* g_kl_rows, const long long* off, int B, int D, int S,
                             int T_max, long long rows, float* g_mean, float* g_logvar, cudaStream_t st);
cudaError_t launch_collate_scan(const float* data, const float* grid, const int32_t* index, int B, int F, int T_full,
                                int max_time, int32_t* lengths, float* times, cudaStream_t st);
cudaError_t launch_collate_gather(const float* data, const int32_t* index, const long long* off, int B, int F, int T_full,
                                  int max_time, float* x, cudaStream_t st);

// GP posterior imputation (gpkl_impute.cu), SURVEY.md S8(f) row 2
size_t impute_smem_bytes(int nd_max, int ns);
cudaError_t launch_impute(int kernel, int B, int D, int nd_max, int ns, const float* z_obs, const float* t_obs,
                          const int32_t* n_obs, const float* t_full, const float* eps, const long long* off, float ell,
                          float noise, float* out, int32_t* status, cudaStream_t st);

}  // namespace gpkl
