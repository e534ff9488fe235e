// GP posterior conditioning / imputation: predictive mean and sample of a latent sequence on a
// full time grid given its values at the observed time points.  Reference replaced:
//   sample_given_part_latent   src/Models/FullGP_and_GPdecoder_dynamic_time_analysis.py:40-56
//   post_gp_sample             src/Models/FullGP_and_GPdecoder_dynamic_time_analysis.py:96-111  (loop over sequences and latent rows)
// with kernel_function / kernel_matrix of the same file (:8-22): K(t1,t2) = (1-noise) k(t1-t2) + noise wherever t1 == t2
// EXACTLY (also between an observed and a full-grid point), entries rounded to float32 (:17).
//
//   L   = chol(K_dd)                      float32 (numpy factors the float32 matrix in float32, :43-44)
//   Lk  = L^-1 K_ds                       float32 (:46-47)
//   mu  = Lk^T L^-1 z_d                   float32 (:48)
//   C   = K_ss + 1e-15 I - Lk^T Lk        float64 (:49-50: the float64 identity promotes the sum)
//   f   = mu + chol(C) eps                float64 (:50-51), returned as float32
//
// The reference recomputes all of this for every latent row of a sequence although the matrices depend on the time stamps
// only (time_char = 1.0 for every row, :12); here one CTA factors them ONCE per sequence in shared memory and then serves
// the D rows.  As in the reference the second Cholesky runs even when only the mean is requested (:50 precedes the early
// return :53): a non-positive pivot -- what happens whenever an observed time coincides with a full-grid time, where C is
// exactly singular and the reference raises LinAlgError -- bumps *status; the mean is still written (it is well defined),
// the sample rows become NaN.
#include <stdint.h>

#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int NT = 256;

template <int KERNEL>
__device__ __forceinline__ float kval(float t1, float t2, double ell, double noise) {
  // kernel_function :8-14 evaluates in Python floats (float64) and kernel_matrix stores float32
  const double nz = (t1 == t2) ? noise : 0.0;
  const double d = (double)t1 - (double)t2;
  const double k = (KERNEL == GPKL_KERNEL_RBF) ? exp(-(d * d) / (2.0 * ell * ell)) : 1.0 / (1.0 + d * d / (ell * ell));
  return (float)((1.0 - nz) * k + nz);
}

template <int KERNEL>
__global__ void __launch_bounds__(NT) impute_kernel(int B, int D, int nd_max, int ns, const float* __restrict__ z_obs,
                                                    const float* __restrict__ t_obs, const int32_t* __restrict__ n_obs,
                                                    const float* __restrict__ t_full, const float* __restrict__ eps,
                                                    const long long* __restrict__ off, float ell, float noise,
                                                    float* __restrict__ out, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int bad;
  const int b = blockIdx.x, tid = threadIdx.x;
  const int nd = max(n_obs[b], 0);
  double* C = reinterpret_cast<double*>(smem_raw);           // [ns][ns] lower triangle used, row-major
  float* Kd = reinterpret_cast<float*>(C + (size_t)ns * ns);  // [nd][nd]  L(i,k) at Kd[i*nd + k]
  float* Lk = Kd + (size_t)nd_max * nd_max;                  // [nd][ns]
  float* td = Lk + (size_t)nd_max * ns;                      // [nd_max]
  float* tf = td + nd_max;                                   // [ns]
  float* v = tf + ns;                                        // [32][nd_max]: L^-1 z of 32 latent rows at a time
  if (tid == 0) bad = 0;
  for (int i = tid; i < nd; i += NT) td[i] = t_obs[(size_t)b * nd_max + i];
  for (int i = tid; i < ns; i += NT) tf[i] = t_full[(size_t)b * ns + i];
  __syncthreads();
  for (int e = tid; e < nd * nd; e += NT) { const int i = e / nd, k = e - i * nd; Kd[e] = kval<KERNEL>(td[i], td[k], ell, noise); }
  for (int e = tid; e < nd * ns; e += NT) { const int i = e / ns, j = e - i * ns; Lk[e] = kval<KERNEL>(td[i], tf[j], ell, noise); }
  for (int e = tid; e < ns * ns; e += NT) {
    const int i = e / ns, j = e - i * ns;
    C[e] = (double)kval<KERNEL>(tf[i], tf[j], ell, noise) + (i == j ? 1e-15 : 0.0);
  }
  __syncthreads();
  // L = chol(K_dd), float32, right-looking (one barrier pair per column; n <= ~128)
  for (int c = 0; c < nd; ++c) {
    const float d = Kd[c * nd + c];
    if (tid == 0 && !(d > 0.0f)) bad = 1;
    const float sd = sqrtf(d);
    __syncthreads();
    for (int i = c + tid; i < nd; i += NT) Kd[i * nd + c] = (i == c) ? sd : Kd[i * nd + c] / sd;
    __syncthreads();
    for (int e = tid; e < (nd - c - 1) * (nd - c - 1); e += NT) {
      const int i = c + 1 + e / (nd - c - 1), k = c + 1 + e % (nd - c - 1);
      if (k <= i) Kd[i * nd + k] = fmaf(-Kd[i * nd + c], Kd[k * nd + c], Kd[i * nd + k]);
    }
    __syncthreads();
  }
  // Lk = L^-1 K_ds: one full-grid column per thread, forward substitution down the observed rows
  for (int j = tid; j < ns; j += NT) {
    for (int i = 0; i < nd; ++i) {
      float acc = Lk[i * ns + j];
      for (int k = 0; k < i; ++k) acc = fmaf(-Kd[i * nd + k], Lk[k * ns + j], acc);
      Lk[i * ns + j] = acc / Kd[i * nd + i];
    }
  }
  __syncthreads();
  // C = K_ss + 1e-15 I - Lk^T Lk (the product in float32 like np.dot on float32 arrays, the sum in float64), lower triangle
  for (int e = tid; e < ns * ns; e += NT) {
    const int i = e / ns, j = e - i * ns;
    if (j > i) continue;
    float acc = 0.0f;
    for (int k = 0; k < nd; ++k) acc = fmaf(Lk[k * ns + i], Lk[k * ns + j], acc);
    C[e] -= (double)acc;
  }
  __syncthreads();
  // chol(C), float64, in place; a non-positive pivot poisons the sample (NaN) and is reported
  for (int c = 0; c < ns; ++c) {
    const double d = C[c * ns + c];
    if (tid == 0 && !(d > 0.0)) bad = 1;
    const double sd = sqrt(d);
    __syncthreads();
    for (int i = c + tid; i < ns; i += NT) C[i * ns + c] = (i == c) ? sd : C[i * ns + c] / sd;
    __syncthreads();
    for (int e = tid; e < (ns - c - 1) * (ns - c - 1); e += NT) {
      const int i = c + 1 + e / (ns - c - 1), k = c + 1 + e % (ns - c - 1);
      if (k <= i) C[i * ns + k] -= C[i * ns + c] * C[k * ns + c];
    }
    __syncthreads();
  }
  if (tid == 0 && bad && status) atomicAdd(status, 1);
  // the D latent rows of this sequence, 32 at a time
  const long long r0 = off[b];
  for (int d0 = 0; d0 < D; d0 += 32) {
    const int nrow = min(32, D - d0);
    for (int r = tid; r < nrow; r += NT) {  // v = L^-1 z_d (forward substitution), one latent row per thread
      float* vr = v + (size_t)r * nd_max;
      for (int i = 0; i < nd; ++i) {
        float acc = z_obs[(size_t)(r0 + i) * D + d0 + r];
        for (int k = 0; k < i; ++k) acc = fmaf(-Kd[i * nd + k], vr[k], acc);
        vr[i] = acc / Kd[i * nd + i];
      }
    }
    __syncthreads();
    for (int e = tid; e < nrow * ns; e += NT) {
      const int i = e / nrow, r = e - i * nrow;  // consecutive threads: consecutive latent dims (coalesced output)
      const float* vr = v + (size_t)r * nd_max;
      float mu = 0.0f;
      for (int k = 0; k < nd; ++k) mu = fmaf(Lk[k * ns + i], vr[k], mu);
      double f = (double)mu;
      if (eps) {
        const float* er = eps + ((size_t)b * D + d0 + r) * ns;
        for (int k = 0; k <= i; ++k) f += C[i * ns + k] * (double)er[k];
      }
      out[((size_t)b * ns + i) * D + d0 + r] = (float)f;
    }
    __syncthreads();
  }
}

}  // namespace

size_t impute_smem_bytes(int nd_max, int ns) {
  return (size_t)ns * ns * sizeof(double) +
         ((size_t)nd_max * nd_max + (size_t)nd_max * ns + (size_t)nd_max + ns + 32 * (size_t)nd_max) * sizeof(float);
}

cudaError_t launch_impute(int kernel, int B, int D, int nd_max, int ns, const float* z_obs, const float* t_obs,
                          const int32_t* n_obs, const float* t_full, const float* eps, const long long* off, float ell,
                          float noise, float* out, int32_t* status, cudaStream_t st) {
  const size_t smem = impute_smem_bytes(nd_max, ns);
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  auto kern = kernel == GPKL_KERNEL_RBF ? impute_kernel<GPKL_KERNEL_RBF> : impute_kernel<GPKL_KERNEL_CAUCHY>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<B, NT, smem, st>>>(B, D, nd_max, ns, z_obs, t_obs, n_obs, t_full, eps, off, ell, noise, out, status);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gpkl
