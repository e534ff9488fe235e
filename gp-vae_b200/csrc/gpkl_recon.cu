// Reconstruction term of the ELBO: the step right after the GP-prior path in the reference,
//   recon = sum_rows mean_s  -sum_f [ x log(1e-10 + xd) + (1 - x) log(1 - xd + 1e-10) ]
// (src/Models/Full_GP_VAE_dynamic_time.py:323-327 tiles x over the S samples, :349 the Bernoulli NLL per row in
// float32, :350-356 cast to float64, mean over samples, sum over time and batch; loss = recon + beta*KL at :360).
// One streaming pass: HBM-bound (4 B of xd per element, x re-read S times from L2), 128-bit loads, one warp per row
// chunk, float32 per-thread partials combined in float64, fixed-order final reduction (deterministic).
#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int RB_THREADS = 256;

// row of xd -> row of x:  sequence b owns xd rows [S*off[b], S*off[b+1]) as S blocks of T_b rows (gp_vae_sample layout)
__device__ __forceinline__ long long x_row_of(long long row, const long long* __restrict__ off, int B, int S) {
  int lo = 0, hi = B;  // find b with S*off[b] <= row < S*off[b+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((long long)S * off[mid] <= row) lo = mid;
    else hi = mid;
  }
  const long long Tb = off[lo + 1] - off[lo];
  const long long local = row - (long long)S * off[lo];
  return off[lo] + (Tb > 0 ? local % Tb : 0);
}

// __logf (MUFU.LG2 * ln 2, ~2 ulp) and __fdividef keep this stage memory-bound; the accurate logf / IEEE division
// made it instruction-bound (45 % / 30 % of HBM peak).  Both are far inside the 1e-5 tolerance of the summed loss.
__device__ __forceinline__ float nll_elem(float x, float xd) {
  return x * __logf(1e-10f + xd) + (1.0f - x) * __logf(1.0f - xd + 1e-10f);
}

constexpr int CHUNK = 1024;  // floats per (row, chunk) work item: 8 independent 128-bit loads per lane and operand

__global__ void __launch_bounds__(RB_THREADS) recon_fwd_kernel(const float* __restrict__ x, const float* __restrict__ xd,
                                                               const long long* __restrict__ off, int B, int F, int S,
                                                               long long rows, double* __restrict__ partials) {
  __shared__ double red[32];
  const int lane = threadIdx.x & 31, wpb = RB_THREADS / 32;
  const long long warp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * wpb;
  const int nchunk = (F + CHUNK - 1) / CHUNK;
  double total = 0.0;
  for (long long item = warp; item < rows * nchunk; item += nwarps) {
    const long long row = item / nchunk;
    const int f0 = (int)(item - row * nchunk) * CHUNK;
    const float* __restrict__ xr = x + x_row_of(row, off, B, S) * F;
    const float* __restrict__ dr = xd + row * F;
    float acc = 0.0f;
    if ((F & 3) == 0) {
      float4 a[8], d4[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = f0 + 4 * lane + 128 * e;
        const bool ok = f < F;
        a[e] = ok ? *reinterpret_cast<const float4*>(xr + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        d4[e] = ok ? *reinterpret_cast<const float4*>(dr + f) : make_float4(0.5f, 0.5f, 0.5f, 0.5f);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        if (f0 + 4 * lane + 128 * e < F)
          acc += nll_elem(a[e].x, d4[e].x) + nll_elem(a[e].y, d4[e].y) + nll_elem(a[e].z, d4[e].z) + nll_elem(a[e].w, d4[e].w);
      }
    } else {
      const int fe = (f0 + CHUNK < F) ? f0 + CHUNK : F;
      for (int f = f0 + lane; f < fe; f += 32) acc += nll_elem(xr[f], dr[f]);
    }
    total += (double)acc;
  }
  total = block_sum(total, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = total;
}

__global__ void recon_final_kernel(const double* __restrict__ partials, int n, int S, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *out = -acc / (double)S;
}

__device__ __forceinline__ float nll_grad(float x, float xd, float gs) {
  return gs * (__fdividef(x, 1e-10f + xd) - __fdividef(1.0f - x, 1.0f - xd + 1e-10f));
}

__global__ void __launch_bounds__(RB_THREADS) recon_bwd_kernel(const float* __restrict__ x, const float* __restrict__ xd,
                                                               const long long* __restrict__ off, int B, int F, int S,
                                                               long long rows, const double* __restrict__ g_recon,
                                                               float* __restrict__ g_xd) {
  const int lane = threadIdx.x & 31, wpb = RB_THREADS / 32;
  const long long warp = (long long)blockIdx.x * wpb + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * wpb;
  const int nchunk = (F + CHUNK - 1) / CHUNK;
  const float gs = -(float)((g_recon ? *g_recon : 1.0) / (double)S);
  for (long long item = warp; item < rows * nchunk; item += nwarps) {
    const long long row = item / nchunk;
    const int f0 = (int)(item - row * nchunk) * CHUNK;
    const float* __restrict__ xr = x + x_row_of(row, off, B, S) * F;
    const float* __restrict__ dr = xd + row * F;
    float* __restrict__ gr = g_xd + row * F;
    if ((F & 3) == 0) {
      float4 a[8], d4[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = f0 + 4 * lane + 128 * e;
        const bool ok = f < F;
        a[e] = ok ? *reinterpret_cast<const float4*>(xr + f) : make_float4(0.f, 0.f, 0.f, 0.f);
        d4[e] = ok ? *reinterpret_cast<const float4*>(dr + f) : make_float4(0.5f, 0.5f, 0.5f, 0.5f);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int f = f0 + 4 * lane + 128 * e;
        if (f < F)
          *reinterpret_cast<float4*>(gr + f) = make_float4(nll_grad(a[e].x, d4[e].x, gs), nll_grad(a[e].y, d4[e].y, gs),
                                                           nll_grad(a[e].z, d4[e].z, gs), nll_grad(a[e].w, d4[e].w, gs));
      }
    } else {
      const int fe = (f0 + CHUNK < F) ? f0 + CHUNK : F;
      for (int f = f0 + lane; f < fe; f += 32) gr[f] = nll_grad(xr[f], dr[f], gs);
    }
  }
}

}  // namespace

int recon_grid(long long items) {
  const long long want = (items + (RB_THREADS / 32) - 1) / (RB_THREADS / 32);
  const long long cap = (long long)kNumSMs * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

cudaError_t launch_recon_fwd(const float* x, const float* xd, const long long* off, int B, int F, int S, long long rows,
                             double* partials, double* out, cudaStream_t st) {
  const int grid = recon_grid(rows * ((F + 1023) / 1024));
  prof_begin(false, st);
  recon_fwd_kernel<<<grid, RB_THREADS, 0, st>>>(x, xd, off, B, F, S, rows, partials);
  prof_end(false, st);
  recon_final_kernel<<<1, 256, 0, st>>>(partials, grid, S, out);
  note_launch(2);
  return cudaGetLastError();
}

cudaError_t launch_recon_bwd(const float* x, const float* xd, const long long* off, int B, int F, int S, long long rows,
                             const double* g_recon, float* g_xd, cudaStream_t st) {
  prof_begin(true, st);
  recon_bwd_kernel<<<recon_grid(rows * ((F + 1023) / 1024)), RB_THREADS, 0, st>>>(x, xd, off, B, F, S, rows, g_recon, g_xd);
  prof_end(true, st);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gpkl
