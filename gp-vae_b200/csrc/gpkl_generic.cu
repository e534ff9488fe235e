// Generic tier: one CTA per (sequence, latent-dim) pair, any T.
//
// Correctness-first implementation of the whole path (V1 full-GP posterior, V2 diagonal posterior;
// RBF / Cauchy; S samples; ragged lengths; d/d ell_p) that every faster tier is tested against and
// that serves sizes the specialised tiers do not cover.  The two T x T work matrices live in shared
// memory when they fit (T <= ~160) and otherwise in a per-CTA slot of the caller's workspace that
// stays L2 resident; either way no T x T matrix is an input or output of the op.
//
// Buffer layout ("two triangles in one rectangle"): a buffer is (T+1) x ld floats, ld odd >= T.
//   LC(B,i,k) = B[k*ld + i]      lower-triangular factor L, column-major (i >= k)
//   XR(B,i,k) = B[(i+1)*ld + k]  lower-triangular inverse X = L^-1, row-major  (i >= k)
// so L and L^-1 of the same matrix coexist, column sweeps of L and row sweeps of X are both
// unit-stride across threads, and an odd ld keeps strided accesses bank-conflict free.
//
// Reference being replaced: tf_kernel / gp_vae_sample / gp_kl_div,
// src/Models/Full_GP_VAE_dynamic_time.py:149-172, :174-195, :242-260; V2
// src/Models/VAE_GPprior_diag_cov.py:64-71, :100-119; backward = TF autodiff (:361) restated with
// closed forms.
#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int NT = 256;
constexpr int WCH = 8;  // columns of W = X_p^T [A|a] processed per chunk in the d/d ell_p path

#define LC(Bm, i, k) (Bm)[(size_t)(k) * ld + (i)]
#define XR(Bm, i, k) (Bm)[(size_t)((i) + 1) * ld + (k)]

template <int KERNEL>
__device__ void build_K(float* __restrict__ Bm, int ld, int T, const float* __restrict__ t, float ell, float sig,
                        float noise) {
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int k = e / T, i = e - k * T;
    if (i >= k) {
      float v = kern_val<KERNEL>(t[i] - t[k], ell, sig);
      if (i == k) v += noise;
      LC(Bm, i, k) = v;
    }
  }
}

// Right-looking Cholesky, in place in the LC triangle; dg receives diag(L).
__device__ void chol_inplace(float* __restrict__ Bm, int ld, int T, float* __restrict__ dg, int* bad) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = 0; j < T; ++j) {
    __syncthreads();
    const float d = LC(Bm, j, j);
    const float s = sqrtf(d);
    const float rinv = 1.0f / s;
    if (threadIdx.x == 0) {
      dg[j] = s;
      if (!(d > 0.0f)) *bad = 1;
    }
    for (int i = j + 1 + threadIdx.x; i < T; i += blockDim.x) LC(Bm, i, j) *= rinv;
    __syncthreads();
    for (int k = j + 1 + warp; k < T; k += nw) {
      const float lkj = LC(Bm, k, j);
      for (int i = k + lane; i < T; i += 32) LC(Bm, i, k) -= LC(Bm, i, j) * lkj;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < T; j += blockDim.x) LC(Bm, j, j) = dg[j];
  __syncthreads();
}

__device__ __forceinline__ int warp_min(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// In place:  LC(Bq) <- L_p^-1 LC(Bq)  (A = L_p^-1 L_q, lower triangular) and a <- L_p^-1 a.
// One column per thread (the vector is column index T); rows advance in lock step so that the
// L_p(i,k) operand is a broadcast.
__device__ void trsm_cols(const float* __restrict__ Bp, float* __restrict__ Bq, int ld, int T,
                          const float* __restrict__ dgp, float* __restrict__ a, bool with_matrix) {
  const int ncol = T + 1;
  for (int c0 = 0; c0 < ncol; c0 += blockDim.x) {
    const int cc = c0 + threadIdx.x;
    const bool isvec = (cc == T);
    const bool active = (cc < ncol) && (isvec || with_matrix);
    const int cstart = isvec ? 0 : cc;
    float* col = isvec ? a : (Bq + (size_t)cc * ld);
    const int kmin = warp_min(active ? cstart : T);
    for (int i = kmin; i < T; ++i) {
      const bool on = active && i >= cstart;
      float s = on ? col[i] : 0.0f;
      for (int k = kmin; k < i; ++k) {
        const float l = LC(Bp, i, k);
        const float x = (active && k >= cstart) ? col[k] : 0.0f;
        s = fmaf(-l, x, s);
      }
      if (on) col[i] = s / dgp[i];
    }
  }
  __syncthreads();
}

// XR(B) <- L^-1 where L = LC(B) (same buffer, disjoint triangles).
__device__ void trinv(float* __restrict__ Bm, int ld, int T, const float* __restrict__ dg) {
  for (int c0 = 0; c0 < T; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    const bool active = c < T;
    const int kmin = warp_min(active ? c : T);
    for (int i = kmin; i < T; ++i) {
      float s = (i == c) ? 1.0f : 0.0f;
      for (int k = kmin; k < i; ++k) {
        const float l = LC(Bm, i, k);
        const float x = (active && k >= c) ? XR(Bm, k, c) : 0.0f;
        s = fmaf(-l, x, s);
      }
      if (active && i >= c) XR(Bm, i, c) = s / dg[i];
    }
  }
  __syncthreads();
}

// sum_{k != l} dK(k,l)/d ell * sum_{i >= max(k,l)} U(i,k) V(i,l), U/V in XR layout.
// Returns this thread's partial sums for up to two lengthscales (second skipped when ell2 <= 0).
template <int KERNEL>
__device__ void tri_contract(const float* __restrict__ U, const float* __restrict__ V, int ld, int T,
                             const float* __restrict__ t, float sig, float ell1, float ell2, double& o1,
                             double& o2) {
  const float inv_sig = 1.0f / sig;
  const float il31 = 1.0f / (ell1 * ell1 * ell1);
  const float il32 = ell2 > 0.0f ? 1.0f / (ell2 * ell2 * ell2) : 0.0f;
  double a1 = 0.0, a2 = 0.0;
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int l = e / T, k = e - l * T;
    if (k == l) continue;
    float dot = 0.0f;
    for (int i = max(k, l); i < T; ++i) dot = fmaf(XR(U, i, k), XR(V, i, l), dot);
    const float dt = t[k] - t[l];
    const float k1 = kern_val<KERNEL>(dt, ell1, sig);
    a1 += (double)dot * (double)kern_dell<KERNEL>(dt, k1, il31, inv_sig);
    if (ell2 > 0.0f) {
      const float k2 = kern_val<KERNEL>(dt, ell2, sig);
      a2 += (double)dot * (double)kern_dell<KERNEL>(dt, k2, il32, inv_sig);
    }
  }
  o1 = a1;
  o2 = a2;
}

// W = B^-1 for upper-bidiagonal B (diag b, super-diagonal c): upper triangular, W(i,j) = Wm[i*ld + j], j >= i.
// One column per thread by back substitution: W(j,j) = 1/b_j, W(i,j) = -c_i W(i+1,j) / b_i.
__device__ void bidiag_inverse(float* __restrict__ Wm, int ld, int T, const float* __restrict__ b,
                               const float* __restrict__ c) {
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    float w = 1.0f / b[j];
    Wm[(size_t)j * ld + j] = w;
    for (int i = j - 1; i >= 0; --i) {
      w = -c[i] * w / b[i];
      Wm[(size_t)i * ld + j] = w;
    }
  }
  __syncthreads();
}

struct SmemPlan {
  double* red;
  float *t, *m, *a, *al, *dgp, *dgq, *gzs, *pd, *hd, *var, *w, *v, *u, *wc;
  float *Bp, *Bq;
};

__device__ SmemPlan carve(unsigned char* base, int T, int S, bool mats_in_smem, int ld, float* slot) {
  SmemPlan s;
  s.red = reinterpret_cast<double*>(base);
  float* f = reinterpret_cast<float*>(base + 32 * sizeof(double));
  s.t = f; f += T;
  s.m = f; f += T;
  s.a = f; f += T;
  s.al = f; f += T;
  s.dgp = f; f += T;
  s.dgq = f; f += T;
  s.gzs = f; f += T;
  s.pd = f; f += T;
  s.hd = f; f += T;
  s.var = f; f += T;
  s.w = f; f += (size_t)S * T;
  s.v = f; f += (size_t)S * T;
  s.u = f; f += (size_t)S * T;
  s.wc = f; f += (size_t)WCH * T;
  const size_t mat = (size_t)(T + 1) * ld;
  if (mats_in_smem) {
    s.Bp = f;
    s.Bq = f + mat;
  } else {
    s.Bp = slot;
    s.Bq = slot + mat;
  }
  return s;
}

// INSMEM: the two work matrices live in shared memory (a compile-time fact, so that their accesses are LDS / STS and not
// generic loads -- the run-time selection between a shared and a global carve-up made every matrix access generic).
template <int KERNEL, bool INSMEM>
__global__ void __launch_bounds__(NT) fwd_generic(Params P, int Tcap, int ld_cap, int) {
  constexpr bool mats_in_smem = INSMEM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int bad;
  const GpklDesc& d = P.d;
  if (P.skip_if_shared && P.prior_flag && *P.prior_flag != 0) return;  // fallback launch behind the V3 hot tier (gpkl_bidiag.cu)
  const int npairs = d.B * d.D;
  const int S = d.S;
  const float noise = d.noise;
  const float sig = (float)(1.0 - (double)noise);
  float* slot = mats_in_smem ? nullptr : (P.scratch ? P.scratch + (size_t)blockIdx.x * P.scratch_stride : nullptr);
  for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
    const int b = p / d.D, dd = p - b * d.D;
    const int T = P.lengths[b];
    const int64_t r0 = P.offsets[b];
    __syncthreads();
    if (T <= 0) {
      if (threadIdx.x == 0) {
        P.kl_pairs[p] = 0.0f;
        if (P.logdets) { P.logdets[2 * p] = 0.0f; P.logdets[2 * p + 1] = 0.0f; }
      }
      continue;
    }
    const int ld = (T & 1) ? T : T + 1;
    SmemPlan s = carve(smem_raw, Tcap, S, mats_in_smem, ld_cap, slot);
    if (threadIdx.x == 0) bad = 0;
    for (int i = threadIdx.x; i < T; i += NT) {
      s.t[i] = P.times[(size_t)b * d.T_max + i];
      s.m[i] = P.mean[(size_t)(r0 + i) * d.D + dd];
      s.a[i] = s.m[i];
      if (d.posterior == GPKL_POST_DIAG) s.var[i] = P.aux[(size_t)(r0 + i) * d.D + dd];  // logvar
      if (d.posterior == GPKL_POST_BIDIAG) {  // diagonal b_i > 0 and super-diagonal c_i of the precision factor B
        s.var[i] = P.aux[((size_t)(r0 + i) * d.D + dd) * 2];
        s.hd[i] = P.aux[((size_t)(r0 + i) * d.D + dd) * 2 + 1];
      }
    }
    for (int e = threadIdx.x; e < S * T; e += NT) {
      const int sidx = e / T, i = e - sidx * T;
      s.v[e] = eps_value(P, ((size_t)p * S + sidx) * d.T_max + i);
    }
    __syncthreads();
    const float lp = P.ell_p[dd];
    build_K<KERNEL>(s.Bp, ld, T, s.t, lp, sig, noise);
    chol_inplace(s.Bp, ld, T, s.dgp, &bad);
    double part = 0.0, ldp = 0.0, ldq = 0.0;
    if (d.posterior == GPKL_POST_GP) {
      const float lq = P.ell_q[dd];
      build_K<KERNEL>(s.Bq, ld, T, s.t, lq, sig, noise);
      chol_inplace(s.Bq, ld, T, s.dgq, &bad);
      // z_s = m + L_q eps_s  (row i per thread; LC(i,k) unit stride across threads)
      for (int i = threadIdx.x; i < T; i += NT) {
        for (int sidx = 0; sidx < S; ++sidx) {
          float acc = 0.0f;
          const float* ev = s.v + (size_t)sidx * T;
          for (int k = 0; k <= i; ++k) acc = fmaf(LC(s.Bq, i, k), ev[k], acc);
          P.z[((size_t)S * r0 + (size_t)sidx * T + i) * d.D + dd] = s.m[i] + acc;
        }
      }
      __syncthreads();
      trsm_cols(s.Bp, s.Bq, ld, T, s.dgp, s.a, true);
      // 2 KL = sum_i f(A_ii) + sum_{i>k} A_ik^2 + sum a_i^2
      for (int e = threadIdx.x; e < T * T; e += NT) {
        const int k = e / T, i = e - k * T;
        if (i > k) {
          const float x = LC(s.Bq, i, k);
          part += (double)x * (double)x;
        }
      }
      for (int i = threadIdx.x; i < T; i += NT) {
        part += diag_term((double)s.dgq[i] / (double)s.dgp[i]);
        part += (double)s.a[i] * (double)s.a[i];
        ldp += 2.0 * log((double)s.dgp[i]);
        ldq += 2.0 * log((double)s.dgq[i]);
      }
    } else if (d.posterior == GPKL_POST_BIDIAG) {
      // V3 (extension): q = N(m, (B^T B)^-1), B upper bidiagonal (diag s.var, super s.hd)
      //   KL = 1/2 [ ||L^-1 B^-1||_F^2 - T + log|K| + 2 sum log b_i + ||L^-1 m||^2 ],   z = m + B^-1 eps
      trinv(s.Bp, ld, T, s.dgp);
      bidiag_inverse(s.Bq, ld, T, s.var, s.hd);  // W = B^-1, upper triangular, W(i,j) = Bq[i*ld + j]
      for (int e = threadIdx.x; e < T * T; e += NT) {
        const int i = e / T, j = e - i * T;
        const int kmax = min(i, j);
        float mij = 0.0f;
        for (int k = 0; k <= kmax; ++k) mij = fmaf(XR(s.Bp, i, k), s.Bq[(size_t)k * ld + j], mij);
        part += (double)mij * (double)mij;
      }
      for (int i = threadIdx.x; i < T; i += NT) {
        float ai = 0.0f;
        for (int k = 0; k <= i; ++k) ai = fmaf(XR(s.Bp, i, k), s.m[k], ai);
        const double lb = log((double)s.var[i]), lpd = log((double)s.dgp[i]);
        part += (double)ai * (double)ai - 1.0 + 2.0 * lpd + 2.0 * lb;
        ldp += 2.0 * lpd;
        ldq -= 2.0 * lb;  // log|Sigma_q| = -2 sum log b
        for (int sidx = 0; sidx < S; ++sidx) {
          const float* ev = s.v + (size_t)sidx * T;
          float acc = s.m[i];
          for (int j = i; j < T; ++j) acc = fmaf(s.Bq[(size_t)i * ld + j], ev[j], acc);
          P.z[((size_t)S * r0 + (size_t)sidx * T + i) * d.D + dd] = acc;
        }
      }
    } else {  // GPKL_POST_DIAG
      trinv(s.Bp, ld, T, s.dgp);
      for (int i = threadIdx.x; i < T; i += NT) {
        float h = 0.0f, ai = 0.0f;
        for (int k = i; k < T; ++k) { const float x = XR(s.Bp, k, i); h = fmaf(x, x, h); }
        for (int k = 0; k <= i; ++k) ai = fmaf(XR(s.Bp, i, k), s.m[k], ai);
        const float lv = s.var[i];
        const float vv = expf(lv), sd = expf(0.5f * lv);
        part += (double)h * (double)vv - 1.0 - (double)lv + (double)ai * (double)ai + 2.0 * log((double)s.dgp[i]);
        ldp += 2.0 * log((double)s.dgp[i]);
        ldq += (double)lv;
        for (int sidx = 0; sidx < S; ++sidx)
          P.z[((size_t)S * r0 + (size_t)sidx * T + i) * d.D + dd] = s.m[i] + sd * s.v[(size_t)sidx * T + i];
      }
    }
    part = block_sum(part, s.red);
    if (P.logdets) {
      ldp = block_sum(ldp, s.red);
      ldq = block_sum(ldq, s.red);
    }
    if (threadIdx.x == 0) {
      P.kl_pairs[p] = (float)(0.5 * part);
      if (P.logdets) { P.logdets[2 * p] = (float)ldp; P.logdets[2 * p + 1] = (float)ldq; }
      if (bad && P.status) atomicAdd(P.status, 1);
    }
  }
}

template <int KERNEL, bool INSMEM>
__global__ void __launch_bounds__(NT) bwd_generic(Params P, int Tcap, int ld_cap, int) {
  constexpr bool mats_in_smem = INSMEM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int bad;
  const GpklDesc& d = P.d;
  const int npairs = d.B * d.D;
  const int S = d.S;
  const float noise = d.noise;
  const float sig = (float)(1.0 - (double)noise);
  const bool want_lp = (d.flags & GPKL_FLAG_GRAD_ELL_P) != 0;
  if (P.skip_if_shared && P.prior_flag && *P.prior_flag != 0) return;  // fallback launch behind the V3 hot tier
  float* slot = mats_in_smem ? nullptr : (P.scratch ? P.scratch + (size_t)blockIdx.x * P.scratch_stride : nullptr);
  const double g_sum = P.g_kl_sum ? *P.g_kl_sum : 1.0;
  for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
    const int b = p / d.D, dd = p - b * d.D;
    const int T = P.lengths[b];
    const int64_t r0 = P.offsets[b];
    __syncthreads();
    if (T <= 0) {
      if (threadIdx.x == 0) {
        if (P.gq_pairs) P.gq_pairs[p] = 0.0f;
        if (P.gp_pairs) P.gp_pairs[p] = 0.0f;
      }
      continue;
    }
    const int ld = (T & 1) ? T : T + 1;
    SmemPlan s = carve(smem_raw, Tcap, S, mats_in_smem, ld_cap, slot);
    const float g = (float)(g_sum + (P.g_kl_pairs ? (double)P.g_kl_pairs[p] : 0.0));
    if (threadIdx.x == 0) bad = 0;
    for (int i = threadIdx.x; i < T; i += NT) {
      s.t[i] = P.times[(size_t)b * d.T_max + i];
      s.m[i] = P.mean[(size_t)(r0 + i) * d.D + dd];
      if (d.posterior == GPKL_POST_DIAG) s.var[i] = P.aux[(size_t)(r0 + i) * d.D + dd];
      if (d.posterior == GPKL_POST_BIDIAG) {
        s.var[i] = P.aux[((size_t)(r0 + i) * d.D + dd) * 2];
        s.hd[i] = P.aux[((size_t)(r0 + i) * d.D + dd) * 2 + 1];
      }
      float gs = 0.0f;
      for (int sidx = 0; sidx < S; ++sidx) {
        const float gz = P.g_z ? P.g_z[((size_t)S * r0 + (size_t)sidx * T + i) * d.D + dd] : 0.0f;
        s.u[(size_t)sidx * T + i] = gz;
        gs += gz;
      }
      s.gzs[i] = gs;
    }
    for (int e = threadIdx.x; e < S * T; e += NT) {
      const int sidx = e / T, i = e - sidx * T;
      s.v[e] = eps_value(P, ((size_t)p * S + sidx) * d.T_max + i);
    }
    __syncthreads();
    const float lp = P.ell_p[dd];
    build_K<KERNEL>(s.Bp, ld, T, s.t, lp, sig, noise);
    chol_inplace(s.Bp, ld, T, s.dgp, &bad);
    trinv(s.Bp, ld, T, s.dgp);
    // a = X_p m ; alpha = X_p^T a = K_p^-1 m
    for (int i = threadIdx.x; i < T; i += NT) {
      float ai = 0.0f;
      for (int k = 0; k <= i; ++k) ai = fmaf(XR(s.Bp, i, k), s.m[k], ai);
      s.a[i] = ai;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < T; k += NT) {
      float al = 0.0f;
      for (int i = k; i < T; ++i) al = fmaf(XR(s.Bp, i, k), s.a[i], al);
      s.al[k] = al;
      P.g_mean[(size_t)(r0 + k) * d.D + dd] = g * al + s.gzs[k];
    }
    __syncthreads();

    if (d.posterior == GPKL_POST_GP) {
      const float lq = P.ell_q[dd];
      build_K<KERNEL>(s.Bq, ld, T, s.t, lq, sig, noise);
      chol_inplace(s.Bq, ld, T, s.dgq, &bad);
      // t1 = <K_p^-1, dK/d ell> for ell_q (and ell_p)
      double t1q, t1p;
      tri_contract<KERNEL>(s.Bp, s.Bp, ld, T, s.t, sig, lq, want_lp ? lp : -1.0f, t1q, t1p);
      // w_s = L_q^T g_z,s ; pd_i = 1/2 sum_s w_si eps_si - g/2
      for (int k = threadIdx.x; k < T; k += NT) {
        float pdk = 0.0f;
        for (int sidx = 0; sidx < S; ++sidx) {
          const float* uu = s.u + (size_t)sidx * T;
          float wk = 0.0f;
          for (int i = k; i < T; ++i) wk = fmaf(LC(s.Bq, i, k), uu[i], wk);
          s.w[(size_t)sidx * T + k] = wk;
          pdk = fmaf(wk, s.v[(size_t)sidx * T + k], pdk);
        }
        s.pd[k] = 0.5f * pdk - 0.5f * g;
      }
      __syncthreads();
      double t3 = 0.0;
      if (want_lp) {
        // A = X_p L_q into LC(Bp) (L_p is dead); then W = X_p^T [A | a] in column chunks and
        // t3 = <W W^T, dK_p/d ell_p>  ==  <K_p^-1 (K_q + m m^T) K_p^-1, dK_p/d ell_p>
        for (int e = threadIdx.x; e < T * T; e += NT) {
          const int c = e / T, i = e - c * T;
          if (i >= c) {
            float acc = 0.0f;
            for (int k = c; k <= i; ++k) acc = fmaf(XR(s.Bp, i, k), LC(s.Bq, k, c), acc);
            LC(s.Bp, i, c) = acc;
          }
        }
        __syncthreads();
        const float inv_sig = 1.0f / sig, il3 = 1.0f / (lp * lp * lp);
        for (int c0 = 0; c0 <= T; c0 += WCH) {
          const int nc = min(WCH, T + 1 - c0);
          for (int e = threadIdx.x; e < nc * T; e += NT) {
            const int cc = e / T, k = e - cc * T;
            const int c = c0 + cc;
            float acc = 0.0f;
            if (c < T) {
              for (int i = max(k, c); i < T; ++i) acc = fmaf(XR(s.Bp, i, k), LC(s.Bp, i, c), acc);
            } else {
              acc = s.al[k];
            }
            s.wc[(size_t)cc * T + k] = acc;
          }
          __syncthreads();
          for (int e = threadIdx.x; e < T * T; e += NT) {
            const int l = e / T, k = e - l * T;
            if (k == l) continue;
            float dot = 0.0f;
            for (int cc = 0; cc < nc; ++cc) dot = fmaf(s.wc[(size_t)cc * T + k], s.wc[(size_t)cc * T + l], dot);
            const float dt = s.t[k] - s.t[l];
            const float kv = kern_val<KERNEL>(dt, lp, sig);
            t3 += (double)dot * (double)kern_dell<KERNEL>(dt, kv, il3, inv_sig);
          }
          __syncthreads();
        }
      }
      trinv(s.Bq, ld, T, s.dgq);
      // C' = (Phi(sum_s w_s eps_s^T) - g/2 I) X_q  via running prefix sums down each column, into XR(Bp)
      for (int l = threadIdx.x; l < T; l += NT) {
        for (int i = l; i < T; ++i) XR(s.Bp, i, l) = s.pd[i] * XR(s.Bq, i, l);
        for (int sidx = 0; sidx < S; ++sidx) {
          const float* ww = s.w + (size_t)sidx * T;
          const float* vv = s.v + (size_t)sidx * T;
          float cum = 0.0f;
          for (int i = l; i < T; ++i) {
            XR(s.Bp, i, l) = fmaf(ww[i], cum, XR(s.Bp, i, l));
            cum = fmaf(vv[i], XR(s.Bq, i, l), cum);
          }
        }
      }
      __syncthreads();
      double t2, unused;
      tri_contract<KERNEL>(s.Bq, s.Bp, ld, T, s.t, sig, lq, -1.0f, t2, unused);
      const double gq = block_sum(0.5 * (double)g * t1q + t2, s.red);
      double gp = 0.0;
      if (want_lp) gp = block_sum(0.5 * (double)g * (t1p - t3), s.red);
      if (threadIdx.x == 0) {
        P.gq_pairs[p] = (float)gq;
        if (want_lp) P.gp_pairs[p] = (float)gp;
      }
    } else if (d.posterior == GPKL_POST_BIDIAG) {
      // W-bar = g K^-1 W + sum_s g_z,s eps_s^T (upper part), B-bar = -W^T W-bar W^T on the two diagonals,
      // plus g/b_i from the log-determinant term.
      bidiag_inverse(s.Bq, ld, T, s.var, s.hd);
      // H = K^-1 = X^T X, lower triangle, into the (dead) LC triangle of Bp
      for (int e = threadIdx.x; e < T * T; e += NT) {
        const int k = e / T, i = e - k * T;
        if (i >= k) {
          float dot = 0.0f;
          for (int r = i; r < T; ++r) dot = fmaf(XR(s.Bp, r, i), XR(s.Bp, r, k), dot);
          LC(s.Bp, i, k) = dot;
        }
      }
      __syncthreads();
      // U^T = triu(W-bar)^T into the XR triangle of Bq:  XR(Bq, l, k) = W-bar[k][l], k <= l
      for (int e = threadIdx.x; e < T * T; e += NT) {
        const int l = e / T, k = e - l * T;
        if (k <= l) {
          float q = 0.0f;
          for (int a = 0; a <= l; ++a) {
            const float h = (k >= a) ? LC(s.Bp, k, a) : LC(s.Bp, a, k);
            q = fmaf(h, s.Bq[(size_t)a * ld + l], q);
          }
          float ge = 0.0f;
          for (int sidx = 0; sidx < S; ++sidx) ge = fmaf(s.u[(size_t)sidx * T + k], s.v[(size_t)sidx * T + l], ge);
          XR(s.Bq, l, k) = g * q + ge;
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < T; i += NT) {
        float bb = 0.0f, cc = 0.0f;
        for (int l = i; l < T; ++l) {
          float pl = 0.0f;  // (W^T triu(W-bar))[i][l]
          for (int k = 0; k <= i; ++k) pl = fmaf(s.Bq[(size_t)k * ld + i], XR(s.Bq, l, k), pl);
          bb = fmaf(s.Bq[(size_t)i * ld + l], pl, bb);
          if (l >= i + 1) cc = fmaf(s.Bq[(size_t)(i + 1) * ld + l], pl, cc);
        }
        P.g_aux[((size_t)(r0 + i) * d.D + dd) * 2] = -bb + g / s.var[i];
        P.g_aux[((size_t)(r0 + i) * d.D + dd) * 2 + 1] = (i + 1 < T) ? -cc : 0.0f;
      }
    } else {  // GPKL_POST_DIAG
      for (int i = threadIdx.x; i < T; i += NT) {
        float h = 0.0f;
        for (int k = i; k < T; ++k) { const float x = XR(s.Bp, k, i); h = fmaf(x, x, h); }
        const float lv = s.var[i];
        const float vv = expf(lv), sd = expf(0.5f * lv);
        float ge = 0.0f;
        for (int sidx = 0; sidx < S; ++sidx) ge = fmaf(s.u[(size_t)sidx * T + i], s.v[(size_t)sidx * T + i], ge);
        P.g_aux[(size_t)(r0 + i) * d.D + dd] = 0.5f * g * (h * vv - 1.0f) + 0.5f * sd * ge;
        s.hd[i] = vv;  // variance, reused below
      }
      if (want_lp) {
        __syncthreads();
        // H = K^-1 = X^T X, full symmetric, into Bq (row k at Bq[k*ld + l])
        for (int e = threadIdx.x; e < T * T; e += NT) {
          const int l = e / T, k = e - l * T;
          float dot = 0.0f;
          for (int i = max(k, l); i < T; ++i) dot = fmaf(XR(s.Bp, i, k), XR(s.Bp, i, l), dot);
          s.Bq[(size_t)l * ld + k] = dot;
        }
        __syncthreads();
        const float inv_sig = 1.0f / sig, il3 = 1.0f / (lp * lp * lp);
        double acc = 0.0;
        for (int e = threadIdx.x; e < T * T; e += NT) {
          const int l = e / T, k = e - l * T;
          if (k == l) continue;
          float dot = 0.0f;
          for (int i = 0; i < T; ++i) dot = fmaf(s.Bq[(size_t)i * ld + k] * s.hd[i], s.Bq[(size_t)i * ld + l], dot);
          const float dt = s.t[k] - s.t[l];
          const float kv = kern_val<KERNEL>(dt, lp, sig);
          const float dk = kern_dell<KERNEL>(dt, kv, il3, inv_sig);
          acc += ((double)s.Bq[(size_t)l * ld + k] - (double)dot - (double)s.al[k] * (double)s.al[l]) * (double)dk;
        }
        const double gp = block_sum(0.5 * (double)g * acc, s.red);
        if (threadIdx.x == 0) P.gp_pairs[p] = (float)gp;
      }
    }
    if (threadIdx.x == 0 && bad && P.status) atomicAdd(P.status, 1);
  }
}

}  // namespace

size_t generic_smem_bytes(int T, int S, bool mats_in_smem) {
  const int ld = (T & 1) ? T : T + 1;
  size_t fl = (size_t)10 * T + (size_t)3 * S * T + (size_t)WCH * T;
  if (mats_in_smem) fl += (size_t)2 * (T + 1) * ld;
  return 32 * sizeof(double) + fl * sizeof(float);
}

size_t generic_slot_floats(int T) {
  const int ld = (T & 1) ? T : T + 1;
  return (size_t)2 * (T + 1) * ld;
}

static int generic_grid(const GpklDesc& d, bool in_smem, size_t smem) {
  const int npairs = d.B * d.D;
  int per_sm = in_smem ? (int)((size_t)(227 * 1024) / (smem + 1024)) : 1;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  const int cap = 148 * per_sm;
  return npairs < cap ? (npairs > 0 ? npairs : 1) : cap;
}

int generic_slots(const GpklDesc& d) {
  const size_t smem = generic_smem_bytes(d.T_max, d.S, true);
  if (smem <= kMaxDynSmem) return 0;
  return 148;
}

template <typename K>
static cudaError_t launch(K kern, const Params& P, int grid, size_t smem, int in_smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int T = P.d.T_max;
  const int ld = (T & 1) ? T : T + 1;
  prof_begin(P.g_mean != nullptr, st);
  kern<<<grid, NT, smem, st>>>(P, T, ld, in_smem);
  prof_end(P.g_mean != nullptr, st);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_generic(const Params& P, bool backward, cudaStream_t st) {
  const GpklDesc& d = P.d;
  bool in_smem = generic_smem_bytes(d.T_max, d.S, true) <= kMaxDynSmem;
  const size_t smem = generic_smem_bytes(d.T_max, d.S, in_smem);
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  const int grid = in_smem ? generic_grid(d, true, smem) : min(148, max(1, d.B * d.D));
  if (!backward) {
    if (in_smem)
      return d.kernel == GPKL_KERNEL_RBF ? launch(fwd_generic<GPKL_KERNEL_RBF, true>, P, grid, smem, in_smem, st)
                                         : launch(fwd_generic<GPKL_KERNEL_CAUCHY, true>, P, grid, smem, in_smem, st);
    return d.kernel == GPKL_KERNEL_RBF ? launch(fwd_generic<GPKL_KERNEL_RBF, false>, P, grid, smem, in_smem, st)
                                       : launch(fwd_generic<GPKL_KERNEL_CAUCHY, false>, P, grid, smem, in_smem, st);
  }
  if (in_smem)
    return d.kernel == GPKL_KERNEL_RBF ? launch(bwd_generic<GPKL_KERNEL_RBF, true>, P, grid, smem, in_smem, st)
                                       : launch(bwd_generic<GPKL_KERNEL_CAUCHY, true>, P, grid, smem, in_smem, st);
  return d.kernel == GPKL_KERNEL_RBF ? launch(bwd_generic<GPKL_KERNEL_RBF, false>, P, grid, smem, in_smem, st)
                                     : launch(bwd_generic<GPKL_KERNEL_CAUCHY, false>, P, grid, smem, in_smem, st);
}

}  // namespace gpkl
