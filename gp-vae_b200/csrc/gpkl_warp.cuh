// Warp tier: T <= 64.  A group of LP lanes owns one (sequence, latent-dim) pair; each lane keeps R rows
// (or columns) of the T x T matrix in REGISTERS (TM = LP*R >= T), so the O(T^3) work is FFMA on
// registers and the only shared-memory traffic is one 128-bit broadcast load per 4*R FMAs.
//
//   (LP,R) = (8,1) T<=8 [4 pairs/warp], (16,1) T<=16 [2 pairs/warp], (32,1) T<=32, (16,3) T<=48
//   [2 pairs/warp], (32,2) T<=64.   Lane `lig` of a group owns rows/columns lig + LP*j, j < R.
//
// Forward (V1):   rows of K_p built in registers -> right-looking Cholesky (column broadcast through a
//   double-buffered shared vector, one __syncwarp per column) -> a = L_p^-1 m by a shuffle column sweep
//   -> L_p rows parked in shared memory (packed lower, 16 B aligned rows) -> same for K_q, z = m + L_q eps
//   from the register rows -> L_q transposed through shared memory so each lane owns COLUMNS -> forward
//   substitution A = L_p^-1 L_q with L_p as 128-bit broadcast operand -> KL = 1/2[sum_offdiag A^2 +
//   sum f(A_ii) + |a|^2] with the cancelling diagonal part in float64.
// Backward (V1):  recompute the factors the same way; X_p = L_p^-1 and X_q = L_q^-1 as register columns;
//   alpha = X_p^T a; C' = (Phi(w eps^T) - g/2 I) X_q by a running prefix sum down each register column;
//   d/d ell_q = sum_{k != l} dK_q(k,l) [ g/2 (X_p^T X_p)_kl + (X_q^T C')_kl ]  with the second operand's
//   columns broadcast from shared memory.  Nothing T x T touches HBM.
//
// Reference replaced: tf_kernel / gp_vae_sample / gp_kl_div (src/Models/Full_GP_VAE_dynamic_time.py:
// 149-172, :174-195, :242-260), V2 gp_kl_div / vae_sample (src/Models/VAE_GPprior_diag_cov.py:64-71,
// :100-119) and TF autodiff through them (:361).
#pragma once
#include "gpkl_common.cuh"
#include <string.h>

#include "gpkl_launch.h"

namespace gpkl {
namespace {

#ifndef GPKL_WPC
#define GPKL_WPC 4
#endif
constexpr int WPC = GPKL_WPC;  // warps per CTA

// Packed lower-triangular storage with 16-byte aligned rows: row i has capacity 4*(i/4+1) floats.
__host__ __device__ constexpr int poff(int i) { return 8 * (i >> 2) * ((i >> 2) + 1) + (i & 3) * 4 * ((i >> 2) + 1); }

template <int LP, int R>
struct Geo {
  static constexpr int TM = LP * R;
  static constexpr int PK = poff(TM);
  static constexpr int G = 32 / LP;  // groups (pairs) per warp
  // packed FFMA2 (see fma2): PACK = false selects the scalar-FMA form of the same arithmetic
  static constexpr bool PACK = true;
  // per-group shared floats: 2 packed buffers, 2 column buffers, ts, dgp, dgq, dinv, as, pd + 3*S*TM
  static constexpr int fixed_floats() { return 2 * PK + 2 * TM + 6 * TM; }
};

template <int LP>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = LP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LP, int R>
struct Smem {
  static constexpr int TM = LP * R;
  float *bufA, *bufB, *col, *ts, *dgp, *dgq, *dinv, *as, *pd, *vS, *uS, *wS;
  __device__ Smem(float* base, int S) {
    bufA = base; base += Geo<LP, R>::PK;
    bufB = base; base += Geo<LP, R>::PK;
    col = base; base += 2 * TM;
    ts = base; base += TM;
    dgp = base; base += TM;
    dgq = base; base += TM;
    dinv = base; base += TM;
    as = base; base += TM;
    pd = base; base += TM;
    vS = base; base += S * TM;
    uS = base; base += S * TM;
    wS = base;
  }
};

// Rows of K(t, ell) (lower triangle incl. diagonal, zeros above; identity on padded rows/cols).
template <int LP, int R, int KERNEL>
__device__ __forceinline__ void build_rows(float (&a)[R][LP * R], const float (&trow)[R], const float* __restrict__ ts,
                                           int lig, int T, int Tw, float ell, float sig, float noise) {
  constexpr int TM = LP * R;
  const KernC<KERNEL> kc(ell, sig);
#pragma unroll
  for (int k4 = 0; k4 < TM; k4 += 4) {
    if (k4 < Tw) {
      const float4 t4 = *reinterpret_cast<const float4*>(ts + k4);
      const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k4 + e;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int r = lig + LP * j;
          if (k > LP * j + LP - 1) {  // above the diagonal for every lane of this slot (compile-time): no exp
            a[j][k] = 0.0f;
            continue;
          }
          float v = kc.val(trow[j] - tv[e]);
          if (k == r) v += noise;
          a[j][k] = (k <= r && r < T) ? v : ((k == r) ? 1.0f : 0.0f);
        }
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
#pragma unroll
        for (int j = 0; j < R; ++j) a[j][k4 + e] = (k4 + e == lig + LP * j) ? 1.0f : 0.0f;
    }
  }
}

// Right-looking Cholesky on register rows.  col: 2*TM floats of shared memory (double buffered).
template <int LP, int R>
__device__ __forceinline__ void chol_rows(float (&a)[R][LP * R], int lig, int T, int Tw, float* __restrict__ col,
                                          float* __restrict__ dg, int& bad) {
  constexpr int TM = LP * R;
  // (column steps are guarded in groups of 4: a padded column inside a live group is an identity column and
  //  factors to itself)
#pragma unroll
  for (int j = 0; j < TM; ++j) {
    if ((j & ~3) < Tw) {
      float* cb = col + (j & 1) * TM;
#pragma unroll
      for (int jj = 0; jj < R; ++jj) cb[lig + LP * jj] = a[jj][j];
      __syncwarp();
      const float d = cb[j];
      float rs = rsqrtf(d);
      rs = rs * fmaf(-0.5f * d, rs * rs, 1.5f);  // one Newton step: 1/sqrt(d) to ~1 ulp, branch-free
      const float sd = d * rs;
      const float rd = rs * rs;
      if (j < T && !(d > 0.0f)) bad = 1;
      float s[R];
#pragma unroll
      for (int jj = 0; jj < R; ++jj) s[jj] = (lig + LP * jj > j) ? a[jj][j] * rd : 0.0f;
      float ns[R];
#pragma unroll
      for (int jj = 0; jj < R; ++jj) ns[jj] = -s[jj];
      // (the trailing update is guarded per 16 columns only: finer guards cost more in branches than the
      //  skipped identity-padding columns save)
#pragma unroll
      for (int k16 = (j + 1) & ~15; k16 < TM; k16 += 16) {
        if (k16 < Tw) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int k4 = k16 + 4 * q;
            if (k4 < ((j + 1) & ~3) || k4 >= TM) continue;  // compile-time after unrolling
            const float4 c4 = *reinterpret_cast<const float4*>(cb + k4);
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
            for (int e = 0; e < 4; e += 2) {
              if (k4 + e > j) {  // both columns of the pair are trailing columns: one packed FMA per row
#pragma unroll
                for (int jj = 0; jj < R; ++jj) fma2<Geo<LP, R>::PACK>(a[jj][k4 + e], a[jj][k4 + e + 1], ns[jj], ns[jj], cv[e], cv[e + 1]);
              } else if (k4 + e + 1 > j) {
#pragma unroll
                for (int jj = 0; jj < R; ++jj) a[jj][k4 + e + 1] = fmaf(ns[jj], cv[e + 1], a[jj][k4 + e + 1]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int jj = 0; jj < R; ++jj) a[jj][j] *= rs;
      if (lig == (j % LP)) {
        a[j / LP][j] = sd;
        dg[j] = sd;
      }
    } else if (lig == (j % LP)) {
      dg[j] = 1.0f;  // identity padding beyond the longest sequence of this warp
    }
  }
}

// b <- L^-1 b by a column sweep over register rows; solution also written to out[] (shared).
template <int LP, int R>
__device__ __forceinline__ void sweep_vec(const float (&a)[R][LP * R], float (&b)[R], int lig, int Tw,
                                          const float* __restrict__ dinv, float* __restrict__ out) {
  constexpr int TM = LP * R;
#pragma unroll
  for (int k = 0; k < TM; ++k) {
    {  // (no guard: padded rows/columns are identity and b is zero there, so the sweep is a no-op on them)
      const float cand = b[k / LP] * dinv[k];
      const float xk = __shfl_sync(0xffffffffu, cand, k % LP, LP);
#pragma unroll
      for (int jj = 0; jj < R; ++jj) b[jj] = fmaf(-a[jj][k], xk, b[jj]);
      if (lig == (k % LP)) out[k] = xk;
    }
  }
}

// Park register rows in shared memory, packed lower (padding in the last 4-group of a row is zero).
template <int LP, int R>
__device__ __forceinline__ void store_rows(const float (&a)[R][LP * R], int lig, float* __restrict__ buf) {
  constexpr int TM = LP * R;
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    const int r = lig + LP * jj;
    float* row = buf + poff(r);
#pragma unroll
    for (int g = 0; g < TM / 4; ++g) {
      if (4 * g <= r)
        *reinterpret_cast<float4*>(row + 4 * g) = make_float4(a[jj][4 * g], a[jj][4 * g + 1], a[jj][4 * g + 2], a[jj][4 * g + 3]);
    }
  }
}

// x[jj][i] <- element (i, c_jj) of the packed lower matrix in buf (0 above the diagonal).
template <int LP, int R>
__device__ __forceinline__ void load_cols(float (&x)[R][LP * R], int lig, const float* __restrict__ buf) {
  constexpr int TM = LP * R;
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int c = lig + LP * jj;
      x[jj][i] = (i >= c) ? buf[poff(i) + c] : 0.0f;
    }
}

// In place forward substitution on register columns: x <- L^-1 x, L packed lower in shared memory.
template <int LP, int R>
__device__ __forceinline__ void solve_cols(float (&x)[R][LP * R], const float* __restrict__ Lpk,
                                           const float* __restrict__ dinv, int Tw) {
  constexpr int TM = LP * R;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    if ((i & ~3) < Tw) {  // groups of 4 rows: a padded row inside a live group is an identity row
      const float* row = Lpk + poff(i);
      float acc[R][2];
#pragma unroll
      for (int jj = 0; jj < R; ++jj) acc[jj][0] = acc[jj][1] = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < i; k4 += 4) {
        const float4 l4 = *reinterpret_cast<const float4*>(row + k4);
        const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int e = 0; e < 4; e += 2) {
          if (k4 + e + 1 < i) {  // both rows of the pair are already solved: one packed FMA per column
#pragma unroll
            for (int jj = 0; jj < R; ++jj) fma2<Geo<LP, R>::PACK>(acc[jj][0], acc[jj][1], lv[e], lv[e + 1], x[jj][k4 + e], x[jj][k4 + e + 1]);
          } else if (k4 + e < i) {
#pragma unroll
            for (int jj = 0; jj < R; ++jj) acc[jj][0] = fmaf(lv[e], x[jj][k4 + e], acc[jj][0]);
          }
        }
      }
      const float di = dinv[i];
#pragma unroll
      for (int jj = 0; jj < R; ++jj) x[jj][i] = (x[jj][i] - (acc[jj][0] + acc[jj][1])) * di;
    }
  }
}

// Column c of a lower-triangular matrix (zeros for i < c) -> packed row TM-1-c, reversed in i, so that
// other lanes can stream it back with aligned 128-bit broadcast loads at static register positions.
// Element i sits at position (TM-1-i)^1: reversed by PAIRS, so that a loaded pair (q.x, q.y) lines up with the
// register pair (x[even], x[odd]) of the packed FMA in contract_cols.
__device__ __forceinline__ int poff_dyn(int i) {
  const int q = i >> 2, r = i & 3;
  return 8 * q * (q + 1) + 4 * r * (q + 1);
}

template <int LP, int R>
__device__ __forceinline__ void store_cols_rev(const float (&x)[R][LP * R], int lig, float* __restrict__ buf) {
  constexpr int TM = LP * R;
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    const int ip = TM - 1 - (lig + LP * jj);
    float* row = buf + poff_dyn(ip);
#pragma unroll
    for (int g = 0; g < TM / 4; ++g) {
      if (4 * g <= ip)
        *reinterpret_cast<float4*>(row + 4 * g) =
            make_float4(x[jj][TM - 2 - 4 * g], x[jj][TM - 1 - 4 * g], x[jj][TM - 4 - 4 * g], x[jj][TM - 3 - 4 * g]);
    }
  }
}

// sum_{l != c} dK(c,l)/d ell * <x_c, M_l>  for each register column c; M columns in reversed-packed shared.
// KINV (shared-prior path): kinv = this sequence's K_p^-1 (row-major, stride TM; symmetric, lane c reads ITS row c),
// the weight of dK(c,l) becomes hg * K_p^-1(c,l) + <x_c, M_l> (hg = g/2: the prior-side term t1 rides along), and
// alpha[jj] = sum_l K_p^-1(c,l) mvec[l] is accumulated on the way.
template <int LP, int R, int KERNEL, bool KINV = false>
__device__ __forceinline__ float contract_cols(const float (&x)[R][LP * R], const float* __restrict__ M,
                                               const float* __restrict__ ts, const float (&tcol)[R], int lig, int T,
                                               float ell, float sig, const float* __restrict__ kinv = nullptr,
                                               float hg = 0.0f, const float* __restrict__ mvec = nullptr,
                                               float* alpha = nullptr) {
  constexpr int TM = LP * R;
  const KernC<KERNEL> kc(ell, sig);
  float acc = 0.0f;
  float al[R];
#pragma unroll
  for (int jj = 0; jj < R; ++jj) al[jj] = 0.0f;
  // Rows of M are grouped by capacity: rows ip = 4q .. 4q+3 (l = TM-1-ip) hold q+1 four-groups.  The q loop is
  // unrolled (compile-time trip counts and offsets, no per-group guards), the four rows of a group are a real loop
  // (the same code four times: instruction-cache hits instead of straight-line fetch).
#pragma unroll
  for (int q = TM / 4 - 1; q >= 0; --q) {
    const float* base = M + 8 * q * (q + 1);  // poff(4q)
#pragma unroll 1
    for (int r = 3; r >= 0; --r) {
      const int l = TM - 1 - (4 * q + r);
      if (l < T) {
        float kin[R];
        if (KINV) {
#pragma unroll
          for (int jj = 0; jj < R; ++jj) kin[jj] = __ldg(kinv + (size_t)(lig + LP * jj) * TM + l);
        }
        const float* row = base + 4 * r * (q + 1);
        float dot[R], dot1[R];
#pragma unroll
        for (int jj = 0; jj < R; ++jj) dot[jj] = dot1[jj] = 0.0f;
#pragma unroll
        for (int g = 0; g <= q; ++g) {
          const float4 qv = *reinterpret_cast<const float4*>(row + 4 * g);
#pragma unroll
          for (int jj = 0; jj < R; ++jj) {
            fma2<Geo<LP, R>::PACK>(dot[jj], dot1[jj], qv.x, qv.y, x[jj][TM - 2 - 4 * g], x[jj][TM - 1 - 4 * g]);
            fma2<Geo<LP, R>::PACK>(dot[jj], dot1[jj], qv.z, qv.w, x[jj][TM - 4 - 4 * g], x[jj][TM - 3 - 4 * g]);
          }
        }
#pragma unroll
        for (int jj = 0; jj < R; ++jj) dot[jj] += dot1[jj];
        const float tl = ts[l];
        if (KINV) {
          const float ml = mvec[l];
#pragma unroll
          for (int jj = 0; jj < R; ++jj) {
            dot[jj] = fmaf(hg, kin[jj], dot[jj]);
            al[jj] = fmaf(kin[jj], ml, al[jj]);
          }
        }
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
          const int c = lig + LP * jj;
          const float dt = tcol[jj] - tl;
          const float dk = kc.dell(dt, kc.val_fast(dt));
          if (c != l && c < T) acc = fmaf(dot[jj], dk, acc);
        }
      }
    }
  }
  if (KINV) {
#pragma unroll
    for (int jj = 0; jj < R; ++jj) alpha[jj] = al[jj];
  }
  return acc;
}


// ---- d/d ell_p in the warp tier (GPKL_FLAG_GRAD_ELL_P; trainable prior of Full_GP_VAE_fixed_for_MovMnist.py:96) ------------
//   K_p-bar = g/2 (K_p^-1 - K_p^-1 (K_q + m m^T) K_p^-1),   d/d ell_p = <K_p-bar, dK_p/d ell_p> = g/2 (t1p - t3)
//   t1p = <X_p^T X_p, dK_p>  (contract_cols with the prior's length scale),
//   t3  = sum_j v_j^T dK_p v_j + alpha^T dK_p alpha,   V = K_p^-1 L_q = X_p^T A,  A = L_p^-1 L_q,  alpha = K_p^-1 m.
// V(l, c) = <A_c, X_p,l> is the dot product contract_cols forms (register column x table column); its values are parked in a
// TM x TM table, reloaded as register columns, and the quadratic forms run against a packed table of dK_p/d ell_p.

// Vs[l*TM + c] = <x_c, M_l> for every l (0 for l >= T); M columns reversed-packed as for contract_cols.
template <int LP, int R>
__device__ __forceinline__ void dots_cols(const float (&x)[R][LP * R], const float* __restrict__ M, int lig, int T,
                                          float* __restrict__ Vs) {
  constexpr int TM = LP * R;
#pragma unroll
  for (int q = TM / 4 - 1; q >= 0; --q) {
    const float* base = M + 8 * q * (q + 1);
#pragma unroll 1
    for (int r = 3; r >= 0; --r) {
      const int l = TM - 1 - (4 * q + r);
      const float* row = base + 4 * r * (q + 1);
      float dot[R], dot1[R];
#pragma unroll
      for (int jj = 0; jj < R; ++jj) dot[jj] = dot1[jj] = 0.0f;
#pragma unroll
      for (int g = 0; g <= q; ++g) {
        const float4 qv = *reinterpret_cast<const float4*>(row + 4 * g);
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
          fma2<Geo<LP, R>::PACK>(dot[jj], dot1[jj], qv.x, qv.y, x[jj][TM - 2 - 4 * g], x[jj][TM - 1 - 4 * g]);
          fma2<Geo<LP, R>::PACK>(dot[jj], dot1[jj], qv.z, qv.w, x[jj][TM - 4 - 4 * g], x[jj][TM - 3 - 4 * g]);
        }
      }
#pragma unroll
      for (int jj = 0; jj < R; ++jj) Vs[l * TM + lig + LP * jj] = (l < T) ? dot[jj] + dot1[jj] : 0.0f;
    }
  }
}

// Packed lower table (poff rows, zero diagonal and padding) of dK(l,k)/d ell, rows dealt to the lanes; zero beyond T.
template <int LP, int R, int KERNEL>
__device__ __forceinline__ void dk_table(float* __restrict__ Kd, const float* __restrict__ ts, int lig, int T, float ell, float sig) {
  constexpr int TM = LP * R;
  const KernC<KERNEL> kc(ell, sig);
  for (int l = lig; l < TM; l += LP) {
    float* row = Kd + poff_dyn(l);
    const float tl = ts[l];
    const int cap = 4 * ((l >> 2) + 1);
    for (int k = 0; k < cap; ++k) {
      const float dt = tl - ts[k];
      row[k] = (k < l && l < T) ? kc.dell(dt, kc.val_fast(dt)) : 0.0f;
    }
  }
}

// sum over this lane's columns c of  2 sum_{l > k} dK(l,k) V(l,c) V(k,c)
template <int LP, int R>
__device__ __forceinline__ float quad_cols(const float* __restrict__ Vs, const float* __restrict__ Kd, int lig) {
  constexpr int TM = LP * R;
  float v[R][TM];
#pragma unroll
  for (int l = 0; l < TM; ++l)
#pragma unroll
    for (int jj = 0; jj < R; ++jj) v[jj][l] = Vs[l * TM + lig + LP * jj];
  float qf = 0.0f;
#pragma unroll
  for (int l = 1; l < TM; ++l) {
    const float* row = Kd + poff(l);
    float acc[R][2];
#pragma unroll
    for (int jj = 0; jj < R; ++jj) acc[jj][0] = acc[jj][1] = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 <= l; k4 += 4) {  // (the diagonal entry and the padding of the row are zeros)
      const float4 d4 = *reinterpret_cast<const float4*>(row + k4);
#pragma unroll
      for (int jj = 0; jj < R; ++jj) {
        fma2<Geo<LP, R>::PACK>(acc[jj][0], acc[jj][1], d4.x, d4.y, v[jj][k4], v[jj][k4 + 1]);
        fma2<Geo<LP, R>::PACK>(acc[jj][0], acc[jj][1], d4.z, d4.w, v[jj][k4 + 2], v[jj][k4 + 3]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < R; ++jj) qf = fmaf(v[jj][l], acc[jj][0] + acc[jj][1], qf);
  }
  return 2.0f * qf;
}

// this lane's share of  2 sum_{l > k} dK(l,k) al[l] al[k]  (rows l = the lane's columns)
template <int LP, int R>
__device__ __forceinline__ float quad_vec(const float* __restrict__ al, const float* __restrict__ Kd, int lig) {
  float qf = 0.0f;
#pragma unroll
  for (int jj = 0; jj < R; ++jj) {
    const int l = lig + LP * jj;
    const float* row = Kd + poff_dyn(l);
    float acc = 0.0f;
    for (int k4 = 0; k4 <= l; k4 += 4) {
      const float4 d4 = *reinterpret_cast<const float4*>(row + k4);
      const float4 a4 = *reinterpret_cast<const float4*>(al + k4);
      acc = fmaf(d4.x, a4.x, fmaf(d4.y, a4.y, fmaf(d4.z, a4.z, fmaf(d4.w, a4.w, acc))));
    }
    qf = fmaf(al[l], acc, qf);
  }
  return 2.0f * qf;
}

__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- shared-prior fast path ------------------------------------------------------------------------------------
// The reference's prior length scales are one constant for all latent dims (prior_time_chars,
// Full_GP_VAE_dynamic_time.py:114), so the D pairs of a sequence share K_p.  A pre-pass (prior_inv64_small_kernel,
// gpkl_prior64.cu: one CTA per SEQUENCE, float64 sweep inverse) leaves a record in the workspace; the per-pair kernels
// then only factor K_q:
//   forward   KL = 1/2 [tr(K_p^-1 (K_q + m m^T)) - T + log|K_p| - log|K_q|] with the trace taken entrywise against the
//             FLOAT64 K_p^-1 while K_q is generated (the reference's own formula, :250-259)
//   backward  alpha = K_p^-1 m and t1 = <K_p^-1, dK_q/d ell> straight from K_p^-1 rows rounded to float32.
// Backward record (floats): [KI, KI + TM*TM) K_p^-1, full symmetric, identity on the padding ([0, KI) unused).
template <int LP, int R>
struct PriorRec {
  static constexpr int TM = LP * R, PK = Geo<LP, R>::PK;
  static constexpr int KI = PK + TM, SIZE = PK + TM + TM * TM;
};

struct PairInfo {
  int p, b, d, T, lig;
  long long r0;
  bool active;
};

template <int LP>
__device__ __forceinline__ PairInfo pair_info(const Params& P) {
  PairInfo pi;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int G = 32 / LP;
  pi.lig = lane % LP;
  pi.p = (blockIdx.x * (blockDim.x >> 5) + warp) * G + lane / LP;  // (blockDim.x / 32 = WPC, fewer for the d/d ell_p variant)
  pi.active = pi.p < P.d.B * P.d.D;
  pi.b = pi.active ? pi.p / P.d.D : 0;
  pi.d = pi.active ? pi.p - pi.b * P.d.D : 0;
  pi.T = pi.active ? P.lengths[pi.b] : 0;
  pi.r0 = pi.active ? P.offsets[pi.b] : 0;
  return pi;
}

template <int LP, int R, int KERNEL, int POST>
__global__ void __launch_bounds__(WPC * 32) fwd_warp(Params P, int group_floats) {
  constexpr int TM = LP * R;
  extern __shared__ __align__(16) float smem_f[];
  const GpklDesc& d = P.d;
  const int S = d.S;
  const PairInfo pi = pair_info<LP>(P);
  const int gslot = (threadIdx.x >> 5) * (32 / LP) + (threadIdx.x & 31) / LP;
  Smem<LP, R> sm(smem_f + (size_t)gslot * group_floats, S);
  const int T = pi.T, lig = pi.lig;
  const int Tw = warp_max(T);
  if (Tw == 0) {
    if (pi.active && lig == 0) {
      P.kl_pairs[pi.p] = 0.0f;
      if (P.logdets) { P.logdets[2 * pi.p] = 0.0f; P.logdets[2 * pi.p + 1] = 0.0f; }
    }
    return;
  }
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  float trow[R], mrow[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int r = lig + LP * j;
    const bool ok = r < T;
    trow[j] = ok ? P.times[(size_t)pi.b * d.T_max + r] : 0.0f;
    mrow[j] = ok ? P.mean[(size_t)(pi.r0 + r) * d.D + pi.d] : 0.0f;
    sm.ts[r] = trow[j];
    for (int s = 0; s < S; ++s) sm.vS[s * TM + r] = ok ? eps_value(P, ((size_t)pi.p * S + s) * d.T_max + r) : 0.0f;
  }
  __syncwarp();
  int bad = 0;
  float a[R][TM];
  // shared-prior fast path: the pre-pass found one ell_p for all latent dims and left this sequence's record
  const bool shared = (POST == GPKL_POST_GP) && P.prior != nullptr && *P.prior_flag != 0;
  if (!shared) {
    const float lp = pi.active ? P.ell_p[pi.d] : 1.0f;
    build_rows<LP, R, KERNEL>(a, trow, sm.ts, lig, T, Tw, lp, sig, noise);
    chol_rows<LP, R>(a, lig, T, Tw, sm.col, sm.dgp, bad);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j) sm.dinv[lig + LP * j] = 1.0f / sm.dgp[lig + LP * j];
    __syncwarp();
    float bvec[R];
#pragma unroll
    for (int j = 0; j < R; ++j) bvec[j] = mrow[j];
    sweep_vec<LP, R>(a, bvec, lig, Tw, sm.dinv, sm.as);
    store_rows<LP, R>(a, lig, sm.bufA);
    __syncwarp();
  }
  double part = 0.0, ldp = 0.0, ldq = 0.0;
  if (POST == GPKL_POST_GP) {
    const float lq = pi.active ? P.ell_q[pi.d] : 1.0f;
    build_rows<LP, R, KERNEL>(a, trow, sm.ts, lig, T, Tw, lq, sig, noise);
    double tr0 = 0.0, tr1 = 0.0, tr2 = 0.0, tr3 = 0.0, ldp_rec = 0.0;
    if (shared) {
      // Shared prior: the record of this sequence is K_p^-1 in FLOAT64 (lower triangle, column-major, pitch prior64_pitch(T_max), diagonal
      // halved) + log|K_p| (prior_inv64_small_kernel, gpkl_prior64.cu), and the KL is the reference's own formula
      // (Full_GP_VAE_dynamic_time.py:250-259)
      //     KL = 1/2 [ tr(K_p^-1 (K_q + m m^T)) - T + log|K_p| - log|K_q| ],   tr = 2 sum_{k<=r} Kinv'_rk (K_q,rk + m_r m_k),
      // accumulated HERE from the rows of K_q this lane has just built (before the factorisation overwrites them): no
      // triangular product A = L_p^-1 L_q, no transposition of L_q, no a = L_p^-1 m.  Three FP64 operations per entry.
      griddep_wait();
      const double* __restrict__ kinv = reinterpret_cast<const double*>(P.prior + (size_t)pi.b * P.prior_stride);
      double* m64 = reinterpret_cast<double*>(sm.bufA);  // the mean as doubles (bufA is idle on this path)
      constexpr int TRB = LP < 16 ? LP : 16;
      const int ldk = prior64_pitch(d.T_max);
#pragma unroll
      for (int j = 0; j < R; ++j) m64[lig + LP * j] = (double)mrow[j];
      __syncwarp();
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int r = lig + LP * j;
        const double mr = (double)mrow[j];
        const bool rin = r < T;
#pragma unroll
        for (int kb = 0; kb < LP * j + LP; kb += TRB) {  // (columns up to the last row of this register slot)
          if (kb < Tw) {
            // one batch of record entries in flight per trip (no branch between the loads: the L2 round trips overlap)
            double kd[TRB];
#pragma unroll
            for (int e = 0; e < TRB; ++e) kd[e] = (kb + e <= r && rin) ? __ldg(kinv + (size_t)(kb + e) * ldk + r) : 0.0;
#pragma unroll
            for (int e = 0; e < TRB; e += 4) {
              tr0 = fma(kd[e], fma(mr, m64[kb + e], (double)a[j][kb + e]), tr0);
              tr1 = fma(kd[e + 1], fma(mr, m64[kb + e + 1], (double)a[j][kb + e + 1]), tr1);
              tr2 = fma(kd[e + 2], fma(mr, m64[kb + e + 2], (double)a[j][kb + e + 2]), tr2);
              tr3 = fma(kd[e + 3], fma(mr, m64[kb + e + 3], (double)a[j][kb + e + 3]), tr3);
            }
          }
        }
      }
      ldp_rec = __ldg(kinv + ldk * ldk);
    }
    chol_rows<LP, R>(a, lig, T, Tw, sm.col, sm.dgq, bad);
    // z_s = m + L_q eps_s from the register rows
    for (int s = 0; s < S; ++s) {
      float zz[R], zz1[R];
#pragma unroll
      for (int j = 0; j < R; ++j) { zz[j] = mrow[j]; zz1[j] = 0.0f; }
#pragma unroll
      for (int k4 = 0; k4 < TM; k4 += 4) {
        {  // (eps is zero on padded steps: no guard needed)
          const float4 e4 = *reinterpret_cast<const float4*>(sm.vS + s * TM + k4);
#pragma unroll
          for (int j = 0; j < R; ++j) {
            // entries above the diagonal of a register row hold update garbage (never part of L): mask them
            const int r = lig + LP * j;
            fma2<Geo<LP, R>::PACK>(zz[j], zz1[j], k4 <= r ? a[j][k4] : 0.0f, k4 + 1 <= r ? a[j][k4 + 1] : 0.0f, e4.x, e4.y);
            fma2<Geo<LP, R>::PACK>(zz[j], zz1[j], k4 + 2 <= r ? a[j][k4 + 2] : 0.0f, k4 + 3 <= r ? a[j][k4 + 3] : 0.0f, e4.z, e4.w);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int r = lig + LP * j;
        if (r < T) P.z[((size_t)S * pi.r0 + (size_t)s * T + r) * d.D + pi.d] = zz[j] + zz1[j];
      }
    }
    if (!shared) {
      store_rows<LP, R>(a, lig, sm.bufB);
      __syncwarp();
      float (&x)[R][TM] = a;  // reuse the registers: columns of L_q
      load_cols<LP, R>(x, lig, sm.bufB);
      float ssq = 0.0f;
      solve_cols<LP, R>(x, sm.bufA, sm.dinv, Tw);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int jj = 0; jj < R; ++jj) {
          const float v = (i == lig + LP * jj) ? 0.0f : x[jj][i];
          ssq = fmaf(v, v, ssq);
        }
      part = (double)ssq;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int r = lig + LP * j;
        if (r < T) {
          const double lpd = (double)sm.dgp[r], lqd = (double)sm.dgq[r];
          const double av = (double)sm.as[r];
          part += diag_term(lqd / lpd) + av * av;
          if (P.logdets) {  // (float64 logarithms: only when the caller asked for the log-determinants)
            ldp += 2.0 * log(lpd);
            ldq += 2.0 * log(lqd);
          }
        }
      }
    } else {
      // log|K_q| = 2 log prod diag L_q: the product in float64 (<= 64 factors in (0.03, 2): no underflow), ONE logarithm per pair
      __syncwarp();
      double pq = 1.0;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int r = lig + LP * j;
        if (r < T) pq *= (double)sm.dgq[r];
      }
#pragma unroll
      for (int o = LP / 2; o > 0; o >>= 1) pq *= __shfl_xor_sync(0xffffffffu, pq, o);
      part = 2.0 * ((tr0 + tr1) + (tr2 + tr3));
      if (lig == 0) {
        const double lq2 = 2.0 * log(pq);
        part += ldp_rec - lq2 - (double)T;
        ldp = ldp_rec;
        ldq = lq2;
      }
    }
  } else {  // diagonal posterior: X = L_p^-1 columns, h_c = |X[:,c]|^2
    float (&x)[R][TM] = a;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int jj = 0; jj < R; ++jj) x[jj][i] = (i == lig + LP * jj) ? 1.0f : 0.0f;
    solve_cols<LP, R>(x, sm.bufA, sm.dinv, Tw);
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int c = lig + LP * jj;
      float h = 0.0f;
#pragma unroll
      for (int i = 0; i < TM; ++i) h = fmaf(x[jj][i], x[jj][i], h);
      if (c < T) {
        const float lv = P.aux[(size_t)(pi.r0 + c) * d.D + pi.d];
        const float vv = expf(lv), sd = expf(0.5f * lv);
        const double lpd = (double)sm.dgp[c], av = (double)sm.as[c];
        part += (double)h * (double)vv - 1.0 - (double)lv + av * av + 2.0 * log(lpd);
        ldp += 2.0 * log(lpd);
        ldq += (double)lv;
        for (int s = 0; s < S; ++s)
          P.z[((size_t)S * pi.r0 + (size_t)s * T + c) * d.D + pi.d] = mrow[jj] + sd * sm.vS[s * TM + c];
      }
    }
  }
  part = group_sum<LP>(part);
  if (P.logdets) {
    ldp = group_sum<LP>(ldp);
    ldq = group_sum<LP>(ldq);
  }
  bad = __any_sync(0xffffffffu, bad && pi.active) ? 1 : 0;
  if (pi.active && lig == 0) {
    P.kl_pairs[pi.p] = (float)(0.5 * part);
    if (P.logdets) { P.logdets[2 * pi.p] = (float)ldp; P.logdets[2 * pi.p + 1] = (float)ldq; }
  }
  if (bad && P.status && (threadIdx.x & 31) == 0) atomicAdd(P.status, 1);
}

// GLP: also d/d ell_p (per-pair prior path only; extra shared memory behind the per-sample vectors: L_q rows, the dK_p table,
// the V table, 1/diag L_p, alpha -- glp_floats())
template <int LP, int R, int KERNEL, int POST, bool GLP = false>
__global__ void __launch_bounds__(WPC * 32) bwd_warp(Params P, int group_floats) {
  constexpr int TM = LP * R;
  extern __shared__ __align__(16) float smem_f[];
  const GpklDesc& d = P.d;
  const int S = d.S;
  const PairInfo pi = pair_info<LP>(P);
  const int gslot = (threadIdx.x >> 5) * (32 / LP) + (threadIdx.x & 31) / LP;
  Smem<LP, R> sm(smem_f + (size_t)gslot * group_floats, S);
  const int T = pi.T, lig = pi.lig;
  const int Tw = warp_max(T);
  if (Tw == 0) {
    if (pi.active && lig == 0 && P.gq_pairs) P.gq_pairs[pi.p] = 0.0f;
    if (GLP && pi.active && lig == 0) P.gp_pairs[pi.p] = 0.0f;
    return;
  }
  // GLP extras
  float* const bufC = sm.wS + S * TM;
  float* const Kd = bufC + Geo<LP, R>::PK;
  float* const Vs = Kd + Geo<LP, R>::PK;
  float* const dinvp = Vs + TM * TM;
  float* const alv = dinvp + TM;
  float t1p = 0.0f, t3 = 0.0f;
  const float lp = pi.active ? P.ell_p[pi.d] : 1.0f;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const float g = pi.active ? (float)((P.g_kl_sum ? *P.g_kl_sum : 1.0) + (P.g_kl_pairs ? (double)P.g_kl_pairs[pi.p] : 0.0)) : 0.0f;
  float trow[R], mrow[R], gzs[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int r = lig + LP * j;
    const bool ok = r < T;
    trow[j] = ok ? P.times[(size_t)pi.b * d.T_max + r] : 0.0f;
    mrow[j] = ok ? P.mean[(size_t)(pi.r0 + r) * d.D + pi.d] : 0.0f;
    sm.ts[r] = trow[j];
    float gs = 0.0f;
    for (int s = 0; s < S; ++s) {
      sm.vS[s * TM + r] = ok ? eps_value(P, ((size_t)pi.p * S + s) * d.T_max + r) : 0.0f;
      const float gz = (ok && P.g_z) ? P.g_z[((size_t)S * pi.r0 + (size_t)s * T + r) * d.D + pi.d] : 0.0f;
      sm.uS[s * TM + r] = gz;
      gs += gz;
    }
    gzs[j] = gs;
  }
  __syncwarp();
  int bad = 0;
  float a[R][TM];
  float (&x)[R][TM] = a;
  float hdiag[R];
  // shared-prior fast path: the pre-pass found one ell_p for all latent dims and left this sequence's K_p^-1
  const bool shared = (POST == GPKL_POST_GP) && P.prior != nullptr && *P.prior_flag != 0;
  if (!shared) {
    build_rows<LP, R, KERNEL>(a, trow, sm.ts, lig, T, Tw, lp, sig, noise);
    chol_rows<LP, R>(a, lig, T, Tw, sm.col, sm.dgp, bad);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j) {
      sm.dinv[lig + LP * j] = 1.0f / sm.dgp[lig + LP * j];
      if (GLP) dinvp[lig + LP * j] = sm.dinv[lig + LP * j];
    }
    __syncwarp();
    float bvec[R];
#pragma unroll
    for (int j = 0; j < R; ++j) bvec[j] = mrow[j];
    sweep_vec<LP, R>(a, bvec, lig, Tw, sm.dinv, sm.as);
    store_rows<LP, R>(a, lig, sm.bufA);
    __syncwarp();
    // X_p = L_p^-1 as register columns
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int jj = 0; jj < R; ++jj) x[jj][i] = (i == lig + LP * jj) ? 1.0f : 0.0f;
    solve_cols<LP, R>(x, sm.bufA, sm.dinv, Tw);
    // alpha_c = <X_p[:,c], a> ; g_mean = g alpha + sum_s g_z
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      float al = 0.0f, h = 0.0f, al1 = 0.0f, h1 = 0.0f;
#pragma unroll
      for (int k4 = 0; k4 < TM; k4 += 4) {
        {  // (no guard: padded rows of X_p are identity rows and a is zero there)
          const float4 a4 = *reinterpret_cast<const float4*>(sm.as + k4);
          fma2<Geo<LP, R>::PACK>(al, al1, x[jj][k4], x[jj][k4 + 1], a4.x, a4.y);
          fma2<Geo<LP, R>::PACK>(al, al1, x[jj][k4 + 2], x[jj][k4 + 3], a4.z, a4.w);
          if (POST == GPKL_POST_DIAG) {
            fma2<Geo<LP, R>::PACK>(h, h1, x[jj][k4], x[jj][k4 + 1], x[jj][k4], x[jj][k4 + 1]);
            fma2<Geo<LP, R>::PACK>(h, h1, x[jj][k4 + 2], x[jj][k4 + 3], x[jj][k4 + 2], x[jj][k4 + 3]);
          }
        }
      }
      hdiag[jj] = h + h1;
      al += al1;
      const int c = lig + LP * jj;
      if (c < T) P.g_mean[(size_t)(pi.r0 + c) * d.D + pi.d] = g * al + gzs[jj];
      if (GLP) alv[c] = (c < T) ? al : 0.0f;
    }
  }
  if (POST == GPKL_POST_DIAG) {
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int c = lig + LP * jj;
      if (c < T) {
        const float lv = P.aux[(size_t)(pi.r0 + c) * d.D + pi.d];
        const float vv = expf(lv), sd = expf(0.5f * lv);
        float ge = 0.0f;
        for (int s = 0; s < S; ++s) ge = fmaf(sm.uS[s * TM + c], sm.vS[s * TM + c], ge);
        P.g_aux[(size_t)(pi.r0 + c) * d.D + pi.d] = 0.5f * g * (hdiag[jj] * vv - 1.0f) + 0.5f * sd * ge;
      }
    }
  } else {
    const float lq = pi.active ? P.ell_q[pi.d] : 1.0f;
    // t1 = sum_{k != l} dK_q(k,l) (X_p^T X_p)_kl
    float t1 = 0.0f;
    if (!shared) {
      store_cols_rev<LP, R>(x, lig, sm.bufB);
      __syncwarp();
      t1 = contract_cols<LP, R, KERNEL>(x, sm.bufB, sm.ts, trow, lig, T, lq, sig);
      if (GLP) t1p = contract_cols<LP, R, KERNEL>(x, sm.bufB, sm.ts, trow, lig, T, lp, sig);
      __syncwarp();
    }
    // factor K_q
    build_rows<LP, R, KERNEL>(a, trow, sm.ts, lig, T, Tw, lq, sig, noise);
    chol_rows<LP, R>(a, lig, T, Tw, sm.col, sm.dgq, bad);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < R; ++j) sm.dinv[lig + LP * j] = 1.0f / sm.dgq[lig + LP * j];
    if (!GLP) {
      store_rows<LP, R>(a, lig, sm.bufA);
    } else {
      // L_q rows parked in bufC (bufA still holds L_p, bufB the columns of X_p): A = L_p^-1 L_q as register columns,
      // V = X_p^T A into the table, then the quadratic forms against dK_p/d ell_p
      store_rows<LP, R>(a, lig, bufC);
      dk_table<LP, R, KERNEL>(Kd, sm.ts, lig, T, lp, sig);
      __syncwarp();
      load_cols<LP, R>(x, lig, bufC);
      solve_cols<LP, R>(x, sm.bufA, dinvp, Tw);
      dots_cols<LP, R>(x, sm.bufB, lig, T, Vs);
      __syncwarp();
      t3 = quad_cols<LP, R>(Vs, Kd, lig) + quad_vec<LP, R>(alv, Kd, lig);
      __syncwarp();
      for (int e = lig * 4; e < Geo<LP, R>::PK; e += LP * 4)  // L_q rows -> bufA, where the rest of the kernel expects them
        *reinterpret_cast<float4*>(sm.bufA + e) = *reinterpret_cast<const float4*>(bufC + e);
    }
    __syncwarp();
    // w_s = L_q^T g_z,s (column reads of the packed rows) ; pd = 1/2 sum_s w_s eps_s - g/2
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int k = lig + LP * jj;
      float pdk = 0.0f;
      for (int s = 0; s < S; ++s) {
        float wk = 0.0f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {  // (padded rows of L_q are identity rows and g_z is zero there)
          const float l = (i >= k) ? sm.bufA[poff(i) + k] : 0.0f;
          wk = fmaf(l, sm.uS[s * TM + i], wk);
        }
        sm.wS[s * TM + k] = wk;
        pdk = fmaf(wk, sm.vS[s * TM + k], pdk);
      }
      sm.pd[k] = 0.5f * pdk - 0.5f * g;
    }
    __syncwarp();
    // X_q columns
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int jj = 0; jj < R; ++jj) x[jj][i] = (i == lig + LP * jj) ? 1.0f : 0.0f;
    solve_cols<LP, R>(x, sm.bufA, sm.dinv, Tw);
    // C'[:,l] = pd .* X_q[:,l] + sum_s w_s .* prefix(eps_s .* X_q[:,l]) -> reversed-packed rows of bufB
#pragma unroll
    for (int jj = 0; jj < R; ++jj) {
      const int ip = TM - 1 - (lig + LP * jj);
      float* row = sm.bufB + poff_dyn(ip);
      float cum = 0.0f;  // sample 0 fused; further samples are added below
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        {  // (no guard: on padded steps eps = w = 0 and X_q is the identity, every term stays finite and is
           //  multiplied by a zero of X_q in the contraction)
          const float xi = x[jj][i];
          float cv = fmaf(sm.pd[i], xi, sm.wS[i] * cum);
          cum = fmaf(sm.vS[i], xi, cum);
          if (TM - 1 - i <= (ip | 3)) row[(TM - 1 - i) ^ 1] = cv;
        }
      }
      for (int s = 1; s < S; ++s) {
        float cs = 0.0f;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          const float xi = x[jj][i];
          if (TM - 1 - i <= (ip | 3)) row[(TM - 1 - i) ^ 1] = fmaf(sm.wS[s * TM + i], cs, row[(TM - 1 - i) ^ 1]);
          cs = fmaf(sm.vS[s * TM + i], xi, cs);
        }
      }
    }
    __syncwarp();
    float t2;
    if (!shared) {
      t2 = contract_cols<LP, R, KERNEL>(x, sm.bufB, sm.ts, trow, lig, T, lq, sig);
    } else {
      // The record (this sequence's K_p^-1) is first needed here, after the whole K_q chain, so the pre-pass overlaps
      // it (griddep_wait).  alpha = K_p^-1 m and t1 = <K_p^-1, dK_q/d ell> ride along in the same contraction loop
      // (same dK weights), so t2 below already contains g/2 * t1.
      griddep_wait();
      using Rec = PriorRec<LP, R>;
      const float* __restrict__ rec = P.prior + (size_t)pi.b * P.prior_stride;
#pragma unroll
      for (int j = 0; j < R; ++j) sm.as[lig + LP * j] = mrow[j];
      __syncwarp();
      float alpha[R];
      t2 = contract_cols<LP, R, KERNEL, true>(x, sm.bufB, sm.ts, trow, lig, T, lq, sig, rec + Rec::KI, 0.5f * g, sm.as, alpha);
#pragma unroll
      for (int jj = 0; jj < R; ++jj) {
        const int c = lig + LP * jj;
        if (c < T) P.g_mean[(size_t)(pi.r0 + c) * d.D + pi.d] = g * alpha[jj] + gzs[jj];
      }
    }
    const double gq = group_sum<LP>(0.5 * (double)g * (double)t1 + (double)t2);
    if (pi.active && lig == 0) P.gq_pairs[pi.p] = (float)gq;
    if (GLP) {
      const double gp = group_sum<LP>(0.5 * (double)g * ((double)t1p - (double)t3));
      if (pi.active && lig == 0) P.gp_pairs[pi.p] = (float)gp;
    }
  }
  bad = __any_sync(0xffffffffu, bad && pi.active) ? 1 : 0;
  if (bad && P.status && (threadIdx.x & 31) == 0) atomicAdd(P.status, 1);
}

template <int LP, int R>
constexpr int glp_floats() { return 2 * Geo<LP, R>::PK + LP * R * LP * R + 2 * LP * R; }

template <int LP, int R>
size_t warp_smem_bytes(int S, bool glp = false, int wpc = WPC) {
  return (size_t)wpc * (32 / LP) * (Geo<LP, R>::fixed_floats() + 3 * S * LP * R + (glp ? glp_floats<LP, R>() : 0)) * sizeof(float);
}

template <int LP, int R, int KERNEL, int POST, bool BWD>
cudaError_t launch_cfg(const Params& P, cudaStream_t st) {
  const int S = P.d.S;
  const bool glp = BWD && POST == GPKL_POST_GP && (P.d.flags & GPKL_FLAG_GRAD_ELL_P) != 0;
  const int group_floats = Geo<LP, R>::fixed_floats() + 3 * S * LP * R + (glp ? glp_floats<LP, R>() : 0);
  int wpc = WPC;  // the d/d ell_p variant's tables may not leave room for WPC pairs' worth of shared memory per CTA
  while (glp && wpc > 1 && warp_smem_bytes<LP, R>(S, glp, wpc) > kMaxDynSmem) wpc >>= 1;
  const size_t smem = warp_smem_bytes<LP, R>(S, glp, wpc);
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  const int npairs = P.d.B * P.d.D;
  const int per_cta = wpc * (32 / LP);
  const int grid = (npairs + per_cta - 1) / per_cta;
  cudaError_t e;
  bool pdl = false;
  // (the profiling events bracket pre-pass + per-pair kernel: an event recorded between the two would break the
  //  programmatic dependency and serialise them)
  prof_begin(BWD, st);
  if (POST == GPKL_POST_GP && P.prior != nullptr) {  // shared-prior pre-pass: one lane group per sequence
    // shared-prior pre-pass, one CTA per SEQUENCE (gpkl_prior64.cu): float64 K_p^-1 and log|K_p| for the forward's trace,
    // K_p^-1 rounded to float32 (full rows) for the backward's contraction
    e = BWD ? launch_prior_inv64_small(P, st, PriorRec<LP, R>::KI, LP * R) : launch_prior_inv64_small(P, st, 0, 0);
    if (e != cudaSuccess) return e;
    pdl = pdl_enabled();
  }
  // With a pre-pass in front, the per-pair kernel is its programmatic dependent: it starts while the pre-pass
  // runs and waits (griddepcontrol.wait) only where it first touches the records.
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(wpc * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if constexpr (!BWD) {
    auto kern = fwd_warp<LP, R, KERNEL, POST>;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaLaunchKernelEx(&cfg, kern, P, group_floats);
    prof_end(false, st);
  } else {
    void (*kern)(Params, int) = bwd_warp<LP, R, KERNEL, POST>;
    if constexpr (POST == GPKL_POST_GP) {
      if (glp) kern = bwd_warp<LP, R, KERNEL, POST, true>;
    }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaLaunchKernelEx(&cfg, kern, P, group_floats);
    prof_end(true, st);
  }
  if (e != cudaSuccess) return e;
  note_launch();
  return cudaGetLastError();
}

template <int LP, int R, bool BWD>
cudaError_t launch_kp(const Params& P, cudaStream_t st) {
  const bool rbf = P.d.kernel == GPKL_KERNEL_RBF;
  const bool gp = P.d.posterior == GPKL_POST_GP;
  if (rbf && gp) return launch_cfg<LP, R, GPKL_KERNEL_RBF, GPKL_POST_GP, BWD>(P, st);
  if (rbf && !gp) return launch_cfg<LP, R, GPKL_KERNEL_RBF, GPKL_POST_DIAG, BWD>(P, st);
  if (!rbf && gp) return launch_cfg<LP, R, GPKL_KERNEL_CAUCHY, GPKL_POST_GP, BWD>(P, st);
  return launch_cfg<LP, R, GPKL_KERNEL_CAUCHY, GPKL_POST_DIAG, BWD>(P, st);
}

}  // namespace
}  // namespace gpkl
