// V3 hot tier: the bidiagonal-precision posterior (north_star (c); SURVEY.md Appendix A.3 -- an extension, NOT in the
// reference: parity is pinned by the in-repo float64 oracle only) for T <= 64 under the shared prior, one WARP per
// (sequence, latent-dim) pair, O(T^2) per pair.
//
//   q = N(m, Sigma),  Sigma = (B^T B)^-1,  B upper bidiagonal (diagonal b_i > 0, super-diagonal c_i),  W = B^-1
//   z = m + W eps                                   -- ONE back substitution per sample (bidiagonal solve), O(T)
//   KL = 1/2 [ tr(K_p^-1 Sigma) - T + log|K_p| + 2 sum log b_i + m^T K_p^-1 m ]
//
// The generic tier forms W densely and evaluates ||L_p^-1 W||_F^2 (O(T^3) per pair, per-pair prior factorisation).  Here
//  * K_p^-1 comes from the per-SEQUENCE float64 sweep pre-pass of the register tier (gpkl_prior64.cu), shared by the D pairs;
//  * Sigma is never formed: it is SEMISEPARABLE.  With g_i = -c_i / b_i,
//        Sigma_ii = 1/b_i^2 + g_i^2 Sigma_{i+1,i+1},     Sigma_ik = (g_k g_{k+1} ... g_{i-1}) Sigma_ii   (k < i),
//    so lane r walks its row of the record from the diagonal to the left with one running product (float64, three FP64
//    operations per entry) and  tr(K_p^-1 (Sigma + m m^T))  falls out in one pass over the lower triangle;
//  * backward:  d KL / d B = -(B M),  M = Sigma K_p^-1 Sigma, needed on the two diagonals only.  (B M)^T = Sigma Y with
//    Y = K_p^-1 W, and the columns of W obey  W[:,k+1] = h_k W[:,k] + e_{k+1} / b_{k+1}  (h_k = -c_k / b_{k+1}), so lane r
//    builds row r of Y by a recurrence along k (parked k-major in shared memory), and lane i contracts column i of Y with rows
//    i and i+1 of Sigma through the same running products:  O(T^2) per pair, no T x T product anywhere.
//    The sample path adds  -lambda_s w_s^T  on the two diagonals,  lambda = B^-T g_z (forward substitution), w = B^-1 eps.
// If ell_p differs between latent dims (device flag) these kernels return at once and the generic tier, launched behind
// them with skip_if_shared, does the work.
#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int BW = 4;        // warps (pairs) per CTA
constexpr int TMX = 64;      // longest sequence of this tier

// Per-warp shared memory, sized by the batch (tm = T_max rounded up to 8, S samples): doubles g[tm+1] (g_i = -c_i / b_i, 0 beyond
// the sequence), sd[tm+1] (Sigma_ii, sd[T] = 0), m[tm]; floats b, c, h (h_i = -c_i / b_{i+1}), rb (1 / b_i), w[S][tm]
// (w_s = B^-1 eps_s), lam[S][tm] (lambda_s = B^-T g_z,s; backward), Y[tm][tm+1] (backward; odd pitch: lane i reads column i
// conflict-free).
struct BdSm {
  double *g, *sd, *m;
  float *b, *c, *h, *rb, *w, *lam, *Y;
  int tm;
  __host__ __device__ static size_t bytes(int tm, int S, bool backward) {
    return (size_t)(3 * tm + 2) * sizeof(double) + ((size_t)4 * tm + (size_t)(backward ? 2 : 1) * S * tm + (backward ? (size_t)tm * (tm + 1) : 0) + 2) * sizeof(float);
  }
  __device__ BdSm(unsigned char* base, int tm_, int S, bool backward) : tm(tm_) {
    g = reinterpret_cast<double*>(base);
    sd = g + tm + 1;
    m = sd + tm + 1;
    b = reinterpret_cast<float*>(m + tm);
    c = b + tm;
    h = c + tm;
    rb = h + tm;
    w = rb + tm;
    lam = w + (size_t)S * tm;
    Y = backward ? lam + (size_t)S * tm : nullptr;
  }
};
__host__ __device__ inline size_t bd_warp_bytes(int tm, int S, bool backward) { return (BdSm::bytes(tm, S, backward) + 15) / 16 * 16; }

struct BdPair {
  int p, b, d, T;
  long long r0;
  bool active;
};

__device__ __forceinline__ BdPair bd_pair(const Params& P) {
  BdPair q;
  q.p = blockIdx.x * BW + (threadIdx.x >> 5);
  q.active = q.p < P.d.B * P.d.D;
  q.b = q.active ? q.p / P.d.D : 0;
  q.d = q.active ? q.p - q.b * P.d.D : 0;
  q.T = q.active ? P.lengths[q.b] : 0;
  q.r0 = q.active ? P.offsets[q.b] : 0;
  return q;
}

// b, c, m of the pair into shared memory; g, h, 1/b; Sigma_ii by the backward recurrence (lane 0); w_s by back substitution
// (lanes 1..S), eps from the caller's tensor or the in-kernel generator.
__device__ __forceinline__ void bd_load(const Params& P, const BdPair& q, const BdSm& s, int lane) {
  const GpklDesc& d = P.d;
  const int T = q.T;
  const int tm = s.tm;
  for (int i = lane; i < tm; i += 32) {
    const bool ok = i < T;
    const float bi = ok ? P.aux[((size_t)(q.r0 + i) * d.D + q.d) * 2] : 1.0f;
    const float ci = (i + 1 < T) ? P.aux[((size_t)(q.r0 + i) * d.D + q.d) * 2 + 1] : 0.0f;
    s.b[i] = bi;
    s.c[i] = ci;
    s.rb[i] = 1.0f / bi;
    s.g[i] = -(double)ci / (double)bi;
    s.m[i] = ok ? (double)P.mean[(size_t)(q.r0 + i) * d.D + q.d] : 0.0;
  }
  if (lane == 0) s.g[tm] = 0.0;
  __syncwarp();
  for (int i = lane; i < tm; i += 32) s.h[i] = (i + 1 < T) ? -s.c[i] / s.b[i + 1] : 0.0f;
  if (lane == 0) {
    double acc = 0.0;
    s.sd[T] = 0.0;
    for (int i = T - 1; i >= 0; --i) {
      const double rb = 1.0 / (double)s.b[i], gi = s.g[i];
      acc = fma(gi * gi, acc, rb * rb);
      s.sd[i] = acc;
    }
  } else if (lane <= d.S) {
    const int sx = lane - 1;
    float wv = 0.0f;
    for (int i = T - 1; i >= 0; --i) {  // w_i = (eps_i - c_i w_{i+1}) / b_i
      const float e = eps_value(P, ((size_t)q.p * d.S + sx) * d.T_max + i);
      wv = __fmul_rn(__fmaf_rn(-s.c[i], wv, e), s.rb[i]);  // (explicit: no contraction with the inlined noise generator)
      s.w[sx * tm + i] = wv;
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(BW * 32) fwd_bidiag(Params P) {
  extern __shared__ __align__(16) unsigned char bd_raw[];
  if (*P.prior_flag == 0) return;  // ell_p differs between latent dims: the generic tier launched behind does the work
  const GpklDesc& d = P.d;
  const int lane = threadIdx.x & 31;
  const int tm = prior64_pitch(d.T_max);
  const BdSm s(bd_raw + (size_t)(threadIdx.x >> 5) * bd_warp_bytes(tm, d.S, false), tm, d.S, false);
  const BdPair q = bd_pair(P);
  const int T = q.T;
  if (!q.active) return;
  if (T <= 0) {
    if (lane == 0) {
      P.kl_pairs[q.p] = 0.0f;
      if (P.logdets) { P.logdets[2 * q.p] = 0.0f; P.logdets[2 * q.p + 1] = 0.0f; }
    }
    return;
  }
  bd_load(P, q, s, lane);
  for (int sx = 0; sx < d.S; ++sx)
    for (int i = lane; i < T; i += 32)
      P.z[((size_t)d.S * q.r0 + (size_t)sx * T + i) * d.D + q.d] = (float)s.m[i] + s.w[sx * tm + i];
  // tr(K_p^-1 (Sigma + m m^T)) = 2 sum_{k <= r} Kinv'_rk (Sigma_rk + m_r m_k)   (record: diagonal halved, gpkl_prior64.cu)
  const int ldk = prior64_pitch(d.T_max);
  const double* __restrict__ kinv = reinterpret_cast<const double*>(P.prior + (size_t)q.b * P.prior_stride);
  double tr = 0.0, slb = 0.0;
  {
    // rows r0 = lane and r1 = lane + 32; the columns k run from T-1 down to 0 for the whole warp (coalesced record reads);
    // a row joins at its diagonal (k == r), where its running Sigma_rk starts as Sigma_rr
    const int ra = lane, rb2 = lane + 32;
    const bool ina = ra < T, inb = rb2 < T;
    const double ma = ina ? s.m[ra] : 0.0, mb = inb ? s.m[rb2] : 0.0;
    const double sa = ina ? s.sd[ra] : 0.0, sb = inb ? s.sd[rb2] : 0.0;
    double pa = 0.0, pb = 0.0, acc0 = 0.0, acc1 = 0.0;
    for (int kb = T - 1; kb >= 0; kb -= 8) {
      double ka[8], kq[8];  // eight columns in flight per row
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb - e;
        ka[e] = (k >= 0 && k <= ra && ina) ? __ldg(kinv + (size_t)k * ldk + ra) : 0.0;
        kq[e] = (k >= 0 && k <= rb2 && inb) ? __ldg(kinv + (size_t)k * ldk + rb2) : 0.0;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb - e;
        if (k >= 0) {  // (warp-uniform)
          const double mk = s.m[k], gk = k > 0 ? s.g[k - 1] : 0.0;
          pa = (k == ra) ? sa : pa;
          pb = (k == rb2) ? sb : pb;
          acc0 = fma(ka[e], fma(ma, mk, pa), acc0);
          acc1 = fma(kq[e], fma(mb, mk, pb), acc1);
          pa *= gk;
          pb *= gk;
        }
      }
    }
    tr = acc0 + acc1;
    if (ina) slb += log((double)s.b[ra]);
    if (inb) slb += log((double)s.b[rb2]);
  }
  tr = warp_sum(tr);
  slb = warp_sum(slb);
  if (lane == 0) {
    const double ldp = __ldg(kinv + (size_t)ldk * ldk);
    P.kl_pairs[q.p] = (float)(0.5 * (2.0 * tr - (double)T + ldp + 2.0 * slb));
    if (P.logdets) { P.logdets[2 * q.p] = (float)ldp; P.logdets[2 * q.p + 1] = (float)(-2.0 * slb); }
  }
}

// Backward record: K_p^-1 rounded to float32, full symmetric tm x tm (tm = prior64_pitch(T_max); identity on the padding),
// at float offset 0.
__global__ void __launch_bounds__(BW * 32) bwd_bidiag(Params P) {
  extern __shared__ __align__(16) unsigned char bd_raw[];
  if (*P.prior_flag == 0) return;
  const GpklDesc& d = P.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tm = prior64_pitch(d.T_max), YLD = tm + 1;
  const BdSm s(bd_raw + (size_t)warp * bd_warp_bytes(tm, d.S, true), tm, d.S, true);
  float* Y = s.Y;  // Y[k * YLD + r]
  const BdPair q = bd_pair(P);
  const int T = q.T, S = d.S;
  if (!q.active || T <= 0) return;
  bd_load(P, q, s, lane);
  const float g = (float)((P.g_kl_sum ? *P.g_kl_sum : 1.0) + (P.g_kl_pairs ? (double)P.g_kl_pairs[q.p] : 0.0));
  // lambda_s = B^-T g_z,s by forward substitution (lanes 0..S-1): lambda_i = (gz_i - c_{i-1} lambda_{i-1}) / b_i
  if (lane < S) {
    float lv = 0.0f;
    for (int i = 0; i < T; ++i) {
      const float gz = P.g_z ? P.g_z[((size_t)S * q.r0 + (size_t)lane * T + i) * d.D + q.d] : 0.0f;
      lv = (gz - (i > 0 ? s.c[i - 1] : 0.0f) * lv) * s.rb[i];
      s.lam[lane * tm + i] = lv;
    }
  }
  const float* __restrict__ kinv = P.prior + (size_t)q.b * P.prior_stride;  // kinv[k * tm + r], symmetric
  // row r of Y = K_p^-1 W along k, and alpha_r = (K_p^-1 m)_r in the same pass
  for (int r = lane; r < T; r += 32) {
    float y = 0.0f, al = 0.0f;
    for (int kb = 0; kb < T; kb += 8) {  // eight record entries in flight (tm is a multiple of 8: the padding reads are in bounds)
      float kv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) kv[e] = __ldg(kinv + (size_t)(kb + e) * tm + r);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = kb + e;
        if (k < T) {
          y = fmaf(k > 0 ? s.h[k - 1] : 0.0f, y, kv[e] * s.rb[k]);
          al = fmaf(kv[e], (float)s.m[k], al);
          Y[k * YLD + r] = y;
        }
      }
    }
    float gzs = 0.0f;
    if (P.g_z)
      for (int sx = 0; sx < S; ++sx) gzs += P.g_z[((size_t)S * q.r0 + (size_t)sx * T + r) * d.D + q.d];
    P.g_mean[(size_t)(q.r0 + r) * d.D + q.d] = fmaf(g, al, gzs);
  }
  __syncwarp();
  // lane i: A = sum_{r <= i} (g_r ... g_{i-1}) Y_ri,  Bq = sum_{r > i} (g_{i+1} ... g_{r-1}) Sigma_rr Y_ri
  //   (B M)_ii = Sigma_ii A + g_i Bq,   (B M)_{i,i+1} = g_i Sigma_{i+1,i+1} A + Bq
  for (int i = lane; i < T; i += 32) {
    const float* __restrict__ yc = Y + (size_t)i * YLD;
    float A = 0.0f, pr = 1.0f;
    for (int r = i; r >= 0; --r) {
      A = fmaf(pr, yc[r], A);
      pr *= r > 0 ? (float)s.g[r - 1] : 0.0f;
    }
    float Bq = 0.0f, qr = 1.0f;
    for (int r = i + 1; r < T; ++r) {
      Bq = fmaf(qr * (float)s.sd[r], yc[r], Bq);
      qr *= (float)s.g[r];
    }
    const float gi = (float)s.g[i];
    const float d1 = fmaf((float)s.sd[i], A, gi * Bq);
    const float d2 = fmaf(gi * (float)s.sd[i + 1], A, Bq);
    float gb = g * (s.rb[i] - d1), gc = -g * d2;
    for (int sx = 0; sx < S; ++sx) {  // sample path: -lambda_i w_i, -lambda_i w_{i+1}
      gb = fmaf(-s.lam[sx * tm + i], s.w[sx * tm + i], gb);
      if (i + 1 < T) gc = fmaf(-s.lam[sx * tm + i], s.w[sx * tm + i + 1], gc);
    }
    float* ga = P.g_aux + ((size_t)(q.r0 + i) * d.D + q.d) * 2;
    ga[0] = gb;
    ga[1] = (i + 1 < T) ? gc : 0.0f;
  }
}

}  // namespace

bool bidiag_tier_supports(const GpklDesc& d) {
  return d.posterior == GPKL_POST_BIDIAG && d.T_max >= 1 && d.T_max <= TMX && d.S <= 8 && !(d.flags & GPKL_FLAG_GRAD_ELL_P);
}

cudaError_t launch_bidiag(const Params& P, bool backward, cudaStream_t st) {
  if (!P.prior || !P.prior_flag) return cudaErrorInvalidValue;
  cudaError_t e = backward ? launch_prior_inv64_small(P, st, 0, prior64_pitch(P.d.T_max)) : launch_prior_inv64_small(P, st, 0, 0);
  if (e != cudaSuccess) return e;
  const int npairs = P.d.B * P.d.D;
  const int grid = (npairs + BW - 1) / BW;
  const size_t smem = BW * bd_warp_bytes(prior64_pitch(P.d.T_max), P.d.S, backward);
  void (*kern)(Params) = backward ? bwd_bidiag : fwd_bidiag;  // (the kernel kind only enters through the pre-pass)
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, BW * 32, smem, st>>>(P);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gpkl
