// Per-SEQUENCE prior record of the tile tier's forward pass: K_p^-1 and log|K_p| in float64.
//
// The reference evaluates  tr(K_p^-1 K_q)  in float64 with an LU inverse of the float32-built K_p
// (src/Models/Full_GP_VAE_dynamic_time.py:250-254), once per (sequence, latent-dim) pair although K_p is the same for the D
// pairs of a sequence (prior_time_chars is one constant, :114).  Round 1 evaluated the trace as ||L_p^-1 L_q||_F^2 -- a
// T^3/6-FMA triangular product PER PAIR, half of the forward pass at T = 512 -- because a float32 K_p^-1 loses
// cond(K) * eps in the entrywise form.  With K_p^-1 in float64 the entrywise form is exact to ~1e-13 and costs O(T^2) per
// pair:   tr(K_p^-1 (K_q + m m^T)) = sum_ij Kinv_ij (K_q,ij + m_i m_j),  fused into the pass that generates K_q (gpkl_tile.cu).
// The float64 inverse costs T^3/2 DFMA per SEQUENCE (C4: 1,024 sequences against 65,536 pairs).
//
// Algorithm: blocked symmetric SWEEP operator, in place on the lower triangle (column-major, A[c*TP + i], i >= c), NB = 16
// pivots at a time.  For the pivot block P = A_kk:  H = P^-1 (scalar sweeps in shared memory, the pivots give log|K_p|),
// W = the block row/column (all other indices), V = W H, then ONE uniform rank-NB update of the whole lower triangle
// A_ij -= V_i W_j^T (rows of the pivot block carry W = 0), A_ik <- V_i, A_kk <- -H.  After all blocks A = -K_p^-1 (negated at
// the end).  Every pivot block is a Schur complement of an SPD matrix, so no pivoting is needed.  The update runs in 4 x 4
// register tiles with both operands contraction-major in shared memory ([c][i]: a warp reads consecutive rows -- no bank
// conflicts -- and the other operand is a broadcast), the next tile of A is prefetched while the current one is computed.
// SMALL BATCHES: a thread-block CLUSTER of CS CTAs (8, 4 or 2 when there are fewer sequences than SMs / CS) works on one
// sequence: every CTA factors the pivot block and forms W, V redundantly in its own shared memory (small), the rank-NB update
// of the triangle -- the O(T^3) part -- is dealt tile by tile over the cluster, and two cluster barriers per panel order the
// update against the write-back (16 sequences: 3.4 -> 1.5 ms).
// No tensor cores (north_star); DFMA on the FP64 pipe.
#include <string.h>

#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

#define P64_TICK(k) if (P.dbg && blockIdx.x == 0 && tid == 0) { const long long now_ = clock64(); P.dbg[k] += now_ - tk_; tk_ = now_; }

constexpr int NB = 16;
constexpr int NT = 256;

__host__ __device__ inline int p64_tp(int T_max) {
  int TP = (T_max + 63) / 64 * 64;
  return TP < 64 ? 64 : TP;
}
__host__ __device__ inline int p64_ldw(int TP) { return TP + 2; }  // operand row pitch (doubles): 16-byte aligned, rows 4 banks apart

__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster; global-memory writes before it are visible to every CTA after it -- in L2: every read
// of the matrix in this kernel is an L1-bypassing ld.global.cg (another SM may have rewritten part of a cached line)
__device__ __forceinline__ void cluster_sync_all() {
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int KERNEL>
__global__ void __launch_bounds__(NT, 1) prior_inv64_kernel(Params P) {
  extern __shared__ __align__(16) double smd[];
  griddep_launch_dependents();  // a per-pair kernel launched as programmatic dependent may start; it reads the records after its griddep_wait()
  if (*P.prior_flag == 0) return;  // ell_p differs between latent dims: the per-pair kernels take their per-pair path
  const GpklDesc& d = P.d;
  const int TP = p64_tp(d.T_max), ldw = p64_ldw(TP);
  double* Wt = smd;
  double* Vt = Wt + (size_t)NB * ldw;
  double* Pm = Vt + (size_t)NB * ldw;  // [NB][NB + 1]
  double* piv = Pm + NB * (NB + 1);    // [TP]
  float* ts = reinterpret_cast<float*>(piv + TP);
  const int tid = threadIdx.x;
  const int CS = (int)cluster_size(), cr = (int)cluster_rank();
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const KernC<KERNEL> kc(P.ell_p[0], sig);
  // (every CTA of a cluster takes the same trips through this loop and the panel loop: the cluster barriers line up)
  for (int b = blockIdx.x / CS; b < d.B; b += gridDim.x / CS) {
    const int n = P.lengths[b];
    double* A = reinterpret_cast<double*>(P.prior + (size_t)b * P.prior_stride);
    __syncthreads();
    if (n <= 0) {
      if (tid == 0 && cr == 0) A[(size_t)TP * TP] = 0.0;
      continue;
    }
    for (int i = tid; i < n; i += NT) ts[i] = P.times[(size_t)b * d.T_max + i];
    __syncthreads();
    // K_p, lower triangle: the float32 kernel values (as every tier builds them) cast to float64
    for (int c = cr; c < n; c += CS) {
      const float tc = ts[c];
      for (int i = c + tid; i < n; i += NT) A[(size_t)c * TP + i] = (double)(kc.val(ts[i] - tc) + (i == c ? noise : 0.0f));
    }
    cluster_sync_all();
    const int n4 = (n + 3) >> 2;
    long long tk_ = clock64();
    for (int k0 = 0; k0 < n; k0 += NB) {
      const int nb = n - k0 < NB ? n - k0 : NB;
      __syncthreads();
      // ---- pivot block -> Pm (full symmetric; identity beyond nb) ----------------------------------------------------
      {
        const int r = tid >> 4, c = tid & 15;
        const int hi = r > c ? r : c, lo = r > c ? c : r;
        Pm[r * (NB + 1) + c] = (r < nb && c < nb) ? __ldcg(A + (size_t)(k0 + lo) * TP + k0 + hi) : (r == c ? 1.0 : 0.0);
      }
      // ---- W^T: the block row / column of every other index (zero for the block's own rows and beyond n).  The loads are
      // issued in batches of 8 / 16 before their values are stored: each is an L2 round trip (L1 is bypassed)
      for (int e0 = tid; e0 < k0 * NB; e0 += 8 * NT) {  // rows above the block: A(k0+c, i) stored as A[i*TP + k0 + c]
        double v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int e = e0 + q * NT, i = e >> 4, c = e & 15;
          v[q] = (e < k0 * NB && c < nb) ? __ldcg(A + (size_t)i * TP + k0 + c) : 0.0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int e = e0 + q * NT, i = e >> 4, c = e & 15;
          if (e < k0 * NB) Wt[(size_t)c * ldw + i] = v[q];
        }
      }
      for (int i = k0 + tid; i < 4 * n4; i += NT) {
        double v[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) v[c] = (c < nb && i >= k0 + nb && i < n) ? __ldcg(A + (size_t)(k0 + c) * TP + i) : 0.0;
#pragma unroll
        for (int c = 0; c < NB; ++c) Wt[(size_t)c * ldw + i] = v[c];
      }
      __syncthreads();
      P64_TICK(0)
      // ---- H = P^-1 by scalar sweeps (Pm <- -P^-1); pivots -> log|K_p| --------------------------------------------------
      {
        const int r = tid >> 4, c = tid & 15;
        for (int s = 0; s < nb; ++s) {
          const double dv = Pm[s * (NB + 1) + s];
          const double x = Pm[r * (NB + 1) + c], prs = Pm[r * (NB + 1) + s], psc = Pm[s * (NB + 1) + c];
          __syncthreads();
          const double inv = 1.0 / dv;
          Pm[r * (NB + 1) + c] = (r == s) ? (c == s ? -inv : psc * inv) : (c == s ? prs * inv : x - prs * psc * inv);
          if (tid == 0) piv[k0 + s] = dv;  // (log|K_p| = sum of the logs of the pivots, taken after the last panel)
          __syncthreads();
        }
      }
      P64_TICK(1)
      // ---- V = W H = -W Pm ---------------------------------------------------------------------------------------------
      for (int i = tid; i < 4 * n4; i += NT) {
        double v[NB];
#pragma unroll
        for (int c = 0; c < NB; ++c) v[c] = 0.0;
#pragma unroll 4
        for (int cp = 0; cp < NB; ++cp) {
          const double wv = -Wt[(size_t)cp * ldw + i];
#pragma unroll
          for (int c = 0; c < NB; ++c) v[c] = fma(wv, Pm[cp * (NB + 1) + c], v[c]);
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) Vt[(size_t)c * ldw + i] = v[c];
      }
      __syncthreads();
      P64_TICK(2)
      // ---- A_ij -= V_i W_j^T over the lower triangle, 4 x 4 tiles walked column by column, next tile prefetched ----------
      {
        int J4 = 0, r = cr * NT + tid;
        auto settle = [&]() {
          while (J4 < n4 && r >= n4 - J4) { r -= n4 - J4; ++J4; }
        };
        settle();
        double a[4][4];
        auto load_tile = [&](int I4t, int J4t) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const double* ptr = A + (size_t)(4 * J4t + jj) * TP + 4 * I4t;
            const double2 lo = __ldcg(reinterpret_cast<const double2*>(ptr)), hi = __ldcg(reinterpret_cast<const double2*>(ptr + 2));
            a[jj][0] = lo.x; a[jj][1] = lo.y; a[jj][2] = hi.x; a[jj][3] = hi.y;
          }
        };
        if (J4 < n4) load_tile(J4 + r, J4);
        while (J4 < n4) {
          const int I4 = J4 + r, cJ4 = J4;
          double cur[4][4];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) cur[jj][ii] = a[jj][ii];
          r += NT * CS;
          settle();
          if (J4 < n4) load_tile(J4 + r, J4);
          const double* vp = Vt + 4 * I4;
          const double* wp = Wt + 4 * cJ4;
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            const double2 v0 = *reinterpret_cast<const double2*>(vp + (size_t)c * ldw), v1 = *reinterpret_cast<const double2*>(vp + (size_t)c * ldw + 2);
            const double2 w0 = *reinterpret_cast<const double2*>(wp + (size_t)c * ldw), w1 = *reinterpret_cast<const double2*>(wp + (size_t)c * ldw + 2);
            const double vv[4] = {v0.x, v0.y, v1.x, v1.y}, ww[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)
#pragma unroll
              for (int ii = 0; ii < 4; ++ii) cur[jj][ii] = fma(-ww[jj], vv[ii], cur[jj][ii]);
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            double* ptr = A + (size_t)(4 * cJ4 + jj) * TP + 4 * I4;
            *reinterpret_cast<double2*>(ptr) = make_double2(cur[jj][0], cur[jj][1]);
            *reinterpret_cast<double2*>(ptr + 2) = make_double2(cur[jj][2], cur[jj][3]);
          }
        }
      }
      P64_TICK(3)
      cluster_sync_all();  // every tile of the update (they also rewrite the block's rows / columns unchanged) is done
      P64_TICK(4)
      // ---- write back: A_ik <- V_i (row part and column part), A_kk <- Pm = -H ---------------------------------------------
      if (cr == 0) {
        for (int e = tid; e < k0 * NB; e += NT) {
          const int i = e >> 4, c = e & 15;
          if (c < nb) A[(size_t)i * TP + k0 + c] = Vt[(size_t)c * ldw + i];
        }
        for (int c = 0; c < nb; ++c)
          for (int i = k0 + nb + tid; i < n; i += NT) A[(size_t)(k0 + c) * TP + i] = Vt[(size_t)c * ldw + i];
        const int r = tid >> 4, c = tid & 15;
        if (r < nb && c <= r) A[(size_t)(k0 + c) * TP + k0 + r] = Pm[r * (NB + 1) + c];
      }
      P64_TICK(5)
      cluster_sync_all();  // the write-back is visible to the next panel's loads in every CTA
      P64_TICK(6)
    }
    // A = -K_p^-1 -> K_p^-1
    for (int c = cr; c < n; c += CS)
      for (int i = c + tid; i < n; i += NT) A[(size_t)c * TP + i] = -__ldcg(A + (size_t)c * TP + i);
    if (cr == 0) {  // (every CTA holds all pivots: it swept every block itself)
      double ld = 0.0;
      int bad = 0;
      for (int i = tid; i < n; i += NT) {
        const double dv = piv[i];
        ld += log(dv);
        if (!(dv > 0.0)) bad = 1;
      }
      ld = block_sum(ld, Wt);  // (Wt is idle: scratch of the reduction)
      bad = __syncthreads_or(bad);
      if (tid == 0) {
        A[(size_t)TP * TP] = ld;
        if (bad && P.status) atomicAdd(P.status, 1);
      }
    }
  }
}

// n scalar symmetric sweeps of the register-resident lower triangle (prior_inv64_small_kernel): QN entries per thread.
// Pivot step s: the owners of sym(., s) publish the pivot column, one barrier, every entry takes
//   x(r,c) -= col[r] col[c] / d      (r, c != s),      x(r,s) = col[r] / d,      x(s,s) = -1 / d
// branch-free (the operand of the multiply and the result are selected), so the QN entries of a thread overlap.
template <int QN, int Q, int TP>
__device__ __forceinline__ void small_sweep(double (&x)[Q], const int (&rr)[Q], const int (&cc)[Q], int n,
                                            double (*pc)[TP + 2], double* piv) {
  static_assert(QN <= Q, "entries per thread");
  for (int s = 0; s < n; ++s) {
    double* col = pc[s & 1];
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      const bool rs = rr[q] == s, cs = cc[q] == s;
      if (rs | cs) col[rs ? cc[q] : rr[q]] = x[q];
    }
    __syncthreads();
    const double dv = col[s];
    double inv = (double)__frcp_rn((float)dv);  // pivots of an SPD float32-built matrix: well inside the float range
    const double e1 = fma(-dv, inv, 1.0);       // 1/d = inv (1 + e + e^2 + ...), |e| ~ 1e-7: two terms
    inv = fma(inv, fma(e1, e1, e1), inv);
    const double ninv = -inv;
#pragma unroll
    for (int q = 0; q < QN; ++q) {
      const bool rs = rr[q] == s, cs = cc[q] == s;
      const double psc = col[cc[q]], prs = col[rr[q]];
      const double t = (rs ? psc : prs) * inv;
      const double upd = fma(-t, psc, x[q]);
      x[q] = (rs | cs) ? ((rs & cs) ? ninv : t) : upd;
    }
    if (threadIdx.x == 0) piv[s] = dv;
  }
}

// T_max <= 64 (the register tier's sizes; consumers: fwd_warp / bwd_warp, gpkl_warp.cuh).  The lower triangle lives in
// REGISTERS: the n(n+1)/2 entries are dealt to the 256 threads (<= 9 each, column-major packed so that the record is written
// coalesced), and one scalar symmetric sweep per pivot needs only the pivot column sym(., s), which its owners publish through
// a double-buffered shared vector: ONE barrier per pivot, 2 FP64 operations per entry and pivot, the pivot's reciprocal by
// MUFU.RCP + two Newton steps.
// Forward record (f32_tm == 0): -(sweep result) = K_p^-1 in float64, lower triangle column-major, THE DIAGONAL
// HALVED (the consumer's trace is 2 * sum_{k <= r}), log|K_p| at [pitch * pitch]; pitch = prior64_pitch(T_max) (T_max rounded up to 8).
// Backward record (f32_tm = TM of the consumer's lane geometry): K_p^-1 rounded to float32, full symmetric TM x TM at float
// offset f32_off (PriorRec::KI), identity on the padding.
// Several CTAs per SM; the per-pair kernel launched behind it (programmatic dependent launch) starts at once and waits where
// it first reads a record.
// NTH threads per sequence for sequences of at most NMAX points (64 / 16, 128 / 32, 256 / 64): the pivot chain is the same,
// but barriers, the final reduction and the idle lanes of a short sequence are cheaper with fewer warps.
template <int KERNEL, int NTH, int NMAX>
__global__ void __launch_bounds__(NTH) prior_inv64_small_kernel(Params P, int f32_off, int f32_tm) {
  constexpr int TP = 64, NT = NTH, Q = (NMAX * (NMAX + 1) / 2 + NT - 1) / NT;
  __shared__ double pc[2][TP + 2];
  __shared__ double piv[TP];
  __shared__ double red[32];
  __shared__ float ts[TP];
  griddep_launch_dependents();
  if (*P.prior_flag == 0) return;
  const GpklDesc& d = P.d;
  const int tid = threadIdx.x;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const KernC<KERNEL> kc(P.ell_p[0], sig);
  for (int i = tid; i < 2 * (TP + 2); i += NT) (&pc[0][0])[i] = 0.0;
  for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
    const int n = P.lengths[b];
    float* rec = P.prior + (size_t)b * P.prior_stride;
    double* A = reinterpret_cast<double*>(rec);
    const int ldk = prior64_pitch(d.T_max);  // column pitch of the float64 record
    __syncthreads();
    if (f32_tm) {  // padding of the float32 record: identity
      for (int e = tid; e < f32_tm * f32_tm; e += NT) {
        const int r = e / f32_tm, c = e - r * f32_tm;
        if (r >= n || c >= n) rec[f32_off + e] = r == c ? 1.0f : 0.0f;
      }
    }
    if (n <= 0) {
      if (tid == 0 && !f32_tm) A[(size_t)ldk * ldk] = 0.0;
      continue;
    }
    if (tid < n) ts[tid] = P.times[(size_t)b * d.T_max + tid];
    __syncthreads();
    const int ne = n * (n + 1) / 2;
    const int qn = (ne + NT - 1) / NT;  // entries per thread of this sequence (the last one may be a dummy)
    int rr[Q], cc[Q];
    double x[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const int e = tid + q * NT;
      rr[q] = cc[q] = TP;  // dummy entry: never a pivot row / column, reads the spare slot of the pivot vector
      x[q] = 0.0;
      if (e < ne) {
        // column-major packed lower triangle = the row-major packed one read backwards and mirrored
        const int f = ne - 1 - e;
        int rp = (int)((sqrtf(8.0f * (float)f + 1.0f) - 1.0f) * 0.5f);
        while (rp * (rp + 1) / 2 > f) --rp;
        while ((rp + 1) * (rp + 2) / 2 <= f) ++rp;
        const int cp = f - rp * (rp + 1) / 2;
        rr[q] = n - 1 - cp;
        cc[q] = n - 1 - rp;
        x[q] = (double)(kc.val(ts[rr[q]] - ts[cc[q]]) + (rr[q] == cc[q] ? noise : 0.0f));
      }
    }
    // (compile-time entry counts: no branch between the entries of a pivot step, their latencies overlap)
    if (qn <= 1) small_sweep<1, Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 2) small_sweep<(Q < 2 ? Q : 2), Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 3) small_sweep<(Q < 3 ? Q : 3), Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 4) small_sweep<(Q < 4 ? Q : 4), Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 5) small_sweep<(Q < 5 ? Q : 5), Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 6) small_sweep<(Q < 6 ? Q : 6), Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 7) small_sweep<(Q < 7 ? Q : 7), Q, TP>(x, rr, cc, n, pc, piv);
    else if (qn == 8) small_sweep<(Q < 8 ? Q : 8), Q, TP>(x, rr, cc, n, pc, piv);
    else small_sweep<Q, Q, TP>(x, rr, cc, n, pc, piv);
    if (!f32_tm) {
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (rr[q] < TP) A[(size_t)cc[q] * ldk + rr[q]] = (rr[q] == cc[q] ? -0.5 : -1.0) * x[q];
    } else {
#pragma unroll
      for (int q = 0; q < Q; ++q)
        if (rr[q] < TP) {
          const float v = (float)(-x[q]);
          rec[f32_off + cc[q] * f32_tm + rr[q]] = v;
          rec[f32_off + rr[q] * f32_tm + cc[q]] = v;
        }
    }
    __syncthreads();
    double ld = 0.0;
    int bad = 0;
    if (tid < n) {
      ld = log(piv[tid]);
      if (!(piv[tid] > 0.0)) bad = 1;
    }
    ld = block_sum(ld, red);
    bad = __syncthreads_or(bad);
    if (tid == 0) {
      if (!f32_tm) A[(size_t)ldk * ldk] = ld;
      if (bad && P.status) atomicAdd(P.status, 1);
    }
  }
}

}  // namespace

size_t prior64_record_floats(int T_max) {
  const size_t TP = T_max <= 64 ? (size_t)prior64_pitch(T_max) : (size_t)p64_tp(T_max);
  return 2 * (TP * TP + 2);
}

// register tier's pre-pass; f32_tm == 0: float64 forward record, else the float32 backward record (see the kernel)
cudaError_t launch_prior_inv64_small(const Params& P, cudaStream_t st, int f32_off, int f32_tm) {
  if (P.d.T_max > 64 || f32_tm > 64) return cudaErrorInvalidValue;
  const size_t need = f32_tm ? (size_t)f32_off + (size_t)f32_tm * f32_tm : prior64_record_floats(P.d.T_max);
  if (P.prior_stride < need) return cudaErrorInvalidValue;
  const bool rbf = P.d.kernel == GPKL_KERNEL_RBF;
  void (*ks)(Params, int, int);
  int nth;
  if (P.d.T_max <= 16) {
    nth = 64;
    ks = rbf ? prior_inv64_small_kernel<GPKL_KERNEL_RBF, 64, 16> : prior_inv64_small_kernel<GPKL_KERNEL_CAUCHY, 64, 16>;
  } else if (P.d.T_max <= 32) {
    nth = 128;
    ks = rbf ? prior_inv64_small_kernel<GPKL_KERNEL_RBF, 128, 32> : prior_inv64_small_kernel<GPKL_KERNEL_CAUCHY, 128, 32>;
  } else {
    nth = 256;
    ks = rbf ? prior_inv64_small_kernel<GPKL_KERNEL_RBF, 256, 64> : prior_inv64_small_kernel<GPKL_KERNEL_CAUCHY, 256, 64>;
  }
  const int cap = (1024 / nth) * kNumSMs;
  ks<<<P.d.B < cap ? P.d.B : cap, nth, 0, st>>>(P, f32_off, f32_tm);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_prior_inv64(const Params& P, cudaStream_t st) {
  // (T_max <= 64 is the register tier's: launch_prior_inv64_small, whose record differs -- pitch prior64_pitch(), diagonal halved)
  if (P.d.T_max <= 64) return cudaErrorInvalidValue;
  const int TP = p64_tp(P.d.T_max);
  const size_t smem = ((size_t)2 * NB * p64_ldw(TP) + NB * (NB + 1) + TP) * sizeof(double) + (size_t)TP * sizeof(float);
  if (smem > kMaxDynSmem || P.prior_stride < prior64_record_floats(P.d.T_max)) return cudaErrorInvalidValue;
  void (*kern)(Params) = P.d.kernel == GPKL_KERNEL_RBF ? prior_inv64_kernel<GPKL_KERNEL_RBF> : prior_inv64_kernel<GPKL_KERNEL_CAUCHY>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // cluster size: as many CTAs per sequence as it takes to fill the chip with a small batch (8, 4, 2), one CTA per sequence
  // once there are more sequences than SMs (measured at B = 1024, T = 512: 23 ms with one CTA per sequence, 56 ms with
  // clusters of four -- the update's tile loads are L2 round trips either way and the redundant per-panel work is not free)
  const int cs = P.d.B <= 18 ? 8 : (P.d.B <= 37 ? 4 : (P.d.B <= 74 ? 2 : 1));
  int nclu = kNumSMs / cs;
  if (nclu > P.d.B) nclu = P.d.B;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(nclu * cs);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = cs;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, P);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace gpkl
