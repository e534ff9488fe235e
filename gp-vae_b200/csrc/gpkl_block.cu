// Block tier: one CTA per (sequence, latent-dim) pair, T > 64.  Three regimes (launch_kp):
//   resident    both work matrices in shared memory (T <= 144); prior and posterior chains on two half-CTA groups
//   one-buffer  the shared-prior kernels (SH): K_p factored once per sequence by a pre-pass, ONE work matrix per pair
//               (T <= 208), 1-3 CTAs per SM, look-ahead panel Cholesky, tiles dealt by contraction length
//   GEMM path   work matrices in per-CTA workspace slots (T <= 512), 64-wide panels in shared memory updated by a staged
//               128x64 cp.async GEMM tile; beyond 512 the resident code reading the slots through L1/L2
//
// Loop-based by design (the register-resident warp tier is instruction-fetch bound beyond T ~ 32, see
// profiles/r01_c2_warp_tier_unrolled_ncu_summary.txt): every O(T^3) phase is the same inner loop,
//
//     tile_update:  acc[4][4] (+/-)= sum_j  U[j*ldu + r0..r0+3] (x) V[j*ldv + c0..c0+3]
//
// i.e. a 4x4 register tile fed by two 128-bit shared-memory loads per 16 FMAs, over operands stored
// "k-major" (contraction index outermost).  A buffer is (TP+1) x ld floats (TP = T rounded up to 16,
// ld = TP+4) holding two triangles at once:
//     LC(i,k) = B[k*ld + i]      factor L, column-major  (i >= k)   -> k-major in the column index
//     XR(i,k) = B[(i+1)*ld + k]  inverse/solution X, row-major (i >= k) -> k-major in the row index
// Phases
//   Cholesky   left-looking by 16-column panels; K entries are GENERATED in the accumulator init (never
//              stored), panel -= L[:,0:j0] L[panel,0:j0]^T via tile_update, the 16x16 diagonal block is
//              factored and inverted by one warp in registers (shuffles), rows below are multiplied by the
//              inverse block (tile_update again).  The mean vector rides along as an extra matrix row, so
//              a = L_p^-1 m falls out of the factorisation.
//   Solve      X = L^-1 B by 16-row blocks: tile_update against the already solved rows, then the stored
//              inverse diagonal block (tile_update).  B = L_q gives A = L_p^-1 L_q, B = I gives L^-1.
//   Contract   sum_{k != l} dK(k,l)/d ell * (X_U^T X_V)_kl : tile_update over the row index, kernel
//              derivative evaluated in the epilogue (never stored).
// Math: SURVEY.md Appendix A; reference replaced: src/Models/Full_GP_VAE_dynamic_time.py:149-172, :174-195,
// :242-260 (+ TF autodiff :361); V2 src/Models/VAE_GPprior_diag_cov.py:64-71, :100-119.
#include <stdlib.h>
#include <string.h>

#include "gpkl_common.cuh"
#include "gpkl_diag.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int NB = 16;

#ifdef GPKL_PANEL_TRACE  // debug build only (tools/phase_trace.py): per-phase cycles of the panel loops
// Every thread keeps its own counters in registers (a thread-0-only probe makes the warp diverge, and the
// shuffles of diag_factor then take the compiler's slow divergent path); thread 0 of CTA 0 publishes them.
__device__ long long g_ptrace[8];
#define PT_DECL long long pt_t0 = clock64(), pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PT_ADD(slot) do { const long long pt_t1 = clock64(); pt_acc[slot] += pt_t1 - pt_t0; pt_t0 = pt_t1; } while (0)
#define PT_FLUSH do { if (blockIdx.x == 0 && g.tid == 0 && g.bar <= 1) for (int i_ = 0; i_ < 8; ++i_) g_ptrace[i_] += pt_acc[i_]; } while (0)
#else
#define PT_DECL
#define PT_ADD(slot)
#define PT_FLUSH
#endif

// Address-space hint.  The resident kernels keep their work matrices in shared memory, but the phase functions below are
// shared with the workspace (global memory) paths and are not inlined, so without the hint every operand access is a
// generic LD.E / ST.E (64-bit address arithmetic, longer latency) instead of LDS / STS: 1186 of the 1721 128-bit loads
// of the one-buffer backward kernel were generic (cuobjdump -sass).  SM = "the matrix operands are in shared memory".
#define GPKL_SHARED_HINT(p) __builtin_assume(__isShared(p))

template <int SGN>
__device__ __forceinline__ void tile_fma(float (&acc)[4][4], const float4& u4, const float4& v4) {
  const float u[4] = {SGN > 0 ? u4.x : -u4.x, SGN > 0 ? u4.y : -u4.y, SGN > 0 ? u4.z : -u4.z, SGN > 0 ? u4.w : -u4.w};
  const float v[4] = {v4.x, v4.y, v4.z, v4.w};
  // packed FFMA2: (acc[r][c], acc[r][c+1]) += u[r] * (v[c], v[c+1]); the scalar u[r] is a broadcast operand and
  // its sign a modifier of the instruction, so a 4x4 tile step is 8 instructions instead of 16
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; c += 2) fma2(acc[r][c], acc[r][c + 1], u[r], u[r], v[c], v[c + 1]);
}

// acc (+/-)= sum_{j in [ja, jb)} U[j*ldu + 0..3] (x) V[j*ldv + 0..3].  Main loop in groups of 8 steps with all
// 16 operand loads issued before the FMAs (memory-level parallelism when the operands come from L1/L2).
// VNEG: V is walked with stride -ldv (the row-reversed C' of the one-buffer backward).
template <int SGN, bool VNEG = false>
__device__ __forceinline__ void tile_update(float (&acc)[4][4], const float* __restrict__ U, int ldu,
                                            const float* __restrict__ V, int ldv, int ja, int jb) {
  const ptrdiff_t sv = VNEG ? -(ptrdiff_t)ldv : (ptrdiff_t)ldv;
  const float* up = U + (size_t)ja * ldu;
  const float* vp = V + (ptrdiff_t)ja * sv;
  int j = ja;
  for (; j + 8 <= jb; j += 8) {
    float4 u4[8], v4[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      u4[e] = *reinterpret_cast<const float4*>(up + (size_t)e * ldu);
      v4[e] = *reinterpret_cast<const float4*>(vp + (ptrdiff_t)e * sv);
    }
    up += (size_t)8 * ldu;
    vp += (ptrdiff_t)8 * sv;
#pragma unroll
    for (int e = 0; e < 8; ++e) tile_fma<SGN>(acc, u4[e], v4[e]);
  }
  for (; j < jb; ++j) {
    const float4 u4 = *reinterpret_cast<const float4*>(up);
    const float4 v4 = *reinterpret_cast<const float4*>(vp);
    up += ldu;
    vp += sv;
    tile_fma<SGN>(acc, u4, v4);
  }
}

// A warp-aligned group of threads that runs one chain of phases with its own barrier.  bar == 0: the whole
// CTA (__syncthreads); bar 1/2: the two half-CTA groups that factor K_p and K_q concurrently.
struct Grp {
  int tid, nt, bar;
};
// Barrier ids must be compile-time constants: a register barrier id makes ptxas reserve all 16 hardware
// barriers for the CTA (one CTA per SM).  DUAL = false kernels only ever use barrier 0.
template <bool DUAL>
__device__ __forceinline__ void grp_sync(const Grp& g) {
  if (!DUAL || g.bar == 0) {
    __syncthreads();
  } else if (g.bar == 1) {
    asm volatile("bar.sync 1, %0;" ::"r"(g.nt) : "memory");
  } else {
    asm volatile("bar.sync 2, %0;" ::"r"(g.nt) : "memory");
  }
}

// Work items sorted by decreasing cost are dealt to the warps of a group in chunks of 32, boustrophedon (round r goes
// w = 0..nw-1, round r+1 back): every warp gets the same total and the lanes of a warp the same trip count.  Phases
// end at a barrier, so the slowest warp is the phase time.  Returns the item index of this thread in round r.
__device__ __forceinline__ int dealt_index(int r, const Grp& g) {
  const int w = g.tid >> 5, nw = g.nt >> 5;
  return (r * nw + ((r & 1) ? nw - 1 - w : w)) * 32 + (g.tid & 31);
}

// Workspace variant (T > ~144): operand slabs are staged global -> shared with cp.async, double buffered.
// staging floats: >= 2*32*128 for the 64x64 staged contraction, and stg + pan (16*(TP+4) floats, TP >= 160 on the
// GEMM path) must hold the GM_NS stages of the GEMM tile: 9728 + 16*164 = 12352 >= 4*16*192 = 12288
constexpr int STG_FLOATS = 9728;
constexpr int GM_KC = 16;                             // GEMM phase: contraction steps per stage
constexpr int GM_SLD = 192;                           // floats per staged step: 128 A values | 64 B values
constexpr int GM_NS = 4;                              // pipeline depth (4 x 12 KB <= stg + pan, see STG_FLOATS)
static_assert(STG_FLOATS + 16 * 164 >= GM_NS * GM_KC * GM_SLD, "staging area too small for the GEMM pipeline");
static_assert(STG_FLOATS >= 2 * 32 * 128, "staging area too small for the staged contraction");
constexpr int NBL = 64;                               // large-T panel width (columns / rows per GEMM phase)
constexpr int GEMM_TMAX = 512;                        // the shared-memory panel fits up to this T
constexpr int GEMM_TMIN = 144;                        // below this the per-panel overheads outweigh the GEMM phase


struct Lay {  // shared-memory carve-up (floats), identical on host and device
  int TP, ld, nP, S;
  __host__ __device__ Lay(int Tmax, int S_) : S(S_) {
    // beyond the shared-memory resident sizes the padded size is a multiple of the tile tier's 64 x 64 tiles
    // (gpkl_tile.cu reads the prior records this tier's pre-pass writes); Tact (16-rounded) still bounds the work
    const int al = Tmax > 208 ? 64 : NB;
    TP = (Tmax + al - 1) / al * al;
    if (TP < NB) TP = NB;
    ld = TP + 4;
    nP = TP / NB;
  }
  __host__ __device__ size_t buf() const { return (size_t)(TP + 1) * ld; }
  // shared floats; resident = the two work matrices live in shared memory (else in a workspace slot)
  // dual: the two chains run concurrently and need a staging panel each
  __host__ __device__ bool dual(bool resident) const { return resident && TP > 64; }
  // gemm: workspace path whose 64-wide panels live in shared memory and are updated by the staged GEMM tile
  __host__ __device__ bool gemm(bool resident) const { return !resident && TP > GEMM_TMIN && TP <= GEMM_TMAX; }
  // onebuf: the resident shared-prior kernels need ONE work matrix (B2: L_q / X_q, and in backward the row-reversed
  // C' in the triangle L_q vacates; the prior record is read from global memory through L1), single chain; their
  // second staging panel is the look-ahead panel of the factorisation
  __host__ __device__ size_t floats(bool resident, bool onebuf = false) const {
    return 64 + (resident ? (onebuf ? 1 : 2) * buf() : (size_t)STG_FLOATS) + (gemm(resident) ? (size_t)NBL * ld : 0) +
           ((dual(resident) || onebuf) ? 2 : 1) * (size_t)NB * ld + ld + 9 * (size_t)TP + 3 * (size_t)S * TP;
  }
};

struct Sm {
  double* red;
  float *B1, *B2, *stg, *wide, *pan, *pan2, *rdp, *rdq, *ts, *dgp, *dgq, *aa, *al, *pd, *gzs, *mm, *u, *v, *w;
  __device__ Sm(float* base, const Lay& L, float* slot, bool onebuf = false) {
    red = reinterpret_cast<double*>(base); base += 64;
    stg = nullptr;
    wide = nullptr;
    if (slot) {  // large T: matrices in this CTA's workspace slot; operand slabs staged through stg
      B1 = slot;
      B2 = slot + L.buf();
      wide = base;  // 64-wide panel (Cholesky) / 64-row block (solve) of the GEMM path
      if (L.gemm(false)) base += (size_t)NBL * L.ld;
      stg = base; base += STG_FLOATS;  // immediately followed by pan: GEMM phases stage through stg + pan
    } else if (onebuf) {
      B1 = nullptr;
      B2 = base; base += L.buf();
    } else {
      B1 = base; base += L.buf();
      B2 = base; base += L.buf();
    }
    pan = base; base += (size_t)NB * L.ld;
    pan2 = base;
    if (L.dual(slot == nullptr) || onebuf) base += (size_t)NB * L.ld;
    ts = base; base += L.ld;
    rdp = base; base += L.TP;
    rdq = base; base += L.TP;
    dgp = base; base += L.TP;
    dgq = base; base += L.TP;
    aa = base; base += L.TP;
    al = base; base += L.TP;
    pd = base; base += L.TP;
    gzs = base; base += L.TP;
    mm = base; base += L.TP;
    u = base; base += (size_t)L.S * L.TP;
    v = base; base += (size_t)L.S * L.TP;
    w = base;
  }
};

// Left-looking panel Cholesky with fused kernel-matrix generation.  Result: LC triangle of Bm, inverse
// diagonal blocks in inv[nP][256], diag(L) in dg.  extra: also carry row TP = m^T (gives L^-1 m).
// Bm is a VIEW: element (i,k) at Bm[k*ldm + i].  Columns [c_begin, c_end) are factored; contributions of columns
// < kstart are assumed applied already, and with FROM_VIEW the starting values are read from the view (the
// GEMM phase of the large-T path left them there) instead of being generated.
// XRC (resident buffers only): the factor is ALSO written row-major into the XR triangle of the same buffer
// (element (i,k) at Bm[(i+1)*ldm + k]) -- the k-major operand of the shared-prior product A = L_p^-1 L_q.
// Tile phase of one 16-column panel at j0: every (4-row x 4-column) tile of rows >= j0 is initialised (INIT 0: kernel
// entries generated on the fly / the extra m^T row / identity padding; INIT 1: read from the view Bm; INIT 2: read from the
// staging panel `src`, a partial result of this same phase), updated with  acc -= sum_{k in [ka,kb)} L[rows,k] L[cols,k]
// and parked in the staging panel `dst` (column-major, stride ld).  (tid, NT): the threads that share the phase.
template <int KERNEL, int INIT>
__device__ __forceinline__ void panel_tiles(const float* __restrict__ Bm, int ldm, int j0, int ka, int kb, const Lay& L, int T,
                                            bool extra, const float* __restrict__ ts, const float* __restrict__ mm,
                                            const KernC<KERNEL>& kc, float noise, const float* __restrict__ src,
                                            float* __restrict__ dst, int tid, int NT) {
  const int cg = tid & 3, rg = tid >> 2, NRG = NT >> 2;
  const int ld = L.ld, TP = L.TP;
  const int Tact = (T + NB - 1) / NB * NB;
  const int rows_end = extra ? TP + 4 : Tact;
  const int cb = j0 + 4 * cg;
  for (int rb = j0 + 4 * rg; rb < rows_end; rb += 4 * NRG) {
    if (rb >= Tact && rb < TP) continue;  // identity padding rows
    float acc[4][4];
    if (INIT == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 a4 = *reinterpret_cast<const float4*>(src + (size_t)(4 * cg + c) * ld + rb);
        acc[0][c] = a4.x; acc[1][c] = a4.y; acc[2][c] = a4.z; acc[3][c] = a4.w;
      }
    } else if (INIT == 1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 a4 = *reinterpret_cast<const float4*>(Bm + (size_t)(cb + c) * ldm + rb);
        acc[0][c] = a4.x; acc[1][c] = a4.y; acc[2][c] = a4.z; acc[3][c] = a4.w;
      }
    } else if (rb + 3 < T && cb + 3 < T) {  // all-real tile: no padding logic
      const float4 tr4 = *reinterpret_cast<const float4*>(ts + rb);
      const float4 tc4 = *reinterpret_cast<const float4*>(ts + cb);
      const float tr[4] = {tr4.x, tr4.y, tr4.z, tr4.w}, tc[4] = {tc4.x, tc4.y, tc4.z, tc4.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = kc.val(tr[r] - tc[c]);
      if (rb == cb) {
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][r] += noise;
      }
    } else if (rb == TP) {  // the extra row block: row TP carries m^T, rows TP+1..TP+3 are zero
      const float4 m4 = *reinterpret_cast<const float4*>(mm + cb);
      const float mv[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[0][c] = (cb + c < T) ? mv[c] : 0.0f;
        acc[1][c] = acc[2][c] = acc[3][c] = 0.0f;
      }
    } else {  // tile touching the identity padding: branch-free selects (divergent branches cost ~20 cycles each)
      const float4 tr4 = *reinterpret_cast<const float4*>(ts + rb);
      const float4 tc4 = *reinterpret_cast<const float4*>(ts + cb);
      const float tr[4] = {tr4.x, tr4.y, tr4.z, tr4.w}, tc[4] = {tc4.x, tc4.y, tc4.z, tc4.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = rb + r, k = cb + c;
          const float v = kc.val(tr[r] - tc[c]) + (i == k ? noise : 0.0f);
          acc[r][c] = (i < T && k < T) ? v : (i == k ? 1.0f : 0.0f);
        }
    }
    tile_update<-1>(acc, Bm + rb, ldm, Bm + cb, ldm, ka, kb);
#pragma unroll
    for (int c = 0; c < 4; ++c)
      *reinterpret_cast<float4*>(dst + (size_t)(4 * cg + c) * ld + rb) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
  }
}

// Left-looking panel Cholesky with fused kernel-matrix generation.  Result: LC triangle of Bm, diag(L) in dg, 1/diag(L)
// in rdg.  extra: also carry row TP = m^T (gives L^-1 m).
// Bm is a VIEW: element (i,k) at Bm[k*ldm + i].  Columns [c_begin, c_end) are factored; contributions of columns
// < kstart are assumed applied already, and with FROM_VIEW the starting values are read from the view (the
// GEMM phase of the large-T path left them there) instead of being generated.
// XRC (resident buffers only): the factor is ALSO written row-major into the XR triangle of the same buffer
// (element (i,k) at Bm[(i+1)*ldm + k]) -- the k-major operand of the shared-prior product A = L_p^-1 L_q.
// pan2 != NULL: LOOK-AHEAD.  The serial 16x16 diagonal factor (one warp, ~3 K cycles) used to idle the other warps at a
// barrier; now, while warp 0 factors the diagonal block of panel j, the other warps already apply the finished columns
// [kstart, j0) to panel j+1 (into pan2), and only the 16 columns of panel j itself are applied after its rows are solved.
template <int KERNEL, bool DUAL, bool FROM_VIEW, bool XRC = false, bool SM = false>
__device__ __noinline__ void chol_panels(float* __restrict__ Bm, int ldm, int c_begin, int c_end, int kstart, const Lay& L,
                                         int T, bool extra, const float* __restrict__ ts, const float* __restrict__ mm,
                                         float ell, float sig, float noise, float* __restrict__ pan,
                                         float* __restrict__ dg, float* __restrict__ rdg, int* bad, Grp g,
                                         float* __restrict__ pan2 = nullptr) {
  if (SM) {
    GPKL_SHARED_HINT(Bm); GPKL_SHARED_HINT(pan); GPKL_SHARED_HINT(dg); GPKL_SHARED_HINT(rdg); GPKL_SHARED_HINT(ts);
    GPKL_SHARED_HINT(mm);
    if (pan2) GPKL_SHARED_HINT(pan2);
  }
  const int tid = g.tid, NT = g.nt;
  const int ld = L.ld, TP = L.TP;  // ld: stride of the staging panel `pan`
  const KernC<KERNEL> kc(ell, sig);
  // rows beyond the last real row are identity padding and decouple: only panels that contain real rows matter
  const int Tact = (T + NB - 1) / NB * NB;
  const bool ahead = pan2 != nullptr && NT > 32;
  PT_DECL;
  if (ahead) panel_tiles<KERNEL, FROM_VIEW ? 1 : 0>(Bm, ldm, c_begin, kstart, c_begin, L, T, extra, ts, mm, kc, noise, nullptr, pan, tid, NT);
  for (int j0 = c_begin; j0 < c_end; j0 += NB) {
    const bool more = j0 + NB < c_end;
    if (!ahead) panel_tiles<KERNEL, FROM_VIEW ? 1 : 0>(Bm, ldm, j0, kstart, j0, L, T, extra, ts, mm, kc, noise, nullptr, pan, tid, NT);
    PT_ADD(0);
    grp_sync<DUAL>(g);
    PT_ADD(1);
    if (tid < 32) diag_factor<XRC>(Bm, ldm, j0, T, pan, ld, dg, rdg, bad);
    else if (ahead && more)
      panel_tiles<KERNEL, FROM_VIEW ? 1 : 0>(Bm, ldm, j0 + NB, kstart, j0, L, T, extra, ts, mm, kc, noise, nullptr, pan2, tid - 32, NT - 32);
    PT_ADD(2);
    grp_sync<DUAL>(g);
    PT_ADD(1);
    // rows below the diagonal block: L[i, j0:j0+16] = pan[i, :] L_dd^-T, one row per thread
    const int nbelow = Tact - j0 - NB;
    const int nrows = nbelow + (extra ? 1 : 0);
    for (int t = tid; t < nrows; t += NT) {
      const int i = t < nbelow ? j0 + NB + t : TP;
      float b[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) b[c] = pan[(size_t)c * ld + i];
      diag_solve16(b, Bm, ldm, j0, rdg);
#pragma unroll
      for (int c = 0; c < 16; ++c) Bm[(size_t)(j0 + c) * ldm + i] = b[c];
      if (XRC && i < TP) {
        float* xr = Bm + (size_t)(i + 1) * ldm + j0;
#pragma unroll
        for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(xr + c) = make_float4(b[c], b[c + 1], b[c + 2], b[c + 3]);
      }
    }
    PT_ADD(3);
    grp_sync<DUAL>(g);
    PT_ADD(1);
    if (ahead && more)  // the 16 columns just finished, applied to the partial panel j+1
      panel_tiles<KERNEL, 2>(Bm, ldm, j0 + NB, j0, j0 + NB, L, T, extra, ts, mm, kc, noise, pan2, pan, tid, NT);
  }
  PT_FLUSH;
  // identity padding: diag entries for rows in [Tact, TP) (only dg / rdg are consulted for them)
  if (c_end >= Tact) {
    for (int i = Tact + tid; i < TP; i += NT) { dg[i] = 1.0f; rdg[i] = 1.0f; }
  }
  grp_sync<DUAL>(g);
}

// Resident path: the whole factorisation, kernel entries generated on the fly.
template <int KERNEL, bool DUAL, bool XRC = false, bool SM = false>
__device__ __forceinline__ void chol_block(float* __restrict__ Bm, const Lay& L, int T, bool extra, const float* __restrict__ ts,
                                           const float* __restrict__ mm, float ell, float sig, float noise,
                                           float* __restrict__ pan, float* __restrict__ dg, float* __restrict__ rdg, int* bad,
                                           Grp g, float* __restrict__ pan2 = nullptr) {
  const int Tact = (T + NB - 1) / NB * NB;
  chol_panels<KERNEL, DUAL, false, XRC, SM>(Bm, L.ld, 0, Tact, 0, L, T, extra, ts, mm, ell, sig, noise, pan, dg, rdg, bad, g, pan2);
}

// X = L^-1 B by 16-row blocks into the XR triangle of Xb.  L: LC triangle of Lb with inverse diagonal
// reciprocals rdgL.  IDENT: B = I, else B = LC triangle of Bb.  Returns this thread's partial sum of squares
// of the strictly-lower entries of X (rows/cols < T).
// Xb is a VIEW of the solution: X(i,k) at Xb[(i+1)*ldx + k].  Rows [r_begin, r_end) are solved; contributions of
// rows < kstart are assumed applied already and with FROM_VIEW the starting values are read from the view (the GEMM
// phase of the large-T path left them there) instead of the right-hand side.
template <bool IDENT, bool DUAL, bool FROM_VIEW, bool SM = false>
__device__ __noinline__ float solve_rows(const float* __restrict__ Lb, const float* __restrict__ rdgL, const float* __restrict__ Bb,
                                         float* __restrict__ Xb, int ldx, int r_begin, int r_end, int kstart, const Lay& L,
                                         int T, float* __restrict__ pan, Grp g) {
  if (SM) {
    GPKL_SHARED_HINT(Lb); GPKL_SHARED_HINT(rdgL); GPKL_SHARED_HINT(Xb); GPKL_SHARED_HINT(pan);
    if (!IDENT) GPKL_SHARED_HINT(Bb);
  }
  const int tid = g.tid, NT = g.nt;
  const int ld = L.ld;
  float ssq = 0.0f;
  PT_DECL;
  for (int i0 = r_begin; i0 < r_end; i0 += NB) {
    const int ntile = 4 * (i0 / 4 + 4);
    for (int id = tid; id < ntile; id += NT) {
      const int rt = id & 3, ct = id >> 2;
      const int rb = i0 + 4 * rt, cb = 4 * ct;
      // (tiles entirely above the diagonal are staged too -- as zeros -- because the diagonal-block
      //  solve below reads every staged row of its column)
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int row = rb + r, col = cb + c;
          if (FROM_VIEW) acc[r][c] = Xb[(size_t)(row + 1) * ldx + col];
          else acc[r][c] = (col <= row) ? (IDENT ? (row == col ? 1.0f : 0.0f) : Bb[(size_t)col * ld + row]) : 0.0f;
        }
      // contraction over already solved rows k in [max(cb, kstart), i0); X[k][col] exists only for col <= k
      int k = cb < i0 ? cb : i0;
      if (k < kstart) k = kstart;
      const int khead = (k + 3 < i0) ? k + 3 : i0;
      for (; k < khead; ++k) {
        const float4 u4 = *reinterpret_cast<const float4*>(Lb + (size_t)k * ld + rb);
        const float4 v4 = *reinterpret_cast<const float4*>(Xb + (size_t)(k + 1) * ldx + cb);
        const float u[4] = {u4.x, u4.y, u4.z, u4.w};
        const float v[4] = {cb <= k ? v4.x : 0.0f, cb + 1 <= k ? v4.y : 0.0f, cb + 2 <= k ? v4.z : 0.0f,
                            cb + 3 <= k ? v4.w : 0.0f};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(-u[r], v[c], acc[r][c]);
      }
      tile_update<-1>(acc, Lb + rb, ld, Xb + ldx + cb, ldx, k, i0);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(pan + (size_t)(4 * rt + r) * ld + cb) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    }
    PT_ADD(4);
    grp_sync<DUAL>(g);
    PT_ADD(5);
    // diagonal block: X[i0:i0+16, col] = L_dd^-1 staged[:, col], one column per thread
    for (int col = tid; col < i0 + NB; col += NT) {
      float b[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) b[r] = pan[(size_t)r * ld + col];
      diag_solve16(b, Lb, ld, i0, rdgL);
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const int row = i0 + r;
        if (col <= row) {
          Xb[(size_t)(row + 1) * ldx + col] = b[r];
          if (col < row && row < T) ssq = fmaf(b[r], b[r], ssq);
        }
      }
    }
    PT_ADD(6);
    grp_sync<DUAL>(g);
    PT_ADD(5);
  }
  PT_FLUSH;
  return ssq;
}

// Resident path: X = L^-1 B over all rows.
template <bool IDENT, bool DUAL, bool SM = false>
__device__ __forceinline__ float solve_block(const float* __restrict__ Lb, const float* __restrict__ rdgL,
                                             const float* __restrict__ Bb, float* __restrict__ Xb, const Lay& L, int T,
                                             float* __restrict__ pan, Grp g) {
  const int Tact = (T + NB - 1) / NB * NB;
  return solve_rows<IDENT, DUAL, false, SM>(Lb, rdgL, Bb, Xb, L.ld, 0, Tact, 0, L, T, pan, g);
}

// ---- GEMM-structured phases of the workspace path (144 < T <= 512, 256 threads) -------------------------------
// acc[8][4] (+/-)= sum_{k in [k0,k1)} A(arow0 + 8*ty + r, k) * B(bcol0 + 4*tx + c, k),  ty = tid/16, tx = tid%16.
// Operands are k-major in global memory (A(i,k) at Ag[k*lda + i], B(j,k) at Bg[k*ldb + j]) and are staged 16 steps
// at a time with cp.async, double buffered: one stage = 768 16-byte copies, ~512 FMAs per thread per barrier pair.
// Rows >= arow_lim / columns >= bcol_lim are zero-filled; with A_TRI also A(i,k) for i > k (the row-major lower
// triangle of an already solved X, whose other triangle holds a different matrix).
template <int SGN, bool A_TRI, bool B_TRI = false>
__device__ __forceinline__ void gemm_tile_128x64(float (&acc)[8][4], const float* __restrict__ Ag, int lda, int arow0,
                                                 int arow_lim, const float* __restrict__ Bg, int ldb, int bcol0,
                                                 int bcol_lim, int k0, int k1, float* __restrict__ stg) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int nch = (k1 - k0 + GM_KC - 1) / GM_KC;
  // Each thread copies the same three 16-byte slots of every stage (16 steps x 48 float4 = 768 slots / 256 threads):
  // their stage offset, global pointer, contraction step and limits are set up ONCE per call and advanced by one
  // stage per issue, so a stage costs ~10 instructions per slot instead of the full index arithmetic.
  const float* gp[3];
  int kq[3], klim[3], rce[3], so[3], gstep[3];
#pragma unroll
  for (int rnd = 0; rnd < 3; ++rnd) {
    const int q = tid + 256 * rnd;
    const int kk = q / 48, e = q - kk * 48;
    const bool isA = e < 32;
    const int rc = isA ? arow0 + 4 * e : bcol0 + 4 * (e - 32);
    const bool inb = rc < (isA ? arow_lim : bcol_lim);
    const bool tri = isA ? A_TRI : B_TRI;
    const int ldx = isA ? lda : ldb;
    klim[rnd] = inb ? k1 : -0x40000000;        // out-of-range rows / columns: never valid (zero-filled)
    rce[rnd] = tri ? rc : -0x20000000;         // triangular operand: elements rc .. rc+3 exist while <= k
    so[rnd] = kk * GM_SLD + (isA ? 4 * e : 128 + 4 * (e - 32));
    kq[rnd] = k0 + kk;
    gstep[rnd] = GM_KC * ldx;
    gp[rnd] = (isA ? Ag : Bg) + (size_t)kq[rnd] * ldx + rc;
  }
  int issued = 0;
  auto issue = [&](int c) {  // called with c = 0, 1, 2, ... in order
    float* buf = stg + (issued & (GM_NS - 1)) * (GM_KC * GM_SLD);
    if (c < nch) {
#pragma unroll
      for (int rnd = 0; rnd < 3; ++rnd) {
        int v = kq[rnd] - rce[rnd] + 1;
        v = v > 4 ? 4 : v;
        const int valid = (kq[rnd] < klim[rnd] && v > 0) ? v : 0;
        cp_async16(buf + so[rnd], valid ? gp[rnd] : Ag, 4 * valid);
        gp[rnd] += gstep[rnd];
        kq[rnd] += GM_KC;
      }
    }
    ++issued;
    cp_async_commit();  // (possibly empty) group: keeps the group count uniform
  };
  // GM_NS-deep pipeline, one barrier per stage: the barrier of iteration c publishes stage c and also guarantees
  // that every thread is done with stage c-1, whose buffer the copy issued right after it overwrites.
#pragma unroll
  for (int c = 0; c < GM_NS - 1; ++c) issue(c);
  for (int c = 0; c < nch; ++c) {
    cp_async_wait<GM_NS - 2>();
    __syncthreads();
    issue(c + GM_NS - 1);
    const float* buf = stg + (c % GM_NS) * (GM_KC * GM_SLD);
#pragma unroll
    for (int kk = 0; kk < GM_KC; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(buf + kk * GM_SLD + 8 * ty);
      const float4 a1 = *reinterpret_cast<const float4*>(buf + kk * GM_SLD + 8 * ty + 4);
      const float4 b4 = *reinterpret_cast<const float4*>(buf + kk * GM_SLD + 128 + 4 * tx);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int cc = 0; cc < 4; cc += 2) {
          const float ar = SGN > 0 ? a[r] : -a[r];
          fma2(acc[r][cc], acc[r][cc + 1], ar, ar, b[cc], b[cc + 1]);
        }
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // the staging area (stg + pan) is free again
}

// Large-T Cholesky: 64-column panels.  GEMM phase: panel = K - L[:, 0:J0] L[J0:J0+64, 0:J0]^T into the shared-memory
// panel `wide` (K generated in the accumulator init); then the panel is factored in shared memory by the 16-column
// code on a view of `wide`; then its lower triangle is copied to the global factor.
// XRC: the panel is ALSO copied out row-major into the XR triangle (element (i,k) at Bm[(i+1)*ld + k]).
template <int KERNEL, bool XRC = false>
__device__ __noinline__ void chol_gemm(float* __restrict__ Bm, const Lay& L, int T, bool extra, const float* __restrict__ ts,
                                       const float* __restrict__ mm, float ell, float sig, float noise,
                                       float* __restrict__ pan, float* __restrict__ wide, float* __restrict__ stg,
                                       float* __restrict__ dg, float* __restrict__ rdg, int* bad) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ld = L.ld, TP = L.TP;
  const KernC<KERNEL> kc(ell, sig);
  const int Tact = (T + NB - 1) / NB * NB;
  const int rows_end = extra ? TP + 4 : Tact;
  const Grp all{tid, (int)blockDim.x, 0};
  for (int J0 = 0; J0 < Tact; J0 += NBL) {
    const int ncols = (Tact - J0 < NBL) ? Tact - J0 : NBL;
    float* Pv = wide - (size_t)J0 * ld;  // view: column c (absolute) at Pv[c*ld + row]
    const int cb = J0 + 4 * tx;
    for (int rb0 = J0; rb0 < rows_end; rb0 += 128) {
      float acc[8][4];
      float tc[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) tc[c] = (cb + c < TP) ? ts[cb + c] : 0.0f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int i = rb0 + 8 * ty + r;
        const float ti = (i < ld) ? ts[i] : 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int k = cb + c;
          const float v = kc.val(ti - tc[c]) + (i == k ? noise : 0.0f);
          const float pad = (i == k) ? 1.0f : ((extra && i == TP && k < T) ? mm[k < TP ? k : 0] : 0.0f);
          acc[r][c] = (i < T && k < T) ? v : pad;
        }
      }
      gemm_tile_128x64<-1, false>(acc, Bm, ld, rb0, rows_end, Bm, ld, J0, J0 + ncols, 0, J0, stg);
      const int row0 = rb0 + 8 * ty;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (4 * tx + c < ncols) {
          float* dst = Pv + (size_t)(cb + c) * ld + row0;
          if (row0 < rows_end) *reinterpret_cast<float4*>(dst) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
          if (row0 + 4 < rows_end) *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][c], acc[5][c], acc[6][c], acc[7][c]);
        }
      }
    }
    __syncthreads();
    chol_panels<KERNEL, false, true>(Pv, ld, J0, J0 + ncols, J0, L, T, extra, ts, mm, ell, sig, noise, pan, dg, rdg, bad, all);
    // copy the finished panel's lower triangle (and the extra rows) to the global factor
    const int nr4 = rows_end >> 2;
    for (int idx = tid; idx < ncols * nr4; idx += blockDim.x) {
      const int cl = idx / nr4, r4 = 4 * (idx - cl * nr4);
      const int c = J0 + cl;
      if (r4 + 3 < c) continue;              // above the diagonal: the other triangle of the buffer
      if (r4 >= Tact && r4 < TP) continue;   // identity padding rows are never written
      const float4 v = *reinterpret_cast<const float4*>(Pv + (size_t)c * ld + r4);
      float* dst = Bm + (size_t)c * ld + r4;
      if (r4 >= c) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        if (r4 + 1 >= c) dst[1] = v.y;
        if (r4 + 2 >= c) dst[2] = v.z;
        dst[3] = v.w;
      }
    }
    if (XRC) {
      // row-major copy: consecutive threads take consecutive rows (conflict-free panel reads), each writes the four
      // columns c4..c4+3 of its row as one 16-byte store (scalars where the group straddles the diagonal)
      const int nrows = Tact - J0;
      for (int idx = tid; idx < (ncols >> 2) * nrows; idx += blockDim.x) {
        const int cq = idx / nrows, r = J0 + (idx - cq * nrows);
        const int c4 = J0 + 4 * cq;
        if (c4 > r) continue;
        const float v0 = Pv[(size_t)c4 * ld + r], v1 = Pv[(size_t)(c4 + 1) * ld + r], v2 = Pv[(size_t)(c4 + 2) * ld + r],
                    v3 = Pv[(size_t)(c4 + 3) * ld + r];
        float* dst = Bm + (size_t)(r + 1) * ld + c4;
        if (c4 + 3 <= r) {
          *reinterpret_cast<float4*>(dst) = make_float4(v0, v1, v2, v3);
        } else {
          dst[0] = v0;
          if (c4 + 1 <= r) dst[1] = v1;
          if (c4 + 2 <= r) dst[2] = v2;
        }
      }
    }
    __syncthreads();
  }
}

// Large-T solve X = L^-1 B by 64-row blocks: GEMM phase rows[I0:I0+64] = B - L[rows, 0:I0] X[0:I0, :] into the
// shared-memory block `wide` (X operand zero-filled above its diagonal), then the 64x64 diagonal part by the 16-row
// code on a view of `wide`, then the rows are copied to the global XR triangle.
template <bool IDENT>
__device__ __noinline__ float solve_gemm(const float* __restrict__ Lb, const float* __restrict__ rdgL,
                                         const float* __restrict__ Bb, float* __restrict__ Xb, const Lay& L, int T,
                                         float* __restrict__ pan, float* __restrict__ wide, float* __restrict__ stg) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ld = L.ld;
  const int Tact = (T + NB - 1) / NB * NB;
  const Grp all{tid, (int)blockDim.x, 0};
  float ssq = 0.0f;
  for (int I0 = 0; I0 < Tact; I0 += NBL) {
    const int nrows = (Tact - I0 < NBL) ? Tact - I0 : NBL;
    const int ncols = I0 + nrows;
    float* Qv = wide - (size_t)(I0 + 1) * ld;  // X view: X(i,k) at Qv[(i+1)*ld + k]
    for (int cb0 = 0; cb0 < ncols; cb0 += 128) {
      float acc[8][4];  // acc[a][b] = X[I0 + 4*tx + b][cb0 + 8*ty + a]
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int col = cb0 + 8 * ty + a, row = I0 + 4 * tx + b;
          acc[a][b] = (row < ncols && col <= row) ? (IDENT ? (row == col ? 1.0f : 0.0f) : Bb[(size_t)col * ld + row]) : 0.0f;
        }
      // A operand: X(k, col) for k < I0 (global, zero above its diagonal); B operand: L(I0 + r, k)
      gemm_tile_128x64<-1, true>(acc, Xb + ld, ld, cb0, ld, Lb, ld, I0, I0 + nrows, cb0, I0, stg);
      const int col0 = cb0 + 8 * ty;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int row = I0 + 4 * tx + b;
        if (row < ncols && col0 < ld) {
          float* dst = Qv + (size_t)(row + 1) * ld + col0;
          *reinterpret_cast<float4*>(dst) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
          if (col0 + 4 < ld) *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4][b], acc[5][b], acc[6][b], acc[7][b]);
        }
      }
    }
    __syncthreads();
    ssq += solve_rows<IDENT, false, true>(Lb, rdgL, Bb, Qv, ld, I0, I0 + nrows, I0, L, T, pan, all);
    const int nc4 = ncols >> 2;
    for (int idx = tid; idx < nrows * nc4; idx += blockDim.x) {
      const int r = idx / nc4, c4 = 4 * (idx - r * nc4);
      const int row = I0 + r;
      if (c4 > row) continue;
      const float4 v = *reinterpret_cast<const float4*>(wide + (size_t)r * ld + c4);
      float* dst = Xb + (size_t)(row + 1) * ld + c4;
      if (c4 + 3 <= row) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        dst[0] = v.x;
        if (c4 + 1 <= row) dst[1] = v.y;
        if (c4 + 2 <= row) dst[2] = v.z;
      }
    }
    __syncthreads();
  }
  return ssq;
}

// sum_{k != l, k,l < T} dK(k,l)/d ell * sum_{i} XU[i][k] XV[i][l]   (XR triangles of Ub / Vb); thread partial.
// kinv != NULL (shared-prior path): the weight of dK(k,l) becomes hg * kinv[k*ld + l] + (U^T V)_kl, i.e. the prior-side
// term g/2 <K_p^-1, dK_q/d ell> rides along in the same epilogue (kinv = this sequence's K_p^-1 record, row stride ld).
// VREV (one-buffer backward): V = C' is stored row- and column-reversed in the triangle of Vb that L_q vacated,
// C'(i,l) at Vb[(TP-1-i)*ld + (TP-1-l)] (l <= i), so that a row of C' is still one aligned float4 per 4 columns; the
// tile then holds its columns in reverse order (acc[r][c] <-> l = lb + 3 - c).
template <int KERNEL, bool VREV = false, bool SM = false>
__device__ __noinline__ double contract_block(const float* __restrict__ Ub, const float* __restrict__ Vb, const Lay& L, int T,
                                 const float* __restrict__ ts, float ell, float sig, Grp g,
                                 const float* __restrict__ kinv = nullptr, float hg = 0.0f) {
  if (SM) { GPKL_SHARED_HINT(Ub); GPKL_SHARED_HINT(Vb); GPKL_SHARED_HINT(ts); }
  const int NT = g.nt;
  const int ld = L.ld, TP = L.TP;
  const int nk = (T + 3) / 4;
  const KernC<KERNEL> kc(ell, sig);
  double total = 0.0;
  // tiles enumerated by shell m = max(kt, lt) (2m+1 tiles, all with the contraction range [4m, T)), longest first
  for (int rr = 0; rr * NT < nk * nk; ++rr) {
    const int n = dealt_index(rr, g);
    if (n >= nk * nk) continue;
    int m = (int)sqrtf((float)n);
    while ((m + 1) * (m + 1) <= n) ++m;
    while (m * m > n) --m;
    const int pos = n - m * m;
    const int kt = pos <= m ? m : pos - m - 1, lt = pos <= m ? pos : m;
    const int kb = 4 * kt, lb = 4 * lt;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    int i = kb > lb ? kb : lb;
    const int ihead = (i + 3 < T) ? i + 3 : T;
    // row i of V: VREV ? Vb + (TP-1-i)*ld + (TP-4-lb) (components l = lb+3 .. lb) : Vb + (i+1)*ld + lb
    const float* __restrict__ v0 = VREV ? Vb + (size_t)(TP - 1) * ld + (TP - 4 - lb) : Vb + ld + lb;
    for (; i < ihead; ++i) {
      const float4 u4 = *reinterpret_cast<const float4*>(Ub + (size_t)(i + 1) * ld + kb);
      const float4 v4 = *reinterpret_cast<const float4*>(VREV ? v0 - (size_t)i * ld : v0 + (size_t)i * ld);
      const float u[4] = {kb <= i ? u4.x : 0.0f, kb + 1 <= i ? u4.y : 0.0f, kb + 2 <= i ? u4.z : 0.0f, kb + 3 <= i ? u4.w : 0.0f};
      float v[4];
      if (VREV) {
        v[0] = lb + 3 <= i ? v4.x : 0.0f; v[1] = lb + 2 <= i ? v4.y : 0.0f; v[2] = lb + 1 <= i ? v4.z : 0.0f; v[3] = lb <= i ? v4.w : 0.0f;
      } else {
        v[0] = lb <= i ? v4.x : 0.0f; v[1] = lb + 1 <= i ? v4.y : 0.0f; v[2] = lb + 2 <= i ? v4.z : 0.0f; v[3] = lb + 3 <= i ? v4.w : 0.0f;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(u[r], v[c], acc[r][c]);
    }
    tile_update<1, VREV>(acc, Ub + ld + kb, ld, v0, ld, i, T);
    if (kinv) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(kinv + (size_t)(kb + r) * ld + lb));
        if (VREV) {
          acc[r][0] = fmaf(hg, q.w, acc[r][0]); acc[r][1] = fmaf(hg, q.z, acc[r][1]);
          acc[r][2] = fmaf(hg, q.y, acc[r][2]); acc[r][3] = fmaf(hg, q.x, acc[r][3]);
        } else {
          acc[r][0] = fmaf(hg, q.x, acc[r][0]); acc[r][1] = fmaf(hg, q.y, acc[r][1]);
          acc[r][2] = fmaf(hg, q.z, acc[r][2]); acc[r][3] = fmaf(hg, q.w, acc[r][3]);
        }
      }
    }
    const float4 tk4 = *reinterpret_cast<const float4*>(ts + kb);
    const float4 tl4 = *reinterpret_cast<const float4*>(ts + lb);
    const float tk[4] = {tk4.x, tk4.y, tk4.z, tk4.w};
    const float tl[4] = {VREV ? tl4.w : tl4.x, VREV ? tl4.z : tl4.y, VREV ? tl4.y : tl4.z, VREV ? tl4.x : tl4.w};
    float part = 0.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = kb + r, l = lb + (VREV ? 3 - c : c);
        const float dt = tk[r] - tl[c];
        const float dk = kc.dell(dt, kc.val_fast(dt));
        part = fmaf((k < T && l < T && k != l) ? acc[r][c] : 0.0f, dk, part);  // branch-free
      }
    total += (double)part;
  }
  return total;
}

// Large-T contraction on the staged GEMM tile: 128 (k) x 64 (l) output blocks, contraction over the row index i
// of the two row-major lower triangles (both zero-filled above their diagonals while staging), kernel derivative
// in the epilogue.
// SYM: U == V, so (U^T V) is symmetric: only blocks that contain entries with k > l are computed, weighted twice.
template <int KERNEL, bool SYM>
__device__ __noinline__ double contract_gemm(const float* __restrict__ Ub, const float* __restrict__ Vb, const Lay& L, int T,
                                             const float* __restrict__ ts, float ell, float sig, float* __restrict__ stg,
                                             const float* __restrict__ kinv = nullptr, float hg = 0.0f) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ld = L.ld;
  const KernC<KERNEL> kc(ell, sig);
  const int nkb = (T + 127) / 128, nlb = (T + 63) / 64;
  double total = 0.0;
  for (int bp = 0; bp < nkb * nlb; ++bp) {
    const int kb0 = (bp % nkb) * 128, lb0 = (bp / nkb) * 64;
    if (SYM && lb0 >= kb0 + 127) continue;  // no entry with k > l in this block
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    const int i0 = (kb0 > lb0 ? kb0 : lb0) / GM_KC * GM_KC;  // X[i][k] = 0 for i < k
    gemm_tile_128x64<1, true, true>(acc, Ub + ld, ld, kb0, ld, Vb + ld, ld, lb0, ld, i0, T, stg);
    float tk[8], tl[4];
#pragma unroll
    for (int r = 0; r < 8; ++r) tk[r] = (kb0 + 8 * ty + r < T) ? ts[kb0 + 8 * ty + r] : 0.0f;
#pragma unroll
    for (int c = 0; c < 4; ++c) tl[c] = (lb0 + 4 * tx + c < T) ? ts[lb0 + 4 * tx + c] : 0.0f;
    if (!SYM && kinv) {  // prior-side term: acc += hg * K_p^-1 block (eight independent 16-byte loads, issued together)
      float4 kq[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int k = kb0 + 8 * ty + r, l = lb0 + 4 * tx;
        kq[r] = (k < L.TP && l < L.TP) ? __ldg(reinterpret_cast<const float4*>(kinv + (size_t)k * ld + l))
                                       : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        acc[r][0] = fmaf(hg, kq[r].x, acc[r][0]); acc[r][1] = fmaf(hg, kq[r].y, acc[r][1]);
        acc[r][2] = fmaf(hg, kq[r].z, acc[r][2]); acc[r][3] = fmaf(hg, kq[r].w, acc[r][3]);
      }
    }
    float part = 0.0f;
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = kb0 + 8 * ty + r, l = lb0 + 4 * tx + c;
        const float dt = tk[r] - tl[c];
        const float dk = kc.dell(dt, kc.val_fast(dt));
        const bool on = SYM ? (k < T && l < k) : (k < T && l < T && k != l);
        part = fmaf(on ? acc[r][c] : 0.0f, SYM ? 2.0f * dk : dk, part);
      }
    total += (double)part;
  }
  return total;
}

// Workspace variant of contract_block: 64x64 output blocks (one 4x4 tile per thread, 256 threads); the 64 U
// columns and 64 V columns of 32 rows at a time are staged with cp.async (zero-filled above the diagonal, so
// the other triangle of the buffer never leaks in) and consumed from shared memory.
template <int KERNEL>
__device__ __noinline__ double contract_block_staged(const float* __restrict__ Ub, const float* __restrict__ Vb,
                                                     const Lay& L, int T, const float* __restrict__ ts, float ell,
                                                     float sig, float* __restrict__ stg,
                                                     const float* __restrict__ kinv = nullptr, float hg = 0.0f) {
  constexpr int KC = 32, SLD = 128;
  const int tid = threadIdx.x;  // blockDim.x == 256
  const int ld = L.ld;
  const int nb = (T + 63) / 64;
  const KernC<KERNEL> kc(ell, sig);
  const int kt = tid & 15, lt = tid >> 4;
  double total = 0.0;
  for (int bp = 0; bp < nb * nb; ++bp) {
    const int kb0 = (bp % nb) * 64, lb0 = (bp / nb) * 64;
    const int kb = kb0 + 4 * kt, lb = lb0 + 4 * lt;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    const int i_begin = kb0 > lb0 ? kb0 : lb0;  // X[i][k] = 0 for i < k
    const int nch = (T - i_begin + KC - 1) / KC;
    auto issue = [&](int c) {
      float* buf = stg + (c & 1) * (KC * SLD);
      for (int q = tid; q < KC * 32; q += 256) {
        const int ii = q >> 5, seg = q & 31;  // seg 0..15: U columns, 16..31: V columns
        const int i = i_begin + c * KC + ii;
        const bool isU = seg < 16;
        const int col = (isU ? kb0 : lb0) + 4 * (seg & 15);
        const float* base = isU ? Ub : Vb;
        int valid = (i < T) ? (i - col + 1) : 0;  // columns col..col+3 exist while <= i
        valid = valid < 0 ? 0 : (valid > 4 ? 4 : valid);
        const float* src = valid ? base + (size_t)(i + 1) * ld + col : base;
        cp_async16(buf + ii * SLD + 4 * seg, src, 4 * valid);
      }
      cp_async_commit();
    };
    if (nch > 0) issue(0);
    for (int c = 0; c < nch; ++c) {
      if (c + 1 < nch) issue(c + 1);
      else cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
      const float* buf = stg + (c & 1) * (KC * SLD);
#pragma unroll 8
      for (int ii = 0; ii < KC; ++ii) {
        const float4 u4 = *reinterpret_cast<const float4*>(buf + ii * SLD + 4 * kt);
        const float4 v4 = *reinterpret_cast<const float4*>(buf + ii * SLD + 64 + 4 * lt);
        tile_fma<1>(acc, u4, v4);
      }
      __syncthreads();
    }
    float part = 0.0f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int k = kb + r, l = lb + c;
        const bool ok = k < T && l < T && k != l;
        const float dt = ok ? ts[k] - ts[l] : 0.0f;
        const float dk = kc.dell(dt, kc.val_fast(dt));
        float w = acc[r][c];
        if (kinv && ok) w = fmaf(hg, __ldg(kinv + (size_t)k * ld + l), w);
        part = fmaf(ok ? w : 0.0f, dk, part);
      }
    total += (double)part;
  }
  return total;
}


// ---- shared-prior fast path ------------------------------------------------------------------------------------
// The reference's prior length scales are one constant for all latent dims (prior_time_chars,
// Full_GP_VAE_dynamic_time.py:114): the D pairs of a sequence share K_p.  A pre-pass (prior_block, one CTA per
// SEQUENCE) factors it once and leaves a record in the workspace; the per-pair kernels then only factor K_q.
//   forward record  [0, TP*ld)            X_p = L_p^-1 as a full square, COLUMN-major (X(i,k) at rec[k*ld + i]), exact
//                                         zeros above the diagonal, identity on the padding
//                   [TP*ld, TP*ld + TP)   diag L_p
//   backward record [0, TP*ld)            K_p^-1 = X_p^T X_p, full symmetric square, row stride ld
// Forward: A = L_p^-1 L_q is the PRODUCT X_p L_q (no substitution chain, no barriers); the contraction index needs
// L_q row-major, which the factorisation leaves in the XR triangle (XRC).  Backward: alpha = K_p^-1 m, and the prior
// term g/2 <K_p^-1, dK_q/d ell> rides in the epilogue of the posterior contraction (kinv argument).
__host__ __device__ inline size_t rec_dg_offset(const Lay& L) { return (size_t)L.TP * L.ld; }

// sum_{c < i < T} A(i,c)^2, A = X L:  Xc column-major with explicit zeros above the diagonal (X(i,k) at Xc[k*ldx+i]),
// Lq's row-major copy in the XR triangle of Lb (L(k,c) at Lb[(k+1)*ld + c], garbage above the diagonal).  Thread partial.
template <bool SM = false>
__device__ __noinline__ float product_ssq_block(const float* __restrict__ Xc, int ldx, const float* __restrict__ Lb,
                                                const Lay& L, int T, Grp g) {
  if (SM) GPKL_SHARED_HINT(Lb);
  const int ld = L.ld;
  const int nt = (T + 3) / 4;
  const int ntri = nt * (nt + 1) / 2;
  float ssq = 0.0f;
  // lower tiles (it >= ct) enumerated by diagonal, longest contraction first: diagonal it - ct = nt-1-q holds q+1 tiles
  for (int rr = 0; rr * g.nt < ntri; ++rr) {
    const int n = dealt_index(rr, g);
    if (n >= ntri) continue;
    int q = (int)((sqrtf(8.0f * (float)n + 1.0f) - 1.0f) * 0.5f);
    while ((q + 1) * (q + 2) / 2 <= n) ++q;
    while (q * (q + 1) / 2 > n) --q;
    const int ct = n - q * (q + 1) / 2, it = ct + nt - 1 - q;
    const int rb = 4 * it, cb = 4 * ct;
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    // k in [cb, rb + 4): L(k,c) = 0 for k < c, X(i,k) = 0 for k > i.  Head steps touch L's diagonal: mask the garbage.
    int k = cb;
    const int kend = rb + 4;
    const int khead = (cb + 3 < kend) ? cb + 3 : kend;
    for (; k < khead; ++k) {
      const float4 u4 = *reinterpret_cast<const float4*>(Xc + (size_t)k * ldx + rb);
      const float4 v4 = *reinterpret_cast<const float4*>(Lb + (size_t)(k + 1) * ld + cb);
      const float u[4] = {u4.x, u4.y, u4.z, u4.w};
      const float v[4] = {cb <= k ? v4.x : 0.0f, cb + 1 <= k ? v4.y : 0.0f, cb + 2 <= k ? v4.z : 0.0f, cb + 3 <= k ? v4.w : 0.0f};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(u[r], v[c], acc[r][c]);
    }
    tile_update<1>(acc, Xc + rb, ldx, Lb + ld + cb, ld, k, kend);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = rb + r, cc = cb + c;
        const float v = (cc < i && i < T) ? acc[r][c] : 0.0f;
        ssq = fmaf(v, v, ssq);
      }
  }
  return ssq;
}

// The same on the staged GEMM tile (large T): 128 (i) x 64 (c) output blocks, A operand = the record (global, zeros
// above the diagonal), B operand = Lq's row-major copy in the workspace slot (zero-filled above its diagonal).
__device__ __noinline__ float product_ssq_gemm(const float* __restrict__ Xc, const float* __restrict__ Lb, const Lay& L, int T,
                                               float* __restrict__ stg) {
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int ld = L.ld;
  const int Tact = (T + NB - 1) / NB * NB;
  const int nib = (T + 127) / 128, ncb = (T + 63) / 64;
  float ssq = 0.0f;
  for (int bp = 0; bp < nib * ncb; ++bp) {
    const int i0 = (bp % nib) * 128, c0 = (bp / nib) * 64;
    if (c0 > i0 + 127) continue;  // entirely above the diagonal
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
    const int k0 = c0 / GM_KC * GM_KC;
    const int k1 = (i0 + 128 < Tact) ? i0 + 128 : Tact;
    gemm_tile_128x64<1, false, true>(acc, Xc, ld, i0, L.TP, Lb + ld, ld, c0, ld, k0, k1, stg);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = i0 + 8 * ty + r, cc = c0 + 4 * tx + c;
        const float v = (cc < i && i < T) ? acc[r][c] : 0.0f;
        ssq = fmaf(v, v, ssq);
      }
  }
  return ssq;
}

__device__ __forceinline__ void load_pair(const Params& P, int p, int b, int dd, int T, long long r0, const Lay& L,
                                          Sm& s, bool backward) {
  const GpklDesc& d = P.d;
  const int S = d.S, TP = L.TP;
  for (int i = threadIdx.x; i < L.ld; i += blockDim.x) s.ts[i] = (i < T) ? P.times[(size_t)b * d.T_max + i] : 0.0f;
  for (int i = threadIdx.x; i < TP; i += blockDim.x) {
    s.mm[i] = (i < T) ? P.mean[(size_t)(r0 + i) * d.D + dd] : 0.0f;
    if (backward) {
      float gs = 0.0f;
      for (int sx = 0; sx < S; ++sx) {
        const float gz = (i < T && P.g_z) ? P.g_z[((size_t)S * r0 + (size_t)sx * T + i) * d.D + dd] : 0.0f;
        s.u[(size_t)sx * TP + i] = gz;
        gs += gz;
      }
      s.gzs[i] = gs;
    }
    for (int sx = 0; sx < S; ++sx)
      s.v[(size_t)sx * TP + i] = (i < T) ? eps_value(P, ((size_t)p * S + sx) * d.T_max + i) : 0.0f;
  }
}

// Thread groups of a CTA.  With >= 128 threads and the full-GP posterior the CTA splits into two halves that
// run the independent prior chain (K_p) and posterior chain (K_q) concurrently on named barriers 1 / 2; with
// fewer threads (small T, many CTAs per SM) one group runs both chains back to back.
struct Groups {
  Grp all, chain;
  bool g0, g1, dual;
  __device__ Groups(bool want_dual) {
    const int tid = threadIdx.x, nt = blockDim.x, half = nt >> 1;
    dual = want_dual && nt >= 128;
    all = Grp{tid, nt, 0};
    if (dual) {
      g0 = tid < half;
      g1 = !g0;
      chain = Grp{g0 ? tid : tid - half, half, g0 ? 1 : 2};
    } else {
      g0 = g1 = true;
      chain = all;
    }
  }
};


// Pre-pass of the shared-prior path: one CTA per sequence (grid-stride).  Factors K_p(ell_p[0]) with the same code as
// the per-pair path, inverts the factor and writes the record (layout above).  WANT_KINV: backward record.
template <int KERNEL, bool SLOT, bool WANT_KINV>
__global__ void __launch_bounds__(256, SLOT ? 1 : 2) prior_block(Params P, int use_slot) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ int bad;
  griddep_launch_dependents();  // the per-pair kernel may start; it reads the records after its griddep_wait()
  if (*P.prior_flag == 0) return;  // ell_p differs between latent dims (offsets_kernel checked): per-pair path
  const GpklDesc& d = P.d;
  const Lay L(d.T_max, d.S);
  // use_slot: 0 = resident two-buffer carve-up, 1 = workspace slot, 2 = resident ONE-buffer carve-up (sizes whose two-buffer
  // layout does not fit shared memory but whose one-buffer shared-prior kernels do: 145 <= T <= 208)
  const bool onebuf = use_slot == 2;
  Sm s(smem_f, L, SLOT ? (use_slot == 1 ? P.scratch + (size_t)blockIdx.x * P.scratch_stride : nullptr) : nullptr, onebuf);
  float* const W = onebuf ? s.B2 : s.B1;  // the work matrix of the pre-pass
  const int TP = L.TP, ld = L.ld;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const float lp = P.ell_p[0];
  const Grp all{tid, nt, 0};
  const bool gm = SLOT && L.gemm(false);
  for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
    const int T = max(P.lengths[b], 0);
    const int Tact = (T + NB - 1) / NB * NB;
    float* __restrict__ rec = P.prior + (size_t)b * P.prior_stride;
    __syncthreads();
    if (tid == 0) bad = 0;
    for (int i = tid; i < ld; i += nt) s.ts[i] = (i < T) ? P.times[(size_t)b * d.T_max + i] : 0.0f;
    for (int i = tid; i < TP; i += nt) s.mm[i] = 0.0f;
    __syncthreads();
    if (T > 0) {
      if (gm) {
        chol_gemm<KERNEL>(W, L, T, false, s.ts, s.mm, lp, sig, noise, s.pan, s.wide, s.stg, s.dgp, s.rdp, &bad);
        (void)solve_gemm<true>(W, s.rdp, nullptr, W, L, T, s.pan, s.wide, s.stg);
      } else {
        chol_block<KERNEL, false, false, !SLOT>(W, L, T, false, s.ts, s.mm, lp, sig, noise, s.pan, s.dgp, s.rdp, &bad, all,
                                  (onebuf || (!use_slot && L.dual(true))) ? s.pan2 : nullptr);
        (void)solve_block<true, false, !SLOT>(W, s.rdp, nullptr, W, L, T, s.pan, all);
      }
      __syncthreads();
    }
    if (!WANT_KINV) {
      for (int idx = tid; idx < TP * ld; idx += nt) {
        const int k = idx / ld, i = idx - k * ld;
        float v;
        if (i < Tact && k < Tact) v = (i >= k) ? W[(size_t)(i + 1) * ld + k] : 0.0f;
        else v = (i == k) ? 1.0f : 0.0f;
        rec[idx] = v;
      }
      for (int i = tid; i < TP; i += nt) rec[rec_dg_offset(L) + i] = (i < Tact) ? s.dgp[i] : 1.0f;
    } else {
      // identity outside the factored range, then K_p^-1 = X^T X over it
      for (int idx = tid; idx < TP * ld; idx += nt) {
        const int k = idx / ld, l = idx - k * ld;
        if (k >= Tact || l >= Tact) rec[idx] = (k == l) ? 1.0f : 0.0f;
      }
      if (gm) {
        const int tx = tid & 15, ty = tid >> 4;
        const int nkb = (Tact + 127) / 128, nlb = (Tact + 63) / 64;
        for (int bp = 0; bp < nkb * nlb; ++bp) {
          const int kb0 = (bp % nkb) * 128, lb0 = (bp / nkb) * 64;
          float acc[8][4];
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
          const int i0 = (kb0 > lb0 ? kb0 : lb0) / GM_KC * GM_KC;
          gemm_tile_128x64<1, true, true>(acc, W + ld, ld, kb0, ld, W + ld, ld, lb0, ld, i0, Tact, s.stg);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const int k = kb0 + 8 * ty + r, l = lb0 + 4 * tx;
            if (k < Tact && l < Tact)
              *reinterpret_cast<float4*>(rec + (size_t)k * ld + l) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
          }
        }
      } else {
        const int nk = Tact / 4;
        for (int id = tid; id < nk * nk; id += nt) {
          const int lt = id % nk, kt = id / nk;
          const int kb = 4 * kt, lb = 4 * lt;
          float acc[4][4];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[r][c] = 0.0f;
          int i = kb > lb ? kb : lb;
          const int ihead = (i + 3 < Tact) ? i + 3 : Tact;
          for (; i < ihead; ++i) {
            const float4 u4 = *reinterpret_cast<const float4*>(W + (size_t)(i + 1) * ld + kb);
            const float4 v4 = *reinterpret_cast<const float4*>(W + (size_t)(i + 1) * ld + lb);
            const float u[4] = {kb <= i ? u4.x : 0.0f, kb + 1 <= i ? u4.y : 0.0f, kb + 2 <= i ? u4.z : 0.0f, kb + 3 <= i ? u4.w : 0.0f};
            const float v[4] = {lb <= i ? v4.x : 0.0f, lb + 1 <= i ? v4.y : 0.0f, lb + 2 <= i ? v4.z : 0.0f, lb + 3 <= i ? v4.w : 0.0f};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
              for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(u[r], v[c], acc[r][c]);
          }
          tile_update<1>(acc, W + ld + kb, ld, W + ld + lb, ld, i, Tact);
#pragma unroll
          for (int r = 0; r < 4; ++r)
            *reinterpret_cast<float4*>(rec + (size_t)(kb + r) * ld + lb) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        }
      }
    }
    if (tid == 0 && bad && P.status) atomicAdd(P.status, 1);
  }
}

// SH (resident sizes only): the shared-prior kernel proper.  It holds ONE work matrix in shared memory (two CTAs per
// SM up to T = 144) and reads the prior record from global memory; it returns at once when the device flag says the
// prior is not shared, and the ordinary kernel launched behind it (P.skip_if_shared) returns at once when it is.
// Off-diagonal part of  tr(K_p^-1 (K_q + m m^T))  against the float64 record (fwd_block, shared-prior one-buffer path):
// sum over i > j of Kinv(i,j) (K_q(i,j) + m_i m_j), K_q regenerated.  Columns j and T-1-j are paired (T-1 entries below the
// diagonal together), one pair per warp trip, lanes along the rows (coalesced record reads), UB row trips (compile time: no
// branch between them, so the kernel evaluations and the FP64 chains of a trip overlap).  All record entries of a column
// pair are in flight before their first use, and the next pair's are issued before this pair's arithmetic.
template <int KERNEL, int UB>
__device__ __forceinline__ void trace_offdiag(const double* __restrict__ kinv, int ldk, int T, const float* __restrict__ ts,
                                              const float* __restrict__ mm, const KernC<KERNEL>& kq, double& tro0, double& tro1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int npair = (T + 1) / 2;
  auto fetch = [&](int jj, double (&kd)[UB]) {
    const int j2 = T - 1 - jj, n1 = T - 1 - jj, n2 = j2 != jj ? jj : 0;
    const double* __restrict__ c1 = kinv + (size_t)jj * ldk + jj + 1;
    const double* __restrict__ c2 = kinv + (size_t)j2 * ldk + j2 + 1 - n1;
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int r = lane + 32 * u;
      kd[u] = (jj < npair && r < n1 + n2) ? __ldg((r < n1 ? c1 : c2) + r) : 0.0;
    }
  };
  double kd[UB], kn[UB];
  fetch(warp, kd);
  for (int jj = warp; jj < npair; jj += nw) {
    fetch(jj + nw, kn);
    const int j2 = T - 1 - jj, n1 = T - 1 - jj;
    const float t1 = ts[jj], t2 = ts[j2];
    const double m1 = (double)mm[jj], m2 = (double)mm[j2];
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int r = lane + 32 * u;
      const bool first = r < n1;
      int i = first ? jj + 1 + r : j2 + 1 + r - n1;
      i = i < T ? i : T - 1;  // lanes beyond the pair: kd = 0
      const double v = fma((double)mm[i], first ? m1 : m2, (double)kq.val(ts[i] - (first ? t1 : t2)));
      if (u & 1) tro1 = fma(kd[u], v, tro1);
      else tro0 = fma(kd[u], v, tro0);
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) kd[u] = kn[u];
  }
}

template <int KERNEL, int POST, bool DUAL, bool SLOT, bool SH = false>
__global__ void __launch_bounds__(256, SLOT ? 1 : 2) fwd_block(Params P, int use_slot) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ int bad;
  const GpklDesc& d = P.d;
  if (SH && *P.prior_flag == 0) return;
  if (!SH && P.skip_if_shared && *P.prior_flag != 0) return;
  const Lay L(d.T_max, d.S);
  // (resident instantiations pass a literal NULL slot: the work-matrix pointers are then provably shared-memory addresses)
  Sm s(smem_f, L, SLOT ? (use_slot ? P.scratch + (size_t)blockIdx.x * P.scratch_stride : nullptr) : nullptr, SH);
  const int S = d.S, TP = L.TP, ld = L.ld;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const Groups G(DUAL);
  // shared-prior fast path: the pre-pass found one ell_p for all latent dims and left per-sequence records
  const bool shared = SH || ((POST == GPKL_POST_GP) && !DUAL && SLOT && P.prior != nullptr && *P.prior_flag != 0);
  for (int p = blockIdx.x; p < d.B * d.D; p += gridDim.x) {
    const int b = p / d.D, dd = p - b * d.D;
    const int T = P.lengths[b];
    const long long r0 = P.offsets[b];
    __syncthreads();
    if (T <= 0) {
      if (threadIdx.x == 0) {
        P.kl_pairs[p] = 0.0f;
        if (P.logdets) { P.logdets[2 * p] = 0.0f; P.logdets[2 * p + 1] = 0.0f; }
      }
      continue;
    }
    if (threadIdx.x == 0) bad = 0;
    phase_mark(P, 0);
    load_pair(P, p, b, dd, T, r0, L, s, false);
    __syncthreads();
    phase_mark(P, 1);
    const bool gm = SLOT && L.gemm(false);  // large T: GEMM-structured phases on shared-memory panels
    if (!shared) {
      if (gm) chol_gemm<KERNEL>(s.B1, L, T, true, s.ts, s.mm, P.ell_p[dd], sig, noise, s.pan, s.wide, s.stg, s.dgp, s.rdp, &bad);
      else if (G.g0) chol_block<KERNEL, DUAL, false, !SLOT>(s.B1, L, T, true, s.ts, s.mm, P.ell_p[dd], sig, noise, s.pan, s.dgp, s.rdp, &bad, G.chain);
    }
    phase_mark(P, 2);
    double part = 0.0, ldp = 0.0, ldq = 0.0;
    if (POST == GPKL_POST_GP && shared) {
      // ---- shared-prior path: only K_q is factored here (all threads); L_p^-1 and diag L_p come from the record
      if (gm) chol_gemm<KERNEL, true>(s.B2, L, T, false, s.ts, s.mm, P.ell_q[dd], sig, noise, s.pan, s.wide, s.stg, s.dgq, s.rdq, &bad);
      else if (SH && (use_slot & 4))  // (float64-record path: no product, so no row-major copy of the factor either)
        chol_block<KERNEL, false, false, true>(s.B2, L, T, false, s.ts, s.mm, P.ell_q[dd], sig, noise, s.pan, s.dgq, s.rdq, &bad, G.all, s.pan2);
      else chol_block<KERNEL, false, true, !SLOT>(s.B2, L, T, false, s.ts, s.mm, P.ell_q[dd], sig, noise, s.pan, s.dgq, s.rdq, &bad, G.all,
                                           SH ? s.pan2 : nullptr);
      phase_mark(P, 3);
      for (int i = threadIdx.x; i < T; i += blockDim.x) {  // z_s = m + L_q eps_s (loads batched 8 ahead: L_q may be global)
        for (int sx = 0; sx < S; ++sx) {
          const float* ev = s.v + (size_t)sx * TP;
          float acc = s.mm[i], acc1 = 0.0f;
          int k = 0;
          for (; k + 8 <= i + 1; k += 8) {
            float lv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) lv[e] = s.B2[(size_t)(k + e) * ld + i];
#pragma unroll
            for (int e = 0; e < 8; e += 2) { acc = fmaf(lv[e], ev[k + e], acc); acc1 = fmaf(lv[e + 1], ev[k + e + 1], acc1); }
          }
          for (; k <= i; ++k) acc = fmaf(s.B2[(size_t)k * ld + i], ev[k], acc);
          P.z[((size_t)S * r0 + (size_t)sx * T + i) * d.D + dd] = acc + acc1;
        }
      }
      phase_mark(P, 4);
      griddep_wait();  // the record is first needed here: the pre-pass overlaps the K_q factorisation
      const float* __restrict__ rec = P.prior + (size_t)b * P.prior_stride;
      if (SH && (use_slot & 4)) {
        // ---- float64 record (gpkl_prior64.cu): K_p^-1, lower triangle column-major with pitch ldk, log|K_p| behind it.
        // The reference's own formula (Full_GP_VAE_dynamic_time.py:250-259)
        //     KL = 1/2 [ tr(K_p^-1 (K_q + m m^T)) - T + log|K_p| - log|K_q| ],   tr = sum_ij Kinv_ij (K_q,ij + m_i m_j)
        // with the K_q entries REGENERATED (T^2/2 kernel evaluations: nothing next to the factorisation): the T^3/3-flop
        // triangular product A = L_p^-1 L_q and a = L_p^-1 m are gone.  Columns j and T-1-j are paired (T-1 entries below
        // the diagonal together), one pair per warp trip, lanes along the rows: coalesced record reads.
        phase_mark(P, 7);
        const int ldk = (d.T_max + 63) / 64 * 64;
        const double* __restrict__ kinv = reinterpret_cast<const double*>(rec);
        const KernC<KERNEL> kq(P.ell_q[dd], sig);
        double tro0 = 0.0, tro1 = 0.0, trd = 0.0;
        switch ((d.T_max + 30) / 32) {  // row trips per column pair (grid-uniform; T_max <= 208 for the one-buffer sizes)
          case 1: case 2: trace_offdiag<KERNEL, 2>(kinv, ldk, T, s.ts, s.mm, kq, tro0, tro1); break;
          case 3: trace_offdiag<KERNEL, 3>(kinv, ldk, T, s.ts, s.mm, kq, tro0, tro1); break;
          case 4: trace_offdiag<KERNEL, 4>(kinv, ldk, T, s.ts, s.mm, kq, tro0, tro1); break;
          case 5: trace_offdiag<KERNEL, 5>(kinv, ldk, T, s.ts, s.mm, kq, tro0, tro1); break;
          case 6: trace_offdiag<KERNEL, 6>(kinv, ldk, T, s.ts, s.mm, kq, tro0, tro1); break;
          default: trace_offdiag<KERNEL, 7>(kinv, ldk, T, s.ts, s.mm, kq, tro0, tro1); break;
        }
        const float kdiag = kq.val(0.0f) + noise;
        for (int i = threadIdx.x; i < T; i += blockDim.x) {
          const double mi = (double)s.mm[i];
          trd = fma(__ldg(kinv + (size_t)i * ldk + i), fma(mi, mi, (double)kdiag), trd);
          ldq += 2.0 * log((double)s.dgq[i]);
        }
        part = trd + 2.0 * (tro0 + tro1) - ldq;
        if (threadIdx.x == 0) {
          ldp = __ldg(kinv + (size_t)ldk * ldk);
          part += ldp - (double)T;
        }
        phase_mark(P, 8);
        phase_mark(P, 5);
      } else {
      for (int i = threadIdx.x; i < TP; i += blockDim.x) s.dgp[i] = __ldg(rec + rec_dg_offset(L) + i);
      phase_mark(P, 7);
      float ssq;
      if (gm) {
        for (int i = threadIdx.x; i < T; i += blockDim.x) {  // a = L_p^-1 m from the record (coalesced over i)
          float a0 = 0.0f, a1 = 0.0f;
          int k = 0;
          for (; k + 8 <= i + 1; k += 8) {
            float xv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) xv[e] = __ldg(rec + (size_t)(k + e) * ld + i);
#pragma unroll
            for (int e = 0; e < 8; e += 2) { a0 = fmaf(xv[e], s.mm[k + e], a0); a1 = fmaf(xv[e + 1], s.mm[k + e + 1], a1); }
          }
          for (; k <= i; ++k) a0 = fmaf(__ldg(rec + (size_t)k * ld + i), s.mm[k], a0);
          s.aa[i] = a0 + a1;
        }
        ssq = product_ssq_gemm(rec, s.B2, L, T, s.stg);
      } else {
        // the record (this sequence's L_p^-1, shared by its D pairs) is read in place: L1/L2 resident, no staging copy
        for (int i = threadIdx.x; i < T; i += blockDim.x) {
          float a0 = 0.0f, a1 = 0.0f;
          int k = 0;
          for (; k + 8 <= i + 1; k += 8) {
            float xv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) xv[e] = __ldg(rec + (size_t)(k + e) * ld + i);
#pragma unroll
            for (int e = 0; e < 8; e += 2) { a0 = fmaf(xv[e], s.mm[k + e], a0); a1 = fmaf(xv[e + 1], s.mm[k + e + 1], a1); }
          }
          for (; k <= i; ++k) a0 = fmaf(__ldg(rec + (size_t)k * ld + i), s.mm[k], a0);
          s.aa[i] = a0 + a1;
        }
        phase_mark(P, 8);
        ssq = product_ssq_block<!SLOT>(rec, ld, s.B2, L, T, G.all);
      }
      __syncthreads();
      phase_mark(P, 5);
      part = (double)ssq;
      for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const double lpd = (double)s.dgp[i], lqd = (double)s.dgq[i];
        const double av = (double)s.aa[i];
        part += diag_term(lqd / lpd) + av * av;
        if (P.logdets) {  // float64 logs only when the log-determinants are requested (they are not part of the KL form used)
          ldp += 2.0 * log(lpd);
          ldq += 2.0 * log(lqd);
        }
      }
      }
    } else if (POST == GPKL_POST_GP) {
      if (gm) chol_gemm<KERNEL>(s.B2, L, T, false, s.ts, s.mm, P.ell_q[dd], sig, noise, s.pan, s.wide, s.stg, s.dgq, s.rdq, &bad);
      if (G.g1) {
        if (!gm)
          chol_block<KERNEL, DUAL, false, !SLOT>(s.B2, L, T, false, s.ts, s.mm, P.ell_q[dd], sig, noise, G.dual ? s.pan2 : s.pan, s.dgq,
                                   s.rdq, &bad, G.chain);
        for (int i = G.chain.tid; i < T; i += G.chain.nt) {  // z_s = m + L_q eps_s
          for (int sx = 0; sx < S; ++sx) {
            const float* ev = s.v + (size_t)sx * TP;
            float acc = s.mm[i];
            for (int k = 0; k <= i; ++k) acc = fmaf(s.B2[(size_t)k * ld + i], ev[k], acc);
            P.z[((size_t)S * r0 + (size_t)sx * T + i) * d.D + dd] = acc;
          }
        }
      }
      if (G.dual) __syncthreads();  // join the two chains
      phase_mark(P, 4);
      const float ssq = gm ? solve_gemm<false>(s.B1, s.rdp, s.B2, s.B1, L, T, s.pan, s.wide, s.stg)
                           : solve_block<false, DUAL, !SLOT>(s.B1, s.rdp, s.B2, s.B1, L, T, s.pan, G.all);
      phase_mark(P, 5);
      part = (double)ssq;
      for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const double lpd = (double)s.dgp[i], lqd = (double)s.dgq[i];
        const double av = (double)s.B1[(size_t)i * ld + TP];  // a_i = (L_p^-1 m)_i rides in the extra row
        part += diag_term(lqd / lpd) + av * av;
        if (P.logdets) {  // float64 logs only when the log-determinants are requested (they are not part of the KL form used)
          ldp += 2.0 * log(lpd);
          ldq += 2.0 * log(lqd);
        }
      }
    } else {
      if (gm) (void)solve_gemm<true>(s.B1, s.rdp, nullptr, s.B1, L, T, s.pan, s.wide, s.stg);
      else (void)solve_block<true, DUAL, !SLOT>(s.B1, s.rdp, nullptr, s.B1, L, T, s.pan, G.all);
      for (int i = threadIdx.x; i < T; i += blockDim.x) {
        float h = 0.0f;
        for (int k = i; k < T; ++k) { const float x = s.B1[(size_t)(k + 1) * ld + i]; h = fmaf(x, x, h); }
        const float lv = P.aux[(size_t)(r0 + i) * d.D + dd];
        const float vv = expf(lv), sd = expf(0.5f * lv);
        const double lpd = (double)s.dgp[i], av = (double)s.B1[(size_t)i * ld + TP];
        part += (double)h * (double)vv - 1.0 - (double)lv + av * av + 2.0 * log(lpd);
        ldp += 2.0 * log(lpd);
        ldq += (double)lv;
        for (int sx = 0; sx < S; ++sx)
          P.z[((size_t)S * r0 + (size_t)sx * T + i) * d.D + dd] = s.mm[i] + sd * s.v[(size_t)sx * TP + i];
      }
    }
    part = block_sum(part, s.red);
    if (P.logdets) {
      ldp = block_sum(ldp, s.red);
      ldq = block_sum(ldq, s.red);
    }
    phase_mark(P, 6);
#ifdef GPKL_PANEL_TRACE
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0)
      for (int i = 0; i < 8; ++i) { P.dbg[32 + i] = g_ptrace[i]; g_ptrace[i] = 0; }
#endif
    if (threadIdx.x == 0) {
      P.kl_pairs[p] = (float)(0.5 * part);
      if (P.logdets) { P.logdets[2 * p] = (float)ldp; P.logdets[2 * p + 1] = (float)ldq; }
      if (bad && P.status) atomicAdd(P.status, 1);
    }
  }
}

template <int KERNEL, int POST, bool DUAL, bool SLOT, bool SH = false>
__global__ void __launch_bounds__(256, SLOT ? 1 : 2) bwd_block(Params P, int use_slot) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ int bad;
  const GpklDesc& d = P.d;
  if (SH && *P.prior_flag == 0) return;
  if (!SH && P.skip_if_shared && *P.prior_flag != 0) return;
  const Lay L(d.T_max, d.S);
  // (resident instantiations pass a literal NULL slot: the work-matrix pointers are then provably shared-memory addresses)
  Sm s(smem_f, L, SLOT ? (use_slot ? P.scratch + (size_t)blockIdx.x * P.scratch_stride : nullptr) : nullptr, SH);
  const int S = d.S, TP = L.TP, ld = L.ld;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const double g_sum = P.g_kl_sum ? *P.g_kl_sum : 1.0;
  const Groups G(DUAL);
  // shared-prior fast path: the pre-pass found one ell_p for all latent dims and left per-sequence records
  const bool shared = SH || ((POST == GPKL_POST_GP) && !DUAL && SLOT && P.prior != nullptr && *P.prior_flag != 0);
  for (int p = blockIdx.x; p < d.B * d.D; p += gridDim.x) {
    const int b = p / d.D, dd = p - b * d.D;
    const int T = P.lengths[b];
    const long long r0 = P.offsets[b];
    __syncthreads();
    if (T <= 0) {
      if (threadIdx.x == 0 && P.gq_pairs) P.gq_pairs[p] = 0.0f;
      continue;
    }
    const float g = (float)(g_sum + (P.g_kl_pairs ? (double)P.g_kl_pairs[p] : 0.0));
    if (threadIdx.x == 0) bad = 0;
    phase_mark(P, 16);
    load_pair(P, p, b, dd, T, r0, L, s, true);
    __syncthreads();
    phase_mark(P, 17);
    const float lp = P.ell_p[dd];
    const float lq = (POST == GPKL_POST_GP) ? P.ell_q[dd] : lp;
    const bool gm = SLOT && L.gemm(false);  // large T: GEMM-structured phases on shared-memory panels
    double t1 = 0.0;
    if (!shared && G.g0) {  // ---- prior chain: L_p, X_p = L_p^-1, alpha = K_p^-1 m, t1 = <K_p^-1, dK_q/d ell>
      if (gm) chol_gemm<KERNEL>(s.B1, L, T, true, s.ts, s.mm, lp, sig, noise, s.pan, s.wide, s.stg, s.dgp, s.rdp, &bad);
      else chol_block<KERNEL, DUAL, false, !SLOT>(s.B1, L, T, true, s.ts, s.mm, lp, sig, noise, s.pan, s.dgp, s.rdp, &bad, G.chain);
      phase_mark(P, 18);
      if (gm) (void)solve_gemm<true>(s.B1, s.rdp, nullptr, s.B1, L, T, s.pan, s.wide, s.stg);
      else (void)solve_block<true, DUAL, !SLOT>(s.B1, s.rdp, nullptr, s.B1, L, T, s.pan, G.chain);
      phase_mark(P, 19);
      // a = L_p^-1 m (the extra row of the factor) into shared memory once, then alpha = X_p^T a with the
      // column loads issued 8 rows ahead
      for (int i = G.chain.tid; i < T; i += G.chain.nt) s.aa[i] = s.B1[(size_t)i * ld + TP];
      grp_sync<DUAL>(G.chain);
      for (int k = G.chain.tid; k < T; k += G.chain.nt) {
        const float* __restrict__ xp = s.B1 + ld + k;  // X_p(i, k) at xp[i*ld]
        float al = 0.0f;
        for (int i0 = k; i0 < T; i0 += 8) {
          float xv[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) xv[e] = (i0 + e < T) ? xp[(size_t)(i0 + e) * ld] : 0.0f;
#pragma unroll
          for (int e = 0; e < 8; ++e) al = fmaf(xv[e], (i0 + e < T) ? s.aa[i0 + e] : 0.0f, al);
        }
        P.g_mean[(size_t)(r0 + k) * d.D + dd] = g * al + s.gzs[k];
      }
      phase_mark(P, 20);
      if (POST == GPKL_POST_GP)
        t1 = gm ? contract_gemm<KERNEL, true>(s.B1, s.B1, L, T, s.ts, lq, sig, s.stg)
           : SLOT ? contract_block_staged<KERNEL>(s.B1, s.B1, L, T, s.ts, lq, sig, s.stg)
                      : contract_block<KERNEL, false, !SLOT>(s.B1, s.B1, L, T, s.ts, lq, sig, G.chain);
      phase_mark(P, 21);
    }
    if (POST == GPKL_POST_DIAG) {
      for (int i = threadIdx.x; i < T; i += blockDim.x) {
        float h = 0.0f;
        for (int k = i; k < T; ++k) { const float x = s.B1[(size_t)(k + 1) * ld + i]; h = fmaf(x, x, h); }
        const float lv = P.aux[(size_t)(r0 + i) * d.D + dd];
        const float vv = expf(lv), sd = expf(0.5f * lv);
        float ge = 0.0f;
        for (int sx = 0; sx < S; ++sx) ge = fmaf(s.u[(size_t)sx * TP + i], s.v[(size_t)sx * TP + i], ge);
        P.g_aux[(size_t)(r0 + i) * d.D + dd] = 0.5f * g * (h * vv - 1.0f) + 0.5f * sd * ge;
      }
    } else {
      if (G.g1) {  // ---- posterior chain: L_q, w = L_q^T g_z, X_q = L_q^-1
        if (gm) chol_gemm<KERNEL>(s.B2, L, T, false, s.ts, s.mm, lq, sig, noise, s.pan, s.wide, s.stg, s.dgq, s.rdq, &bad);
        else
          chol_block<KERNEL, DUAL, false, !SLOT>(s.B2, L, T, false, s.ts, s.mm, lq, sig, noise, G.dual ? s.pan2 : s.pan, s.dgq, s.rdq,
                                   &bad, G.chain, SH ? s.pan2 : nullptr);
        phase_mark(P, 22);
        if (SLOT) {
          // L_q lives in global memory: one WARP per column k, lanes stride down the column (coalesced), shuffle-reduce
          const int lane = G.chain.tid & 31, wid = G.chain.tid >> 5, nw = G.chain.nt >> 5;
          for (int k = wid; k < T; k += nw) {
            float pdk = 0.0f;
            for (int sx = 0; sx < S; ++sx) {
              const float* uu = s.u + (size_t)sx * TP;
              const float* __restrict__ colk = s.B2 + (size_t)k * ld;
              float wk = 0.0f;
              for (int i = k + lane; i < T; i += 32) wk = fmaf(colk[i], uu[i], wk);
              wk = warp_sum(wk);
              if (lane == 0) s.w[(size_t)sx * TP + k] = wk;
              pdk = fmaf(wk, s.v[(size_t)sx * TP + k], pdk);
            }
            if (lane == 0) s.pd[k] = 0.5f * pdk - 0.5f * g;
          }
        } else {
          for (int k = G.chain.tid; k < T; k += G.chain.nt) {
            float pdk = 0.0f;
            for (int sx = 0; sx < S; ++sx) {
              const float* uu = s.u + (size_t)sx * TP;
              float wk = 0.0f;
              for (int i = k; i < T; ++i) wk = fmaf(s.B2[(size_t)k * ld + i], uu[i], wk);
              s.w[(size_t)sx * TP + k] = wk;
              pdk = fmaf(wk, s.v[(size_t)sx * TP + k], pdk);
            }
            s.pd[k] = 0.5f * pdk - 0.5f * g;
          }
        }
        phase_mark(P, 23);
        if (gm) (void)solve_gemm<true>(s.B2, s.rdq, nullptr, s.B2, L, T, s.pan, s.wide, s.stg);
        else (void)solve_block<true, DUAL, !SLOT>(s.B2, s.rdq, nullptr, s.B2, L, T, G.dual ? s.pan2 : s.pan, G.chain);
        phase_mark(P, 24);
      }
      __syncthreads();  // join: X_p (dead after t1), X_q, w, pd are complete
      phase_mark(P, 28);
      // C' = (Phi(sum_s w_s eps_s^T) - g/2 I) X_q by running prefix sums down each column, into XR1.  Loads are
      // issued 8 rows ahead of the dependent prefix arithmetic (the matrices may live in global memory).
      for (int l = threadIdx.x; l < T; l += blockDim.x) {
        const float* __restrict__ xq = s.B2 + ld + l;  // X_q(i, l) at xq[i*ld]
        // C'(i, l) at cp[i*cs]: the XR triangle of B1, or (SH, one buffer) row/column-reversed in the triangle of B2
        // that L_q vacated once w and X_q were formed
        float* __restrict__ cp = SH ? s.B2 + (size_t)(TP - 1) * ld + (TP - 1 - l) : s.B1 + ld + l;
        const ptrdiff_t cs = SH ? -(ptrdiff_t)ld : (ptrdiff_t)ld;
        if (S == 1) {
          // one sample (the reference's default, :318): every operand of the 8-row batch is loaded before the dependent
          // prefix chain (the per-row loads of pd / w / eps inside the chain cost a shared-memory round trip per row)
          float cum = 0.0f;
          for (int i0 = l; i0 < T; i0 += 8) {
            float xv[8], pv[8], wv[8], ev[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int i = i0 + e;
              const bool ok = i < T;
              const int ii = ok ? i : l;
              xv[e] = ok ? xq[(size_t)i * ld] : 0.0f;
              pv[e] = s.pd[ii];
              wv[e] = s.w[ii];
              ev[e] = s.v[ii];
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int i = i0 + e;
              const float cv = fmaf(wv[e], cum, pv[e] * xv[e]);
              cum = fmaf(ev[e], xv[e], cum);
              if (i < T) cp[(ptrdiff_t)i * cs] = cv;
            }
          }
          continue;
        }
        for (int sx0 = 0; sx0 < S; sx0 += 4) {
          const int ns = (S - sx0 < 4) ? S - sx0 : 4;
          float cum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
          for (int i0 = l; i0 < T; i0 += 8) {
            float xv[8], base[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int i = i0 + e;
              xv[e] = (i < T) ? xq[(size_t)i * ld] : 0.0f;
              base[e] = (sx0 == 0) ? 0.0f : ((i < T) ? cp[(ptrdiff_t)i * cs] : 0.0f);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int i = i0 + e;
              if (i < T) {
                float cv = (sx0 == 0) ? s.pd[i] * xv[e] : base[e];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  if (q < ns) {
                    cv = fmaf(s.w[(size_t)(sx0 + q) * TP + i], cum[q], cv);
                    cum[q] = fmaf(s.v[(size_t)(sx0 + q) * TP + i], xv[e], cum[q]);
                  }
                }
                cp[(ptrdiff_t)i * cs] = cv;
              }
            }
          }
        }
      }
      phase_mark(P, 29);
      __syncthreads();
      phase_mark(P, 25);
      // shared-prior path: alpha = K_p^-1 m from the record (K_p^-1 symmetric: coalesced over k), and the prior term
      // g/2 <K_p^-1, dK_q/d ell> rides in the epilogue of the contraction below (t1 stays 0)
      const float* __restrict__ kinv = nullptr;
      if (shared) {
        griddep_wait();  // the record is first needed here: the pre-pass overlaps the whole K_q chain
        kinv = P.prior + (size_t)b * P.prior_stride;
        for (int k = threadIdx.x; k < T; k += blockDim.x) {
          float a0 = 0.0f, a1 = 0.0f;
          int l = 0;
          for (; l + 8 <= T; l += 8) {
            float xv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) xv[e] = __ldg(kinv + (size_t)(l + e) * ld + k);
#pragma unroll
            for (int e = 0; e < 8; e += 2) { a0 = fmaf(xv[e], s.mm[l + e], a0); a1 = fmaf(xv[e + 1], s.mm[l + e + 1], a1); }
          }
          for (; l < T; ++l) a0 = fmaf(__ldg(kinv + (size_t)l * ld + k), s.mm[l], a0);
          P.g_mean[(size_t)(r0 + k) * d.D + dd] = g * (a0 + a1) + s.gzs[k];
        }
      }
      const float hg = 0.5f * g;
      const double t2 = gm ? contract_gemm<KERNEL, false>(s.B2, s.B1, L, T, s.ts, lq, sig, s.stg, kinv, hg)
                        : SLOT ? contract_block_staged<KERNEL>(s.B2, s.B1, L, T, s.ts, lq, sig, s.stg, kinv, hg)
                        : SH ? contract_block<KERNEL, true, true>(s.B2, s.B2, L, T, s.ts, lq, sig, G.all, kinv, hg)
                             : contract_block<KERNEL, false, !SLOT>(s.B2, s.B1, L, T, s.ts, lq, sig, G.all, kinv, hg);
      phase_mark(P, 26);
      const double gq = block_sum(0.5 * (double)g * t1 + t2, s.red);
      phase_mark(P, 27);
      if (threadIdx.x == 0) P.gq_pairs[p] = (float)gq;
    }
#ifdef GPKL_PANEL_TRACE
    if (P.dbg && blockIdx.x == 0 && threadIdx.x == 0)
      for (int i = 0; i < 8; ++i) { P.dbg[40 + i] = g_ptrace[i]; g_ptrace[i] = 0; }
#endif
    if (threadIdx.x == 0 && bad && P.status) atomicAdd(P.status, 1);
  }
}

template <int KERNEL, int POST>
cudaError_t launch_kp(const Params& P_in, bool backward, cudaStream_t st) {
  Params P = P_in;
  const Lay L(P.d.T_max, P.d.S);
  const bool resident = block_tier_resident(P.d);
  const size_t smem = L.floats(resident) * sizeof(float);
  if (smem > kMaxDynSmem) return cudaErrorInvalidValue;
  if (!resident && !P.scratch) return cudaErrorInvalidValue;
  const int npairs = P.d.B * P.d.D;
  const int nt = P.d.T_max <= 64 ? 64 : 256;
  int grid;
  if (resident) {
    int per_sm = (int)(kMaxDynSmem / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2048 / nt) per_sm = 2048 / nt;
    const int cap = kNumSMs * per_sm * 4;
    grid = npairs < cap ? npairs : cap;
  } else {
    grid = npairs < kBlockSlots ? npairs : kBlockSlots;  // one workspace slot per CTA
  }
  // Shared-prior path (records in the workspace): served for the resident and GEMM-path sizes.
  //  * GEMM path: one kernel decides on the device (flag) between the shared and the per-pair prior.
  //  * resident sizes: the shared-prior kernel proper (SH: one work matrix in shared memory, two CTAs per SM up to
  //    T = 144) is launched behind the pre-pass and returns at once if the device flag says "not shared"; the ordinary
  //    per-pair kernel is launched behind it and returns at once if the flag says "shared" (no host read of ell_p).
  const bool share = POST == GPKL_POST_GP && P_in.prior != nullptr && (resident || L.gemm(false));
  if (!share) { P.prior = nullptr; P.prior_flag = nullptr; }
  // one-buffer residency: the shared-prior kernels of sizes whose TWO work matrices do not fit shared memory but whose
  // ONE does (145 <= T <= 208) still run shared-memory resident (one CTA per SM); only their per-pair fallback uses the
  // workspace path.  GPKL_ONEBUF_TMAX (experiments) lowers the limit; 0 switches the extension off.
  static const int env_tmax = [] { const char* e = getenv("GPKL_ONEBUF_TMAX"); return e ? atoi(e) : 1 << 30; }();
  const size_t smem1 = L.floats(true, true) * sizeof(float);
  const bool sh_resident = share && nt == 256 ? (resident || (smem1 <= kMaxDynSmem && L.TP <= env_tmax)) : (share && resident);
  const bool tile = share && !resident && !sh_resident && tile_tier_supports(P.d, backward);
  const bool dual = POST == GPKL_POST_GP && L.dual(resident);
  void (*kern)(Params, int);
  if (!resident) kern = backward ? bwd_block<KERNEL, POST, false, true> : fwd_block<KERNEL, POST, false, true>;
  else if (!backward) kern = dual ? fwd_block<KERNEL, POST, true, false> : fwd_block<KERNEL, POST, false, false>;
  else kern = dual ? bwd_block<KERNEL, POST, true, false> : bwd_block<KERNEL, POST, false, false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  // (the profiling events bracket pre-pass + per-pair kernel: an event between them would break the programmatic
  //  dependency and serialise them)
  prof_begin(backward, st);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(nt);
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 0;
  if (share && !(tile && !backward)) {  // (the tile tier's forward pass has a pre-pass of its own: float64 K_p^-1, gpkl_prior64.cu)
    void (*pk)(Params, int);
    if (!sh_resident) pk = backward ? prior_block<KERNEL, true, true> : prior_block<KERNEL, true, false>;
    else pk = backward ? prior_block<KERNEL, false, true> : prior_block<KERNEL, false, false>;
    // pre-pass carve-up: two buffers where they fit (its own look-ahead panel included), else the one-buffer layout
    const size_t psmem = !sh_resident ? smem : (resident ? smem : smem1);
    e = cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    if (e != cudaSuccess) return e;
    const int pgrid = sh_resident ? P.d.B : (P.d.B < kBlockSlots ? P.d.B : kBlockSlots);
    // forward of the one-buffer kernels (T > 64): float64 K_p^-1 records (gpkl_prior64.cu) and the entrywise trace
    const bool rec64 = sh_resident && !backward && nt == 256 && POST == GPKL_POST_GP;
    if (rec64) {
      e = launch_prior_inv64(P, st);
      if (e != cudaSuccess) return e;
    } else {
      pk<<<pgrid, 256, psmem, st>>>(P, !sh_resident ? 1 : (resident ? 0 : 2));
      note_launch();
    }
    if (sh_resident) {
      if (POST == GPKL_POST_GP) {  // (the SH instantiations exist for the GP posterior only)
        void (*sk)(Params, int) = backward ? bwd_block<KERNEL, GPKL_POST_GP, false, false, true>
                                           : fwd_block<KERNEL, GPKL_POST_GP, false, false, true>;
        e = cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) return e;
        static const int env_nt = [] { const char* e = getenv("GPKL_SH_NT"); return e ? atoi(e) : 0; }();
        // CTA size of the one-buffer kernels: what counts is pairs in flight per SM (latency-bound phases), so fewer
        // threads per pair where shared memory admits a third CTA (measured at the end of session 3, fwd | bwd ms at
        // 128 / 192 / 256 threads: T=96 2.33|4.06, 2.75|4.53, 2.88|4.80; T=128 2.61|4.32, 2.43|4.06, 2.55|4.13;
        // T=160 (one CTA per SM) 3.84|6.30, 3.31|5.80, 3.17|5.59); GPKL_SH_NT overrides (experiments)
        const int def_nt = nt != 256 ? nt : (L.TP <= 112 ? 128 : (resident ? 192 : 256));
        const int snt = (env_nt >= 32 && env_nt <= 256 && env_nt % 32 == 0) ? env_nt : def_nt;
        int per_sm = (int)(kMaxDynSmem / (smem1 + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2048 / snt) per_sm = 2048 / snt;
        if (per_sm > 65536 / (snt * 128)) per_sm = 65536 / (snt * 128);
        const int cap = kNumSMs * per_sm * 4;
        cfg.gridDim = dim3(npairs < cap ? npairs : cap);
        cfg.blockDim = dim3(snt);
        cfg.dynamicSmemBytes = smem1;
        // the pre-pass overlaps the K_q chain of the per-pair kernel (programmatic dependent launch)
        cfg.numAttrs = pdl_enabled() ? 1 : 0;
        e = cudaLaunchKernelEx(&cfg, sk, P, rec64 ? 4 : 0);
        note_launch();
        if (e != cudaSuccess) return e;
      }
      // per-pair fallback behind it: no records, returns at once when the flag says "shared"
      P.prior = nullptr;
      P.skip_if_shared = 1;
    }
  }
  if (tile) {
    // tile tier (gpkl_tile.cu): the shared-prior kernel of 208 < T <= 512; it returns at once when the device flag says
    // "not shared", and the per-pair kernel launched behind it returns at once when it says "shared"
    e = launch_tile(P, backward, st);
    if (e != cudaSuccess) return e;
    P.prior = nullptr;
    P.skip_if_shared = 1;
  }
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(nt);
  cfg.dynamicSmemBytes = smem;
  cfg.numAttrs = 0;
  e = cudaLaunchKernelEx(&cfg, kern, P, resident ? 0 : 1);
  prof_end(backward, st);
  note_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace

bool block_tier_resident(const GpklDesc& d) {
  const Lay L(d.T_max, d.S);
  return L.floats(true) * sizeof(float) <= kMaxDynSmem;
}

size_t block_slot_floats(const GpklDesc& d) {
  const Lay L(d.T_max, d.S);
  return block_tier_resident(d) ? 0 : 2 * L.buf();
}

bool block_tier_supports(const GpklDesc& d, bool backward) {
  if (d.T_max < 1) return false;
  if (backward && (d.flags & GPKL_FLAG_GRAD_ELL_P)) return false;  // d/d ell_p is served by the generic tier
  if (d.posterior != GPKL_POST_GP && d.posterior != GPKL_POST_DIAG) return false;
  const Lay L(d.T_max, d.S);
  return L.floats(false) * sizeof(float) <= kMaxDynSmem;
}

cudaError_t launch_block(const Params& P, bool backward, cudaStream_t st) {
  const bool rbf = P.d.kernel == GPKL_KERNEL_RBF;
  const bool gp = P.d.posterior == GPKL_POST_GP;
  if (rbf && gp) return launch_kp<GPKL_KERNEL_RBF, GPKL_POST_GP>(P, backward, st);
  if (rbf && !gp) return launch_kp<GPKL_KERNEL_RBF, GPKL_POST_DIAG>(P, backward, st);
  if (!rbf && gp) return launch_kp<GPKL_KERNEL_CAUCHY, GPKL_POST_GP>(P, backward, st);
  return launch_kp<GPKL_KERNEL_CAUCHY, GPKL_POST_DIAG>(P, backward, st);
}

}  // namespace gpkl
