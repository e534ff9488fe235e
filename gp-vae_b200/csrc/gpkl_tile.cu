// Tile tier: the shared-prior GP posterior for 208 < T <= 512 (BASELINE config C4: T = 512), one CTA (256 threads) per
// (sequence, latent-dim) pair, one pair in flight per SM.
//
// Why a tier of its own.  A T = 512 factor is 0.5 MB packed: it fits neither the registers of a warp nor the shared
// memory of an SM, and the block tier's GEMM path kept TWO padded squares per pair in 2 x 148 workspace slots (627 MB,
// five times the L2) -- 300-580x the algorithmic DRAM bytes (profiles/traffic.json, round 1).  Here a pair owns ONE
// tile-packed lower triangle (nT (nT+1)/2 tiles of 64 x 64 floats, 576 KB at T = 512; 148 CTAs -> 85 MB < 126 MB L2;
// the backward adds the triangle of C') that stays L2-resident while it is reused pair after pair, and every O(T^3)
// phase streams 4 KB operand chunks L2 -> shared memory with bulk asynchronous copies (cp.async.bulk, the TMA engine;
// SASS: UBLKCP) completing on mbarriers, two stages per thread group.  Distributed shared memory was weighed and not
// used: a 4-CTA cluster would hold the triangle, but the measured DSMEM bandwidth (17-21 B/clk/SM,
// B300_MICROARCH.md) is half of what L2 delivers per SM (~42 B/clk) and every panel would have to be broadcast to
// three peers, whereas from L2 each chunk is read once by the one CTA that needs it.
//
// Micro-kernel.  A GROUP of 64 threads (two warps) owns a 64 x 64 output tile, 8 x 8 per thread (two 4-row blocks x two
// 4-column blocks, 32 apart), operands contraction-major in shared memory ([k][64]): four 128-bit loads (each one
// wavefront: 8 row-threads are the fast lane index, so a warp reads 128 contiguous bytes of A and 64 of B) feed 32
// packed FFMA2.  The CTA is four such groups working on DIFFERENT tiles, each with its own 2-stage chunk ring, so the
// only CTA-wide barriers are at phase boundaries.
//
// Tile layouts (all "contraction index major", so any 16 consecutive contraction steps are one contiguous 4 KB chunk):
//     L  tile (I,K):  [k][i]   (column-major)   operand of the panel update  raw(I,J) = K(I,J) - sum_K L(I,K) L(J,K)^T
//     X  tile (I,K):  [i][k]   (row-major)      X = L^-1, operand of the inverse and of the contraction
//     C' tile (I,L):  [i][l]   (row-major)
// Diagonal tiles carry explicit zeros in their other half and the rows / columns beyond a sequence's length are the
// identity, so no phase masks operands.
//
// Phases per pair (S = samples, T_b <= T_max ragged; nTb = ceil(T_b / 64)):
//   forward   for each 64-column panel J:  raw panel (split-K over the groups, partials reduced in shared memory with the
//             kernel matrix GENERATED in the same pass) -> 64 x 64 diagonal block factored + inverted in shared memory
//             -> rows below = raw x L_JJ^-T (micro-kernel, operands resident) -> the finished panel, row-major in shared
//             memory, is at once the B operand of  A(:,J) = X_p(:,J:) L_q(J:,J)  (A operand: the prior record,
//             streamed) and of z = m + L_q eps.  Nothing of L_q is ever stored row-major in global memory.
//   backward  the same factorisation (+ w = L_q^T g_z per panel), then X_q = L_q^-1 IN PLACE by 64-row blocks
//             (GEMM against the rows already inverted, then L_II^-1 from the left), C' by column prefix sums, and the
//             contraction  sum_kl dK_kl (g/2 K_p^-1 + X_q^T C')_kl  with the kernel derivative in the epilogue.
// The per-sequence prior records (L_p^-1, diag L_p, K_p^-1) come from the block tier's pre-pass, unchanged.
// Reference replaced: tf.cholesky / tf.matrix_inverse / tf.linalg.logdet / tf.matmul of
// src/Models/Full_GP_VAE_dynamic_time.py:165, :250-254 and TF autodiff through them (:361).
#include <stdlib.h>
#include <string.h>

#include "gpkl_common.cuh"
#include "gpkl_diag64.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int TS = 64;             // tile edge
constexpr int TF = TS * TS;        // floats per tile (16 KB)
constexpr int KC = 8;              // contraction steps per staged chunk
constexpr int CH = KC * TS;        // floats per operand chunk (2 KB)
constexpr int UC = TS / KC;        // chunks per 64-step tile of the contraction index
constexpr int STAGE_F = 2 * CH;    // one stage: A chunk | B chunk
constexpr int WSTG_F = 2 * STAGE_F;  // two stages per warp (8 KB); the two warps of a pair together hold one tile
constexpr int NW = 8;              // workers (warps) per CTA
constexpr int NTHR = 256;
constexpr int NVEC = 8;            // per-pair vectors of TP floats besides the per-sample ones

// ---- shared-memory carve-up (floats), identical on host and device ------------------------------------------------
struct TLay {
  int TP, nT, S;
  __host__ __device__ TLay(int Tmax, int S_) : S(S_) {
    TP = (Tmax + TS - 1) / TS * TS;
    if (TP < TS) TP = TS;
    nT = TP / TS;
  }
  __host__ __device__ int ntri() const { return nT * (nT + 1) / 2; }
  __host__ __device__ size_t nvec() const { return (size_t)NVEC + 2 * (size_t)(S > 1 ? S - 1 : 0); }
  __host__ __device__ size_t floats() const {
    return 64 /*red*/ + 32 /*mbarriers*/ + 16 /*flush counters*/ + 64 /*rdl*/ + 576 /*diag-block scratch*/ + (size_t)nT * TF + (size_t)NW * WSTG_F + TF +
           nvec() * TP;
  }
};

__host__ __device__ inline int tri(int I, int J) { return I * (I + 1) / 2 + J; }

// ---- mbarrier / bulk-copy (TMA engine) primitives -------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion counted in bytes on an mbarrier (16-byte aligned addresses and size)
__device__ __forceinline__ void bulk_g2s(float* smem_dst, const float* gsrc, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// One chunk of a staged contraction: arm `bar` with 2 x `bytes` and start the two bulk copies -- PREDICATED on lane 0 inside one
// asm block instead of a branch: the other lanes skip three instructions, there is no divergent region to open and re-converge
// (measured: the 230-300 cycles per chunk between the two __syncwarp of the issue step did not change, so they are the cost of
// the three asynchronous operations themselves, not of the branch), and the addresses are computed warp-uniformly.
__device__ __forceinline__ void issue_pair(int lane, float* dstA, const float* srcA, float* dstB, const float* srcB, unsigned bytes,
                                           uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.eq.s32 p, %0, 0;\n"
      "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %2;\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%3], [%4], %7, [%1];\n"
      "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%5], [%6], %7, [%1];\n"
      "}\n" ::"r"(lane),
      "r"(smem_u32(bar)), "r"(2 * bytes), "r"(smem_u32(dstA)), "l"(srcA), "r"(smem_u32(dstB)), "l"(srcB), "r"(bytes)
      : "memory");
}
// shared -> global bulk copy (bulk async-group)
__device__ __forceinline__ void bulk_s2g(float* gdst, const float* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes (st.shared / st.global by threads) ordered before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- worker (one warp) ------------------------------------------------------------------------------------------------
struct WCtx {
  int w, lane, ty, tx;  // warp, lane, micro-tile coordinates: rows 4ty+{0..3} (+32), cols 4tx+{0..3} (+16, +32, +48)
  float* stg;           // this warp's two stages (WSTG_F floats)
  uint64_t* bar;        // the two "stage full" barriers
  unsigned use0, use1;  // fills consumed per stage (phase parity)
  long long* dbg;       // developer aid: lane 0 of warp 0 of CTA 0 splits the staged loops into wait / compute / issue cycles
};

// named barrier of one warp PAIR (ids must be literals: a register id makes ptxas reserve all 16 hardware barriers)
__device__ __forceinline__ void pair_sync(int q) {
  switch (q) {
    case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
    case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
    case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
    default: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
  }
}

__device__ __forceinline__ void acc_zero(float (&acc)[8][16]) {
#pragma unroll
  for (int r = 0; r < 8; ++r)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[r][c] = 0.0f;
}

// acc[r][c] += sum_{kk < KC} A[kk][row(r)] * B[kk][col(c)]  (operand rows are 64 floats apart).  8 x 16 outputs per thread:
// six 128-bit loads (24 floats) feed 64 packed FFMA2 -- 0.75 bytes of shared-memory delivery per FMA; the 8 x 8 tile of the
// first version needed 1.0, exactly the SM's 128 B/clk at FP32 peak, and ran at 45-60 % of it.
__device__ __forceinline__ void mk_chunk(float (&acc)[8][16], const float* __restrict__ As, const float* __restrict__ Bs, int ty,
                                         int tx) {
  const float* ap = As + 4 * ty;
  const float* bp = Bs + 4 * tx;
#pragma unroll
  for (int kk = 0; kk < KC; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(ap + kk * TS);
    const float4 a1 = *reinterpret_cast<const float4*>(ap + kk * TS + 32);
    const float4 b0 = *reinterpret_cast<const float4*>(bp + kk * TS);
    const float4 b1 = *reinterpret_cast<const float4*>(bp + kk * TS + 16);
    const float4 b2 = *reinterpret_cast<const float4*>(bp + kk * TS + 32);
    const float4 b3 = *reinterpret_cast<const float4*>(bp + kk * TS + 48);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c < 16; c += 2) fma2(acc[r][c], acc[r][c + 1], a[r], a[r], b[c], b[c + 1]);
  }
}

// mk_chunk restricted to the accumulator rows r >= R0 (R0 = 0 or 4: the rows 32.. of the tile) and the column groups
// c >> 2 >= C0 (columns 16 C0 ..): the other operand entries are structural zeros of a triangular factor (the inverse of a
// diagonal block has no entries (k, c) with k > c), so neither their loads nor their FMAs are issued.
template <int R0, int C0>
__device__ __forceinline__ void mk_chunk_part(float (&acc)[8][16], const float* __restrict__ As, const float* __restrict__ Bs,
                                              int ty, int tx) {
  const float* ap = As + 4 * ty;
  const float* bp = Bs + 4 * tx;
#pragma unroll
  for (int kk = 0; kk < KC; ++kk) {
    float a[8], b[16];
#pragma unroll
    for (int h = R0 / 4; h < 2; ++h) {
      const float4 v = *reinterpret_cast<const float4*>(ap + kk * TS + 32 * h);
      a[4 * h] = v.x; a[4 * h + 1] = v.y; a[4 * h + 2] = v.z; a[4 * h + 3] = v.w;
    }
#pragma unroll
    for (int g = C0; g < 4; ++g) {
      const float4 v = *reinterpret_cast<const float4*>(bp + kk * TS + 16 * g);
      b[4 * g] = v.x; b[4 * g + 1] = v.y; b[4 * g + 2] = v.z; b[4 * g + 3] = v.w;
    }
#pragma unroll
    for (int r = R0; r < 8; ++r)
#pragma unroll
      for (int c = 4 * C0; c < 16; c += 2) fma2(acc[r][c], acc[r][c + 1], a[r], a[r], b[c], b[c + 1]);
  }
}

__device__ __forceinline__ int mrow(int ty, int r) { return 4 * ty + (r & 3) + 32 * (r >> 2); }
__device__ __forceinline__ int mcol(int tx, int c) { return 4 * tx + (c & 3) + 16 * (c >> 2); }

// tile stores / read-modify-writes from the micro-kernel's registers.
// column-major [c][i]: the 8 row-threads of a quarter warp touch 128 contiguous bytes (conflict-free, coalesced)
template <int SGN>
__device__ __forceinline__ void store_colmajor(float* __restrict__ dst, const float (&acc)[8][16], int ty, int tx) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float* d = dst + (size_t)mcol(tx, c) * TS + 4 * ty;
    *reinterpret_cast<float4*>(d) = make_float4(SGN * acc[0][c], SGN * acc[1][c], SGN * acc[2][c], SGN * acc[3][c]);
    *reinterpret_cast<float4*>(d + 32) = make_float4(SGN * acc[4][c], SGN * acc[5][c], SGN * acc[6][c], SGN * acc[7][c]);
  }
}
__device__ __forceinline__ void sub_colmajor(float* __restrict__ dst, const float (&acc)[8][16], int ty, int tx) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float* d = dst + (size_t)mcol(tx, c) * TS + 4 * ty;
    float4 lo = *reinterpret_cast<float4*>(d), hi = *reinterpret_cast<float4*>(d + 32);
    lo.x -= acc[0][c]; lo.y -= acc[1][c]; lo.z -= acc[2][c]; lo.w -= acc[3][c];
    hi.x -= acc[4][c]; hi.y -= acc[5][c]; hi.z -= acc[6][c]; hi.w -= acc[7][c];
    *reinterpret_cast<float4*>(d) = lo;
    *reinterpret_cast<float4*>(d + 32) = hi;
  }
}
__device__ __forceinline__ void add_colmajor(float (&acc)[8][16], const float* __restrict__ src, int ty, int tx) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float* d = src + (size_t)mcol(tx, c) * TS + 4 * ty;
    const float4 lo = *reinterpret_cast<const float4*>(d), hi = *reinterpret_cast<const float4*>(d + 32);
    acc[0][c] += lo.x; acc[1][c] += lo.y; acc[2][c] += lo.z; acc[3][c] += lo.w;
    acc[4][c] += hi.x; acc[5][c] += hi.y; acc[6][c] += hi.z; acc[7][c] += hi.w;
  }
}
// row-major [i][c]
template <int SGN>
__device__ __forceinline__ void store_rowmajor(float* __restrict__ dst, const float (&acc)[8][16], int ty, int tx) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float* d = dst + (size_t)mrow(ty, r) * TS + 4 * tx;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(d + 16 * j) =
          make_float4(SGN * acc[r][4 * j], SGN * acc[r][4 * j + 1], SGN * acc[r][4 * j + 2], SGN * acc[r][4 * j + 3]);
  }
}
__device__ __forceinline__ void sub_rowmajor(float* __restrict__ dst, const float (&acc)[8][16], int ty, int tx) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    float* d = dst + (size_t)mrow(ty, r) * TS + 4 * tx;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 v = *reinterpret_cast<float4*>(d + 16 * j);
      v.x -= acc[r][4 * j]; v.y -= acc[r][4 * j + 1]; v.z -= acc[r][4 * j + 2]; v.w -= acc[r][4 * j + 3];
      *reinterpret_cast<float4*>(d + 16 * j) = v;
    }
  }
}

// Staged contraction of one job: n chunks; issue(c, stage, bar, lane) is called by every lane (warp-uniform addresses) and
// starts the bulk copies of chunk c into `stage` (A chunk at stage, B chunk at stage + CH), lane 0 arming `bar` with their
// byte count (issue_pair: predicated, no divergent region).  B_RES: the B operand is resident in shared memory, bres(c) returns its chunk.
template <class IssueF>
__device__ __forceinline__ void prime_chunks(WCtx& W, int n, IssueF issue) {
  if (n > 0) issue(0, W.stg, &W.bar[0], W.lane);
  if (n > 1) issue(1, W.stg + STAGE_F, &W.bar[1], W.lane);
  __syncwarp();
}
struct NoPrep {
  __device__ __forceinline__ void operator()(int, float*) const {}
};
// prep(c, Bs): every lane, after chunk c has landed and before it is multiplied (transforms the staged B chunk in place)
template <bool B_RES, class IssueF, class BresF, class PrepF = NoPrep>
__device__ __forceinline__ void run_chunks(float (&acc)[8][16], WCtx& W, int n, IssueF issue, BresF bres, bool primed = false,
                                           PrepF prep = PrepF()) {
  if (!primed) prime_chunks(W, n, issue);
  for (int c = 0; c < n; ++c) {
    const int s = c & 1;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (W.dbg) t0 = clock64();
    if (s) { mbar_wait(&W.bar[1], W.use1 & 1u); ++W.use1; }
    else { mbar_wait(&W.bar[0], W.use0 & 1u); ++W.use0; }
    if (W.dbg) t1 = clock64();
    const float* As = W.stg + s * STAGE_F;
    const float* Bs = B_RES ? bres(c) : As + CH;
    if (!B_RES) prep(c, W.stg + s * STAGE_F + CH);
    mk_chunk(acc, As, Bs, W.ty, W.tx);
    __syncwarp();  // every lane is done with stage s
    if (W.dbg) t2 = clock64();
    if (c + 2 < n) issue(c + 2, W.stg + s * STAGE_F, &W.bar[s], W.lane);
    __syncwarp();
    if (W.dbg) { const long long t3 = clock64(); W.dbg[32] += t1 - t0; W.dbg[33] += t2 - t1; W.dbg[34] += t3 - t2; W.dbg[35] += 1; }
  }
}

// Ordered flush of split contractions: the parts of one destination tile are applied in part order (bit-reproducible, and
// exclusive access to the tile) through a shared-memory counter per tile.  Every lane calls both.
__device__ __forceinline__ void flush_begin(volatile int* cnt, int part) {
  if (part > 0) {
    while (*cnt < part) __nanosleep(40);  // (a tight spin of up to seven warps on one word starves the flushing warp's loads)
  }
  __threadfence_block();
}
__device__ __forceinline__ void flush_end(volatile int* cnt, int part, int lane) {
  __threadfence_block();
  __syncwarp();
  if (lane == 0) *cnt = part + 1;
}
// first warp whose chunk range [w U / 8, (w+1) U / 8) reaches beyond chunk index `off` (the first contributor to the tile
// whose chunks start at off)
__device__ __forceinline__ int first_warp_after(int U, int off) {
  int w = 0;
  while (((w + 1) * U) / NW <= off) ++w;
  return w;
}
// Chunk ranges when warps 6 and 7 (the "diagonal team") join a phase late: they get `st` chunks each at the END of the
// sequence, warps 0..5 share the rest evenly.  Range of warp w is [team_bound(w), team_bound(w + 1)).
__device__ __forceinline__ int team_bound(int U, int st, int w) {
  const int R = U - 2 * st;
  return w <= 6 ? (w * R) / 6 : (w == 7 ? R + st : U);
}
__device__ __forceinline__ int team_first_after(int U, int st, int off) {
  int w = 0;
  while (team_bound(U, st, w + 1) <= off) ++w;
  return w;
}
__device__ __forceinline__ void team_sync() { asm volatile("bar.sync 5, 64;" ::: "memory"); }

// Phase clock (developer aid, tools/tile_trace.py): thread 0 of CTA 0 accumulates the cycles between ticks into
// dbg[48 + k]; a no-op (NULL) for every other thread and whenever no trace buffer is set.
struct PhClock {
  long long t0;
  long long* dbg;
  long long* dbg2;  // the trace buffer for every thread of CTA 0 (the diagonal team clocks itself)
  __device__ __forceinline__ void start() { if (dbg) t0 = clock64(); }
  __device__ __forceinline__ void tick(int k) {
    if (dbg) { const long long t = clock64(); dbg[48 + k] += t - t0; t0 = t; }
  }
};

// Everything a pair's phases share.
struct Sm {
  double* red;
  uint64_t* bars;
  int* cnt;
  float *rdl, *dsm, *panel, *stg, *linv;
  float *ts, *mm, *dgq, *v0, *v1, *v2, *v3, *v4, *eps, *zacc;  // v0..v4: direction-specific vectors (see kernels)
  __device__ Sm(float* base, const TLay& L) {
    red = reinterpret_cast<double*>(base); base += 64;
    bars = reinterpret_cast<uint64_t*>(base); base += 32;
    cnt = reinterpret_cast<int*>(base); base += 16;
    rdl = base; base += 64;
    dsm = base; base += 576;
    panel = base; base += (size_t)L.nT * TF;
    stg = base; base += (size_t)NW * WSTG_F;
    linv = base; base += TF;
    ts = base; base += L.TP;
    mm = base; base += L.TP;
    dgq = base; base += L.TP;
    v0 = base; base += L.TP;
    v1 = base; base += L.TP;
    v2 = base; base += L.TP;
    eps = base; base += (size_t)L.S * L.TP;     // NVEC counts one eps and one zacc/extra vector; S > 1 adds 2 (S-1)
    zacc = base; base += (size_t)L.S * L.TP;
    v3 = zacc;  // backward (S == 1): the z accumulator's slot is free
    v4 = dgq;   // backward: diag L_q is not needed beyond the factorisation's own 64-entry window
  }
};

template <int KERNEL>
struct Pair {
  int T, nTb, Tact;
  float noise;
  KernC<KERNEL> kc;
  __device__ Pair(int T_, float ell, float sig, float noise_) : T(T_), nTb((T_ + TS - 1) / TS), Tact((T_ + TS - 1) / TS * TS),
                                                                 noise(noise_), kc(ell, sig) {}
  // K(i, c..) for four consecutive rows i0..i0+3 at column c (absolute indices); identity beyond the sequence
  __device__ __forceinline__ float4 kgen4(int c, int i0, const float* __restrict__ ts) const {
    const float4 t4 = *reinterpret_cast<const float4*>(ts + i0);
    const float tc = ts[c];
    const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i0 + e;
      const float v = kc.val(tv[e] - tc) + (i == c ? noise : 0.0f);
      o[e] = (i < T && c < T) ? v : (i == c ? 1.0f : 0.0f);
    }
    return make_float4(o[0], o[1], o[2], o[3]);
  }
};

// ---- the 64 x 64 diagonal block: factor + invert + publish by the DIAGONAL TEAM (warps 6 and 7, named barrier 5) ----------
// D: the block, column-major (D[c*64 + i], i >= c valid).  Out: dgl = diag(L), rdl = 1/diag(L); gt = the global L tile
// (column-major, zeros above the diagonal); D = L ROW-major (zeros above the diagonal); LT[c'][c] = Linv[c][c'] (the
// contraction-major operand of both uses: rows-below = raw Linv^T and X = Linv Y), zeros where c < c'.  Rsw, Xsw: one
// tile of scratch each.  Tl: number of real rows of the block (may exceed 64).  See gpkl_diag64.cuh for the algorithm.
__device__ __forceinline__ void diag_block64(float* __restrict__ D, float* __restrict__ gt, float* __restrict__ dgl,
                                             float* __restrict__ rdl, int Tl, int* bad, float* __restrict__ LT,
                                             float* __restrict__ Rsw, float* __restrict__ Xsw, float* __restrict__ small) {
  const int t = (int)threadIdx.x - (NTHR - 64);  // 0..63
  factor_invert64_rows(D, gt, dgl, rdl, Tl, bad, Rsw, Xsw, small, small + 64, t, [] { team_sync(); });
  for (int e = t; e < TF / 4; e += 64) {
    const int i = e >> 4, g = e & 15;
    *reinterpret_cast<float4*>(D + i * TS + 4 * g) = *reinterpret_cast<const float4*>(Rsw + swz64(i, g));
    *reinterpret_cast<float4*>(LT + i * TS + 4 * g) = *reinterpret_cast<const float4*>(Xsw + swz64(i, g));
  }
  fence_async();  // the scratch tiles are stage areas (thread writes); bulk copies overwrite them next
  team_sync();
}

// ---- the factorisation of K_q by 64-column panels -------------------------------------------------------------------------
// Lg: this CTA's tile-packed triangle in global memory (L tiles column-major).  After panel J: panel tiles I >= J hold the
// finished columns of the panel ROW-major (zeros above the diagonal), LinvT = L_JJ^-T operand, dgq[64J..] = diag.  hook(J)
// runs with the panel in place (all threads; the loop issues the CTA barrier behind it).
// Panel update: a tile's 8J chunks (64 x 64 x 8 contraction steps each) are dealt to the warps as contiguous ranges of ONE
// chunk sequence; a warp walks its range from the last tile to the first (so the tile it shares with its predecessor is
// flushed late and the one it shares with its successor early) and subtracts each finished segment from the tile, which
// all threads pre-filled with the generated kernel matrix, in part order (flush_begin / flush_end).
// The diagonal tile is updated first, by all warps; then warps 6-7 factor and invert it (the serial chain of the panel,
// ~20 K cycles) WHILE the other warps update the tiles below it, and join them for a smaller share when they are done.
// The team's factor + invert + publish in units of one worker warp's chunk time (~2.3-2.6 K cycles).  The routine measures 17 K
// cycles stand-alone but ~30 K next to six streaming workers, i.e. 10-13 chunks: with the first estimate (4) the team was dealt
// more of the panel update than it could finish in time and the workers waited for it.  Measured on C4 (sequences/s):
// 4 -> 1743, 6 -> 1762, 8 -> 1788, 10 -> 1802, 11 -> 1801, 13 -> 1799, 16 -> 1793  (GPKL_DIAG_CHUNKS overrides: experiments).
constexpr int kDiagChunks = 10;
__device__ int g_diag_chunks = kDiagChunks;

// kinv (forward only, else NULL): this sequence's K_p^-1 in float64 (lower triangle, column-major, pitch ldk; gpkl_prior64.cu);
// the pass that generates K_q also accumulates  tr(K_p^-1 (K_q + m m^T)) = sum_ij Kinv_ij (K_q,ij + m_i m_j)  into tr (per
// thread, float64; lower entries count twice) -- the reference's own float64 trace (Full_GP_VAE_dynamic_time.py:250-254).
template <int KERNEL, class HookF>
__device__ __forceinline__ void chol_panels_tile(const Pair<KERNEL>& pr, float* __restrict__ Lg, Sm& s, WCtx& W, int* bad,
                                                 PhClock& pc, HookF hook, const double* __restrict__ kinv = nullptr, int ldk = 0,
                                                 double* tr = nullptr) {
  const int tid = threadIdx.x;
  const int nTb = pr.nTb;
  const float* Arow = nullptr;  // operands of the current segment (bulk-copy sources)
  const float* Brow = nullptr;
  auto issue = [&](int c, float* st, uint64_t* bar, int lane) {
    issue_pair(lane, st, Arow + (size_t)c * CH, st + CH, Brow + (size_t)c * CH, CH * 4, bar);
  };
  auto nobres = [](int) { return (const float*)nullptr; };
  for (int J = 0; J < nTb; ++J) {
    const int m = nTb - J;  // tiles I = J .. nTb-1 of this panel
    const int n = UC * J;   // chunks per tile
    // ---- (1a) diagonal tile: K(J,J) - sum_{K<J} L(J,K) L(J,K)^T, all warps -------------------------------------------------
    const int dlo = (W.w * n) / NW, dhi = ((W.w + 1) * n) / NW;
    if (dhi > dlo) {  // the first copies fly while the kernel matrix is generated
      Arow = Brow = Lg + (size_t)tri(J, 0) * TF + (size_t)dlo * CH;
      prime_chunks(W, dhi - dlo, issue);
    }
    if (tid < NW) s.cnt[tid] = 0;
    if (!kinv) {
      for (int ti = 0; ti < m; ++ti) {
        float* dst = s.panel + (size_t)(J + ti) * TF;
        for (int e = tid * 4; e < TF; e += NTHR * 4) {
          const int c = e >> 6, i = e & 63;
          *reinterpret_cast<float4*>(dst + e) = pr.kgen4(TS * J + c, TS * (J + ti) + i, s.ts);
        }
      }
    } else {
      // each thread: four 4-row pieces per tile (columns c0 + 16 it, rows i0..i0+3); the K_p^-1 entries of the NEXT tile are
      // in flight (L2 round trips) while the current one is generated and accumulated
      constexpr int NIT = TF / (NTHR * 4);
      const int c0 = (tid * 4) >> 6, i0 = (tid * 4) & 63;
      const double* mm64 = reinterpret_cast<const double*>(s.v0);  // the mean as doubles (forward: v0 | v1, filled by the kernel)
      double2 q[NIT][2];
      auto fetch = [&](int ti) {
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const double* kp = kinv + (size_t)(TS * J + c0 + 16 * it) * ldk + TS * (J + ti) + i0;
          q[it][0] = __ldg(reinterpret_cast<const double2*>(kp));
          q[it][1] = __ldg(reinterpret_cast<const double2*>(kp + 2));
        }
      };
      fetch(0);
      // FP64 is the scarce pipe here (measured ~16 DFMA per clock per SM on this part): three FP64 operations per entry -- one
      // conversion and two DFMA, the means are kept as doubles in shared memory (mm64) -- and weight-2 / weight-1
      // accumulators instead of a multiply; masks (diagonal tile, rows beyond T) are selects on the K_p^-1 entry
      double acc2 = 0.0, acc1 = 0.0;
      for (int ti = 0; ti < m; ++ti) {
        float* dst = s.panel + (size_t)(J + ti) * TF;
        double2 qc[NIT][2];
#pragma unroll
        for (int it = 0; it < NIT; ++it) { qc[it][0] = q[it][0]; qc[it][1] = q[it][1]; }
        if (ti + 1 < m) fetch(ti + 1);
        const int row = TS * (J + ti) + i0;
        const double mi[4] = {mm64[row], mm64[row + 1], mm64[row + 2], mm64[row + 3]};
        const bool inner = ti > 0 && TS * (J + ti) + TS <= pr.T;  // (uniform) a tile below the diagonal one, all rows real
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          const int col = TS * J + c0 + 16 * it;
          const float4 kv = pr.kgen4(col, row, s.ts);
          *reinterpret_cast<float4*>(dst + (c0 + 16 * it) * TS + i0) = kv;
          const double mc = mm64[col];
          const float kf[4] = {kv.x, kv.y, kv.z, kv.w};
          double kd[4] = {qc[it][0].x, qc[it][0].y, qc[it][1].x, qc[it][1].y};
          if (inner) {
#pragma unroll
            for (int e = 0; e < 4; ++e) acc2 = fma(kd[e], fma(mi[e], mc, (double)kf[e]), acc2);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = row + e;
              const double k2 = (i > col && i < pr.T) ? kd[e] : 0.0, k1 = (i == col && i < pr.T) ? kd[e] : 0.0;
              const double v = fma(mi[e], mc, (double)kf[e]);
              acc2 = fma(k2, v, acc2);
              acc1 = fma(k1, v, acc1);
            }
          }
        }
      }
      const double acc = 2.0 * acc2 + acc1;
      *tr += acc;
    }
    __syncthreads();
    pc.tick(11);
    if (dhi > dlo) {
      float acc[8][16];
      acc_zero(acc);
      run_chunks<false>(acc, W, dhi - dlo, issue, nobres, true);
      pc.tick(12);
      const int part = W.w - first_warp_after(n, 0);
      volatile int* cnt = s.cnt;
      flush_begin(cnt, part);
      sub_colmajor(s.panel + (size_t)J * TF, acc, W.ty, W.tx);
      flush_end(cnt, part, W.lane);
      pc.tick(13);
    }
    if (J > 0) __syncthreads();
    pc.tick(1);
    // ---- (1b) + (2): the tiles below the diagonal one || the diagonal block factored, inverted, published ---------------------
    float* D = s.panel + (size_t)J * TF;
    float* gdiag = Lg + (size_t)tri(J, J) * TF;
    const bool team = J > 0 && m > 1;
    if (!team) {  // first panel (nothing to update) or last one (no tiles below): the team works alone, all stage areas are idle
      if (W.w >= 6) diag_block64(D, gdiag, s.dgq + TS * J, s.rdl, pr.T - TS * J, bad, s.linv, s.stg, s.stg + TF, s.dsm);
      __syncthreads();
    } else {
      const int U = (m - 1) * n;  // chunks of the tiles below the diagonal one
      const int kdc = g_diag_chunks;
      int st = (U + 2 * kdc) / NW - kdc;  // the team's share once it has finished the diagonal block
      if (st < 0) st = 0;
      if (W.w >= 6) {
        float* mine = s.stg + (size_t)6 * WSTG_F;  // the team's own stage areas (one tile) as scratch; panel tile 0 is free (J > 0)
        const long long tt0 = (pc.dbg2 && tid == NTHR - 64) ? clock64() : 0;
        diag_block64(D, gdiag, s.dgq + TS * J, s.rdl, pr.T - TS * J, bad, s.linv, mine, s.panel, s.dsm);
        if (pc.dbg2 && tid == NTHR - 64) pc.dbg2[57] += clock64() - tt0;
      }
      const int lo = team_bound(U, st, W.w), hi = team_bound(U, st, W.w + 1);
      if (hi > lo) {
        for (int ti = (hi - 1) / n; ti >= lo / n; --ti) {
          const int c0 = (lo > ti * n ? lo : ti * n) - ti * n, c1 = (hi < (ti + 1) * n ? hi : (ti + 1) * n) - ti * n;
          float acc[8][16];
          acc_zero(acc);
          Arow = Lg + (size_t)tri(J + 1 + ti, 0) * TF + (size_t)c0 * CH;  // chunk (K, q) of a row of tiles: contiguous
          Brow = Lg + (size_t)tri(J, 0) * TF + (size_t)c0 * CH;
          run_chunks<false>(acc, W, c1 - c0, issue, nobres);
          const int part = W.w - team_first_after(U, st, ti * n);
          volatile int* cnt = s.cnt + 1 + ti;
          flush_begin(cnt, part);
          sub_colmajor(s.panel + (size_t)(J + 1 + ti) * TF, acc, W.ty, W.tx);
          flush_end(cnt, part, W.lane);
        }
      }
      __syncthreads();
    }
    pc.tick(2);
    // ---- (3) rows below: L(I,J) = raw(I,J) L_JJ^-T, one tile per warp, operands resident ---------------------------------
    for (int ti = 1 + W.w; ti < m; ti += NW) {
      float* tile = s.panel + (size_t)(J + ti) * TF;
      float acc[8][16];
      acc_zero(acc);
      // operands swapped (acc[c][i] = L(i,c)): the conflict-free "column-major" store pattern then writes the ROW-major
      // tile the product / z / w read, and the global column-major L tile takes the strided pattern (no banks there)
      // (L_JJ^-T has no entries (k, c) with k > c: the contraction steps k >= 32 only reach the columns c >= 32, i.e. the
      //  second half of the accumulator rows)
#pragma unroll 1
      for (int q = 0; q < UC / 2; ++q) mk_chunk(acc, s.linv + q * CH, tile + q * CH, W.ty, W.tx);
#pragma unroll 1
      for (int q = UC / 2; q < UC; ++q) mk_chunk_part<4, 0>(acc, s.linv + q * CH, tile + q * CH, W.ty, W.tx);
      store_rowmajor<1>(Lg + (size_t)tri(J + ti, J) * TF, acc, W.ty, W.tx);
      __syncwarp();  // the warp has read the whole raw tile
      store_colmajor<1>(tile, acc, W.ty, W.tx);
    }
    fence_async();  // the L tiles written above are read by bulk copies from the next panel on
    __syncthreads();
    pc.tick(3);
    // ---- (4) per-panel consumers of the finished panel -----------------------------------------------------------------
    hook(J);
    __syncthreads();
    pc.tick(4);
  }
}

__device__ __forceinline__ void load_vectors(const Params& P, int p, int b, int dd, int T, long long r0, const TLay& L, Sm& s,
                                             bool backward) {
  const GpklDesc& d = P.d;
  const int S = d.S, TP = L.TP;
  for (int i = threadIdx.x; i < TP; i += NTHR) {
    s.ts[i] = (i < T) ? P.times[(size_t)b * d.T_max + i] : 0.0f;
    s.mm[i] = (i < T) ? P.mean[(size_t)(r0 + i) * d.D + dd] : 0.0f;
    for (int sx = 0; sx < S; ++sx) {
      s.eps[(size_t)sx * TP + i] = (i < T) ? eps_value(P, ((size_t)p * S + sx) * d.T_max + i) : 0.0f;
      if (!backward) s.zacc[(size_t)sx * TP + i] = s.mm[i];
    }
    if (backward) {  // S == 1
      s.v0[i] = (i < T && P.g_z) ? P.g_z[((size_t)r0 + i) * d.D + dd] : 0.0f;  // u = g_z
      s.v3[i] = 0.0f;                                                         // running column sums of C'
    }
  }
}

// out[j] = sum_{l < lim(j)} R[l*ldr + j] * m[l] for j < T (lim = T, or j+1 for a lower-triangular record whose upper part is
// exact zeros): the two matrix-vector products against the per-sequence prior record (a = L_p^-1 m, alpha = K_p^-1 m).
// 128 groups of four columns (one 128-bit load per row, a warp reads 512 contiguous bytes) x 2 halves of the row range,
// sixteen loads in flight per thread; the halves meet in `scratch` (>= 1024 floats).  All 256 threads; TP <= 512.
__device__ __forceinline__ void matvec_record(const float* __restrict__ R, int ldr, const float* __restrict__ m, int T, bool lower,
                                              float* __restrict__ out, float* __restrict__ scratch) {
  const int tid = threadIdx.x, j = 4 * (tid & 127), h = tid >> 7;
  float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (j < T) {
    const int lend = lower ? (T < j + 4 ? T : j + 4) : T;
    const int l0 = h ? lend / 2 : 0, l1 = h ? lend : lend / 2;
    const float* rp = R + (size_t)l0 * ldr + j;
    int l = l0;
    for (; l + 16 <= l1; l += 16) {
      float4 x[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) x[e] = __ldg(reinterpret_cast<const float4*>(rp + (size_t)e * ldr));
      rp += (size_t)16 * ldr;
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const float mv = m[l + e];
        acc.x = fmaf(x[e].x, mv, acc.x); acc.y = fmaf(x[e].y, mv, acc.y); acc.z = fmaf(x[e].z, mv, acc.z); acc.w = fmaf(x[e].w, mv, acc.w);
      }
    }
    for (; l < l1; ++l) {
      const float4 x = __ldg(reinterpret_cast<const float4*>(rp));
      rp += ldr;
      const float mv = m[l];
      acc.x = fmaf(x.x, mv, acc.x); acc.y = fmaf(x.y, mv, acc.y); acc.z = fmaf(x.z, mv, acc.z); acc.w = fmaf(x.w, mv, acc.w);
    }
  }
  *reinterpret_cast<float4*>(scratch + h * 512 + j) = acc;
  __syncthreads();
  if (h == 0 && j < T) {
    const float4 a = *reinterpret_cast<const float4*>(scratch + j), b = *reinterpret_cast<const float4*>(scratch + 512 + j);
    *reinterpret_cast<float4*>(out + j) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
  fence_async();  // `scratch` is a stage area: bulk copies overwrite it next
  __syncthreads();
}

// ---- forward ---------------------------------------------------------------------------------------------------------------
template <int KERNEL>
__global__ void __launch_bounds__(NTHR, 1) fwd_tile(Params P) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ int bad;
  if (*P.prior_flag == 0) return;  // ell_p differs between latent dims: the per-pair kernel launched behind this one serves
  const GpklDesc& d = P.d;
  const TLay L(d.T_max, d.S);
  Sm s(smem_f, L);
  const int tid = threadIdx.x;
  WCtx W;
  W.w = tid >> 5; W.lane = tid & 31;
  W.ty = W.lane & 7; W.tx = W.lane >> 3;
  W.stg = s.stg + (size_t)W.w * WSTG_F;
  W.bar = s.bars + 2 * W.w;
  W.use0 = W.use1 = 0;
  W.dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg : nullptr;
  if (tid == 0) {
    for (int i = 0; i < 2 * NW; ++i) mbar_init(&s.bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int S = d.S, TP = L.TP;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  float* Lg = P.scratch + (size_t)blockIdx.x * P.scratch_stride;
  PhClock pc;
  pc.dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg : nullptr;
  pc.dbg2 = (P.dbg && blockIdx.x == 0) ? P.dbg : nullptr;
  pc.start();
  for (int p = blockIdx.x; p < d.B * d.D; p += gridDim.x) {
    const int b = p / d.D, dd = p - b * d.D;
    const int T = P.lengths[b];
    const long long r0 = P.offsets[b];
    __syncthreads();
    if (T <= 0) {
      if (tid == 0) {
        P.kl_pairs[p] = 0.0f;
        if (P.logdets) { P.logdets[2 * p] = 0.0f; P.logdets[2 * p + 1] = 0.0f; }
      }
      continue;
    }
    if (tid == 0) bad = 0;
    load_vectors(P, p, b, dd, T, r0, L, s, false);
    {
      double* mm64 = reinterpret_cast<double*>(s.v0);  // (v0 and v1 are adjacent: TP doubles)
      for (int i = tid; i < TP; i += NTHR) mm64[i] = (i < T) ? (double)P.mean[(size_t)(r0 + i) * d.D + dd] : 0.0;
    }
    // this sequence's float64 record: K_p^-1 (lower, column-major, pitch TP) and log|K_p| (gpkl_prior64.cu)
    const double* __restrict__ kinv = reinterpret_cast<const double*>(P.prior + (size_t)b * P.prior_stride);
    __syncthreads();
    const Pair<KERNEL> pr(T, P.ell_q[dd], sig, noise);
    const int nTb = pr.nTb;
    double tr = 0.0;  // this thread's share of tr(K_p^-1 (K_q + m m^T))
    pc.tick(0);
    chol_panels_tile<KERNEL>(pr, Lg, s, W, &bad, pc, [&](int J) {
      const int m = nTb - J;
      // z_s += L_q(:, panel J) eps_s(panel J): one row per thread, 128-bit row reads started at a lane-dependent column
      for (int idx = tid; idx < m * TS; idx += NTHR) {
        const float* row = s.panel + (size_t)J * TF + (size_t)idx * TS;  // tiles are contiguous: row idx of the panel
        for (int sx = 0; sx < S; ++sx) {
          const float* ev = s.eps + (size_t)sx * TP + TS * J;
          float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int cc = ((j + idx) & 15) * 4;
            const float4 l4 = *reinterpret_cast<const float4*>(row + cc);
            const float4 e4 = *reinterpret_cast<const float4*>(ev + cc);
            a0 = fmaf(l4.x, e4.x, a0); a1 = fmaf(l4.y, e4.y, a1); a0 = fmaf(l4.z, e4.z, a0); a1 = fmaf(l4.w, e4.w, a1);
          }
          s.zacc[(size_t)sx * TP + TS * J + idx] += a0 + a1;
        }
      }
    }, kinv, TP, &tr);
    // ---- z out, KL = 1/2 [tr(K_p^-1 (K_q + m m^T)) - T + log|K_p| - log|K_q|] ----------------------------------------------
    for (int i = tid; i < T; i += NTHR)
      for (int sx = 0; sx < S; ++sx) P.z[((size_t)S * r0 + (size_t)sx * T + i) * d.D + dd] = s.zacc[(size_t)sx * TP + i];
    double ldq = 0.0;
    for (int i = tid; i < T; i += NTHR) ldq += 2.0 * log((double)s.dgq[i]);
    const double trs = block_sum(tr, s.red);
    ldq = block_sum(ldq, s.red);
    if (tid == 0) {
      const double ldp = kinv[(size_t)TP * TP];
      P.kl_pairs[p] = (float)(0.5 * (trs - (double)T + ldp - ldq));
      if (P.logdets) { P.logdets[2 * p] = (float)ldp; P.logdets[2 * p + 1] = (float)ldq; }
      if (bad && P.status) atomicAdd(P.status, 1);
    }
    pc.tick(5);
    if (pc.dbg) pc.dbg[63] += 1;
  }
}

// ---- backward (S == 1) --------------------------------------------------------------------------------------------------------
template <int KERNEL>
__global__ void __launch_bounds__(NTHR, 1) bwd_tile(Params P) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ int bad;
  if (*P.prior_flag == 0) return;
  const GpklDesc& d = P.d;
  const TLay L(d.T_max, d.S);
  Sm s(smem_f, L);
  const int tid = threadIdx.x;
  WCtx W;
  W.w = tid >> 5; W.lane = tid & 31;
  W.ty = W.lane & 7; W.tx = W.lane >> 3;
  W.stg = s.stg + (size_t)W.w * WSTG_F;
  W.bar = s.bars + 2 * W.w;
  W.use0 = W.use1 = 0;
  W.dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg : nullptr;
  if (tid == 0) {
    for (int i = 0; i < 2 * NW; ++i) mbar_init(&s.bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int TP = L.TP;
  const float noise = d.noise, sig = (float)(1.0 - (double)noise);
  const int ldr = TP + 4;
  const double g_sum = P.g_kl_sum ? *P.g_kl_sum : 1.0;
  float* Lg = P.scratch + (size_t)blockIdx.x * P.scratch_stride;  // L tiles, overwritten in place by X tiles
  float* Vg = Lg + (size_t)L.ntri() * TF;                         // L_JJ^-T of the nT diagonal blocks
  float* cumtab = Vg + (size_t)L.nT * TF;                         // [nT][TP]: column sums of eps .* X over the rows above a row block
  float* u = s.v0;    // g_z
  float* w = s.v1;    // L_q^T g_z
  float* pd = s.v2;   // 1/2 w eps - g/2
  float* cum = s.v3;  // running column sums  sum_{j < i} eps_j X(j, l)
  PhClock pc;
  pc.dbg = (P.dbg && blockIdx.x == 0 && tid == 0) ? P.dbg : nullptr;
  pc.dbg2 = (P.dbg && blockIdx.x == 0) ? P.dbg : nullptr;
  pc.start();
  for (int p = blockIdx.x; p < d.B * d.D; p += gridDim.x) {
    const int b = p / d.D, dd = p - b * d.D;
    const int T = P.lengths[b];
    const long long r0 = P.offsets[b];
    __syncthreads();
    if (T <= 0) {
      if (tid == 0 && P.gq_pairs) P.gq_pairs[p] = 0.0f;
      continue;
    }
    const float g = (float)(g_sum + (P.g_kl_pairs ? (double)P.g_kl_pairs[p] : 0.0));
    const float hg = 0.5f * g;
    if (tid == 0) bad = 0;
    load_vectors(P, p, b, dd, T, r0, L, s, true);
    __syncthreads();
    const float lq = P.ell_q[dd];
    const Pair<KERNEL> pr(T, lq, sig, noise);
    const int nTb = pr.nTb;
    pc.tick(0);
    chol_panels_tile<KERNEL>(pr, Lg, s, W, &bad, pc, [&](int J) {
      const int m = nTb - J;
      // w(panel J) = L_q(:, panel J)^T g_z: four row-interleaved partial sums per column, summed in a fixed order
      {
        const int c = tid & 63, part = tid >> 6;
        const float* col = s.panel + (size_t)J * TF + c;
        float acc = 0.0f;
        for (int idx = part; idx < m * TS; idx += 4) acc = fmaf(col[(size_t)idx * TS], u[TS * J + idx], acc);
        s.stg[part * TS + c] = acc;  // (the stage areas are idle between the phases)
      }
      // L_JJ^-T is needed again when X_q = L_q^-1 is formed: keep it in the slot instead of inverting the block twice
      {
        float* vg = Vg + (size_t)J * TF;
        for (int e = tid * 4; e < TF; e += NTHR * 4) *reinterpret_cast<float4*>(vg + e) = *reinterpret_cast<const float4*>(s.linv + e);
      }
      __syncthreads();
      if (tid < TS) {
        const float wk = (s.stg[tid] + s.stg[TS + tid]) + (s.stg[2 * TS + tid] + s.stg[3 * TS + tid]);
        w[TS * J + tid] = wk;
        pd[TS * J + tid] = 0.5f * wk * s.eps[TS * J + tid] - hg;
      }
      fence_async();  // the stage area served as scratch (thread writes); bulk copies overwrite it next
    });
    // ---- X_q = L_q^-1 in place, 64-row blocks; C' alongside ----------------------------------------------------------------
    for (int I = 0; I < nTb; ++I) {
      // L_II^-T (kept by the factorisation) -> linv (the operand of X(I,C) = L_II^-1 Y), and transposed -- through an
      // XOR-swizzled scratch tile, both passes conflict-free -- into panel tile I: the diagonal X tile, row-major
      {
        const float* src = Vg + (size_t)I * TF;
        for (int e = tid * 4; e < TF; e += NTHR * 4) *reinterpret_cast<float4*>(s.linv + e) = __ldcg(reinterpret_cast<const float4*>(src + e));
        if (tid < NW) s.cnt[tid] = 0;
      }
      if (tid == 0) bulk_wait_read0();  // the bulk stores of the previous row block have finished reading the panel
      __syncthreads();
      {
        float* tmp = s.stg;
        for (int e = tid; e < TF; e += NTHR) {
          const int c = e >> 6, r = e & 63;  // LT[c][r] = Linv[r][c]
          tmp[r * TS + (c ^ (r & 31))] = s.linv[e];
        }
        __syncthreads();
        float* Xd = s.panel + (size_t)I * TF;
        for (int e = tid; e < TF; e += NTHR) {
          const int r = e >> 6, c = e & 63;
          Xd[e] = tmp[r * TS + (c ^ (r & 31))];
        }
      }
      fence_async();                     // stage areas were used as scratch by threads
      __syncthreads();
      pc.tick(6);
      // off-diagonal tiles C < I:  Y(C) = -sum_{K=C}^{I-1} L(I,K) X(K,C).  Tile C has 8 (I-C) chunks; the chunk sequence of the
      // row block is cut into 8 equal ranges (one per warp, walked from its last tile to its first), segments flushed into
      // the shared-memory tile in part order -- the scheme of the panel update above
      if (I > 0) {
        const int U = UC * I * (I + 1) / 2;
        const int lo = (W.w * U) / NW, hi = ((W.w + 1) * U) / NW;
        if (hi > lo) {
          for (int C = I - 1; C >= 0; --C) {
            const int off = UC * (C * I - C * (C - 1) / 2), n = UC * (I - C);  // chunks of tiles 0..C-1, of tile C
            if (off >= hi || off + n <= lo) continue;
            const int c0 = (lo > off ? lo : off) - off, c1 = (hi < off + n ? hi : off + n) - off;
            float acc[8][16];
            acc_zero(acc);
            const float* Arow = Lg + (size_t)tri(I, C) * TF;  // L(I,C), L(I,C+1), ... are consecutive tiles
            run_chunks<false>(acc, W, c1 - c0,
                              [&](int c, float* st, uint64_t* bar, int lane) {
                                const int cc = c0 + c, K = C + cc / UC, qq = cc % UC;  // operands swapped: acc[c][i] = sum_k X(k,c) L(i,k)
                                issue_pair(lane, st + CH, Arow + (size_t)cc * CH, st, Lg + (size_t)tri(K, C) * TF + (size_t)qq * CH, CH * 4, bar);
                              },
                              [](int) { return (const float*)nullptr; });
            const int part = W.w - first_warp_after(U, off);
            volatile int* cnt = s.cnt + C;
            float* Y = s.panel + (size_t)C * TF;
            flush_begin(cnt, part);  // (the transposed accumulators land ROW-major through the conflict-free pattern)
            if (part == 0) store_colmajor<-1>(Y, acc, W.ty, W.tx);
            else sub_colmajor(Y, acc, W.ty, W.tx);
            flush_end(cnt, part, W.lane);
          }
        }
      }
      __syncthreads();
      pc.tick(14);
      // X(I,C) = L_II^-1 Y(C), one tile per warp, operands resident
      for (int C = W.w; C < I; C += NW) {
        float* Y = s.panel + (size_t)C * TF;
        float acc[8][16];
        acc_zero(acc);
        // acc[c][r] = X(r,c); L_II^-1 has no entries (r, k) with k > r: chunk q (k >= 8 q) only reaches the accumulator
        // column groups r >= 16 g with 16 g + 15 >= 8 q
#pragma unroll 1
        for (int q = 0; q < 2; ++q) mk_chunk(acc, Y + q * CH, s.linv + q * CH, W.ty, W.tx);
#pragma unroll 1
        for (int q = 2; q < 4; ++q) mk_chunk_part<0, 1>(acc, Y + q * CH, s.linv + q * CH, W.ty, W.tx);
#pragma unroll 1
        for (int q = 4; q < 6; ++q) mk_chunk_part<0, 2>(acc, Y + q * CH, s.linv + q * CH, W.ty, W.tx);
#pragma unroll 1
        for (int q = 6; q < 8; ++q) mk_chunk_part<0, 3>(acc, Y + q * CH, s.linv + q * CH, W.ty, W.tx);
        __syncwarp();  // the warp has read all of Y
        store_colmajor<1>(Y, acc, W.ty, W.tx);
      }
      fence_async();  // panel tiles (X row block I) are read by the bulk stores below
      __syncthreads();
      pc.tick(7);
      if (tid == 0) {  // X(I, 0..I) -> global, in place of L(I, 0..I) (every read of that row of L tiles is complete)
        bulk_s2g(Lg + (size_t)tri(I, 0) * TF, s.panel, (unsigned)((I + 1) * TF * 4));
        bulk_commit();
      }
      // C'(i,l) = w_i sum_{j<i} eps_j X(j,l) + pd_i X(i,l) is NOT stored: the contraction rebuilds its chunks from the X chunks
      // it streams anyway (see below); all it needs from here are the column sums over the rows ABOVE each row block
      for (int l = tid; l < TS * (I + 1); l += NTHR) {
        const int Cb = l >> 6, lc = l & 63;
        const float* xs = s.panel + (size_t)Cb * TF + lc;
        float cm = cum[l];
        cumtab[(size_t)I * TP + l] = cm;
#pragma unroll 8
        for (int r = 0; r < TS; ++r) cm = fmaf(s.eps[TS * I + r], xs[r * TS], cm);
        cum[l] = cm;
      }
      __syncthreads();
      pc.tick(8);
    }
    if (tid == 0) bulk_wait0();  // X tiles are in global memory
    __syncthreads();
    // ---- d/d mean: alpha = K_p^-1 m from the record (symmetric: coalesced over k) -------------------------------------------
    const float* __restrict__ kinv = P.prior + (size_t)b * P.prior_stride;
    matvec_record(kinv, ldr, s.mm, T, false, s.v4, s.stg);
    for (int k = tid; k < T; k += NTHR) P.g_mean[(size_t)(r0 + k) * d.D + dd] = g * s.v4[k] + u[k];
    __syncthreads();
    pc.tick(9);
    // ---- contraction: sum_{k != l} dK_kl (hg K_p^-1 + X_q^T C')_kl, tiles by shell max(kt,lt) (longest first), boustrophedon --
    double total = 0.0;
    for (int blk = 0; blk * NW < nTb * nTb; ++blk) {
      const int n = blk * NW + ((blk & 1) ? NW - 1 - W.w : W.w);
      if (n >= nTb * nTb) continue;
      int mx = (int)sqrtf((float)n);
      while ((mx + 1) * (mx + 1) <= n) ++mx;
      while (mx * mx > n) --mx;
      const int pos = n - mx * mx;
      const int kt = pos <= mx ? mx : pos - mx - 1, lt = pos <= mx ? pos : mx;
      float acc[8][16];
      acc_zero(acc);
      // B operand: the chunk of X(Ib, lt) (rows i = 64 mx + 8 c .., columns l of tile lt) turned into the chunk of C' in place:
      // C'(i,l) = w_i cs(l) + pd_i X(i,l),  cs(l) += eps_i X(i,l)  -- two columns per lane, the running sums carried in
      // registers from chunk to chunk (they start at the table entry of row block mx).  No C' matrix exists in memory: the
      // backward slot is the X triangle + the diagonal-block inverses (720 KB per CTA at T = 512, 107 MB for 148 CTAs: L2)
      float2 cs = *reinterpret_cast<const float2*>(cumtab + (size_t)mx * TP + TS * lt + 2 * W.lane);
      run_chunks<false>(acc, W, UC * (nTb - mx),
                        [&](int c, float* st, uint64_t* bar, int lane) {
                          const int Ib = mx + c / UC, qq = c % UC;
                          issue_pair(lane, st, Lg + (size_t)tri(Ib, kt) * TF + (size_t)qq * CH, st + CH,
                                     Lg + (size_t)tri(Ib, lt) * TF + (size_t)qq * CH, CH * 4, bar);
                        },
                        [](int) { return (const float*)nullptr; }, false,
                        [&](int c, float* Bs) {
                          const int i0 = TS * mx + KC * c;
                          float2* bp = reinterpret_cast<float2*>(Bs) + W.lane;
                          const float4 w0 = *reinterpret_cast<const float4*>(w + i0), w1 = *reinterpret_cast<const float4*>(w + i0 + 4);
                          const float4 p0 = *reinterpret_cast<const float4*>(pd + i0), p1 = *reinterpret_cast<const float4*>(pd + i0 + 4);
                          const float4 e0 = *reinterpret_cast<const float4*>(s.eps + i0), e1 = *reinterpret_cast<const float4*>(s.eps + i0 + 4);
                          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                          const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                          const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
                          for (int r = 0; r < KC; ++r) {
                            const float wi = wv[r], pi = pv[r], ei = ev[r];
                            const float2 x = bp[r * (TS / 2)];
                            bp[r * (TS / 2)] = make_float2(fmaf(wi, cs.x, pi * x.x), fmaf(wi, cs.y, pi * x.y));
                            cs.x = fmaf(ei, x.x, cs.x);
                            cs.y = fmaf(ei, x.y, cs.y);
                          }
                          __syncwarp();
                        });
      float tl[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) tl[c] = s.ts[TS * lt + mcol(W.tx, c)];
      float part = 0.0f;
      // K_p^-1 rows: the loads of row r+1 are issued before row r is consumed (each is an L2 round trip)
      float4 kn[4];
      {
        const float* krow = kinv + (size_t)(TS * kt + mrow(W.ty, 0)) * ldr + TS * lt + 4 * W.tx;
#pragma unroll
        for (int j = 0; j < 4; ++j) kn[j] = __ldg(reinterpret_cast<const float4*>(krow + 16 * j));
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int k = TS * kt + mrow(W.ty, r);
        const float tk = s.ts[k];
        float4 kq[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) kq[j] = kn[j];
        if (r + 1 < 8) {
          const float* krow = kinv + (size_t)(TS * kt + mrow(W.ty, r + 1)) * ldr + TS * lt + 4 * W.tx;
#pragma unroll
          for (int j = 0; j < 4; ++j) kn[j] = __ldg(reinterpret_cast<const float4*>(krow + 16 * j));
        }
        const float kv[16] = {kq[0].x, kq[0].y, kq[0].z, kq[0].w, kq[1].x, kq[1].y, kq[1].z, kq[1].w,
                              kq[2].x, kq[2].y, kq[2].z, kq[2].w, kq[3].x, kq[3].y, kq[3].z, kq[3].w};
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int l = TS * lt + mcol(W.tx, c);
          const float dt = tk - tl[c];
          const float dk = pr.kc.dval_fast(dt);
          const float wv = fmaf(hg, kv[c], acc[r][c]);
          part = fmaf((k < T && l < T && k != l) ? wv : 0.0f, dk, part);
        }
      }
      total += (double)part;
    }
    const double gq = block_sum(total, s.red);
    if (tid == 0) {
      P.gq_pairs[p] = (float)gq;
      if (bad && P.status) atomicAdd(P.status, 1);
    }
    pc.tick(10);
    if (pc.dbg) pc.dbg[63] += 1;
  }
}

size_t tile_smem_bytes(const GpklDesc& d) { return TLay(d.T_max, d.S).floats() * sizeof(float); }

}  // namespace

bool tile_tier_supports(const GpklDesc& d, bool backward) {
  static const bool off = [] { const char* e = getenv("GPKL_TILE"); return e && e[0] == '0'; }();
  if (off) return false;
  if (d.posterior != GPKL_POST_GP || d.T_max <= 208 || d.T_max > 512) return false;
  if (backward && d.S != 1) return false;
  return tile_smem_bytes(d) <= kMaxDynSmem;
}

size_t tile_slot_floats(const GpklDesc& d) {
  const TLay L(d.T_max, d.S);
  return ((size_t)L.ntri() + (size_t)L.nT) * TF + (size_t)L.nT * L.TP;  // L / X tiles, L_JJ^-T of the diagonal blocks, column-sum table
}

// The shared-prior kernel of the tile tier; the caller has launched the block tier's pre-pass (records in P.prior) before
// it and launches the per-pair kernel (skip_if_shared) behind it.
// The slots are compact (576 KB per CTA at T = 512 in forward, 148 CTAs -> 85 MB; backward 1.28 MB per CTA) and reused pair
// after pair: DRAM READS of the forward kernel are down to the inputs and the prior records (21 KB per pair at T = 512, ncu).
// DRAM WRITES are not: the L2 of this part writes every dirty tile back once (596 KB per pair in forward -- exactly the
// bytes the kernel stores), with or without a persisting access-policy window over the slots (tried: no change), so
// "nothing T x T reaches HBM" holds for reads only; keeping writes on chip needs the triangle in shared memory / DSMEM.
cudaError_t launch_tile(const Params& P_in, bool backward, cudaStream_t st) {
  Params P = P_in;
  const size_t smem = tile_smem_bytes(P.d);
  if (!P.scratch || P.scratch_stride < tile_slot_floats(P.d)) return cudaErrorInvalidValue;
  const int npairs = P.d.B * P.d.D;
  const int grid = npairs < kNumSMs ? npairs : kNumSMs;
  // compact slots: this tier's own stride inside the block tier's slot area (forward needs the L triangle only)
  const TLay L(P.d.T_max, P.d.S);
  P.scratch_stride = backward ? tile_slot_floats(P.d) : (size_t)L.ntri() * TF;
  if (!backward) {  // the forward pass reads float64 records of its own (the block tier's pre-pass is not launched for it)
    const cudaError_t pe = launch_prior_inv64(P, st);
    if (pe != cudaSuccess) return pe;
  }
  static const int env_dc = [] { const char* e = getenv("GPKL_DIAG_CHUNKS"); return e ? atoi(e) : 0; }();
  if (env_dc > 0) cudaMemcpyToSymbolAsync(g_diag_chunks, &env_dc, sizeof(int), 0, cudaMemcpyHostToDevice, st);
  void (*kern)(Params);
  if (P.d.kernel == GPKL_KERNEL_RBF) kern = backward ? bwd_tile<GPKL_KERNEL_RBF> : fwd_tile<GPKL_KERNEL_RBF>;
  else kern = backward ? bwd_tile<GPKL_KERNEL_CAUCHY> : fwd_tile<GPKL_KERNEL_CAUCHY>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<grid, NTHR, smem, st>>>(P);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gpkl
