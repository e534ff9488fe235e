// The 64 x 64 diagonal block of the tile tier (gpkl_tile.cu): Cholesky factor AND inverse in one sweep by ONE warp pair
// (64 threads, thread t = row t).
//
// The block is the serial chain of every 64-column panel, so what counts is its latency, not its flops (2 x 64^3/6 FMAs).
// The first version (16-column sub-panels: one warp's shuffle Cholesky of a 16 x 16 block, per-thread substitutions, 4 x 4
// trailing tiles, three barriers per sub-panel, then a separate inversion) took 29-35 K cycles per block.  Two things were
// measured with tools/micro/diag64_bench.cu on the way here:
//   * fully unrolled register code (4.5 K instructions) is no faster: ONE warp running straight-line code misses the
//     instruction cache on every 128-byte line (~40 cycles each; the inverse took 12.6 K cycles cold against 4.6 K warm);
//   * a column-at-a-time elimination pays a shared-memory round trip per pivot (publish -> __syncwarp -> load -> MUFU.RCP
//     -> Newton -> scale: ~190 cycles per column whatever the amount of update work).
// So:
//   * BLOCKS OF 8 COLUMNS, one register window per thread that is shifted by 8 after each block: the code of one block is
//     fetched once and executed 8 times (the trailing update is a run-time loop over the block's 8 columns, 48 instructions).
//   * the 8 x 8 diagonal sub-block is broadcast through shared memory ONCE per block and every thread factors it
//     REDUNDANTLY in registers (LDL^T-style: pivots d_j, unscaled columns, multipliers through 1/d_j = MUFU.RCP + one Newton
//     step; the square roots are off the chain and applied when finished entries are written out), then eliminates its own
//     row's 8 entries locally: two shared-memory round trips per 8 columns instead of one per column.
//   * THE INVERSE RIDES ALONG: the same elimination applied to the rows of an identity appended below the block turns them
//     into L^-T (row t of it = column t of L^-1 -- exactly the operand layout the callers want).  Row t of the block is
//     finished at column t and row t of the identity is zero before column t, so each thread needs ONE window: its matrix
//     row up to the 8-column block holding its diagonal entry, its identity row from there on.
//   * rows leave the registers through XOR-swizzled scratch tiles (a thread writes 128-bit pieces of its own row; plain row
//     pitch 64 would be an 8-way bank conflict); the caller compacts them in one conflict-free pass.
// Reference step replaced: the pivot loop inside tf.cholesky (src/Models/Full_GP_VAE_dynamic_time.py:165) and the
// tf.matrix_inverse of :250 restricted to one diagonal block.
#pragma once
#include "gpkl_common.cuh"

namespace gpkl {

#ifndef GPKL_D64_TICK
#define GPKL_D64_TICK(k)
#endif

__device__ __forceinline__ float rcp_newton(float d) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return fmaf(r, fmaf(-d, r, 1.0f), r);
}

// float4 group g of row i of a swizzled 64 x 64 scratch tile
__device__ __forceinline__ int swz64(int i, int g) { return i * 64 + 4 * (g ^ (i & 15)); }

// 64 threads (t = 0..63), synchronised by sync() (a barrier over exactly these threads).
// D: the block, column-major (D[c*64 + i]), lower part valid on entry; overwritten with the unscaled columns (scratch).
// Out: gt = L (global tile, column-major, zeros above the diagonal); Rsw = L row-major and Xsw = L^-T row-major (row t =
// column t of L^-1), both in the swizzled scratch layout; dgl[c] = L(c,c), rdl[c] = 1/L(c,c).  Dg: 64 floats and Ss: 512
// floats of scratch.  Tl: real rows of the block (non-positive pivots there set *bad).
template <class SyncF>
__device__ __forceinline__ void factor_invert64_rows(float* __restrict__ D, float* __restrict__ gt, float* __restrict__ dgl,
                                                     float* __restrict__ rdl, int Tl, int* bad, float* __restrict__ Rsw,
                                                     float* __restrict__ Xsw, float* __restrict__ Dg, float* __restrict__ Ss,
                                                     int t, SyncF sync) {
  float w[64];
#pragma unroll
  for (int k = 0; k < 64; ++k) w[k] = D[k * 64 + t];
  const int myblk = t >> 3, me = t & 7;
#pragma unroll 1
  for (int blk = 0; blk < 8; ++blk) {
    const bool own = myblk == blk, before = myblk > blk;
    // (1) the 8 x 8 diagonal sub-block to every thread
    if (own) {
      *reinterpret_cast<float4*>(Dg + me * 8) = make_float4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<float4*>(Dg + me * 8 + 4) = make_float4(w[4], w[5], w[6], w[7]);
    }
    sync();
    float g[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 lo = *reinterpret_cast<const float4*>(Dg + i * 8);
      g[i][0] = lo.x; g[i][1] = lo.y; g[i][2] = lo.z; g[i][3] = lo.w;
      if (i >= 4) {
        const float4 hi = *reinterpret_cast<const float4*>(Dg + i * 8 + 4);
        g[i][4] = hi.x; g[i][5] = hi.y; g[i][6] = hi.z; g[i][7] = hi.w;
      }
    }
    GPKL_D64_TICK(0)
    // (2) redundant LDL^T of the sub-block: g[i][j] (i > j) = unscaled column entries, dj / rj = pivots and reciprocals
    float dj[8], rj[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dj[j] = g[j][j];
      rj[j] = rcp_newton(dj[j]);
#pragma unroll
      for (int i = j + 1; i < 8; ++i) {
        const float m = -(g[i][j] * rj[j]);
#pragma unroll
        for (int k = j + 1; k <= i; ++k) g[i][k] = fmaf(m, g[k][j], g[i][k]);
      }
    }
    GPKL_D64_TICK(1)
    // (3) this thread's row through the block's 8 columns: its matrix row (before / own) or its identity row (after), and
    // for the owners the start of their identity row; multipliers for the trailing update
    float v[8], xo[8], sv[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      v[cc] = w[cc];
      xo[cc] = (cc == me) ? 1.0f : 0.0f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float mv = -(v[j] * rj[j]), mx = -(xo[j] * rj[j]);
      sv[j] = own ? mx : mv;
#pragma unroll
      for (int cc = j + 1; cc < 8; ++cc) {
        v[cc] = fmaf(mv, g[cc][j], v[cc]);
        xo[cc] = fmaf(mx, g[cc][j], xo[cc]);
      }
    }
    GPKL_D64_TICK(2)
    // (4) publish the block's unscaled columns (zeros above the diagonal) and the finished, SCALED entries
    float oa[8], ox[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc) {
      const int c = 8 * blk + cc;
      const bool lower = before || (own && cc <= me);
      D[c * 64 + t] = lower ? v[cc] : 0.0f;
      Ss[cc * 64 + t] = sv[cc];
      const float d = dj[cc];
      float rs = rsqrtf(d);
      rs = rs * fmaf(-0.5f * d, rs * rs, 1.5f);  // one Newton step: 1/sqrt(d) to ~1 ulp
      if (t == cc) {
        dgl[c] = d * rs;
        rdl[c] = rs;
        if (c < Tl && !(d > 0.0f)) *bad = 1;
      }
      oa[cc] = lower ? v[cc] * rs : 0.0f;
      ox[cc] = (before ? 0.0f : (own ? xo[cc] : v[cc])) * rs;
      gt[c * 64 + t] = oa[cc];
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      *reinterpret_cast<float4*>(Rsw + swz64(t, 2 * blk + h)) = make_float4(oa[4 * h], oa[4 * h + 1], oa[4 * h + 2], oa[4 * h + 3]);
      *reinterpret_cast<float4*>(Xsw + swz64(t, 2 * blk + h)) = make_float4(ox[4 * h], ox[4 * h + 1], ox[4 * h + 2], ox[4 * h + 3]);
    }
    GPKL_D64_TICK(3)
    sync();
    GPKL_D64_TICK(4)
    if (blk == 7) break;
    // (5) trailing update of the window (columns 8 blk + 8 ..): w[k] += sum_cc s_cc u_cc(8 blk + k).  The owners' window
    // becomes their identity row (zeros so far).  Window entries beyond column 63 read the next column's words: never used.
    if (own) {
#pragma unroll
      for (int k = 8; k < 64; ++k) w[k] = 0.0f;
    }
    const float* ub = D + (8 * blk) * 64 + 8 * blk;
#pragma unroll 1
    for (int cc = 0; cc < 8; ++cc) {
      const float s = Ss[cc * 64 + t];
      const float* up = ub + cc * 64;
#pragma unroll
      for (int gq = 2; gq < 16; ++gq) {
        if (gq >= 10 && blk >= 4) break;  // (uniform) the second half of the window is beyond column 63
        const float4 u4 = *reinterpret_cast<const float4*>(up + 4 * gq);
        fma2(w[4 * gq], w[4 * gq + 1], s, s, u4.x, u4.y);
        fma2(w[4 * gq + 2], w[4 * gq + 3], s, s, u4.z, u4.w);
      }
    }
    GPKL_D64_TICK(5)
#pragma unroll
    for (int k = 0; k + 8 < 64; ++k) w[k] = w[k + 8];
    GPKL_D64_TICK(6)
  }
}

}  // namespace gpkl
