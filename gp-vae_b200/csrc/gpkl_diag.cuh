// 16x16 diagonal-block primitives shared by the block tier (gpkl_block.cu) and the tile tier (gpkl_tile.cu):
// the register/shuffle Cholesky of one diagonal block by one warp and the 16-vector forward substitution.
// Reference step replaced: the pivot loop inside tf.cholesky, src/Models/Full_GP_VAE_dynamic_time.py:165.
#pragma once
#include "gpkl_common.cuh"

namespace gpkl {

// One warp: factor the 16x16 diagonal block held in pan (columns 0..15, rows j0..j0+15) in registers with
// shuffles; writes L_dd into the LC triangle of Bm, diag(L) into dg and 1/diag(L) into rdg.
// XRC: also write the block row-major into the XR triangle (element (i,k) at Bm[(i+1)*ld + k]).
template <bool XRC = false>
__device__ __forceinline__ void diag_factor(float* __restrict__ Bm, int ld, int j0, int T, const float* __restrict__ pan,
                                            int ldpan, float* __restrict__ dg, float* __restrict__ rdg, int* bad) {
  const int lane = threadIdx.x & 31, l = lane & 15;
  float a[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) a[c] = (c <= l) ? pan[c * ldpan + j0 + l] : 0.0f;
  float dgv = 1.0f, rdv = 1.0f;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float d = __shfl_sync(0xffffffffu, a[c], c, 16);
    float rs = rsqrtf(d);
    rs = rs * fmaf(-0.5f * d, rs * rs, 1.5f);  // one Newton step: 1/sqrt(d) to ~1 ulp
    const float sd = d * rs;
    if (j0 + c < T && !(d > 0.0f)) *bad = 1;
    const float lc = (l > c) ? a[c] * rs : ((l == c) ? sd : 0.0f);
    a[c] = lc;
    if (l == c) { dgv = sd; rdv = rs; }
#pragma unroll
    for (int k = c + 1; k < 16; ++k) {
      const float lk = __shfl_sync(0xffffffffu, lc, k, 16);
      a[k] = fmaf(-lc, lk, a[k]);
    }
  }
  if (lane < 16) {
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c <= l) {
        Bm[(size_t)(j0 + c) * ld + j0 + l] = a[c];
        if (XRC) Bm[(size_t)(j0 + l + 1) * ld + j0 + c] = a[c];
      }
    dg[j0 + l] = dgv;
    rdg[j0 + l] = rdv;
  }
}

// x <- L_dd^-1 b for one 16-vector held in registers (right-looking substitution); L_dd is the 16x16
// diagonal block at (d0,d0) of the LC triangle of Lb, read as 128-bit broadcasts; rdg = 1/diag(L).
__device__ __forceinline__ void diag_solve16(float (&b)[16], const float* __restrict__ Lb, int ld, int d0,
                                             const float* __restrict__ rdg) {
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    const float xc = b[c] * rdg[d0 + c];
    b[c] = xc;
    const float* col = Lb + (size_t)(d0 + c) * ld + d0;  // L[d0.., d0+c]
#pragma unroll
    for (int g = (c + 1) / 4; g < 4; ++g) {
      const float4 l4 = *reinterpret_cast<const float4*>(col + 4 * g);
      const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        if (4 * g + e > c) fma2(b[4 * g + e], b[4 * g + e + 1], -xc, -xc, lv[e], lv[e + 1]);
        else if (4 * g + e + 1 > c) b[4 * g + e + 1] = fmaf(-xc, lv[e + 1], b[4 * g + e + 1]);
      }
    }
  }
}

}  // namespace gpkl
