// Developer microbenchmark: latency of the 64x64 diagonal-block routines (gpkl_diag64.cuh), cold vs warm instruction cache.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I gp-vae_b200/csrc -o /tmp/diag64_bench tools/micro/diag64_bench.cu
#include <cstdio>
#include <vector>
#include <cmath>
__device__ long long g_tk[8]; __device__ long long g_last;
#define GPKL_D64_TICK(k) if (t == 0) { long long now = clock64(); g_tk[k] += now - g_last; g_last = now; }
#include "gpkl_diag64.cuh"
using namespace gpkl;

__global__ void __launch_bounds__(64, 1) k(const float* K, float* gt, float* out, long long* clk, int reps) {
  extern __shared__ __align__(16) float sm[];
  float* D = sm; float* Rsw = sm + 4096; float* Xsw = sm + 8192; float* dgl = sm + 12288; float* rdl = dgl + 64;
  __shared__ int bad;
  const int t = threadIdx.x;
  for (int r = 0; r < reps; ++r) {
    for (int e = t; e < 4096; e += 64) D[e] = K[e];
    __syncthreads();
    long long t0 = clock64();
    if (t == 0) g_last = t0;
    factor_invert64_rows(D, gt, dgl, rdl, 64, &bad, Rsw, Xsw, dgl + 128, dgl + 256, t, [] { __syncthreads(); });
    __syncthreads();
    long long t1 = clock64();
    long long t2 = 0, t3 = t1;
    if (t == 0) { clk[3 * r] = t1 - t0; clk[3 * r + 1] = t2; clk[3 * r + 2] = t3 - t2; }
  }
  for (int e = t; e < 4096; e += 64) { out[e] = gt[e]; out[4096 + e] = Xsw[e]; out[8192 + e] = Rsw[e]; }
}

int main() {
  std::vector<float> K(4096);
  for (int c = 0; c < 64; ++c) for (int i = 0; i < 64; ++i) { float d = float(i - c); K[c * 64 + i] = 0.999f / (1.0f + d * d) + (i == c ? 0.001f : 0.0f); }
  float *dK, *dout, *dgt; long long* dclk; const int reps = 6;
  cudaMalloc(&dK, 4096 * 4); cudaMalloc(&dgt, 4096 * 4); cudaMalloc(&dout, 3 * 4096 * 4); cudaMalloc(&dclk, reps * 3 * 8);
  cudaMemcpy(dK, K.data(), 4096 * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  k<<<1, 64, 60000>>>(dK, dgt, dout, dclk, reps);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  std::vector<long long> clk(reps * 3); std::vector<float> out(3 * 4096);
  long long tk[8]; cudaMemcpyFromSymbol(tk, g_tk, sizeof(tk)); for (int i = 0; i < 7; ++i) printf("tick %d: %lld per block\n", i, tk[i] / (reps * 8));
  cudaMemcpy(clk.data(), dclk, reps * 3 * 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(out.data(), dout, 3 * 4096 * 4, cudaMemcpyDeviceToHost);
  for (int r = 0; r < reps; ++r) printf("rep %d: factor+inverse %lld cycles (%lld %lld)\n", r, clk[3 * r], clk[3 * r + 1], clk[3 * r + 2]);
  // check: L L^T = K (lower), X L = I
  double e1 = 0, e2 = 0;
  for (int i = 0; i < 64; ++i) for (int c = 0; c <= i; ++c) {
    double s = 0; for (int k2 = 0; k2 <= c; ++k2) s += (double)out[k2 * 64 + i] * out[k2 * 64 + c];
    e1 = fmax(e1, fabs(s - K[c * 64 + i]));
  }
  // Xsw swizzled: row t (= column t of Linv), group g at t*64 + 4*(g ^ (t&15))
  for (int t2 = 0; t2 < 64; ++t2) for (int i = 0; i < 64; ++i) {
    double s = 0;
    for (int k2 = 0; k2 < 64; ++k2) { // sum_k L(i,k) Linv(k,t)
      const float xv = out[4096 + t2 * 64 + 4 * ((k2 >> 2) ^ (t2 & 15)) + (k2 & 3)];
      const float lv = (k2 <= i) ? out[k2 * 64 + i] : 0.0f;
      s += (double)lv * xv;
    }
    e2 = fmax(e2, fabs(s - (i == t2 ? 1.0 : 0.0)));
  }
  printf("max |LL^T - K| = %.3g, max |L Linv - I| = %.3g\n", e1, e2);
  return 0;
}
