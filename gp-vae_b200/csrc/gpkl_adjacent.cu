// Rows either side of the GP-prior path: streaming, HBM-bound kernels.
//
// (1) GP-recognition sampler epilogue -- src/Models/GP_recog_VAE_prior.py: tf_kernel_approx (:137-168) samples
//     z = m + (chol(K) + diag(sqrt(exp(logvar)))) eps and the model uses the standard N(0,I) KL per time row
//     (standard_vae_kl :65-70, negated at :274-276).  The chol(K) eps part IS the sample of the fused GP op (same L_q), so
//     these kernels only add the diagonal term, evaluate the row KL and, backward, the logvar / mean adjoints.
//     One warp per row of mean [total_T, D]: lanes stride over the latent dims, so mean / logvar / z / g_z are
//     read and written in full 128-byte lines.
// (2) Ragged batch producer -- src/Models/DataHandler.py:129-156 (_prep_dataset) + :111-127 (data_batch): from the
//     -1-masked array [N, F, T_full] and the shared time grid, the packed rows x [sum_T, F], the zero-padded time
//     stamps [B, max_time] and the lengths [B] of one batch.  A warp per sequence compacts the valid time points
//     (ballot + popc) for lengths and time stamps; after the scan of the lengths a warp per (sequence, 32 features)
//     transposes [F, T_full] -> [T_b, F] through its own shared-memory tile so that both the read (along t) and the
//     write (along the packed rows) are coalesced.
#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {
namespace {

constexpr int RG_THREADS = 256;

// sequence owning row r of mean:  off[b] <= r < off[b+1]
__device__ __forceinline__ int seq_of_row(long long r, const long long* __restrict__ off, int B) {
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= r) lo = mid;
    else hi = mid;
  }
  return lo;
}

// z += sqrt(v) * eps ;  kl_rows[r] = -1/2 sum_d (1 + log(1e-10 + v) - m^2 - v),  v = exp(logvar)
__global__ void __launch_bounds__(RG_THREADS) recog_fwd_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                                               const float* __restrict__ eps, const long long* __restrict__ off,
                                                               int B, int D, int S, int T_max, long long rows,
                                                               float* __restrict__ z, float* __restrict__ kl_rows) {
  const int lane = threadIdx.x & 31, wpb = RG_THREADS / 32;
  const long long nwarps = (long long)gridDim.x * wpb;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += nwarps) {
    const int b = seq_of_row(r, off, B);
    const long long r0 = off[b];
    const int T = (int)(off[b + 1] - r0), i = (int)(r - r0);
    double acc = 0.0;
    for (int d = lane; d < D; d += 32) {
      const float m = mean[r * D + d], lv = logvar[r * D + d];
      const float v = expf(lv);
      const float sd = sqrtf(v);
      acc += (double)(1.0f + logf(1e-10f + v) - m * m - v);
      const float* __restrict__ e = eps + ((size_t)b * D + d) * S * T_max + i;
      for (int s = 0; s < S; ++s) {
        float* zp = z + ((size_t)S * r0 + (size_t)s * T + i) * D + d;
        *zp = fmaf(sd, e[(size_t)s * T_max], *zp);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) kl_rows[r] = (float)(-0.5 * acc);
  }
}

// g_mean += g_r m ;  g_logvar = 1/2 sqrt(v) sum_s g_z eps  -  g_r/2 (v/(1e-10+v) - v),   g_r = g_kl_sum + g_kl_rows[r]
__global__ void __launch_bounds__(RG_THREADS) recog_bwd_kernel(const float* __restrict__ mean, const float* __restrict__ logvar,
                                                               const float* __restrict__ eps, const float* __restrict__ g_z,
                                                               const double* __restrict__ g_kl_sum,
                                                               const float* __restrict__ g_kl_rows,
                                                               const long long* __restrict__ off, int B, int D, int S,
                                                               int T_max, long long rows, float* __restrict__ g_mean,
                                                               float* __restrict__ g_logvar) {
  const int lane = threadIdx.x & 31, wpb = RG_THREADS / 32;
  const long long nwarps = (long long)gridDim.x * wpb;
  const double gs = g_kl_sum ? *g_kl_sum : 1.0;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += nwarps) {
    const int b = seq_of_row(r, off, B);
    const long long r0 = off[b];
    const int T = (int)(off[b + 1] - r0), i = (int)(r - r0);
    const float g = (float)(gs + (g_kl_rows ? (double)g_kl_rows[r] : 0.0));
    for (int d = lane; d < D; d += 32) {
      const float m = mean[r * D + d], lv = logvar[r * D + d];
      const float v = expf(lv);
      const float sd = sqrtf(v);
      float ge = 0.0f;
      if (g_z) {
        const float* __restrict__ e = eps + ((size_t)b * D + d) * S * T_max + i;
        for (int s = 0; s < S; ++s)
          ge = fmaf(g_z[((size_t)S * r0 + (size_t)s * T + i) * D + d], e[(size_t)s * T_max], ge);
      }
      g_mean[r * D + d] = fmaf(g, m, g_mean[r * D + d]);
      g_logvar[r * D + d] = 0.5f * sd * ge - 0.5f * g * (__fdiv_rn(v, 1e-10f + v) - v);
    }
  }
}

// ---- ragged batch producer ----------------------------------------------------------------------------------------
// Pass 1, one warp per sequence of the batch: valid(t) = data[seq, 0, t] > -1  (DataHandler.py:143); writes lengths[b]
// and times[b, k] = grid[t_k] zero padded to max_time (:149-151).
__global__ void __launch_bounds__(RG_THREADS) collate_scan_kernel(const float* __restrict__ data, const float* __restrict__ grid,
                                                                  const int32_t* __restrict__ index, int B, int F, int T_full,
                                                                  int max_time, int32_t* __restrict__ lengths,
                                                                  float* __restrict__ times) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (RG_THREADS / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  const size_t seq = index ? (size_t)index[b] : (size_t)b;
  const float* __restrict__ row0 = data + seq * F * T_full;
  int count = 0;
  for (int t0 = 0; t0 < T_full; t0 += 32) {
    const int t = t0 + lane;
    const bool ok = t < T_full && row0[t] > -1.0f;
    const unsigned mask = __ballot_sync(0xffffffffu, ok);
    const int k = count + __popc(mask & ((1u << lane) - 1u));
    if (ok && k < max_time) times[(size_t)b * max_time + k] = grid[t];
    count += __popc(mask);
  }
  if (count > max_time) count = max_time;
  for (int k = count + lane; k < max_time; k += 32) times[(size_t)b * max_time + k] = 0.0f;
  if (lane == 0) lengths[b] = count;
}

// Pass 2, one warp per (sequence, chunk of CF features), no CTA-wide barrier: for every 32 time steps the warp rebuilds
// the validity mask from feature 0 (one coalesced load), loads the CF x 32 source values with lanes along t (coalesced,
// independent of the mask, so all of them are in flight together), parks the kept ones in its own shared-memory tile at
// their compacted row, and writes the tile out with lanes along the packed [row, feature] order -- for F <= CF the
// destination span x[off[b]+k0 .. , :] is one contiguous run (:145 "reshape(...,[15,-1]).T").
constexpr int CF = 32;
__global__ void __launch_bounds__(RG_THREADS) collate_gather_kernel(const float* __restrict__ data, const int32_t* __restrict__ index,
                                                                    const long long* __restrict__ off, int B, int F, int T_full,
                                                                    int max_time, float* __restrict__ x) {
  __shared__ float tiles[RG_THREADS / 32][32][CF + 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = RG_THREADS / 32;
  float (*tile)[CF + 1] = tiles[wid];
  const int fchunks = (F + CF - 1) / CF;
  const long long nitems = (long long)B * fchunks;
  for (long long it = (long long)blockIdx.x * wpb + wid; it < nitems; it += (long long)gridDim.x * wpb) {
    const int b = (int)(it / fchunks);
    const int f0 = (int)(it - (long long)b * fchunks) * CF;
    const int fc = min(CF, F - f0);
    const size_t seq = index ? (size_t)index[b] : (size_t)b;
    const float* __restrict__ src = data + seq * F * T_full;
    float* __restrict__ dst = x + (size_t)off[b] * F + f0;
    int count = 0;
    for (int t0 = 0; t0 < T_full && count < max_time; t0 += 32) {
      const int t = t0 + lane;
      const bool ok = t < T_full && src[t] > -1.0f;
      const unsigned mask = __ballot_sync(0xffffffffu, ok);
      const int kl = __popc(mask & ((1u << lane) - 1u));  // row of this lane's time step inside the tile
      const int n = min(__popc(mask), max_time - count);
      const bool keep = ok && kl < n;
#pragma unroll 8
      for (int f = 0; f < fc; ++f) {
        const float v = t < T_full ? src[(size_t)(f0 + f) * T_full + t] : 0.0f;
        if (keep) tile[kl][f] = v;
      }
      __syncwarp();
      if (fc == F) {  // whole rows: one contiguous run of n*F floats
        for (int e = lane; e < n * F; e += 32) dst[(size_t)count * F + e] = tile[e / F][e - (e / F) * F];
      } else {
        for (int r = 0; r < n; ++r)
          if (lane < fc) dst[(size_t)(count + r) * F + lane] = tile[r][lane];
      }
      __syncwarp();
      count += n;
    }
  }
}

int stream_grid(long long warps) {
  const long long want = (warps + (RG_THREADS / 32) - 1) / (RG_THREADS / 32);
  const long long cap = (long long)kNumSMs * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace

cudaError_t launch_recog_fwd(const float* mean, const float* logvar, const float* eps, const long long* off, int B, int D,
                             int S, int T_max, long long rows, float* z, float* kl_rows, cudaStream_t st) {
  recog_fwd_kernel<<<stream_grid(rows), RG_THREADS, 0, st>>>(mean, logvar, eps, off, B, D, S, T_max, rows, z, kl_rows);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_recog_bwd(const float* mean, const float* logvar, const float* eps, const float* g_z,
                             const double* g_kl_sum, const float* g_kl_rows, const long long* off, int B, int D, int S,
                             int T_max, long long rows, float* g_mean, float* g_logvar, cudaStream_t st) {
  recog_bwd_kernel<<<stream_grid(rows), RG_THREADS, 0, st>>>(mean, logvar, eps, g_z, g_kl_sum, g_kl_rows, off, B, D, S,
                                                             T_max, rows, g_mean, g_logvar);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_collate_scan(const float* data, const float* grid, const int32_t* index, int B, int F, int T_full,
                                int max_time, int32_t* lengths, float* times, cudaStream_t st) {
  const int wpb = RG_THREADS / 32;
  collate_scan_kernel<<<(B + wpb - 1) / wpb, RG_THREADS, 0, st>>>(data, grid, index, B, F, T_full, max_time, lengths, times);
  note_launch();
  return cudaGetLastError();
}

cudaError_t launch_collate_gather(const float* data, const int32_t* index, const long long* off, int B, int F, int T_full,
                                  int max_time, float* x, cudaStream_t st) {
  const long long items = (long long)B * ((F + CF - 1) / CF);
  const long long want = (items + (RG_THREADS / 32) - 1) / (RG_THREADS / 32), cap = (long long)kNumSMs * 6;  // 6 CTAs/SM by shared memory
  const int grid = (int)(want < cap ? (want > 0 ? want : 1) : cap);
  collate_gather_kernel<<<grid, RG_THREADS, 0, st>>>(data, index, off, B, F, T_full, max_time, x);
  note_launch();
  return cudaGetLastError();
}

}  // namespace gpkl
