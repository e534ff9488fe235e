// C ABI of libgpkl.so (include/gpkl.h): argument checking, workspace carving, tier dispatch, the small
// prologue (row offsets) and epilogue (deterministic reductions) kernels, and the host-buffer step.
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "gpkl_common.cuh"
#include "gpkl_launch.h"

namespace gpkl {

// ---- measurement hooks ---------------------------------------------------------------------------
long long* g_dbg = nullptr;  // device buffer for the phase trace (internal; set by gpkl_debug_set_trace)
namespace {
constexpr int kProfRing = 1024;
struct ProfState {
  bool on = false;
  cudaEvent_t ev[2][kProfRing][2];
  bool created = false;
  int n[2] = {0, 0};
};
ProfState g_prof;
std::atomic<long long> g_launches{0};  // kernel launches issued by this library (bench.py: gpu_launches)
}  // namespace

void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("GPKL_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
void prof_begin(bool backward, cudaStream_t st) {
  if (!g_prof.on || g_prof.n[backward] >= kProfRing) return;
  cudaEventRecord(g_prof.ev[backward][g_prof.n[backward]][0], st);
}
void prof_end(bool backward, cudaStream_t st) {
  if (!g_prof.on || g_prof.n[backward] >= kProfRing) return;
  cudaEventRecord(g_prof.ev[backward][g_prof.n[backward]][1], st);
  g_prof.n[backward]++;
}

namespace {

// Independent FFMA chains: 16 accumulators per thread, 2 flops per FMA.
__global__ void __launch_bounds__(256) fp32_peak_kernel(float* __restrict__ sink, int iters) {
  float a[16];
  const float x = 1.0f + 1e-7f * threadIdx.x, y = 1e-9f * (blockIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 0.001f * (i + threadIdx.x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

constexpr size_t kAlign = 256;
inline size_t align_up(size_t x) { return (x + kAlign - 1) / kAlign * kAlign; }

// offsets[b] = sum_{b' < b} lengths[b'] ; single CTA, two-level scan.
// Also the shared-prior decision: *prior_flag = 1 iff ell_p[0] == ... == ell_p[D-1] (prior_flag may be NULL).
__global__ void offsets_kernel(const int32_t* __restrict__ lengths, int B, int64_t* __restrict__ offsets,
                               const float* __restrict__ ell_p = nullptr, int D = 0, int32_t* __restrict__ prior_flag = nullptr) {
  __shared__ int64_t warp_tot[32];
  const int nt = blockDim.x, tid = threadIdx.x;
  if (prior_flag) {
    bool uni = true;
    const float l0 = ell_p[0];
    for (int i = tid; i < D; i += nt) uni = uni && (ell_p[i] == l0);
    const int all = __syncthreads_and(uni ? 1 : 0);
    if (tid == 0) *prior_flag = all ? 1 : 0;
  }
  const int chunk = (B + nt - 1) / nt;
  const int lo = min(B, tid * chunk), hi = min(B, lo + chunk);
  int64_t local = 0;
  for (int i = lo; i < hi; ++i) local += max(lengths[i], 0);
  // inclusive scan of `local` across the block
  int64_t v = local;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  if (lane == 31) warp_tot[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int64_t w = (lane < (nt >> 5)) ? warp_tot[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t n = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += n;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  int64_t run = v - local + (warp > 0 ? warp_tot[warp - 1] : 0);
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += max(lengths[i], 0);
  }
  if (tid == nt - 1) offsets[B] = warp_tot[(nt >> 5) - 1];
}

// kl_sum = sum_p kl_pairs[p] in float64, fixed order (Full_GP_VAE_dynamic_time.py:228).
__global__ void sum_pairs_kernel(const float* __restrict__ x, int n, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {  // 128-bit loads, two in flight per thread (the loop is a latency chain)
    const float4* __restrict__ x4 = reinterpret_cast<const float4*>(x);
    const int n4 = n >> 2, bd = blockDim.x;
    for (int i = threadIdx.x; i < n4; i += 2 * bd) {
      const float4 a = x4[i];
      const float4 b = i + bd < n4 ? x4[i + bd] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
      acc += ((double)a.x + (double)a.y) + ((double)a.z + (double)a.w);
      acc += ((double)b.x + (double)b.y) + ((double)b.z + (double)b.w);
    }
    for (int i = 4 * n4 + threadIdx.x; i < n; i += bd) acc += (double)x[i];
  } else {
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)x[i];
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) *out = acc;
}

// g_ell[d] = sum_b pairs[b*D + d]; one warp per d, fixed order.
__global__ void sum_over_batch_kernel(const float* __restrict__ pairs, int B, int D, float* __restrict__ out) {
  const int d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (d >= D) return;
  double acc = 0.0;
  for (int b = threadIdx.x & 31; b < B; b += 32) acc += (double)pairs[(size_t)b * D + d];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) out[d] = (float)acc;
}

struct Workspace {
  int64_t* offsets;
  float* gq_pairs;
  float* gp_pairs;
  float* scratch;
  size_t scratch_stride;
  int32_t* prior_flag;  // shared-prior fast path: device flag + per-sequence records (NULL when not eligible)
  float* prior;
  size_t prior_stride;
  size_t total;
};

// The shared-prior fast path applies to the GP posterior when no d/d ell_p is requested (that goes to the generic
// tier) and the caller did not force the per-pair factorisation.
bool prior_sharing_eligible(const GpklDesc& d) {
  if (d.flags & (GPKL_FLAG_GRAD_ELL_P | GPKL_FLAG_PER_PAIR_PRIOR)) return false;
  if (d.tier == GPKL_TIER_GENERIC || d.B <= 0) return false;
  if (d.posterior == GPKL_POST_BIDIAG) return d.tier != GPKL_TIER_BLOCK && bidiag_tier_supports(d);  // V3 hot tier (T <= 64)
  return d.posterior == GPKL_POST_GP;
}

Workspace plan(const GpklDesc& d, void* base) {
  Workspace w;
  const size_t P = (size_t)d.B * d.D;
  size_t off = 0;
  unsigned char* b = static_cast<unsigned char*>(base);
  w.offsets = reinterpret_cast<int64_t*>(b + off);
  off += align_up(((size_t)d.B + 1) * sizeof(int64_t));
  w.gq_pairs = reinterpret_cast<float*>(b + off);
  off += align_up(P * sizeof(float));
  w.gp_pairs = reinterpret_cast<float*>(b + off);
  off += align_up(P * sizeof(float));
  // per-CTA matrix slots for sizes whose work matrices do not fit shared memory; sized for whichever
  // tier (block: kBlockSlots CTAs, generic: 148 CTAs) may serve this descriptor
  w.scratch = reinterpret_cast<float*>(b + off);
  const size_t gs = generic_slots(d) ? generic_slot_floats(d.T_max) : 0, bs = block_slot_floats(d);
  w.scratch_stride = gs > bs ? gs : bs;
  const size_t nslots = w.scratch_stride ? (size_t)(kBlockSlots > generic_slots(d) ? kBlockSlots : generic_slots(d)) : 0;
  off += align_up(nslots * w.scratch_stride * sizeof(float));
  w.prior_flag = nullptr;
  w.prior = nullptr;
  w.prior_stride = 0;
  if (prior_sharing_eligible(d)) {
    w.prior_flag = reinterpret_cast<int32_t*>(b + off);
    off += align_up(sizeof(int32_t));
    w.prior = reinterpret_cast<float*>(b + off);
    w.prior_stride = prior_record_floats(d.T_max);
    off += align_up((size_t)d.B * w.prior_stride * sizeof(float));
  }
  w.total = off;
  return w;
}

int check_desc(const GpklDesc* d) {
  if (!d) return GPKL_ERR_NULL;
  if (d->B < 0 || d->D <= 0 || d->T_max < 0 || d->S < 1 || d->total_T < 0) return GPKL_ERR_DESC;
  if (d->kernel != GPKL_KERNEL_RBF && d->kernel != GPKL_KERNEL_CAUCHY) return GPKL_ERR_DESC;
  if (d->posterior != GPKL_POST_GP && d->posterior != GPKL_POST_DIAG && d->posterior != GPKL_POST_BIDIAG)
    return GPKL_ERR_DESC;
  if (d->tier < GPKL_TIER_AUTO || d->tier > GPKL_TIER_BLOCK) return GPKL_ERR_DESC;
  if (!(d->noise >= 0.0f) || !(d->noise < 1.0f)) return GPKL_ERR_DESC;
  // V3 (bidiagonal-precision posterior) has no d/d ell_p path
  if (d->posterior == GPKL_POST_BIDIAG && (d->flags & GPKL_FLAG_GRAD_ELL_P)) return GPKL_ERR_UNSUPPORTED;
  if ((int64_t)d->B * d->D > (int64_t)1 << 30) return GPKL_ERR_DESC;
  return GPKL_OK;
}

int dispatch(const Params& P, bool backward, cudaStream_t st) {
  // tier selection (GPKL_TIER_AUTO): register-resident warp tier for T <= 64, block tier above (matrices in
  // shared memory up to T ~ 144, in an L2-backed workspace slot beyond), generic tier for the combinations
  // the specialised tiers do not implement (d/d ell_p).  Explicit requests are honoured or refused, never silently rerouted.
  cudaError_t e;
  if (P.d.posterior == GPKL_POST_BIDIAG && P.prior != nullptr &&
      (P.d.tier == GPKL_TIER_AUTO || P.d.tier == GPKL_TIER_WARP)) {
    // V3 hot tier (T <= 64, one warp per pair, O(T^2)); if the device finds ell_p non-uniform it returns at once and the
    // generic tier launched behind it does the work (and returns at once otherwise)
    e = launch_bidiag(P, backward, st);
    if (e == cudaSuccess) {
      Params Q = P;
      Q.skip_if_shared = 1;
      e = launch_generic(Q, backward, st);
    }
  } else if (P.d.tier == GPKL_TIER_WARP) {
    if (!warp_tier_supports(P.d, backward)) return GPKL_ERR_UNSUPPORTED;
    e = launch_warp(P, backward, st);
  } else if (P.d.tier == GPKL_TIER_BLOCK) {
    if (!block_tier_supports(P.d, backward)) return GPKL_ERR_UNSUPPORTED;
    e = launch_block(P, backward, st);
  } else if (P.d.tier == GPKL_TIER_AUTO && P.d.T_max <= 64 && warp_tier_supports(P.d, backward)) {
    e = launch_warp(P, backward, st);
  } else if (P.d.tier == GPKL_TIER_AUTO && block_tier_supports(P.d, backward)) {
    e = launch_block(P, backward, st);
  } else {
    e = launch_generic(P, backward, st);
  }
  return e == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

}  // namespace
}  // namespace gpkl

using namespace gpkl;

extern "C" int gpkl_version(void) { return GPKL_VERSION; }

extern "C" const char* gpkl_strerror(int code) {
  switch (code) {
    case GPKL_OK: return "ok";
    case GPKL_ERR_NULL: return "gpkl: a required pointer is NULL";
    case GPKL_ERR_DESC: return "gpkl: inconsistent descriptor";
    case GPKL_ERR_UNSUPPORTED: return "gpkl: combination not supported";
    case GPKL_ERR_WORKSPACE: return "gpkl: workspace too small (see gpkl_workspace_bytes)";
    case GPKL_ERR_CUDA: return "gpkl: CUDA launch failed";
    default: return "gpkl: unknown error code";
  }
}

extern "C" size_t gpkl_workspace_bytes(const GpklDesc* desc) {
  if (check_desc(desc) != GPKL_OK) return 0;
  return plan(*desc, nullptr).total;
}

extern "C" int gpkl_forward(const GpklDesc* desc, const float* mean, const float* times, const int32_t* lengths,
                            const float* ell_q, const float* ell_p, const float* aux, const float* eps, float* z,
                            float* kl_pairs, double* kl_sum, float* logdets, int32_t* status, void* workspace,
                            size_t ws_bytes, void* stream) {
  int rc = check_desc(desc);
  if (rc != GPKL_OK) return rc;
  const GpklDesc& d = *desc;
  if (d.B == 0) {  // empty batch: only the scalar is defined
    if (!kl_sum) return GPKL_ERR_NULL;
    cudaMemsetAsync(kl_sum, 0, sizeof(double), static_cast<cudaStream_t>(stream));
    if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), static_cast<cudaStream_t>(stream));
    return GPKL_OK;
  }
  if (!mean || !times || !lengths || !ell_p || !eps || !z || !kl_pairs || !kl_sum || !workspace) return GPKL_ERR_NULL;
  if (d.posterior == GPKL_POST_GP && !ell_q) return GPKL_ERR_NULL;
  if (d.posterior != GPKL_POST_GP && !aux) return GPKL_ERR_NULL;
  const Workspace w = plan(d, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  if (d.B == 0) {
    cudaMemsetAsync(kl_sum, 0, sizeof(double), st);
    return GPKL_OK;
  }
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, d.B, w.offsets, ell_p, d.D, w.prior_flag);
  note_launch();
  Params P;
  memset(&P, 0, sizeof(P));
  P.d = d;
  P.mean = mean; P.times = times; P.lengths = lengths; P.ell_q = ell_q; P.ell_p = ell_p; P.aux = aux; P.eps = eps;
  if (d.flags & GPKL_FLAG_PHILOX_EPS) { P.eps = nullptr; P.eps_seed = reinterpret_cast<const unsigned long long*>(eps); }
  P.z = z; P.kl_pairs = kl_pairs; P.logdets = logdets; P.status = status;
  P.offsets = w.offsets;
  P.scratch = w.scratch_stride ? w.scratch : nullptr;
  P.scratch_stride = w.scratch_stride;
  P.prior = w.prior; P.prior_stride = w.prior_stride; P.prior_flag = w.prior_flag;
  P.dbg = g_dbg;
  rc = dispatch(P, false, st);
  if (rc != GPKL_OK) return rc;
  sum_pairs_kernel<<<1, 1024, 0, st>>>(kl_pairs, d.B * d.D, kl_sum);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

extern "C" int gpkl_backward(const GpklDesc* desc, const float* mean, const float* times, const int32_t* lengths,
                             const float* ell_q, const float* ell_p, const float* aux, const float* eps,
                             const float* g_z, const double* g_kl_sum, const float* g_kl_pairs, float* g_mean,
                             float* g_ell_q, float* g_ell_p, float* g_aux, int32_t* status, void* workspace,
                             size_t ws_bytes, void* stream) {
  int rc = check_desc(desc);
  if (rc != GPKL_OK) return rc;
  const GpklDesc& d = *desc;
  const bool want_lp = (d.flags & GPKL_FLAG_GRAD_ELL_P) != 0;
  if (d.B == 0) {
    cudaStream_t st0 = static_cast<cudaStream_t>(stream);
    if (g_ell_q) cudaMemsetAsync(g_ell_q, 0, sizeof(float) * d.D, st0);
    if (want_lp && g_ell_p) cudaMemsetAsync(g_ell_p, 0, sizeof(float) * d.D, st0);
    if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), st0);
    return GPKL_OK;
  }
  if (!mean || !times || !lengths || !ell_p || !eps || !g_mean || !workspace) return GPKL_ERR_NULL;
  if (d.posterior == GPKL_POST_GP && (!ell_q || !g_ell_q)) return GPKL_ERR_NULL;
  if (d.posterior != GPKL_POST_GP && (!aux || !g_aux)) return GPKL_ERR_NULL;
  if (want_lp && !g_ell_p) return GPKL_ERR_NULL;
  const Workspace w = plan(d, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  if (d.B == 0) {
    if (g_ell_q) cudaMemsetAsync(g_ell_q, 0, sizeof(float) * d.D, st);
    if (want_lp) cudaMemsetAsync(g_ell_p, 0, sizeof(float) * d.D, st);
    return GPKL_OK;
  }
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, d.B, w.offsets, ell_p, d.D, w.prior_flag);
  note_launch();
  Params P;
  memset(&P, 0, sizeof(P));
  P.d = d;
  P.mean = mean; P.times = times; P.lengths = lengths; P.ell_q = ell_q; P.ell_p = ell_p; P.aux = aux; P.eps = eps;
  if (d.flags & GPKL_FLAG_PHILOX_EPS) { P.eps = nullptr; P.eps_seed = reinterpret_cast<const unsigned long long*>(eps); }
  P.g_z = g_z; P.g_kl_sum = g_kl_sum; P.g_kl_pairs = g_kl_pairs;
  P.g_mean = g_mean; P.g_aux = g_aux; P.gq_pairs = w.gq_pairs; P.gp_pairs = w.gp_pairs; P.status = status;
  P.offsets = w.offsets;
  P.scratch = w.scratch_stride ? w.scratch : nullptr;
  P.scratch_stride = w.scratch_stride;
  P.prior = w.prior; P.prior_stride = w.prior_stride; P.prior_flag = w.prior_flag;
  P.dbg = g_dbg;
  rc = dispatch(P, true, st);
  if (rc != GPKL_OK) return rc;
  const int wpb = 8;
  if (d.posterior == GPKL_POST_GP) {
    sum_over_batch_kernel<<<(d.D + wpb - 1) / wpb, wpb * 32, 0, st>>>(w.gq_pairs, d.B, d.D, g_ell_q);
    note_launch();
  } else if (g_ell_q) {
    cudaMemsetAsync(g_ell_q, 0, sizeof(float) * d.D, st);
  }
  if (want_lp) {
    sum_over_batch_kernel<<<(d.D + wpb - 1) / wpb, wpb * 32, 0, st>>>(w.gp_pairs, d.B, d.D, g_ell_p);
    note_launch();
  }
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// internal, not declared in include/gpkl.h: device buffer (>= 64 int64) receiving CTA 0's phase clocks
extern "C" void gpkl_debug_set_trace(void* dev_buf) { g_dbg = static_cast<long long*>(dev_buf); }

extern "C" int64_t gpkl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int gpkl_profile_enable(int on) {
  if (on && !g_prof.created) {
    for (int k = 0; k < 2; ++k)
      for (int i = 0; i < kProfRing; ++i)
        for (int j = 0; j < 2; ++j)
          if (cudaEventCreate(&g_prof.ev[k][i][j]) != cudaSuccess) return GPKL_ERR_CUDA;
    g_prof.created = true;
  }
  g_prof.on = on != 0;
  g_prof.n[0] = g_prof.n[1] = 0;
  return GPKL_OK;
}

extern "C" int gpkl_profile_read(double* fwd_ms, int32_t* fwd_launches, double* bwd_ms, int32_t* bwd_launches) {
  double tot[2] = {0.0, 0.0};
  for (int k = 0; k < 2; ++k) {
    for (int i = 0; i < g_prof.n[k]; ++i) {
      float ms = 0.0f;
      if (cudaEventSynchronize(g_prof.ev[k][i][1]) != cudaSuccess) return GPKL_ERR_CUDA;
      if (cudaEventElapsedTime(&ms, g_prof.ev[k][i][0], g_prof.ev[k][i][1]) != cudaSuccess) return GPKL_ERR_CUDA;
      tot[k] += ms;
    }
  }
  if (fwd_ms) *fwd_ms = tot[0];
  if (fwd_launches) *fwd_launches = g_prof.n[0];
  if (bwd_ms) *bwd_ms = tot[1];
  if (bwd_launches) *bwd_launches = g_prof.n[1];
  g_prof.n[0] = g_prof.n[1] = 0;
  return GPKL_OK;
}

extern "C" int gpkl_fp32_peak_launch(float* sink, int32_t iters, double* flops, void* stream) {
  if (!sink || iters <= 0) return GPKL_ERR_NULL;
  const int grid = kNumSMs * 8, block = 256;
  fp32_peak_kernel<<<grid, block, 0, static_cast<cudaStream_t>(stream)>>>(sink, iters);
  note_launch();
  if (flops) *flops = (double)grid * block * (double)iters * 16.0 * 2.0;
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// ---- reconstruction term ------------------------------------------------------------------------
namespace {
struct ReconWs {
  int64_t* offsets;
  double* partials;
  size_t total;
};
ReconWs plan_recon(int B, void* base) {
  ReconWs w;
  unsigned char* b = static_cast<unsigned char*>(base);
  size_t off = 0;
  w.offsets = reinterpret_cast<int64_t*>(b + off);
  off += align_up(((size_t)B + 1) * sizeof(int64_t));
  w.partials = reinterpret_cast<double*>(b + off);
  off += align_up((size_t)kNumSMs * 8 * sizeof(double));
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t gpkl_recon_workspace_bytes(int32_t B) { return B < 0 ? 0 : plan_recon(B, nullptr).total; }

extern "C" int gpkl_recon_forward(int32_t B, int32_t F, int32_t S, int64_t total_T, const float* x, const float* x_decode,
                                  const int32_t* lengths, double* recon, void* workspace, size_t ws_bytes, void* stream) {
  if (B < 0 || F <= 0 || S < 1 || total_T < 0) return GPKL_ERR_DESC;
  if (!recon) return GPKL_ERR_NULL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B == 0 || total_T == 0) {
    cudaMemsetAsync(recon, 0, sizeof(double), st);
    return GPKL_OK;
  }
  if (!x || !x_decode || !lengths || !workspace) return GPKL_ERR_NULL;
  const ReconWs w = plan_recon(B, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, B, w.offsets);
  note_launch();
  const cudaError_t e = launch_recon_fwd(x, x_decode, reinterpret_cast<const long long*>(w.offsets), B, F, S,
                                         (long long)S * total_T, w.partials, recon, st);
  return e == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

extern "C" int gpkl_recon_backward(int32_t B, int32_t F, int32_t S, int64_t total_T, const float* x,
                                   const float* x_decode, const int32_t* lengths, const double* g_recon,
                                   float* g_x_decode, void* workspace, size_t ws_bytes, void* stream) {
  if (B < 0 || F <= 0 || S < 1 || total_T < 0) return GPKL_ERR_DESC;
  if (B == 0 || total_T == 0) return GPKL_OK;
  if (!x || !x_decode || !lengths || !g_x_decode || !workspace) return GPKL_ERR_NULL;
  const ReconWs w = plan_recon(B, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, B, w.offsets);
  note_launch();
  const cudaError_t e = launch_recon_bwd(x, x_decode, reinterpret_cast<const long long*>(w.offsets), B, F, S,
                                         (long long)S * total_T, g_recon, g_x_decode, st);
  return e == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// ---- GP-recognition sampler ---------------------------------------------------------------
// z = m + (chol(K(t, ell)) + diag(sqrt(exp(logvar)))) eps  and the per-row standard KL (GP_recog_VAE_prior.py:137-168,
// :65-70, :274-276).  The chol(K) eps part is the sample of the fused GP op (gpkl_forward with the prior set to the
// posterior's own lengthscales, whose KL output is discarded into the workspace); recog_fwd/bwd_kernel add the rest.
namespace {
struct RecogWs {
  void* inner;
  size_t inner_bytes;
  int64_t* offsets;
  float* kl_pairs;
  double* scalars;  // [0] discarded kl_sum, [1] constant 0 (upstream gradient of the discarded KL)
  size_t total;
};
GpklDesc recog_inner_desc(const GpklDesc& d) {
  GpklDesc in = d;
  in.posterior = GPKL_POST_GP;
  in.flags = d.flags & ~GPKL_FLAG_GRAD_ELL_P;
  return in;
}
RecogWs plan_recog(const GpklDesc& d, void* base) {
  RecogWs w;
  unsigned char* b = static_cast<unsigned char*>(base);
  size_t off = 0;
  const GpklDesc in = recog_inner_desc(d);
  w.inner = b;
  w.inner_bytes = plan(in, nullptr).total;
  off += align_up(w.inner_bytes);
  w.offsets = reinterpret_cast<int64_t*>(b + off);
  off += align_up(((size_t)d.B + 1) * sizeof(int64_t));
  w.kl_pairs = reinterpret_cast<float*>(b + off);
  off += align_up((size_t)d.B * d.D * sizeof(float));
  w.scalars = reinterpret_cast<double*>(b + off);
  off += align_up(2 * sizeof(double));
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t gpkl_recog_workspace_bytes(const GpklDesc* desc) {
  if (!desc) return 0;
  const GpklDesc in = recog_inner_desc(*desc);
  if (check_desc(&in) != GPKL_OK) return 0;
  return plan_recog(*desc, nullptr).total;
}

extern "C" int gpkl_recog_forward(const GpklDesc* desc, const float* mean, const float* logvar, const float* times,
                                  const int32_t* lengths, const float* ell, const float* eps, float* z, float* kl_rows,
                                  double* kl_sum, int32_t* status, void* workspace, size_t ws_bytes, void* stream) {
  if (!desc) return GPKL_ERR_NULL;
  const GpklDesc in = recog_inner_desc(*desc);
  int rc = check_desc(&in);
  if (rc != GPKL_OK) return rc;
  if (in.flags & GPKL_FLAG_PHILOX_EPS) return GPKL_ERR_UNSUPPORTED;  // the epilogue kernel reads eps
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!kl_sum) return GPKL_ERR_NULL;
  if (in.B == 0 || in.total_T == 0) {
    cudaMemsetAsync(kl_sum, 0, sizeof(double), st);
    if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), st);
    return GPKL_OK;
  }
  if (!mean || !logvar || !times || !lengths || !ell || !eps || !z || !kl_rows || !workspace) return GPKL_ERR_NULL;
  const RecogWs w = plan_recog(*desc, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  rc = gpkl_forward(&in, mean, times, lengths, ell, ell, nullptr, eps, z, w.kl_pairs, w.scalars, nullptr, status, w.inner,
                    w.inner_bytes, stream);
  if (rc != GPKL_OK) return rc;
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, in.B, w.offsets);
  note_launch();
  if (launch_recog_fwd(mean, logvar, eps, reinterpret_cast<const long long*>(w.offsets), in.B, in.D, in.S, in.T_max,
                       in.total_T, z, kl_rows, st) != cudaSuccess)
    return GPKL_ERR_CUDA;
  sum_pairs_kernel<<<1, 1024, 0, st>>>(kl_rows, (int)in.total_T, kl_sum);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

extern "C" int gpkl_recog_backward(const GpklDesc* desc, const float* mean, const float* logvar, const float* times,
                                   const int32_t* lengths, const float* ell, const float* eps, const float* g_z,
                                   const double* g_kl_sum, const float* g_kl_rows, float* g_mean, float* g_logvar,
                                   float* g_ell, int32_t* status, void* workspace, size_t ws_bytes, void* stream) {
  if (!desc) return GPKL_ERR_NULL;
  const GpklDesc in = recog_inner_desc(*desc);
  int rc = check_desc(&in);
  if (rc != GPKL_OK) return rc;
  if (in.flags & GPKL_FLAG_PHILOX_EPS) return GPKL_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (in.B == 0 || in.total_T == 0) {
    if (g_ell) cudaMemsetAsync(g_ell, 0, sizeof(float) * in.D, st);
    if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), st);
    return GPKL_OK;
  }
  if (!mean || !logvar || !times || !lengths || !ell || !eps || !g_mean || !g_logvar || !g_ell || !workspace)
    return GPKL_ERR_NULL;
  const RecogWs w = plan_recog(*desc, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  cudaMemsetAsync(w.scalars, 0, 2 * sizeof(double), st);
  // sample path only: the discarded KL gets upstream gradient 0, so g_mean = sum_s g_z and g_ell = <L-bar, dL/d ell>
  rc = gpkl_backward(&in, mean, times, lengths, ell, ell, nullptr, eps, g_z, w.scalars + 1, nullptr, g_mean, g_ell,
                     nullptr, nullptr, status, w.inner, w.inner_bytes, stream);
  if (rc != GPKL_OK) return rc;
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, in.B, w.offsets);
  note_launch();
  if (launch_recog_bwd(mean, logvar, eps, g_z, g_kl_sum, g_kl_rows, reinterpret_cast<const long long*>(w.offsets), in.B,
                       in.D, in.S, in.T_max, in.total_T, g_mean, g_logvar, st) != cudaSuccess)
    return GPKL_ERR_CUDA;
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// ---- ragged batch producer ---------------------------------------------------------------
namespace {
struct CollateWs {
  int64_t* offsets;
  size_t total;
};
CollateWs plan_collate(int B, int max_time, void* base) {
  CollateWs w;
  unsigned char* b = static_cast<unsigned char*>(base);
  size_t off = 0;
  w.offsets = reinterpret_cast<int64_t*>(b + off);
  off += align_up(((size_t)B + 1) * sizeof(int64_t));
  (void)max_time;
  w.total = off;
  return w;
}
}  // namespace

extern "C" size_t gpkl_collate_workspace_bytes(int32_t B, int32_t max_time) {
  return (B < 0 || max_time < 0) ? 0 : plan_collate(B, max_time, nullptr).total;
}

extern "C" int gpkl_collate(int32_t N, int32_t F, int32_t T_full, int32_t B, int32_t max_time, const float* data,
                            const float* time_grid, const int32_t* index, float* x, float* times, int32_t* lengths,
                            int64_t* total_T, void* workspace, size_t ws_bytes, void* stream) {
  if (N < 0 || F <= 0 || T_full < 0 || B < 0 || max_time < 0) return GPKL_ERR_DESC;
  if (!index && B > N) return GPKL_ERR_DESC;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B == 0) {
    if (total_T) cudaMemsetAsync(total_T, 0, sizeof(int64_t), st);
    return GPKL_OK;
  }
  if (!data || !time_grid || !x || !times || !lengths || !workspace) return GPKL_ERR_NULL;
  const CollateWs w = plan_collate(B, max_time, workspace);
  if (ws_bytes < w.total) return GPKL_ERR_WORKSPACE;
  if (launch_collate_scan(data, time_grid, index, B, F, T_full, max_time, lengths, times, st) != cudaSuccess)
    return GPKL_ERR_CUDA;
  offsets_kernel<<<1, 1024, 0, st>>>(lengths, B, w.offsets);
  note_launch();
  if (launch_collate_gather(data, index, reinterpret_cast<const long long*>(w.offsets), B, F, T_full, max_time, x, st) != cudaSuccess)
    return GPKL_ERR_CUDA;
  if (total_T) cudaMemcpyAsync(total_T, w.offsets + B, sizeof(int64_t), cudaMemcpyDeviceToDevice, st);
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// ---- the GPKL_FLAG_PHILOX_EPS noise stream, materialised ---------------------------------------------------------------
namespace gpkl {
static __global__ void philox_fill_kernel(const unsigned long long* __restrict__ seed, long long n, float* __restrict__ eps) {
  const unsigned long long sd = *seed;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    eps[e] = philox_normal(sd, (unsigned long long)e);
}
}  // namespace gpkl

extern "C" int gpkl_philox_normal(const uint64_t* seed_dev, int64_t n, float* eps, void* stream) {
  if (n < 0) return GPKL_ERR_DESC;
  if (n == 0) return GPKL_OK;
  if (!seed_dev || !eps) return GPKL_ERR_NULL;
  const int grid = (int)((n + 255) / 256 < (int64_t)kNumSMs * 8 ? (n + 255) / 256 : (int64_t)kNumSMs * 8);
  philox_fill_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const unsigned long long*>(seed_dev), (long long)n, eps);
  note_launch();
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// ---- GP posterior imputation -----------------------------------------------------------------
extern "C" size_t gpkl_impute_workspace_bytes(int32_t B) {
  return B < 0 ? 0 : align_up(((size_t)B + 1) * sizeof(int64_t));
}

extern "C" int gpkl_impute(int32_t B, int32_t D, int32_t n_obs_max, int32_t n_full, int32_t kernel, float ell, float noise,
                           const float* z_obs, const float* t_obs, const int32_t* n_obs, const float* t_full,
                           const float* eps, float* out, int32_t* status, void* workspace, size_t ws_bytes, void* stream) {
  if (B < 0 || D <= 0 || n_obs_max < 0 || n_full < 0) return GPKL_ERR_DESC;
  if (kernel != GPKL_KERNEL_RBF && kernel != GPKL_KERNEL_CAUCHY) return GPKL_ERR_DESC;
  if (!(ell > 0.0f) || !(noise >= 0.0f) || !(noise < 1.0f)) return GPKL_ERR_DESC;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (status) cudaMemsetAsync(status, 0, sizeof(int32_t), st);
  if (B == 0 || n_full == 0) return GPKL_OK;
  if (!z_obs || !t_obs || !n_obs || !t_full || !out || !workspace) return GPKL_ERR_NULL;
  if (ws_bytes < gpkl_impute_workspace_bytes(B)) return GPKL_ERR_WORKSPACE;
  if (impute_smem_bytes(n_obs_max, n_full) > kMaxDynSmem) return GPKL_ERR_UNSUPPORTED;
  int64_t* off = static_cast<int64_t*>(workspace);
  offsets_kernel<<<1, 1024, 0, st>>>(n_obs, B, off);
  note_launch();
  const cudaError_t e = launch_impute(kernel, B, D, n_obs_max, n_full, z_obs, t_obs, n_obs, t_full, eps,
                                      reinterpret_cast<const long long*>(off), ell, noise, out, status, st);
  return e == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}

// ---- host-buffer step ---------------------------------------------------------------------------
namespace {
struct Staging {
  float *mean, *times, *ell_q, *ell_p, *aux, *eps, *g_z, *z, *kl_pairs, *g_mean, *g_ell_q, *g_ell_p, *g_aux;
  int32_t* lengths;
  double* kl_sum;
  void* ws;
  size_t ws_bytes, total;
};
Staging plan_staging(const GpklDesc& d, void* base) {
  Staging s;
  unsigned char* b = static_cast<unsigned char*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { unsigned char* p = b + off; off += align_up(bytes); return p; };
  const size_t rows = (size_t)d.total_T, P = (size_t)d.B * d.D;
  const size_t auxw = d.posterior == GPKL_POST_DIAG ? 1 : (d.posterior == GPKL_POST_BIDIAG ? 2 : 0);
  s.mean = (float*)take(rows * d.D * 4);
  s.times = (float*)take((size_t)d.B * d.T_max * 4);
  s.lengths = (int32_t*)take((size_t)d.B * 4);
  s.ell_q = (float*)take((size_t)d.D * 4);
  s.ell_p = (float*)take((size_t)d.D * 4);
  s.aux = (float*)take(rows * d.D * auxw * 4);
  s.eps = (float*)take(P * d.S * d.T_max * 4);
  s.g_z = (float*)take(rows * d.S * d.D * 4);
  s.z = (float*)take(rows * d.S * d.D * 4);
  s.kl_pairs = (float*)take(P * 4);
  s.kl_sum = (double*)take(8);
  s.g_mean = (float*)take(rows * d.D * 4);
  s.g_ell_q = (float*)take((size_t)d.D * 4);
  s.g_ell_p = (float*)take((size_t)d.D * 4);
  s.g_aux = (float*)take(rows * d.D * auxw * 4);
  s.ws_bytes = plan(d, nullptr).total;
  s.ws = take(s.ws_bytes);
  s.total = off;
  return s;
}
}  // namespace

extern "C" size_t gpkl_step_host_bytes(const GpklDesc* desc) {
  if (check_desc(desc) != GPKL_OK) return 0;
  return plan_staging(*desc, nullptr).total;
}

extern "C" int gpkl_step_host(const GpklDesc* desc, const float* mean_host, const float* times_host,
                              const int32_t* lengths_host, const float* ell_q_host, const float* ell_p_host,
                              const float* aux_host, const float* eps_host, const float* g_z_host, float* z_host,
                              float* kl_pairs_host, double* kl_sum_host, float* g_mean_host, float* g_ell_q_host,
                              float* g_ell_p_host, float* g_aux_host, void* staging, size_t staging_bytes,
                              void* stream) {
  int rc = check_desc(desc);
  if (rc != GPKL_OK) return rc;
  const GpklDesc& d = *desc;
  if (!mean_host || !times_host || !lengths_host || !ell_p_host || !eps_host || !staging) return GPKL_ERR_NULL;
  if (d.posterior == GPKL_POST_GP && !ell_q_host) return GPKL_ERR_NULL;
  if (d.posterior != GPKL_POST_GP && !aux_host) return GPKL_ERR_NULL;
  const Staging s = plan_staging(d, staging);
  if (staging_bytes < s.total) return GPKL_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t rows = (size_t)d.total_T, P = (size_t)d.B * d.D;
  const size_t auxw = d.posterior == GPKL_POST_DIAG ? 1 : (d.posterior == GPKL_POST_BIDIAG ? 2 : 0);
  auto h2d = [&](void* dst, const void* src, size_t bytes) {
    if (src && bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st);
  };
  auto d2h = [&](void* dst, const void* src, size_t bytes) {
    if (dst && bytes) cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st);
  };
  h2d(s.mean, mean_host, rows * d.D * 4);
  h2d(s.times, times_host, (size_t)d.B * d.T_max * 4);
  h2d(s.lengths, lengths_host, (size_t)d.B * 4);
  h2d(s.ell_q, ell_q_host, (size_t)d.D * 4);
  h2d(s.ell_p, ell_p_host, (size_t)d.D * 4);
  h2d(s.aux, aux_host, rows * d.D * auxw * 4);
  h2d(s.eps, eps_host, (d.flags & GPKL_FLAG_PHILOX_EPS) ? sizeof(uint64_t) : P * d.S * d.T_max * 4);
  h2d(s.g_z, g_z_host, rows * d.S * d.D * 4);
  rc = gpkl_forward(desc, s.mean, s.times, s.lengths, ell_q_host ? s.ell_q : nullptr, s.ell_p,
                    aux_host ? s.aux : nullptr, s.eps, s.z, s.kl_pairs, s.kl_sum, nullptr, nullptr, s.ws, s.ws_bytes,
                    stream);
  if (rc != GPKL_OK) return rc;
  rc = gpkl_backward(desc, s.mean, s.times, s.lengths, ell_q_host ? s.ell_q : nullptr, s.ell_p,
                     aux_host ? s.aux : nullptr, s.eps, g_z_host ? s.g_z : nullptr, nullptr, nullptr, s.g_mean,
                     s.g_ell_q, s.g_ell_p, s.g_aux, nullptr, s.ws, s.ws_bytes, stream);
  if (rc != GPKL_OK) return rc;
  d2h(z_host, s.z, rows * d.S * d.D * 4);
  d2h(kl_pairs_host, s.kl_pairs, P * 4);
  d2h(kl_sum_host, s.kl_sum, 8);
  d2h(g_mean_host, s.g_mean, rows * d.D * 4);
  if (d.posterior == GPKL_POST_GP) d2h(g_ell_q_host, s.g_ell_q, (size_t)d.D * 4);
  if (d.flags & GPKL_FLAG_GRAD_ELL_P) d2h(g_ell_p_host, s.g_ell_p, (size_t)d.D * 4);
  d2h(g_aux_host, s.g_aux, rows * d.D * auxw * 4);
  return cudaGetLastError() == cudaSuccess ? GPKL_OK : GPKL_ERR_CUDA;
}
