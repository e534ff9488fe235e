/* gpkl.h -- C ABI of the B200-native GP-prior KL path for GP-VAE (libgpkl.so).
 *
 * Drop-in boundary for ONE path of ethanev/GP-VAE: the GP-prior term of the ELBO.  The reference has
 * no plugin/FFI layer; the boundary is the four Python call sites in its main()
 *   prior_kernels(...)    src/Models/Full_GP_VAE_dynamic_time.py:332  (def :100-130)
 *   approx_kernels(...)   src/Models/Full_GP_VAE_dynamic_time.py:335  (def :60-98)
 *   gp_vae_sample(...)    src/Models/Full_GP_VAE_dynamic_time.py:339  (def :174-195)
 *   calc_gp_kl(...)       src/Models/Full_GP_VAE_dynamic_time.py:340  (def :197-260)
 * (V2, diagonal posterior: vae_sample / calc_gp_kl, src/Models/VAE_GPprior_diag_cov.py:203-204,
 *  defs :64-71, :73-119) and TensorFlow's autodiff through them (:361).  Because all four share the
 * Cholesky factor of K_q they are ONE fused forward entry point here, plus one backward entry point.
 *
 * Conventions
 *   - plain C: POD descriptor, raw pointers, sizes; no C++/torch types, no exceptions.
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all buffers,
 *     including the workspace; the library allocates nothing on the data path.  The only process-global state is
 *     measurement-side and none of it influences results: an atomic launch counter (gpkl_launch_count), the CUDA
 *     events of the optional kernel profiler (gpkl_profile_enable / _read, off by default and NOT thread-safe: one
 *     measuring thread) and the internal phase-trace pointer of the developer tools.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no host synchronisation, no
 *     allocation => CUDA-graph capturable and re-entrant.
 *   - outputs are fully overwritten (no caller zero-fill needed); reductions are deterministic.
 *   - layouts are the reference's:
 *       mean   [total_T, D] f32   rows sequence-major then time (encoder output, :329)
 *       times  [B, T_max]   f32   zero padded on the right (DataHandler.py:151)
 *       lengths[B]          i32   T_b, 0 <= T_b <= T_max            (:322)
 *       ell_q, ell_p [D]    f32   posterior / prior lengthscales    (:72, :114)
 *       eps    [B, D, S, T_max] f32  N(0,1) draws (tf.random_normal at :166, made explicit)
 *       aux    posterior DIAG: logvar [total_T, D]; BIDIAG: [total_T, D, 2]; GP: unused (NULL)
 *       z      [S*total_T, D] f32  per sequence S blocks of [T_b, D]  (:186-194)
 *       kl_pairs [B*D] f32  pair index p = b*D + d (:216-217, :239);  kl_sum f64 scalar (:228)
 *   - return value: 0 on success, a GPKL_ERR_* code otherwise (see gpkl_strerror).  A non-positive-
 *     definite pivot yields NaN in that pair's outputs and increments *status (if non-NULL) on the
 *     device -- the analogue of TF's "Cholesky decomposition was not successful".
 */
#ifndef GPKL_H_
#define GPKL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPKL_VERSION 1

enum { GPKL_KERNEL_RBF = 0, GPKL_KERNEL_CAUCHY = 1 };
/* GPKL_POST_BIDIAG (V3, bidiagonal-precision posterior q = N(m, (B^T B)^-1); an extension named by north_star, not in the
 * reference): T_max <= 64 and S <= 8 under GPKL_TIER_AUTO / GPKL_TIER_WARP run the V3 hot tier (gpkl_bidiag.cu: one warp per
 * pair, bidiagonal solves + semiseparable recurrences against the per-sequence float64 K_p^-1, O(T^2) per pair); longer
 * sequences, a non-uniform ell_p (decided on the device) and GPKL_FLAG_PER_PAIR_PRIOR run the dense generic tier. */
enum { GPKL_POST_GP = 0, GPKL_POST_DIAG = 1, GPKL_POST_BIDIAG = 2 };
enum { GPKL_TIER_AUTO = 0, GPKL_TIER_GENERIC = 1, GPKL_TIER_WARP = 2, GPKL_TIER_BLOCK = 3 };
enum {
  GPKL_FLAG_GRAD_ELL_P = 1,     /* backward also produces d/d ell_p (Full_GP_VAE_fixed_for_MovMnist.py:96) */
  /* The reference's prior length scales are one constant for every latent dimension
   * (prior_time_chars = tf.constant(1.0, [latent_size, 1]), Full_GP_VAE_dynamic_time.py:114), so K_p of a
   * sequence is the same matrix for its D pairs ("each batch has the SAME matrix", :108-109).  When the device
   * finds ell_p[0] == ... == ell_p[D-1] the hot tiers factor K_p once per SEQUENCE (a small pre-pass into the
   * workspace) instead of once per pair.  This flag forces the per-pair factorisation (A/B tests). */
  GPKL_FLAG_PER_PAIR_PRIOR = 2,
  /* Production-mode noise: no eps tensor exists.  The `eps` argument of gpkl_forward / gpkl_backward is then a DEVICE
   * pointer to ONE uint64 seed (read by the kernels at run time, so a captured CUDA graph can be replayed with a new
   * seed), and element ((b*D + d)*S + s)*T_max + t of the virtual eps tensor is generated in the kernels with
   * Philox4x32-10 + cuRAND's Box-Muller (definition: gpkl_common.cuh, "counter-based N(0,1) noise"; host restatement
   * pinned to the Random123 known-answer vectors: tests/test_philox_cpu.py).  It replaces tf.random_normal inside
   * tf_kernel (Full_GP_VAE_dynamic_time.py:166) without the HBM round trip of an explicit eps.  (gpkl_step_host: eps_host
   * points to the seed on the host; gpkl_recog_*: not supported, their epilogue kernels read eps.)
   * gpkl_philox_normal writes the same stream into memory (tests, callers that want to see the draws). */
  GPKL_FLAG_PHILOX_EPS = 4
};

enum {
  GPKL_OK = 0,
  GPKL_ERR_NULL = 1,         /* a required pointer is NULL */
  GPKL_ERR_DESC = 2,         /* inconsistent descriptor (negative sizes, unknown enum) */
  GPKL_ERR_UNSUPPORTED = 3,  /* combination not implemented (e.g. tier cannot hold this T) */
  GPKL_ERR_WORKSPACE = 4,    /* ws_bytes < gpkl_workspace_bytes(desc) */
  GPKL_ERR_CUDA = 5          /* a CUDA runtime call failed (launch error) */
};

typedef struct GpklDesc {
  int32_t B;         /* sequences in the batch (batch_size, :312) */
  int32_t D;         /* latent dimensions (latent_size, :317) */
  int32_t T_max;     /* row stride of times and innermost stride of eps */
  int32_t S;         /* samples per pair (number_samples, :318) */
  int64_t total_T;   /* sum of lengths = rows of mean */
  int32_t kernel;    /* GPKL_KERNEL_* ; RBF == tf_kernel :162 */
  int32_t posterior; /* GPKL_POST_*   */
  float noise;       /* jitter, 1e-3 in tf_kernel :154 ; 0 reproduces the numpy kernel_matrix of V2 */
  int32_t flags;     /* GPKL_FLAG_* */
  int32_t tier;      /* GPKL_TIER_* ; AUTO picks by T_max */
  int32_t reserved;
} GpklDesc;

int gpkl_version(void);
const char* gpkl_strerror(int code);

/* Bytes of device workspace forward/backward need for this descriptor (max of both). */
size_t gpkl_workspace_bytes(const GpklDesc* desc);

/* Fused forward: kernel build + Cholesky + sample + KL.
 * Replaces prior_kernels + approx_kernels + gp_vae_sample + calc_gp_kl (see header comment).
 * logdets [B*D, 2] = (log|K_p|, log|K_q|) per pair, may be NULL.  status may be NULL. */
int gpkl_forward(const GpklDesc* desc, const float* mean, const float* times, const int32_t* lengths,
                 const float* ell_q, const float* ell_p, const float* aux, const float* eps,
                 float* z, float* kl_pairs, double* kl_sum, float* logdets, int32_t* status,
                 void* workspace, size_t ws_bytes, void* stream);

/* Backward of  Loss = g_kl_sum*kl_sum + <g_kl_pairs, kl_pairs> + <g_z, z>  (replaces TF autodiff, :361).
 * g_z [S*total_T, D] may be NULL (== 0); g_kl_sum is a DEVICE f64 scalar, NULL == 1.0;
 * g_kl_pairs [B*D] may be NULL (== 0).  Outputs: g_mean [total_T, D]; g_ell_q [D] (GP posterior only,
 * may be NULL otherwise); g_ell_p [D] only written when GPKL_FLAG_GRAD_ELL_P (else may be NULL);
 * g_aux like aux (DIAG/BIDIAG only).  The factors are recomputed on chip: nothing T x T is read
 * from or written to HBM between forward and backward. */
int gpkl_backward(const GpklDesc* desc, const float* mean, const float* times, const int32_t* lengths,
                  const float* ell_q, const float* ell_p, const float* aux, const float* eps,
                  const float* g_z, const double* g_kl_sum, const float* g_kl_pairs,
                  float* g_mean, float* g_ell_q, float* g_ell_p, float* g_aux, int32_t* status,
                  void* workspace, size_t ws_bytes, void* stream);

/* Device staging area a host-buffer step needs (inputs + outputs + workspace), in bytes. */
size_t gpkl_step_host_bytes(const GpklDesc* desc);

/* One forward+backward step with HOST buffers (pinned memory recommended): copies the inputs to the
 * device staging area, runs gpkl_forward + gpkl_backward with g_kl_sum = 1, copies z, kl_pairs, kl_sum,
 * g_mean, g_ell_q, g_ell_p back, all on `stream` (asynchronous; synchronise the stream before reading).
 * g_z_host / aux_host / g_aux_host / g_ell_p_host may be NULL as for the device entry points. */
int gpkl_step_host(const GpklDesc* desc, const float* mean_host, const float* times_host,
                   const int32_t* lengths_host, const float* ell_q_host, const float* ell_p_host,
                   const float* aux_host, const float* eps_host, const float* g_z_host,
                   float* z_host, float* kl_pairs_host, double* kl_sum_host, float* g_mean_host,
                   float* g_ell_q_host, float* g_ell_p_host, float* g_aux_host,
                   void* staging, size_t staging_bytes, void* stream);

/* ---- reconstruction term (first "next" row after the path) ------------------------------------
 * recon = sum over rows of mean_s  -sum_f [ x log(1e-10 + xd) + (1-x) log(1 - xd + 1e-10) ]      (float64 scalar)
 * replaces src/Models/Full_GP_VAE_dynamic_time.py:323-327 (x tiled over the S samples) and :349-356; the loss is
 * recon + beta * kl_sum (:360).  x [total_T, F] f32 targets, x_decode [S*total_T, F] f32 in the layout of z (per
 * sequence S blocks of [T_b, F]); lengths [B] i32.  Backward: g_x_decode = d(g_recon*recon)/d x_decode, g_recon a
 * DEVICE f64 scalar (NULL == 1).  Workspace: gpkl_recon_workspace_bytes(B). */
size_t gpkl_recon_workspace_bytes(int32_t B);
int gpkl_recon_forward(int32_t B, int32_t F, int32_t S, int64_t total_T, const float* x, const float* x_decode,
                       const int32_t* lengths, double* recon, void* workspace, size_t ws_bytes, void* stream);
int gpkl_recon_backward(int32_t B, int32_t F, int32_t S, int64_t total_T, const float* x, const float* x_decode,
                        const int32_t* lengths, const double* g_recon, float* g_x_decode, void* workspace,
                        size_t ws_bytes, void* stream);

/* ---- GP-recognition sampler ------------------------------------------------------------
 * Replaces approx_kernels / tf_kernel_approx / gp_vae_sample / standard_vae_kl of src/Models/GP_recog_VAE_prior.py
 * (:72-117, :137-168, :170-191, :65-70; call sites :274-284):
 *   z_s  = m + (chol(K(t, ell)) + diag(sqrt(exp(logvar)))) eps_s        per (sequence, latent-dim) pair
 *   kl_rows[r] = -1/2 sum_d (1 + log(1e-10 + exp(logvar)) - mean^2 - exp(logvar))   per row r of mean (:69, :276)
 *   kl_sum = sum_r kl_rows[r]  (:277)
 * desc: B, D, T_max, S, total_T, kernel, noise, tier as for gpkl_forward; `posterior` is ignored.  logvar [total_T, D],
 * ell [D] = approx_time_chars (:81), eps [B, D, S, T_max], z [S*total_T, D] in the layout of gpkl_forward.
 * Backward of  Loss = g_kl_sum*kl_sum + <g_kl_rows, kl_rows> + <g_z, z>  (g_kl_sum DEVICE f64, NULL == 1; g_kl_rows,
 * g_z may be NULL == 0): g_mean, g_logvar [total_T, D], g_ell [D]. */
size_t gpkl_recog_workspace_bytes(const GpklDesc* desc);
int gpkl_recog_forward(const GpklDesc* desc, const float* mean, const float* logvar, const float* times,
                       const int32_t* lengths, const float* ell, const float* eps, float* z, float* kl_rows,
                       double* kl_sum, int32_t* status, void* workspace, size_t ws_bytes, void* stream);
int gpkl_recog_backward(const GpklDesc* desc, const float* mean, const float* logvar, const float* times,
                        const int32_t* lengths, const float* ell, const float* eps, const float* g_z,
                        const double* g_kl_sum, const float* g_kl_rows, float* g_mean, float* g_logvar, float* g_ell,
                        int32_t* status, void* workspace, size_t ws_bytes, void* stream);

/* ---- ragged batch producer ----------------------------------------------------------------
 * Replaces SyntheticDataHandler._prep_dataset + data_batch (src/Models/DataHandler.py:129-156, :111-127) for one batch:
 * data [N, F, T_full] f32 with -1 marking missing time points (:143-145; validity is read from feature 0, the reference
 * requires the mask to be the same for every feature), time_grid [T_full] f32 (data['time']), index [B] i32 = the
 * sequences of this batch (NULL == 0..B-1).  Outputs: x [sum_T, F] rows sequence-major then time (capacity
 * B*min(T_full, max_time) rows), times [B, max_time] zero padded on the right (:149-151), lengths [B], *total_T (device
 * int64, may be NULL).  Sequences longer than max_time are truncated (the reference's np.pad raises there). */
size_t gpkl_collate_workspace_bytes(int32_t B, int32_t max_time);
int gpkl_collate(int32_t N, int32_t F, int32_t T_full, int32_t B, int32_t max_time, const float* data,
                 const float* time_grid, const int32_t* index, float* x, float* times, int32_t* lengths,
                 int64_t* total_T, void* workspace, size_t ws_bytes, void* stream);

/* ---- GP posterior conditioning / imputation -------------------------------------------------
 * Replaces sample_given_part_latent / post_gp_sample (src/Models/FullGP_and_GPdecoder_dynamic_time_analysis.py:40-56,
 * :96-111) for a whole batch: the latent values z_obs [sum n_obs, D] (rows sequence-major then time, the layout of `mean`)
 * at the observed time points t_obs [B, n_obs_max] (n_obs [B] valid per sequence) are conditioned on, and the predictive
 * mean (eps == NULL) or one sample (eps [B, D, n_full] N(0,1), the np.random.normal of :51 made explicit) on the full
 * grid t_full [B, n_full] is written to out [B*n_full, D] (per sequence n_full rows, post_gp_sample's concatenation
 * :108-110).  Kernel as kernel_function :8-14: (1-noise) k + noise wherever two time stamps coincide exactly, entries
 * rounded to float32; the reference fixes ell = 1, noise = 1e-3, RBF.  Matrices are factored once per SEQUENCE (they do
 * not depend on the latent row) in shared memory; n_obs_max and n_full are limited by it (about 110 each) ->
 * GPKL_ERR_UNSUPPORTED beyond.  The posterior covariance K_ss + 1e-15 I - Lk^T Lk is factored even for the mean, as the
 * reference does (:50 precedes :53): if it is not positive definite -- always the case when an observed time point
 * coincides with a full-grid point, where the reference raises LinAlgError -- *status is incremented; the mean is still
 * written, sample rows are NaN. */
/* eps[e] = the e-th N(0,1) draw of the GPKL_FLAG_PHILOX_EPS stream for *seed_dev (device uint64), e in [0, n). */
int gpkl_philox_normal(const uint64_t* seed_dev, int64_t n, float* eps, void* stream);

size_t gpkl_impute_workspace_bytes(int32_t B);
int gpkl_impute(int32_t B, int32_t D, int32_t n_obs_max, int32_t n_full, int32_t kernel, float ell, float noise,
                const float* z_obs, const float* t_obs, const int32_t* n_obs, const float* t_full, const float* eps,
                float* out, int32_t* status, void* workspace, size_t ws_bytes, void* stream);

/* ---- measurement hooks (bench.py; not part of the data path) ------------------------------------
 * These are the only process-global state in the library and are not thread safe.
 * gpkl_launch_count: kernels this library has launched since load (bench.py's gpu_launches).
 * gpkl_profile_enable(1): forward/backward additionally record CUDA events on `stream` immediately
 *   around their dominant kernel launch (ring of 1024 pairs each); gpkl_profile_read synchronises
 *   those events and returns summed milliseconds and launch counts since the last read. */
int64_t gpkl_launch_count(void);
int gpkl_profile_enable(int on);
int gpkl_profile_read(double* fwd_ms, int32_t* fwd_launches, double* bwd_ms, int32_t* bwd_launches);

/* FP32 CUDA-core peak microbenchmark: launches one kernel of independent FFMA chains on `stream`
 * (148*8 CTAs x 256 threads x iters x 16 FMAs) and stores the flop count of that launch in *flops.
 * The caller times it with CUDA events; sink (device, >= 148*8*256 floats) keeps the work alive. */
int gpkl_fp32_peak_launch(float* sink, int32_t iters, double* flops, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GPKL_H_ */
