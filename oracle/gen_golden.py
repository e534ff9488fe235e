"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN CODE.

Run in the dev container (needs /root/reference, which does not exist on the GPU box):

    python oracle/gen_golden.py

The reference (TensorFlow-1.x scripts) has no tests or golden vectors, so the pins
are made here: the unmodified reference modules are imported under oracle/tf1_stub.py (a torch-backed
TF1 look-alike) and their hot-path functions are called exactly as main() calls them
(Full_GP_VAE_dynamic_time.py:332-340, Full_GP_VAE_fixed_for_MovMnist.py:291-299,
VAE_GPprior_diag_cov.py:195-204).  Inputs, the recorded tf.random_normal draws, outputs and
autograd gradients are stored; tests/ then compare oracle/gp_kl_oracle.py and the CUDA op with them.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GPKL_REFERENCE", "/root/reference/src/Models")
OUT = os.path.join(ROOT, "tests", "golden")

sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import tf1_stub  # noqa: E402

tf = tf1_stub.install()
import Full_GP_VAE_dynamic_time as dyn  # noqa: E402  (reference, unmodified)
import Full_GP_VAE_fixed_for_MovMnist as fixed  # noqa: E402
import VAE_GPprior_diag_cov as diagcov  # noqa: E402
import GP_recog_VAE_prior as recog  # noqa: E402
import DataHandler as datahandler  # noqa: E402


class _Vars:
    """Make the next tf.Variable(...) calls return prepared leaf tensors (so lengthscales can be set)."""

    def __init__(self, *tensors):
        self.queue = list(tensors)

    def __enter__(self):
        self.old = tf.Variable
        tf.Variable = lambda *a, **k: self.queue.pop(0)
        return self

    def __exit__(self, *exc):
        tf.Variable = self.old


def _seeded_noise(seed):
    g = torch.Generator().manual_seed(seed)
    tf1_stub.RANDOM_LOG.clear()
    tf1_stub.RANDOM_SOURCE = lambda shape: torch.randn(shape, generator=g, dtype=torch.float32)
    tf.RANDOM_SOURCE = tf1_stub.RANDOM_SOURCE


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    conv = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print("wrote", name, {k: (v.shape, str(v.dtype)) for k, v in conv.items()})


def run_dynamic(name, times, lengths, mean, ell_q, S, seed, with_gz=True):
    """V1, ragged/irregular: the four calls of Full_GP_VAE_dynamic_time.main() (:332-340)."""
    B, T_max = times.shape
    D = mean.shape[1]
    lengths_t = torch.tensor(lengths, dtype=torch.int32)
    mean = mean.clone().requires_grad_(True)
    lq = ell_q.clone().reshape(D, 1).requires_grad_(True)
    _seeded_noise(seed)
    prior_kernel, prior_chars = dyn.prior_kernels(times, lengths_t, D, B)
    n_prior_draws = len(tf1_stub.RANDOM_LOG)
    with _Vars(lq):
        approx_kernel, chol_noise, approx_chars = dyn.approx_kernels(times, lengths_t, D, B, S)
    draws = tf1_stub.RANDOM_LOG[n_prior_draws:]
    assert len(draws) == B * D
    eps = torch.zeros(B, D, S, T_max)
    for b in range(B):
        for d in range(D):
            r = draws[b * D + d]            # [T_b, S]   (tf_kernel, :166)
            eps[b, d, :, : r.shape[0]] = r.t()
    z = dyn.gp_vae_sample(mean, chol_noise, lengths_t, B, S, D)
    kl_sum, kl = dyn.calc_gp_kl(mean, lengths_t, approx_kernel, prior_kernel, B, D)
    g = torch.Generator().manual_seed(seed + 1)
    g_z = torch.randn(z.shape, generator=g, dtype=torch.float32) if with_gz else torch.zeros_like(z)
    loss = kl_sum + (g_z.to(torch.float64) * z.to(torch.float64)).sum()
    loss.backward()
    _save(name, variant="v1", kernel="rbf", posterior="gp", noise=1e-3, S=S,
          times=times, lengths=np.asarray(lengths, np.int32), mean=mean, ell_q=lq.reshape(-1),
          ell_p=prior_chars.reshape(-1), eps=eps, g_z=g_z,
          z=z, kl_sum=kl_sum, kl_pairs=kl.reshape(-1), g_mean=mean.grad, g_ell_q=lq.grad.reshape(-1))


def run_fixed(name, B, D, mean, ell_q, ell_p, seed):
    """V1, fixed T=20 with TRAINABLE prior lengthscales (Full_GP_VAE_fixed_for_MovMnist.py:96, :291-299)."""
    T = 20
    prior_sequence = np.array([float(i) + 1 for i in range(T)], dtype=np.float32)
    test_sequences = np.array([[float(i) + 1 for i in range(T)] for _ in range(B)], dtype=np.float32)
    mean = mean.clone().requires_grad_(True)
    lq = ell_q.clone().reshape(D, 1).requires_grad_(True)
    lp = ell_p.clone().reshape(D, 1).requires_grad_(True)
    _seeded_noise(seed)
    with _Vars(lp):
        prior_kernel = fixed.prior_kernels(prior_sequence, D, B)
    n_prior_draws = len(tf1_stub.RANDOM_LOG)
    with _Vars(lq):
        approx_kernel, chol_noise = fixed.approx_kernels(test_sequences, D, B)
    draws = tf1_stub.RANDOM_LOG[n_prior_draws:]
    eps = torch.stack([d.reshape(T) for d in draws]).reshape(B, D, 1, T)
    z = fixed.gp_vae_sample(mean, chol_noise, B)
    kl_sum, kl = fixed.calc_gp_kl(mean, approx_kernel, prior_kernel, B)
    g = torch.Generator().manual_seed(seed + 1)
    g_z = torch.randn(z.shape, generator=g, dtype=torch.float32)
    loss = kl_sum + (g_z.to(torch.float64) * z.to(torch.float64)).sum()
    loss.backward()
    _save(name, variant="v1_fixed", kernel="rbf", posterior="gp", noise=1e-3, S=1,
          times=torch.from_numpy(test_sequences), lengths=np.full(B, T, np.int32), mean=mean,
          ell_q=lq.reshape(-1), ell_p=lp.reshape(-1), eps=eps, g_z=g_z, z=z, kl_sum=kl_sum,
          kl_pairs=kl.reshape(-1), g_mean=mean.grad, g_ell_q=lq.grad.reshape(-1), g_ell_p=lp.grad.reshape(-1))


def run_diag(name):
    """V2 golden G2 of SURVEY.md Appendix B: VAE_GPprior_diag_cov.calc_gp_kl with its own numpy
    kernel_matrix(20, 1.0) (:153-165, :195-204); B=5, T=20, D=100 are hard-coded there (:93-98)."""
    B, T, D = 5, 20, 100
    K = torch.from_numpy(diagcov.kernel_matrix(T, 1.0))
    mean = torch.linspace(-1, 1, B * T * D, dtype=torch.float32).reshape(B * T, D).clone().requires_grad_(True)
    logvar = torch.linspace(-2, 0.5, B * T * D, dtype=torch.float32).reshape(B * T, D).flip(0).clone().requires_grad_(True)
    kl_sum, kl = diagcov.calc_gp_kl(mean, logvar, K)
    kl_sum.backward()
    _seeded_noise(77)
    z = diagcov.vae_sample(mean.detach(), logvar.detach(), K, [T] * B)
    eps_rows = tf1_stub.RANDOM_LOG[-1]                    # [B*T, D]  (vae_sample :68)
    eps = eps_rows.reshape(B, T, D).permute(0, 2, 1).reshape(B, D, 1, T).contiguous()
    times = torch.arange(T, dtype=torch.float32).repeat(B, 1)
    _save(name, variant="v2", kernel="rbf", posterior="diag", noise=0.0, S=1,
          times=times, lengths=np.full(B, T, np.int32), mean=mean, logvar=logvar,
          ell_p=np.ones(D, np.float32), ell_q=np.ones(D, np.float32), eps=eps, K=K, z=z,
          kl_sum=kl_sum, kl_pairs=kl.reshape(-1), g_mean=mean.grad, g_logvar=logvar.grad)


def run_recon(name):
    """Reconstruction term + loss: the reference has them inline in main() (Full_GP_VAE_dynamic_time.py:323-327,
    :349-356, :360), so those SOURCE LINES are read from the reference at generation time and executed under the stub
    (nothing of the reference is copied into this repository)."""
    import textwrap
    src = open(os.path.join(REF, "Full_GP_VAE_dynamic_time.py")).read().split("\n")
    code = textwrap.dedent("\n".join(src[322:327] + src[348:356] + [src[359]]))
    g = torch.Generator().manual_seed(21)
    lengths = [4, 2, 5]
    F, S = 7, 2
    total = sum(lengths)
    x = (torch.rand(total, F, generator=g) < 0.3).float()
    x_decode = (torch.rand(S * total, F, generator=g) * 0.96 + 0.02).clone().requires_grad_(True)
    ns = dict(tf=tf, x=x, sequence_sizes_placeholder=torch.tensor(lengths, dtype=torch.int32), number_samples=S,
              x_decode=x_decode, beta=0.7, sum_gp_kl=torch.tensor(12.5, dtype=torch.float64))
    exec(code, ns)
    ns["sum_recon_loss"].backward()
    _save(name, variant="recon", S=S, lengths=np.asarray(lengths, np.int32), x=x, x_decode=x_decode,
          recon=ns["sum_recon_loss"], loss=ns["loss"], beta=0.7, kl=12.5, g_x_decode=x_decode.grad)


def run_recog(name, times, lengths, mean, logvar, ell, S, seed):
    """GP-recognition sampler: the calls of GP_recog_VAE_prior.main() (:274-284) -- standard_vae_kl, approx_kernels
    (tf_kernel_approx: float64 Cholesky + diag sqrt(var)), gp_vae_sample -- and autograd through them (:305)."""
    B, T_max = times.shape
    D = mean.shape[1]
    lengths_t = torch.tensor(lengths, dtype=torch.int32)
    mean = mean.clone().requires_grad_(True)
    logvar = logvar.clone().requires_grad_(True)
    lq = ell.clone().reshape(D, 1).requires_grad_(True)
    _seeded_noise(seed)
    kl = recog.standard_vae_kl(mean, logvar, D)
    kl = tf.scalar_mul(-1.0, tf.cast(kl, tf.float64))                      # :275-276
    sum_kl = tf.reduce_sum(kl)                                              # :277
    n0 = len(tf1_stub.RANDOM_LOG)
    with _Vars(lq):
        approx_kernel, chol_noise, chars = recog.approx_kernels(times, lengths_t, D, B, S, logvar)
    draws = tf1_stub.RANDOM_LOG[n0:]
    assert len(draws) == B * D
    eps = torch.zeros(B, D, S, T_max)
    for b in range(B):
        for d in range(D):
            r = draws[b * D + d]            # [T_b, S]   (tf_kernel_approx :159)
            eps[b, d, :, : r.shape[0]] = r.t()
    z = recog.gp_vae_sample(mean, chol_noise, lengths_t, B, S, D)
    g = torch.Generator().manual_seed(seed + 1)
    g_z = torch.randn(z.shape, generator=g, dtype=torch.float32)
    g_rows = torch.randn(kl.shape, generator=g, dtype=torch.float32)
    loss = 0.7 * sum_kl + (g_rows.to(torch.float64) * kl).sum() + (g_z.to(torch.float64) * z.to(torch.float64)).sum()
    loss.backward()
    _save(name, variant="recog", kernel="rbf", noise=1e-3, S=S, times=times, lengths=np.asarray(lengths, np.int32),
          mean=mean, logvar=logvar, ell=lq.reshape(-1), eps=eps, g_z=g_z, g_kl_rows=g_rows, g_kl_sum=0.7, z=z,
          kl_rows=kl.reshape(-1), kl_sum=sum_kl, g_mean=mean.grad, g_logvar=logvar.grad, g_ell=lq.grad.reshape(-1))


def run_collate(name):
    """Ragged batch producer: SyntheticDataHandler (DataHandler.py:96-162) on a synthetic -1-masked array with the
    reference's hard-coded 15 features (:145); two consecutive data_batch('train') calls."""
    rng = np.random.RandomState(31)
    N, F, T_full, max_time, bs = 8, 15, 45, 45, 3
    grid = np.linspace(0.0, 60.0, T_full).astype(np.float32)
    x = rng.rand(N, F, T_full).astype(np.float32)
    for i in range(N):
        drop = rng.rand(T_full) < 0.35
        x[i][:, drop] = -1.0
    h = datahandler.SyntheticDataHandler({"x": x.copy(), "time": grid}, max_time, batch_size=bs, train_fraction=0.75)
    out = {}
    for k in range(2):
        bx, bt, bl = h.data_batch("train")
        out["x%d" % k], out["times%d" % k], out["lengths%d" % k] = bx, bt, np.asarray(bl, np.int32)
        out["index%d" % k] = np.arange(k * bs, (k + 1) * bs, dtype=np.int32)
    _save(name, variant="collate", data=x, time_grid=grid, max_time=max_time, **out)


def run_impute(name, seed):
    """GP posterior imputation: sample_given_part_latent (FullGP_and_GPdecoder_dynamic_time_analysis.py:40-56) called for
    every latent row of every sequence exactly as post_gp_sample (:96-111) does, with np.random.normal replaced by recorded
    draws.  Observed time stamps sit BETWEEN the points of the full grid (the reference's own subset-of-grid input makes
    its posterior covariance singular and it raises LinAlgError -- recorded in the fixture as `coincident_raises`)."""
    import FullGP_and_GPdecoder_dynamic_time_analysis as ana  # noqa: E402  (reference, unmodified)
    rng = np.random.RandomState(seed)
    full = [float(i) + 1 for i in range(20)]                                   # x_space of the reference (:25)
    obs_times = [sorted(rng.choice(np.arange(1, 20), size=n, replace=False) + 0.5) for n in (7, 10, 4)]
    B, D = len(obs_times), 6
    t_s_matrix = [rng.randn(D, len(t)).astype(np.float32) for t in obs_times]
    eps = rng.randn(B, D, len(full)).astype(np.float32)
    draws = []
    old = np.random.normal
    np.random.normal = lambda *a, **k: draws.pop(0)
    try:
        draws[:] = [np.zeros((len(full), 1)) for _ in range(B * D)]
        mean = ana.post_gp_sample(t_s_matrix, obs_times, full, mean=True)
        draws[:] = [eps[b, d].reshape(-1, 1).astype(np.float64) for b in range(B) for d in range(D)]
        sample = ana.post_gp_sample(t_s_matrix, obs_times, full, mean=False)
    finally:
        np.random.normal = old
    raised = False
    try:
        ana.sample_given_part_latent(t_s_matrix[0][0][:5], [1.0, 3.0, 4.0, 8.0, 10.0], full, mean=True)
    except np.linalg.LinAlgError:
        raised = True
    nd_max = max(len(t) for t in obs_times)
    t_obs = np.zeros((B, nd_max), np.float32)
    for b, t in enumerate(obs_times):
        t_obs[b, : len(t)] = t
    z_obs = np.concatenate([m.T for m in t_s_matrix], 0)
    _save(name, variant="impute", t_obs=t_obs, n_obs=np.asarray([len(t) for t in obs_times], np.int32),
          t_full=np.tile(np.asarray(full, np.float32), (B, 1)), z_obs=z_obs, eps=eps, mean_out=mean, sample_out=sample,
          coincident_raises=np.asarray(raised))


def main():
    torch.manual_seed(0)
    # G1 -- SURVEY.md Appendix B golden case (regular grid, B=3 D=4 T=6)
    B, D, T = 3, 4, 6
    run_dynamic("g1_v1_regular", torch.arange(T, dtype=torch.float32).repeat(B, 1), [T] * B,
                torch.linspace(-1, 1, B * T * D, dtype=torch.float32).reshape(B * T, D),
                torch.tensor([0.5, 1.0, 2.0, 3.0]), S=1, seed=11)
    # G3 -- ragged lengths, irregular times, S=2 samples (DataHandler.py:143-151 layout)
    g = torch.Generator().manual_seed(5)
    lengths = [9, 5, 7, 2]
    B, D, T_max = 4, 3, 9
    times = torch.cumsum(torch.rand(B, T_max, generator=g) + 0.5, 1).float()
    for b, L in enumerate(lengths):
        times[b, L:] = 0
    run_dynamic("g3_v1_ragged_s2", times, lengths, torch.randn(sum(lengths), D, generator=g),
                torch.tensor([0.7, 1.3, 2.2]), S=2, seed=12)
    # G5 -- the reference's toy configuration shape: T<=45 ragged, D=2, l=[9,3]-like, spacing 1.36
    g = torch.Generator().manual_seed(6)
    lengths = [45, 31, 38, 27, 33]
    B, D, T_max = 5, 2, 45
    grid = torch.linspace(0, 60, 45)
    times = torch.zeros(B, T_max)
    for b, L in enumerate(lengths):
        keep = torch.sort(torch.randperm(45, generator=g)[:L]).values
        times[b, :L] = grid[keep]
    run_dynamic("g5_v1_toy_shape", times, lengths, 0.5 * torch.randn(sum(lengths), D, generator=g),
                torch.tensor([2.0, 1.5]), S=1, seed=13)
    # G4 -- fixed T=20 model, trainable prior lengthscales -> d/d l_p
    g = torch.Generator().manual_seed(7)
    B, D = 2, 3
    run_fixed("g4_v1_fixed_prior_grad", B, D, torch.randn(B * 20, D, generator=g),
              torch.tensor([0.8, 1.0, 1.7]), torch.tensor([1.0, 1.4, 0.9]), seed=14)
    # G2 -- V2 diagonal posterior
    run_diag("g2_v2_diag")
    # G6 -- reconstruction term and beta-weighted loss (next row after the path)
    run_recon("g6_recon_loss")
    # G7 -- GP-recognition sampler (S8(f) row 3): ragged, irregular, S=2
    g = torch.Generator().manual_seed(8)
    lengths = [8, 5, 11]
    B, D, T_max = 3, 4, 11
    times = torch.cumsum(torch.rand(B, T_max, generator=g) + 0.5, 1).float()
    for b, L in enumerate(lengths):
        times[b, L:] = 0
    run_recog("g7_recog_ragged_s2", times, lengths, torch.randn(sum(lengths), D, generator=g),
              0.5 * torch.randn(sum(lengths), D, generator=g) - 0.5, torch.tensor([1.0, 0.6, 1.8, 2.5]), S=2, seed=15)
    # G8 -- the reference's own configuration of that model: T=20 grid 0..19, ell = 1 (:81), S=1
    g = torch.Generator().manual_seed(9)
    B, D, T = 3, 6, 20
    run_recog("g8_recog_grid", torch.arange(T, dtype=torch.float32).repeat(B, 1), [T] * B,
              torch.randn(B * T, D, generator=g), 0.3 * torch.randn(B * T, D, generator=g) - 1.0, torch.ones(D), S=1, seed=16)
    # G9 -- ragged batch producer (S8(f) row 4)
    run_collate("g9_collate")
    # G10 -- GP posterior imputation (S8(f) row 2): the reference's sample_given_part_latent / post_gp_sample
    run_impute("g10_impute", 23)


if __name__ == "__main__":
    main()
