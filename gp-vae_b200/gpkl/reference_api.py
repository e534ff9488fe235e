"""The reference's four call sites, by name, over the fused op.

In the reference main() (src/Models/Full_GP_VAE_dynamic_time.py:332-340)

    prior_kernel, prior_time_chars         = prior_kernels(sequences, sizes, latent_size, batch_size)
    approx_kernel, chol_noise, approx_chars = approx_kernels(sequences, sizes, latent_size, batch_size, S)
    latent_sample                           = gp_vae_sample(latent_mean, chol_noise, sizes, batch_size, S, latent_size)
    sum_gp_kl, gp_kl                        = calc_gp_kl(latent_mean, sizes, approx_kernel, prior_kernel, batch_size, latent_size)

materialise [B*D, T_max^2] kernel tensors in memory.  Here the two *_kernels calls only record WHAT
to build (times, lengths, lengthscales) in light handles -- no T x T matrix is ever formed in HBM --
and the first of gp_vae_sample / calc_gp_kl that runs launches the fused CUDA op once; the other
returns the cached half.  Same argument order and meaning as the reference functions.
"""
import torch

from .ops import gp_prior_kl


class KernelHandle:
    """What prior_kernels / approx_kernels return in place of the [B*D, T_max^2] tensor."""

    def __init__(self, path, role, time_chars):
        self.path, self.role, self.time_chars = path, role, time_chars


class GPPriorPath:
    def __init__(self, latent_size, *, kernel="rbf", noise=1e-3, prior_lengthscale=1.0, approx_lengthscale=1.0,
                 train_prior=False, device="cuda:0", tier="auto"):
        dev = torch.device(device)
        # tf.Variable(tf.constant(1.0, shape=[latent_size,1]), name='approx_time_chars')   (:72)
        self.approx_time_chars = torch.nn.Parameter(torch.full((latent_size,), float(approx_lengthscale), device=dev))
        # tf.constant(1.0, ...) (:114) -- a tf.Variable in Full_GP_VAE_fixed_for_MovMnist.py:96
        p = torch.full((latent_size,), float(prior_lengthscale), device=dev)
        self.prior_time_chars = torch.nn.Parameter(p) if train_prior else p
        self.latent_size, self.kernel, self.noise, self.tier = latent_size, kernel, noise, tier
        self._seq = None
        self._cache = None

    def parameters(self):
        ps = [self.approx_time_chars]
        if isinstance(self.prior_time_chars, torch.nn.Parameter):
            ps.append(self.prior_time_chars)
        return ps

    def _note(self, sequences, sequence_sizes, latent_size, batch_size):
        assert latent_size == self.latent_size and sequences.shape[0] == batch_size
        self._seq = (sequences.contiguous().float(), sequence_sizes.to(torch.int32).contiguous())
        self._cache = None

    def prior_kernels(self, sequences, sequence_sizes, latent_size, batch_size):
        self._note(sequences, sequence_sizes, latent_size, batch_size)
        return KernelHandle(self, "prior", self.prior_time_chars), self.prior_time_chars

    def approx_kernels(self, sequences, sequence_sizes, latent_size, batch_size, number_samples, eps=None):
        self._note(sequences, sequence_sizes, latent_size, batch_size)
        self.number_samples, self._eps = number_samples, eps
        h = KernelHandle(self, "approx", self.approx_time_chars)
        return h, h, self.approx_time_chars  # (approx_kernel, chol_noise, time_chars)

    def _run(self, mean):
        if self._cache is None or self._cache[0] is not mean:
            times, lengths = self._seq
            out = gp_prior_kl(mean.contiguous(), times, lengths, self.approx_time_chars, self.prior_time_chars,
                              self._eps, kernel=self.kernel, noise=self.noise, S=self.number_samples, tier=self.tier)
            self._cache = (mean, out)
        return self._cache[1]

    def gp_vae_sample(self, mean, noise_chol_full_time, sequence_sizes, batch_size, number_samples, latent_size):
        assert isinstance(noise_chol_full_time, KernelHandle) and number_samples == self.number_samples
        return self._run(mean)[0]

    def calc_gp_kl(self, mean, sequence_sizes, approx_linear_kernel, prior_kernel, batch_size, latent_size):
        assert isinstance(approx_linear_kernel, KernelHandle) and isinstance(prior_kernel, KernelHandle)
        _, kl_sum, kl_pairs = self._run(mean)
        return kl_sum, kl_pairs


class GPRecogPath:
    """The GP-recognition model's call sites (src/Models/GP_recog_VAE_prior.py:274-284) over gp_recog_sample:

        kl                                   = standard_vae_kl(latent_mean, latent_log_var, latent_size)
        approx_kernel, chol_noise, time_chars = approx_kernels(sequences, sizes, latent_size, batch_size, S, latent_log_var)
        latent_sample                        = gp_vae_sample(latent_mean, chol_noise, sizes, batch_size, S, latent_size)

    approx_kernels records what to build; gp_vae_sample launches the fused op; standard_vae_kl returns the reference's
    value (+1/2 sum(1 + log(1e-10+var) - mean^2 - var) per row, i.e. MINUS the KL, :69) from the same launch when the
    sample has already been drawn for this (mean, log_var), else from a launch of its own."""

    def __init__(self, latent_size, *, kernel="rbf", noise=1e-3, approx_lengthscale=1.0, device="cuda:0", tier="auto"):
        dev = torch.device(device)
        # tf.Variable(tf.constant(1.0, shape=[latent_size,1]), name='approx_time_chars')   (:81)
        self.approx_time_chars = torch.nn.Parameter(torch.full((latent_size,), float(approx_lengthscale), device=dev))
        self.latent_size, self.kernel, self.noise, self.tier = latent_size, kernel, noise, tier
        self._seq = None
        self._cache = None

    def parameters(self):
        return [self.approx_time_chars]

    def approx_kernels(self, sequences, sequence_sizes, latent_size, batch_size, number_samples, encode_log_var, eps=None):
        assert latent_size == self.latent_size and sequences.shape[0] == batch_size
        self._seq = (sequences.contiguous().float(), sequence_sizes.to(torch.int32).contiguous())
        self._logvar, self.number_samples, self._eps = encode_log_var, number_samples, eps
        self._cache = None
        h = KernelHandle(self, "approx", self.approx_time_chars)
        return h, h, self.approx_time_chars

    def _run(self, mean, logvar):
        from .ops import gp_recog_sample
        if self._cache is None or self._cache[0] is not mean or self._cache[1] is not logvar:
            times, lengths = self._seq
            out = gp_recog_sample(mean.contiguous(), logvar.contiguous(), times, lengths, self.approx_time_chars, self._eps,
                                  kernel=self.kernel, noise=self.noise, S=self.number_samples, tier=self.tier)
            self._cache = (mean, logvar, out)
        return self._cache[2]

    def gp_vae_sample(self, mean, noise_chol_full_time, sequence_sizes, batch_size, number_samples, latent_size):
        assert isinstance(noise_chol_full_time, KernelHandle) and number_samples == self.number_samples
        return self._run(mean, self._logvar)[0]

    def standard_vae_kl(self, mean, log_var, latent_size):
        assert self._seq is not None, "call approx_kernels first (it records the time stamps the fused op needs)"
        return -self._run(mean, log_var)[2]


class SyntheticDataHandlerGPU:
    """SyntheticDataHandler.data_batch (src/Models/DataHandler.py:111-127) with the array resident on the GPU: the
    -1-masked data [N, F, T_full] and data['time'] are uploaded once; every data_batch() collates the next batch_size
    sequences on the device (gpkl.collate_batch) and returns (batch_xs, batch_time_steps, batch_lengths) as CUDA tensors.
    Reshuffling at the end of an epoch permutes an index vector instead of the data (:116-117, :137)."""

    def __init__(self, data, max_time, batch_size=5, device="cuda:0", generator=None):
        dev = torch.device(device)
        self.x = torch.as_tensor(data["x"], dtype=torch.float32).to(dev).contiguous()
        self.time = torch.as_tensor(data["time"], dtype=torch.float32).to(dev).contiguous()
        self.max_time, self.batch_size, self.generator = max_time, batch_size, generator
        self.order = torch.arange(self.x.shape[0], dtype=torch.int32, device=dev)
        self.counter = 0

    def data_batch(self, data_name="train"):
        from .ops import collate_batch
        n = self.x.shape[0]
        if self.counter + self.batch_size > n:
            self.order = torch.randperm(n, generator=self.generator).to(torch.int32).to(self.x.device)
            self.counter = 0
        index = self.order[self.counter:self.counter + self.batch_size].contiguous()
        self.counter += self.batch_size
        return collate_batch(self.x, self.time, index, self.max_time)


# ---- GP posterior imputation by the reference's names (FullGP_and_GPdecoder_dynamic_time_analysis.py) -------------------
def post_gp_sample(t_s_matrix, sequence_times, full_sequence_times, mean=False, *, eps=None, device="cuda:0"):
    """post_gp_sample (:96-111): t_s_matrix is a list of [D, n_obs_b] arrays (one per sequence, latent rows x kept time
    points), sequence_times the kept time stamps per sequence, full_sequence_times the grid to fill in.  Returns the
    reference's [B*n_full, D] array (numpy, float32).  eps [B, D, n_full] replaces np.random.normal (:51) for reproducible
    runs; mean=True returns the predictive mean."""
    import numpy as np
    from .ops import gp_posterior_impute
    dev = torch.device(device)
    B = len(t_s_matrix)
    D = int(np.asarray(t_s_matrix[0]).shape[0])
    n_obs = [len(t) for t in sequence_times]
    nd_max, ns = max(n_obs), len(full_sequence_times)
    t_obs = torch.zeros(B, nd_max, dtype=torch.float32)
    for b, t in enumerate(sequence_times):
        t_obs[b, : n_obs[b]] = torch.as_tensor(np.asarray(t, dtype=np.float32))
    z = torch.cat([torch.as_tensor(np.asarray(m, dtype=np.float32)).t() for m in t_s_matrix], 0).contiguous()  # [sum n_obs, D]
    t_full = torch.as_tensor(np.asarray(full_sequence_times, dtype=np.float32)).repeat(B, 1).contiguous()
    if not mean and eps is None:
        eps = torch.randn(B, D, ns)
    e = None if mean else torch.as_tensor(np.asarray(eps, dtype=np.float32)).contiguous().to(dev)
    out = gp_posterior_impute(z.to(dev), t_obs.to(dev), torch.tensor(n_obs, dtype=torch.int32, device=dev), t_full.to(dev), e)
    return out.cpu().numpy()


def sample_given_part_latent(z_d, z_d_times, full_seq_times, mean=False, *, eps=None, device="cuda:0"):
    """sample_given_part_latent (:40-56) for ONE latent row: returns [1, n_full]."""
    import numpy as np
    e = None if eps is None else np.asarray(eps, dtype=np.float32).reshape(1, 1, -1)
    out = post_gp_sample([np.asarray(z_d, dtype=np.float32).reshape(1, -1)], [z_d_times], full_seq_times, mean, eps=e, device=device)
    return out.reshape(1, -1)
