"""PyTorch-facing operator for the GP-prior KL path.

One fused op replaces the reference's four calls -- prior_kernels, approx_kernels, gp_vae_sample,
calc_gp_kl (src/Models/Full_GP_VAE_dynamic_time.py:332-340; V2: vae_sample/calc_gp_kl,
src/Models/VAE_GPprior_diag_cov.py:203-204) -- because they share the Cholesky factor of K_q, and a
hand-written backward replaces TF autodiff through them (:361).  PyTorch is plumbing here (device
memory, streams, autograd bookkeeping); all arithmetic happens in libgpkl.so.
"""
import ctypes

import torch

from . import _lib
from ._lib import GpklDesc, KERNELS, POSTERIORS, TIERS, FLAG_GRAD_ELL_P, FLAG_PER_PAIR_PRIOR, FLAG_PHILOX_EPS

# One workspace per (device, stream): the C ABI takes a caller-owned workspace and two launches on different streams must
# not share one.  The cache keeps the largest workspace ever requested per stream; release_workspaces() drops them all
# (e.g. between experiments with very different sizes: the C4 workspace is 2.9 GB).
_WS = {}


def release_workspaces():
    """Free every cached workspace (they are re-created on demand)."""
    _WS.clear()


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _make_desc(B, D, T_max, S, total_T, kernel, posterior, noise, flags=0, tier="auto"):
    return GpklDesc(int(B), int(D), int(T_max), int(S), int(total_T), KERNELS[kernel], POSTERIORS[posterior],
                    float(noise), int(flags), TIERS[tier], 0)


def workspace_bytes(desc):
    n = _lib.lib().gpkl_workspace_bytes(ctypes.byref(desc))
    if n == 0:
        _lib.check(_lib.lib().gpkl_forward(ctypes.byref(desc), *([None] * 12), None, 0, None))  # raises the descriptor error
    return n


def _workspace(desc, device):
    n = workspace_bytes(desc)
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws, n


def _is_seed(eps):
    """eps given as a seed (0-d / 1-element int64 CUDA tensor): the kernels generate the noise (GPKL_FLAG_PHILOX_EPS)."""
    return isinstance(eps, torch.Tensor) and eps.dtype == torch.int64


def _check_inputs(mean, times, lengths, ell_q, ell_p, eps, aux, posterior, S):
    if not mean.is_cuda:
        raise RuntimeError("gpkl: tensors must live on a CUDA device (there is no CPU implementation)")
    B, T_max = times.shape
    total_T, D = mean.shape
    assert lengths.shape == (B,) and lengths.dtype == torch.int32, "lengths must be int32 [B]"
    if _is_seed(eps):
        assert eps.numel() == 1 and eps.is_cuda, "the seed is one int64 on the device"
        eps = None
    else:
        assert eps.shape == (B, D, S, T_max), "eps must be [B, D, S, T_max]"
    assert ell_p.shape == (D,)
    if posterior == "gp":
        assert ell_q is not None and ell_q.shape == (D,)
    else:
        assert aux is not None and aux.shape[:2] == (total_T, D)
    for t in (mean, times, ell_q, ell_p, eps, aux):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous()), "float32 contiguous tensors required"
    return B, D, T_max, total_T


def gp_prior_kl_forward(mean, times, lengths, ell_q, ell_p, eps, *, aux=None, kernel="rbf", posterior="gp",
                        noise=1e-3, S=1, tier="auto", want_logdets=False, want_status=False, shared_prior=True):
    """Raw forward through the C ABI (gpkl_forward).  Returns dict(z, kl_sum, kl_pairs[, logdets, status]).
    shared_prior=False forces the per-pair prior factorisation (GPKL_FLAG_PER_PAIR_PRIOR); by default the library
    factors K_p once per sequence when it finds ell_p identical for all latent dims (the reference's prior)."""
    B, D, T_max, total_T = _check_inputs(mean, times, lengths, ell_q, ell_p, eps, aux, posterior, S)
    dev = mean.device
    desc = _make_desc(B, D, T_max, S, total_T, kernel, posterior, noise,
                      (0 if shared_prior else FLAG_PER_PAIR_PRIOR) | (FLAG_PHILOX_EPS if _is_seed(eps) else 0), tier)
    ws, n = _workspace(desc, dev)
    z = torch.empty(S * total_T, D, dtype=torch.float32, device=dev)
    kl_pairs = torch.empty(B * D, dtype=torch.float32, device=dev)
    kl_sum = torch.empty((), dtype=torch.float64, device=dev)
    logdets = torch.empty(B * D, 2, dtype=torch.float32, device=dev) if want_logdets else None
    status = torch.zeros((), dtype=torch.int32, device=dev) if want_status else None
    rc = _lib.lib().gpkl_forward(ctypes.byref(desc), _ptr(mean), _ptr(times), _ptr(lengths), _ptr(ell_q), _ptr(ell_p),
                                 _ptr(aux), _ptr(eps), _ptr(z), _ptr(kl_pairs), _ptr(kl_sum), _ptr(logdets),
                                 _ptr(status), _ptr(ws), n, _stream(dev))
    _lib.check(rc)
    out = {"z": z, "kl_sum": kl_sum, "kl_pairs": kl_pairs}
    if want_logdets:
        out["logdets"] = logdets
    if want_status:
        out["status"] = status
    return out


def gp_prior_kl_backward(mean, times, lengths, ell_q, ell_p, eps, g_z, g_kl_sum=None, g_kl_pairs=None, *, aux=None,
                         kernel="rbf", posterior="gp", noise=1e-3, S=1, tier="auto", grad_ell_p=False,
                         out=None, shared_prior=True):
    """Raw backward through the C ABI (gpkl_backward).  g_kl_sum: 0-d float64 CUDA tensor or None (== 1).
    `out` may carry preallocated g_mean / g_ell_q / g_ell_p / g_aux (e.g. views into a gradient bucket)."""
    B, D, T_max, total_T = _check_inputs(mean, times, lengths, ell_q, ell_p, eps, aux, posterior, S)
    dev = mean.device
    desc = _make_desc(B, D, T_max, S, total_T, kernel, posterior, noise,
                      (FLAG_GRAD_ELL_P if grad_ell_p else 0) | (0 if shared_prior else FLAG_PER_PAIR_PRIOR) |
                      (FLAG_PHILOX_EPS if _is_seed(eps) else 0), tier)
    ws, n = _workspace(desc, dev)
    out = dict(out or {})
    g_mean = out.get("g_mean")
    if g_mean is None:
        g_mean = torch.empty_like(mean)
    g_ell_q = out.get("g_ell_q")
    if g_ell_q is None:
        g_ell_q = torch.empty(D, dtype=torch.float32, device=dev)
    g_ell_p = out.get("g_ell_p")
    if g_ell_p is None and grad_ell_p:
        g_ell_p = torch.empty(D, dtype=torch.float32, device=dev)
    g_aux = out.get("g_aux")
    if g_aux is None and aux is not None:
        g_aux = torch.empty_like(aux)
    if g_z is not None:
        assert g_z.shape == (S * total_T, D) and g_z.dtype == torch.float32 and g_z.is_contiguous()
    if g_kl_sum is not None:
        assert g_kl_sum.dtype == torch.float64 and g_kl_sum.numel() == 1 and g_kl_sum.is_cuda
    if g_kl_pairs is not None:
        assert g_kl_pairs.shape == (B * D,) and g_kl_pairs.dtype == torch.float32 and g_kl_pairs.is_contiguous()
    rc = _lib.lib().gpkl_backward(ctypes.byref(desc), _ptr(mean), _ptr(times), _ptr(lengths), _ptr(ell_q),
                                  _ptr(ell_p), _ptr(aux), _ptr(eps), _ptr(g_z), _ptr(g_kl_sum), _ptr(g_kl_pairs),
                                  _ptr(g_mean), _ptr(g_ell_q), _ptr(g_ell_p), _ptr(g_aux), None, _ptr(ws), n,
                                  _stream(dev))
    _lib.check(rc)
    return {"g_mean": g_mean, "g_ell_q": g_ell_q if posterior == "gp" else None, "g_ell_p": g_ell_p, "g_aux": g_aux}


class GpPriorKL(torch.autograd.Function):
    """(mean, times, lengths, ell_q, ell_p, eps, aux) -> (z, kl_sum, kl_pairs), differentiable in
    mean, ell_q, ell_p and aux.  Factors are recomputed on chip in backward (nothing T x T is saved)."""

    @staticmethod
    def forward(ctx, mean, times, lengths, ell_q, ell_p, eps, aux, kernel, posterior, noise, S, tier):
        o = gp_prior_kl_forward(mean, times, lengths, ell_q, ell_p, eps, aux=aux, kernel=kernel, posterior=posterior,
                                noise=noise, S=S, tier=tier)
        ctx.save_for_backward(mean, times, lengths, ell_q, ell_p, eps, aux)
        ctx.cfg = (kernel, posterior, noise, S, tier)
        ctx.mark_non_differentiable()
        return o["z"], o["kl_sum"], o["kl_pairs"]

    @staticmethod
    def backward(ctx, g_z, g_kl_sum, g_kl_pairs):
        mean, times, lengths, ell_q, ell_p, eps, aux = ctx.saved_tensors
        kernel, posterior, noise, S, tier = ctx.cfg
        grad_ell_p = ctx.needs_input_grad[4]
        if g_kl_sum is None:
            g_kl_sum = torch.zeros((), dtype=torch.float64, device=mean.device)
        g = gp_prior_kl_backward(mean, times, lengths, ell_q, ell_p, eps,
                                 None if g_z is None else g_z.contiguous(), g_kl_sum.to(torch.float64).reshape(()),
                                 None if g_kl_pairs is None else g_kl_pairs.contiguous(), aux=aux, kernel=kernel,
                                 posterior=posterior, noise=noise, S=S, tier=tier, grad_ell_p=grad_ell_p)
        return (g["g_mean"], None, None, g["g_ell_q"] if ctx.needs_input_grad[3] else None,
                g["g_ell_p"] if grad_ell_p else None, None, g["g_aux"] if aux is not None else None,
                None, None, None, None, None)


def philox_normal(seed, n, device="cuda:0"):
    """The first n draws of the GPKL_FLAG_PHILOX_EPS stream for `seed`, materialised (gpkl_philox_normal)."""
    dev = torch.device(device)
    sd = seed if isinstance(seed, torch.Tensor) else torch.tensor([int(seed)], dtype=torch.int64, device=dev)
    out = torch.empty(int(n), dtype=torch.float32, device=dev)
    _lib.check(_lib.lib().gpkl_philox_normal(_ptr(sd), int(n), _ptr(out), _stream(dev)))
    return out


def gp_prior_kl(mean, times, lengths, ell_q, ell_p, eps=None, *, aux=None, kernel="rbf", posterior="gp", noise=1e-3,
                S=1, tier="auto", generator=None, seed=None):
    """Fused GP-prior path.  Returns (z [S*sum_T, D] f32, kl_sum f64 scalar, kl_pairs [B*D] f32).

    eps [B, D, S, T_max]: explicit N(0,1) noise (reproducible parity runs).  eps=None: the noise of the reference's
    tf.random_normal inside tf_kernel (Full_GP_VAE_dynamic_time.py:166) is generated INSIDE the kernels from a 64-bit seed
    (Philox4x32-10 + Box-Muller, no eps tensor in HBM; forward and backward regenerate the same values): `seed` (int or
    1-element int64 CUDA tensor), else one drawn from `generator` / torch's default generator."""
    if eps is None:
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator).item())
        eps = seed if isinstance(seed, torch.Tensor) else torch.tensor([int(seed)], dtype=torch.int64, device=mean.device)
    if posterior != "gp" and ell_q is None:
        ell_q = ell_p
    return GpPriorKL.apply(mean, times, lengths, ell_q, ell_p, eps, aux, kernel, posterior, float(noise), int(S), tier)


class HostStep:
    """Forward+backward with HOST (pinned) buffers through gpkl_step_host: the end-to-end call a caller
    holding CPU tensors makes.  Owns the device staging area and pinned result buffers."""

    def __init__(self, B, D, T_max, S, total_T, *, kernel="rbf", posterior="gp", noise=1e-3, grad_ell_p=False,
                 tier="auto", device="cuda:0", shared_prior=True):
        self.device = torch.device(device)
        self.desc = _make_desc(B, D, T_max, S, total_T, kernel, posterior, noise,
                               (FLAG_GRAD_ELL_P if grad_ell_p else 0) | (0 if shared_prior else FLAG_PER_PAIR_PRIOR), tier)
        n = _lib.lib().gpkl_step_host_bytes(ctypes.byref(self.desc))
        if n == 0:
            raise RuntimeError("gpkl: bad descriptor")
        self.nbytes = n
        self.staging = torch.empty(n, dtype=torch.uint8, device=self.device)
        pin = dict(pin_memory=True)
        self.z = torch.empty(S * total_T, D, dtype=torch.float32, **pin)
        self.kl_pairs = torch.empty(B * D, dtype=torch.float32, **pin)
        self.kl_sum = torch.empty((), dtype=torch.float64, **pin)
        self.g_mean = torch.empty(total_T, D, dtype=torch.float32, **pin)
        self.g_ell_q = torch.empty(D, dtype=torch.float32, **pin)
        self.g_ell_p = torch.empty(D, dtype=torch.float32, **pin) if grad_ell_p else None
        self.g_aux = torch.empty(total_T, D, dtype=torch.float32, **pin) if posterior == "diag" else None
        self.posterior = posterior
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def __call__(self, mean, times, lengths, ell_q, ell_p, eps, g_z=None, aux=None, *, full_outputs=True):
        """All arguments are CPU tensors (pinned for async copies).  Results land in self.* (pinned) once the
        current stream is synchronised.  full_outputs=False reads back only kl_sum and the lengthscale grads."""
        for t in (mean, times, lengths, ell_q, ell_p, eps, g_z, aux):
            assert t is None or (not t.is_cuda and t.is_contiguous())
        z = self.z if full_outputs else None
        klp = self.kl_pairs if full_outputs else None
        gm = self.g_mean if full_outputs else None
        ga = self.g_aux if full_outputs else None
        rc = _lib.lib().gpkl_step_host(ctypes.byref(self.desc), _ptr(mean), _ptr(times), _ptr(lengths), _ptr(ell_q),
                                       _ptr(ell_p), _ptr(aux), _ptr(eps), _ptr(g_z), _ptr(z), _ptr(klp),
                                       _ptr(self.kl_sum), _ptr(gm), _ptr(self.g_ell_q), _ptr(self.g_ell_p), _ptr(ga),
                                       _ptr(self.staging), self.nbytes, _stream(self.device))
        _lib.check(rc)
        ins = [mean, times, lengths, ell_q, ell_p, eps, g_z, aux]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in ins if t is not None)
        outs = [z, klp, self.kl_sum, gm, self.g_ell_q if self.posterior == "gp" else None, self.g_ell_p, ga]
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in outs if t is not None)
        return self


# ---- reconstruction term + beta-weighted loss ---------------------------------------------
def _recon_ws(B, device):
    n = _lib.lib().gpkl_recon_workspace_bytes(int(B))
    key = ("recon", device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 1 << 16), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws, n


class BernoulliRecon(torch.autograd.Function):
    """recon(x, x_decode) of Full_GP_VAE_dynamic_time.py:349-356 as one streaming CUDA pass (float64 scalar)."""

    @staticmethod
    def forward(ctx, x, x_decode, lengths, S):
        if not x.is_cuda:
            raise RuntimeError("gpkl: tensors must live on a CUDA device (there is no CPU implementation)")
        total_T, F = x.shape
        B = lengths.shape[0]
        assert x_decode.shape == (S * total_T, F) and lengths.dtype == torch.int32
        assert x.dtype == torch.float32 and x_decode.dtype == torch.float32 and x.is_contiguous() and x_decode.is_contiguous()
        out = torch.empty((), dtype=torch.float64, device=x.device)
        ws, n = _recon_ws(B, x.device)
        _lib.check(_lib.lib().gpkl_recon_forward(B, F, S, total_T, _ptr(x), _ptr(x_decode), _ptr(lengths), _ptr(out),
                                                 _ptr(ws), n, _stream(x.device)))
        ctx.save_for_backward(x, x_decode, lengths)
        ctx.S = S
        return out

    @staticmethod
    def backward(ctx, g):
        x, x_decode, lengths = ctx.saved_tensors
        total_T, F = x.shape
        B = lengths.shape[0]
        gx = torch.empty_like(x_decode)
        ws, n = _recon_ws(B, x.device)
        g = g.to(torch.float64).reshape(()).contiguous()
        _lib.check(_lib.lib().gpkl_recon_backward(B, F, ctx.S, total_T, _ptr(x), _ptr(x_decode), _ptr(lengths), _ptr(g),
                                                  _ptr(gx), _ptr(ws), n, _stream(x.device)))
        return None, gx, None, None


def bernoulli_recon(x, x_decode, lengths, S=1):
    """sum_recon_loss of the reference (Full_GP_VAE_dynamic_time.py:349-356): Bernoulli NLL with the 1e-10 guards,
    mean over the S samples, sum over time and batch; differentiable in x_decode."""
    return BernoulliRecon.apply(x, x_decode, lengths, int(S))


def elbo_loss(x, x_decode, lengths, kl_sum, beta=1.0, S=1):
    """loss = recon + beta * KL  (Full_GP_VAE_dynamic_time.py:358-360; beta warm-up: syndata/GP_VAE_syn_data.py:361-364)."""
    return bernoulli_recon(x, x_decode, lengths, S) + beta * kl_sum


# ---- GP-recognition sampler ----------------------------------------------------------------
def _named_ws(name, n, device):
    key = (name, device, torch.cuda.current_stream(device).cuda_stream)
    ws = _WS.get(key)
    if ws is None or ws.numel() < n:
        ws = torch.empty(max(n, 1 << 16), dtype=torch.uint8, device=device)
        _WS[key] = ws
    return ws


def _recog_setup(mean, logvar, times, lengths, ell, eps, kernel, noise, S, tier):
    if not mean.is_cuda:
        raise RuntimeError("gpkl: tensors must live on a CUDA device (there is no CPU implementation)")
    B, T_max = times.shape
    total_T, D = mean.shape
    assert logvar.shape == (total_T, D) and ell.shape == (D,) and eps.shape == (B, D, S, T_max)
    assert lengths.shape == (B,) and lengths.dtype == torch.int32
    for t in (mean, logvar, times, ell, eps):
        assert t.dtype == torch.float32 and t.is_contiguous(), "float32 contiguous tensors required"
    desc = _make_desc(B, D, T_max, S, total_T, kernel, "gp", noise, 0, tier)
    n = _lib.lib().gpkl_recog_workspace_bytes(ctypes.byref(desc))
    if n == 0:
        raise RuntimeError("gpkl: inconsistent descriptor")
    return desc, _named_ws("recog", n, mean.device), n


class GpRecogSample(torch.autograd.Function):
    """(mean, logvar, times, lengths, ell, eps) -> (z, kl_sum, kl_rows): the sampler and KL of GP_recog_VAE_prior.py
    (:65-70, :137-168, :170-191), differentiable in mean, logvar and ell (approx_time_chars, :81)."""

    @staticmethod
    def forward(ctx, mean, logvar, times, lengths, ell, eps, kernel, noise, S, tier):
        desc, ws, n = _recog_setup(mean, logvar, times, lengths, ell, eps, kernel, noise, S, tier)
        dev = mean.device
        z = torch.empty(S * mean.shape[0], mean.shape[1], dtype=torch.float32, device=dev)
        kl_rows = torch.empty(mean.shape[0], dtype=torch.float32, device=dev)
        kl_sum = torch.empty((), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib().gpkl_recog_forward(ctypes.byref(desc), _ptr(mean), _ptr(logvar), _ptr(times), _ptr(lengths),
                                                 _ptr(ell), _ptr(eps), _ptr(z), _ptr(kl_rows), _ptr(kl_sum), None,
                                                 _ptr(ws), n, _stream(dev)))
        ctx.save_for_backward(mean, logvar, times, lengths, ell, eps)
        ctx.cfg = (kernel, noise, S, tier)
        return z, kl_sum, kl_rows

    @staticmethod
    def backward(ctx, g_z, g_kl_sum, g_kl_rows):
        mean, logvar, times, lengths, ell, eps = ctx.saved_tensors
        kernel, noise, S, tier = ctx.cfg
        desc, ws, n = _recog_setup(mean, logvar, times, lengths, ell, eps, kernel, noise, S, tier)
        dev = mean.device
        if g_kl_sum is None:
            g_kl_sum = torch.zeros((), dtype=torch.float64, device=dev)
        g_kl_sum = g_kl_sum.to(torch.float64).reshape(()).contiguous()
        g_z = None if g_z is None else g_z.contiguous()
        g_kl_rows = None if g_kl_rows is None else g_kl_rows.contiguous()
        g_mean, g_logvar = torch.empty_like(mean), torch.empty_like(logvar)
        g_ell = torch.empty_like(ell)
        _lib.check(_lib.lib().gpkl_recog_backward(ctypes.byref(desc), _ptr(mean), _ptr(logvar), _ptr(times), _ptr(lengths),
                                                  _ptr(ell), _ptr(eps), _ptr(g_z), _ptr(g_kl_sum), _ptr(g_kl_rows),
                                                  _ptr(g_mean), _ptr(g_logvar), _ptr(g_ell), None, _ptr(ws), n,
                                                  _stream(dev)))
        return g_mean, g_logvar, None, None, g_ell, None, None, None, None, None


def gp_recog_sample(mean, logvar, times, lengths, ell, eps=None, *, kernel="rbf", noise=1e-3, S=1, tier="auto",
                    generator=None):
    """GP-recognition sampler: z = mean + (chol(K(times, ell)) + diag(sqrt(exp(logvar)))) eps and the standard KL of the
    rows (GP_recog_VAE_prior.py:274-284).  Returns (z [S*sum_T, D] f32, kl_sum f64 scalar, kl_rows [sum_T] f32)."""
    if eps is None:
        B, T_max = times.shape
        eps = torch.randn(B, mean.shape[1], S, T_max, device=mean.device, dtype=torch.float32, generator=generator)
    return GpRecogSample.apply(mean, logvar, times, lengths, ell, eps, kernel, float(noise), int(S), tier)


# ---- ragged batch producer -----------------------------------------------------------------
def collate_batch(data, time_grid, index=None, max_time=None, batch_size=None):
    """GPU-side SyntheticDataHandler._prep_dataset + data_batch (DataHandler.py:111-156) for one batch.
    data [N, F, T_full] f32 CUDA with -1 at missing time points, time_grid [T_full], index [B] int32 (None: the first
    batch_size sequences).  Returns (x [sum_T, F], times [B, max_time], lengths [B] int32) -- the feed of
    Full_GP_VAE_dynamic_time.py:377-378.  One host read (sum_T) sizes the returned view of x."""
    if not data.is_cuda:
        raise RuntimeError("gpkl: tensors must live on a CUDA device (there is no CPU implementation)")
    N, F, T_full = data.shape
    max_time = T_full if max_time is None else int(max_time)
    B = int(index.shape[0]) if index is not None else int(N if batch_size is None else batch_size)
    assert data.dtype == torch.float32 and data.is_contiguous() and time_grid.shape == (T_full,)
    assert time_grid.dtype == torch.float32 and time_grid.is_contiguous()
    assert index is None or (index.dtype == torch.int32 and index.is_contiguous())
    dev = data.device
    x = torch.empty(B * min(T_full, max_time), F, dtype=torch.float32, device=dev)
    times = torch.empty(B, max_time, dtype=torch.float32, device=dev)
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    total = torch.zeros((), dtype=torch.int64, device=dev)
    n = _lib.lib().gpkl_collate_workspace_bytes(B, max_time)
    ws = _named_ws("collate", n, dev)
    _lib.check(_lib.lib().gpkl_collate(N, F, T_full, B, max_time, _ptr(data), _ptr(time_grid), _ptr(index), _ptr(x),
                                       _ptr(times), _ptr(lengths), _ptr(total), _ptr(ws), n, _stream(dev)))
    return x[: int(total.item())], times, lengths


# ---- GP posterior imputation ----------------------------------------------------------------
def gp_posterior_impute(z_obs, t_obs, n_obs, t_full, eps=None, *, kernel="rbf", ell=1.0, noise=1e-3, want_status=False):
    """Predictive mean (eps=None) or one sample (eps [B, D, n_full]) of every latent row of every sequence on the full time
    grid t_full [B, n_full], given the rows' values z_obs [sum n_obs, D] at the observed times t_obs [B, n_obs_max]
    (n_obs [B] int32 valid per sequence): sample_given_part_latent / post_gp_sample of the reference
    (FullGP_and_GPdecoder_dynamic_time_analysis.py:40-56, :96-111) for a whole batch.  Returns out [B*n_full, D]
    (and the device status counter: > 0 when the posterior covariance was not positive definite, where the reference raises)."""
    if not z_obs.is_cuda:
        raise RuntimeError("gpkl: tensors must live on a CUDA device (there is no CPU implementation)")
    B, nd_max = t_obs.shape
    Bf, ns = t_full.shape
    D = z_obs.shape[1]
    assert Bf == B and n_obs.shape == (B,) and n_obs.dtype == torch.int32
    for t in (z_obs, t_obs, t_full, eps):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous()), "float32 contiguous tensors required"
    assert eps is None or eps.shape == (B, D, ns)
    dev = z_obs.device
    out = torch.empty(B * ns, D, dtype=torch.float32, device=dev)
    status = torch.zeros((), dtype=torch.int32, device=dev)
    n = _lib.lib().gpkl_impute_workspace_bytes(B)
    ws = _named_ws("impute", n, dev)
    _lib.check(_lib.lib().gpkl_impute(B, D, nd_max, ns, KERNELS[kernel], float(ell), float(noise), _ptr(z_obs), _ptr(t_obs),
                                      _ptr(n_obs), _ptr(t_full), _ptr(eps), _ptr(out), _ptr(status), _ptr(ws), n,
                                      _stream(dev)))
    return (out, status) if want_status else out
