"""GPU: size-independent properties at BASELINE.json's full sizes (the oracle would take minutes there).
C2 = T=48, D=35, B=256 (8,960 pairs); C1 = T=10, D=256, B=64; C4's sequence length T=512 on a slice of its batch."""
import pytest
import torch

import gp_kl_oracle as orc

pytestmark = pytest.mark.gpu


def _run(c, dev, **kw):
    import gpkl
    d = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    f = gpkl.gp_prior_kl_forward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], **kw)
    b = gpkl.gp_prior_kl_backward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], d["g_z"], **kw)
    torch.cuda.synchronize()
    return f, b


@pytest.mark.parametrize("B,D,T,kernel", [(256, 35, 48, "rbf"), (64, 256, 10, "cauchy"), (4, 64, 512, "cauchy")])
def test_full_size_invariants(cuda_device, B, D, T, kernel):
    dev = cuda_device
    c = orc.synthetic_batch(B, D, T, 1, seed=1234)
    f, b = _run(c, dev, kernel=kernel)
    # finite, KL >= 0 for every pair, the scalar is the float64 sum of the per-pair values
    assert torch.isfinite(f["kl_pairs"]).all() and torch.isfinite(f["z"]).all() and torch.isfinite(b["g_mean"]).all()
    assert float(f["kl_pairs"].min()) >= 0.0
    assert abs(float(f["kl_sum"]) - float(f["kl_pairs"].double().sum())) <= 1e-9 * abs(float(f["kl_sum"]))
    # deterministic: a second run is bit-identical (fixed-order reductions, no atomics on the data path)
    f2, b2 = _run(c, dev, kernel=kernel)
    assert torch.equal(f["kl_pairs"], f2["kl_pairs"]) and torch.equal(f["z"], f2["z"])
    assert torch.equal(b["g_mean"], b2["g_mean"]) and torch.equal(b["g_ell_q"], b2["g_ell_q"])
    # permuting the sequences permutes the per-pair outputs and leaves the summed lengthscale gradient unchanged
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(3))
    cp = dict(c)
    cp["times"], cp["eps"], cp["lengths"] = c["times"][perm], c["eps"][perm], c["lengths"][perm]
    cp["mean"] = c["mean"].reshape(B, T, D)[perm].reshape(B * T, D).contiguous()
    cp["g_z"] = c["g_z"].reshape(B, T, D)[perm].reshape(B * T, D).contiguous()
    fp, bp = _run(cp, dev, kernel=kernel)
    assert torch.equal(fp["kl_pairs"].reshape(B, D), f["kl_pairs"].reshape(B, D)[perm.to(dev)])
    assert torch.equal(fp["z"].reshape(B, T, D), f["z"].reshape(B, T, D)[perm.to(dev)])
    assert torch.allclose(bp["g_ell_q"], b["g_ell_q"], rtol=2e-5, atol=1e-6 * float(b["g_ell_q"].abs().max()))


@pytest.mark.parametrize("B,D,T", [(256, 35, 48), (8, 16, 160), (2, 8, 512)])
def test_identical_prior_and_posterior_gives_zero_kl(cuda_device, B, D, T):
    """q = p (same lengthscale, zero mean): KL = 0 per pair, and its gradient w.r.t. l_q vanishes."""
    c = orc.synthetic_batch(B, D, T, 1, seed=5)
    c["ell_q"] = c["ell_p"].clone()
    c["mean"] = torch.zeros_like(c["mean"])
    c["g_z"] = torch.zeros_like(c["g_z"])
    f, b = _run(c, cuda_device)
    assert float(f["kl_pairs"].abs().max()) < 1e-4 * T
    assert float(b["g_mean"].abs().max()) == 0.0
    assert float(b["g_ell_q"].abs().max()) < 2e-3 * B * T


@pytest.mark.parametrize("T,Tpad", [(10, 16), (48, 64), (100, 144), (200, 272)])
def test_padding_invariance(cuda_device, T, Tpad):
    """A batch stored with a larger T_max (zero padded times / eps, DataHandler.py:151) gives the same result,
    although it may be served by a different tier."""
    B, D = 6, 5
    c = orc.synthetic_batch(B, D, T, 1, ragged=True, seed=7)
    f, b = _run(c, cuda_device)
    cp = dict(c)
    cp["times"] = torch.zeros(B, Tpad)
    cp["times"][:, :T] = c["times"]
    cp["eps"] = torch.zeros(B, D, 1, Tpad)
    cp["eps"][..., :T] = c["eps"]
    fp, bp = _run(cp, cuda_device)

    def close(a, b_, tol):
        return float((a.double() - b_.double()).abs().max()) <= tol * float(b_.double().abs().max())
    assert close(fp["kl_pairs"], f["kl_pairs"], 2e-5) and close(fp["z"], f["z"], 1e-5)
    assert close(bp["g_mean"], b["g_mean"], 1e-4) and close(bp["g_ell_q"], b["g_ell_q"], 1e-4)


def test_sample_is_linear_in_eps(cuda_device):
    """z = m + L_q eps: z(eps1 + eps2) - m = (z(eps1) - m) + (z(eps2) - m)  (C2 size)."""
    c = orc.synthetic_batch(256, 35, 48, 1, seed=9)
    g = torch.Generator().manual_seed(10)
    e2 = torch.randn(c["eps"].shape, generator=g)
    z1 = _run(c, cuda_device)[0]["z"]
    z2 = _run(dict(c, eps=e2), cuda_device)[0]["z"]
    z12 = _run(dict(c, eps=c["eps"] + e2), cuda_device)[0]["z"]
    m = c["mean"].to(cuda_device)
    assert float(((z12 - m) - (z1 - m) - (z2 - m)).abs().max()) < 2e-5
