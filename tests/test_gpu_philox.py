"""In-kernel noise (GPKL_FLAG_PHILOX_EPS): the device stream against the host restatement of Philox4x32-10 + Box-Muller
(oracle/philox_ref.py, pinned to the Random123 known-answer vectors), and the seeded op against the explicit-eps op in
every tier."""
import numpy as np
import pytest
import torch

import gp_kl_oracle as orc
import philox_ref
from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 1234, 2 ** 40 + 17, 2 ** 63 - 1])
def test_device_stream_matches_host_philox(cuda_device, seed):
    import gpkl
    n = 100003
    got = gpkl.philox_normal(seed, n, device=cuda_device).cpu().numpy()
    want = philox_ref.philox_normal(seed, n)
    # same uint32 draws; logf / sincosf differ from numpy's in the last ulp
    assert np.abs(got - want).max() < 2e-6 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("tier,T,S", [("warp", 12, 2), ("warp", 48, 1), ("block", 100, 2), ("block", 160, 1), ("auto", 300, 1),
                                      ("auto", 512, 1), ("generic", 33, 3)])
def test_seeded_op_equals_explicit_eps(cuda_device, tier, T, S):
    """The op with a seed reproduces, bit for bit, the op fed with the materialised stream (forward and backward)."""
    import gpkl
    dev = cuda_device
    B, D = 3, 4
    case = orc.synthetic_batch(B, D, T, S, ragged=True, seed=800 + T, grid=True)
    c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    seed = torch.tensor([987654321 + T], dtype=torch.int64, device=dev)
    eps = gpkl.philox_normal(seed, B * D * S * T, device=dev).reshape(B, D, S, T)
    a = (c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"])
    f0 = gpkl.gp_prior_kl_forward(*a, eps, S=S, tier=tier)
    f1 = gpkl.gp_prior_kl_forward(*a, seed, S=S, tier=tier)
    b0 = gpkl.gp_prior_kl_backward(*a, eps, c["g_z"], S=S, tier=tier)
    b1 = gpkl.gp_prior_kl_backward(*a, seed, c["g_z"], S=S, tier=tier)
    torch.cuda.synchronize()
    assert torch.equal(f0["z"], f1["z"]) and torch.equal(f0["kl_pairs"], f1["kl_pairs"])
    assert torch.equal(b0["g_mean"], b1["g_mean"]) and torch.equal(b0["g_ell_q"], b1["g_ell_q"])


def test_autograd_op_draws_in_kernel(cuda_device):
    """gp_prior_kl(eps=None): no eps tensor; same seed -> same sample, gradients flow, oracle agrees on the host stream."""
    import gpkl
    dev = cuda_device
    B, D, T = 4, 5, 20
    case = orc.synthetic_batch(B, D, T, 1, ragged=False, seed=77, grid=True)
    c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    mean = c["mean"].clone().requires_grad_(True)
    lq = c["ell_q"].clone().requires_grad_(True)
    z1, kl1, _ = gpkl.gp_prior_kl(mean, c["times"], c["lengths"], lq, c["ell_p"], seed=42)
    z2, _, _ = gpkl.gp_prior_kl(mean, c["times"], c["lengths"], lq, c["ell_p"], seed=42)
    z3, _, _ = gpkl.gp_prior_kl(mean, c["times"], c["lengths"], lq, c["ell_p"], seed=43)
    (kl1 + (c["g_z"].double() * z1.double()).sum()).backward()
    torch.cuda.synchronize()
    assert torch.equal(z1, z2) and not torch.equal(z1, z3)
    eps = torch.from_numpy(philox_ref.philox_normal(42, B * D * T)).reshape(B, D, 1, T)
    out, grads = orc.gp_prior_kl_grads(case["mean"], case["times"], case["lengths"], case["ell_q"], case["ell_p"], eps,
                                       case["g_z"])
    assert rel_err(z1, out["z"]) < 1e-5 and rel_err(mean.grad, grads["mean"]) < 1e-4
    assert rel_err(lq.grad, grads["ell_q"]) < 1e-4


@pytest.mark.parametrize("T,S", [(9, 3), (48, 1), (64, 2)])
def test_seeded_op_equals_explicit_eps_v3_hot_tier(cuda_device, T, S):
    """The V3 hot tier (gpkl_bidiag.cu) draws its noise in the kernel too: seed == materialised stream, bit for bit."""
    import gpkl
    dev = cuda_device
    B, D = 3, 4
    case = orc.synthetic_batch(B, D, T, S, ragged=True, seed=850 + T, posterior="bidiag", grid=True)
    c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    seed = torch.tensor([13579 + T], dtype=torch.int64, device=dev)
    eps = gpkl.philox_normal(seed, B * D * S * T, device=dev).reshape(B, D, S, T)
    a = (c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"])
    kw = dict(aux=c["aux"], posterior="bidiag", S=S, tier="auto")
    f0 = gpkl.gp_prior_kl_forward(*a, eps, **kw)
    f1 = gpkl.gp_prior_kl_forward(*a, seed, **kw)
    b0 = gpkl.gp_prior_kl_backward(*a, eps, c["g_z"], **kw)
    b1 = gpkl.gp_prior_kl_backward(*a, seed, c["g_z"], **kw)
    torch.cuda.synchronize()
    assert torch.equal(f0["z"], f1["z"]) and torch.equal(f0["kl_pairs"], f1["kl_pairs"])
    assert torch.equal(b0["g_mean"], b1["g_mean"]) and torch.equal(b0["g_aux"], b1["g_aux"])
