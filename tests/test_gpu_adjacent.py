"""GPU: the rows either side of the path through the C ABI -- GP-recognition sampler
(GP_recog_VAE_prior.py) and ragged batch producer (DataHandler.py) -- vs the reference-executed fixtures and the oracle."""
import numpy as np
import pytest
import torch

import gp_kl_oracle as orc
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-5        # z / KL (north_star tolerance of the path)
TOL_GRAD = 1e-4


def _run_recog(case, dev, S, kernel="rbf", tier="auto"):
    import gpkl
    mean = case["mean"].to(dev).requires_grad_(True)
    logvar = case["logvar"].to(dev).requires_grad_(True)
    ell = case["ell"].to(dev).requires_grad_(True)
    z, kl_sum, kl_rows = gpkl.gp_recog_sample(mean, logvar, case["times"].to(dev), case["lengths"].to(dev).to(torch.int32),
                                              ell, case["eps"].to(dev), kernel=kernel, S=S, tier=tier)
    loss = float(case["g_kl_sum"]) * kl_sum + (case["g_kl_rows"].to(dev).double() * kl_rows.double()).sum() + \
        (case["g_z"].to(dev).double() * z.double()).sum()
    loss.backward()
    torch.cuda.synchronize()
    return z, kl_sum, kl_rows, mean.grad, logvar.grad, ell.grad


@pytest.mark.parametrize("name", ["g7_recog_ragged_s2", "g8_recog_grid"])
def test_recog_golden(cuda_device, name):
    g = load_golden(name)
    z, kl_sum, kl_rows, gm, glv, gl = _run_recog(g, cuda_device, g["S"])
    assert rel_err(z, g["z"]) < TOL
    assert rel_err(kl_rows, g["kl_rows"]) < TOL
    assert abs(float(kl_sum) - float(g["kl_sum"])) < TOL * abs(float(g["kl_sum"]))
    assert rel_err(gm, g["g_mean"]) < TOL_GRAD
    assert rel_err(glv, g["g_logvar"]) < TOL_GRAD
    assert rel_err(gl, g["g_ell"]) < TOL_GRAD


def _recog_case(B, D, T, S, ragged, seed, grid=True):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.randint((T + 1) // 2, T + 1, (B,), generator=g, dtype=torch.int32) if ragged else torch.full((B,), T, dtype=torch.int32)
    if B > 1 and ragged:
        lengths[0] = T
    times = torch.arange(T, dtype=torch.float32).repeat(B, 1) if grid else torch.cumsum(torch.rand(B, T, generator=g) + 0.5, 1)
    for b in range(B):
        times[b, int(lengths[b]):] = 0
    total = int(lengths.sum())
    return {"mean": torch.randn(total, D, generator=g), "logvar": 0.5 * torch.randn(total, D, generator=g) - 0.5,
            "times": times.contiguous(), "lengths": lengths, "ell": torch.exp(0.2 * torch.randn(D, generator=g)),
            "eps": torch.randn(B, D, S, T, generator=g), "g_z": torch.randn(S * total, D, generator=g),
            "g_kl_rows": torch.randn(total, generator=g), "g_kl_sum": 0.6}


@pytest.mark.parametrize("B,D,T,S,ragged,kernel,tier", [
    (1, 1, 1, 1, False, "rbf", "auto"), (3, 5, 7, 2, True, "rbf", "auto"), (4, 35, 48, 1, True, "rbf", "auto"),
    (5, 100, 20, 1, False, "rbf", "auto"), (2, 3, 100, 1, True, "rbf", "auto"), (2, 2, 200, 2, False, "cauchy", "auto"),
    (3, 4, 33, 1, True, "cauchy", "generic"), (2, 6, 64, 3, False, "rbf", "warp"), (2, 3, 80, 1, True, "rbf", "block")])
def test_recog_vs_oracle(cuda_device, B, D, T, S, ragged, kernel, tier):
    case = _recog_case(B, D, T, S, ragged, 1000 * B + T)
    z, kl_sum, kl_rows, gm, glv, gl = _run_recog(case, cuda_device, S, kernel=kernel, tier=tier)
    out, grads = orc.gp_recog_grads(case["mean"], case["logvar"], case["times"], case["lengths"], case["ell"], case["eps"],
                                    case["g_z"], case["g_kl_sum"], case["g_kl_rows"], kernel=kernel, S=S)
    assert rel_err(z, out["z"]) < TOL
    assert rel_err(kl_rows, out["kl_rows"]) < TOL
    assert abs(float(kl_sum) - float(out["kl_sum"])) < TOL * abs(float(out["kl_sum"]))
    assert rel_err(gm, grads["mean"]) < TOL_GRAD
    assert rel_err(glv, grads["logvar"]) < TOL_GRAD
    assert rel_err(gl, grads["ell"]) < TOL_GRAD


def test_recog_empty_and_zero_length(cuda_device):
    import gpkl
    dev = cuda_device
    case = _recog_case(3, 2, 6, 1, False, 5)
    case["lengths"][1] = 0
    keep = torch.cat([torch.arange(0, 6), torch.arange(12, 18)])
    for k in ("mean", "logvar", "g_z"):
        case[k] = case[k][keep].contiguous()
    case["g_kl_rows"] = case["g_kl_rows"][keep].contiguous()
    z, kl_sum, kl_rows, gm, glv, gl = _run_recog(case, dev, 1)
    out, grads = orc.gp_recog_grads(case["mean"], case["logvar"], case["times"], case["lengths"], case["ell"], case["eps"],
                                    case["g_z"], case["g_kl_sum"], case["g_kl_rows"], S=1)
    assert rel_err(z, out["z"]) < TOL and rel_err(gl, grads["ell"]) < TOL_GRAD and rel_err(glv, grads["logvar"]) < TOL_GRAD
    z, kl_sum, kl_rows = gpkl.gp_recog_sample(torch.zeros(0, 2, device=dev), torch.zeros(0, 2, device=dev),
                                              torch.zeros(0, 4, device=dev), torch.zeros(0, dtype=torch.int32, device=dev),
                                              torch.ones(2, device=dev), torch.zeros(0, 2, 1, 4, device=dev))
    assert z.shape == (0, 2) and float(kl_sum) == 0.0


def test_collate_golden(cuda_device):
    import gpkl
    dev = cuda_device
    g = load_golden("g9_collate")
    data, grid, mt = g["data"].to(dev), g["time_grid"].to(dev), int(g["max_time"])
    for k in range(2):
        x, times, lengths = gpkl.collate_batch(data, grid, g["index%d" % k].to(dev).to(torch.int32), mt)
        assert torch.equal(x.cpu(), g["x%d" % k]) and torch.equal(times.cpu(), g["times%d" % k])
        assert torch.equal(lengths.cpu(), g["lengths%d" % k])


@pytest.mark.parametrize("N,F,T_full,max_time,B,p_drop", [(1, 1, 1, 1, 1, 0.0), (5, 15, 45, 45, 5, 0.3), (9, 35, 48, 60, 4, 0.5),
                                                          (6, 3, 100, 40, 6, 0.2), (4, 70, 33, 33, 3, 1.0),
                                                          (300, 35, 48, 48, 256, 0.4)])
def test_collate_vs_oracle(cuda_device, N, F, T_full, max_time, B, p_drop):
    """Bit exact (pure data movement); includes all-missing sequences (length 0), truncation at max_time, F not a
    multiple of the tile, a permuted batch index."""
    import gpkl
    dev = cuda_device
    rng = np.random.RandomState(N * 7 + F)
    data = rng.rand(N, F, T_full).astype(np.float32)
    for i in range(N):
        data[i][:, rng.rand(T_full) < p_drop] = -1.0
    grid = np.cumsum(rng.rand(T_full) + 0.5).astype(np.float32)
    index = rng.permutation(N)[:B].astype(np.int32)
    xo, to, lo = orc.collate_batch(data, grid, index, max_time)
    x, times, lengths = gpkl.collate_batch(torch.from_numpy(data).to(dev), torch.from_numpy(grid).to(dev),
                                           torch.from_numpy(index).to(dev), max_time)
    assert x.shape == xo.shape and (x.cpu().numpy() == xo).all()
    assert (times.cpu().numpy() == to).all() and (lengths.cpu().numpy() == lo).all()
    # the collated batch feeds the path directly
    if int(lengths.min()) > 0 and F <= 64:
        ell = torch.ones(F, device=dev)
        z, kl_sum, _ = gpkl.gp_prior_kl(x.contiguous(), times, lengths, ell, ell)
        assert z.shape == x.shape and torch.isfinite(kl_sum)


def test_recog_reference_call_sites(cuda_device):
    """GP_recog_VAE_prior.main()'s three calls by name (reference_api.GPRecogPath) reproduce golden G7."""
    import gpkl
    dev = cuda_device
    g = load_golden("g7_recog_ragged_s2")
    B, D, S = g["times"].shape[0], g["mean"].shape[1], g["S"]
    path = gpkl.GPRecogPath(D, device=dev)
    with torch.no_grad():
        path.approx_time_chars.copy_(g["ell"].to(dev))
    mean, logvar = g["mean"].to(dev), g["logvar"].to(dev)
    sizes = g["lengths"].to(dev)
    approx_kernel, chol_noise, chars = path.approx_kernels(g["times"].to(dev), sizes, D, B, S, logvar, eps=g["eps"].to(dev))
    z = path.gp_vae_sample(mean, chol_noise, sizes, B, S, D)
    kl = path.standard_vae_kl(mean, logvar, D)          # the reference's sign: minus the KL (:69)
    assert rel_err(z, g["z"]) < TOL and rel_err(-kl, g["kl_rows"]) < TOL


def test_data_handler_gpu(cuda_device):
    """SyntheticDataHandlerGPU.data_batch returns the reference handler's first two batches (golden G9)."""
    import gpkl
    g = load_golden("g9_collate")
    h = gpkl.SyntheticDataHandlerGPU({"x": g["data"][:6], "time": g["time_grid"]}, int(g["max_time"]), batch_size=3,
                                     device=cuda_device)
    for k in range(2):
        x, times, lengths = h.data_batch("train")
        assert torch.equal(x.cpu(), g["x%d" % k]) and torch.equal(times.cpu(), g["times%d" % k])
        assert torch.equal(lengths.cpu(), g["lengths%d" % k])
    x, times, lengths = h.data_batch("train")            # third call wraps around: reshuffled epoch
    assert x.shape[0] == int(lengths.sum())
