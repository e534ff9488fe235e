"""GPU: reconstruction term + beta-weighted loss through the C ABI vs the reference-executed
fixture and the oracle."""
import pytest
import torch

import gp_kl_oracle as orc
from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def test_recon_golden(cuda_device):
    import gpkl
    dev = cuda_device
    g = load_golden("g6_recon_loss")
    xd = g["x_decode"].to(dev).requires_grad_(True)
    kl = torch.tensor(g["kl"], dtype=torch.float64, device=dev)
    loss = gpkl.elbo_loss(g["x"].to(dev), xd, g["lengths"].to(dev), kl, beta=g["beta"], S=g["S"])
    loss.backward()
    assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert rel_err(xd.grad, g["g_x_decode"]) < 1e-5


@pytest.mark.parametrize("B,T,F,S,ragged", [(1, 1, 1, 1, False), (3, 5, 7, 2, True), (4, 8, 64, 1, False), (5, 20, 4096, 1, False),
                                            (2, 9, 35, 3, True), (64, 8, 12288, 1, False)])
def test_recon_vs_oracle(cuda_device, B, T, F, S, ragged):
    import gpkl
    dev = cuda_device
    g = torch.Generator().manual_seed(B * 100 + T)
    lengths = torch.randint((T + 1) // 2, T + 1, (B,), generator=g, dtype=torch.int32) if ragged else torch.full((B,), T, dtype=torch.int32)
    total = int(lengths.sum())
    x = (torch.rand(total, F, generator=g) < 0.2).float()
    xd = torch.rand(S * total, F, generator=g) * 0.98 + 0.01
    xo = xd.clone().requires_grad_(True)
    ref = orc.bernoulli_recon(x, xo, lengths, S)
    (0.37 * ref).backward()
    xg = xd.to(dev).requires_grad_(True)
    out = gpkl.bernoulli_recon(x.to(dev), xg, lengths.to(dev), S)
    (0.37 * out).backward()
    assert abs(float(out) - float(ref)) < 1e-5 * abs(float(ref))
    assert rel_err(xg.grad, xo.grad) < 1e-5
    # deterministic
    out2 = gpkl.bernoulli_recon(x.to(dev), xg.detach(), lengths.to(dev), S)
    assert float(out2) == float(out)
