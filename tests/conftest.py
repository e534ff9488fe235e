"""Shared test plumbing.  `-m "not gpu"` runs on the CPU-only dev box; `-m gpu` runs on a B200."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-vae_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    """Fixture made by oracle/gen_golden.py from the reference's own functions."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "US":
            out[k] = str(v)
        elif v.ndim == 0 and k in ("S",):
            out[k] = int(v)
        elif v.ndim == 0 and k in ("noise", "beta", "kl"):
            out[k] = float(v)
        else:
            out[k] = torch.from_numpy(v)
    return out


def rel_err(a, b):
    """max|a-b| / max|b| -- scale-relative error used for vectors whose entries may cross zero."""
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1).cpu()
    if b.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
