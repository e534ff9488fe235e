"""Developer aid: the tile tier (shared prior, 208 < T <= 512) against the per-pair kernels and the float64 oracle,
every output's error printed (no asserts), plus kernel times.   python tools/tile_debug.py [T ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gp-vae_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import gp_kl_oracle as orc  # noqa: E402
from conftest import rel_err  # noqa: E402
from gpu_util import run_cuda, run_oracle  # noqa: E402

dev = torch.device("cuda:0")
Ts = [int(a) for a in sys.argv[1:]] or [256, 209, 300, 384, 512]
for T in Ts:
    for ragged in (False, True):
        for kernel in ("rbf", "cauchy"):
            case = orc.synthetic_batch(2, 3, T, 1, ragged=ragged, seed=6000 + T, grid=True)
            f1, b1 = run_cuda(case, dev, kernel=kernel, tier="auto", grad_ell_p=False, shared_prior=True)
            f0, b0 = run_cuda(case, dev, kernel=kernel, tier="auto", grad_ell_p=False, shared_prior=False)
            out, grads = run_oracle(case, kernel=kernel)
            e = {
                "kl/pp": rel_err(f1["kl_pairs"], f0["kl_pairs"]), "kl/or": rel_err(f1["kl_pairs"], out["kl_pairs"]),
                "z/pp": rel_err(f1["z"], f0["z"]), "z/or": rel_err(f1["z"], out["z"]),
                "ldq/or": rel_err(f1["logdets"][:, 1], out["logdet_q"]),
                "gm/pp": rel_err(b1["g_mean"], b0["g_mean"]), "gm/or": rel_err(b1["g_mean"], grads["mean"]),
                "glq/pp": rel_err(b1["g_ell_q"], b0["g_ell_q"]), "glq/or": rel_err(b1["g_ell_q"], grads["ell_q"]),
                "status": int(f1["status"]),
            }
            print("T=%d ragged=%d %s: %s" % (T, ragged, kernel, {k: ("%.2e" % v if isinstance(v, float) else v) for k, v in e.items()}),
                  flush=True)
