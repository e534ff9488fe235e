"""Worst-case CUDA-vs-oracle errors per tier and input kind: the numbers behind the tolerances of
tests/gpu_util.py, printed instead of asserted.    python tools/parity_report.py > profiles/rNN_parity_residuals.txt
Inputs: the SURVEY S8(d) synthetic batches, on the reference's own time grid 0..T-1 (strict 1e-5 / 1e-4 in the tests) and on
irregular times cumsum(U(0.5, 1.5)) (cond(K) ~ 1e3; the tests allow 1e-5 + 4 x the reference's own rounding floor there)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gp-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch
import gp_kl_oracle as orc
from gpu_util import compare

dev = torch.device('cuda:0')
CASES = [  # (tier, T list, B, D)
    ("warp", [8, 10, 20, 32, 48, 64], 6, 5),
    ("block", [65, 96, 128, 144, 160, 208], 3, 4),
    ("auto", [209, 256, 300, 384, 512], 2, 3),   # tile tier
    ("auto", [600, 768], 1, 2),                  # block tier from workspace slots (per-pair prior)
]
print("%-6s %-7s %-5s %-9s | %-9s %-9s %-9s %-9s %-9s | floor: %-9s %-9s" % (
    "tier", "kernel", "T", "times", "kl_pairs", "logdet_q", "z", "g_mean", "g_ell_q", "kl", "grad"))
worst = {}
for tier, Ts, B, D in CASES:
    for kernel in ("rbf", "cauchy"):
        for T in Ts:
            for grid in (True, False):
                case = orc.synthetic_batch(B, D, T, 1, ragged=True, seed=4000 + T, grid=grid)
                e = compare(case, dev, floor=True, kernel=kernel, S=1, tier=tier, grad_ell_p=False)
                kind = "grid" if grid else "irregular"
                print("%-6s %-7s %-5d %-9s | %-9.2e %-9.2e %-9.2e %-9.2e %-9.2e | %-16.2e %-9.2e" % (
                    tier, kernel, T, kind, e["kl_pairs"], e["logdet_q"], e["z"], e["g_mean"], e["g_ell_q"],
                    e["floor"]["kl"], e["floor"]["grad"]), flush=True)
                w = worst.setdefault((tier if T <= 512 else "slots", kind), {"kl": 0.0, "grad": 0.0, "fkl": 0.0, "fgrad": 0.0})
                w["kl"] = max(w["kl"], e["kl_pairs"], e["logdet_q"], e["z"])
                w["grad"] = max(w["grad"], e["g_mean"], e["g_ell_q"])
                w["fkl"] = max(w["fkl"], e["floor"]["kl"]); w["fgrad"] = max(w["fgrad"], e["floor"]["grad"])
print()
print("worst case per tier / input kind (KL, log-det, z | gradients | the reference's own float32-K rounding floor):")
for (tier, kind), w in worst.items():
    print("  %-6s %-9s  %.2e | %.2e | floor %.2e / %.2e   %s" % (
        tier, kind, w["kl"], w["grad"], w["fkl"], w["fgrad"],
        "strict 1e-5 / 1e-4 met" if w["kl"] < 1e-5 and w["grad"] < 1e-4 else "above strict; within 1e-5 + 4 x floor" if
        w["kl"] < 1e-5 + 4 * w["fkl"] and w["grad"] < 1e-4 + 4 * w["fgrad"] else "OUTSIDE"))
