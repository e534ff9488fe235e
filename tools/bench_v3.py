"""V3 (bidiagonal-precision posterior) hot tier vs generic tier: forward + backward kernel time on c2-shaped batches.
    python tools/bench_v3.py > profiles/rNN_v3_hot_tier.txt      (on a B200)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gp-vae_b200"), os.path.join(ROOT, "oracle")]
import torch, gpkl, gp_kl_oracle as orc

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


print("V3 forward+backward, median of 20 (CUDA events, 256 MiB L2 flush between iterations)")
print("%-28s %-10s %-12s %-12s %-10s %-14s" % ("shape", "tier", "fwd ms", "bwd ms", "seq/s", "launches/step"))
L = gpkl._lib.lib()
for (B, D, T, S) in ((256, 35, 48, 1), (64, 256, 10, 1), (1024, 64, 64, 1), (1024, 64, 16, 1)):
    c = orc.synthetic_batch(B, D, T, S, ragged=False, seed=T, posterior="bidiag", grid=True)
    d = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    for tier in ("auto", "generic"):
        f = lambda: gpkl.gp_prior_kl_forward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], aux=d["aux"],
                                             posterior="bidiag", S=S, tier=tier)
        g = lambda: gpkl.gp_prior_kl_backward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], d["g_z"],
                                              aux=d["aux"], posterior="bidiag", S=S, tier=tier, grad_ell_p=False)
        n0 = L.gpkl_launch_count() if hasattr(L, "gpkl_launch_count") else 0
        f(); g()
        n1 = L.gpkl_launch_count() if hasattr(L, "gpkl_launch_count") else 0
        tf, tb = timed(f), timed(g)
        print("%-28s %-10s %-12.4f %-12.4f %-10.0f %-14d" % ("B=%d D=%d T=%d S=%d" % (B, D, T, S), tier, tf, tb, B / ((tf + tb) * 1e-3), n1 - n0))
