"""Profiling driver: a few forward+backward launches of a C4-shaped batch (T=512, D=64, Cauchy; B sequences) so that
every SM runs the tile tier with several pairs each.   python tools/prof_tile.py [B] [reps] [T]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gp-vae_b200'), os.path.join(ROOT, 'oracle')]
import torch, gpkl, gp_kl_oracle as orc
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 512
dev = torch.device('cuda:0')
case = orc.synthetic_batch(B, 64, T, 1, seed=1)
c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
args = (c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"], c["eps"])
for _ in range(reps):
    f = gpkl.gp_prior_kl_forward(*args, kernel="cauchy")
    b = gpkl.gp_prior_kl_backward(*args, c["g_z"], kernel="cauchy")
torch.cuda.synchronize()
print("kl_sum", float(f["kl_sum"]), "g_ell_q[0]", float(b["g_ell_q"][0]))
