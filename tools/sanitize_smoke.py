"""One forward+backward through every tier / path at small batch sizes (run under compute-sanitizer --tool memcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gp-vae_b200"), os.path.join(ROOT, "oracle")]
import torch, gpkl, gp_kl_oracle as orc
dev = torch.device("cuda:0")
cases = [(3, 5, 7, 2, "auto", "gp"), (2, 3, 16, 1, "auto", "gp"), (2, 3, 31, 1, "auto", "gp"), (2, 3, 48, 1, "auto", "gp"),
         (2, 2, 64, 1, "auto", "gp"), (2, 2, 80, 1, "auto", "gp"), (1, 2, 144, 1, "auto", "gp"), (2, 2, 150, 1, "auto", "gp"),
         (1, 2, 160, 1, "auto", "gp"), (1, 2, 272, 1, "auto", "gp"), (1, 1, 512, 1, "auto", "gp"), (1, 1, 530, 1, "auto", "gp"),
         (2, 2, 40, 1, "block", "gp"), (2, 3, 20, 1, "generic", "gp"), (2, 3, 20, 2, "auto", "diag"), (2, 2, 100, 1, "auto", "diag"),
         (1, 2, 200, 1, "auto", "diag"), (2, 3, 12, 1, "auto", "bidiag")]
for B, D, T, S, tier, post in cases:
    c = orc.synthetic_batch(B, D, T, S, ragged=True, seed=T, posterior=post)
    d = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    lp = tier == "generic"
    f = gpkl.gp_prior_kl_forward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], aux=d["aux"], posterior=post, S=S, tier=tier)
    b = gpkl.gp_prior_kl_backward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], d["g_z"], aux=d["aux"],
                                  posterior=post, S=S, tier=tier, grad_ell_p=lp)
    torch.cuda.synchronize()
    print("ok", B, D, T, S, tier, post, float(f["kl_sum"]))
x = (torch.rand(11, 7, device=dev) < 0.3).float()
xd = (torch.rand(22, 7, device=dev) * 0.9 + 0.05).requires_grad_(True)
out = gpkl.bernoulli_recon(x, xd, torch.tensor([4, 2, 5], dtype=torch.int32, device=dev), 2)
out.backward()
torch.cuda.synchronize()
print("ok recon", float(out))
