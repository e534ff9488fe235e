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
# rows either side of the path: GP-recognition sampler and ragged batch producer
for B, D, T, S in ((3, 5, 9, 2), (2, 3, 100, 1)):
    g = torch.Generator().manual_seed(T)
    lengths = torch.randint((T + 1) // 2, T + 1, (B,), generator=g, dtype=torch.int32)
    total = int(lengths.sum())
    times = torch.arange(T, dtype=torch.float32).repeat(B, 1)
    for b in range(B):
        times[b, int(lengths[b]):] = 0
    mean = torch.randn(total, D, generator=g).to(dev).requires_grad_(True)
    logvar = (0.3 * torch.randn(total, D, generator=g) - 0.5).to(dev).requires_grad_(True)
    ell = torch.ones(D, device=dev, requires_grad=True)
    z, kl_sum, kl_rows = gpkl.gp_recog_sample(mean, logvar, times.to(dev), lengths.to(dev), ell,
                                              torch.randn(B, D, S, T, generator=g).to(dev), S=S)
    (kl_sum + z.double().sum()).backward()
    torch.cuda.synchronize()
    print("ok recog", B, D, T, S, float(kl_sum))
import numpy as np
rng = np.random.RandomState(3)
for N, F, Tf, mt in ((5, 15, 45, 45), (4, 70, 33, 20), (3, 3, 100, 100)):
    data = rng.rand(N, F, Tf).astype(np.float32)
    for i in range(N):
        data[i][:, rng.rand(Tf) < 0.4] = -1.0
    x, times, lengths = gpkl.collate_batch(torch.from_numpy(data).to(dev), torch.arange(Tf, dtype=torch.float32, device=dev),
                                           torch.arange(N, dtype=torch.int32, device=dev), mt)
    torch.cuda.synchronize()
    print("ok collate", N, F, Tf, mt, tuple(x.shape))
