import ctypes, os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gp-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch, gpkl, gp_kl_oracle as orc
dev = torch.device('cuda:0')
L = gpkl._lib.lib()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
for B in (16, 148):
    case = orc.synthetic_batch(B, 2, 512, 1, seed=1)
    c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    args = (c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"], c["eps"])
    gpkl.gp_prior_kl_forward(*args, kernel="cauchy"); torch.cuda.synchronize()
    buf.zero_()
    L.gpkl_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    gpkl.gp_prior_kl_forward(*args, kernel="cauchy"); torch.cuda.synchronize()
    L.gpkl_debug_set_trace(None)
    t = buf.cpu().tolist()
    print("B=%d per panel cycles (CTA 0): load P,W %d | sweeps %d | V %d | update %d | cbar1 %d | writeback %d | cbar2 %d" % tuple([B] + [x // 32 for x in t[:7]]))
