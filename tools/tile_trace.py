"""Developer aid: per-phase cycles of CTA 0 in the tile tier (thread 0's clock, accumulated over its pairs).
    python tools/tile_trace.py [T ...]      env TRACE_B (sequences, default 37), TRACE_D (latent dims, default 16)"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'gp-vae_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import torch, gpkl, gp_kl_oracle as orc
dev = torch.device('cuda:0')
L = gpkl._lib.lib()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
NB = int(os.environ.get("TRACE_B", "37"))
ND = int(os.environ.get("TRACE_D", "16"))
FW = ['load+a', 'chol.diagtile', 'chol.gemm||diag', 'chol.rows', 'hook(z+product)', 'final']
BW = ['load', 'chol.diagtile', 'chol.gemm||diag', 'chol.rows', 'hook(w)', 'final-unused', 'inv.diag', 'inv.gemm', 'inv.store+Cprime', 'alpha', 'contraction+final']
for T in [int(a) for a in sys.argv[1:]] or [512]:
    case = orc.synthetic_batch(NB, ND, T, 1, seed=1)
    c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    args = (c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"], c["eps"])
    for name, fn, labels in (("fwd", lambda: gpkl.gp_prior_kl_forward(*args, kernel="cauchy"), FW),
                             ("bwd", lambda: gpkl.gp_prior_kl_backward(*args, c["g_z"], kernel="cauchy"), BW)):
        fn(); torch.cuda.synchronize()
        buf.zero_()
        L.gpkl_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
        fn(); torch.cuda.synchronize()
        L.gpkl_debug_set_trace(None)
        t = buf.cpu().tolist()
        n = max(t[63], 1)
        tot = (sum(t[48:48 + len(labels)]) - (t[57] if name == 'bwd' else 0) + t[59] + t[60] + t[61] + (t[62] if name == 'bwd' else 0)) / n  # (+ the diag-tile phase and the staged inverse, clocked separately; slot 57 is the team's own clock)
        print('T=%d %s: %d pairs by CTA 0, %.0f cycles/pair: ' % (T, name, t[63], tot) +
              ' | '.join('%s %.0f' % (lab, t[48 + i] / n) for i, lab in enumerate(labels)), flush=True)
        print('      diag-tile phase split: prefill %.0f | warp 0 k-loop %.0f | warp 0 flush %.0f' % (t[59] / n, t[60] / n, t[61] / n), flush=True)
        print('      diagonal team (warps 6-7): %.0f cycles per pair in factor + invert + publish (6 panels)' % (t[57] / n), flush=True)
        if name == "bwd":
            print('      inverse: staged products %.0f (the rest of inv.gemm is the L_II^-1 multiply)' % (t[62] / n), flush=True)
        if t[35]:
            print('      staged loops of warp 0 (cycles per chunk of 8 steps, %d chunks): wait %.0f | compute %.0f | issue %.0f' % (
                t[35], t[32] / t[35], t[33] / t[35], t[34] / t[35]), flush=True)
