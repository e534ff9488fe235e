"""Per-phase cycle breakdown of CTA 0's FIRST pair in the block tier (internal debug hook)."""
import ctypes, sys
sys.path[:0] = ['/root/repo', '/root/repo/gp-vae_b200', '/root/repo/oracle', '/root/repo/tests']
import torch, gpkl, gp_kl_oracle as orc
from gpu_util import run_cuda
dev = torch.device('cuda:0')
L = gpkl._lib.lib()
buf = torch.zeros(64, dtype=torch.int64, device=dev)
import os
NB = int(os.environ.get("TRACE_B", "37"))   # 37 x 4 = 148 pairs: one per SM; more -> the trace shows CTA 0's LAST pair (steady state)
ND = int(os.environ.get("TRACE_D", "4"))
for T in [int(a) for a in sys.argv[1:]] or [48, 128]:
    case = orc.synthetic_batch(NB, ND, T, 1, seed=1)  # 148 pairs: a full grid (single-CTA launches fetch-throttle)
    buf.zero_()
    L.gpkl_debug_set_trace(ctypes.c_void_p(buf.data_ptr()))
    run_cuda(case, dev, tier='block', grad_ell_p=False)
    L.gpkl_debug_set_trace(None)
    t = buf.cpu().tolist()
    f = [t[i + 1] - t[i] for i in range(0, 6)]
    b = [t[i + 1] - t[i] for i in range(16, 27)]
    print('T=%d fwd cycles: load %d chol_p %d chol_q %d z %d solve %d reduce %d | total %d' % tuple([T] + f + [t[6] - t[0]]))
    if t[7] and t[8]:
        print('T=%d fwd shared-prior path: load %d chol_q %d z %d wait+diag %d a %d product %d reduce %d' % (
            T, t[1] - t[0], t[3] - t[1], t[4] - t[3], t[7] - t[4], t[8] - t[7], t[5] - t[8], t[6] - t[5]))
    print('T=%d bwd cycles: load %d chol_p %d inv_p %d alpha %d t1 %d chol_q %d w %d inv_q %d Cprime %d t2 %d red %d | total %d'
          % tuple([T] + b + [t[27] - t[16]]))
    if t[28] and t[29]:
        print('T=%d bwd C-prime phase: join barrier %d, thread 0 loop %d, end barrier %d' % (T, t[28] - t[24], t[29] - t[28], t[25] - t[29]))
    if any(t[32:48]):
        names = ['chol.tiles', 'chol.sync', 'chol.diag', 'chol.rows', 'solve.tiles', 'solve.sync', 'solve.diag', 'diag.loop']
        print('   panel loops (GPKL_PANEL_TRACE build): fwd ' + ' '.join('%s %d' % (n, v) for n, v in zip(names, t[32:40])))
        print('                                         bwd ' + ' '.join('%s %d' % (n, v) for n, v in zip(names, t[40:48])))
