"""HBM roofline of the reconstruction-term kernels (streaming stage).
Prints GB/s (algorithmic bytes / CUDA-event time) against MEASURED_PEAKS.json's copy bandwidth."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gp-vae_b200")]
import torch, gpkl
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, B, T, F, S in (("sprites-shape 512x8 rows, 64x64x3", 512, 8, 12288, 1), ("moving-mnist shape B=256, T=20, 64x64", 256, 20, 4096, 1),
                         ("same, S=4 samples", 256, 20, 4096, 4)):
    lengths = torch.full((B,), T, dtype=torch.int32, device=dev)
    x = (torch.rand(B * T, F, device=dev) < 0.2).float()
    xd = (torch.rand(S * B * T, F, device=dev) * 0.98 + 0.01).requires_grad_(True)
    import ctypes
    L = gpkl._lib.lib()
    tf, tb = [], []
    for it in range(8):
        L.gpkl_profile_enable(1)   # CUDA events recorded by the library right around the streaming kernel
        flush.zero_()
        out = gpkl.bernoulli_recon(x, xd, lengths, S)
        flush.zero_()
        out.backward()
        torch.cuda.synchronize()
        fm, bm, nf, nb = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_int32(0), ctypes.c_int32(0)
        L.gpkl_profile_read(ctypes.byref(fm), ctypes.byref(nf), ctypes.byref(bm), ctypes.byref(nb))
        L.gpkl_profile_enable(0)
        xd.grad = None
        if it >= 3:
            tf.append(fm.value); tb.append(bm.value)
    n = S * B * T * F
    bf = 4.0 * n + 4.0 * B * T * F          # read x_decode once, x once (re-reads of x across samples hit L2)
    bb = 8.0 * n + 4.0 * B * T * F          # read x_decode, write gradient, read x
    f, b = min(tf), min(tb)
    print("%-42s fwd %.1f us %.0f GB/s (%.0f%% of %.0f)   bwd %.1f us %.0f GB/s (%.0f%%)" % (
        name, f * 1e3, bf / f / 1e6, 100 * bf / f / 1e6 / peak, peak, b * 1e3, bb / b / 1e6, 100 * bb / b / 1e6 / peak))
