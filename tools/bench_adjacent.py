"""HBM roofline of the streaming kernels either side of the path: the GP-recognition
epilogue (recog_fwd/bwd_kernel) and the ragged batch producer (collate_scan/gather_kernel).  CUDA events on the launch
stream around the whole C-ABI call (the recog call includes the fused GP sample it rides on, reported separately),
256 MiB L2 flush before every timed call; bytes are algorithmic."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gp-vae_b200")]
import torch, gpkl
dev = torch.device("cuda:0")
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk)).get("hbm_gbs", 6650.0) if os.path.exists(pk) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=8, warm=3):
    best = 1e30
    for it in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if it >= warm:
            best = min(best, a.elapsed_time(b))
    return best


print("# collate: read data[B,F,T_full] + write x[sum_T,F] + times; peak %.0f GB/s" % peak)
for name, N, F, T_full in (("physionet-shape B=4096 F=35 T=48", 4096, 35, 48), ("toy shape B=65536 F=15 T=45", 65536, 15, 45),
                           ("wide B=2048 F=784 T=64", 2048, 784, 64)):
    data = torch.rand(N, F, T_full, device=dev)
    drop = torch.rand(N, 1, T_full, device=dev) < 0.3
    data = torch.where(drop.expand_as(data), torch.full_like(data, -1.0), data).contiguous()
    grid = torch.arange(T_full, dtype=torch.float32, device=dev)
    index = torch.arange(N, dtype=torch.int32, device=dev)
    L = gpkl._lib.lib()
    import ctypes
    from gpkl.ops import _ptr, _stream
    x = torch.empty(N * T_full, F, device=dev); times = torch.empty(N, T_full, device=dev)
    lengths = torch.empty(N, dtype=torch.int32, device=dev); total = torch.zeros((), dtype=torch.int64, device=dev)
    n = L.gpkl_collate_workspace_bytes(N, T_full)
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    ms = timed(lambda: L.gpkl_collate(N, F, T_full, N, T_full, _ptr(data), _ptr(grid), _ptr(index), _ptr(x), _ptr(times),
                                      _ptr(lengths), _ptr(total), _ptr(ws), n, _stream(dev)))
    kept = int(total.item())
    byt = 4.0 * kept * F * 2 + 4.0 * N * T_full * 2     # kept values read + written; mask row read, times written
    print("%-36s %.1f us  %.0f GB/s (%.0f%% of peak; 4 launches)" % (name, ms * 1e3, byt / ms / 1e6, 100 * byt / ms / 1e6 / peak))

print("# recog sampler: whole call = fused GP sample (compute-bound) + streaming epilogue (read mean, logvar, eps, z; write z, kl)")
for name, B, D, T, S in (("reference config B=5 D=100 T=20", 5, 100, 20, 1), ("c2 shape B=256 D=35 T=48", 256, 35, 48, 1),
                         ("T=8 D=256 B=4096", 4096, 256, 8, 1)):
    lengths = torch.full((B,), T, dtype=torch.int32, device=dev)
    times = torch.arange(T, dtype=torch.float32, device=dev).repeat(B, 1).contiguous()
    mean = torch.randn(B * T, D, device=dev); logvar = torch.randn(B * T, D, device=dev) * 0.3 - 0.5
    ell = torch.ones(D, device=dev); eps = torch.randn(B, D, S, T, device=dev)
    f = timed(lambda: gpkl.gp_recog_sample(mean, logvar, times, lengths, ell, eps, S=S))
    g = timed(lambda: gpkl.gp_prior_kl_forward(mean, times, lengths, ell, ell, eps, S=S))
    el = B * T * D
    byt = 4.0 * el * (2 + 3 * S) + 4.0 * B * T
    print("%-36s call %.1f us, of which fused GP forward %.1f us; epilogue+scan+sum ~%.1f us for %.1f MB (%.0f GB/s)" % (
        name, f * 1e3, g * 1e3, (f - g) * 1e3, byt / 1e6, byt / max(f - g, 1e-6) / 1e6))
