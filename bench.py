#!/usr/bin/env python
"""bench.py -- GP-prior KL hot path (forward + backward) throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2] [--impl ours|reference]

A "step" is one forward+backward pass of the hot path (kernel build, Cholesky, sample, KL, and the
hand-written adjoints) over one batch of synthetic input of the named BASELINE.json config; at N>1 each
rank owns its own batch of that size (weak scaling) and the lengthscale gradients are all-reduced over
NCCL inside the step.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gp-vae_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# The CPU legs (--impl reference, cpu_baseline) use every host thread.  torchrun exports OMP_NUM_THREADS=1 to its ranks; the
# thread count has to be fixed through the environment BEFORE torch is imported: torch.set_num_threads() after the
# import makes this image's oneMKL fail inside the LU inverse at T = 512 ("Parameter 6 was incorrect on entry to DLASWP").
if "--impl" in sys.argv and "reference" in sys.argv and os.environ.get("RANK", "0") == "0":
    os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = str(os.cpu_count() or 1)

import torch  # noqa: E402

# BASELINE.json configs (per-GPU batch).  c3 is the 8-GPU config: 512 sequences sharded 64 per GPU.
WORKLOADS = {
    "c1": dict(T=10, D=256, B=64, kernel="cauchy", desc="HMNIST-shape T=10 D=256 B=64 Cauchy"),
    "c2": dict(T=48, D=35, B=256, kernel="rbf", desc="Physionet-shape T=48 D=35 B=256 RBF"),
    "c3": dict(T=8, D=256, B=64, kernel="rbf", desc="Sprites-shape T=8 D=256 B=64/GPU (512 over 8) RBF"),
    "c4": dict(T=512, D=64, B=1024, kernel="cauchy", desc="long-sequence T=512 D=64 B=1024 Cauchy"),
}
for _t in (16, 32, 64, 96, 128, 160, 192, 256, 384, 512, 768, 1024):
    WORKLOADS["t%d" % _t] = dict(T=_t, D=64, B=max(4, min(1024, (1 << 22) // (_t * _t))), kernel="rbf",
                                 desc="sweep T=%d D=64" % _t)


def model_flops_pair(T, posterior="gp"):
    """V1 forward T^3 (2 Cholesky + 1 triangular solve, T^3/3 each), backward 2 T^3."""
    f = float(T) ** 3
    return (f, 2.0 * f) if posterior == "gp" else (2.0 / 3.0 * f, 4.0 / 3.0 * f)


def executed_flops(T, B, D, shared_prior):
    """Flops the implementation has to execute per launch (forward, backward).  Per-pair prior: the model
    flops (T^3, 2 T^3 per pair).  Shared prior (ell_p one value for all latent dims, the reference's
    prior_time_chars constant, Full_GP_VAE_dynamic_time.py:114): per pair only the K_q side -- forward chol K_q
    (T^3/3) + the product L_p^-1 L_q (T^3/3); backward chol K_q, L_q^-1 and the contraction X_q^T C' (T^3/3 +
    T^3/3 + 2T^3/3) -- plus once per SEQUENCE chol K_p + L_p^-1 (2T^3/3; backward also X^T X, T^3/3 by symmetry)."""
    f = float(T) ** 3
    if not shared_prior:
        return B * D * f, B * D * 2.0 * f
    if tile_forward(T):
        # forward (every shared-prior tier up to T = 512 since round 2: warp, one-buffer block, tile): the trace tr(K_p^-1 K_q) is
        # evaluated entrywise against a float64 K_p^-1 (O(T^2) per pair), so the only FP32 O(T^3) work per pair is chol K_q; the
        # float64 per-sequence inverse (T^3 DP flops per sequence, FP64 pipe) is inside the timed launch but not counted as
        # FP32 work.  Backward: chol K_q, L_q^-1, contraction per pair; the per-sequence float32 K_p^-1 of the block / tile
        # tiers (T^3) -- the register tier (T <= 64) takes its backward record from the float64 sweep too.
        return B * D * (1.0 / 3.0) * f, B * D * (4.0 / 3.0) * f + (B * f if T > 64 else 0.0)
    return B * D * (2.0 / 3.0) * f + B * (2.0 / 3.0) * f, B * D * (4.0 / 3.0) * f + B * f


def tile_forward(T):
    """Forward through the float64 K_p^-1 record (name kept from the round in which only the tile tier had it)."""
    return T <= 512


def algo_bytes_pair(T, S=1):
    """fwd read m, eps, write z (12T) + KL (4); bwd read m, eps, g_z (12T), write g_m (4T): 28T+16 (S=1)."""
    return 28.0 * T + 16.0


def make_case(w, seed, S=1):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_kl_oracle as orc  # input generator only (shared with the tests); no oracle compute here
    return orc.synthetic_batch(w["B"], w["D"], w["T"], S, ragged=False, seed=seed)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.stop_flag = False
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(device_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.h = h
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def sample(self):
        if not self.ok:
            return
        try:
            self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
            try:
                r = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def run(self):
        while not self.stop_flag:
            self.sample()
            time.sleep(0.002)

    def result(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _cpu_sample(orc, w, Bs, Ds, seed=1234):
    """A bounded sample of the workload for the CPU legs: Bs sequences x the first Ds latent dims (every (sequence, dim)
    pair is independent, so seq/s of the full width is the sample's pairs/s / D)."""
    case = orc.synthetic_batch(Bs, w["D"], w["T"], 1, seed=seed)
    if Ds < w["D"]:
        case = dict(case)
        for k in ("mean", "g_z"):
            case[k] = case[k][:, :Ds].contiguous()
        case["eps"] = case["eps"][:, :Ds].contiguous()
        case["ell_q"], case["ell_p"] = case["ell_q"][:Ds].contiguous(), case["ell_p"][:Ds].contiguous()
    return case


def _cpu_call(orc, case, w):
    orc.gp_prior_kl_grads(case["mean"], case["times"], case["lengths"], case["ell_q"], case["ell_p"], case["eps"],
                          case["g_z"], kernel=w["kernel"])


def _cpu_plan(orc, w, per_call_budget_s):
    """Probe with a tiny sample (2 pairs at large T), then size (Bs, Ds) so that one call takes about per_call_budget_s."""
    T, D, B = w["T"], w["D"], w["B"]
    Bs, Ds = (1, min(D, 2)) if T >= 128 else (min(B, 2), D)
    case = _cpu_sample(orc, w, Bs, Ds)
    t0 = time.perf_counter()
    _cpu_call(orc, case, w)
    per_pair = max(time.perf_counter() - t0, 1e-4) / (Bs * Ds)
    pairs = max(1, int(per_call_budget_s / per_pair))
    pairs = min(pairs, max(1, (1 << 28) // (T * T)))  # the float64 oracle keeps ~10 T x T matrices per pair alive
    if pairs >= D:
        Ds, Bs = D, max(1, min(B, pairs // D))
    else:
        Ds, Bs = max(1, pairs), 1
    return Bs, Ds


def cpu_port_throughput(w, budget_s, threads, seed=1234):
    """Oracle (float64 port of the reference's algorithm) forward+backward on a bounded sample of the
    workload, on the host cores.  Returns (seq/s, description of the sample)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_kl_oracle as orc
    T, D = w["T"], w["D"]
    Bs, Ds = _cpu_plan(orc, w, budget_s / 6.0)
    case = _cpu_sample(orc, w, Bs, Ds, seed)
    times = []
    t_start = time.perf_counter()
    for _ in range(5):
        t0 = time.perf_counter()
        _cpu_call(orc, case, w)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    times.sort()
    med = times[len(times) // 2]
    val = Bs * (Ds / float(D)) / med
    return val, "%d sequence(s) x %d of %d latent dims (T=%d) of the %d-sequence batch, %d reps, median" % (
        Bs, Ds, D, T, w["B"], len(times)), Bs, med


def run_reference(args, w, rank, world):
    """--impl reference: the reference's CPU algorithm (float64 oracle port: the reference is TF1 Python and
    cannot travel) on all host threads; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gp_kl_oracle as orc
    cores = torch.get_num_threads()
    T, D = w["T"], w["D"]
    budget = 100.0 / max(1, args.steps + args.warmup)
    Bs, Ds = _cpu_plan(orc, w, budget)
    case = _cpu_sample(orc, w, Bs, Ds)

    def step():
        _cpu_call(orc, case, w)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = Bs * (Ds / float(D)) / dt
    sample = "%d sequence(s) x %d of %d latent dims per step (of %d sequences; pairs are independent)" % (Bs, Ds, D, w["B"])
    print(json.dumps({
        "impl": "reference", "metric": "GP-prior KL fwd+bwd sequences/s", "value": val, "unit": "sequences/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "sample": sample},
        "cpu_baseline": {"value": val, "unit": "sequences/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def verbatim_reference_timing():
    """the UNMODIFIED reference timed under the TF1 stub at its own configuration (B=5, D=100, T=20).
    The reference cannot travel to the GPU box, so the record was measured in the dev container by the committed
    oracle/time_verbatim.py; the float64 port is timed here on the same configuration for scale."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "r02_verbatim_reference_cpu.json")))
    except Exception:
        return None
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gp_kl_oracle as orc
        case = orc.synthetic_batch(5, 100, 20, 1, ragged=False, seed=1234, grid=True)
        a = (case["mean"], case["times"], case["lengths"], case["ell_q"], case["ell_p"], case["eps"], case["g_z"])
        orc.gp_prior_kl_grads(*a)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            orc.gp_prior_kl_grads(*a)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        rec["oracle_port_same_config_this_box"] = {"value": 5 / ts[2], "unit": "sequences/s", "cores": torch.get_num_threads()}
    except Exception:
        pass
    return rec


def measure_device(gpkl, L, w, dev, steps, warmup, cfg, *, world=1, flush=None, grad_ell_p=False, use_graph=False):
    """Device-resident fwd+bwd steps of one workload: CUDA events per step on the launch stream, L2 flushed between timed
    steps (outside the events), async all-reduce of the lengthscale-gradient bucket at N>1 (waited two steps later and
    drained inside the timed region).  use_graph: the step's launches are captured once and replayed (launch-bound sizes).
    Returns per-rank totals; the caller takes the max over ranks."""
    import ctypes
    import torch.distributed as dist
    from gpkl.parallel import GradBucket
    T, D, B = w["T"], w["D"], w["B"]
    rank = int(os.environ.get("RANK", "0"))
    case = make_case(w, 1234 + rank)
    c = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    buckets = [GradBucket(D, dev), GradBucket(D, dev)]
    pending = [None, None]
    one = torch.ones((), dtype=torch.float64, device=dev)
    g_mean = torch.empty_like(c["mean"])
    step_no = [0]

    def compute(bk):
        f = gpkl.gp_prior_kl_forward(c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"], c["eps"], **cfg)
        out = bk.out_views()
        out["g_mean"] = g_mean
        gpkl.gp_prior_kl_backward(c["mean"], c["times"], c["lengths"], c["ell_q"], c["ell_p"], c["eps"], c["g_z"],
                                  one, None, grad_ell_p=grad_ell_p, out=out, **cfg)
        bk.kl.copy_(f["kl_sum"].to(torch.float32).reshape(1))

    graphs = [None, None]
    if use_graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for k in (0, 1):
                compute(buckets[k])  # warms the workspace cache of the capture stream
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph_nodes = 0
        for k in (0, 1):
            g = torch.cuda.CUDAGraph()
            n0 = L.gpkl_launch_count()
            with torch.cuda.graph(g, stream=side):
                compute(buckets[k])
            graph_nodes = L.gpkl_launch_count() - n0  # the library's kernels captured per step (replays do not pass the counter)
            graphs[k] = g

    def step():
        k = step_no[0] & 1
        step_no[0] += 1
        if pending[k] is not None:
            pending[k].wait()  # the reduction that last used this bucket is done (stream-level wait)
            pending[k] = None
        if use_graph:
            graphs[k].replay()
        else:
            compute(buckets[k])
        pending[k] = buckets[k].all_reduce(async_op=True)

    def drain():
        for k in (0, 1):
            if pending[k] is not None:
                pending[k].wait()
                pending[k] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
        if flush is not None:
            flush.zero_()
    drain()
    L.gpkl_profile_enable(0 if use_graph else 1)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = L.gpkl_launch_count()
    sampler.sample()
    sampler.start()
    evs = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
        if flush is not None:
            flush.zero_()  # evict the step's data from L2 between timed iterations (outside the events)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    drain()  # the last all-reduces finish inside the timed region
    e1.record()
    evs.append((e0, e1))
    barrier()
    sampler.stop_flag = True
    sampler.sample()
    per_step_launches = None
    lib_launches = L.gpkl_launch_count() - launches0
    if use_graph:
        lib_launches = graph_nodes * steps
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    fwd_ms, bwd_ms = ctypes.c_double(0), ctypes.c_double(0)
    nf, nb = ctypes.c_int32(0), ctypes.c_int32(0)
    L.gpkl_profile_read(ctypes.byref(fwd_ms), ctypes.byref(nf), ctypes.byref(bwd_ms), ctypes.byref(nb))
    L.gpkl_profile_enable(0)
    return {"total_ms": total_ms, "fwd_ms": fwd_ms.value / max(nf.value, 1), "bwd_ms": bwd_ms.value / max(nb.value, 1),
            "lib_launches": int(lib_launches), "clocks": sampler.result(), "case": case, "dev_case": c}


def roofline_block(w, B, fwd_ms, bwd_ms, peak, shared_prior, step_ms, hbm_peak, hbm_src, traffic):
    T, D = w["T"], w["D"]
    npairs = B * D
    f_fwd, f_bwd = model_flops_pair(T)
    x_fwd, x_bwd = executed_flops(T, B, D, shared_prior)
    ach_bwd = x_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0
    ach_fwd = x_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else 0.0
    mod_bwd = npairs * f_bwd / (bwd_ms * 1e-3) / 1e12 if bwd_ms > 0 else 0.0
    mod_fwd = npairs * f_fwd / (fwd_ms * 1e-3) / 1e12 if fwd_ms > 0 else 0.0
    return {
        "bound": "fp32", "kernel": "backward (Cholesky / inverse / contraction adjoints)", "achieved": ach_bwd, "peak": peak,
        "unit": "TFLOP/s", "frac": ach_bwd / peak if peak > 0 else None, "traffic": traffic,
        "peak_source": "FFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP32 entry; "
                       "nominal 148x128x2x1.965GHz = 74.4)",
        "algorithmic": ("shared prior: %d pairs x 4/3 T^3 + %d sequences x T^3 flops" % (npairs, B if T > 64 else 0)) if shared_prior
                       else "%d pairs x 2*T^3 flops" % npairs,
        "model_frac": mod_bwd / peak if peak > 0 else None, "shared_prior": bool(shared_prior),
        "launch_ms": bwd_ms,
        "fwd_bwd_frac": (x_fwd + x_bwd) / ((fwd_ms + bwd_ms) * 1e-3) / 1e12 / peak if peak > 0 and fwd_ms + bwd_ms > 0 else None,
        "hbm_frac": npairs * algo_bytes_pair(T) / ((fwd_ms + bwd_ms) * 1e-3) / 1e9 / hbm_peak if fwd_ms + bwd_ms > 0 else None,
        "hbm_peak_source": hbm_src,
        "forward": {"achieved": ach_fwd, "frac": ach_fwd / peak if peak > 0 else None,
                    "model_frac": mod_fwd / peak if peak > 0 else None, "launch_ms": fwd_ms,
                    "algorithmic": (("shared prior, float64 K_p^-1 record: %d pairs x 1/3 T^3 FP32 flops (+ %d sequences x T^3 "
                                     "FP64 flops, not counted)" % (npairs, B)) if tile_forward(T) else
                                    ("shared prior: %d pairs x 2/3 T^3 + %d sequences x 2/3 T^3 flops" % (npairs, B)))
                                   if shared_prior else "%d pairs x T^3 flops" % npairs},
        "kernel_share_of_step": (fwd_ms + bwd_ms) / step_ms if step_ms > 0 else None,
    }


def elbo_step_bench(gpkl, w, dev, world, steps=3, warmup=2):
    """The metric BASELINE.json names: one full ELBO forward+backward+Adam step.  Stock torch.nn encoder / decoder (out of
    scope for kernels, SURVEY.md S2: per-time-step MLPs 35 -> 128 -> D and D -> 128 -> 35, the shape of the reference's
    dense nets, Full_GP_VAE_dynamic_time.py:27-58, :262-292) around the fused GP-prior op and the reconstruction / loss op
    (:349-360), Adam(2e-4) (:361); at N>1 every parameter gradient plus the lengthscale gradients are summed with one
    NCCL all-reduce per step.  Returns seq/s (whole job) and the share of the step spent in the GP-prior kernels."""
    import ctypes
    import torch.distributed as dist
    L = gpkl._lib.lib()
    T, D, B = w["T"], w["D"], w["B"]
    F = 35
    rank = int(os.environ.get("RANK", "0"))
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    case = make_case(w, 1234 + rank)
    times, lengths = case["times"].to(dev), case["lengths"].to(dev)
    total_T = int(case["lengths"].sum())
    x = (torch.rand(total_T, F, generator=g) < 0.2).to(torch.float32).to(dev)
    torch.manual_seed(7)
    enc = torch.nn.Sequential(torch.nn.Linear(F, 128), torch.nn.ReLU(), torch.nn.Linear(128, D)).to(dev)
    dec = torch.nn.Sequential(torch.nn.Linear(D, 128), torch.nn.ReLU(), torch.nn.Linear(128, F), torch.nn.Sigmoid()).to(dev)
    ell_q = torch.nn.Parameter(case["ell_q"].to(dev))
    ell_p = case["ell_p"].to(dev)
    params = list(enc.parameters()) + list(dec.parameters()) + [ell_q]
    opt = torch.optim.Adam(params, lr=2e-4)
    step_no = [0]

    def one():
        opt.zero_grad(set_to_none=True)
        mean = enc(x)
        step_no[0] += 1  # production mode: the noise is drawn inside the kernels from a per-step seed (no eps tensor)
        z, kl_sum, _ = gpkl.gp_prior_kl(mean, times, lengths, ell_q, ell_p, None, kernel=w["kernel"], seed=1000 + step_no[0])
        xd = dec(z)
        loss = gpkl.elbo_loss(x, xd, lengths, kl_sum, beta=1.0)
        loss.backward()
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
            off = 0
            for p in params:
                n = p.numel()
                p.grad.copy_(flat[off:off + n].view_as(p))
                off += n
        opt.step()
        return loss

    for _ in range(warmup):
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    L.gpkl_profile_enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = one()
    e1.record()
    torch.cuda.synchronize()
    fwd_ms, bwd_ms = ctypes.c_double(0), ctypes.c_double(0)
    nf, nb = ctypes.c_int32(0), ctypes.c_int32(0)
    L.gpkl_profile_read(ctypes.byref(fwd_ms), ctypes.byref(nf), ctypes.byref(bwd_ms), ctypes.byref(nb))
    L.gpkl_profile_enable(0)
    ms = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    kl_ms = (fwd_ms.value + bwd_ms.value) / steps
    return {"value": B * world / (ms * 1e-3), "unit": "sequences/s", "ms_per_step": ms, "steps": steps,
            "gp_prior_kernels_ms": kl_ms, "gp_prior_share_of_step": kl_ms / ms if ms > 0 else None,
            "loss": float(loss), "model": "per-time-step MLP %d-128-%d encoder, %d-128-%d decoder (stock torch.nn), "
                                          "Bernoulli recon + KL, Adam(2e-4)" % (F, D, D, F)}


_T0 = time.perf_counter()


def note(msg):
    """Progress line on stderr (the JSON line on stdout stays the only stdout output)."""
    sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg))
    sys.stderr.flush()


def v3_bench(gpkl, dev, flush):
    """Forward + backward of the V3 posterior on the c2 shape (T=48, D=35, B=256), median of 20, CUDA events, L2 flushed."""
    import gp_kl_oracle as orc
    ws = WORKLOADS["c2"]
    c = orc.synthetic_batch(ws["B"], ws["D"], ws["T"], 1, ragged=False, seed=77, posterior="bidiag", grid=True)
    d = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    row = {"workload": "V3 bidiagonal-precision posterior on the c2 shape (T=%d D=%d B=%d)" % (ws["T"], ws["D"], ws["B"])}
    for tier in ("auto", "generic"):
        def f():
            gpkl.gp_prior_kl_forward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], aux=d["aux"],
                                     posterior="bidiag", S=1, tier=tier)

        def g():
            gpkl.gp_prior_kl_backward(d["mean"], d["times"], d["lengths"], d["ell_q"], d["ell_p"], d["eps"], d["g_z"],
                                      aux=d["aux"], posterior="bidiag", S=1, tier=tier, grad_ell_p=False)
        ms = []
        for fn in (f, g):
            for _ in range(3):
                fn()
            ts = []
            for _ in range(20):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record()
                b.synchronize()
                ts.append(a.elapsed_time(b))
            ts.sort()
            ms.append(ts[10])
        row["hot_tier" if tier == "auto" else "generic_tier"] = {"fwd_ms": ms[0], "bwd_ms": ms[1], "value": ws["B"] / ((ms[0] + ms[1]) * 1e-3),
                                                                 "unit": "sequences/s"}
    return row


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS),
                    help="headline workload; c4 (T=512, D=64, B=1024, Cauchy) is the config BASELINE.json quotes the "
                         "'1/2/4/8 B200; Cholesky %% of FP32 peak' metric on and it fits one GPU")
    ap.add_argument("--tier", default="auto")
    ap.add_argument("--grad-ell-p", action="store_true", help="also produce d/d ell_p (fixed-T model)")
    ap.add_argument("--per-pair-prior", action="store_true",
                    help="A/B: force the per-pair prior factorisation (GPKL_FLAG_PER_PAIR_PRIOR) instead of the "
                         "shared-prior fast path")
    ap.add_argument("--graph", action="store_true",
                    help="replay the step's forward+backward from a CUDA graph (launch-bound workloads, e.g. c3 over 8 GPUs; "
                         "the all-reduce stays an eager NCCL call behind the replay; no per-kernel times in this mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the short T-sweep of kernel FP32 fractions")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short c2 / c1 / c3 lines (N=1 only)")
    ap.add_argument("--no-elbo", action="store_true", help="skip the full ELBO step (stock torch encoder/decoder + Adam)")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch.distributed as dist
    import gpkl
    import ctypes
    L = gpkl._lib.lib()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)

    T, D, B = w["T"], w["D"], w["B"]
    cfg = dict(kernel=w["kernel"], posterior="gp", noise=1e-3, S=1, tier=args.tier, shared_prior=not args.per_pair_prior)
    # the synthetic inputs carry the reference's prior (ell_p = 1 for every latent dim), so the library takes its
    # shared-prior path unless told otherwise; T > 512 (slot tier), d/d ell_p and the generic tier factor per pair
    shared_prior = (not args.per_pair_prior) and (not args.grad_ell_p) and T <= 512 and args.tier != "generic"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # 256 MiB > 126 MB L2

    # ---- FP32 CUDA-core peak (FFMA microbenchmark, timed alone) --------------------------------------
    sink = torch.empty(148 * 8 * 256, dtype=torch.float32, device=dev)
    flops = ctypes.c_double(0.0)
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    best_peak = 0.0
    for i in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.gpkl_fp32_peak_launch(ctypes.c_void_p(sink.data_ptr()), 20000, ctypes.byref(flops), st)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            best_peak = max(best_peak, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)

    note("FP32 peak %.1f TFLOP/s; timing %s" % (best_peak, args.workload))
    # ---- device-resident timed region ----------------------------------------------------------------
    m = measure_device(gpkl, L, w, dev, args.steps, warmup, cfg, world=world, flush=flush, grad_ell_p=args.grad_ell_p,
                       use_graph=args.graph)
    case = m["case"]
    host = {k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in case.items()}
    # our launches: the library's own count + per step the KL copy into the bucket (+ the NCCL all-reduce at N>1)
    launches = m["lib_launches"] + args.steps + (args.steps if world > 1 else 0)
    tmax = torch.tensor([m["total_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax) / args.steps
    value = B * world / (ms_per_step * 1e-3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    note("device-resident: %.2f ms/step" % ms_per_step)
    # ---- end-to-end: host (pinned) buffers in, EVERY result out (z, KL per pair, d/d mean, d/d ell), copies timed -----
    total_T = case["mean"].shape[0]
    hs = gpkl.HostStep(B, D, T, 1, total_T, kernel=w["kernel"], grad_ell_p=args.grad_ell_p, tier=args.tier,
                       device=dev, shared_prior=not args.per_pair_prior)
    from gpkl.parallel import GradBucket
    bucket = GradBucket(D, dev)
    res_host = torch.empty(2 * D + 1, dtype=torch.float32).pin_memory()

    def e2e_step():
        hs(host["mean"], host["times"], host["lengths"], host["ell_q"], host["ell_p"], host["eps"], host["g_z"],
           full_outputs=True)
        extra = 0
        if world > 1:  # the ranks' lengthscale gradients and KL are summed before the caller reads them
            torch.cuda.current_stream(dev).synchronize()
            bucket.g_ell_q.copy_(hs.g_ell_q, non_blocking=True)
            bucket.kl.copy_(hs.kl_sum.to(torch.float32).reshape(1), non_blocking=True)
            bucket.all_reduce()
            res_host.copy_(bucket.flat, non_blocking=True)
            extra = res_host.numel() * 4
        return hs.h2d_bytes, hs.d2h_bytes + extra
    for _ in range(2):
        h2d_b, d2h_b = e2e_step()
        torch.cuda.synchronize()
    barrier()
    e2e_ms = 0.0
    n_e2e = max(3, min(args.steps, 10 if ms_per_step > 50.0 else 200))
    for _ in range(n_e2e):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h2d_b, d2h_b = e2e_step()
        e1.record()
        e1.synchronize()   # the caller reads the step's results every step
        e2e_ms += e0.elapsed_time(e1)
    barrier()
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_val = B * world / (float(t2) / n_e2e * 1e-3)
    del hs

    note("end to end: %.2f ms/step" % (float(t2) / n_e2e))
    # ---- roofline of the dominant kernel (backward) -------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tr.get(args.workload, {}).get("bwd_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = roofline_block(w, B, m["fwd_ms"], m["bwd_ms"], best_peak, shared_prior, m["total_ms"] / args.steps,
                              hbm_peak, hbm_src, traffic)

    out = {
        "metric": "GP-prior KL fwd+bwd sequences/s", "value": value, "unit": "sequences/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload + ": " + w["desc"], "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": "dp%d (sequences sharded; async all-reduce of lengthscale grads overlapped with the next step)" % world,
                   "l2": "256 MiB flush between timed iterations", "tier": args.tier, "cuda_graph": bool(args.graph),
                   "grad_ell_p": bool(args.grad_ell_p)},
        "clocks": m["clocks"],
        "e2e": {"value": e2e_val, "unit": "sequences/s", "h2d_bytes_per_step": int(h2d_b),
                "d2h_bytes_per_step": int(d2h_b), "ms_per_step": float(t2) / n_e2e, "steps": n_e2e,
                "outputs_copied_back": "z, kl_pairs, kl_sum, g_mean, g_ell_q"},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    # ---- the BASELINE metric itself: full ELBO step (encoder + GP-prior op + decoder + recon + Adam) ----------------------
    if not args.no_elbo:
        try:
            out["elbo_step"] = elbo_step_bench(gpkl, w, dev, world)
        except Exception as ex:  # never lose the headline line to the auxiliary measurement
            out["elbo_step"] = {"error": repr(ex)[:200]}
        note("elbo step done")
    # ---- secondary workloads (N=1): the other single-GPU BASELINE configs, eager and CUDA-graph replay ---------------
    if world == 1 and not args.no_secondary:
        sec = []
        for name in ("c2", "c1", "c3"):
            if name == args.workload:
                continue
            ws = WORKLOADS[name]
            scfg = dict(kernel=ws["kernel"], posterior="gp", noise=1e-3, S=1, tier=args.tier, shared_prior=not args.per_pair_prior)
            row = {"workload": name + ": " + ws["desc"]}
            for mode in ("eager", "graph"):
                r = measure_device(gpkl, L, ws, dev, 50, 5, scfg, world=1, flush=flush, use_graph=(mode == "graph"))
                msps = r["total_ms"] / 50
                row[mode] = {"value": ws["B"] / (msps * 1e-3), "ms_per_step": msps}
                if mode == "eager":
                    rf = roofline_block(ws, ws["B"], r["fwd_ms"], r["bwd_ms"], best_peak, True, msps, hbm_peak, hbm_src, None)
                    row["roofline"] = {k: rf[k] for k in ("frac", "model_frac", "hbm_frac", "kernel_share_of_step", "launch_ms")}
                    row["roofline"]["forward_frac"] = rf["forward"]["frac"]
                    row["roofline"]["forward_launch_ms"] = rf["forward"]["launch_ms"]
            # the same kernels under graph replay (kernel times from the eager pass: replays carry no per-kernel events)
            row["graph"]["kernel_share_of_step"] = row["roofline"]["kernel_share_of_step"] * row["eager"]["ms_per_step"] / row["graph"]["ms_per_step"]
            sec.append(row)
            note("secondary %s done" % name)
        out["secondary"] = sec
    # ---- V3 (bidiagonal-precision posterior, north_star (c)) on the c2 shape: hot tier (gpkl_bidiag.cu) next to the generic tier ----
    if world == 1 and not args.no_secondary:
        try:
            out["v3"] = v3_bench(gpkl, dev, flush)
        except Exception as ex:
            out["v3"] = {"error": repr(ex)[:200]}
        note("v3 done")
    # ---- short T-sweep (N=1): FP32 fraction of the forward / backward kernels at T >= 128 --------------------------------
    if world == 1 and not args.no_sweep:
        sweep = []
        one = torch.ones((), dtype=torch.float64, device=dev)
        fwd_ms, bwd_ms = ctypes.c_double(0), ctypes.c_double(0)
        nf, nb = ctypes.c_int32(0), ctypes.c_int32(0)
        for Ts, Bs in ((128, 128), (256, 64), (384, 32), (512, 16)):
            ws = dict(T=Ts, D=64, B=Bs, kernel="rbf")
            cs = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in make_case(ws, 4321).items()}
            scfg = dict(kernel="rbf", posterior="gp", noise=1e-3, S=1, tier=args.tier, shared_prior=not args.per_pair_prior)

            def sstep():
                gpkl.gp_prior_kl_forward(cs["mean"], cs["times"], cs["lengths"], cs["ell_q"], cs["ell_p"], cs["eps"], **scfg)
                gpkl.gp_prior_kl_backward(cs["mean"], cs["times"], cs["lengths"], cs["ell_q"], cs["ell_p"], cs["eps"],
                                          cs["g_z"], one, None, **scfg)
            for _ in range(3):
                sstep()
            torch.cuda.synchronize()
            L.gpkl_profile_enable(1)
            for _ in range(3):
                sstep()
                flush.zero_()
            torch.cuda.synchronize()
            L.gpkl_profile_read(ctypes.byref(fwd_ms), ctypes.byref(nf), ctypes.byref(bwd_ms), ctypes.byref(nb))
            L.gpkl_profile_enable(0)
            ff, fb = model_flops_pair(Ts)
            xf, xb = executed_flops(Ts, Bs, 64, not args.per_pair_prior)
            pairs = Bs * 64
            fm, bm = fwd_ms.value / max(nf.value, 1), bwd_ms.value / max(nb.value, 1)
            sweep.append({"T": Ts, "pairs": pairs, "fwd_ms": fm, "bwd_ms": bm,
                          "fwd_frac": xf / (fm * 1e-3) / 1e12 / best_peak,
                          "bwd_frac": xb / (bm * 1e-3) / 1e12 / best_peak,
                          "fwd_model_frac": pairs * ff / (fm * 1e-3) / 1e12 / best_peak,
                          "bwd_model_frac": pairs * fb / (bm * 1e-3) / 1e12 / best_peak})
        out["roofline"]["sweep"] = sweep
        note("T sweep done")
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, sample, _, _ = cpu_port_throughput(w, args.cpu_budget, torch.get_num_threads())
        out["cpu_baseline"] = {"value": val, "unit": "sequences/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": sample}
        note("cpu baseline done")
        vb = verbatim_reference_timing()
        if vb is not None:
            out["cpu_baseline"]["verbatim"] = vb
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
