"""TEST / MEASUREMENT INFRASTRUCTURE ONLY -- time the UNMODIFIED reference's GP-prior path under oracle/tf1_stub.py at the
reference's own configuration (batch_size 5, latent_size 100, max_time 20, one sample:
/root/reference/src/Models/Full_GP_VAE_dynamic_time.py:312-318), forward + backward, on this machine's host cores.

Run in the dev container (needs /root/reference, which does not travel to the GPU box):

    python oracle/time_verbatim.py            # writes profiles/r02_verbatim_reference_cpu.json

bench.py copies the committed record into cpu_baseline.verbatim and times the float64 oracle port on the same
configuration on the GPU box's cores beside it.
"""
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GPKL_REFERENCE", "/root/reference/src/Models")
sys.path.insert(0, HERE)
sys.path.insert(0, REF)
import tf1_stub  # noqa: E402

tf = tf1_stub.install()
import Full_GP_VAE_dynamic_time as dyn  # noqa: E402  (reference, unmodified)
import gp_kl_oracle as orc  # noqa: E402

B, D, T, S = 5, 100, 20, 1


def one_step(case):
    times, lengths = case["times"], case["lengths"]
    mean = case["mean"].clone().requires_grad_(True)
    lq = case["ell_q"].clone().reshape(D, 1).requires_grad_(True)
    old = tf.Variable
    tf.Variable = lambda *a, **k: lq
    try:
        prior_kernel, _ = dyn.prior_kernels(times, lengths, D, B)                       # :332
        approx_kernel, chol_noise, _ = dyn.approx_kernels(times, lengths, D, B, S)      # :335
    finally:
        tf.Variable = old
    z = dyn.gp_vae_sample(mean, chol_noise, lengths, B, S, D)                           # :339
    kl_sum, _ = dyn.calc_gp_kl(mean, lengths, approx_kernel, prior_kernel, B, D)        # :340
    (kl_sum + (case["g_z"].double() * z.double()).sum()).backward()                     # TF autodiff, :361
    return float(kl_sum)


def main():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case = orc.synthetic_batch(B, D, T, S, ragged=False, seed=1234, grid=True)
    one_step(case)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        kl = one_step(case)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    # the float64 port on the same inputs, same machine
    orc.gp_prior_kl_grads(case["mean"], case["times"], case["lengths"], case["ell_q"], case["ell_p"], case["eps"], case["g_z"])
    tp = []
    for _ in range(5):
        t0 = time.perf_counter()
        orc.gp_prior_kl_grads(case["mean"], case["times"], case["lengths"], case["ell_q"], case["ell_p"], case["eps"], case["g_z"])
        tp.append(time.perf_counter() - t0)
    tp.sort()
    rec = {"what": "unmodified Full_GP_VAE_dynamic_time.{prior_kernels, approx_kernels, gp_vae_sample, calc_gp_kl} + autograd "
                   "backward under oracle/tf1_stub.py (torch eager standing in for TF1 graph ops)",
           "config": "B=%d D=%d T=%d S=%d (the reference's own, :312-318)" % (B, D, T, S),
           "where": "dev container (the reference does not travel to the GPU box)", "cores": cores,
           "seconds_per_step_median_of_5": med, "value": B / med, "unit": "sequences/s", "kl_sum": kl,
           "oracle_port_same_inputs_same_machine": {"seconds_per_step_median_of_5": tp[len(tp) // 2],
                                                    "value": B / tp[len(tp) // 2]}}
    out = os.path.join(ROOT, "profiles", "r02_verbatim_reference_cpu.json")
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()
