"""Helpers shared by the GPU parity tests: run the CUDA path (through the C ABI) and the oracle on the
same inputs and report scale-relative errors."""
import torch

import gp_kl_oracle as orc
from conftest import rel_err

TOL_KL = 1e-5      # north_star: log-det and KL within 1e-5 relative (float32 kernels vs float64 reference)
TOL_GRAD = 1e-4    # north_star: gradients within 1e-4
TOL_Z = 1e-5


def to_dev(case, dev):
    return {k: (v.to(dev).contiguous() if isinstance(v, torch.Tensor) else v) for k, v in case.items()}


def run_cuda(case, dev, *, kernel="rbf", posterior="gp", noise=1e-3, S=1, tier="auto", grad_ell_p=True,
             g_kl_pairs=None, g_kl_sum=1.0):
    import gpkl
    c = to_dev(case, dev)
    aux = c.get("aux")
    if aux is None:
        aux = c.get("logvar")
    fwd = gpkl.gp_prior_kl_forward(c["mean"], c["times"], c["lengths"].to(torch.int32), c["ell_q"], c["ell_p"],
                                   c["eps"], aux=aux, kernel=kernel, posterior=posterior, noise=noise, S=S, tier=tier,
                                   want_logdets=True, want_status=True)
    gks = torch.tensor(float(g_kl_sum), dtype=torch.float64, device=dev)
    gkp = None if g_kl_pairs is None else g_kl_pairs.to(dev).float().contiguous()
    bwd = gpkl.gp_prior_kl_backward(c["mean"], c["times"], c["lengths"].to(torch.int32), c["ell_q"], c["ell_p"],
                                    c["eps"], c.get("g_z"), gks, gkp, aux=aux, kernel=kernel, posterior=posterior,
                                    noise=noise, S=S, tier=tier, grad_ell_p=grad_ell_p)
    torch.cuda.synchronize()
    return fwd, bwd


def run_oracle(case, *, kernel="rbf", posterior="gp", noise=1e-3, S=1, g_kl_pairs=None, g_kl_sum=1.0):
    aux = case.get("aux")
    if aux is None:
        aux = case.get("logvar")
    return orc.gp_prior_kl_grads(case["mean"], case["times"], case["lengths"], case["ell_q"], case["ell_p"],
                                 case["eps"], case.get("g_z"), g_kl_sum, g_kl_pairs, aux=aux, kernel=kernel,
                                 posterior=posterior, noise=noise, S=S)


def compare(case, dev, **cfg):
    """Returns dict of scale-relative errors CUDA vs oracle."""
    fwd, bwd = run_cuda(case, dev, **cfg)
    ocfg = {k: v for k, v in cfg.items() if k not in ("tier", "grad_ell_p")}
    out, grads = run_oracle(case, **ocfg)
    errs = {
        "kl_pairs": rel_err(fwd["kl_pairs"], out["kl_pairs"]),
        "kl_sum": abs(float(fwd["kl_sum"]) - float(out["kl_sum"])) / max(abs(float(out["kl_sum"])), 1e-300),
        "logdet_p": rel_err(fwd["logdets"][:, 0], out["logdet_p"]),
        "z": rel_err(fwd["z"], out["z"]),
        "g_mean": rel_err(bwd["g_mean"], grads["mean"]),
        "status": int(fwd["status"]),
    }
    if cfg.get("posterior", "gp") == "gp":
        errs["logdet_q"] = rel_err(fwd["logdets"][:, 1], out["logdet_q"])
        errs["g_ell_q"] = rel_err(bwd["g_ell_q"], grads["ell_q"])
    else:
        errs["g_aux"] = rel_err(bwd["g_aux"], grads["aux"])
    if cfg.get("grad_ell_p", True):
        errs["g_ell_p"] = rel_err(bwd["g_ell_p"], grads["ell_p"])
    return errs


def assert_parity(errs, tag=""):
    bad = {}
    for k, v in errs.items():
        if k == "status":
            if v != 0:
                bad[k] = v
            continue
        tol = TOL_GRAD if k.startswith("g_") else (TOL_Z if k == "z" else TOL_KL)
        if not (v < tol):
            bad[k] = v
    assert not bad, "%s parity failures %s (all: %s)" % (tag, bad, errs)
