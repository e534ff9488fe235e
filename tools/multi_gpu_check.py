"""N-rank GPU result == oracle on the whole batch (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Each rank runs the CUDA path on its shard of the sequences; lengthscale gradients and KL are summed with one
NCCL all-reduce (gpkl.parallel.GradBucket); rank 0 compares with the float64 oracle on the full batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gp-vae_b200"), os.path.join(ROOT, "oracle")]
import torch, torch.distributed as dist
import gpkl, gp_kl_oracle as orc
from gpkl.parallel import GradBucket, shard_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, D, T, S = 4 * world + 1, 6, 40, 1
full = orc.synthetic_batch(B, D, T, S, ragged=True, seed=5)
mine = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in shard_batch(full, rank, world).items()}
f = gpkl.gp_prior_kl_forward(mine["mean"], mine["times"], mine["lengths"], mine["ell_q"], mine["ell_p"], mine["eps"])
bucket = GradBucket(D, dev)
g = gpkl.gp_prior_kl_backward(mine["mean"], mine["times"], mine["lengths"], mine["ell_q"], mine["ell_p"], mine["eps"],
                              mine["g_z"], out=bucket.out_views())
bucket.kl.copy_(f["kl_sum"].float().reshape(1))
bucket.all_reduce()
gm = [torch.empty(0)] * world
sizes = [None] * world
dist.all_gather_object(sizes, g["g_mean"].shape[0])
parts = [torch.empty(n, D, device=dev) for n in sizes]
dist.all_gather(parts, g["g_mean"])
if rank == 0:
    out, grads = orc.gp_prior_kl_grads(full["mean"], full["times"], full["lengths"], full["ell_q"], full["ell_p"],
                                       full["eps"], full["g_z"])
    rel = lambda a, b: float((a.double().cpu() - b.double()).abs().max() / b.double().abs().max())
    e = {"g_ell_q": rel(bucket.g_ell_q, grads["ell_q"]), "kl_sum": abs(float(bucket.kl) - float(out["kl_sum"])) / float(out["kl_sum"]),
         "g_mean": rel(torch.cat(parts), grads["mean"])}
    print("multi_gpu_check world=%d B=%d errs=%s" % (world, B, e))
    assert e["g_ell_q"] < 1e-4 and e["kl_sum"] < 2e-5 and e["g_mean"] < 1e-4
    print("OK")
dist.destroy_process_group()
