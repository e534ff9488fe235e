"""Data-parallel plumbing for the GP-prior path: one process per GPU, batch (sequence) sharding, and the
one exchange step the path has -- the sum of the lengthscale gradients over ranks.

Every (sequence, latent-dim) pair is independent in forward (the reference's loop at
src/Models/Full_GP_VAE_dynamic_time.py:222-226 is a pure map; the only reduction is reduce_sum, :228),
so sequences are partitioned across ranks with no data-path collective.  In backward the per-rank
d/d ell_q[D], d/d ell_p[D] (and the scalar KL for logging) are summed with ONE all-reduce; the backward
kernel's epilogue writes them straight into the bucket tensor that NCCL reduces (no copy in between).
"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Contiguous, balanced [lo, hi) of sequences owned by `rank` (first B % world ranks get one extra)."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(case, rank, world):
    """Slice a whole-batch dict (mean, times, lengths, eps, g_z, aux; ell_* replicated) to this rank's sequences."""
    lengths = case["lengths"]
    B = lengths.shape[0]
    lo, hi = shard_range(B, rank, world)
    offs = torch.zeros(B + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(lengths.to(torch.int64).cpu(), 0)
    r0, r1 = int(offs[lo]), int(offs[hi])
    S = case["eps"].shape[2]
    out = dict(case)
    out["lengths"] = lengths[lo:hi].contiguous()
    out["times"] = case["times"][lo:hi].contiguous()
    out["eps"] = case["eps"][lo:hi].contiguous()
    out["mean"] = case["mean"][r0:r1].contiguous()
    if case.get("g_z") is not None:
        out["g_z"] = case["g_z"][S * r0:S * r1].contiguous()
    if case.get("aux") is not None:
        out["aux"] = case["aux"][r0:r1].contiguous()
    return out


class GradBucket:
    """[g_ell_q (D) | g_ell_p (D) | kl_sum (1, as f32)] -- the tensor the backward kernel's reduction writes
    into and the all-reduce sums in place."""

    def __init__(self, D, device):
        self.D = D
        self.flat = torch.zeros(2 * D + 1, dtype=torch.float32, device=device)

    @property
    def g_ell_q(self):
        return self.flat[: self.D]

    @property
    def g_ell_p(self):
        return self.flat[self.D: 2 * self.D]

    @property
    def kl(self):
        return self.flat[2 * self.D:]

    def out_views(self):
        return {"g_ell_q": self.g_ell_q, "g_ell_p": self.g_ell_p}

    def all_reduce(self, group=None, async_op=False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None
