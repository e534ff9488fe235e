"""CPU, world_size 2 over gloo: the data-parallel host logic (gpkl.parallel) -- sequence sharding with no
forward collective, one all-reduce of the lengthscale gradients -- reproduces the single-process result.
The oracle stands in for the CUDA op here (tests may use it as the checker/stand-in; no GPU on this box)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ragged, out_dir):
    for p in (ROOT, os.path.join(ROOT, "gp-vae_b200"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import gp_kl_oracle as orc
    from gpkl.parallel import GradBucket, shard_batch, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    B, D, T, S = 7, 3, 9, 2
    full = orc.synthetic_batch(B, D, T, S, ragged=ragged, seed=77)          # same on every rank
    mine = shard_batch(full, rank, world)
    lo, hi = shard_range(B, rank, world)
    assert mine["lengths"].shape[0] == hi - lo
    out, grads = orc.gp_prior_kl_grads(mine["mean"], mine["times"], mine["lengths"], mine["ell_q"], mine["ell_p"],
                                       mine["eps"], mine["g_z"], S=S)
    bucket = GradBucket(D, torch.device("cpu"))
    bucket.g_ell_q.copy_(grads["ell_q"].float())
    bucket.g_ell_p.copy_(grads["ell_p"].float())
    bucket.kl.copy_(out["kl_sum"].float().reshape(1))
    bucket.all_reduce()
    torch.save({"bucket": bucket.flat.clone(), "g_mean": grads["mean"], "z": out["z"], "kl_pairs": out["kl_pairs"],
                "lo": lo, "hi": hi}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.parametrize("ragged", [False, True])
def test_two_ranks_match_single_process(tmp_path, ragged):
    import gp_kl_oracle as orc
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), ragged, str(tmp_path)), nprocs=world, join=True)
    B, D, T, S = 7, 3, 9, 2
    full = orc.synthetic_batch(B, D, T, S, ragged=ragged, seed=77)
    out, grads = orc.gp_prior_kl_grads(full["mean"], full["times"], full["lengths"], full["ell_q"], full["ell_p"],
                                       full["eps"], full["g_z"], S=S)
    r = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % k)) for k in range(world)]
    # every rank holds the same reduced bucket = whole-batch lengthscale gradients and KL
    assert torch.equal(r[0]["bucket"], r[1]["bucket"])
    ref = torch.cat([grads["ell_q"].float(), grads["ell_p"].float(), out["kl_sum"].float().reshape(1)])
    assert torch.allclose(r[0]["bucket"], ref, rtol=2e-6, atol=1e-6)
    # per-sequence outputs concatenate in rank order with no exchange
    assert torch.allclose(torch.cat([x["g_mean"] for x in r]), grads["mean"], rtol=0, atol=0)
    assert torch.allclose(torch.cat([x["z"] for x in r]), out["z"], rtol=0, atol=0)
    assert torch.allclose(torch.cat([x["kl_pairs"] for x in r]), out["kl_pairs"], rtol=0, atol=0)
    assert r[0]["lo"] == 0 and r[0]["hi"] == r[1]["lo"] and r[1]["hi"] == B


def test_shard_range_covers_batch():
    from gpkl.parallel import shard_range
    for B in (0, 1, 5, 8, 13):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
