"""world_size-2 gloo test of the only multi-rank logic on the path: chain sharding + the final all-reduce of the
moment sums must reproduce the single-process pooled statistics / ESS."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, means, seconds, out):
    import pdmpflux_b200 as p
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    off, cnt = p.dist.shard(means.shape[0], rank, world)
    sums = p.dist.moment_sums(torch.from_numpy(means[off:off + cnt]), torch.from_numpy(seconds[off:off + cnt]))
    p.dist.all_reduce_sums(sums)
    if rank == 0:
        st = p.dist.ess_from_sums(sums)
        out.put({k: np.asarray(v) for k, v in st.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_shard_covers_all_chains():
    import pdmpflux_b200 as p
    for total, world in ((4096, 1), (4096, 8), (65536, 8), (10, 4), (7, 8)):
        blocks = [p.dist.shard(total, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
        for (o0, c0), (o1, _) in zip(blocks, blocks[1:]):
            assert o0 + c0 == o1


def test_two_rank_moment_reduction_matches_single_process():
    import pdmpflux_b200 as p
    g = np.random.default_rng(0)
    C, d = 101, 6
    means = g.standard_normal((C, d)) * 0.1
    seconds = 1.0 + 0.05 * g.standard_normal((C, d))
    ref = p.dist.ess_from_sums(p.dist.moment_sums(means, seconds))
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, means, seconds, out)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = out.get(timeout=120)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    for k in ("mean", "var", "ess"):
        assert np.allclose(got[k], ref[k], rtol=1e-12, atol=0)
    assert got["chains"] == C
    # definition check: ESS = C * pooled variance / between-chain variance of the chain means
    assert np.allclose(ref["ess"], C * (seconds.mean(0) - means.mean(0) ** 2) / means.var(0, ddof=1))
