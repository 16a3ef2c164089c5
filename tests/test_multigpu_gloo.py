"""CPU: the N > 1 host path with a world_size-2 gloo group: shard by global replicate index, all-reduce the flat
summary.  The per-rank batches are produced by the oracle here (no GPU in this container); on the GPU box the same
code all-reduces the plan's device summary over NCCL (bench.py)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from unbiasedmcmc_b200 import _abi as A


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nrep, q):
    import torch.distributed as dist
    from oracle import oracle_py as O
    from tests import cases
    from unbiasedmcmc_b200 import multigpu as MG
    from unbiasedmcmc_b200 import registry as R
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mdl, bins = cases.c1_model(), cases.c1_bins()
    off = MG.shard_offset(0, rank, world, nrep)
    res = O.sample_unbiasedestimator(mdl, R.Draws(seed=42), h_kind=A.UBM_H_BOTH, bins=bins, k=20, m=80, lag=5,
                                     nrep=nrep, rep_offset=off)
    summ = res["summary"].copy()
    MG.allreduce_summary(summ)
    if rank == 0:
        q.put(summ)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_summary_allreduce_matches_single_process():
    from oracle import oracle_py as O
    from tests import cases
    from unbiasedmcmc_b200 import multigpu as MG
    from unbiasedmcmc_b200 import registry as R
    O.lib()
    world, nrep = 2, 300
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nrep, q)) for r in range(world)]
    for p in procs:
        p.start()
    summ = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mdl, bins = cases.c1_model(), cases.c1_bins()
    full = O.sample_unbiasedestimator(mdl, R.Draws(seed=42), h_kind=A.UBM_H_BOTH, bins=bins, k=20, m=80, lag=5,
                                      nrep=world * nrep, rep_offset=0)["summary"]
    assert np.array_equal(summ[:6], full[:6])                  # counts, sum tau, sum tau^2, cost, iterations: integers
    np.testing.assert_allclose(summ, full, rtol=1e-12, atol=1e-12)
    st = MG.summary_stats(summ, 2, bins.nbins)
    assert st["n_done"] == world * nrep and abs(st["bin_mean"].sum() - 1.0) < 0.05


def test_split_helpers():
    from unbiasedmcmc_b200 import multigpu as MG
    parts = MG.split_replicates(10, 4)
    assert parts == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert MG.shard_offset(2, 1, 4, 100) == 900
