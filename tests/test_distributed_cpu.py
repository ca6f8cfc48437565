"""N>1 host logic on CPU: two gloo ranks shard a Monte-Carlo ensemble, replay the reference's rand() stream for their own
block and reduce ensemble statistics; the union must equal the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from ravenhydroframework_b200 import api, ensemble

MODEL = os.path.join(H.GOLDEN, "hmets_mc", "model")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m0, nloc = ensemble.shard_members(4, rank, world)
        model = api.RavenModel(MODEL)
        model.flatten(nloc)
        samples = ensemble.sample_shard(model, m0, nloc)
        # a per-member scalar series standing in for the hydrographs: the oracle's outlet flow of each local member
        q_local = np.stack([H.oracle_run(model, 60, member=i, record_state=False)["qout_rec"][:, 0] for i in range(nloc)])
        stats = ensemble.ensemble_statistics(torch.from_numpy(q_local))
        q.put((rank, m0, nloc, samples, {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in stats.items()}))
    finally:
        dist.destroy_process_group()


def test_shard_members_partition():
    for n in (1, 4, 7, 4096):
        for w in (1, 2, 3, 8):
            blocks = [ensemble.shard_members(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and sum(b[1] for b in blocks) == n
            for a, b in zip(blocks, blocks[1:]):
                assert a[0] + a[1] == b[0]
            assert max(b[1] for b in blocks) - min(b[1] for b in blocks) <= 1


def test_two_rank_ensemble_equals_single_process():
    # single process: all four members
    model = api.RavenModel(MODEL)
    model.flatten(4)
    ref_samples = ensemble.sample_shard(model, 0, 4)
    q_all = np.stack([H.oracle_run(model, 60, member=i, record_state=False)["qout_rec"][:, 0] for i in range(4)])
    ref_stats = ensemble.ensemble_statistics(torch.from_numpy(q_all))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 2), (2, 2)]
    assert np.array_equal(np.concatenate([r[3] for r in res]), ref_samples)        # same rand() stream on every rank
    for r in res:
        for k in ("mean", "std", "min", "max"):
            assert np.allclose(r[4][k], ref_stats[k].numpy(), rtol=1e-12, atol=1e-15), k
        assert r[4]["n"] == 4
