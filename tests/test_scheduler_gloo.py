"""Multi-rank host logic of the utterance sharder on CPU: world_size 2, gloo.

The data path has NO collective: each rank plans its own share from the same length list and the host gathers the
audio.  What needs a multi-process test is therefore the plan itself -- every utterance lands on exactly one rank,
the expected frame load balances -- and the only collectives bench.py issues (barrier, MAX of the timings, SUM of
the work counters)."""
import os
import socket

import numpy as np
import pytest

from phoonnx_b200 import scheduler


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(2)                      # every rank derives the same workload (bench.py does too)
        lengths = rs.randint(64, 257, size=(1000,)).astype(np.int64)
        batches = scheduler.plan(lengths, world, rank, max_ids=8192)
        mine = np.concatenate(batches)
        # what bench.py reduces: device time (MAX) and work counters (SUM)
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([float(lengths[mine].sum()), float(len(mine))], dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        dist.barrier()
        # ownership mask, summed over ranks: each utterance exactly once
        own = torch.zeros(1000, dtype=torch.int32)
        own[torch.from_numpy(mine)] += 1
        dist.all_reduce(own, op=dist.ReduceOp.SUM)
        q.put((rank, int(lengths[mine].sum()), [int(b.sum()) for b in [lengths[x] for x in batches]], t.item(), w.tolist(),
               bool((own == 1).all()), int(lengths.sum())))
    finally:
        dist.destroy_process_group()


def test_sharding_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ids0, b0, tmax0, w0, once0, total), (r1, ids1, b1, tmax1, w1, once1, _) = res
    assert once0 and once1                                 # every utterance on exactly one rank
    assert ids0 + ids1 == total and w0 == w1 == [float(total), 1000.0]
    assert tmax0 == tmax1 == 2.0                           # MAX over ranks is what both report
    assert abs(ids0 - ids1) / total < 0.02                 # greedy LPT deal balances the id (≈ frame) load
    assert max(b0 + b1) <= 8192                            # device batches respect the id budget


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_plan_partitions_and_balances(world):
    rs = np.random.RandomState(5)
    lengths = rs.randint(64, 257, size=(4096,)).astype(np.int64)
    seen = np.zeros(4096, np.int32)
    loads = []
    for rank in range(world):
        batches = scheduler.plan(lengths, world, rank)
        idx = np.concatenate(batches) if batches else np.zeros((0,), np.int64)
        seen[idx] += 1
        loads.append(int(lengths[idx].sum()))
        for b in batches:                                   # length-bucketed: similar lengths share a device batch
            assert int(lengths[b].sum()) <= 32768
    assert (seen == 1).all()
    assert (max(loads) - min(loads)) / max(loads) < 0.04


def _gather_worker(rank, world, name, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    rs = np.random.RandomState(2)
    n_utts, hop = 200, 256
    lengths = rs.randint(64, 257, size=(n_utts,)).astype(np.int64)
    frames_of = lambda g: 1 + (int(lengths[g]) * 7 + 3 * g) % 11            # "durations": any deterministic function both ranks agree on
    g = bench.HostGather(name, world, rank, n_utts, 1 << 22, create=False)
    try:
        for step in (1, 2):
            results = []
            for bidx in scheduler.plan(lengths, world, rank, max_ids=4096):
                alen = np.asarray([frames_of(int(i)) * hop for i in bidx], np.int64)
                audio = np.concatenate([np.full((int(n),), float(i) + 0.5 * step, np.float32) for i, n in zip(bidx, alen)])
                results.append((bidx, audio, alen))
            total = g.gather(step, hop, results)
        if rank == 0:
            offs = np.concatenate([[0], np.cumsum([frames_of(i) * hop for i in range(n_utts)])])
            ok = total == int(offs[-1]) and all((g.audio[offs[i]:offs[i + 1]] == float(i) + 1.0).all() for i in range(n_utts))
            q.put(("ok", bool(ok), total))
        else:
            q.put(("done", True, total))
    finally:
        g.close(False)


def test_host_gather_two_ranks_original_order():
    """bench.py --scaling strong ("only the host gathers the audio"): two processes place their shards' audio into
    ONE shared host buffer in original utterance order, hand-shaking through flags in the same segment -- no collective."""
    import multiprocessing as mp
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    name = f"vits_test_gather_{os.getpid()}"
    owner = bench.HostGather(name, 2, 0, 200, 1 << 22, create=True)
    try:
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        ps = [ctx.Process(target=_gather_worker, args=(r, 2, name, q)) for r in range(2)]
        for p in ps:
            p.start()
        got = [q.get(timeout=120) for _ in ps]
        for p in ps:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert sorted(x[0] for x in got) == ["done", "ok"] and all(x[1] for x in got)
        assert got[0][2] == got[1][2] > 0
    finally:
        owner.close(True)
