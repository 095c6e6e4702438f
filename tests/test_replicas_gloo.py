"""CPU, world_size 2 over gloo: the N > 1 path is "replicas only" — contiguous shards of independent frame
pairs, no data-path collective; torch.distributed carries only the barrier and the max-over-ranks timing."""
import os
import socket

import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "vi-slam_b200"))
    sys.path.insert(0, root)
    import numpy as np
    import torch.distributed as dist
    from vislam_b200 import replicas
    from oracle import vso
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_pairs = 7
    lo, hi = replicas.shard_range(n_pairs, rank, world)
    # every rank matches its own shard of independent pairs with the CPU oracle (stand-in for its GPU)
    rng = np.random.default_rng(123)
    d1 = rng.integers(0, 256, (n_pairs, 40, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (n_pairs, 40, 32), dtype=np.uint8)
    mine = {k: vso.knn2_hamming(d1[k], d2[k])[0][:, 0].tolist() for k in range(lo, hi)}
    fake_ms = 10.0 + 5.0 * rank
    dist.barrier()
    ms = replicas.max_over_ranks(fake_ms, dist)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))       # host-side concatenation of per-pair results only
    if rank == 0:
        q.put((ms, gathered, replicas.aggregate_throughput(100, ms, world)))
    dist.destroy_process_group()


def test_two_replicas_gloo():
    import numpy as np
    import torch.multiprocessing as mp
    from oracle import vso
    from vislam_b200 import replicas
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, gathered, thr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 15.0                                       # slowest rank defines the step time
    assert thr == pytest.approx(2 * 100 / 15e-3)
    assert [(g[0], g[1]) for g in gathered] == [(0, 4), (4, 7)]
    rng = np.random.default_rng(123)
    d1 = rng.integers(0, 256, (7, 40, 32), dtype=np.uint8)
    d2 = rng.integers(0, 256, (7, 40, 32), dtype=np.uint8)
    merged = {}
    for g in gathered:
        merged.update(g[2])
    assert sorted(merged) == list(range(7))
    for k in range(7):                                      # sharded == unsharded, no cross-rank exchange needed
        assert merged[k] == vso.knn2_hamming(d1[k], d2[k])[0][:, 0].tolist()


def test_shard_range_covers_everything():
    from vislam_b200 import replicas
    for n in (0, 1, 7, 8192, 1999):
        for world in (1, 2, 4, 8):
            spans = [replicas.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    with pytest.raises(ValueError):
        replicas.shard_range(10, 2, 2)


def test_gpu_local_cpus_restores_affinity():
    """The NUMA binding used while pinning host buffers never leaves the process bound: affinity is restored on exit, and
    without NVML / a GPU (this container) it is a no-op that says so."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "vi-slam_b200"))
    from vislam_b200 import replicas
    before = os.sched_getaffinity(0)
    with replicas.gpu_local_cpus(0) as g:
        inside = os.sched_getaffinity(0)
        assert inside <= before and len(inside) >= 1
        if not g.bound:
            assert inside == before
    assert os.sched_getaffinity(0) == before
