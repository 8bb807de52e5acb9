"""world_size-2 gloo tests (CPU) of the N>1 path: batch sharding + one all-reduce of dW.

The CUDA kernels cannot run here; the oracle stands in for the per-rank compute, which is
exactly what the data-parallel logic has to be agnostic to: dW is a sum over the batch, so
all-reduce(sum) of the shard gradients must equal the gradient of the whole batch, and y / dX
of a shard must equal the corresponding rows of the full result.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import make_weight
from inverse_flow_b200 import parallel
from oracle import oracle


def test_shard_bounds_cover_the_batch_exactly():
    for B in (0, 1, 5, 100, 257):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                       # same data on every rank
        C, H, W, k, groups = 4, 6, 5, 3, 1
        x = rng.standard_normal((B, C, H, W))
        g = rng.standard_normal((B, C, H, W))
        w = make_weight(rng, C, C, k, k, 0.1, np.float64)
        lo, hi = parallel.shard_bounds(B, world, rank)
        y = oracle.inverse(x[lo:hi], w, groups)
        dx, dw = oracle.backward(g[lo:hi], y, w, groups)
        bucket = torch.from_numpy(dw.reshape(-1).copy())
        # what the fused peer all-reduce kernel computes on the GPUs (csrc/ifk_comm.cu): every rank gathers all
        # buckets and adds them in rank order -> bit-identical on every rank, equal to the collective sum
        gathered = [torch.empty_like(bucket) for _ in range(world)]
        dist.all_gather(gathered, bucket)
        ordered = parallel.rank_order_sum(gathered)
        parallel.allreduce_gradients(bucket)
        ranks_agree = [torch.empty_like(ordered) for _ in range(world)]
        dist.all_gather(ranks_agree, ordered)
        same_everywhere = all(torch.equal(ranks_agree[0], r) for r in ranks_agree)
        close_to_collective = torch.allclose(ordered, bucket, rtol=1e-12, atol=1e-12)
        t = parallel.max_over_ranks(float(rank + 1))
        y_full = oracle.inverse(x, w, groups)
        dx_full, dw_full = oracle.backward(g, y_full, w, groups)
        ok = (np.array_equal(y, y_full[lo:hi]) and np.array_equal(dx, dx_full[lo:hi])
              and np.allclose(bucket.numpy(), dw_full.reshape(-1), rtol=1e-12, atol=1e-12)
              and t == float(world) and same_everywhere and close_to_collective)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8])          # ragged and even shards
def test_two_rank_allreduce_of_shard_gradients_equals_full_batch(B):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(0, True), (1, True)]


def test_allreduce_is_a_noop_without_a_process_group():
    t = torch.ones(4)
    assert parallel.allreduce_gradients(t) is t and torch.all(t == 1)
    assert parallel.max_over_ranks(3.0) == 3.0
