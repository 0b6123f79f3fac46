"""world_size-2 gloo test of the multi-rank host logic on CPU: shards cover the grid exactly once and the packed
all-reduce of per-shard partial results equals the unsharded result.  The per-shard stand-in numbers come from
the CPU oracle (allowed in tests/); on the GPU box the same code path runs over NCCL (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest

from co_monitor_b200.distributed import PackedResults, shard_range


def test_shards_partition_the_grid():
    for n in (0, 1, 7, 270000, 981_000_001):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, lo, hi, out_dir):
    import torch.distributed as dist

    from oracle import stage1 as o1

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(hi - lo, rank, world)
    w, n, _ = o1.run_range("C", lo + a, lo + b, processes=1, chunk=500)
    # Reff-bin style histogram of the weights (bin = voxel index mod 16 stands in for the Reff bin)
    hist = np.bincount((np.arange(lo + a, lo + b) % 16), weights=w, minlength=16)
    res = PackedResults()
    res.add("weight_sum", np.array([w.sum()]))
    res.add("hist", hist)
    res.count("passing_rays", int(n.sum()))
    res.count("visible_voxels", int((n > 0).sum()))
    res.count("voxels", b - a)
    res.all_reduce()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), weight_sum=res.floats["weight_sum"], hist=res.floats["hist"],
             counters=np.array([res.counters[k] for k in ("passing_rays", "visible_voxels", "voxels")]),
             shard_w=w)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_reduce_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp

    from oracle import stage1 as o1

    lo, hi = 131000, 131600
    port = _free_port()
    mp.spawn(_worker, args=(2, port, lo, hi, str(tmp_path)), nprocs=2, join=True)
    w, n, _ = o1.run_range("C", lo, hi, processes=1, chunk=500)
    hist = np.bincount((np.arange(lo, hi) % 16), weights=w, minlength=16)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in (0, 1))
    for r in (r0, r1):  # every rank holds the reduced result
        np.testing.assert_allclose(r["weight_sum"], [w.sum()], rtol=1e-13)
        np.testing.assert_allclose(r["hist"], hist, rtol=1e-13, atol=1e-300)
        assert list(r["counters"]) == [int(n.sum()), int((n > 0).sum()), hi - lo]
    assert np.array_equal(np.concatenate((r0["shard_w"], r1["shard_w"])), w)  # per-voxel data stays sharded
    assert n.sum() > 100
