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


def _world_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from co_monitor_b200.volume import exchange_world

    assert exchange_world("nccl") == 1 and exchange_world("auto") == 1  # no process group yet
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    got = [exchange_world(m) for m in ("nccl", "peer", "auto", "none")]
    with open(os.path.join(out_dir, f"world{rank}.txt"), "w") as fh:
        fh.write(" ".join(map(str, got)))
    dist.barrier()
    dist.destroy_process_group()


def test_rank_local_pipelines_exchange_nothing(tmp_path):
    """bench.py's rank-0-only sub-records build whole-lattice pipelines inside a multi-rank job: with exchange="none" they
    see a world of one and issue no collective (a one-sided all-reduce there once cost ten minutes of box time)."""
    import torch.multiprocessing as mp

    from co_monitor_b200.volume import exchange_world

    port = _free_port()
    mp.start_processes(_world_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    for r in range(2):
        assert open(tmp_path / f"world{r}.txt").read().split() == ["2", "2", "2", "1"]
    with pytest.raises(ValueError):
        exchange_world("ring")
    # and bench.py asks for exactly that in every sub-record
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    body = src[src.index("def subrecords("):src.index("def explicit_voxels_record(")]
    assert body.count("VolumePipeline(") == body.count('exchange="none"') >= 3


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


# ---- the production exchange: block-cyclic shards, one all-reduce of the Reff-millimetre histogram, Stage 2 from it ----
SPACING, H, L = 20, 12, 10  # 67 x 40 x 12 = 32 160 voxels x 120 crystal points: seconds for the CPU oracle
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]
WINDOW = (14000, 20000)  # local voxel range per rank that holds visible voxels


def _shard_weights(rank, world):
    """Oracle Stage 1 on a window of this rank's block-cyclic shard: global indices, weights."""
    import contextlib
    import io

    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from oracle import geometry_setup as gs
    from oracle import stage1 as o1

    with contextlib.redirect_stdout(io.StringIO()):
        full = CuboidMesh(SPACING).lattice
    lat = full.shard(world, rank, cols_per_block=8)
    lo, hi = (WINDOW[0] // world, WINDOW[1] // world)
    glob = lat.local_to_global(np.arange(lo, min(hi, lat.n_points)))
    scene = o1.build_scene("C", 10, H, L)
    res = o1.run_chunk(scene, gs.voxel_mesh(gs.load_db(), SPACING, flat_indices=glob))
    return full, glob, res.weight


def _mm_histogram(glob, w, reff_all, bins=4096):
    mm = np.rint(reff_all[glob] * 1000.0)
    keep = np.isfinite(mm) & (w != 0)
    return np.bincount(mm[keep].astype(np.int64), weights=w[keep], minlength=bins)[:bins]


def _flux_from_histogram(hist):
    """Stage 2 from the millimetre histogram with the oracle's own per-voxel arithmetic: one pseudo-voxel per
    non-empty bin, Reff = bin / 1000, weight = H[bin] (what com_histogram_rows + com_emissivity_sweep compute)."""
    from oracle import stage2 as o2

    bins = np.flatnonzero(hist)
    observed = np.column_stack((np.arange(len(bins)), np.zeros((len(bins), 3)), hist[bins]))
    res = o2.emissivity("C", "Z5", 33.7, 2, ["EXCIT", "RECOM"], o2.two_gauss_profile(NE_MAIN, TE_MAIN), observed, bins / 1000.0)
    return np.array([res.flux["EXCIT"], res.flux["RECOM"], res.flux["TOTAL"]]), res.reff_boundary


def _reff20():
    from co_monitor_b200.geometry.reff import load_reff

    return load_reff(SPACING)["Reff [m]"].to_numpy(dtype=float)


def _exchange_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    _, glob, w = _shard_weights(rank, world)
    res = PackedResults()
    res.add("hist_mm", _mm_histogram(glob, w, _reff20()))
    res.count("visible_voxels", int((w != 0).sum()))
    res.all_reduce()  # the single exchange of the path
    flux, boundary = _flux_from_histogram(res.floats["hist_mm"])  # every rank evaluates Stage 2 from the global histogram
    np.savez(os.path.join(out_dir, f"x{rank}.npz"), flux=flux, boundary=boundary, glob=glob, w=w,
             visible=res.counters["visible_voxels"])
    dist.barrier()
    dist.destroy_process_group()


def test_block_cyclic_shards_and_histogram_exchange(tmp_path):
    """Two ranks, block-cyclic shards of the z-columns, one all-reduce of the 4096-bin histogram, Stage 2 from the reduced
    histogram on every rank == the per-voxel Stage-2 oracle on the union of the shards (the multi-GPU path of bench.py and
    scripts/volume_flux.py, with the CPU oracle standing in for the kernels)."""
    import torch.multiprocessing as mp

    from oracle import stage2 as o2

    port = _free_port()
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (np.load(tmp_path / f"x{r}.npz") for r in (0, 1))
    glob = np.concatenate((r0["glob"], r1["glob"]))
    w = np.concatenate((r0["w"], r1["w"]))
    assert len(np.unique(glob)) == len(glob)  # the shards do not overlap
    vis = w != 0
    assert vis.sum() > 50 and int(r0["visible"]) == int(r1["visible"]) == int(vis.sum())
    # reference: per-voxel Stage 2 on the union (rows ordered by voxel index, as a Stage-1 file would hold them)
    order = np.argsort(glob[vis])
    observed = np.column_stack((glob[vis][order], np.zeros((vis.sum(), 3)), w[vis][order]))
    ref = o2.emissivity("C", "Z5", 33.7, 2, ["EXCIT", "RECOM"], o2.two_gauss_profile(NE_MAIN, TE_MAIN), observed, _reff20())
    want = np.array([ref.flux["EXCIT"], ref.flux["RECOM"], ref.flux["TOTAL"]])
    assert want[2] > 0
    for r in (r0, r1):
        np.testing.assert_allclose(r["flux"], want, rtol=1e-12)
        assert float(r["boundary"]) == ref.reff_boundary
