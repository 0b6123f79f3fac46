#!/usr/bin/env python
"""Per-channel C/O-monitor line flux of a dense synthetic observation volume (BASELINE.json config 5), 1..8 GPUs.

    python scripts/volume_flux.py --spacing 0.65                                   # 9.8e8 voxels, one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/volume_flux.py --spacing 0.65                                      # sharded over the box

Everything happens on the device, nothing of size voxels x crystal points (or even voxels x 8 B per channel) is kept:
  Reff      the shipped 15 mm VMEC grid resampled onto this rank's block-cyclic shard of the lattice, as uint16
            millimetre bins (com_reff_resample; --reff quadratic: the analytic flux-surface model instead)
  Stage 1   visibility + weights in launches of --chunk voxels, each followed by one 10 B/voxel pass that adds the launch's
            weights to the channel's millimetre-bin histogram (--histogram fused: the exact kernel adds every passing
            ray to hist[reff_mm[voxel]] itself - fewer bytes, but many atomics on few bins)
  exchange  ONE all-reduce(SUM) over NVLink of the packed [channels, 4096] histograms plus the ray/voxel counters
  Stage 2   histogram -> profile rows -> flux per transition (com_histogram_rows + com_emissivity_sweep), every rank
The timed region (CUDA events, barrier on both sides, max over ranks) covers Stage 1, the exchange and Stage 2 of all
channels; the Reff resampling is timed separately (it is per lattice, not per channel or time slice).

Size-independent checks printed with the result: flux x voxel volume (a Riemann sum of the same integral: agrees with
the 10 mm grid's to a few per cent), the statistics summed over ranks, and - across runs at different N - the flux
itself (identical to ~1e-13: fp64 sums in a different order).
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]  # reader.py:476-479 / __main__.py:35-37
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]


def main():
    import numpy as np
    import torch
    import torch.distributed as dist

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import LineTables, _ptr
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import QuadraticReff, resample_reff_device, source_volume
    from co_monitor_b200.pipeline import LYMAN_ALPHA

    ap = argparse.ArgumentParser()
    ap.add_argument("--spacing", type=float, default=0.65)
    ap.add_argument("--elements", default="BCNO")
    ap.add_argument("--chunk", type=int, default=1 << 26, help="voxels per Stage-1 launch")
    ap.add_argument("--h", type=int, default=30)
    ap.add_argument("--l", type=int, default=20)
    ap.add_argument("--slits", type=int, default=10)
    ap.add_argument("--repeat", type=int, default=3, help="timed repetitions after one warm-up; the best is reported")
    ap.add_argument("--reff", choices=("resample", "quadratic"), default="resample")
    ap.add_argument("--cols-per-block", type=int, default=8)
    ap.add_argument("--candidate-fraction", type=float, default=0.0)
    ap.add_argument("--histogram", choices=("pass", "fused"), default="pass",
                    help="pass: one 10 B/voxel histogram pass over each launch's weights (default); fused: the exact kernel "
                         "adds every passing ray to the histogram itself (more atomics on few bins)")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("volume_flux.py: no CUDA device (the accelerated path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    with contextlib.redirect_stdout(io.StringIO()):
        full = CuboidMesh(a.spacing).lattice
    lat = full.shard(world, rank, cols_per_block=a.cols_per_block)
    n = lat.n_points
    elements = list(a.elements)
    n_ch = len(elements)

    # ---- Reff of this rank's voxels, on the device ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    model = QuadraticReff(15) if a.reff == "quadratic" else None
    source = None if model else source_volume(15)
    reff_mm = None
    for _ in range(2):  # second pass is the timed one
        del reff_mm
        e0.record()
        reff_mm, _ = model.on_device(lat, device=dev) if model else resample_reff_device(lat, device=dev, source=source)
        e1.record()
        torch.cuda.synchronize()
    reff_ms = e0.elapsed_time(e1)

    # ---- scenes, tables, buffers ----
    profile = TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy()
    prof_t = torch.from_numpy(np.ascontiguousarray(profile, dtype=np.float64)).to(dev)
    K = len(profile)
    prof_sorted = int(np.all(np.diff(profile[:, 0]) >= 0))
    prof_max = float(profile[:, 0].max())
    chunks = [(lo, min(n, lo + a.chunk)) for lo in range(0, n, a.chunk)]
    packed = torch.zeros(n_ch * _lib.COM_MM_BINS + n_ch * _lib.N_STATS, dtype=torch.float64, device=dev)  # histograms | counters
    hist = packed[: n_ch * _lib.COM_MM_BINS].view(n_ch, _lib.COM_MM_BINS)
    counters = packed[n_ch * _lib.COM_MM_BINS:].view(n_ch, _lib.N_STATS)
    stats = torch.zeros((n_ch, len(chunks), _lib.N_STATS), dtype=torch.int64, device=dev)
    rows = torch.zeros((n_ch, K), dtype=torch.float64, device=dev)
    boundary = torch.zeros(n_ch, dtype=torch.float64, device=dev)
    flux = torch.zeros((n_ch, 3), dtype=torch.float64, device=dev)
    scenes, tables, plans = [], [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i, el in enumerate(elements):
            scenes.append(make_channel(el, a.slits, a.h, a.l, lattice=full))
            ion, wl = LYMAN_ALPHA[el]
            tables.append(LineTables(el, ion, wl, ["EXCIT", "RECOM"]))
    # one set of Stage-1 buffers, reused by every chunk of every channel (stream order keeps them consistent)
    plan_cap = max(hi - lo for lo, hi in chunks)
    fused = a.histogram == "fused"
    plan_buf = scenes[0].plan(lat, 0, plan_cap, candidate_fraction=a.candidate_fraction, device=dev,
                              reff_bin=reff_mm if fused else None, hist=hist[0] if fused else None)
    if not fused and any(lo % 8 for lo, _ in chunks):
        raise SystemExit("--chunk must be a multiple of 8 voxels for the histogram pass (16-byte aligned loads)")
    stream = torch.cuda.current_stream(dev)
    s_ptr = lambda: __import__("ctypes").c_void_p(stream.cuda_stream)  # noqa: E731

    def run_once():
        packed.zero_()
        for i in range(n_ch):
            plan_buf.scene = scenes[i]
            if fused:
                plan_buf.hist = hist[i]
            for j, (lo, hi) in enumerate(chunks):
                plan_buf.rebind(lo, hi, stats=stats[i, j]).launch()
                if not fused:
                    rc = lib.com_weight_histogram_mm(_ptr(reff_mm[lo:hi]), _ptr(plan_buf.weight), hi - lo, None, _ptr(hist[i]),
                                                     s_ptr())
                    _lib.check(rc, "com_weight_histogram_mm")
            counters[i].copy_(stats[i].sum(0).to(torch.float64))  # exact: counts < 2^53
        if world > 1:
            dist.all_reduce(packed)  # the single exchange: [channels, 4096] histograms + counters
        for i in range(n_ch):
            rc = lib.com_histogram_rows(_ptr(prof_t), K, prof_sorted, prof_max, _ptr(hist[i]), None, _ptr(rows[i]),
                                        _ptr(boundary[i:]), s_ptr())
            _lib.check(rc, "com_histogram_rows")
            rc = lib.com_emissivity_sweep(tables[i].handle, _ptr(prof_t), 1, K, _ptr(rows[i]), 0.02, _ptr(flux[i]), s_ptr())
            _lib.check(rc, "com_emissivity_sweep")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    best = None
    for it in range(a.repeat + 1):
        barrier()
        e0.record()
        run_once()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if stats[..., 6].any().item():
            bad = stats[..., 6].nonzero()[0].tolist()
            raise SystemExit(f"rank {rank}: candidate overflow in (channel, chunk) {bad}; raise --candidate-fraction")
        if it > 0 and (best is None or ms.item() < best):
            best = ms.item()

    if rank == 0:
        names = [f[0] for f in _lib.ComStats._fields_]
        cnt = counters.cpu().numpy()
        fl = flux.cpu().numpy()
        tests = full.n_points_global * a.h * a.l * n_ch
        vol = a.spacing ** 3
        out = {
            "workload": f"config 5: {full.nx}x{full.ny}x{full.nz} lattice ({a.spacing} mm), {n_ch} channels, "
                        f"{a.h * a.l} crystal points, {a.slits} slits; Stage 1 -> mm histogram -> Stage 2",
            "n_gpus": world, "voxels": full.n_points_global, "voxels_per_rank": n, "tests": tests,
            "ms": best, "tests_per_s": tests / (best * 1e-3), "reff_provider": a.reff, "reff_ms_per_rank": reff_ms,
            "launches_per_channel": len(chunks), "exchange": "one all-reduce(SUM) of %d fp64" % packed.numel(),
            "channels": {
                el: {
                    "flux": {"EXCIT": fl[i, 0], "RECOM": fl[i, 1], "TOTAL": fl[i, 2]},
                    "total_emissivity": f"{fl[i, 2]:.2e}",
                    "flux_times_voxel_volume_mm3": fl[i, 2] * vol,
                    "reff_boundary_m": float(boundary[i].item()),
                    **{k: int(cnt[i, names.index(k)]) for k in
                       ("tests", "candidates", "passing_rays", "near_edge_rays", "visible_voxels")},
                } for i, el in enumerate(elements)
            },
        }
        print(json.dumps(out))
    for sc in scenes:
        sc.close()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
