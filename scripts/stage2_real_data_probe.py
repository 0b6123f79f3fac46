#!/usr/bin/env python
"""Stage-2 voxel pass (com_weight_histogram_mm) on REAL data of a fine lattice: Reff millimetres resampled from the shipped
grid (spatially smooth: neighbouring voxels share bins) and the weights of an actual Stage-1 run (visible voxels are
clustered), as opposed to the random synthetic channel of bench.py's roofline leg.

    python scripts/stage2_real_data_probe.py --spacing 1 --element C
"""
import argparse
import contextlib
import ctypes as C
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import resample_reff_device

    ap = argparse.ArgumentParser()
    ap.add_argument("--spacing", type=float, default=1.0)
    ap.add_argument("--element", default="C")
    a = ap.parse_args()
    lib = _lib.load()
    with contextlib.redirect_stdout(io.StringIO()):
        lat = CuboidMesh(a.spacing).lattice
        scene = make_channel(a.element, 10, 30, 20, lattice=lat)
    n = lat.n_points
    reff_mm, _ = resample_reff_device(lat)
    w = torch.zeros(n, dtype=torch.float64, device="cuda")
    chunk = 1 << 26
    plan = scene.plan(lat, 0, min(chunk, n))
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        plan.rebind(lo, hi).launch()
        w[lo:hi].copy_(plan.weight[: hi - lo])
    torch.cuda.synchronize()
    del plan
    hist = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device="cuda")
    times = []
    for _ in range(7):
        hist.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.com_weight_histogram_mm(C.c_void_p(reff_mm.data_ptr()), C.c_void_p(w.data_ptr()), n, None,
                                               C.c_void_p(hist.data_ptr()), None), "com_weight_histogram_mm")
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.mean(times[2:]))
    ok = reff_mm.view(torch.int16) >= 0  # 0xFFFF = -1
    want = torch.bincount(reff_mm.view(torch.int16)[ok].to(torch.int64), weights=w[ok], minlength=_lib.COM_MM_BINS)
    err = float(((hist - want[: _lib.COM_MM_BINS]).abs().max() / want.abs().max()).item())
    print(json.dumps({"element": a.element, "spacing_mm": a.spacing, "voxels": n, "visible_fraction": float((w != 0).float().mean().item()),
                      "ms": ms, "GBps": 10.0 * n / (ms * 1e-3) / 1e9, "launch_ms": times, "max_rel_err_vs_bincount": err}))


if __name__ == "__main__":
    main()
