#!/usr/bin/env python
"""Stage 1 on fine synthetic lattices (BASELINE.json configs 3 and 5): implicit voxel generation, chunked launches.

    python scripts/large_grid.py --spacing 1 --elements C            # 1350 x 800 x 250 = 2.7e8 voxels, 1.62e11 tests
    python scripts/large_grid.py --spacing 0.65 --elements BCNO      # 2076 x 1230 x 384 = 9.8e8 voxels

Prints one JSON line per channel: tests/s (CUDA events), visible voxels, passing rays, the volume-weighted sum
sum(w) * spacing^3 (a spacing-independent integral: it must agree with the 10 mm grid to a few per cent) and the
per-chunk timings.  Nothing of size P x C is materialised; per-voxel outputs are kept per chunk only.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh

    ap = argparse.ArgumentParser()
    ap.add_argument("--spacing", type=float, default=1.0)
    ap.add_argument("--elements", default="C")
    ap.add_argument("--chunk", type=int, default=1 << 26, help="voxels per launch")
    ap.add_argument("--h", type=int, default=30)
    ap.add_argument("--l", type=int, default=20)
    ap.add_argument("--slits", type=int, default=10)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--image", action="store_true", help="also accumulate the 0.1 mm detector image (config 3 extension)")
    ap.add_argument("--kernel-times", action="store_true", help="per-chunk times of the FP32 filter kernel and of the fp64 kernels")
    ap.add_argument("--only-chunks", default="", help="comma-separated chunk indices to run (default: all)")
    ap.add_argument("--mode", default="production", choices=["production", "percull", "nocull"],
                    help="production: column-interval filter + weight-only kernel; percull: per-ray filter with its run-level "
                         "cull; nocull: every ray's detector test evaluated individually (the headline of bench.py)")
    a = ap.parse_args()
    import ctypes as C

    from co_monitor_b200 import _lib

    lib = _lib.load()
    only = {int(x) for x in a.only_chunks.split(",") if x}
    with contextlib.redirect_stdout(io.StringIO()):
        lat = CuboidMesh(a.spacing).lattice
    P = lat.n_points
    for el in a.elements:
        with contextlib.redirect_stdout(io.StringIO()):
            scene = make_channel(el, a.slits, a.h, a.l, lattice=lat, intervals=a.mode == "production",
                                 run_cull=a.mode != "nocull")
        best = None
        img = scene.enable_detector_image(0.1) if a.image else None
        for _ in range(a.repeat):
            if img is not None:
                img.zero_()
            sum_w, vis, rays, cand, near, ms_total, chunks, ktimes, cert = 0.0, 0, 0, 0, 0, 0.0, [], [], 0
            for j, lo in enumerate(range(0, P, a.chunk)):
                if only and j not in only:
                    continue
                hi = min(P, lo + a.chunk)
                if a.kernel_times:
                    lib.com_kernel_timing_enable(1)
                plan = scene.plan(lat, lo, hi)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                plan.launch()
                e1.record()
                torch.cuda.synchronize()
                st = plan.stats()
                if st["overflow"]:
                    raise SystemExit(f"candidate overflow in chunk [{lo},{hi}): {st}")
                ms = e0.elapsed_time(e1)
                if a.kernel_times:
                    ms3, n_l = (C.c_double * 3)(), C.c_int64()
                    lib.com_kernel_timing_read3(ms3, C.byref(n_l))
                    lib.com_kernel_timing_enable(0)
                    ktimes.append([round(ms3[0], 3), round(ms3[1], 3), round(ms3[2], 3), st["candidates"], st["passing_rays"]])
                ms_total += ms
                chunks.append(round(ms, 3))
                sum_w += float(plan.weight.sum().item())
                vis += st["visible_voxels"]
                rays += st["passing_rays"]
                cand += st["candidates"]
                cert += st["certain_candidates"]
                near += st["near_edge_rays"]
                del plan
            tests = P * a.h * a.l
            rec = {"element": el, "mode": a.mode, "spacing_mm": a.spacing, "lattice": [lat.nx, lat.ny, lat.nz], "voxels": P, "tests": tests,
                   "ms": ms_total, "tests_per_s": tests / (ms_total * 1e-3), "visible_voxels": vis, "passing_rays": rays,
                   "fp64_candidates": cand, "certain_candidates": cert, "near_edge_rays": near, "sum_w": sum_w,
                   "sum_w_times_voxel_volume_mm3": sum_w * a.spacing**3, "chunk_ms": chunks}
            if a.kernel_times:
                rec["chunk_filter_ms_weight_ms_exact_ms_candidates_passing"] = ktimes
            if img is not None:
                rec["detector_image"] = {"shape": list(img.shape), "pixel_mm": 0.1, "sum": float(img.sum().item()),
                                         "max_pixel": float(img.max().item()),
                                         "lit_pixels": int((img > 0).sum().item())}
            if best is None or rec["ms"] < best["ms"]:
                best = rec
        print(json.dumps(best))
        scene.close()


if __name__ == "__main__":
    main()
