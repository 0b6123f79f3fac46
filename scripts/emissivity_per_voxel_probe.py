#!/usr/bin/env python
"""The per-voxel Stage-2 kernel (emissivity_kernel: what the Emissivity drop-in runs, reader.py:199-401 row by row) at
1e8 voxels: time, voxels/s and bytes against the HBM peak, flux-only and with the 9 per-voxel output columns.

    python scripts/emissivity_per_voxel_probe.py [--voxels 100000000] [--element C] [--once]

Every voxel is valid (random 3-decimal Reff inside the profile, positive weight): the worst case for the table
lookups, which are gathers into the profile (13 kB), the fractional-abundance grid (the rows the profile's T_e range
reaches) and the PEC nodes (<= 9 kB) - global memory served by L1.  `--once` runs each form once (for ncu).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.pipeline import LYMAN_ALPHA

    ap = argparse.ArgumentParser()
    ap.add_argument("--voxels", type=int, default=100_000_000)
    ap.add_argument("--element", default="C")
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    el = a.element
    tables = cached_line_tables(el, *LYMAN_ALPHA[el], ("EXCIT", "RECOM"), None)
    prof = np.ascontiguousarray(TwoGaussSumProfile([7e13, 0, 0.37, 9.8e12, 0.5, 0.11], [1870, 0, 0.155, 210, 0.38, 0.07])
                                .profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    K, nt = len(prof), 2
    n = a.voxels
    g = torch.Generator(device=dev).manual_seed(0)
    reff = torch.randint(0, int(prof[:, 0].max() * 1000), (n,), generator=g, device=dev).to(torch.float64) / 1000.0
    w = torch.rand(n, generator=g, device=dev, dtype=torch.float64) * 1e-6 + 1e-9
    prof_d = torch.from_numpy(prof).to(dev)
    flux = torch.zeros(nt + 1, dtype=torch.float64, device=dev)
    ws_bytes = lib.com_emissivity_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = 6547.5
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except (OSError, KeyError, ValueError):
        pass
    out = {"voxels": n, "element": el, "profile_rows": K, "hbm_peak_GBps": peak}
    for name, per_voxel in (("flux_only", False), ("per_voxel_columns", True)):
        cols = torch.empty((n, 4 + 2 * nt + 1), dtype=torch.float64, device=dev) if per_voxel else None
        valid = torch.empty(n, dtype=torch.uint8, device=dev) if per_voxel else None

        def run():
            rc = lib.com_emissivity_integrate(
                tables.handle, C.c_void_p(prof_d.data_ptr()), K, 1, float(prof[:, 0].max()), 0.02,
                C.c_void_p(reff.data_ptr()), None, None, C.c_void_p(w.data_ptr()), n, None, None, C.c_void_p(flux.data_ptr()),
                C.c_void_p(cols.data_ptr()) if per_voxel else None, C.c_void_p(valid.data_ptr()) if per_voxel else None,
                C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "com_emissivity_integrate")

        ms = []
        for i in range(1 if a.once else 6):
            flush.fill_(i)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = float(np.mean(ms[1:])) if len(ms) > 1 else ms[0]
        # reff_max pass (8 B/voxel) + evaluation pass (Reff 8 + w 8 [+ 9 columns x 8 + 1 valid byte])
        bytes_alg = n * (8 + 16 + (8 * (4 + 2 * nt + 1) + 1 if per_voxel else 0))
        out[name] = {"ms": t, "voxels_per_s": n / (t * 1e-3), "algorithmic_bytes": bytes_alg,
                     "GBps": bytes_alg / (t * 1e-3) / 1e9, "frac_of_hbm_peak": bytes_alg / (t * 1e-3) / 1e9 / peak,
                     "flux": [float(x) for x in flux.cpu().numpy()]}
        del cols, valid
    print(json.dumps(out))


if __name__ == "__main__":
    main()
