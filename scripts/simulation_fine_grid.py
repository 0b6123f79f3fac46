#!/usr/bin/env python
"""The Stage-1 drop-in on a fine lattice, file included: wall-clock of the constructor
(setup, aperture calibration, chunked launches, compaction, coordinates of the visible voxels) and of the writers.

    python scripts/simulation_fine_grid.py --spacing 1 --element C
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import pandas as pd

    from co_monitor_b200.geometry import stage1_io
    from co_monitor_b200.geometry.simulation import Simulation

    ap = argparse.ArgumentParser()
    ap.add_argument("--spacing", type=float, default=1.0)
    ap.add_argument("--element", default="C")
    ap.add_argument("--pandas-rows", type=int, default=2_000_000, help="rows of the result also written with pandas, for the rate")
    a = ap.parse_args()
    spacing = int(a.spacing) if float(a.spacing).is_integer() else a.spacing
    out = tempfile.mkdtemp(prefix="stage1_")
    Simulation("C", 10, 10, 30, 20, plot=False, verbose=False, output_dir=out)  # CUDA context, library, caches
    t0 = time.perf_counter()
    sim = Simulation(a.element, slits_number=10, distance_between_points=spacing, crystal_height_step=30,
                     crystal_length_step=20, plot=False, verbose=False, output_dir=out)
    t_sim = time.perf_counter() - t0
    rows = len(sim.ddf)
    t0 = time.perf_counter()
    text = sim.save_to_file()
    t_text = time.perf_counter() - t0
    t0 = time.perf_counter()
    binary = sim.save_to_file(binary=True)
    t_bin = time.perf_counter() - t0
    t0 = time.perf_counter()
    back = stage1_io.read(binary, mesh=sim.mesh)
    t_read = time.perf_counter() - t0
    same = back.equals(pd.DataFrame(sim.ddf))
    k = min(rows, a.pandas_rows)
    head = pd.DataFrame(sim.ddf).iloc[:k]
    t0 = time.perf_counter()
    head.to_csv(os.path.join(out, "pandas.csv"), sep=";", header=True, index=False)
    t_pandas = time.perf_counter() - t0
    with open(text, "rb") as f1, open(os.path.join(out, "pandas.csv"), "rb") as f2:
        want = f2.read()
        identical_prefix = f1.read(len(want)) == want if k == rows else f1.read(len(want) - 1) == want[:-1]
    print(json.dumps({
        "element": a.element, "spacing_mm": a.spacing, "voxels": sim.mesh.lattice.n_points, "tests": sim.stats["tests"],
        "visible_voxels": rows, "passing_rays": sim.stats["passing_rays"], "near_edge_rays": sim.stats["near_edge_rays"],
        "simulation_constructor_s": t_sim, "text_file_s": t_text, "text_file_bytes": os.path.getsize(text),
        "text_rows_per_s": rows / t_text, "pandas_rows_per_s": k / t_pandas, "text_equals_pandas_on_first_rows": identical_prefix,
        "binary_file_s": t_bin, "binary_file_bytes": os.path.getsize(binary), "binary_read_s": t_read,
        "binary_round_trip_equal": bool(same), "host_threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
