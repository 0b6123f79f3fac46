#!/usr/bin/env python
"""BASELINE.json config 4: 1000 synthetic T_e profile time slices over cached geometric weights, B/C/N/O.

Weights = the shipped Stage-1 files (device-cached), Reff = 10 mm grid (up-sampled), profiles = emissivity/main.py's
sweep extended from 11 to 1000 slices.  Two ways are timed with CUDA events:
  factorised: one com_weight_histogram pass per channel + com_emissivity_sweep over all slices (what the sweep needs)
  naive     : com_emissivity_integrate once per slice (what looping the reference's Emissivity class amounts to)
Prints one JSON line; slices 0/499/999 are compared with the golden reference fluxes when the fixture is present.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LINES = {"B": ("Z4", 48.6), "C": ("Z5", 33.7), "N": ("Z6", 24.8), "O": ("Z7", 19.0)}
NE = [4e13, 0, 0.937, 1.8e12, 0.5, 0.11]
TE_LO = np.array([500, 0, 0.175, 30, 0.38, 0.07])
TE_HI = np.array([10000, 0, 0.175, 1500, 0.38, 0.07])


def main(n_slices=1000, naive_slices=50):
    import pandas as pd
    import torch

    from co_monitor_b200.emissivity.engine import LineTables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.json_reader import input_root
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import interpolate_reff

    with contextlib.redirect_stdout(io.StringIO()):
        reff_all = interpolate_reff(CuboidMesh(10).lattice, 15)
    profs = np.stack([
        TwoGaussSumProfile(NE, list(TE_LO + (TE_HI - TE_LO) * (s / (n_slices - 1)))).profiles_df[
            ["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy() for s in range(n_slices)])
    profs_d = torch.from_numpy(profs).cuda()
    gold_path = os.path.join(ROOT, "tests", "golden", "stage2_reference.npz")
    gold = np.load(gold_path) if os.path.exists(gold_path) else None
    out = {"n_slices": n_slices, "channels": {}}
    total_fact_ms = 0.0
    for el, (ion, wl) in LINES.items():
        obs = pd.read_csv(input_root() / "geometry" / "observed_plasma_volume" / el /
                          f"{el}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat", sep=";")
        reff = reff_all[obs["idx_sel_plas_points"].to_numpy()]
        keep = ~np.isnan(reff)
        reff_d = torch.from_numpy(np.ascontiguousarray(reff[keep])).cuda()
        w_d = torch.from_numpy(np.ascontiguousarray(obs["total_intensity_fraction"].to_numpy()[keep])).cuda()
        tables = LineTables(el, ion, wl, ["EXCIT", "RECOM"])
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(2):  # warm-up + timed
            e[0].record()
            hist, boundary = tables.weight_histogram(profs[0][:, 0], reff_d, w_d)
            e[1].record()
            flux = tables.sweep(profs_d, hist, 0.02)
            e[2].record()
            torch.cuda.synchronize()
        hist_ms, sweep_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        # naive: one voxel pass per slice
        e[0].record()
        naive = []
        for s in range(naive_slices):
            naive.append(tables.integrate(profs[s], 0.02, reff_d, w_d, per_voxel=False)["flux"])
        e[1].record()
        torch.cuda.synchronize()
        naive_ms = e[0].elapsed_time(e[1]) / naive_slices
        flux_h = flux.cpu().numpy()
        err_naive = float(np.max(np.abs(flux_h[:naive_slices] / np.array(naive) - 1)))
        rec = {"voxels": int(keep.sum()), "histogram_ms": hist_ms, "sweep_ms": sweep_ms, "ms_per_slice_factorised": sweep_ms / n_slices,
               "ms_per_slice_naive_voxel_pass": naive_ms, "max_rel_diff_factorised_vs_naive": err_naive,
               "flux_total_slice0": float(flux_h[0, 2]), "flux_total_last": float(flux_h[-1, 2])}
        if gold is not None and n_slices == 1000 and f"sweep0_{el}_flux" in gold:
            rec["max_rel_err_vs_reference_slices_0_499_999"] = float(max(
                np.max(np.abs(flux_h[s] / gold[f"sweep{s}_{el}_flux"] - 1)) for s in (0, 499, 999)))
        out["channels"][el] = rec
        total_fact_ms += hist_ms + sweep_ms
    out["total_ms_factorised_all_channels"] = total_fact_ms
    out["voxel_slices_per_s_factorised"] = sum(c["voxels"] for c in out["channels"].values()) * n_slices / (total_fact_ms * 1e-3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
