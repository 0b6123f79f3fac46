"""TEST INFRASTRUCTURE - fixture generator for Stage 2 (runs only where /root/reference exists).

Runs the UNMODIFIED reference ``co_monitor.emissivity.reader.Emissivity`` (and its table classes)
through the shims of oracle/_refshim.py, from a scratch working directory laid out the way the
reference expects (``input_files/...`` relative to the CWD):

  * PEC/, kinetic_profiles/, geometry/observed_plasma_volume/  -> the reference's own files
  * fractional_abundance/fractional_abundance_<el>.txt         -> the shipped .dat (fixes the reference's D7)
  * geometry/reff/Reff_coordinates-10_mm.dat                   -> missing blob of the reference; replaced by the
    repo's deterministic trilinear up-sampling of the shipped 15 mm grid (co_monitor_b200.geometry.reff)

and stores what it produces as tests/golden/stage2_reference.npz (per-voxel columns keyed by
``idx_sel_plas_points``, the flux totals, the profile / FA / PEC tables at sampled points).

    python oracle/make_golden_stage2.py                 # tests/golden/stage2_reference.npz
    python oracle/make_golden_stage2.py --transitions   # tests/golden/stage2_transitions_reference.npz
"""
from __future__ import annotations

import contextlib
import io
import os
import shutil
import sys
import tempfile
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from _refshim import REFERENCE_ROOT, install_interp2d, install_stubs  # noqa: E402

install_stubs()
install_interp2d()
warnings.filterwarnings("ignore")

OUT = os.path.join(ROOT, "tests", "golden", "stage2_reference.npz")

LINES = {"B": ("Z4", 48.6), "C": ("Z5", 33.7), "N": ("Z6", 24.8), "O": ("Z7", 19.0)}  # reader.py:469-474
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]    # __main__.py:35-36
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]
NE_SWEEP = [4e13, 0, 0.937, 1.8e12, 0.5, 0.11]  # emissivity/main.py:39-41
TE_SWEEP_LO = [500, 0, 0.175, 30, 0.38, 0.07]
TE_SWEEP_HI = [10000, 0, 0.175, 1500, 0.38, 0.07]


def scratch_cwd() -> str:
    from . import reff as oreff

    tmp = tempfile.mkdtemp(prefix="comonitor_ref_")
    inp = os.path.join(tmp, "input_files")
    os.makedirs(os.path.join(inp, "geometry", "reff"))
    os.makedirs(os.path.join(inp, "fractional_abundance"))
    ref_in = os.path.join(REFERENCE_ROOT, "input_files")
    os.symlink(os.path.join(ref_in, "PEC"), os.path.join(inp, "PEC"))
    os.symlink(os.path.join(ref_in, "kinetic_profiles"), os.path.join(inp, "kinetic_profiles"))
    os.symlink(os.path.join(ref_in, "geometry", "observed_plasma_volume"),
               os.path.join(inp, "geometry", "observed_plasma_volume"))
    for el in "BCNO":
        os.symlink(os.path.join(ref_in, "fractional_abundance", f"fractional_abundance_{el}.dat"),
                   os.path.join(inp, "fractional_abundance", f"fractional_abundance_{el}.txt"))
    with contextlib.redirect_stdout(io.StringIO()):
        oreff.write_file(oreff.file_frame(10, 15), os.path.join(inp, "geometry", "reff", "Reff_coordinates-10_mm.dat"))
    return tmp


def run_emissivity(element, profiles, time="Test", concentration=2, transitions=("EXCIT", "RECOM")):
    from co_monitor.emissivity.reader import Emissivity

    ion, wl = LINES[element]
    with contextlib.redirect_stdout(io.StringIO()):
        em = Emissivity(element, ion, wl, concentration, list(transitions), "Reff_coordinates-10_mm", profiles, time)
    df = em.df_prof_frac_ab_pec.sort_values("idx_sel_plas_points", kind="stable")
    return em, df


def pack(out, tag, em, df, stride=1):
    cols = list(df.columns)
    arr = df.to_numpy(dtype=float)
    out[f"{tag}_columns"] = np.array(cols)
    out[f"{tag}_nrows"] = np.array([len(arr)])
    out[f"{tag}_rows"] = arr[::stride]
    out[f"{tag}_stride"] = np.array([stride])
    w = df["total_intensity_fraction"]
    out[f"{tag}_flux"] = np.array([(df["Emissivity_EXCIT"] * w).sum(), (df["Emissivity_RECOM"] * w).sum(),
                                   (df["Emissivity_TOTAL"] * w).sum()])
    out[f"{tag}_colsum"] = arr.sum(axis=0)
    out[f"{tag}_total_str"] = np.array([em.total_emissivity])
    out[f"{tag}_reff_boundary"] = np.array([em.reff_boundary])


OUT_TRANSITIONS = os.path.join(ROOT, "tests", "golden", "stage2_transitions_reference.npz")
#: other transition lists the reference accepts: single transitions, a different order, and the charge-exchange block
#: of the ADAS files, which reader.py:355-369 treats like EXCIT (its FA column logic only knows "RECOM" / else)
TRANSITION_CASES = [("C", ("EXCIT",)), ("C", ("RECOM",)), ("C", ("EXCIT", "RECOM", "CHEXC")), ("O", ("RECOM", "EXCIT")),
                    ("N", ("CHEXC",))]


def main_transitions():
    from co_monitor.emissivity.kinetic_profiles import TwoGaussSumProfile

    tmp = scratch_cwd()
    old = os.getcwd()
    os.chdir(tmp)
    out = {}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            prof = TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df
        for el, trs in TRANSITION_CASES:
            em, df = run_emissivity(el, prof, transitions=trs)
            tag = f"{el}_{'_'.join(trs)}"
            arr = df.to_numpy(dtype=float)
            w = df["total_intensity_fraction"]
            out[f"{tag}_columns"] = np.array(list(df.columns))
            out[f"{tag}_nrows"] = np.array([len(arr)])
            out[f"{tag}_rows"] = arr[::9]
            out[f"{tag}_colsum"] = arr.sum(axis=0)
            names = ["Emissivity_RECOM" if "RECOM" in t else f"Emissivity_{t}" for t in trs] + ["Emissivity_TOTAL"]
            out[f"{tag}_flux"] = np.array([(df[n] * w).sum() for n in names])
            out[f"{tag}_total_str"] = np.array([em.total_emissivity])
            out[f"{tag}_reff_boundary"] = np.array([em.reff_boundary])
            print(tag, len(df), em.total_emissivity)
    finally:
        os.chdir(old)
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez_compressed(OUT_TRANSITIONS, **out)
    print("wrote", OUT_TRANSITIONS, f"{os.path.getsize(OUT_TRANSITIONS)/1e6:.2f} MB", len(out), "arrays")


def main():
    from co_monitor.emissivity.fractional_abundance import FractionalAbundance
    from co_monitor.emissivity.kinetic_profiles import ExperimentalProfile, TestExperimentalProfile, TwoGaussSumProfile
    from co_monitor.emissivity.pec import PEC

    tmp = scratch_cwd()
    old = os.getcwd()
    os.chdir(tmp)
    out = {}
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            main_prof = TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df
            exp_prof = ExperimentalProfile("report_20181011_012@5_5000_v_1", plot=False).profiles_df
            tst_prof = TestExperimentalProfile("20230215.058", "8.0105").profiles_df
        out["profile_main"] = main_prof.to_numpy()
        out["profile_experimental"] = exp_prof.to_numpy()
        out["profile_test_experimental"] = tst_prof.to_numpy()
        out["profile_columns"] = np.array(list(main_prof.columns))

        # table classes at sampled points
        rng = np.random.default_rng(0)
        for el, (ion, wl) in LINES.items():
            fa = FractionalAbundance(el, ion).df_interpolated_frac_ab
            rows = np.sort(rng.choice(len(fa), 400, replace=False))
            out[f"fa_{el}_columns"] = np.array(list(fa.columns))
            out[f"fa_{el}_n"] = np.array([len(fa)])
            out[f"fa_{el}_rows_idx"] = rows
            out[f"fa_{el}_rows"] = fa.to_numpy()[rows]
            for tr in ("EXCIT", "RECOM"):
                pec = PEC(el, wl, ion[-1], tr, 2000)
                ii = rng.integers(0, 2000, 600)
                jj = rng.integers(0, 2000, 600)
                ii[:4], jj[:4] = (0, 0, 1999, 1999), (0, 1999, 0, 1999)
                out[f"pec_{el}_{tr}_ij"] = np.stack((ii, jj))
                out[f"pec_{el}_{tr}_vals"] = pec.interpolated_pec[ii, jj]
                out[f"pec_{el}_{tr}_nodes_ne"] = pec.ne_nodes
                out[f"pec_{el}_{tr}_nodes_te"] = pec.te_nodes
                out[f"pec_{el}_{tr}_table"] = pec.pec_data_array

        # the four channels with the two-Gaussian profile of __main__.py / reader.main
        for el in "BCNO":
            em, df = run_emissivity(el, main_prof)
            pack(out, f"main_{el}", em, df, stride=1 if el == "C" else 9)
            print(el, "rows", len(df), "TOTAL", em.total_emissivity, "boundary", em.reff_boundary)
        # experimental profile (10 000 rows, 540 distinct Reff: first duplicate wins)
        em, df = run_emissivity("C", exp_prof, time="exp")
        pack(out, "exp_C", em, df, stride=9)
        print("C experimental", len(df), em.total_emissivity)
        em, df = run_emissivity("O", tst_prof, time="tst")
        pack(out, "tst_O", em, df, stride=9)
        print("O test-experimental", len(df), em.total_emissivity)
        # three slices of the 1000-slice temperature sweep (config 4)
        with contextlib.redirect_stdout(io.StringIO()):
            for s in (0, 499, 999):
                te = list(np.array(TE_SWEEP_LO) + (np.array(TE_SWEEP_HI) - np.array(TE_SWEEP_LO)) * (s / 999))
                prof = TwoGaussSumProfile(NE_SWEEP, te).profiles_df
                for el in "CO":
                    em, df = run_emissivity(el, prof, time=s)
                    pack(out, f"sweep{s}_{el}", em, df, stride=97)
    finally:
        os.chdir(old)
        shutil.rmtree(tmp, ignore_errors=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, f"{os.path.getsize(OUT)/1e6:.2f} MB", len(out), "arrays")


if __name__ == "__main__":
    if "--transitions" in sys.argv:
        main_transitions()
    else:
        main()
