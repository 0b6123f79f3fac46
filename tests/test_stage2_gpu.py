"""Stage-2 parity on the GPU: the Emissivity drop-in (CUDA path through the C ABI) against outputs of the
unmodified reference class (tests/golden/stage2_reference.npz) and against the CPU oracle."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

LINES = {"B": ("Z4", 48.6), "C": ("Z5", 33.7), "N": ("Z6", 24.8), "O": ("Z7", 19.0)}
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]
NE_SWEEP = [4e13, 0, 0.937, 1.8e12, 0.5, 0.11]
TE_SWEEP_LO = np.array([500, 0, 0.175, 30, 0.38, 0.07])
TE_SWEEP_HI = np.array([10000, 0, 0.175, 1500, 0.38, 0.07])
RTOL = 1e-12  # fp64 path; differences are rounding-order only (north-star bar: 1e-6 on intensities)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stage2_reference.npz"))


def run(el, profile, tmp_path, time="Test"):
    from co_monitor_b200.emissivity.reader import Emissivity

    ion, wl = LINES[el]
    return Emissivity(el, ion, wl, 2, ["EXCIT", "RECOM"], "Reff_coordinates-10_mm", profile, time,
                      output_dir=tmp_path, verbose=False)


def check(em, gold, tag):
    df = em.df_prof_frac_ab_pec
    assert list(df.columns) == list(gold[f"{tag}_columns"])
    assert np.all(np.diff(df["Reff [m]"].to_numpy()) >= 0)  # sorted by Reff like the reference
    rows = df.sort_values("idx_sel_plas_points", kind="stable").to_numpy(dtype=float)
    stride = int(gold[f"{tag}_stride"][0])
    assert len(rows) == int(gold[f"{tag}_nrows"][0])
    ref = gold[f"{tag}_rows"]
    assert np.array_equal(rows[::stride, 0], ref[:, 0])
    np.testing.assert_allclose(rows[::stride], ref, rtol=RTOL, atol=0)
    np.testing.assert_allclose(rows.sum(axis=0), gold[f"{tag}_colsum"], rtol=1e-11)
    flux = np.array([em.flux["EXCIT"], em.flux["RECOM"], em.flux["TOTAL"]])
    np.testing.assert_allclose(flux, gold[f"{tag}_flux"], rtol=RTOL)
    assert em.total_emissivity == str(gold[f"{tag}_total_str"][0])
    assert em.reff_boundary == float(gold[f"{tag}_reff_boundary"][0])


@pytest.mark.parametrize("el", list("CBNO"))
def test_reference_parity_main_profile(gold, reff10, tmp_path, el):
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile

    em = run(el, TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df, tmp_path)
    check(em, gold, f"main_{el}")
    out = tmp_path / f"Emissivity ({el}_{LINES[el][0]}_Test).dat"
    assert out.exists()
    with open(out) as fh:
        head = [fh.readline() for _ in range(8)]
    assert head[0].startswith("# Index(['idx_sel_plas_points'") and any("Total emissivity: " + em.total_emissivity in h for h in head)
    saved = np.loadtxt(out)
    assert saved.shape == (len(em.df_prof_frac_ab_pec), 15)


def test_reference_parity_experimental_profiles(gold, reff10, tmp_path):
    from co_monitor_b200.emissivity.kinetic_profiles import ExperimentalProfile, TestExperimentalProfile

    check(run("C", ExperimentalProfile("report_20181011_012@5_5000_v_1").profiles_df, tmp_path, "exp"), gold, "exp_C")
    check(run("O", TestExperimentalProfile("20230215.058", "8.0105").profiles_df, tmp_path, "tst"), gold, "tst_O")


def test_sweep_slices_and_factorised_form(gold, reff10, tmp_path):
    """Per-slice Emissivity and the histogram + sweep kernels against the reference on slices 0, 499, 999."""
    import torch

    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile

    slices = (0, 499, 999)
    profs = [TwoGaussSumProfile(NE_SWEEP, list(TE_SWEEP_LO + (TE_SWEEP_HI - TE_SWEEP_LO) * (s / 999))).profiles_df
             for s in slices]
    for el in "CO":
        ems = [run(el, p, tmp_path, s) for s, p in zip(slices, profs)]
        for s, em in zip(slices, ems):
            check(em, gold, f"sweep{s}_{el}")
        tables = cached_line_tables(el, LINES[el][0], LINES[el][1], ("EXCIT", "RECOM"), str(ems[0]._input_root))
        base = ems[0].plas_coords_with_rad_frac
        hist, boundary = tables.weight_histogram(profs[0]["Reff [m]"].to_numpy(), base["Reff [m]"].to_numpy(),
                                                 base["total_intensity_fraction"].to_numpy())
        assert boundary == ems[0].reff_boundary
        batch = np.stack([p[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy() for p in profs])
        flux = tables.sweep(batch, hist, 0.02).cpu().numpy()
        for i, s in enumerate(slices):
            np.testing.assert_allclose(flux[i], gold[f"sweep{s}_{el}_flux"], rtol=1e-11)
        assert abs(float(hist.sum()) - ems[0].df_prof_frac_ab_pec["total_intensity_fraction"].sum()) < 1e-12


def test_oracle_parity_random_reff_and_unsorted_profile(tmp_path):
    """Synthetic Reff (random, with NaNs and exact ties) and a shuffled profile table: literal-argmin path."""
    from co_monitor_b200.emissivity.engine import cached_line_tables
    from oracle import stage2 as o2

    rng = np.random.default_rng(0)
    n = 5000
    reff = np.round(rng.random(n) * 0.6, 3)
    reff[rng.random(n) < 0.1] = np.nan
    reff[:50] = np.round(np.arange(50) * 0.0105, 4)  # x.xxx5 values: equidistant from two profile rows
    w = rng.random(n) * 1e-5
    obs = np.column_stack((np.arange(n), np.zeros((n, 3)), w))
    prof = o2.two_gauss_profile(NE_MAIN, TE_MAIN)
    tables = cached_line_tables("N", "Z6", 24.8, ("EXCIT", "RECOM"), None)
    otab = o2.load_tables("N", "Z6", 24.8, ["EXCIT", "RECOM"])
    for profile in (prof, prof[rng.permutation(len(prof))]):
        ref = o2.emissivity("N", "Z6", 24.8, 2, ["EXCIT", "RECOM"], profile, obs, reff, tables=otab)
        out = tables.integrate(profile, 0.02, reff, w)
        assert np.array_equal(np.flatnonzero(out["valid"]), ref.rows[:, 0].astype(int))
        np.testing.assert_allclose(out["columns"][out["valid"]], ref.rows[:, 6:], rtol=RTOL, atol=0)
        np.testing.assert_allclose(out["flux"], [ref.flux["EXCIT"], ref.flux["RECOM"], ref.flux["TOTAL"]], rtol=RTOL)
        assert out["boundary"] == ref.reff_boundary


def test_find_nearest_reference_unit_test():
    from co_monitor_b200.emissivity.reader import Emissivity

    nrs = [1, 22, 50, 16, 7, 19, 100, 4, 6, 211]
    assert Emissivity._find_nearest(nrs, 75) == 2
    assert Emissivity._find_nearest(nrs, 76) == 6
    assert Emissivity._find_nearest(nrs, 0) == 0
