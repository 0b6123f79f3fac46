"""Stage-2 host table classes (PEC, fractional abundance, kinetic profiles) against fixtures produced by the
unmodified reference classes (tests/golden/stage2_reference.npz, oracle/make_golden_stage2.py)."""
import os

import numpy as np
import pytest

from co_monitor_b200.emissivity.fractional_abundance import FractionalAbundance
from co_monitor_b200.emissivity.kinetic_profiles import ExperimentalProfile, TestExperimentalProfile, TwoGaussSumProfile
from co_monitor_b200.emissivity.pec import PEC

LINES = {"B": ("Z4", 48.6), "C": ("Z5", 33.7), "N": ("Z6", 24.8), "O": ("Z7", 19.0)}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stage2_reference.npz"))


def test_two_gauss_profile_bit_exact(gold):
    df = TwoGaussSumProfile([7e13, 0, 0.37, 9.8e12, 0.5, 0.11], [1870, 0, 0.155, 210, 0.38, 0.07]).profiles_df
    assert list(df.columns) == list(gold["profile_columns"]) == ["Reff [m]", "T_e [eV]", "n_e [m-3]"]
    assert np.array_equal(df.to_numpy(), gold["profile_main"])
    assert len(df) == 539


def test_experimental_profiles(gold):
    exp = ExperimentalProfile("report_20181011_012@5_5000_v_1").profiles_df
    assert exp.shape == (10000, 3)
    assert np.array_equal(exp.to_numpy(), gold["profile_experimental"])
    assert len(np.unique(exp["Reff [m]"])) == 540  # rounded to 3 decimals: duplicates, first one must win
    tst = TestExperimentalProfile("20230215.058", "8.0105").profiles_df
    assert np.array_equal(tst.to_numpy(), gold["profile_test_experimental"])


@pytest.mark.parametrize("el", list("BCNO"))
def test_fractional_abundance(gold, el):
    fa = FractionalAbundance(el, LINES[el][0]).df_interpolated_frac_ab
    assert list(fa.columns) == list(gold[f"fa_{el}_columns"])
    assert len(fa) == int(gold[f"fa_{el}_n"][0]) == 19995
    np.testing.assert_allclose(fa.to_numpy()[gold[f"fa_{el}_rows_idx"]], gold[f"fa_{el}_rows"], rtol=1e-15, atol=0)
    assert fa.iloc[0, 0] == 5 and fa.iloc[-1, 0] == 19999


@pytest.mark.parametrize("el", list("BCNO"))
def test_pec_nodes_and_dense_table(gold, el):
    ion, wl = LINES[el]
    for tr in ("EXCIT", "RECOM"):
        pec = PEC(el, wl, ion[-1], tr, 2000)
        assert np.array_equal(pec.ne_nodes, gold[f"pec_{el}_{tr}_nodes_ne"])
        assert np.array_equal(pec.te_nodes, gold[f"pec_{el}_{tr}_nodes_te"])
        assert np.array_equal(pec.pec_data_array, gold[f"pec_{el}_{tr}_table"])
        ii, jj = gold[f"pec_{el}_{tr}_ij"]
        dense = pec.interpolated_pec
        assert dense.shape == (2000, 2000, 3)
        ref = gold[f"pec_{el}_{tr}_vals"]
        assert np.array_equal(dense[ii, jj, :2], ref[:, :2])  # the linspace axes, exactly
        np.testing.assert_allclose(dense[ii, jj, 2], ref[:, 2], rtol=4e-15, atol=0)  # FITPACK vs plain bilinear


def test_pec_missing_block_raises():
    with pytest.raises(UnboundLocalError):
        PEC("C", 99.9, "5", "EXCIT")
