"""The Stage-1 oracle pinned to the reference: setup restatement against fixtures from the unmodified classes,
and the ray pipeline against the golden Stage-1 file shipped with the reference (sampled here; the full-grid
check of all four channels is `python -m oracle.stage1 --elements BCNO --full`, results in DESIGN.md)."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import geometry_setup as gs
from oracle import stage1 as o1


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "setup_reference.npz"))


def test_setup_restatement_bit_exact(gold):
    db = gs.load_db()
    for el in "BCNO":
        for h, l in ((30, 20), (20, 80), (10, 10)):
            assert np.array_equal(gs.crystal(db, el, h, l).points, gold[f"crystal_{el}_{h}x{l}"])
        crys, plas, vfb = gs.slit_quads(db, el, 10)
        assert np.array_equal(crys, gold[f"collim_{el}_top_crys"]) and np.array_equal(plas, gold[f"collim_{el}_top_plasma"])
        dv, dov = gs.detector(db, el)
        assert np.array_equal(dv, gold[f"detector_{el}_vertices"]) and np.array_equal(dov, gold[f"detector_{el}_ov"])
    pv, pov = gs.port(db)
    assert np.array_equal(pv, gold["port_vertices"]) and np.array_equal(pov, gold["port_ov"])
    assert np.array_equal(gs.voxel_mesh(db, 30), gold["mesh_30"])
    assert np.array_equal(gs.voxel_mesh(db, 10, flat_indices=gold["mesh_10_idx"]), gold["mesh_10_pts"])
    assert gs.mesh_shape(db, 10) == (135, 80, 25)


@pytest.mark.parametrize("el,lo,hi", [("C", 130000, 131500), ("O", 101000, 102000)])
def test_golden_file_sample(el, lo, hi):
    path = os.path.join(os.path.dirname(gs.DEFAULT_DB), "observed_plasma_volume", el,
                        f"{el}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat")
    gold = pd.read_csv(path, sep=";")
    gold = gold[(gold.idx_sel_plas_points >= lo) & (gold.idx_sel_plas_points < hi)].sort_values("idx_sel_plas_points")
    w, n, _ = o1.run_range(el, lo, hi, processes=2, chunk=500)
    vis, pts, wv = o1.voxel_table(el, w, n, lo)
    assert len(gold) > 50
    assert np.array_equal(vis, gold.idx_sel_plas_points.to_numpy())
    assert np.array_equal(pts, gold[["plasma_x", "plasma_y", "plasma_z"]].to_numpy())
    np.testing.assert_allclose(wv, gold.total_intensity_fraction.to_numpy(), rtol=1e-14, atol=0)


def test_parallel_ray_is_rejected_like_nan():
    """|N.u| < 1e-6 -> NaN -> never inside (simulation.py:167-168)."""
    p1, p2, p3 = np.array([0.0, 0, 0]), np.array([1.0, 0, 0]), np.array([1.0, 1, 0])
    src = np.array([[[5.0, 5.0, 1.0]]])
    dst = np.array([[[0.0, 0.0, 1.0]]])
    assert np.isnan(o1.plane_hit(p1, p2, p3, src, dst)).all()
