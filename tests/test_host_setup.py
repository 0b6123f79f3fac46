"""Host geometry setup pinned bit-exactly to fixtures produced by the unmodified reference
(oracle/make_golden_setup.py) and to the reference's own known-answer tests
(tests/unit/geometry/test_rotation_matrix.py)."""
import os

import numpy as np
import pytest

from co_monitor_b200.geometry.collimator import Collimator
from co_monitor_b200.geometry.detector import Detector
from co_monitor_b200.geometry.dispersive_element import DispersiveElement
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
from co_monitor_b200.geometry.port import Port
from co_monitor_b200.geometry.rotation_matrix import rotation_matrix


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "setup_reference.npz"))


# --- the reference's own rotation-matrix tests, verbatim expectations ----------------
def test_rotation_matrix_180_degrees_around_y_axis():
    assert np.allclose(rotation_matrix(np.pi, [0, 1, 0]), [[-1, 0, 0], [0, 1, 0], [0, 0, -1]])


def test_rotation_matrix_45_degrees_around_z_axis():
    r = np.sqrt(2) / 2
    assert np.allclose(rotation_matrix(np.pi / 4, [0, 0, 1]), [[r, -r, 0], [r, r, 0], [0, 0, 1]])


def test_rotation_matrix_90_degrees_around_x_axis():
    assert np.allclose(rotation_matrix(np.pi / 2, [1, 0, 0]), [[1, 0, 0], [0, 0, -1], [0, 1, 0]])


def test_rotation_matrix_if_wrong_axis():
    with pytest.raises(ValueError):
        rotation_matrix(np.pi / 2, [1, 0])


def test_rotation_matrix_if_theta_not_nr():
    with pytest.raises(TypeError):
        rotation_matrix("2", [1, 0, 0])


def test_rotation_matrix_bit_exact(gold):
    assert np.array_equal(rotation_matrix(1, [0, 0, 1]), gold["rot_1rad_z"])
    assert np.array_equal(rotation_matrix(0.3, [0.2, -1.0, 0.5]), gold["rot_odd"])


# --- setup classes ------------------------------------------------------------------
@pytest.mark.parametrize("el", list("BCNO"))
def test_crystal_points_bit_exact(gold, el):
    for h, l in ((30, 20), (20, 80), (10, 10)):
        de = DispersiveElement(el, h, l)
        assert np.array_equal(de.crystal_points, gold[f"crystal_{el}_{h}x{l}"])
    aoi, rmax, radius, alpha = gold[f"crystal_scalars_{el}"]
    assert (float(de.AOI), de.max_reflectivity, de.radius, de.alpha) == (aoi, rmax, radius, alpha)
    assert np.array_equal(de.radius_central_point, gold[f"crystal_rc_{el}"])
    assert np.array_equal(de.B, gold[f"crystal_B_{el}"])
    assert np.array_equal(de.C, gold[f"crystal_C_{el}"])


@pytest.mark.parametrize("el", list("BCNO"))
def test_collimator_bit_exact(gold, el):
    for side in ("top closing side", "bottom closing side"):
        col = Collimator(el, side, 10)
        both, crys, plas = col.read_colim_coord()
        tag = side.split()[0]
        assert np.array_equal(crys, gold[f"collim_{el}_{tag}_crys"])
        assert np.array_equal(plas, gold[f"collim_{el}_{tag}_plasma"])
        assert np.array_equal(col.vector_front_back, gold[f"collim_{el}_{tag}_vfb"])
        assert both.shape == (10, 8, 3)
        assert col.spatial_colimator(crys[0]).shape == (8, 3)


def test_port_and_detector_bit_exact(gold):
    port = Port()
    assert np.array_equal(port.vertices_coordinates, gold["port_vertices"])
    assert np.array_equal(port.orientation_vector, gold["port_ov"])
    assert port.vertices_coordinates.shape == (14, 3)
    for el in "BCNO":
        det = Detector(el)
        assert np.array_equal(det.vertices, gold[f"detector_{el}_vertices"])
        assert np.array_equal(det.orientation_vector, gold[f"detector_{el}_ov"])
        assert det.orientation_vector.shape == (1, 3)


def test_mesh_bit_exact(gold, capsys):
    m30 = CuboidMesh(30)
    assert np.array_equal(m30.outer_cube_mesh, gold["mesh_30"])
    for s in (10, 15):
        mesh = CuboidMesh(s)
        assert mesh.lattice.n_points == int(gold[f"mesh_{s}_n"][0])
        assert np.array_equal(mesh.points_at(gold[f"mesh_{s}_idx"]), gold[f"mesh_{s}_pts"])
    assert "Number of points in the considered volume: 270000 points." in capsys.readouterr().out
    lat = CuboidMesh(10).lattice
    assert (lat.nx, lat.ny, lat.nz) == (135, 80, 25)
    # long index lists are evaluated in blocks on a thread pool: same values as in one piece, and as the reference
    # mesh rows they address
    fine = CuboidMesh(2)
    idx = np.sort(np.random.default_rng(0).choice(fine.lattice.n_points, 700_000, replace=False))
    assert np.array_equal(fine.points_at(idx), fine.points_at(idx, threads=1))
    m10 = CuboidMesh(10)
    every = np.arange(m10.lattice.n_points)
    assert len(every) < 2 * (1 << 18) or np.array_equal(m10.points_at(every), m10.points_at(every, threads=1))


def test_mesh_rejects_coarse_spacing():
    with pytest.raises(AssertionError):
        CuboidMesh(200)


def test_radiation_shield_vertices_bit_exact(gold):
    """The ECRH shields (plot-only in the reference, an optional stop here) are built like the reference builds them."""
    from co_monitor_b200.geometry.radiation_shield import RadiationShield, ecrh_shields

    for chamber in ("upper chamber", "bottom chamber"):
        for shield in ("1st shield", "2nd shield"):
            rs = RadiationShield(chamber, shield)
            ref = gold[f"shield_{chamber.split()[0]}_{shield[:3]}"]
            assert ref.shape == (720, 3) and np.array_equal(rs.vertices_coordinates, ref)
            st = rs.as_stop()
            # the stop's frame addresses the same rim: vertex k of the near rim at centre - axis/2 + r (cos u + sin v)
            k = np.arange(360)
            rim = st["centre"] - 0.5 * st["axis"] + st["radius"] * (np.cos(np.radians(k))[:, None] * st["u"] + np.sin(np.radians(k))[:, None] * st["v"])
            np.testing.assert_allclose(rim, ref[0::2], rtol=0, atol=1e-9)
    assert [round(float(s["centre"][2])) for s in ecrh_shields("C")] == [-80, -97]
    assert [round(float(s["centre"][2])) for s in ecrh_shields("N")] == [-526, -541]
