"""The reference's import paths resolve to the B200 implementation, and the reference's own unit tests
(tests/unit/emissivity/test_emissivity.py, tests/unit/geometry/test_rotation_matrix.py) pass against them verbatim."""
import numpy as np
import pytest


def test_reference_import_paths():
    from co_monitor.emissivity.kinetic_profiles import ExperimentalProfile, TestExperimentalProfile, TwoGaussSumProfile  # noqa: F401
    from co_monitor.emissivity.pec import PEC  # noqa: F401
    from co_monitor.emissivity.reader import Emissivity
    from co_monitor.geometry.collimator import Collimator  # noqa: F401
    from co_monitor.geometry.detector import Detector  # noqa: F401
    from co_monitor.geometry.dispersive_element import DispersiveElement  # noqa: F401
    from co_monitor.geometry.mesh_calculation import CuboidMesh  # noqa: F401
    from co_monitor.geometry.port import Port  # noqa: F401
    from co_monitor.geometry.simulation import Simulation

    import co_monitor_b200.emissivity.reader as impl

    assert Emissivity is impl.Emissivity
    assert callable(Simulation)


# --- tests/unit/emissivity/test_emissivity.py of the reference, expectations verbatim --------------------
list_of_nrs = [1, 22, 50, 16, 7, 19, 100, 4, 6, 211]


def test_find_nearest():
    from co_monitor.emissivity.reader import Emissivity

    assert Emissivity._find_nearest(list_of_nrs, 75) == 2
    assert Emissivity._find_nearest(list_of_nrs, 76) == 6
    assert Emissivity._find_nearest(list_of_nrs, 0) == 0


# --- tests/unit/geometry/test_rotation_matrix.py of the reference ------------------------------------------
def test_rotation_matrix_known_answers_and_errors():
    from co_monitor.geometry.rotation_matrix import rotation_matrix

    assert np.allclose(rotation_matrix(np.pi, [0, 1, 0]), [[-1, 0, 0], [0, 1, 0], [0, 0, -1]])
    r = np.sqrt(2) / 2
    assert np.allclose(rotation_matrix(np.pi / 4, [0, 0, 1]), [[r, -r, 0], [r, r, 0], [0, 0, 1]])
    assert np.allclose(rotation_matrix(np.pi / 2, [1, 0, 0]), [[1, 0, 0], [0, 0, -1], [0, 1, 0]])
    with pytest.raises(ValueError):
        rotation_matrix(np.pi / 2, [1, 0])
    with pytest.raises(TypeError):
        rotation_matrix("2", [1, 0, 0])
