import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long CPU test (oracle full-grid checks)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def _reff10_path():
    """The reference's own 10 mm Reff file is a missing blob; the deterministic trilinear up-sampling of the shipped
    15 mm grid (oracle/reff.py - also what the golden Stage-2 vectors were generated with) stands in for it.  Written
    once next to the shipped grids, where ``Emissivity`` looks for it."""
    import contextlib
    import io

    from co_monitor_b200.geometry.reff import reff_path
    from oracle import reff as oreff

    path = reff_path(10)
    if not path.exists():
        with contextlib.redirect_stdout(io.StringIO()):
            oreff.write_file(oreff.file_frame(10, 15), path)
    return path


@pytest.fixture(scope="session", autouse=True)
def _reff10_file():
    return _reff10_path()


@pytest.fixture(scope="session")
def reff10(_reff10_file):
    """Reff per voxel of the 10 mm lattice."""
    import pandas as pd

    return pd.read_csv(_reff10_file, sep=";").iloc[:, 4].to_numpy(dtype=float)
