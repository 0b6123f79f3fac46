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


@pytest.fixture(scope="session")
def reff10():
    """Reff per voxel of the 10 mm lattice.  The reference's own 10 mm file is a missing blob; the repo's
    deterministic trilinear up-sampling of the shipped 15 mm grid stands in for it (also used when the golden
    Stage-2 vectors were generated).  Written once next to the shipped grids."""
    import contextlib
    import io

    from co_monitor_b200.geometry.reff import reff_path, upsample_reff, write_reff_file

    path = reff_path(10)
    if not path.exists():
        with contextlib.redirect_stdout(io.StringIO()):
            write_reff_file(upsample_reff(15, 10), path)
    import pandas as pd

    return pd.read_csv(path, sep=";").iloc[:, 4].to_numpy(dtype=float)
