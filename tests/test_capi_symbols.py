"""The C-ABI library builds, loads and exports every symbol include/comonitor_sm100.h declares (no GPU needed);
the compute entry points fail loudly - never fall back - when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from co_monitor_b200.csrc.build import build_library

    build_library()
    from co_monitor_b200 import _lib

    return _lib.load()


def declared_functions():
    text = open(os.path.join(ROOT, "include", "comonitor_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(com_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from co_monitor_b200 import _lib

    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "binding list and header disagree"
    declared = int(re.search(r"#define COM_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "comonitor_sm100.h")).read()).group(1))
    assert lib.com_abi_version() == declared == 4


def test_struct_layouts_match_the_header(lib):
    """sizeof() of the ctypes mirrors against the sizes the C compiler sees (via a tiny probe compiled with gcc)."""
    import subprocess
    import tempfile

    from co_monitor_b200 import _lib

    src = '#include <stdio.h>\n#include "comonitor_sm100.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          "sizeof(ComAperture),sizeof(ComSceneDesc),sizeof(ComLattice),sizeof(ComStats),sizeof(ComRayRecord)," \
          "sizeof(ComPecDesc),sizeof(ComTablesDesc),sizeof(ComShield));return 0;}\n"
    with tempfile.TemporaryDirectory() as tmp:
        c = os.path.join(tmp, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(tmp, "probe")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    mine = [C.sizeof(_lib.ComAperture), C.sizeof(_lib.ComSceneDesc), C.sizeof(_lib.ComLattice), C.sizeof(_lib.ComStats),
            _lib.RAY_RECORD_DTYPE.itemsize, C.sizeof(_lib.ComPecDesc), C.sizeof(_lib.ComTablesDesc), C.sizeof(_lib.ComShield)]
    assert mine == sizes


def test_error_strings(lib):
    assert lib.com_error_string(0) == b"ok"
    assert b"Not enough points" in lib.com_error_string(-6)


def test_host_only_aperture_section_matches_qhull(lib):
    """com_aperture_from_slab (pure host code) against scipy's Delaunay.find_simplex, the reference's predicate."""
    from scipy.spatial import Delaunay

    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.apertures import build_apertures
    from co_monitor_b200.geometry.collimator import Collimator
    from co_monitor_b200.geometry.detector import Detector
    from co_monitor_b200.geometry.port import Port

    rng = np.random.default_rng(3)
    for el in "CO":
        col = Collimator(el, "top closing side", 10)
        _, crys, plas = col.read_colim_coord()
        port, det = Port(), Detector(el)
        vfb = col.vector_front_back
        arr, shifts = build_apertures(crys, plas, vfb, port.vertices_coordinates, port.orientation_vector,
                                      det.vertices, det.orientation_vector.reshape(3), calibrate=True)
        assert [arr[i].n_edges for i in (0, 1, 20, 21)] == [4, 4, 14, 4]
        specs = {0: (crys[0], vfb), 7: (plas[3], vfb), 20: (port.vertices_coordinates, port.orientation_vector),
                 21: (det.vertices, det.orientation_vector.reshape(3))}
        for i, (verts, ov) in specs.items():
            ap = arr[i]
            hull = Delaunay(np.concatenate((verts + ov, verts - ov)).reshape(-1, 3))
            V, E = ap.vertices(), ap.edges()
            lo, hi = V.min(0), V.max(0)
            xy = lo - 0.1 * (hi - lo) + rng.random((20000, 2)) * 1.2 * (hi - lo)
            k, t = rng.integers(0, len(V), 20000), rng.random(20000)
            near = V[k] * (1 - t[:, None]) + V[(k + 1) % len(V)] * t[:, None] + rng.normal(0, 1e-3, (20000, 2))
            xy = np.vstack((xy, near))
            p1, e1, e2 = (np.array(list(v)) for v in (ap.p1, ap.e1, ap.e2))
            X = p1 + xy[:, :1] * e1 + xy[:, 1:] * e2
            ref = hull.find_simplex(X) >= 0
            margin = (xy @ E[:, :2].T + E[:, 2]).min(1)
            bad = ref != (margin >= 0)
            # the only allowed disagreements sit inside Qhull's own leeway band (<= 1.5e-7 x simplex altitude)
            assert np.abs(margin[bad]).max(initial=0.0) < 2e-5, (el, i, np.abs(margin[bad]).max())
            one = C.c_int(lib.com_aperture_contains(C.byref(ap), _lib.as_double_ptr(np.ascontiguousarray(X[0]))))
            assert bool(one.value) == bool(margin[0] >= 0)
        # slits are calibrated to Qhull's leeway: shifts are either ~0 or ~1.49e-7 x 100 mm
        s = np.concatenate(shifts[:20])
        assert np.all((np.abs(s) < 1e-10) | (np.abs(s - 1.49e-5) < 2e-6) | (np.abs(s - 1.49e-7) < 2e-8))


def test_degenerate_aperture_is_rejected(lib):
    from co_monitor_b200 import _lib

    verts = np.array([[0.0, 0, 0], [1, 0, 0], [2, 0, 0], [3, 0, 0]])
    with pytest.raises(_lib.ComonitorError) as exc:
        _lib.aperture_from_slab(verts, np.array([0.0, 0, 1]))
    assert exc.value.code == _lib.COM_E_GEOMETRY


def test_compute_fails_loudly_without_a_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.engine import make_channel

    with pytest.raises(_lib.ComonitorError) as exc:
        make_channel("C", 10, 5, 5)
    assert exc.value.code == _lib.COM_E_CUDA
    assert lib.com_measure_fp32_tflops() < 0


def test_driver_build_entry_point():
    """__graft_entry__.build() - what the driver runs on CPU each round - compiles (or finds) the library and passes its
    own checks."""
    import __graft_entry__ as entry

    entry.build(force=False)  # the driver's own call compiles from source (force defaults to True without a GPU)
