"""TEST INFRASTRUCTURE - fixture generator (runs only where /root/reference exists).

Imports the UNMODIFIED reference setup classes (co_monitor/geometry/*.py) and
stores what they produce as tests/golden/setup_reference.npz, so that the host
setup of co_monitor_b200 and the oracle's restatement can be pinned to it on
machines without the reference.

    python oracle/make_golden_setup.py
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _refshim import install_stubs  # noqa: E402

install_stubs()

from co_monitor.geometry.collimator import Collimator  # noqa: E402
from co_monitor.geometry.detector import Detector  # noqa: E402
from co_monitor.geometry.dispersive_element import DispersiveElement  # noqa: E402
from co_monitor.geometry.mesh_calculation import CuboidMesh  # noqa: E402
from co_monitor.geometry.port import Port  # noqa: E402
from co_monitor.geometry.radiation_shield import RadiationShield  # noqa: E402
from co_monitor.geometry.rotation_matrix import rotation_matrix  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "setup_reference.npz")


def main() -> None:
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for el in "BCNO":
            for h, l in ((30, 20), (20, 80), (10, 10)):
                de = DispersiveElement(el, h, l)
                out[f"crystal_{el}_{h}x{l}"] = de.crystal_points
            out[f"crystal_scalars_{el}"] = np.array([de.AOI, de.max_reflectivity, de.radius, de.alpha])
            out[f"crystal_rc_{el}"] = de.radius_central_point
            out[f"crystal_B_{el}"] = de.B
            out[f"crystal_C_{el}"] = de.C
            for side in ("top closing side", "bottom closing side"):
                col = Collimator(el, side, 10)
                both, crys, plas = col.read_colim_coord()
                tag = side.split()[0]
                out[f"collim_{el}_{tag}_crys"] = crys
                out[f"collim_{el}_{tag}_plasma"] = plas
                out[f"collim_{el}_{tag}_vfb"] = col.vector_front_back
            det = Detector(el)
            out[f"detector_{el}_vertices"] = det.vertices
            out[f"detector_{el}_ov"] = det.orientation_vector
        port = Port()
        out["port_vertices"] = port.vertices_coordinates
        out["port_ov"] = port.orientation_vector
        # meshes: 30 mm fully, 10/15 mm at a deterministic sample of flat indices
        m30 = CuboidMesh(30).outer_cube_mesh
        out["mesh_30"] = m30
        rng = np.random.default_rng(0)
        for s in (10, 15):
            m = CuboidMesh(s).outer_cube_mesh
            idx = np.sort(rng.choice(len(m), 4000, replace=False))
            idx[:3] = (0, 1, len(m) - 1)
            out[f"mesh_{s}_n"] = np.array([len(m)])
            out[f"mesh_{s}_idx"] = idx
            out[f"mesh_{s}_pts"] = m[idx]
        RadiationShield.make_shield = lambda self: None  # the PolyData member is plot-only (pyvista is stubbed)
        for chamber in ("upper chamber", "bottom chamber"):
            for shield in ("1st shield", "2nd shield"):
                rs = RadiationShield(chamber, shield)
                out[f"shield_{chamber.split()[0]}_{shield[:3]}"] = rs.vertices_coordinates
        out["rot_1rad_z"] = rotation_matrix(1, [0, 0, 1])
        out["rot_odd"] = rotation_matrix(0.3, [0.2, -1.0, 0.5])
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), f"{os.path.getsize(OUT)/1e3:.0f} kB", len(out), "arrays")


if __name__ == "__main__":
    main()
