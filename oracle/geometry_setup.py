"""TEST INFRASTRUCTURE - CPU oracle, geometry inputs.

Functional restatement of the reference's setup classes, independent of the
product's host code (co_monitor_b200/geometry), so that the oracle can be run on
machines where /root/reference does not exist.  Pinned bit-exactly to fixtures
produced by the unmodified reference (tests/golden/setup_reference.npz, made by
oracle/make_golden_setup.py) in tests/test_oracle_stage1.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product never does.

Reference lines followed:
  rotation            co_monitor/geometry/rotation_matrix.py:3-33
  voxel_mesh          co_monitor/geometry/mesh_calculation.py:77-159
  slit_quads          co_monitor/geometry/collimator.py:46-68, 93-125
  crystal             co_monitor/geometry/dispersive_element.py:34-49, 64-173
  port / detector     co_monitor/geometry/port.py:27-56, detector.py:31-80
"""
from __future__ import annotations

import json
import os
from types import SimpleNamespace

import numpy as np

DEFAULT_DB = os.path.join(
    os.path.dirname(os.path.abspath(__file__)),
    "..", "co_monitor_b200", "input_files", "geometry", "coordinates.json",
)


def load_db(path: str | None = None) -> dict:
    with open(path or DEFAULT_DB) as fh:
        return json.load(fh)


def rotation(theta, axis):
    axis = np.asarray(axis) / np.sqrt(np.dot(axis, axis))
    a = np.cos(theta / 2.0)
    b, c, d = -axis * np.sin(theta / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    return np.array(
        [
            [aa + bb - cc - dd, 2 * (b * c + a * d), 2 * (b * d - a * c)],
            [2 * (b * c - a * d), aa + cc - bb - dd, 2 * (c * d + a * b)],
            [2 * (b * d + a * c), 2 * (c * d - a * b), aa + dd - bb - cc],
        ]
    )


def mesh_axes(db, spacing, dims=(1350, 800, 250)):
    c = np.array(db["plasma"]["central point"])
    axes = []
    for k in range(3):
        lo, hi = c[k] - dims[k] / 2, c[k] + dims[k] / 2
        axes.append(np.linspace(lo, hi, int(abs(lo - hi) // spacing), endpoint=True))
    return c, axes


def voxel_mesh(db, spacing, dims=(1350, 800, 250), flat_indices=None):
    """(P,3) fp64 voxel coordinates in the reference's order, or the rows *flat_indices* of it."""
    c, (x, y, z) = mesh_axes(db, spacing, dims)
    nx, nz = len(x), len(z)
    if flat_indices is None:
        gx, gy, gz = np.meshgrid(x, y, z)
        pts = np.vstack((gx.ravel(), gy.ravel(), gz.ravel())).T
    else:
        flat_indices = np.asarray(flat_indices, dtype=np.int64)
        iz = flat_indices % nz
        ix = (flat_indices // nz) % nx
        iy = flat_indices // (nz * nx)
        pts = np.vstack((x[ix], y[iy], z[iz])).T
    pts = pts - c
    R = rotation(1, np.cross(np.array([1, 0, 0]) / 1.0, np.array([0, 1, 0])))
    return pts.dot(R) + c + [50, 25, 0]


def mesh_shape(db, spacing, dims=(1350, 800, 250)):
    _, (x, y, z) = mesh_axes(db, spacing, dims)
    return len(x), len(y), len(z)


def slit_quads(db, element, n_slits, side="top closing side"):
    col = db["collimator"]["element"][element]
    vfb = np.array(col["vector front-back"])
    t = np.array(col[side]["vector_top_bottom"])
    a1 = np.array(col[side]["vertex"]["A1"])
    a2 = np.array(col[side]["vertex"]["A2"])
    crys = np.empty((n_slits, 4, 3))
    for k in range(n_slits):
        crys[k, 0] = a1 + (t * 2 * k)
        crys[k, 1] = (a1 + t) + (t * 2 * k)
        crys[k, 2] = a2 + (t * 2 * k)
        crys[k, 3] = (a2 + 0.001 + t) + (t * 2 * k)
    return crys, crys + vfb, vfb


def _angle_deg(centre, p1, p2):
    v1, v2 = p1 - centre, p2 - centre
    return np.degrees(np.arccos(np.dot(v1, v2) / (np.linalg.norm(v1) * np.linalg.norm(v2))))


def crystal(db, element, h, l):
    e = db["dispersive element"]["element"][element]
    rc = np.array(e["radius central point"])
    A, B, C = (np.array(e["vertex"][k]) for k in "ABC")
    radius = int(np.sqrt(np.sum((A - rc) ** 2, axis=0)).round(2))
    alpha = _angle_deg(rc, A, B)
    ov = B - C
    hh = np.linspace(0, 20, h)
    arc = np.linspace(0, alpha, num=l) / 180 * np.pi
    X, Y = radius * np.cos(arc), radius * np.sin(arc)
    pts = np.vstack((np.tile(hh, l), np.repeat(X, h), np.repeat(Y, h))).T
    ovec = ov / np.linalg.norm(ov)
    cyl = np.array([1, 0, 0])
    if not np.allclose(cyl, ovec):
        pts = pts.dot(rotation(np.arccos(np.dot(ovec, cyl)), np.cross(ovec, cyl)))
        shift = np.array(rc - ov / 2)
        pts += shift
        ang = _angle_deg(rc, C, pts[0])
        pts -= shift
        pts = pts.dot(rotation(np.deg2rad(ang), ov))
        pts += shift
    return SimpleNamespace(
        points=pts, AOI=float(e["AOI"]), max_reflectivity=e["max reflectivity"],
        radius_central_point=rc, B=B, C=C, radius=radius, alpha=alpha,
    )


def port(db):
    verts = np.vstack([db["port"]["vertex"][k] for k in db["port"]["vertex"]])
    return verts, np.array(db["port"]["orientation vector"])


def detector(db, element):
    d = db["detector"]["element"][element]
    verts = np.vstack([d["vertex"][k] for k in d["vertex"]])
    return verts, np.vstack([d["orientation vector"]])
