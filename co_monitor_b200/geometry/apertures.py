"""Apertures as "plane + convex polygon" (host side, fp64).

The reference decides whether a ray clears an aperture with
``Delaunay(vertices +/- orientation_vector).find_simplex(X) >= 0`` where ``X`` is the
ray's hit point on the aperture plane (co_monitor/geometry/simulation.py:209-271).
``X`` lies in the plane, so the predicate is "X inside the cross-section of the slab's
convex hull with the plane" - a convex polygon, computed exactly by the C library
(``com_aperture_from_slab``).

One subtlety is reproduced here rather than in the kernels.  SciPy's ``find_simplex``
is not an exact hull test: where Qhull covers a (planar) hull face with a *flat*
simplex, the neighbouring simplex accepts points whose barycentric coordinate towards
that face is as low as ``-sqrt(100 * eps) = -1.49e-7``, i.e. points up to
``1.49e-7 x (simplex altitude)`` *outside* the hull (4e-6 mm on the detector, 1.5e-5 mm
on a slit, up to 1e-4 mm on the port).  Which faces get the leeway depends on Qhull's
triangulation, so ``calibrate_to_qhull`` measures it: for every polygon edge it bisects
along the edge normal for the point where the reference's own predicate flips and moves
the edge there.  With the calibrated polygons the fp64 polygon test agrees with
``find_simplex`` to the resolution of fp64 coordinates (~1e-12 mm); without them the
two agree outside a 1e-4 mm band around the edges.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib


def aperture_polygon(vertices: np.ndarray, orientation_vector: np.ndarray) -> _lib.ComAperture:
    """Exact section of ``conv(vertices +/- orientation_vector)`` with the plane through vertices[0:3]."""
    return _lib.aperture_from_slab(vertices, np.asarray(orientation_vector, dtype=float).reshape(3))


def calibrate_to_qhull(ap: _lib.ComAperture, vertices: np.ndarray, orientation_vector: np.ndarray,
                       reach: float = 2e-3, iterations: int = 48) -> np.ndarray:
    """Shift the polygon edges of *ap* (in place) to where ``find_simplex`` actually flips.

    Returns the outward shift applied to every edge in mm (0 for edges Qhull treats exactly).
    """
    from scipy.spatial import Delaunay

    ov = np.asarray(orientation_vector, dtype=float).reshape(3)
    slab = np.concatenate((vertices + ov, vertices - ov)).reshape(-1, 3)
    hull = Delaunay(slab)
    p1, e1, e2 = (np.array(list(v)) for v in (ap.p1, ap.e1, ap.e2))
    V, E = ap.vertices(), ap.edges()
    n = ap.n_edges
    # probe at three points per edge; the leeway is bounded by a line parallel to the edge
    ts = np.array([0.25, 0.5, 0.75])
    mids = (V[:, None, :] * (1 - ts)[None, :, None] + np.roll(V, -1, axis=0)[:, None, :] * ts[None, :, None]).reshape(-1, 2)
    normals = np.repeat(E[:, :2], len(ts), axis=0)  # inward
    lo = np.full(len(mids), -reach)  # outside
    hi = np.full(len(mids), +reach)  # inside

    def inside(s):
        xy = mids + s[:, None] * normals
        X = p1 + xy[:, :1] * e1 + xy[:, 1:] * e2
        return hull.find_simplex(X) >= 0

    ok = inside(hi) & ~inside(lo)
    for _ in range(iterations):
        mid = 0.5 * (lo + hi)
        m = inside(mid)
        hi = np.where(m, mid, hi)
        lo = np.where(m, lo, mid)
    shift = -hi.reshape(n, len(ts))  # > 0: the reference accepts points beyond the exact edge
    ok = ok.reshape(n, len(ts))
    out = np.zeros(n)
    for k in range(n):
        if not ok[k].all():
            continue  # edge shorter than the probe reach or polygon thinner than 2*reach: leave exact
        s = float(np.median(shift[k]))
        if np.ptp(shift[k]) > 1e-9 or abs(s) < 1e-11:
            continue  # not a parallel offset / nothing to correct
        ap.edge[k][2] += s
        out[k] = s
    return out


_APERTURE_CACHE: dict = {}


def build_apertures(slit_crys, slit_plasma, vector_front_back, port_vertices, port_ov, det_vertices, det_ov,
                    calibrate: bool = True):
    """ctypes array of the 2S+2 apertures in the library's order plus the per-edge calibration shifts.

    The calibration against SciPy costs about a second per channel; the result only depends on the vertex arrays, so it
    is kept per process (a second ``Simulation`` of the same channel, or the four shards of a sweep, reuse it)."""
    S = len(slit_crys)
    key = (bool(calibrate),) + tuple(np.ascontiguousarray(a, dtype=float).tobytes() for a in
                                     (slit_crys, slit_plasma, vector_front_back, port_vertices, port_ov, det_vertices, det_ov))
    hit = _APERTURE_CACHE.get(key)
    arr = (_lib.ComAperture * (2 * S + 2))()
    if hit is not None:
        C.memmove(arr, hit[0], len(hit[0]))
        return arr, [s.copy() for s in hit[1]]
    shifts = []
    specs = []
    for k in range(S):
        specs.append((slit_crys[k], vector_front_back))
        specs.append((slit_plasma[k], vector_front_back))
    specs.append((port_vertices, port_ov))
    specs.append((det_vertices, det_ov))
    for i, (verts, ov) in enumerate(specs):
        verts = np.ascontiguousarray(verts, dtype=float)
        ap = aperture_polygon(verts, ov)
        shifts.append(calibrate_to_qhull(ap, verts, ov) if calibrate else np.zeros(ap.n_edges))
        C.memmove(C.byref(arr[i]), C.byref(ap), C.sizeof(ap))
    if len(_APERTURE_CACHE) < 64:
        _APERTURE_CACHE[key] = (bytes(arr), [np.array(s, dtype=float) for s in shifts])
    return arr, shifts
