"""TEST INFRASTRUCTURE - CPU oracle for Stage 1 (visibility + geometric weight).

NumPy/SciPy restatement of ``co_monitor.geometry.simulation.Simulation`` with the
dask/opt_einsum plumbing removed (neither is installed); the
arithmetic, including the in-hull test through the *same* Qhull call
(``scipy.spatial.Delaunay(...).find_simplex``), follows the reference:

  reflect()            simulation.py:273-325   foot on the cylinder axis, unit normal, r = d - 2(d.n)n,
                                               and the reference's round-off (plasma = c - d, crystal = p + d)
  plane_hit()          simulation.py:143-173, 175-207   N = (p2-p1)x(p3-p2), |N.u| < 1e-6 -> NaN
  in_slab()            simulation.py:209-271   Delaunay(v +/- orientation_vector).find_simplex(X) >= 0
  transmission()       simulation.py:327-432   any_k(crys_k & plasma_k) & detector & port
  ray_weights()        simulation.py:483-536, 564-577   distance, AOI with round(deg, 2), weight
  voxel_table()        simulation.py:579-594   per-voxel sum, coordinates .round(1)

PARITY PIN: reproduces the four full-grid golden Stage-1 files shipped with the
reference (input_files/geometry/observed_plasma_volume/{B,C,N,O}/...dat;
(spacing=10, h=30, l=20, S=10)) - identical visible-voxel sets and weights to
<= 4e-15 relative; checked on a sample in tests/test_oracle_stage1.py (fast) and on
the full grid by ``python -m oracle.stage1 --full`` / the slow-marked test.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product never does.
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
from scipy.spatial import Delaunay

from . import geometry_setup as gs

SIGMA_REFLECT = 2.2  # width of the reflectivity profile, simulation.py:568


def build_scene(element, slits_number=10, h=20, l=80, db=None, side="top closing side"):
    """Everything Simulation.__init__ collects before the ray work (simulation.py:77-125)."""
    db = db or gs.load_db()
    cr = gs.crystal(db, element, h, l)
    crys, plas, vfb = gs.slit_quads(db, element, slits_number, side)
    pv, pov = gs.port(db)
    dv, dov = gs.detector(db, element)
    return SimpleNamespace(
        element=element, S=slits_number, h=h, l=l, crystal=cr, crystal_points=cr.points,
        slit_crys=crys, slit_plasma=plas, vfb=vfb, port_v=pv, port_ov=pov, det_v=dv, det_ov=dov,
        area=round(20 * 80 / h / l, 2),  # simulation.py:137-139
        _hulls={},
    )


def crystal_normals(scene):
    """Unit normals of the cylinder at every crystal point (simulation.py:286-312)."""
    cr = scene.crystal
    a = cr.radius_central_point
    b = cr.radius_central_point + (cr.B - cr.C)
    ab = b - a
    foot = np.array([a + np.dot(p - a, ab) / np.dot(ab, ab) * ab for p in scene.crystal_points])
    n = scene.crystal_points - foot
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return n


def reflect(scene, voxels):
    """Return (plasma, crystal, reflected), each (n, C, 3), as the reference stacks them."""
    P = voxels.reshape(-1, 1, 3)
    C = scene.crystal_points.reshape(1, -1, 3)
    n = crystal_normals(scene).reshape(1, -1, 3)
    d = C - P
    r = d - 2 * (d * n).sum(axis=-1, keepdims=True) * n
    return C - d, P + d, C + r


def plane_hit(p1, p2, p3, src, dst, epsilon=1e-6):
    normal = np.cross(p2 - p1, p3 - p2)
    u = src - dst
    ndotu = np.tensordot(normal, u, axes=(0, 2))
    ndotu[np.abs(ndotu) < epsilon] = np.nan
    w = dst - p1
    x = np.tensordot(normal, w, axes=(0, 2))
    si = -(x / ndotu)[:, :, np.newaxis]
    return w + si * u + p1


def _hull(scene, key, verts, ov):
    if key not in scene._hulls:
        slab = np.concatenate((verts + ov, verts - ov)).reshape(-1, 3)
        scene._hulls[key] = Delaunay(slab)
    return scene._hulls[key]


def in_slab(scene, key, points, verts, ov):
    return _hull(scene, key, verts, ov).find_simplex(points) >= 0


def transmission(scene, plasma, crystal, reflected, parts=False):
    """(n, C) bool: the ray mask of simulation.py:408-428."""
    shape = plasma.shape[:2]
    through = np.zeros(shape, bool)
    for k in range(scene.S):
        q = scene.slit_crys[k]
        a = in_slab(scene, ("c", k), plane_hit(q[0], q[1], q[2], plasma, crystal), q, scene.vfb)
        q = scene.slit_plasma[k]
        b = in_slab(scene, ("p", k), plane_hit(q[0], q[1], q[2], plasma, crystal), q, scene.vfb)
        through |= a & b
    pv = scene.port_v
    port = in_slab(scene, "port", plane_hit(pv[0], pv[1], pv[2], plasma, crystal), pv, scene.port_ov)
    dv = scene.det_v
    det = in_slab(scene, "det", plane_hit(dv[0], dv[1], dv[2], reflected, crystal), dv, scene.det_ov)
    if parts:
        return through, port, det
    return through & det & port


def ray_weights(scene, plasma, crystal, reflected):
    """Per-ray distance, AOI and weight for ALL rays of the chunk (mask applied by the caller)."""
    v = plasma - crystal
    r = reflected - crystal
    dist = np.sqrt(np.sum(v**2, axis=-1))
    cosang = (v * r).sum(axis=-1) / (np.linalg.norm(v, axis=2) * np.linalg.norm(r, axis=2))
    aoi = (180 - np.degrees(np.arccos(cosang)).round(2)) / 2
    cr = scene.crystal
    fraction = scene.area / (4 * np.pi * dist**2)
    reflect_prof = cr.max_reflectivity * np.exp(-((aoi - cr.AOI) ** 2) / (2 * SIGMA_REFLECT**2))
    return dist, aoi, fraction * reflect_prof


def detector_image(scene, plasma, crystal, reflected, sel, w_ray, nx, ny, pixel_mm):
    """Extension oracle: 2-D histogram of the detector-plane hit points of the passing rays
    (simulation.py:387-390 gives the hit), in the frame origin = detector vertex 1, e1 along vertex 1 -> 2,
    e2 = n x e1, weight-accumulated; hits are clamped to the image."""
    dv = scene.det_v
    X = plane_hit(dv[0], dv[1], dv[2], reflected, crystal)[sel]
    normal = np.cross(dv[1] - dv[0], dv[2] - dv[1])
    nh = normal / np.linalg.norm(normal)
    e1 = (dv[1] - dv[0]) / np.linalg.norm(dv[1] - dv[0])
    e2 = np.cross(nh, e1)
    x, y = (X - dv[0]) @ e1, (X - dv[0]) @ e2
    px = np.clip(np.floor(x / pixel_mm).astype(int), 0, nx - 1)
    py = np.clip(np.floor(y / pixel_mm).astype(int), 0, ny - 1)
    img = np.zeros((ny, nx))
    np.add.at(img, (py, px), w_ray[sel])
    return img


def run_chunk(scene, voxels, want_rays=False):
    """Stage 1 on a chunk of voxels.  Returns per-voxel weight, per-voxel passing-ray count and,
    optionally, the (n, C) mask with per-ray distance / AOI."""
    plasma, crystal, reflected = reflect(scene, voxels)
    sel = transmission(scene, plasma, crystal, reflected)
    dist, aoi, w = ray_weights(scene, plasma, crystal, reflected)
    w_vox = np.where(sel, w, 0.0).sum(axis=1)
    out = SimpleNamespace(weight=w_vox, nrays=sel.sum(axis=1).astype(np.int64))
    if want_rays:
        out.mask, out.dist, out.aoi, out.w_ray = sel, dist, aoi, w
        out.geometry = (plasma, crystal, reflected)
    return out


# ------------------------------------------------------------------------------------------------
# whole-grid drivers (multiprocessing over 1000-voxel chunks: the reference's own parallelism is
# dask's threaded scheduler over array blocks, simulation.py:87,113,323)
_WORKER = {}


def _init_worker(element, S, h, l, spacing, dims):
    try:  # one BLAS thread per process, otherwise tensordot oversubscribes the cores
        from threadpoolctl import threadpool_limits

        _WORKER["blas_limit"] = threadpool_limits(1)
    except ImportError:  # pragma: no cover
        os.environ["OMP_NUM_THREADS"] = "1"
    _WORKER["scene"] = build_scene(element, S, h, l)
    _WORKER["db"] = gs.load_db()
    _WORKER["spacing"] = spacing
    _WORKER["dims"] = dims


def _work(index_range):
    lo, hi = index_range
    idx = np.arange(lo, hi, dtype=np.int64)
    vox = gs.voxel_mesh(_WORKER["db"], _WORKER["spacing"], _WORKER["dims"], idx)
    res = run_chunk(_WORKER["scene"], vox)
    return lo, res.weight, res.nrays


def run_range(element, lo, hi, S=10, spacing=10, h=30, l=20, dims=(1350, 800, 250),
              processes=None, chunk=1000):
    """Stage 1 for the flat voxel indices [lo, hi) of the lattice; returns (weight, nrays, processes)."""
    import multiprocessing as mp

    processes = processes or os.cpu_count() or 1
    jobs = [(a, min(a + chunk, hi)) for a in range(lo, hi, chunk)]
    weight = np.zeros(hi - lo)
    nrays = np.zeros(hi - lo, np.int64)
    args = (element, S, h, l, spacing, dims)
    if processes == 1:
        _init_worker(*args)
        results = map(_work, jobs)
    else:
        ctx = mp.get_context("fork")
        pool = ctx.Pool(processes, initializer=_init_worker, initargs=args)
        results = pool.imap_unordered(_work, jobs)
    for a, w, n in results:
        weight[a - lo:a - lo + len(w)] = w
        nrays[a - lo:a - lo + len(n)] = n
    if processes != 1:
        pool.close()
        pool.join()
    return weight, nrays, processes


def voxel_table(element, weight, nrays, lo=0, spacing=10, dims=(1350, 800, 250)):
    """Rows of the Stage-1 file for the visible voxels of a run_range() result."""
    vis = np.flatnonzero(nrays > 0) + lo
    pts = gs.voxel_mesh(gs.load_db(), spacing, dims, vis).round(1)
    return vis, pts, weight[vis - lo]


def _main():
    import argparse
    import time

    ap = argparse.ArgumentParser(description="Stage-1 oracle vs the shipped golden files")
    ap.add_argument("--elements", default="C")
    ap.add_argument("--full", action="store_true", help="whole 270 000-voxel grid (about 90 s per element on 8 cores)")
    ap.add_argument("--golden-root", default=os.path.join(os.path.dirname(gs.DEFAULT_DB), "observed_plasma_volume"))
    a = ap.parse_args()
    import pandas as pd

    for el in a.elements:
        hi = 270000 if a.full else 60000
        t0 = time.time()
        w, n, procs = run_range(el, 0, hi)
        dt = time.time() - t0
        gold = pd.read_csv(os.path.join(a.golden_root, el, f"{el}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat"), sep=";")
        gold = gold[gold.idx_sel_plas_points < hi].sort_values("idx_sel_plas_points")
        vis, pts, wv = voxel_table(el, w, n)
        same = np.array_equal(vis, gold.idx_sel_plas_points.to_numpy())
        rel = np.max(np.abs(wv / gold.total_intensity_fraction.to_numpy() - 1)) if same else float("nan")
        coords = same and np.array_equal(pts, gold[["plasma_x", "plasma_y", "plasma_z"]].to_numpy())
        print(f"{el}: voxels [0,{hi}) visible {len(vis)} (golden {len(gold)}) same-set={same} coords-equal={coords} "
              f"max-rel-weight-err={rel:.2e} passing-rays={n.sum()} sum-w={w.sum():.11g} "
              f"{hi*600/dt:.3g} tests/s on {procs} processes")


if __name__ == "__main__":
    _main()
