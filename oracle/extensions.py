"""TEST INFRASTRUCTURE - CPU statements of the optional physics extensions.

These switches are OFF on the parity path; the reference computes none of them, so nothing here is pinned by a
reference output ("parity unpinned", extensions only).  Each function states what the corresponding CUDA code does and
cites the reference lines the extension departs from.

  loglog_pec()         bilinear interpolation of log PEC in (log ne, log Te) through the ADAS nodes, instead of the
                       reference's linear interp2d table read at the nearest node (pec.py:153-187, reader.py:316-341)
                       -> csrc/emissivity.cu::pec_loglog
  shield_vertices()    the reference's model of one ECRH shield: a 1 mm long cylinder, 360 vertices per rim
                       (radiation_shield.py:52-101) - the reference only draws it
  shield_pass()        the shield as a stop: the ray's line against the mid-plane of that cylinder, inside the convex
                       hull the reference builds from the vertices (ConvexHull, radiation_shield.py:103-116)
                       -> csrc/visibility.cu::shield_hit
  run_chunk_physical() Stage 1 of a chunk of voxels with the shields as extra apertures and / or the segment-direction
                       check: the reference intersects infinite lines (simulation.py:143-173); with the check an aperture
                       counts only where the physical ray meets it - between voxel and crystal point for slits, port and
                       shields (line parameter 0 <= s <= 1), beyond the crystal point for the detector (s >= 0)
                       -> csrc/visibility.cu::exact_kernel, COM_PHYS_SEGMENT_CHECK

Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np


def loglog_pec(x_nodes, y_nodes, table, x, y) -> np.ndarray:
    """``pec_interpolation="loglog"``: bilinear in (log x, log y) of log table through the nodes, arguments clamped to
    the node range."""
    x_nodes, y_nodes, table = (np.asarray(a, dtype=float) for a in (x_nodes, y_nodes, table))
    x = np.clip(np.asarray(x, dtype=float), x_nodes[0], x_nodes[-1])
    y = np.clip(np.asarray(y, dtype=float), y_nodes[0], y_nodes[-1])
    i = np.clip(np.searchsorted(x_nodes, x, side="right") - 1, 0, len(x_nodes) - 2)
    j = np.clip(np.searchsorted(y_nodes, y, side="right") - 1, 0, len(y_nodes) - 2)
    u = (np.log(x) - np.log(x_nodes[i])) / (np.log(x_nodes[i + 1]) - np.log(x_nodes[i]))
    v = (np.log(y) - np.log(y_nodes[j])) / (np.log(y_nodes[j + 1]) - np.log(y_nodes[j]))
    c00, c01, c10, c11 = table[i, j], table[i, j + 1], table[i + 1, j], table[i + 1, j + 1]
    L = (1 - u) * (1 - v) * np.log(c00) + (1 - u) * v * np.log(c01) + u * (1 - v) * np.log(c10) + u * v * np.log(c11)
    return np.exp(L)


# ---- ECRH shields and the segment-direction check ---------------------------------------------------------------
def shield_vertices(db, chamber, shield, length=1, nlength=2, alpha=360, nalpha=360):
    """radiation_shield.py:52-101, operation by operation."""
    from .geometry_setup import rotation

    rec = db["ECRH shield"][chamber][shield]
    centre, radius, ovec_raw = np.array(rec["central point"]), rec["radius"], np.array(rec["orientation vector"])
    I = np.linspace(0, length, nlength)
    if int(alpha) == 360:
        A = np.linspace(0, alpha, num=nalpha, endpoint=False) / 180 * np.pi
    else:
        A = np.linspace(0, alpha, num=nalpha) / 180 * np.pi
    X, Y = radius * np.cos(A), radius * np.sin(A)
    points = np.vstack((np.tile(I, nalpha), np.repeat(X, nlength), np.repeat(Y, nlength))).T
    ovec = ovec_raw / np.linalg.norm(ovec_raw)
    cylvec = np.array([1, 0, 0])
    if np.allclose(cylvec, ovec):
        return points
    R = rotation(np.arccos(np.dot(ovec, cylvec)), np.cross(ovec, cylvec))
    return points.dot(R) + centre


def channel_shields(db, element):
    """Vertex arrays of the two shields of the chamber nearest the channel's crystal (C, O: upper; B, N: bottom)."""
    crystal = np.array(db["dispersive element"]["element"][element]["crystal central point"])
    best = min(db["ECRH shield"], key=lambda ch: min(
        np.linalg.norm(np.array(db["ECRH shield"][ch][sh]["central point"]) - crystal) for sh in db["ECRH shield"][ch]))
    return [shield_vertices(db, best, sh) for sh in ("1st shield", "2nd shield")]


def _line_parameter(normal, point, src, dst):
    """s of X = dst + s (src - dst) on the plane (point, normal); NaN for a line parallel to it."""
    u = src - dst
    ndotu = np.tensordot(normal, u, axes=(0, 2))
    ndotu = np.where(np.abs(ndotu) < 1e-12, np.nan, ndotu)
    return -np.tensordot(normal, dst - point, axes=(0, 2)) / ndotu


def shield_pass(verts, plasma, crystal, segment_check=False):
    """(n, C) bool: the line plasma -> crystal meets the mid-plane of the shield's cylinder inside its convex hull."""
    from scipy.spatial import ConvexHull

    axis = verts[1] - verts[0]  # the two axial stations of rim angle 0 (radiation_shield.py:77-80)
    axis = axis / np.linalg.norm(axis)
    mid = verts.mean(axis=0)
    s = _line_parameter(axis, mid, plasma, crystal)
    X = crystal + s[:, :, None] * (plasma - crystal)
    hull = ConvexHull(verts)
    inside = np.ones(s.shape, bool)
    for eq in hull.equations:
        inside &= X @ eq[:3] + eq[3] <= 0.0
    inside &= ~np.isnan(s)
    if segment_check:
        inside &= (s >= 0) & (s <= 1)
    return inside


def run_chunk_physical(scene, voxels, shields=(), segment_check=False):
    """oracle.stage1.run_chunk with the extensions: returns per-voxel weight and passing-ray count plus the number of
    rays the reference would accept but a shield stops / the segment check rejects."""
    from types import SimpleNamespace

    from . import stage1 as o1

    plasma, crystal, reflected = o1.reflect(scene, voxels)
    sel_ref = o1.transmission(scene, plasma, crystal, reflected)
    sel = sel_ref.copy()
    seg_rejected = 0
    if segment_check:
        ok = np.ones(sel.shape, bool)
        pv, dv = scene.port_v, scene.det_v
        planes = [(scene.slit_crys[0], plasma, False), (scene.slit_plasma[0], plasma, False), (pv, plasma, False),
                  (dv, reflected, True)]  # every slit lies in the plane of slit 0 of its side
        for q, src, half_line in planes:
            normal = np.cross(q[1] - q[0], q[2] - q[1])
            s = _line_parameter(normal, q[0], src, crystal)
            ok &= (s >= 0) if half_line else ((s >= 0) & (s <= 1))
        seg_rejected = int((sel & ~ok).sum())
        sel &= ok
    blocked = 0
    if len(shields):
        through = np.ones(sel.shape, bool)
        for verts in shields:
            through &= shield_pass(verts, plasma, crystal, segment_check)
        blocked = int((sel & ~through).sum())
        sel &= through
    dist, aoi, w = o1.ray_weights(scene, plasma, crystal, reflected)
    return SimpleNamespace(weight=np.where(sel, w, 0.0).sum(axis=1), nrays=sel.sum(axis=1).astype(np.int64), mask=sel,
                           reference_mask=sel_ref, shield_blocked=blocked, segment_rejected=seg_rejected)
