"""CPU-side check of the column-interval form of the lattice filter (visibility.cu: interval_kernel_lattice) without a
GPU: the tables come from the library's host-only entry points and the kernel's arithmetic is emulated in NumPy float32
column by column.  Against the fp64 oracle: every ray the oracle accepts lies in an emitted interval (the filter only
over-accepts), every ray the kernel would flag "certain" - its fp64 aperture tests are skipped - is accepted by the
oracle, and most accepted rays are certain."""
import contextlib
import io

import numpy as np
import pytest

from co_monitor_b200.geometry.engine import make_channel
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
from oracle import geometry_setup as gs
from oracle import stage1 as oracle

F = np.float32
BRIGHT_10MM = {"B": (10, 25), "C": (3, 26), "N": (9, 32), "O": (8, 25)}


def form(tab, p0):
    g = tab.astype(F)
    return (g[:, 0] * p0[0] + g[:, 1] * p0[1] + g[:, 2] * p0[2] + g[:, 3]).astype(F)


def dz_dot(tab, dz):
    g = tab.astype(F)
    return (g[:, 0] * dz[0] + g[:, 1] * dz[1] + g[:, 2] * dz[2]).astype(F)


def clip_halfline(lo, hi, c0, dc):
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (-c0 / dc).astype(F)
    lo = np.where(dc > 0, np.maximum(lo, t), lo)
    hi = np.where(dc < 0, np.minimum(hi, t), hi)
    dead = (dc == 0) & (c0 < 0)
    return np.where(dead, F(1), lo).astype(F), np.where(dead, F(0), hi).astype(F)


def clip_ratio(lo, hi, s, n0, dn, w0, dw, lo_b, hi_b):
    lo, hi = clip_halfline(lo, hi, s * (n0 - F(lo_b) * w0), s * (dn - F(lo_b) * dw))
    return clip_halfline(lo, hi, s * (F(hi_b) * w0 - n0), s * (F(hi_b) * dw - dn))


def ratio_inside(s, n, w, lo_b, hi_b):
    return (s * (n - F(lo_b) * w) >= 0) & (s * (F(hi_b) * w - n) >= 0)


def column_intervals(t, p0, dz, nz, n_slits, C):
    """Per crystal point: list of (a, b, ci_lo, ci_hi) voxel-index intervals of one column, as the kernel forms them."""
    det = t["det_forms"]
    xs0, ys0, w0 = (form(det[i], p0)[:C] for i in range(3))
    dxs, dys, dws = (dz_dot(det[i], dz)[:C] for i in range(3))
    s = np.where(w0 < 0, F(-1), F(1))
    lo, hi = np.zeros(C, F), np.full(C, nz - 1, F)
    lo, hi = clip_ratio(lo, hi, s, xs0, dxs, w0, dws, -1, 1)
    lo, hi = clip_ratio(lo, hi, s, ys0, dys, w0, dws, -1, 1)
    d_lo = np.maximum(np.ceil(lo - F(1e-3)).astype(int), 0)
    d_hi = np.minimum(np.floor(hi + F(1e-3)).astype(int), nz - 1)
    empty = ~(lo <= hi + F(2e-3))
    di = t["det_inner"]
    li, hi_i = np.full(C, -1e9, F), np.full(C, 1e9, F)
    if len(t["det_edges"]):  # the detector polygon edge by edge (scaled frame), inner offsets
        for ea, eb, _, ei in t["det_edges"]:
            li, hi_i = clip_halfline(li, hi_i, s * (F(ea) * xs0 + (F(eb) * ys0 + F(ei) * w0)).astype(F),
                                     s * (F(ea) * dxs + (F(eb) * dys + F(ei) * dws)).astype(F))
    else:
        li, hi_i = clip_ratio(li, hi_i, s, xs0, dxs, w0, dws, di[0], di[1])
        li, hi_i = clip_ratio(li, hi_i, s, ys0, dys, w0, dws, di[2], di[3])
    sf = t["slit_forms"]
    Xc0, Yc0, Wc0, Xp0, Yp0 = (form(sf[:, i], p0)[:C] for i in range(5))
    dXc, dYc, dWc, dXp, dYp = (dz_dot(sf[:, i], dz)[:C] for i in range(5))
    pf = t["port_forms"]
    Xq0, Yq0, Wq0 = (form(pf[:, i], p0)[:C] for i in range(3))
    dXq, dYq, dWq = (dz_dot(pf[:, i], dz)[:C] for i in range(3))
    sq = np.where(Wc0 < 0, F(-1), F(1))
    T = F(t["period"])
    rc, rp, ic, ip, pi = t["rect_crys"], t["rect_plasma"], t["inner_crys"], t["inner_plasma"], t["port_inner"]
    out = [[] for _ in range(C)]
    for c in range(C):
        if empty[c] or d_lo[c] > d_hi[c]:
            continue
        ka, kb = F(d_lo[c]), F(d_hi[c])
        xa = (Xc0[c] + ka * dXc[c]) / (Wc0[c] + ka * dWc[c])
        xb = (Xc0[c] + kb * dXc[c]) / (Wc0[c] + kb * dWc[c])
        j_lo = max(0, int(np.ceil((min(xa, xb) - rc[1]) / T - 1e-3)))
        j_hi = min(n_slits - 1, int(np.floor((max(xa, xb) - rc[0]) / T + 1e-3)))
        one = lambda v: np.array([v], F)  # noqa: E731
        for j in range(j_lo, j_hi + 1):
            off = F(j) * T
            lo_j, hi_j = one(ka), one(kb)
            l2, h2 = one(li[c]), one(hi_i[c])
            for side, (X0, Y0, dX, dY) in enumerate(((Xc0[c], Yc0[c], dXc[c], dYc[c]), (Xp0[c], Yp0[c], dXp[c], dYp[c]))):
                for ea, eb, eo, ei in t["slit_edges"][side]:
                    base0, based = F(ea * X0 + eb * Y0), F(ea * dX + eb * dY)
                    co, cin = F(eo - ea * off), F(ei - ea * off)
                    lo_j, hi_j = clip_halfline(lo_j, hi_j, one(sq[c] * (co * Wc0[c] + base0)), one(sq[c] * (co * dWc[c] + based)))
                    l2, h2 = clip_halfline(l2, h2, one(sq[c] * (cin * Wc0[c] + base0)), one(sq[c] * (cin * dWc[c] + based)))
            if not lo_j[0] <= hi_j[0] + F(2e-3):
                continue
            a = max(d_lo[c], int(np.ceil(lo_j[0] - F(1e-3))))
            b = min(d_hi[c], int(np.floor(hi_j[0] + F(1e-3))))
            if a > b:
                continue
            ci_lo, ci_hi = 1, 0
            if l2[0] <= h2[0] and not t["certain_off"]:
                ci_lo = max(a, int(np.ceil(max(l2[0], F(-1e6)) + F(1e-3))))
                ci_hi = min(b, int(np.floor(min(h2[0], F(1e6)) - F(1e-3))))
            if ci_lo <= ci_hi:
                wqa, wqb = Wq0[c] + F(ci_lo) * dWq[c], Wq0[c] + F(ci_hi) * dWq[c]
                ok = pi[0] <= pi[1] and wqa * wqb > 0 and min(abs(wqa), abs(wqb)) > 1e-3
                sp = F(-1) if wqa < 0 else F(1)
                in_rect = ok  # both ends inside the inner rectangle: nothing more to do (convexity)
                for k, wq in ((F(ci_lo), wqa), (F(ci_hi), wqb)):
                    in_rect = in_rect and bool(ratio_inside(sp, Xq0[c] + k * dXq[c], wq, pi[0], pi[1])) and \
                        bool(ratio_inside(sp, Yq0[c] + k * dYq[c], wq, pi[2], pi[3]))
                if ok and not in_rect and len(t["port_edges"]):  # the port polygon edge by edge, inner offsets
                    l3, h3 = one(-1e9), one(1e9)
                    for ea, eb, _, ei in t["port_edges"]:
                        l3, h3 = clip_halfline(l3, h3, one(sp * F(ea * Xq0[c] + F(eb * Yq0[c] + F(ei * Wq0[c])))),
                                               one(sp * F(ea * dXq[c] + F(eb * dYq[c] + F(ei * dWq[c])))))
                    if l3[0] <= h3[0]:
                        ci_lo = max(ci_lo, int(np.ceil(max(l3[0], F(-1e6)) + F(1e-3))))
                        ci_hi = min(ci_hi, int(np.floor(min(h3[0], F(1e6)) - F(1e-3))))
                    else:
                        ok = False
                else:
                    ok = ok and in_rect
                if not ok:
                    ci_lo, ci_hi = 1, 0
            out[c].append((a, b, ci_lo, ci_hi))
    return out


@pytest.mark.parametrize("el,spacing", [("C", 10), ("N", 10), ("O", 10), ("B", 10), ("C", 1), ("N", 1), ("O", 0.65), ("B", 0.65)])
def test_intervals_cover_the_oracle_and_certain_rays_pass(el, spacing):
    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(spacing)
        scene = make_channel(el, 10, 30, 20, lattice=mesh.lattice, host_only=True)
    lat = mesh.lattice
    t = scene.filter_tables()
    assert t["interval_filter"] and t["det_inner"][0] < t["det_inner"][1] and t["port_inner"][0] < t["port_inner"][1]
    assert len(t["port_edges"]) == 14  # the port polygon enters the certain test edge by edge
    assert 4 <= len(t["det_edges"]) <= 8  # ... and so does the detector polygon
    C = 600
    ix10, iy10 = BRIGHT_10MM[el]
    ix, iy = int(round(ix10 * (lat.nx - 1) / 134)), int(round(iy10 * (lat.ny - 1) / 79))
    osc = oracle.build_scene(el, 10, 30, 20)
    dz = (lat.step[2] * lat.rotation[2]).astype(F)
    n_cols = 12 if spacing == 10 else 2
    accepted = certain = emitted = 0
    for q in range(n_cols):
        col = iy * lat.nx + ix + (q if spacing == 10 else 3 * q)
        idx = np.arange(col * lat.nz, (col + 1) * lat.nz)
        ref = oracle.run_chunk(osc, gs.voxel_mesh(gs.load_db(), spacing, flat_indices=idx), want_rays=True)
        p0 = (mesh.points_at([idx[0]])[0] - t["origin"]).astype(F)
        ivs = column_intervals(t, p0, dz, lat.nz, 10, C)
        cand = np.zeros((lat.nz, C), bool)
        cert = np.zeros((lat.nz, C), bool)
        for c, lst in enumerate(ivs):
            for a, b, ci_lo, ci_hi in lst:
                assert not cand[a:b + 1, c].any(), "a ray is emitted twice"
                cand[a:b + 1, c] = True
                if ci_lo <= ci_hi:
                    assert a <= ci_lo and ci_hi <= b
                    cert[ci_lo:ci_hi + 1, c] = True
        assert not (ref.mask & ~cand).any(), "an oracle-accepted ray lies outside every emitted interval"
        assert not (cert & ~ref.mask).any(), "a ray flagged certain is rejected by the oracle"
        accepted += int(ref.mask.sum())
        certain += int(cert.sum())
        emitted += int(cand.sum())
    assert accepted > 50
    assert emitted <= 1.3 * accepted + 50  # selective: few rays beyond the accepted ones
    assert certain > (0.55 if spacing == 10 else 0.93) * accepted, (certain, accepted)
    print(f"{el} {spacing} mm: accepted {accepted}, emitted {emitted}, certain {certain}")
