"""CPU-side check of the FP32 filter (K1a) without a GPU: the filter tables come from the library's host-only entry
point, stages A and B are emulated in NumPy float32 - direct evaluation of the linear forms and forward differences
along z as the lattice kernel advances them - and every ray the fp64 oracle accepts must survive both stages."""
import contextlib
import io

import numpy as np
import pytest

from co_monitor_b200.geometry.engine import make_channel
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
from oracle import geometry_setup as gs
from oracle import stage1 as oracle


def forms(tab, p):
    """value[c, v] = g[c].p[v] + h[c] in float32 (tab: (C,4), p: (V,3))."""
    g, h = tab[:, :3].astype(np.float32), tab[:, 3].astype(np.float32)
    return (g[:, None, 0] * p[None, :, 0] + g[:, None, 1] * p[None, :, 1] + g[:, None, 2] * p[None, :, 2] + h[:, None]).astype(np.float32)


def stage_a(t, p):
    xs, ys, w = (forms(t["det_forms"][i], p) for i in range(3))
    return np.abs(w) - np.maximum(np.abs(xs), np.abs(ys)) >= 0


def certain(t, p, n_slits):
    """The filter's "certain pass" flag (float32, as stage B of the lattice kernel computes it): inside the inner
    rectangle of exactly one slit on both sides."""
    s = t["slit_forms"]
    Xc, Yc, Wc, Xp, Yp = (forms(s[:, i], p) for i in range(5))
    Wp = Wc if t["same_w"] else forms(s[:, 5], p)
    with np.errstate(divide="ignore", invalid="ignore"):
        xc, yc, xp, yp = Xc / Wc, Yc / Wc, Xp / Wp, Yp / Wp
    rc, rp, ic, ip = t["rect_crys"], t["rect_plasma"], t["inner_crys"], t["inner_plasma"]
    T = np.float32(t["period"])
    kf = np.floor(xc / T)
    hits = np.zeros(xc.shape, np.int32)
    inner = np.zeros(xc.shape, bool)
    for dk in (0, 1):
        k = kf + dk
        a, b = xc - k * T, xp - k * T
        ok = (k >= 0) & (k < n_slits) & (a >= rc[0]) & (a <= rc[1]) & (b >= rp[0]) & (b <= rp[1])
        inner |= ok & (hits == 0) & (a >= ic[0]) & (a <= ic[1]) & (b >= ip[0]) & (b <= ip[1])
        hits += ok
    inner &= (hits == 1) & (yc >= ic[2]) & (yc <= ic[3]) & (yp >= ip[2]) & (yp <= ip[3])
    return inner & ~(np.minimum(np.abs(Wc), np.abs(Wp)) < 1e-3)


def certain_everywhere(t, p):
    """Float32 restatement of visibility.cu::certain_everywhere (what the per-ray kernel adds, in its flush, for rays that are
    certain on the collimator): inside every detector edge (scaled frame of the detector forms) and every port edge, inner
    offsets.  Such rays skip every fp64 aperture test: only their weight is computed."""
    F = np.float32
    xs, ys, wd = (forms(t["det_forms"][i], p) for i in range(3))
    sd = np.where(wd < 0, F(-1), F(1))
    ok = np.abs(wd) > F(1e-3)
    for ea, eb, _, ei in t["det_edges"]:
        ok &= sd * (F(ea) * xs + (F(eb) * ys + F(ei) * wd)).astype(F) >= 0
    pf = t["port_forms"]
    Xq, Yq, Wq = (forms(pf[:, i], p) for i in range(3))
    sp = np.where(Wq < 0, F(-1), F(1))
    ok &= np.abs(Wq) - F(1e-3) >= 0
    for ea, eb, _, ei in t["port_edges"]:
        ok &= sp * (F(ea) * Xq + (F(eb) * Yq + F(ei) * Wq)).astype(F) >= 0
    return ok


def stage_b(t, p, n_slits):
    s = t["slit_forms"]
    Xc, Yc, Wc, Xp, Yp = (forms(s[:, i], p) for i in range(5))
    Wp = Wc if t["same_w"] else forms(s[:, 5], p)
    with np.errstate(divide="ignore", invalid="ignore"):
        xc, yc, xp, yp = Xc / Wc, Yc / Wc, Xp / Wp, Yp / Wp
    rc, rp = t["rect_crys"], t["rect_plasma"]
    y_ok = (yc >= rc[2]) & (yc <= rc[3]) & (yp >= rp[2]) & (yp <= rp[3])
    kf = np.floor(xc / np.float32(t["period"]))
    ok = np.zeros(xc.shape, bool)
    for dk in (0, 1):
        k = kf + dk
        a, b = xc - k * np.float32(t["period"]), xp - k * np.float32(t["period"])
        ok |= (k >= 0) & (k < n_slits) & (a >= rc[0]) & (a <= rc[1]) & (b >= rp[0]) & (b <= rp[1])
    grazing = np.minimum(np.abs(Wc), np.abs(Wp)) < 1e-3
    return (ok & y_ok) | grazing


@pytest.mark.parametrize("el,spacing,h,l,lo", [("C", 10, 30, 20, 131000), ("N", 10, 30, 20, 140000), ("O", 15, 10, 10, 40000)])
def test_filter_never_rejects_a_ray_the_oracle_accepts(el, spacing, h, l, lo):
    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(spacing)
        scene = make_channel(el, 10, h, l, lattice=mesh.lattice, host_only=True)
    t = scene.filter_tables()
    assert t["slit_filter"] and not t["disabled"]
    C = h * l
    idx = np.arange(lo, lo + 400)
    vox = mesh.points_at(idx)
    osc = oracle.build_scene(el, 10, h, l)
    ref = oracle.run_chunk(osc, gs.voxel_mesh(gs.load_db(), spacing, flat_indices=idx), want_rays=True)
    assert ref.mask.sum() > 20
    p = (vox - t["origin"]).astype(np.float32)
    a = stage_a(t, p)[:C].T  # (V, C)
    b = stage_b(t, p, 10)[:C].T
    assert not (ref.mask & ~a).any(), "stage A rejected a ray the oracle accepts"
    assert not (ref.mask & ~b).any(), "stage B rejected a ray the oracle accepts"
    # "certain" candidates skip the fp64 slit tests of the exact kernel: the oracle must find every one of them going
    # through the collimator (port and detector are still tested), and most candidates should be certain
    cert = certain(t, p, 10)[:C].T & a & b
    through, port, det = oracle.transmission(osc, *ref.geometry, parts=True)
    assert not (cert & ~through).any(), "a ray flagged certain fails the collimator in the oracle"
    assert cert.sum() > 0.5 * (a & b).sum()
    # ... and those that are also inside the detector's and the port's inner edges skip EVERY fp64 aperture test (they go to
    # the weight-only kernel): the oracle must accept each of them outright, and they should be the bulk of the accepted rays
    assert len(t["det_edges"]) >= 4 and len(t["port_edges"]) == 14
    cert_all = cert & certain_everywhere(t, p)[:C].T
    assert not (cert_all & ~ref.mask).any(), "a ray flagged certain-everywhere is rejected by the oracle"
    assert cert_all.sum() > 0.5 * ref.mask.sum(), (cert_all.sum(), ref.mask.sum())
    # the filter is also selective: few rays beyond the accepted ones survive both stages
    survivors = (a & b).sum()
    assert ref.mask.sum() <= survivors <= 1.3 * ref.mask.sum() + 50
    # forward differences along z, as the lattice kernel advances the forms (re-based every <= 64 voxels)
    nz = mesh.lattice.nz
    col0 = (lo // nz + 1) * nz  # first full column inside the slab
    k = np.arange(min(nz, 64))
    base = (mesh.points_at([col0]) - t["origin"]).astype(np.float32)
    dz = (mesh.lattice.step[2] * mesh.lattice.rotation[2]).astype(np.float32)
    margins = []
    for i in range(3):
        f0 = forms(t["det_forms"][i], base)[:, 0]
        d = (t["det_forms"][i][:, :3].astype(np.float32) @ dz).astype(np.float32)
        acc = np.empty((len(f0), len(k)), np.float32)
        cur = f0.copy()
        for j in range(len(k)):
            acc[:, j] = cur
            cur = (cur + d).astype(np.float32)
        margins.append(acc)
    fd = (np.abs(margins[2]) - np.maximum(np.abs(margins[0]), np.abs(margins[1])) >= 0)[:C].T  # (k, C)
    sel = ref.mask[col0 - lo: col0 - lo + len(k)]
    assert not (sel & ~fd).any(), "forward-differenced stage A rejected a ray the oracle accepts"
    # hierarchical cull of the lattice kernel: a run (whole segment, then groups of 8 voxels) is skipped for a crystal
    # point when |Xs| > |W| at both ends with Xs of one sign (or the same for Ys); end values by fmaf like the kernel
    X, Y, W = margins

    def culled(a, b):
        xa, ya, wa = X[:, a], Y[:, a], W[:, a]
        n = np.float32(b - a)
        d = [(t["det_forms"][i][:, :3].astype(np.float32) @ dz).astype(np.float32) for i in range(3)]
        xb, yb, wb = (xa + n * d[0]).astype(np.float32), (ya + n * d[1]).astype(np.float32), (wa + n * d[2]).astype(np.float32)
        return ((np.abs(xa) > np.abs(wa)) & (np.abs(xb) > np.abs(wb)) & (xa * xb > 0)) | \
               ((np.abs(ya) > np.abs(wa)) & (np.abs(yb) > np.abs(wb)) & (ya * yb > 0))

    runs = [(0, len(k) - 1)] + [(a, a + 7) for a in range(0, len(k) - 7, 8)]
    skipped = 0
    for a, b in runs:
        c = culled(a, b)[:C]  # (C,)
        assert not (sel[a:b + 1] & c[None, :]).any(), f"run [{a},{b}] culled although the oracle accepts a ray in it"
        assert not (fd[a:b + 1] & c[None, :]).any(), "a culled run holds a voxel that passes stage A"
        skipped += int(c.sum())
    assert skipped > 0.3 * C * len(runs)  # even in a column inside the visible region a good part of the pairs go


# brightest voxel of each channel on the 10 mm lattice (ix, iy) - where the z-columns of a fine lattice cross the visible region
BRIGHT_10MM = {"B": (10, 25), "C": (3, 26), "N": (9, 32), "O": (8, 25)}


@pytest.mark.parametrize("el,spacing", [("C", 1), ("N", 1), ("C", 0.65), ("O", 0.65)])
def test_run_cull_on_fine_lattice_columns(el, spacing):
    """The run-level cull on the lattices of configs 3 and 5: z-columns of 250 voxels (segments 64, 64, 64, 58) and of
    384 voxels (six segments of 64), every segment re-based from its own fp64 position as the kernel does
    (visibility.cu: filter_kernel_lattice).  Float32 emulation against the oracle: no oracle-accepted ray in a culled
    run, in the whole-segment test and in the groups of 8, for two columns through the visible region."""
    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(spacing)
        scene = make_channel(el, 10, 30, 20, lattice=mesh.lattice, host_only=True)
    lat = mesh.lattice
    assert lat.nz == (250 if spacing == 1 else 384)
    t = scene.filter_tables()
    C = 600
    ix10, iy10 = BRIGHT_10MM[el]
    ix = int(round(ix10 * (lat.nx - 1) / 134)), int(round(iy10 * (lat.ny - 1) / 79))
    osc = oracle.build_scene(el, 10, 30, 20)
    dz = (lat.step[2] * lat.rotation[2]).astype(np.float32)
    d = [(t["det_forms"][i][:, :3].astype(np.float32) @ dz).astype(np.float32) for i in range(3)]
    accepted = skipped = pairs = 0
    for dx in (0, 3):
        col = ix[1] * lat.nx + ix[0] + dx
        idx = np.arange(col * lat.nz, (col + 1) * lat.nz)
        ref = oracle.run_chunk(osc, gs.voxel_mesh(gs.load_db(), spacing, flat_indices=idx), want_rays=True)
        accepted += int(ref.mask.sum())
        for z0 in range(0, lat.nz, 64):
            n = min(64, lat.nz - z0)
            base = (mesh.points_at([idx[z0]]) - t["origin"]).astype(np.float32)
            cur = [forms(t["det_forms"][i], base)[:, 0] for i in range(3)]
            acc = [np.empty((len(cur[0]), n), np.float32) for _ in range(3)]
            for j in range(n):
                for i in range(3):
                    acc[i][:, j] = cur[i]
                    cur[i] = (cur[i] + d[i]).astype(np.float32)
            X, Y, W = acc
            fd = (np.abs(W) - np.maximum(np.abs(X), np.abs(Y)) >= 0)[:C].T  # (n, C)
            sel = ref.mask[z0:z0 + n]
            assert not (sel & ~fd).any(), "forward-differenced stage A rejected a ray the oracle accepts"

            def culled(a, b):
                m = np.float32(b - a)
                xa, ya, wa = X[:, a], Y[:, a], W[:, a]
                xb, yb, wb = (xa + m * d[0]).astype(np.float32), (ya + m * d[1]).astype(np.float32), (wa + m * d[2]).astype(np.float32)
                return ((np.abs(xa) > np.abs(wa)) & (np.abs(xb) > np.abs(wb)) & (xa * xb > 0)) | \
                       ((np.abs(ya) > np.abs(wa)) & (np.abs(yb) > np.abs(wb)) & (ya * yb > 0))

            runs = [(0, n - 1)] + [(a, a + 7) for a in range(0, n - 7, 8)]
            if n % 8:
                runs.append((n - n % 8, n - 1))  # the ragged tail is one more run
            for a, b in runs:
                c = culled(a, b)[:C]
                assert not (sel[a:b + 1] & c[None, :]).any(), f"z0={z0} run [{a},{b}] culled although the oracle accepts a ray in it"
                assert not (fd[a:b + 1] & c[None, :]).any(), "a culled run holds a voxel that passes stage A"
                skipped += int(c.sum())
                pairs += C
    assert accepted > 50, accepted
    assert skipped > 0.3 * pairs


def test_padding_lanes_never_pass():
    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(20)
        scene = make_channel("C", 10, 7, 13, lattice=mesh.lattice, host_only=True)  # 91 points -> 96 lanes
    t = scene.filter_tables()
    assert t["det_forms"].shape[1] == 96
    p = (mesh.points_at(np.arange(0, mesh.lattice.n_points, 37)) - t["origin"]).astype(np.float32)
    assert not stage_a(t, p)[91:].any()
