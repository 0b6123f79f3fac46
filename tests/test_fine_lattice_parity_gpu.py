"""Oracle parity on the fine lattices of BASELINE.json configs 3 and 5 (1 mm: 1350x800x250, 0.65 mm: 2076x1230x384).

For every channel the whole lattice is run the way production runs it (launches of 2^26 voxels through one
VisibilityPlan), then three slabs of 34 000 consecutive voxels - the centre of the visible region, its edge, and a
dark region; >= 1e5 voxels per channel and lattice - are compared with the NumPy/Qhull oracle (oracle/stage1.py, the
restatement of simulation.py:327-432, 483-594): per-voxel passing-ray counts exactly, weights to 1e-12.  The visible
voxels of the slabs then go through Stage 2 both ways: the production histogram path (device Reff bins -> millimetre
histogram -> profile rows -> flux) against oracle/stage2.py on the same voxels.  One slab is repeated through a 3-rank
block-cyclic shard, and the detector image is checked on a 1 mm slab.

The z-columns of these lattices are 250 and 384 voxels long, i.e. four and six forward-difference segments per column
with different phases: the cases the run-level cull had not been checked on against the oracle.
"""
import contextlib
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SLAB = 34_000
CHUNK = 1 << 26
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]


def _mesh(spacing):
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh

    with contextlib.redirect_stdout(io.StringIO()):
        return CuboidMesh(spacing)


def _scene(el, lattice, **kw):
    from co_monitor_b200.geometry.engine import make_channel

    with contextlib.redirect_stdout(io.StringIO()):
        return make_channel(el, 10, 30, 20, lattice=lattice, **kw)


def whole_lattice(scene, lat, chunk=CHUNK):
    """The production launch pattern: one plan, re-bound over consecutive ranges of *chunk* voxels."""
    import torch

    n = lat.n_points
    nrays = torch.empty(n, dtype=torch.int32, device="cuda")
    weight = torch.empty(n, dtype=torch.float64, device="cuda")
    plan = scene.plan(lat, 0, min(chunk, n))
    totals = {"passing_rays": 0, "visible_voxels": 0, "near_edge_rays": 0}
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        plan.rebind(lo, hi).launch()
        st = plan.stats()
        assert not st["overflow"], st
        for k in totals:
            totals[k] += st[k]
        nrays[lo:hi] = plan.nrays[: hi - lo]
        weight[lo:hi] = plan.weight[: hi - lo]
    return nrays, weight, totals


def pick_slabs(nrays, lat):
    """(name, lo, hi) of three slabs chosen from the visibility mask: around the z-column with the most visible voxels,
    across the x-edge of the visible region in that column's row, and a dark range at the end of the lattice.  The
    starts are deliberately not aligned with z-columns or 64-voxel segments."""
    import torch

    nz, nx = lat.nz, lat.nx
    counts = (nrays.view(-1, nz) > 0).sum(1)
    col_c = int(torch.argmax(counts).item())
    row = col_c // nx
    in_row = counts[row * nx:(row + 1) * nx]
    ix0 = int(torch.nonzero(in_row)[0].item())
    col_e = row * nx + ix0

    def around(col, shift):
        lo = max(0, col * nz - SLAB // 2 + shift)
        return lo, min(lat.n_points, lo + SLAB)

    dark_hi = lat.n_points - 7
    return [("centre", *around(col_c, 13)), ("edge", *around(col_e, 101)), ("dark", dark_hi - SLAB, dark_hi)]


def oracle_slab(el, spacing, lo, hi):
    from oracle import stage1 as o1

    w, nr, _ = o1.run_range(el, lo, hi, S=10, spacing=spacing, h=30, l=20, chunk=500)
    return w, nr


def stage2_flux_device(el, lat, lo, hi, weight_slab):
    """Production Stage 2 of a voxel range: device Reff bins -> millimetre histogram -> rows -> sweep."""
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import _ptr, cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.reff import resample_reff_device
    from co_monitor_b200.pipeline import LYMAN_ALPHA

    lib = _lib.load()
    prof = np.ascontiguousarray(TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    prof_t = torch.from_numpy(prof).cuda()
    K = len(prof)
    ion, wl = LYMAN_ALPHA[el]
    tables = cached_line_tables(el, ion, wl, ("EXCIT", "RECOM"), None)
    reff_mm, _ = resample_reff_device(lat, 15, begin=lo, end=hi)
    w = weight_slab.clone()  # 16-byte aligned copy
    hist = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device="cuda")
    rows = torch.zeros(K, dtype=torch.float64, device="cuda")
    bnd = torch.zeros(1, dtype=torch.float64, device="cuda")
    flux = torch.zeros(3, dtype=torch.float64, device="cuda")
    _lib.check(lib.com_weight_histogram_mm(_ptr(reff_mm), _ptr(w), hi - lo, None, _ptr(hist), None), "com_weight_histogram_mm")
    _lib.check(lib.com_histogram_rows(_ptr(prof_t), K, 1, float(prof[:, 0].max()), _ptr(hist), None, _ptr(rows), _ptr(bnd), None),
               "com_histogram_rows")
    _lib.check(lib.com_emissivity_sweep(tables.handle, _ptr(prof_t), 1, K, _ptr(rows), 0.02, _ptr(flux), None), "com_emissivity_sweep")
    return flux.cpu().numpy(), float(bnd.item()), reff_mm.cpu().numpy()


def stage2_flux_oracle(el, lat, lo, w_ref, nr_ref):
    """oracle/stage2.py on the visible voxels of a slab, Reff from the oracle's own resampling statement."""
    from co_monitor_b200.pipeline import LYMAN_ALPHA
    from oracle import reff as oreff
    from oracle import stage2 as o2

    vis = np.flatnonzero(nr_ref > 0)
    reff = oreff.on_lattice((lat.nx, lat.ny, lat.nz), vis + lo)
    if not np.isfinite(reff).any():
        return None, reff
    ion, wl = LYMAN_ALPHA[el]
    obs = np.column_stack((np.arange(len(vis)), np.zeros((len(vis), 3)), w_ref[vis]))
    res = o2.emissivity(el, ion, wl, 2, ["EXCIT", "RECOM"], o2.two_gauss_profile(NE_MAIN, TE_MAIN), obs, reff)
    return res, reff


@pytest.mark.parametrize("spacing", [1, 0.65])
@pytest.mark.parametrize("el", list("CBNO"))
def test_slabs_against_oracle_stage1_and_stage2(el, spacing):
    import torch

    from oracle import reff as oreff

    mesh = _mesh(spacing)
    lat = mesh.lattice
    assert (lat.nx, lat.ny, lat.nz) == ((1350, 800, 250) if spacing == 1 else (2076, 1230, 384))
    scene = _scene(el, lat)
    nrays, weight, totals = whole_lattice(scene, lat)
    assert totals["visible_voxels"] == int((nrays > 0).sum().item())
    checked_voxels = checked_rays = 0
    for name, lo, hi in pick_slabs(nrays, lat):
        w_ref, nr_ref = oracle_slab(el, spacing, lo, hi)
        nr = nrays[lo:hi].cpu().numpy()
        w = weight[lo:hi].cpu().numpy()
        assert np.array_equal(nr, nr_ref), (name, int((nr != nr_ref).sum()))
        np.testing.assert_allclose(w, w_ref, rtol=1e-12, atol=1e-300)
        if name == "dark":
            assert nr_ref.sum() == 0
        else:
            assert nr_ref.sum() > 1000 and (nr_ref == 0).any(), (name, int(nr_ref.sum()))
        # the same voxels launched on their own (other segment phases, other work items): identical
        direct = scene.visibility(lattice=lat, begin=lo, end=hi)
        assert torch.equal(direct.nrays, nrays[lo:hi])
        assert torch.allclose(direct.weight, weight[lo:hi], rtol=1e-13, atol=0)
        checked_voxels += hi - lo
        checked_rays += int(nr_ref.sum())
        # Stage 2 of the slab
        if name != "dark":
            flux, bnd, reff_mm = stage2_flux_device(el, lat, lo, hi, weight[lo:hi])
            ref, reff_vis = stage2_flux_oracle(el, lat, lo, w_ref, nr_ref)
            vis = np.flatnonzero(nr_ref > 0)
            assert np.array_equal(reff_mm[vis], oreff.to_mm(reff_vis))  # device Reff bins = the oracle's resampling
            if ref is None:
                assert not flux.any()
            else:
                want = np.array([ref.flux["EXCIT"], ref.flux["RECOM"], ref.flux["TOTAL"]])
                np.testing.assert_allclose(flux, want, rtol=1e-11)
                assert bnd == ref.reff_boundary
    assert checked_voxels >= 100_000 and checked_rays > 5000
    scene.close()


@pytest.mark.parametrize("el,spacing", [("N", 0.65), ("C", 1)])
def test_slab_through_three_block_cyclic_shards(el, spacing):
    """The centre slab again, each voxel computed by the rank that owns it in a 3-rank block-cyclic deal."""
    mesh = _mesh(spacing)
    lat = mesh.lattice
    scene = _scene(el, lat)
    # locate the visible region on a coarse pass: column with the most visible voxels of one mid-lattice chunk
    nrays, _, _ = whole_lattice(scene, lat)
    name, lo, hi = pick_slabs(nrays, lat)[0]
    del nrays
    w_ref, nr_ref = oracle_slab(el, spacing, lo, lo + 12_000)
    hi = lo + 12_000
    got_n = np.full(hi - lo, -1, np.int64)
    got_w = np.zeros(hi - lo)
    cpb, nz = 8, lat.nz
    for r in range(3):
        sh = lat.shard(3, r, cols_per_block=cpb)
        # local columns of this rank that intersect the slab's global columns
        cg = np.arange(lo // nz, (hi - 1) // nz + 1)
        mine = cg[(cg // cpb) % 3 == r]
        if len(mine) == 0:
            continue
        cl = (mine // cpb // 3) * cpb + mine % cpb
        a, b = int(cl.min()) * nz, (int(cl.max()) + 1) * nz
        res = scene.visibility(lattice=sh, begin=a, end=b)
        g = sh.local_to_global(np.arange(a, b))
        inside = (g >= lo) & (g < hi)
        got_n[g[inside] - lo] = res.nrays.cpu().numpy()[inside]
        got_w[g[inside] - lo] = res.weight.cpu().numpy()[inside]
    assert np.array_equal(got_n, nr_ref)
    np.testing.assert_allclose(got_w, w_ref, rtol=1e-12, atol=1e-300)
    assert nr_ref.sum() > 500
    scene.close()


def test_detector_image_on_the_1mm_lattice():
    """Config 3's per-pixel detector image on a 1 mm slab against the oracle's 2-D histogram of the hit points."""
    from oracle import geometry_setup as gs
    from oracle import stage1 as oracle

    mesh = _mesh(1)
    lat = mesh.lattice
    scene = _scene("O", lat)
    nrays, _, _ = whole_lattice(scene, lat)
    _, lo, _ = pick_slabs(nrays, lat)[0]
    del nrays
    hi = lo + 3000
    img = scene.enable_detector_image(pixel_mm=0.1)
    ny, nx = img.shape
    res = scene.visibility(lattice=lat, begin=lo, end=hi)
    got = img.cpu().numpy().copy()
    assert abs(got.sum() / float(res.weight.sum().item()) - 1) < 1e-12
    osc = oracle.build_scene("O", 10, 30, 20)
    want = np.zeros((ny, nx))
    rays = 0
    for a in range(lo, hi, 500):
        ref = oracle.run_chunk(osc, gs.voxel_mesh(gs.load_db(), 1, flat_indices=np.arange(a, min(hi, a + 500))), want_rays=True)
        want += oracle.detector_image(osc, *ref.geometry, ref.mask, ref.w_ray, nx, ny, 0.1)
        rays += int(ref.mask.sum())
    assert rays > 100 and rays == res.stats["passing_rays"]
    assert np.abs(got - want).sum() <= 1e-5 * want.sum()
    assert np.isclose(got, want, rtol=1e-10, atol=1e-300).mean() > 0.999
    scene.disable_detector_image()
    scene.close()
