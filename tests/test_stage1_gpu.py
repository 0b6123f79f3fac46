"""Stage-1 parity on the GPU: CUDA path (through the C ABI) vs the golden files shipped with the
reference and vs the CPU oracle on the same voxels."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

GOLD_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "co_monitor_b200",
                         "input_files", "geometry", "observed_plasma_volume")
PASSING_RAYS = {"B": 652069, "C": 613537, "N": 1049965, "O": 901918}  # full-grid oracle runs (oracle/stage1.py --full)


def golden(el):
    df = pd.read_csv(os.path.join(GOLD_ROOT, el, f"{el}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat"), sep=";")
    return df.sort_values("idx_sel_plas_points").reset_index(drop=True)


@pytest.fixture(scope="module")
def sims():
    from co_monitor_b200.geometry.simulation import Simulation

    cache = {}

    def get(el):
        if el not in cache:
            cache[el] = Simulation(el, slits_number=10, distance_between_points=10, crystal_height_step=30,
                                   crystal_length_step=20, plot=False, verbose=False)
        return cache[el]

    return get


@pytest.mark.parametrize("el", list("CBNO"))
def test_golden_file_parity(sims, el):
    """Full 270 000-voxel grid: identical visible set and coordinates, weights <= 1e-12 relative."""
    sim = sims(el)
    gold = golden(el)
    ddf = sim.ddf.compute()
    assert sim.stats["tests"] == 270000 * 600
    assert not sim.stats["overflow"]
    mine, ref = ddf.idx_sel_plas_points.to_numpy(), gold.idx_sel_plas_points.to_numpy()
    extra, missing = np.setdiff1d(mine, ref), np.setdiff1d(ref, mine)
    print(f"{el}: stats {sim.stats} extra {extra} missing {missing}")
    assert len(extra) == 0 and len(missing) == 0, (extra, missing)
    assert sim.stats["passing_rays"] == PASSING_RAYS[el]
    assert sim.stats["visible_voxels"] == len(gold)
    for col in ("plasma_x", "plasma_y", "plasma_z"):
        assert np.array_equal(ddf[col].to_numpy(), gold[col].to_numpy())
    rel = np.abs(ddf.total_intensity_fraction.to_numpy() / gold.total_intensity_fraction.to_numpy() - 1)
    print(f"{el}: max rel weight err {rel.max():.3e}, near-edge rays {sim.stats['near_edge_rays']}")
    assert rel.max() <= 1e-12


def test_filter_is_conservative(sims):
    """FP32 filter on vs off (every ray in fp64): identical ray counts and weights on a visible slab."""
    import torch

    sim = sims("C")
    from co_monitor_b200.geometry.engine import ChannelScene

    lo, hi = 120000, 150000
    a = sim.scene.visibility(lattice=sim.mesh.lattice, begin=lo, end=hi)
    brute = ChannelScene(
        sim.disp_elem_coord, sim.crystal_normals, sim.slit_coord_crys_side, sim.slit_coord_plasma_side,
        sim.collim_vector_front_back, sim.port_vertices_coordinates, sim.port_orientation_vector,
        sim.detector_vertices_coordinates, sim.detector_orientation_vector, origin=sim._crystal_central_point,
        area_pt=sim.crystal_point_area, refl_max=sim.max_reflectivity, aoi0=float(sim.AOI), filter_eps_mm=1e30)
    b = brute.visibility(lattice=sim.mesh.lattice, begin=lo, end=hi)
    assert b.stats["candidates"] == (hi - lo) * 600
    assert a.stats["candidates"] < 0.02 * b.stats["candidates"]
    assert a.stats["passing_rays"] == b.stats["passing_rays"] > 0
    assert torch.equal(a.nrays, b.nrays)
    assert torch.allclose(a.weight, b.weight, rtol=1e-13, atol=0)
    brute.close()


def test_ray_level_parity_with_oracle(sims):
    """Per-ray mask, distance and AOI against the NumPy/Qhull oracle on 1500 voxels of a visible region."""
    from oracle import geometry_setup as gs
    from oracle import stage1 as oracle

    sim = sims("C")
    lo, hi = 131000, 132500
    scene = oracle.build_scene("C", 10, 30, 20)
    vox = gs.voxel_mesh(gs.load_db(), 10, flat_indices=np.arange(lo, hi))
    ref = oracle.run_chunk(scene, vox, want_rays=True)
    res = sim.scene.visibility(lattice=sim.mesh.lattice, begin=lo, end=hi, want_rays=True)
    assert ref.mask.sum() > 1000
    mask = np.zeros((hi - lo) * 600, bool)
    mask[res.rays["ray_index"] - lo * 600] = True
    assert np.array_equal(mask.reshape(hi - lo, 600), ref.mask)
    order = np.argsort(res.rays["ray_index"])
    rays = res.rays[order]
    assert np.allclose(rays["distance"], ref.dist[ref.mask], rtol=1e-14, atol=0)
    assert np.array_equal(rays["aoi"], ref.aoi[ref.mask])
    assert np.allclose(rays["weight"], ref.w_ray[ref.mask], rtol=1e-13, atol=0)
    assert np.allclose(res.weight.cpu().numpy(), ref.weight, rtol=1e-13, atol=1e-300)
    assert np.array_equal(res.nrays.cpu().numpy(), ref.nrays)


def test_explicit_voxels_match_lattice(sims):
    import torch

    sim = sims("C")
    lo, hi = 100000, 140000
    a = sim.scene.visibility(lattice=sim.mesh.lattice, begin=lo, end=hi)
    pts = torch.from_numpy(sim.mesh.points_at(np.arange(0, hi))).cuda()
    b = sim.scene.visibility(voxels=pts, begin=lo, end=hi)
    assert torch.equal(a.nrays, b.nrays)
    assert torch.allclose(a.weight, b.weight, rtol=1e-12, atol=0)


@pytest.mark.parametrize("el", ["C", "O"])
def test_packed_explicit_voxels_ragged_ranges(sims, el):
    """Explicit (P,3) fp64 voxels with the packed float4 copy the filter streams by bulk copies (com_voxels_pack +
    com_visibility_weights_packed) against the lattice path (oracle-checked) and against the unpacked explicit call:
    ranges that start and end off the 512-voxel tiles, a range shorter than one tile, a reused packed array."""
    import ctypes as C

    import torch

    from co_monitor_b200 import _lib

    sim = sims(el)
    scene, lat = sim.scene, sim.mesh.lattice
    hi_all = 150_003
    pts = torch.from_numpy(sim.mesh.points_at(np.arange(0, hi_all))).cuda()
    packed = scene.pack_voxels(pts)
    ref_pack = (pts - torch.tensor(list(scene.desc.origin), dtype=torch.float64, device="cuda")).to(torch.float32)
    assert torch.equal(packed[:, :3], ref_pack) and bool((packed[:, 3] == 1).all())
    for lo, hi in [(100_001, 140_777), (120_000, 120_300), (0, 513), (149_000, hi_all)]:
        a = scene.visibility(lattice=lat, begin=lo, end=hi)
        b = scene.visibility(voxels=pts, begin=lo, end=hi, packed=packed)
        assert torch.equal(a.nrays, b.nrays), (lo, hi)
        assert torch.allclose(a.weight, b.weight, rtol=1e-12, atol=0)
        assert a.stats["passing_rays"] == b.stats["passing_rays"]
    # the unpacked explicit call (tiles converted from fp64 inside the kernel) gives the same candidates and results
    lib = _lib.load()
    lo, hi = 100_001, 140_777
    n = hi - lo
    ws_bytes = lib.com_visibility_workspace_bytes(scene.handle, n, 0.0)
    outs = []
    for use_packed in (False, True):
        w = torch.empty(n, dtype=torch.float64, device="cuda")
        nr = torch.empty(n, dtype=torch.int32, device="cuda")
        st = torch.zeros(_lib.N_STATS, dtype=torch.int64, device="cuda")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
        tail = (lo, hi, C.c_void_p(w.data_ptr()), C.c_void_p(nr.data_ptr()), None, 0, None, None, None, 0,
                C.c_void_p(st.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        if use_packed:
            rc = lib.com_visibility_weights_packed(scene.handle, C.c_void_p(pts.data_ptr()), C.c_void_p(packed.data_ptr()), *tail)
        else:
            rc = lib.com_visibility_weights(scene.handle, None, C.c_void_p(pts.data_ptr()), *tail)
        _lib.check(rc, "explicit voxels")
        torch.cuda.synchronize()
        outs.append((w, nr, st.cpu().tolist()))
    assert torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]
    assert torch.allclose(outs[0][0], outs[1][0], rtol=1e-13, atol=0)
    # misaligned packed array is refused, not read
    with pytest.raises(_lib.ComonitorError):
        bad = torch.empty(hi_all * 4 + 1, dtype=torch.float32, device="cuda")[1:].view(hi_all, 4)
        scene.visibility(voxels=pts, begin=0, end=1000, packed=bad)


def test_reference_attributes(sims):
    sim = sims("C")
    assert sim.crystal_point_area == 2.67
    assert list(sim.ddf.columns) == ["idx_sel_plas_points", "plasma_x", "plasma_y", "plasma_z", "total_intensity_fraction"]
    sel = sim.selected_intersections
    assert sel.shape == (270000, 600) and sel.sum() == PASSING_RAYS["C"]
    assert len(sim.distances) == len(sim.angles_of_incident) == PASSING_RAYS["C"]
    per_voxel = sel.sum(axis=1)
    assert np.array_equal(np.flatnonzero(per_voxel), sim.ddf.idx_sel_plas_points.to_numpy())


def test_no_visible_voxel_exits_like_the_reference():
    from co_monitor_b200.geometry.simulation import Simulation

    with pytest.raises(SystemExit) as exc:
        # a single 1x1 crystal grid point at the crystal corner sees nothing through the collimator at 100 mm spacing
        Simulation("C", slits_number=1, distance_between_points=120, crystal_height_step=1, crystal_length_step=1,
                   plot=False, verbose=False)
    assert "Not enough points" in str(exc.value)


def test_host_buffer_entry_points(sims):
    """com_visibility_weights_host (dense) and com_visible_weights_host (compacted) against the device-buffer path."""
    import ctypes as C

    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.engine import lattice_struct

    sim = sims("C")
    lib = _lib.load()
    lat = lattice_struct(sim.mesh.lattice)
    lo, hi = 100000, 160000
    ref = sim.scene.visibility(lattice=sim.mesh.lattice, begin=lo, end=hi)
    w = np.empty(hi - lo)
    nr = np.empty(hi - lo, dtype=np.uint32)
    st = _lib.ComStats()
    rc = lib.com_visibility_weights_host(C.byref(sim.scene.desc), C.byref(lat), None, lo, hi, C.c_void_p(w.ctypes.data),
                                         C.c_void_p(nr.ctypes.data), None, 0, None, C.byref(st))
    assert rc == 0
    assert np.array_equal(nr, ref.nrays.cpu().numpy().view(np.uint32))
    np.testing.assert_allclose(w, ref.weight.cpu().numpy(), rtol=1e-13, atol=0)
    assert st.passing_rays == ref.stats["passing_rays"] and st.visible_voxels == ref.stats["visible_voxels"]
    idx = np.empty(hi - lo, dtype=np.int64)
    wc = np.empty(hi - lo)
    cnt = C.c_uint64()
    for _ in range(2):  # second call reuses the arena
        rc = lib.com_visible_weights_host(C.byref(sim.scene.desc), C.byref(lat), None, lo, hi, C.c_void_p(idx.ctypes.data),
                                          C.c_void_p(wc.ctypes.data), hi - lo, C.byref(cnt), C.byref(st))
        assert rc == 0
        m = cnt.value
        assert m == st.visible_voxels == int((nr > 0).sum())
        assert np.array_equal(idx[:m], np.flatnonzero(nr) + lo)
        np.testing.assert_allclose(wc[:m], w[nr > 0], rtol=1e-13, atol=0)
    # explicit host voxels, same range
    pts = np.ascontiguousarray(sim.mesh.points_at(np.arange(lo, hi)))
    rc = lib.com_visible_weights_host(C.byref(sim.scene.desc), None, C.c_void_p(pts.ctypes.data), lo, hi,
                                      C.c_void_p(idx.ctypes.data), C.c_void_p(wc.ctypes.data), hi - lo, C.byref(cnt),
                                      C.byref(st))
    assert rc == 0 and cnt.value == m
    # too small a capacity is reported, an empty range is the reference's "Not enough points" condition
    rc = lib.com_visible_weights_host(C.byref(sim.scene.desc), C.byref(lat), None, lo, hi, C.c_void_p(idx.ctypes.data),
                                      C.c_void_p(wc.ctypes.data), 10, C.byref(cnt), C.byref(st))
    assert rc == _lib.COM_E_OVERFLOW
    rc = lib.com_visible_weights_host(C.byref(sim.scene.desc), C.byref(lat), None, 0, 2000, C.c_void_p(idx.ctypes.data),
                                      C.c_void_p(wc.ctypes.data), hi - lo, C.byref(cnt), C.byref(st))
    assert rc == _lib.COM_E_EMPTY and b"Not enough points" in lib.com_last_error()
    assert lib.com_host_arena_release() == 0


def test_block_cyclic_shards_equal_the_unsharded_grid(sims):
    """Three block-cyclic shards (what three ranks would run) reproduce the unsharded per-voxel results exactly."""
    import torch

    sim = sims("C")
    lat = sim.mesh.lattice
    full = sim.scene.visibility(lattice=lat)
    nr_full, w_full = full.nrays.cpu().numpy(), full.weight.cpu().numpy()
    seen = np.zeros(lat.n_points, bool)
    total_rays = 0
    for r in range(3):
        sh = lat.shard(3, r, cols_per_block=7)
        res = sim.scene.visibility(lattice=sh)
        g = sh.local_to_global(np.arange(sh.n_points))
        assert not seen[g].any()
        seen[g] = True
        assert np.array_equal(res.nrays.cpu().numpy(), nr_full[g])
        np.testing.assert_allclose(res.weight.cpu().numpy(), w_full[g], rtol=1e-13, atol=0)
        total_rays += res.stats["passing_rays"]
        # a sub-range of a shard too
        sub = sim.scene.visibility(lattice=sh, begin=1000, end=31000)
        assert np.array_equal(sub.nrays.cpu().numpy(), nr_full[g[1000:31000]])
    assert seen.all() and total_rays == PASSING_RAYS["C"]


@pytest.mark.parametrize("el,spacing,h,l,slits,centre,side", [
    ("O", 10, 20, 80, 10, 140000, "top closing side"),   # class defaults of Simulation: 1600 crystal points (no golden file)
    ("B", 15, 10, 10, 10, 40000, "top closing side"),    # the reference's __main__ settings: 15 mm grid, 100 crystal points
    ("N", 10, 30, 20, 4, 150000, "top closing side"),    # fewer slits
    ("C", 20, 7, 13, 10, 17000, "top closing side"),     # odd crystal sampling: 91 points (padded warp)
    ("C", 10, 30, 20, 10, 135000, "bottom closing side"),  # the collimator position the reference never selects
    ("C", 10, 30, 20, 1, 135000, "top closing side"),    # a single slit: no period to exploit
    ("O", 10, 30, 20, 18, 145000, "top closing side"),   # more slits than the 4-bit slit hint of a candidate can name
])
def test_other_configurations_against_oracle(el, spacing, h, l, slits, centre, side):
    """Ray-exact agreement with the NumPy/Qhull oracle on slabs of configurations that have no shipped golden file,
    with the FP32 filter on, and the same slab with the filter off (every ray through fp64)."""
    import contextlib
    import io

    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from oracle import geometry_setup as gs
    from oracle import stage1 as oracle

    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(spacing)
        scene = make_channel(el, slits, h, l, lattice=mesh.lattice, closing_side=side)
        brute = make_channel(el, slits, h, l, lattice=mesh.lattice, filter_eps_mm=1e30, closing_side=side)
    lat = mesh.lattice
    full = scene.visibility(lattice=lat)
    nr = full.nrays.cpu().numpy()
    vis = np.flatnonzero(nr)
    assert len(vis) > 0
    # slab around the visible voxel nearest to `centre`
    c = int(vis[np.argmin(np.abs(vis - centre))])
    lo = max(0, c - 300)
    hi = min(lat.n_points, lo + 600)
    idx = np.arange(lo, hi)
    osc = oracle.build_scene(el, slits, h, l, side=side)
    ref = oracle.run_chunk(osc, gs.voxel_mesh(gs.load_db(), spacing, flat_indices=idx))
    assert ref.nrays.sum() > 0
    assert np.array_equal(nr[lo:hi], ref.nrays)
    np.testing.assert_allclose(full.weight.cpu().numpy()[lo:hi], ref.weight, rtol=1e-12, atol=1e-300)
    b = brute.visibility(lattice=lat, begin=lo, end=hi)
    assert np.array_equal(b.nrays.cpu().numpy(), ref.nrays)
    assert b.stats["candidates"] == (hi - lo) * h * l
    scene.close()
    brute.close()


@pytest.mark.parametrize("el", list("BNO"))
def test_filter_is_conservative_other_channels(sims, el):
    """Filter on vs off on 20 000 voxels per channel: identical per-voxel ray counts."""
    import contextlib
    import io

    import torch

    from co_monitor_b200.geometry.engine import make_channel

    sim = sims(el)
    lat = sim.mesh.lattice
    with contextlib.redirect_stdout(io.StringIO()):
        brute = make_channel(el, 10, 30, 20, lattice=lat, filter_eps_mm=1e30)
    lo, hi = 125000, 145000
    a = sim.scene.visibility(lattice=lat, begin=lo, end=hi)
    b = brute.visibility(lattice=lat, begin=lo, end=hi)
    assert a.stats["passing_rays"] == b.stats["passing_rays"] > 0
    assert torch.equal(a.nrays, b.nrays)
    brute.close()


def test_fused_reff_bin_histogram(sims):
    """K1b can accumulate the weights straight into a Reff-bin histogram (no per-voxel pass afterwards):
    hist[bin[v]] += w_ray.  Checked against a bincount of the per-voxel weights; bins >= n_bins are skipped."""
    import torch

    sim = sims("N")
    lat = sim.mesh.lattice
    rng = np.random.default_rng(11)
    bins = rng.integers(0, 600, lat.n_points).astype(np.uint16)
    bins[rng.random(lat.n_points) < 0.3] = 0xFFFF
    bins_d = torch.from_numpy(bins.view(np.int16)).cuda()
    hist = torch.zeros(540, dtype=torch.float64, device="cuda")
    lo, hi = 60000, 250000
    res = sim.scene.visibility(lattice=lat, begin=lo, end=hi, reff_bin=bins_d, hist=hist)
    w = res.weight.cpu().numpy()
    b = bins[lo:hi].astype(np.int64)
    keep = b < 540
    ref = np.bincount(b[keep], weights=w[keep], minlength=540)
    np.testing.assert_allclose(hist.cpu().numpy(), ref, rtol=1e-11, atol=1e-300)
    assert ref.sum() > 0


def test_detector_image_extension(sims):
    """Per-pixel detector image (config 3 extension): weight-conserving, and equal to the oracle's histogram of the
    detector hit points on a slab (a ray exactly on a pixel boundary may fall on either side: L1 tolerance)."""
    from oracle import geometry_setup as gs
    from oracle import stage1 as oracle

    sim = sims("O")
    lat = sim.mesh.lattice
    img = sim.scene.enable_detector_image(pixel_mm=0.1)
    ny, nx = img.shape
    assert (nx, ny) in ((266, 67), (267, 68), (267, 67), (266, 68))
    lo, hi = 131000, 132000
    res = sim.scene.visibility(lattice=lat, begin=lo, end=hi)
    got = img.cpu().numpy().copy()
    assert abs(got.sum() / float(res.weight.sum().item()) - 1) < 1e-12
    osc = oracle.build_scene("O", 10, 30, 20)
    ref = oracle.run_chunk(osc, gs.voxel_mesh(gs.load_db(), 10, flat_indices=np.arange(lo, hi)), want_rays=True)
    plasma, crystal, reflected = ref.geometry
    want = oracle.detector_image(osc, plasma, crystal, reflected, ref.mask, ref.w_ray, nx, ny, 0.1)
    assert want.sum() > 0
    assert np.abs(got - want).sum() <= 1e-5 * want.sum()
    same = np.isclose(got, want, rtol=1e-10, atol=1e-300)
    assert same.mean() > 0.999
    # the whole grid: total weight conserved, image confined to the detector
    img.zero_()
    full = sim.scene.visibility(lattice=lat)
    assert abs(float(img.sum().item()) / float(full.weight.sum().item()) - 1) < 1e-12
    sim.scene.disable_detector_image()
    again = sim.scene.visibility(lattice=lat, begin=lo, end=hi)
    assert float(img.sum().item()) == float(img.sum().item())  # untouched after disabling
    assert again.stats["passing_rays"] == res.stats["passing_rays"]


def test_edge_cases_empty_range_overflow_and_bad_arguments(sims):
    """Empty voxel range, candidate-list overflow (reported, then recovered by re-running with the measured count),
    and argument validation through the C ABI."""
    import ctypes as C

    import torch

    from co_monitor_b200 import _lib

    sim = sims("C")
    lat = sim.mesh.lattice
    lib = _lib.load()
    # empty range: nothing to do, statistics all zero
    res = sim.scene.visibility(lattice=lat, begin=1234, end=1234)
    assert res.weight.numel() == 0 and res.stats["passing_rays"] == 0 and res.stats["tests"] == 0
    # a candidate list that is far too small: the first launch overflows, the wrapper re-runs with the measured count
    lo, hi = 120000, 150000
    want = sim.scene.visibility(lattice=lat, begin=lo, end=hi)
    tiny = sim.scene.visibility(lattice=lat, begin=lo, end=hi, candidate_fraction=1e-9)
    assert torch.equal(tiny.nrays, want.nrays) and not tiny.stats["overflow"]
    # ... and the raw entry point reports the overflow instead of hiding it
    from co_monitor_b200.geometry.engine import lattice_struct

    n = hi - lo
    w = torch.empty(n, dtype=torch.float64, device="cuda")
    nr = torch.empty(n, dtype=torch.int32, device="cuda")
    st = torch.zeros(_lib.N_STATS, dtype=torch.int64, device="cuda")
    ws = torch.empty(256 + 24 * 100, dtype=torch.uint8, device="cuda")  # room for 100 candidates (and 100 interval records) only
    ls = lattice_struct(lat)
    rc = lib.com_visibility_weights(sim.scene.handle, C.byref(ls), None, lo, hi, C.c_void_p(w.data_ptr()),
                                    C.c_void_p(nr.data_ptr()), None, 0, None, None, None, 0, C.c_void_p(st.data_ptr()),
                                    C.c_void_p(ws.data_ptr()), ws.numel(), None)
    assert rc == 0
    torch.cuda.synchronize()
    stats = dict(zip([f[0] for f in _lib.ComStats._fields_], st.cpu().tolist()))
    assert stats["overflow"] == 1 and stats["candidates"] > stats["candidate_capacity"] == 100
    # argument validation
    assert lib.com_visibility_weights(None, C.byref(ls), None, 0, 10, C.c_void_p(w.data_ptr()), None, None, 0, None, None,
                                      None, 0, C.c_void_p(st.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), None) == _lib.COM_E_BADARG
    assert lib.com_visibility_weights(sim.scene.handle, C.byref(ls), None, 0, lat.n_points + 1, C.c_void_p(w.data_ptr()),
                                      None, None, 0, None, None, None, 0, C.c_void_p(st.data_ptr()),
                                      C.c_void_p(ws.data_ptr()), ws.numel(), None) == _lib.COM_E_BADARG
    assert b"outside the lattice" in lib.com_last_error()
    with pytest.raises(ValueError):
        sim.scene.visibility(lattice=lat, voxels=torch.zeros((3, 3), dtype=torch.float64, device="cuda"))
    with pytest.raises(ValueError):
        sim.scene.visibility(voxels=torch.zeros((3, 3), dtype=torch.float32, device="cuda"))


def test_chunked_launches_and_result_files(sims, tmp_path):
    """Simulation over several launches gives the same frame; the files it writes: text = pandas' bytes (and so the
    format of the shipped files), dask-style partitions, and the binary side format - all read back to the frame."""
    from co_monitor_b200.geometry import stage1_io
    from co_monitor_b200.geometry.simulation import Simulation

    whole = sims("O")
    chunked = Simulation("O", slits_number=10, distance_between_points=10, crystal_height_step=30, crystal_length_step=20,
                         plot=False, verbose=False, stage1_chunk=47_123, output_dir=tmp_path)  # 6 ragged launches
    a, b = pd.DataFrame(whole.ddf), pd.DataFrame(chunked.ddf)
    assert a.idx_sel_plas_points.equals(b.idx_sel_plas_points)
    np.testing.assert_allclose(b.total_intensity_fraction, a.total_intensity_fraction, rtol=1e-13)
    for key in ("tests", "passing_rays", "visible_voxels", "near_edge_rays"):
        assert chunked.stats[key] == whole.stats[key]
    path = chunked.save_to_file()
    assert path.name == "O_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat"
    want = tmp_path / "pandas.csv"
    b.to_csv(want, sep=";", header=True, index=False)
    assert path.read_bytes() == want.read_bytes()
    pd.testing.assert_frame_equal(stage1_io.read(path), b)
    first = chunked.save_to_file(partitions=4)
    assert first.name.endswith("slit_100.dat") and (tmp_path / first.name.replace("slit_100", "slit_103")).exists()
    pd.testing.assert_frame_equal(stage1_io.read(str(first).replace("slit_100", "slit_10*")), b)
    binary = chunked.save_to_file(binary=True)
    pd.testing.assert_frame_equal(stage1_io.read(binary, mesh=chunked.mesh), b)


@pytest.mark.parametrize("el", list("CBNO"))
def test_the_three_lattice_filters_decide_alike(sims, el):
    """The production column-interval kernel, the per-ray kernel with its run-level cull (COM_FILTER_NO_INTERVALS) and the
    per-ray kernel that evaluates every ray's detector test on its own (COM_FILTER_NO_RUN_CULL): candidate sets may
    differ by what FP32 rounding does inside the 5e-3 mm band, the decisions are identical - and equal to the golden
    file's ray count."""
    import contextlib
    import io

    import torch

    from co_monitor_b200.geometry.engine import make_channel

    sim = sims(el)
    lat = sim.mesh.lattice
    with contextlib.redirect_stdout(io.StringIO()):
        plain = make_channel(el, 10, 30, 20, lattice=lat, run_cull=False)
        culled = make_channel(el, 10, 30, 20, lattice=lat, intervals=False)
    a = sim.scene.visibility(lattice=lat)
    b = plain.visibility(lattice=lat)
    c = culled.visibility(lattice=lat)
    for other in (b, c):
        assert torch.equal(a.nrays, other.nrays)
        torch.testing.assert_close(a.weight, other.weight, rtol=1e-13, atol=0)
        assert other.stats["passing_rays"] == PASSING_RAYS[el]
        # the interval kernel follows the slit polygons edge by edge, the per-ray kernels their bounding rectangles
        assert PASSING_RAYS[el] <= a.stats["candidates"] <= other.stats["candidates"] + 16
    assert a.stats["passing_rays"] == PASSING_RAYS[el]
    # the interval kernel vouches for most passing rays (the fp64 stage then only computes their weight)
    assert a.stats["certain_candidates"] > 0.7 * a.stats["passing_rays"]  # 10 mm voxels: intervals of 1-2 voxels
    assert abs(b.stats["candidates"] - c.stats["candidates"]) <= 16
    # a range that starts and ends inside z-columns
    lo, hi = 100_007, 200_011
    part = sim.scene.visibility(lattice=lat, begin=lo, end=hi)
    assert torch.equal(part.nrays, a.nrays[lo:hi])
    torch.testing.assert_close(part.weight, a.weight[lo:hi], rtol=1e-13, atol=0)
    plain.close()
    culled.close()
