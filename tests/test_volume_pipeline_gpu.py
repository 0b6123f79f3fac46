"""Whole-volume pipeline (csrc/volume.cu, co_monitor_b200/volume.py): Stage 1 in launches -> millimetre histograms ->
(exchange) -> fused Stage 2, device-resident and through the host-buffer C-ABI calls, against the outputs of the
unmodified reference (tests/golden/stage2_reference.npz: per-channel flux and Reff boundary of the 10 mm grid) and the
shipped Stage-1 golden files (ray and voxel counts)."""
import contextlib
import ctypes as C
import io
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ELEMENTS = "BCNO"
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]
PASSING_RAYS = {"B": 652069, "C": 613537, "N": 1049965, "O": 901918}
VISIBLE = {"B": 32667, "C": 27973, "N": 48713, "O": 50285}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stage2_reference.npz"))


@pytest.fixture(scope="module")
def setup():
    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import source_volume
    from co_monitor_b200.pipeline import LYMAN_ALPHA

    with contextlib.redirect_stdout(io.StringIO()):
        lat = CuboidMesh(10).lattice
        scenes = [make_channel(el, 10, 30, 20, lattice=lat) for el in ELEMENTS]
    tables = [cached_line_tables(el, *LYMAN_ALPHA[el], ("EXCIT", "RECOM"), None) for el in ELEMENTS]
    prof = np.ascontiguousarray(TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    yield lat, scenes, tables, prof, source_volume(15)
    for s in scenes:
        s.close()


def check_against_reference(flux, boundary, stats, gold):
    for i, el in enumerate(ELEMENTS):
        np.testing.assert_allclose(flux[i, 0], gold[f"main_{el}_flux"], rtol=1e-11)
        assert boundary[i] == float(gold[f"main_{el}_reff_boundary"][0])
        assert stats[i]["passing_rays"] == PASSING_RAYS[el]
        assert stats[i]["visible_voxels"] == VISIBLE[el]
        assert stats[i]["tests"] == 270000 * 600 and not stats[i]["overflow"]


def test_device_pipeline_matches_the_reference(setup, gold):
    """Unsharded, several ragged launches per channel."""
    from co_monitor_b200.volume import VolumePipeline

    lat, scenes, tables, prof, src = setup
    pipe = VolumePipeline(lat, scenes, tables, prof, 2.0, chunk=70_001)
    assert pipe.volume.chunk % 8 == 0 and pipe.volume.n_launches == 4
    pipe.volume.set_reff_resampled(src)
    flux = pipe.step().cpu().numpy()
    check_against_reference(flux, pipe.boundary.cpu().numpy(), pipe.global_stats(), gold)
    again = pipe.step().cpu().numpy()  # buffers are overwritten, not accumulated
    np.testing.assert_allclose(again, flux, rtol=1e-13)
    # the volume's Reff bins are the Reff provider's
    from co_monitor_b200.geometry.reff import resample_reff_device

    mm, _ = resample_reff_device(lat, 15)
    assert np.array_equal(pipe.volume.reff_mm().cpu().numpy(), mm.cpu().numpy())
    pipe.close()


def test_three_shards_sum_to_the_unsharded_result(setup, gold):
    """What three ranks would do: per-shard packed buffers added (the all-reduce), Stage 2 from the sum."""
    import torch

    from co_monitor_b200.volume import VolumePipeline, _stats_dicts, flux_from_histograms

    lat, scenes, tables, prof, src = setup
    total = None
    for r in range(3):
        pipe = VolumePipeline(lat.shard(3, r, cols_per_block=8), scenes, tables, prof, 2.0, chunk=40_000)
        pipe.volume.set_reff_resampled(src)
        pipe.stage1()
        total = pipe.packed.clone() if total is None else total + pipe.packed
        pipe.close()
    n_ch = len(ELEMENTS)
    from co_monitor_b200 import _lib

    hist = total[: n_ch * _lib.COM_MM_BINS].view(n_ch, _lib.COM_MM_BINS).contiguous()
    stats = _stats_dicts(total[n_ch * _lib.COM_MM_BINS:].cpu().numpy().round().astype(np.uint64))
    flux, bnd = flux_from_histograms(tables, torch.from_numpy(prof[None].copy()).cuda(), hist, 0.02)
    check_against_reference(flux.cpu().numpy(), bnd.cpu().numpy(), stats, gold)


def test_exchange_fused_into_stage2_three_rank_buffers(setup, gold):
    """com_histograms_flux_peers: the kernel that replaces all-reduce + Stage 2 at N > 1 reads every rank's packed buffer
    itself.  Here the three ranks' buffers live on one GPU (same kernel, same pointers-per-rank interface): reference
    flux and counters, sums bit-identical to adding the buffers in rank order, and the many-slice (two-phase) form."""
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.volume import VolumePipeline, _stats_dicts, flux_from_histograms, flux_from_peer_buffers

    lat, scenes, tables, prof, src = setup
    n_ch = len(ELEMENTS)
    bufs = []
    for r in range(3):
        pipe = VolumePipeline(lat.shard(3, r, cols_per_block=8), scenes, tables, prof, 2.0, chunk=40_000)
        pipe.volume.set_reff_resampled(src)
        pipe.stage1()
        bufs.append(pipe.packed.clone())
        pipe.close()
    profiles = torch.from_numpy(prof[None].copy()).cuda()
    flux = torch.zeros((n_ch, 1, 3), dtype=torch.float64, device="cuda")
    bnd = torch.zeros(n_ch, dtype=torch.float64, device="cuda")
    hsum = torch.zeros((n_ch, _lib.COM_MM_BINS), dtype=torch.float64, device="cuda")
    csum = torch.zeros((n_ch, _lib.N_STATS), dtype=torch.float64, device="cuda")
    flux_from_peer_buffers(tables, profiles, [b.data_ptr() for b in bufs], _lib.N_STATS, 0.02, flux, bnd, hsum, csum)
    stats = _stats_dicts(csum.cpu().numpy().round().astype(np.uint64))
    check_against_reference(flux.cpu().numpy(), bnd.cpu().numpy(), stats, gold)
    total = (bufs[0] + bufs[1]) + bufs[2]  # rank order
    assert torch.equal(hsum.ravel(), total[: n_ch * _lib.COM_MM_BINS])
    assert torch.equal(csum.ravel(), total[n_ch * _lib.COM_MM_BINS:])
    # the fused launch equals all-reduce + com_histograms_flux bit for bit
    f2, b2 = flux_from_histograms(tables, profiles, hsum, 0.02)
    assert torch.equal(f2, flux) and torch.equal(b2, bnd)
    # more than four slices: the peers are summed once, then the sweep runs from the local copy
    profs = torch.from_numpy(np.stack([prof * [1, 1 + 0.1 * i, 1 - 0.05 * i] for i in range(6)])).cuda()
    f6 = torch.zeros((n_ch, 6, 3), dtype=torch.float64, device="cuda")
    flux_from_peer_buffers(tables, profs, [b.data_ptr() for b in bufs], _lib.N_STATS, 0.02, f6, bnd, hsum, None)
    g6, _ = flux_from_histograms(tables, profs, hsum, 0.02)
    assert torch.equal(f6, g6) and torch.equal(f6[:, 0], flux[:, 0])
    with pytest.raises(_lib.ComonitorError):  # the two-phase form needs somewhere to keep the sums
        flux_from_peer_buffers(tables, profs, [b.data_ptr() for b in bufs], _lib.N_STATS, 0.02, f6, bnd, None, None)


def test_host_buffer_forms(setup, gold):
    """com_volume_histograms_host + com_histograms_flux_host: host arrays in and out, Reff source uploaded in the call."""
    from co_monitor_b200.volume import ObservationVolume, flux_from_histograms_host

    lat, scenes, tables, prof, src = setup
    vol = ObservationVolume(lat, chunk=100_000)
    hist, stats = vol.histograms_host(scenes, reff_source=src)
    flux, bnd = flux_from_histograms_host(tables, prof, hist, 0.02)
    check_against_reference(flux, bnd, stats, gold)
    hist2, _ = vol.histograms_host(scenes)  # Reff bins kept from the first call
    np.testing.assert_allclose(hist2, hist, rtol=1e-12, atol=1e-300)
    # several time slices in one call: slice s of the sweep equals a single-slice call with that profile
    profs = np.stack([prof, prof * [1, 1.3, 0.7], prof * [1, 0.5, 1.1]])
    f3, _ = flux_from_histograms_host(tables, profs, hist, 0.02)
    for s in range(3):
        f1, _ = flux_from_histograms_host(tables, profs[s], hist, 0.02)
        np.testing.assert_allclose(f3[:, s], f1[:, 0], rtol=1e-13)
    np.testing.assert_allclose(f3[:, 0], flux[:, 0], rtol=1e-13)
    vol.close()
    # a volume without Reff bins refuses to run
    from co_monitor_b200._lib import ComonitorError

    bare = ObservationVolume(lat, chunk=100_000)
    with pytest.raises(ComonitorError):
        bare.histograms_host(scenes)
    bare.close()


def test_fused_stage2_equals_the_two_kernel_form(setup):
    """com_histograms_flux against com_histogram_rows + com_emissivity_sweep on random histograms, unsorted profile too."""
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import _ptr
    from co_monitor_b200.volume import flux_from_histograms

    lat, scenes, tables, prof, src = setup
    lib = _lib.load()
    rng = np.random.default_rng(5)
    hist = np.zeros((4, _lib.COM_MM_BINS))
    for c in range(4):
        nb = rng.integers(50, 700)
        hist[c, rng.choice(800, nb, replace=False)] = rng.uniform(1e-9, 1e-5, nb)
    hist_t = torch.from_numpy(hist).cuda()
    for p in (prof, prof[rng.permutation(len(prof))]):
        p = np.ascontiguousarray(p)
        prof_t = torch.from_numpy(p[None].copy()).cuda()
        flux, bnd = flux_from_histograms(tables, prof_t, hist_t, 0.02)
        srt = int(np.all(np.diff(p[:, 0]) >= 0))
        for c in range(4):
            rows = torch.zeros(len(p), dtype=torch.float64, device="cuda")
            b2 = torch.zeros(1, dtype=torch.float64, device="cuda")
            f2 = torch.zeros(3, dtype=torch.float64, device="cuda")
            _lib.check(lib.com_histogram_rows(_ptr(prof_t), len(p), srt, float(p[:, 0].max()), _ptr(hist_t[c]), None, _ptr(rows),
                                              _ptr(b2), None), "com_histogram_rows")
            _lib.check(lib.com_emissivity_sweep(tables[c].handle, _ptr(prof_t), 1, len(p), _ptr(rows), 0.02, _ptr(f2), None),
                       "com_emissivity_sweep")
            np.testing.assert_allclose(flux[c, 0].cpu().numpy(), f2.cpu().numpy(), rtol=1e-12)
            assert float(bnd[c].item()) == float(b2.item())


def test_handle_variants_of_the_host_calls(setup, reff10):
    """com_visible_weights_host_scene / com_emissivity_integrate_indexed_host_tables (handles kept by the caller) against
    the description-taking calls."""
    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.engine import lattice_struct

    lat, scenes, tables, prof, src = setup
    lib = _lib.load()
    ls = lattice_struct(lat)
    lo, hi = 90_000, 200_000
    sc, tb = scenes[1], tables[1]
    out = {}
    for name in ("desc", "handle"):
        idx, w = np.empty(hi - lo, np.int64), np.empty(hi - lo)
        cnt, st = C.c_uint64(), _lib.ComStats()
        if name == "desc":
            rc = lib.com_visible_weights_host(C.byref(sc.desc), C.byref(ls), None, lo, hi, C.c_void_p(idx.ctypes.data),
                                              C.c_void_p(w.ctypes.data), hi - lo, C.byref(cnt), C.byref(st))
        else:
            rc = lib.com_visible_weights_host_scene(sc.handle, C.byref(ls), None, lo, hi, C.c_void_p(idx.ctypes.data),
                                                    C.c_void_p(w.ctypes.data), hi - lo, C.byref(cnt), C.byref(st))
        assert rc == 0, lib.com_last_error()
        m = cnt.value
        flux = np.zeros(3)
        bnd = C.c_double()
        args = (_lib.as_double_ptr(prof), len(prof), 0.02, _lib.as_double_ptr(reff10), len(reff10), C.c_void_p(idx.ctypes.data),
                _lib.as_double_ptr(w), m, _lib.as_double_ptr(flux), None, None, C.byref(bnd))
        if name == "desc":
            rc = lib.com_emissivity_integrate_indexed_host(C.byref(tb.desc), *args)
        else:
            rc = lib.com_emissivity_integrate_indexed_host_tables(tb.handle, *args)
        assert rc == 0, lib.com_last_error()
        out[name] = (idx[:m].copy(), w[:m].copy(), st.as_dict(), flux, bnd.value)
    a, b = out["desc"], out["handle"]
    assert np.array_equal(a[0], b[0]) and len(a[0]) > 10_000
    np.testing.assert_allclose(a[1], b[1], rtol=1e-13)
    assert a[2]["passing_rays"] == b[2]["passing_rays"] and a[4] == b[4]
    np.testing.assert_allclose(a[3], b[3], rtol=1e-12)
    # too small a capacity: reported, and nothing is written beyond it
    idx = np.full(64, -7, np.int64)
    w = np.zeros(64)
    cnt, st = C.c_uint64(), _lib.ComStats()
    rc = lib.com_visible_weights_host_scene(sc.handle, C.byref(ls), None, lo, hi, C.c_void_p(idx.ctypes.data),
                                            C.c_void_p(w.ctypes.data), 16, C.byref(cnt), C.byref(st))
    assert rc == _lib.COM_E_OVERFLOW and (idx[16:] == -7).all() and np.array_equal(idx[:16], a[0][:16])


def test_full_size_volume_three_filters_agree():
    """BASELINE configs[4] at its full size (0.65 mm lattice, 9.8e8 voxels x 600 crystal points, channel C): the shipped
    column-interval path, the per-ray path with its run-level cull and the every-ray path decide 5.9e11 rays alike - equal
    passing-ray and visible-voxel counts, Reff-millimetre histograms equal to rounding - and the flux is the same from any
    of them.  Size-independent properties stand in for the oracle, which cannot run 5.9e11 rays; slabs of this lattice are
    compared with the oracle ray by ray in tests/test_fine_lattice_parity_gpu.py."""
    import torch

    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import source_volume
    from co_monitor_b200.pipeline import LYMAN_ALPHA
    from co_monitor_b200.volume import VolumePipeline

    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~45 GB of device memory")
    with contextlib.redirect_stdout(io.StringIO()):
        lat = CuboidMesh(0.65).lattice
    assert lat.n_points == 2076 * 1230 * 384
    tables = [cached_line_tables("C", *LYMAN_ALPHA["C"], ("EXCIT", "RECOM"), None)]
    prof = np.ascontiguousarray(TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    results = {}
    pipe = None
    for name, kw in (("interval", {}), ("per_ray_cull", {"intervals": False}), ("every_ray", {"intervals": False, "run_cull": False})):
        with contextlib.redirect_stdout(io.StringIO()):
            scene = make_channel("C", 10, 30, 20, lattice=lat, **kw)
        if pipe is None:
            pipe = VolumePipeline(lat, [scene], tables, prof, 2.0)
            pipe.volume.set_reff_resampled(source_volume(15))
        pipe.scenes = [scene]
        flux = pipe.step().cpu().numpy()[0, 0]
        st = pipe.global_stats()[0]
        assert not st["overflow"] and st["tests"] == lat.n_points * 600
        results[name] = (st, pipe.global_hist.clone(), flux)
        scene.close()
    pipe.close()
    st0, h0, f0 = results["interval"]
    assert st0["passing_rays"] == 2341144738 and st0["visible_voxels"] == 106866897  # the counts of every run so far
    # rays whose angle lies within 1e-9 (units of 0.01 degree) of a rounding boundary - the only ones whose rounded angle
    # could differ from the reference's: 2.3e9 rays x 2e-9 = about five are expected, and found
    assert st0["near_rounding_rays"] < 50
    for name in ("per_ray_cull", "every_ray"):
        st, h, f = results[name]
        assert st["passing_rays"] == st0["passing_rays"] and st["visible_voxels"] == st0["visible_voxels"]
        assert st["candidates"] >= st["passing_rays"]
        torch.testing.assert_close(h, h0, rtol=1e-11, atol=0)
        np.testing.assert_allclose(f, f0, rtol=1e-12)
