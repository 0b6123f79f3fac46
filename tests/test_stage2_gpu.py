"""Stage-2 parity on the GPU: the Emissivity drop-in (CUDA path through the C ABI) against outputs of the
unmodified reference class (tests/golden/stage2_reference.npz) and against the CPU oracle."""
import os

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu

LINES = {"B": ("Z4", 48.6), "C": ("Z5", 33.7), "N": ("Z6", 24.8), "O": ("Z7", 19.0)}
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]
NE_SWEEP = [4e13, 0, 0.937, 1.8e12, 0.5, 0.11]
TE_SWEEP_LO = np.array([500, 0, 0.175, 30, 0.38, 0.07])
TE_SWEEP_HI = np.array([10000, 0, 0.175, 1500, 0.38, 0.07])
RTOL = 1e-12  # fp64 path; differences are rounding-order only (north-star bar: 1e-6 on intensities)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "stage2_reference.npz"))


def run(el, profile, tmp_path, time="Test"):
    from co_monitor_b200.emissivity.reader import Emissivity

    ion, wl = LINES[el]
    return Emissivity(el, ion, wl, 2, ["EXCIT", "RECOM"], "Reff_coordinates-10_mm", profile, time,
                      output_dir=tmp_path, verbose=False)


def check(em, gold, tag):
    df = em.df_prof_frac_ab_pec
    assert list(df.columns) == list(gold[f"{tag}_columns"])
    assert np.all(np.diff(df["Reff [m]"].to_numpy()) >= 0)  # sorted by Reff like the reference
    rows = df.sort_values("idx_sel_plas_points", kind="stable").to_numpy(dtype=float)
    stride = int(gold[f"{tag}_stride"][0])
    assert len(rows) == int(gold[f"{tag}_nrows"][0])
    ref = gold[f"{tag}_rows"]
    assert np.array_equal(rows[::stride, 0], ref[:, 0])
    np.testing.assert_allclose(rows[::stride], ref, rtol=RTOL, atol=0)
    np.testing.assert_allclose(rows.sum(axis=0), gold[f"{tag}_colsum"], rtol=1e-11)
    flux = np.array([em.flux["EXCIT"], em.flux["RECOM"], em.flux["TOTAL"]])
    np.testing.assert_allclose(flux, gold[f"{tag}_flux"], rtol=RTOL)
    assert em.total_emissivity == str(gold[f"{tag}_total_str"][0])
    assert em.reff_boundary == float(gold[f"{tag}_reff_boundary"][0])


@pytest.mark.parametrize("el", list("CBNO"))
def test_reference_parity_main_profile(gold, reff10, tmp_path, el):
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile

    em = run(el, TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df, tmp_path)
    check(em, gold, f"main_{el}")
    out = tmp_path / f"Emissivity ({el}_{LINES[el][0]}_Test).dat"
    assert out.exists()
    with open(out) as fh:
        head = [fh.readline() for _ in range(8)]
    assert head[0].startswith("# Index(['idx_sel_plas_points'") and any("Total emissivity: " + em.total_emissivity in h for h in head)
    saved = np.loadtxt(out)
    assert saved.shape == (len(em.df_prof_frac_ab_pec), 15)


def test_reference_parity_experimental_profiles(gold, reff10, tmp_path):
    from co_monitor_b200.emissivity.kinetic_profiles import ExperimentalProfile, TestExperimentalProfile

    check(run("C", ExperimentalProfile("report_20181011_012@5_5000_v_1").profiles_df, tmp_path, "exp"), gold, "exp_C")
    check(run("O", TestExperimentalProfile("20230215.058", "8.0105").profiles_df, tmp_path, "tst"), gold, "tst_O")


def test_sweep_slices_and_factorised_form(gold, reff10, tmp_path):
    """Per-slice Emissivity and the histogram + sweep kernels against the reference on slices 0, 499, 999."""
    import torch

    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile

    slices = (0, 499, 999)
    profs = [TwoGaussSumProfile(NE_SWEEP, list(TE_SWEEP_LO + (TE_SWEEP_HI - TE_SWEEP_LO) * (s / 999))).profiles_df
             for s in slices]
    for el in "CO":
        ems = [run(el, p, tmp_path, s) for s, p in zip(slices, profs)]
        for s, em in zip(slices, ems):
            check(em, gold, f"sweep{s}_{el}")
        tables = cached_line_tables(el, LINES[el][0], LINES[el][1], ("EXCIT", "RECOM"), str(ems[0]._input_root))
        base = ems[0].plas_coords_with_rad_frac
        hist, boundary = tables.weight_histogram(profs[0]["Reff [m]"].to_numpy(), base["Reff [m]"].to_numpy(),
                                                 base["total_intensity_fraction"].to_numpy())
        assert boundary == ems[0].reff_boundary
        batch = np.stack([p[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy() for p in profs])
        flux = tables.sweep(batch, hist, 0.02).cpu().numpy()
        for i, s in enumerate(slices):
            np.testing.assert_allclose(flux[i], gold[f"sweep{s}_{el}_flux"], rtol=1e-11)
        assert abs(float(hist.sum()) - ems[0].df_prof_frac_ab_pec["total_intensity_fraction"].sum()) < 1e-12


def test_oracle_parity_random_reff_and_unsorted_profile(tmp_path):
    """Synthetic Reff (random, with NaNs and exact ties) and a shuffled profile table: literal-argmin path."""
    from co_monitor_b200.emissivity.engine import cached_line_tables
    from oracle import stage2 as o2

    rng = np.random.default_rng(0)
    n = 5000
    reff = np.round(rng.random(n) * 0.6, 3)
    reff[rng.random(n) < 0.1] = np.nan
    reff[:50] = np.round(np.arange(50) * 0.0105, 4)  # x.xxx5 values: equidistant from two profile rows
    w = rng.random(n) * 1e-5
    obs = np.column_stack((np.arange(n), np.zeros((n, 3)), w))
    prof = o2.two_gauss_profile(NE_MAIN, TE_MAIN)
    tables = cached_line_tables("N", "Z6", 24.8, ("EXCIT", "RECOM"), None)
    otab = o2.load_tables("N", "Z6", 24.8, ["EXCIT", "RECOM"])
    for profile in (prof, prof[rng.permutation(len(prof))]):
        ref = o2.emissivity("N", "Z6", 24.8, 2, ["EXCIT", "RECOM"], profile, obs, reff, tables=otab)
        out = tables.integrate(profile, 0.02, reff, w)
        assert np.array_equal(np.flatnonzero(out["valid"]), ref.rows[:, 0].astype(int))
        np.testing.assert_allclose(out["columns"][out["valid"]], ref.rows[:, 6:], rtol=RTOL, atol=0)
        np.testing.assert_allclose(out["flux"], [ref.flux["EXCIT"], ref.flux["RECOM"], ref.flux["TOTAL"]], rtol=RTOL)
        assert out["boundary"] == ref.reff_boundary


@pytest.mark.parametrize("el,transitions", [("C", ("EXCIT",)), ("C", ("RECOM",)), ("C", ("EXCIT", "RECOM", "CHEXC")),
                                            ("O", ("RECOM", "EXCIT")), ("N", ("CHEXC",))])
def test_reference_parity_other_transition_lists(golden_dir, tmp_path, el, transitions):
    """The drop-in with single transitions, a swapped order and the charge-exchange block - columns, rows, flux and the
    total string of the unmodified reference (tests/golden/stage2_transitions_reference.npz)."""
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.emissivity.reader import Emissivity

    gold = np.load(os.path.join(golden_dir, "stage2_transitions_reference.npz"))
    tag = f"{el}_{'_'.join(transitions)}"
    ion, wl = LINES[el]
    em = Emissivity(el, ion, wl, 2, list(transitions), "Reff_coordinates-10_mm", TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df,
                    output_dir=tmp_path, verbose=False)
    df = em.df_prof_frac_ab_pec
    assert list(df.columns) == list(gold[f"{tag}_columns"])
    rows = df.sort_values("idx_sel_plas_points", kind="stable").to_numpy(dtype=float)
    assert len(rows) == int(gold[f"{tag}_nrows"][0])
    np.testing.assert_allclose(rows[::9], gold[f"{tag}_rows"], rtol=RTOL, atol=0)
    np.testing.assert_allclose(rows.sum(axis=0), gold[f"{tag}_colsum"], rtol=1e-11)
    flux = np.array([em.flux[t] for t in transitions] + [em.flux["TOTAL"]])
    np.testing.assert_allclose(flux, gold[f"{tag}_flux"], rtol=RTOL)
    assert em.total_emissivity == str(gold[f"{tag}_total_str"][0])
    assert em.reff_boundary == float(gold[f"{tag}_reff_boundary"][0])


def test_find_nearest_reference_unit_test():
    from co_monitor_b200.emissivity.reader import Emissivity

    nrs = [1, 22, 50, 16, 7, 19, 100, 4, 6, 211]
    assert Emissivity._find_nearest(nrs, 75) == 2
    assert Emissivity._find_nearest(nrs, 76) == 6
    assert Emissivity._find_nearest(nrs, 0) == 0


def test_device_pipeline_matches_reference_flux(gold, reff10):
    """Stage 1 -> compaction -> Stage 2 without leaving the device (co_monitor_b200.pipeline) reproduces the
    reference's per-channel flux (computed there from the shipped Stage-1 files)."""
    import contextlib
    import io

    import torch

    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.pipeline import LYMAN_ALPHA, ChannelPipeline

    with contextlib.redirect_stdout(io.StringIO()):
        lattice = CuboidMesh(10).lattice
    reff = torch.from_numpy(reff10).cuda()
    prof = TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy()
    for el in "CO":
        ion, wl = LYMAN_ALPHA[el]
        scene = make_channel(el, 10, 30, 20, lattice=lattice)
        tables = cached_line_tables(el, ion, wl, ("EXCIT", "RECOM"), None)
        for mode in ("histogram", "per_voxel"):
            pipe = ChannelPipeline(scene, tables, lattice, reff, 0, lattice.n_points, 2.0, stage2=mode)
            pipe.set_profile(prof)
            flux = pipe.launch().cpu().numpy()
            assert not pipe.check()["overflow"]
            np.testing.assert_allclose(flux, gold[f"main_{el}_flux"], rtol=1e-11)
            if mode == "histogram":
                assert float(pipe.boundary.item()) == float(gold[f"main_{el}_reff_boundary"][0])
            # sharded: two half-ranges with the exchange in between (what N = 2 ranks do over NCCL)
            halves = [ChannelPipeline(scene, tables, lattice, reff, a, b, 2.0, stage2=mode)
                      for a, b in ((0, 135000), (135000, 270000))]
            for h in halves:
                h.set_profile(prof)
                h.launch_geometry()
            if mode == "histogram":  # all-reduce(SUM) of the millimetre histograms: every rank then holds the whole flux
                total_hist = halves[0].hist_mm + halves[1].hist_mm
                for h in halves:
                    h.hist_mm.copy_(total_hist)
                    h.launch_emissivity()
                    np.testing.assert_allclose(h.flux.cpu().numpy(), flux, rtol=1e-12)
            else:  # all-reduce(MAX) of the mapped Reff, then all-reduce(SUM) of the partial fluxes
                m = torch.maximum(halves[0].mapped_max, halves[1].mapped_max)
                total = np.zeros(3)
                for h in halves:
                    h.mapped_max.copy_(m)
                    h.launch_emissivity()
                    total += h.flux.cpu().numpy()
                np.testing.assert_allclose(total, flux, rtol=1e-12)
        scene.close()


def test_volume_pipeline_on_the_fly_histogram_matches_reference_flux(gold):
    """The whole-volume pipeline of scripts/volume_flux.py (config 5) at the 10 mm grid: Reff bins from the device
    resampling kernel, Stage 1 in chunks adding every passing ray to the millimetre histogram on the fly, one histogram
    per shard summed as the all-reduce would, Stage 2 from the histogram - against the reference's flux."""
    import contextlib
    import io

    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import _ptr, cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import resample_reff_device
    from co_monitor_b200.pipeline import LYMAN_ALPHA

    lib = _lib.load()
    with contextlib.redirect_stdout(io.StringIO()):
        full = CuboidMesh(10).lattice
    prof = np.ascontiguousarray(TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    prof_t = torch.from_numpy(prof).cuda()
    K = len(prof)
    for el in "BN":
        ion, wl = LYMAN_ALPHA[el]
        scene = make_channel(el, 10, 30, 20, lattice=full)
        tables = cached_line_tables(el, ion, wl, ("EXCIT", "RECOM"), None)
        for world in (1, 3):
            total = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device="cuda")
            rays = 0
            for rank in range(world):
                lat = full.shard(world, rank, cols_per_block=8)
                reff_mm, _ = resample_reff_device(lat, 15)
                hist = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device="cuda")
                chunk = 40_000  # several launches per shard, the last one ragged
                plan = scene.plan(lat, 0, chunk, reff_bin=reff_mm, hist=hist)
                for lo in range(0, lat.n_points, chunk):
                    plan.rebind(lo, min(lat.n_points, lo + chunk)).launch()
                    st = plan.stats()
                    assert not st["overflow"]
                    rays += st["passing_rays"]
                total += hist
            assert rays == int(gold[f"stage1_{el}_passing_rays"][0]) if f"stage1_{el}_passing_rays" in gold else rays > 0
            rows = torch.zeros(K, dtype=torch.float64, device="cuda")
            bnd = torch.zeros(1, dtype=torch.float64, device="cuda")
            flux = torch.zeros(3, dtype=torch.float64, device="cuda")
            _lib.check(lib.com_histogram_rows(_ptr(prof_t), K, 1, float(prof[:, 0].max()), _ptr(total), None, _ptr(rows),
                                              _ptr(bnd), None), "com_histogram_rows")
            _lib.check(lib.com_emissivity_sweep(tables.handle, _ptr(prof_t), 1, K, _ptr(rows), 0.02, _ptr(flux), None),
                       "com_emissivity_sweep")
            np.testing.assert_allclose(flux.cpu().numpy(), gold[f"main_{el}_flux"], rtol=1e-11)
            assert float(bnd.item()) == float(gold[f"main_{el}_reff_boundary"][0])
        scene.close()


def test_host_buffer_emissivity_call(gold, reff10):
    """com_emissivity_integrate_host (host in, host out) against the reference's C-channel result."""
    import ctypes as C

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from oracle import stage2 as o2

    lib = _lib.load()
    obs = o2.read_observed_volume("C")
    reff = reff10[obs[:, 0].astype(np.int64)]
    keep = ~np.isnan(reff)
    reff, w = np.ascontiguousarray(reff[keep]), np.ascontiguousarray(obs[keep, 4])
    prof = np.ascontiguousarray(TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    tables = cached_line_tables("C", "Z5", 33.7, ("EXCIT", "RECOM"), None)
    n = len(reff)
    flux = np.zeros(3)
    cols = np.zeros((n, 9))
    valid = np.zeros(n, dtype=np.uint8)
    bnd = C.c_double()
    for _ in range(2):
        rc = lib.com_emissivity_integrate_host(C.byref(tables.desc), _lib.as_double_ptr(prof), len(prof), 0.02,
                                               _lib.as_double_ptr(reff), _lib.as_double_ptr(w), n, _lib.as_double_ptr(flux),
                                               C.c_void_p(cols.ctypes.data), C.c_void_p(valid.ctypes.data), C.byref(bnd))
        assert rc == 0
        np.testing.assert_allclose(flux, gold["main_C_flux"], rtol=RTOL)
        assert bnd.value == float(gold["main_C_reff_boundary"][0])
        assert valid.sum() == int(gold["main_C_nrows"][0])
    # the indexed variant does the Reff join inside: same results from the whole Reff table + the voxel indices
    idx_all = np.ascontiguousarray(obs[:, 0].astype(np.int64))
    w_all = np.ascontiguousarray(obs[:, 4])
    flux2, bnd2 = np.zeros(3), C.c_double()
    valid2 = np.zeros(len(idx_all), dtype=np.uint8)
    table = np.ascontiguousarray(reff10)
    rc = lib.com_emissivity_integrate_indexed_host(C.byref(tables.desc), _lib.as_double_ptr(prof), len(prof), 0.02,
                                                   _lib.as_double_ptr(table), len(table), C.c_void_p(idx_all.ctypes.data),
                                                   _lib.as_double_ptr(w_all), len(idx_all), _lib.as_double_ptr(flux2), None,
                                                   C.c_void_p(valid2.ctypes.data), C.byref(bnd2))
    assert rc == 0
    np.testing.assert_allclose(flux2, flux, rtol=1e-13)  # NaN rows are dropped inside; only the summation order differs
    assert bnd2.value == bnd.value and valid2.sum() == valid.sum()
    bad = idx_all.copy()
    bad[5] = len(table)
    assert lib.com_emissivity_integrate_indexed_host(C.byref(tables.desc), _lib.as_double_ptr(prof), len(prof), 0.02,
                                                     _lib.as_double_ptr(table), len(table), C.c_void_p(bad.ctypes.data),
                                                     _lib.as_double_ptr(w_all), len(bad), _lib.as_double_ptr(flux2), None, None,
                                                     None) == _lib.COM_E_BADARG
    ref_rows = gold["main_C_rows"]  # stride 1 for C, sorted by idx
    mine = cols[valid.astype(bool)]
    order = np.argsort(obs[keep][valid.astype(bool), 0], kind="stable")
    np.testing.assert_allclose(mine[order], ref_rows[:, 6:], rtol=RTOL, atol=0)


def test_sweep_accepts_any_tensor_layout(reff10):
    """np.stack of DataFrame.to_numpy() blocks is not C-ordered; device tensors made from it must still be read right."""
    import torch

    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from oracle import stage2 as o2

    obs = o2.read_observed_volume("O")
    reff = reff10[obs[:, 0].astype(np.int64)]
    keep = ~np.isnan(reff)
    reff, w = reff[keep], obs[keep, 4]
    tables = cached_line_tables("O", "Z7", 19.0, ("EXCIT", "RECOM"), None)
    profs = np.stack([TwoGaussSumProfile(NE_SWEEP, list(TE_SWEEP_LO + (TE_SWEEP_HI - TE_SWEEP_LO) * (s / 19))).profiles_df[
        ["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy() for s in range(20)])
    assert not profs.flags["C_CONTIGUOUS"]
    hist, _ = tables.weight_histogram(profs[0][:, 0], torch.from_numpy(reff).cuda(), torch.from_numpy(w).cuda())
    a = tables.sweep(profs, hist, 0.02).cpu().numpy()
    b = tables.sweep(torch.from_numpy(profs).cuda(), hist, 0.02).cpu().numpy()
    assert np.array_equal(a, b)
    for s in (0, 7, 19):
        naive = tables.integrate(profs[s], 0.02, reff, w, per_voxel=False)["flux"]
        np.testing.assert_allclose(a[s], naive, rtol=1e-11)
    assert a[0, 2] != a[19, 2]


def test_millimetre_bin_histogram_fast_path():
    """uint16-millimetre Reff bins (fast 10 B/voxel pass) against the fp64-Reff path and a NumPy histogram:
    odd length (scalar tail), 0xFFFF = outside the LCFS, zero weights = invisible voxels, values beyond the cut."""
    import torch

    from co_monitor_b200.emissivity.engine import cached_line_tables

    rng = np.random.default_rng(5)
    n = 200_003
    mm = rng.integers(0, 700, n).astype(np.uint16)
    mm[rng.random(n) < 0.3] = 0xFFFF
    w = rng.random(n) * 1e-6
    w[rng.random(n) < 0.8] = 0.0
    reff = np.where(mm == 0xFFFF, np.nan, mm / 1000.0)
    prof_reff = np.arange(0, 0.539, 0.001)
    tables = cached_line_tables("C", "Z5", 33.7, ("EXCIT", "RECOM"), None)
    mm_d = torch.from_numpy(mm.view(np.int16)).cuda()
    fast, b_fast = tables.weight_histogram(prof_reff, None, w, reff_mm=mm_d)
    vis = w != 0  # the fp64 path takes the maximum over every voxel it is given: give it the visible ones, like the pipeline
    slow, b_slow = tables.weight_histogram(prof_reff, reff[vis], w[vis])
    assert b_fast == b_slow == 0.538
    np.testing.assert_allclose(fast.cpu().numpy(), slow.cpu().numpy(), rtol=1e-12, atol=1e-300)
    keep = vis & (mm != 0xFFFF) & (mm / 1000.0 < b_fast)
    ref = np.bincount(mm[keep].astype(np.int64), weights=w[keep], minlength=539)[:539]
    np.testing.assert_allclose(fast.cpu().numpy(), ref, rtol=1e-12, atol=1e-300)


def test_loglog_pec_extension(tmp_path):
    """``pec_interpolation="loglog"`` (off by default, not reference behaviour): the kernel against its NumPy statement,
    through the sweep entry point (flux and per-row arithmetic) and through the Emissivity drop-in."""
    import torch

    from co_monitor_b200.emissivity.engine import LineTables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from oracle.extensions import loglog_pec as loglog_on_nodes
    from co_monitor_b200.emissivity.reader import Emissivity

    prof_df = TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df
    prof = prof_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy()
    tables = LineTables("C", "Z5", 33.7, ["EXCIT", "RECOM"], pec_interpolation="loglog")
    K = len(prof)
    rng = np.random.default_rng(3)
    hist = rng.uniform(0, 1e-6, K)
    flux = tables.sweep(prof[None], torch.from_numpy(hist).cuda(), 0.02).cpu().numpy()[0]
    te, ne = prof[:, 1], prof[:, 2]
    fa = tables.fa.df_interpolated_frac_ab
    j = np.abs(fa.iloc[:, 0].to_numpy()[None, :] - te[:, None]).argmin(axis=1)
    fa_ion, fa_last = fa["Z5"].to_numpy()[j], fa.iloc[:, -1].to_numpy()[j]
    want = []
    for pec, f in zip(tables.pecs, (fa_ion, fa_last)):
        val = loglog_on_nodes(pec.ne_nodes, pec.te_nodes, pec.pec_data_array, ne, te)
        want.append(float((f * val * ne**2 * 0.02 * hist).sum()))
    np.testing.assert_allclose(flux[:2], want, rtol=1e-12)
    np.testing.assert_allclose(flux[2], want[0] + want[1], rtol=1e-12)
    # the reference interpolation reads the nearest node of a linear table: same order of magnitude, not the same number
    ref_tables = LineTables("C", "Z5", 33.7, ["EXCIT", "RECOM"])
    ref_flux = ref_tables.sweep(prof[None], torch.from_numpy(hist).cuda(), 0.02).cpu().numpy()[0]
    assert 0.5 < flux[2] / ref_flux[2] < 2.0 and flux[2] != ref_flux[2]
    em = Emissivity("C", "Z5", 33.7, 2, ["EXCIT", "RECOM"], "Reff_coordinates-10_mm", prof_df, output_dir=tmp_path,
                    verbose=False, pec_interpolation="loglog")
    df = em.df_prof_frac_ab_pec
    np.testing.assert_allclose(df["pec_EXCIT"].to_numpy(),
                               loglog_on_nodes(tables.pecs[0].ne_nodes, tables.pecs[0].te_nodes, tables.pecs[0].pec_data_array,
                                               df["n_e [m-3]"].to_numpy(), df["T_e [eV]"].to_numpy()), rtol=1e-12)
    with pytest.raises(ValueError):
        LineTables("C", "Z5", 33.7, ["EXCIT"], pec_interpolation="cubic")
