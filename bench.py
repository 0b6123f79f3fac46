#!/usr/bin/env python
"""Benchmark of the co_monitor hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload = BASELINE.json configs[4]: the dense 0.65 mm observation volume (2076 x 1230 x 384 = 9.8e8 voxels)
x 600 crystal points x 10 slits for all four channels B/C/N/O = 2.35e12 voxel-detector visibility tests per step,
STRONG scaling: the z-columns of the lattice are dealt block-cyclically over the N ranks.  A step is the whole path:

    Stage 1 (visibility + weight) of the rank's shard, launch by launch, each launch folded into the channel's
    Reff-millimetre histogram  ->  the single exchange of the packed [4 x 4096 bins | counters] buffers, fused into
    Stage 2: after one device-side barrier the Stage-2 launch itself reads the ranks' buffers over NVLink (symmetric
    memory), sums them and evaluates EXCIT + RECOM emissivity -> flux (NCCL all-reduce + Stage 2 when symmetric memory
    cannot be set up, and as the timed comparison)

One JSON line (rank 0):
  value            tests/s with every ray's detector test evaluated individually (COM_FILTER_NO_RUN_CULL; SURVEY.md 8(d)
                   wants the rate without culling as the headline), CUDA events, max over ranks, inputs resident in HBM
  production       the same step as shipped (runs of voxels proven dark by the linearity of the test are skipped):
                   rate, ms per step and `flux_time_s` = seconds from nothing to the four channels' flux
  e2e              the same two through the host-buffer C-ABI calls (com_volume_histograms_host: Reff source grid up,
                   histograms + counters down; exchange; com_histograms_flux_host: profile + histograms up, flux down)
  small_grid       configs[1] (10 mm lattice, the golden-parity workload), device step and host-buffer calls
  grid_1mm         configs[2] lattice (1 mm) rate;  sweep_1000  configs[3]: 1000 profile slices from cached weights
  roofline         dominant kernel of the headline step (K1a FP32 every-ray filter) against the issue peak, instruction counts
                   and DRAM bytes MEASURED in this run by an `ncu --metrics` pass over one channel (counters only - never
                   times); roofline_production: the same for the shipped step's interval / weight-only / exact kernels
  explicit_voxels  Stage 1 on explicit (P,3) fp64 voxels with the packed float4 copy the filter streams by cp.async.bulk
  roofline_stage2  the Stage-2 voxel pass (10 B/voxel) on this run's real weights against the measured HBM peak
  cpu_baseline     the NumPy/Qhull oracle on a bounded sample of the same lattice on this box's host cores
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes as C
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ELEMENTS = "BCNO"
H, L, SLITS = 30, 20, 10
METRIC = "voxel-detector visibility tests/s"
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]  # reference __main__.py:35-36
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]
NE_SWEEP = [4e13, 0, 0.937, 1.8e12, 0.5, 0.11]  # emissivity/main.py:39-45
TE_SWEEP_LO = np.array([500, 0, 0.175, 30, 0.38, 0.07])
TE_SWEEP_HI = np.array([10000, 0, 0.175, 1500, 0.38, 0.07])
#: (ix, iy) of each channel's brightest voxel on the 10 mm lattice: where the CPU sample of a fine lattice is taken
BRIGHT_10MM = {"B": (10, 25), "C": (3, 26), "N": (9, 32), "O": (8, 25)}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def workload_name(spacing: float) -> str:
    tag = {0.65: "configs[4]", 1.0: "configs[2]", 10.0: "configs[1]"}.get(float(spacing), "custom")
    return (f"{tag}: dense {spacing} mm observation volume, channels B/C/N/O, Stage 1 (visibility + weight) -> Reff-mm "
            f"histograms -> one reduce -> Stage 2 (EXCIT + RECOM flux); {H}x{L} crystal points, {SLITS} slits")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.thread.join(timeout=2)
        rows = [r for r in self.rows if len(r) >= 9]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in rows if r[3].replace(".", "").isdigit()]
        loaded = [s for s, p in zip(sm, pw) if p > 300.0]  # samples while the GPU was drawing power
        busy = loaded or sm
        reasons = set()
        for r in rows:
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_under_load": len(loaded),
                "power_w_max": max(pw) if pw else None}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference on a bounded sample of the same lattice
def reference_step(element, spacing, voxels_per_core, reff_vol):
    """One bounded sample on the CPU path: Stage 1 (oracle, all host cores) on a range of consecutive voxels through the
    channel's bright region, then Stage 2 (oracle) on the visible voxels of that range.  The reference rebuilds its PEC /
    FA tables in every Emissivity call (reader.py:83-97, 245-261): that fixed cost is counted pro rata (range / lattice)."""
    from co_monitor_b200.pipeline import LYMAN_ALPHA
    from oracle import geometry_setup as gs
    from oracle import reff as oreff
    from oracle import stage1 as o1
    from oracle import stage2 as o2

    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    nx, ny, nz = gs.mesh_shape(gs.load_db(), spacing)
    ix10, iy10 = BRIGHT_10MM[element]
    col = int(round(iy10 * (ny - 1) / 79)) * nx + int(round(ix10 * (nx - 1) / 134))
    n = voxels_per_core * procs
    lo = max(0, min(col * nz - n // 2, nx * ny * nz - n))
    hi = lo + n
    t0 = time.perf_counter()
    w, nr, used = o1.run_range(element, lo, hi, S=SLITS, spacing=spacing, h=H, l=L, processes=procs, chunk=200)
    t1 = time.perf_counter() - t0
    ion, wl = LYMAN_ALPHA[element]
    t0 = time.perf_counter()
    tables = o2.load_tables(element, ion, wl, ["EXCIT", "RECOM"])
    t_tables = time.perf_counter() - t0
    vis = np.flatnonzero(nr > 0)
    t0 = time.perf_counter()
    flux = None
    if len(vis):
        reff = oreff.on_lattice((nx, ny, nz), vis + lo, vol=reff_vol)
        if np.isfinite(reff).any():
            obs = np.column_stack((np.arange(len(vis)), np.zeros((len(vis), 3)), w[vis]))
            flux = o2.emissivity(element, ion, wl, 2, ["EXCIT", "RECOM"], o2.two_gauss_profile(NE_MAIN, TE_MAIN), obs, reff,
                                 tables=tables).flux["TOTAL"]
    t2 = time.perf_counter() - t0
    frac = n / float(nx * ny * nz)
    seconds = t1 + t2 + t_tables * frac
    tests = n * H * L
    return tests / seconds, {"element": element, "voxels": [lo, hi], "tests": tests, "seconds": seconds, "stage1_s": t1,
                              "stage2_s": t2, "tables_s_prorated": t_tables * frac, "processes": used,
                              "cores_visible": cores, "passing_rays": int(nr.sum()), "flux_total": flux}


def cpu_baseline(spacing):
    from oracle import reff as oreff

    value, d = reference_step("C", spacing, 800, oreff.shipped_volume(15))
    return {"value": value, "unit": "tests/s", "cores": d["processes"], "kind": "port",
            "sample": f"C channel, voxels {d['voxels']} of the {spacing} mm lattice x {H * L} crystal points x 22 apertures "
                      f"through the NumPy/Qhull oracle ({d['stage1_s']:.1f} s on {d['processes']} processes of "
                      f"{d['cores_visible']} cores, {d['passing_rays']} passing rays) + Stage-2 oracle on the visible voxels "
                      f"of that range ({d['stage2_s']:.2f} s + {d['tables_s_prorated']:.3f} s of table building pro rata)"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores.  The reference itself needs
    dask / opt_einsum / pyvista (absent, no network), so this is the oracle port of it - the NumPy + Qhull restatement
    that reproduces the reference's shipped outputs exactly.  Each step is a bounded sample of the workload (one channel
    per step, B/C/N/O in turn, a range of consecutive voxels of the same lattice)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import reff as oreff

    vol = oreff.shipped_volume(15)
    per_core = 400 if args.steps + args.warmup > 6 else 800
    vals, details = [], []
    for i in range(args.warmup + args.steps):
        v, d = reference_step(ELEMENTS[i % len(ELEMENTS)], args.spacing, per_core, vol)
        if i >= args.warmup:
            vals.append(v)
            details.append(d)
    tests = float(np.mean([d["tests"] for d in details]))
    seconds = float(np.mean([d["seconds"] for d in details]))
    value = tests / seconds
    sample = (f"per step: one channel (B/C/N/O in turn), {details[-1]['voxels'][1] - details[-1]['voxels'][0]} consecutive "
              f"voxels of the {args.spacing} mm lattice through the channel's bright region x {H * L} crystal points x 22 "
              f"apertures through the NumPy/Qhull oracle on {details[-1]['processes']} processes, then the Stage-2 oracle "
              f"on the visible voxels of that range; last step: {details[-1]}")
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tests/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * seconds, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.spacing), "tests_per_step": tests,
                   "sample": "each step is a bounded sample of the workload (the oracle's cost per test does not depend on "
                             "the data: every ray goes through all 22 apertures); ms_per_step is the measured time of a "
                             "sampled step, value = its tests / its seconds"},
        "cpu_baseline": {"value": value, "unit": "tests/s", "cores": details[-1]["processes"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------------
KERNEL_REGEX = "regex:filter_kernel|interval_kernel_lattice|expand_kernel|weight_kernel|exact_kernel|weight_histogram_mm"


def ncu_counters(spacing, element="C", mode="production", timeout_s=300):
    """Instruction counts and DRAM bytes of one channel's whole-volume Stage 1 + histogram pass, measured now by an
    `ncu --metrics` pass over `bench.py --probe` in a subprocess (hardware counters only; no time from it is used).
    *mode*: "nocull" = the headline step (every ray's detector test evaluated), "production" = the shipped step."""
    import shutil

    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return {"unavailable": "ncu not found"}
    metrics = "smsp__thread_inst_executed.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum"
    cmd = [ncu, "--csv", "--metrics", metrics, "--clock-control", "none", "-k", KERNEL_REGEX, sys.executable,
           os.path.abspath(__file__), "--probe", element, "--probe-mode", mode, "--spacing", str(spacing)]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
    except (subprocess.TimeoutExpired, OSError) as exc:
        return {"unavailable": f"ncu pass failed: {type(exc).__name__}"}
    import csv

    rows = [r for r in csv.reader(proc.stdout.splitlines()) if len(r) > 10]
    if not rows or "Kernel Name" not in rows[0]:
        return {"unavailable": "no ncu output (profiling may be closed on this box)", "rc": proc.returncode,
                "tail": (proc.stdout + proc.stderr)[-300:]}
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("com::", "")
        try:
            val = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        d = agg.setdefault(name, {"launches": 0})
        d[r[im]] = d.get(r[im], 0.0) + val
        if r[im] == "smsp__inst_executed.sum":
            d["launches"] += 1
    probe = None
    for line in proc.stdout.splitlines():
        if line.startswith("{\"probe\""):
            probe = json.loads(line)
    if probe is None:
        return {"unavailable": "the probe printed no summary line", "rc": proc.returncode, "tail": (proc.stdout + proc.stderr)[-300:]}
    return {"kernels": agg, "probe": probe,
            "command": f"ncu --csv --metrics {metrics} --clock-control none -k {KERNEL_REGEX} python bench.py --probe {element} "
                       f"--probe-mode {mode} --spacing {spacing}"}


def run_probe(args):
    """One channel's whole-volume Stage 1 + histogram pass, once (the workload the ncu counter pass profiles)."""
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import cached_line_tables
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import source_volume
    from co_monitor_b200.pipeline import LYMAN_ALPHA
    from co_monitor_b200.volume import ObservationVolume

    el = args.probe
    lat = quiet(CuboidMesh, args.spacing).lattice
    scene = quiet(make_channel, el, SLITS, H, L, lattice=lat, run_cull=args.probe_mode != "nocull")
    cached_line_tables(el, *LYMAN_ALPHA[el], ("EXCIT", "RECOM"), None)
    vol = ObservationVolume(lat, args.chunk)
    vol.set_reff_resampled(source_volume(15))
    hist, stats = vol.histograms([scene])
    torch.cuda.synchronize()
    names = [f[0] for f in _lib.ComStats._fields_]
    st = dict(zip(names, (int(x) for x in stats[0].cpu().numpy())))
    print(json.dumps({"probe": el, "mode": args.probe_mode, "voxels": lat.n_points, "tests": lat.n_points * H * L,
                      "launches": vol.n_launches, "passing_rays": st["passing_rays"], "candidates": st["candidates"],
                      "certain_candidates": st["certain_candidates"]}))


# ---------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import LineTables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import lattice_struct, make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import source_volume
    from co_monitor_b200.pipeline import LYMAN_ALPHA
    from co_monitor_b200.volume import VolumePipeline, flux_from_histograms_host

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the accelerated path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import datetime

        # a rank that dies must not leave the others waiting for ten minutes of box time
        dist.init_process_group("nccl", device_id=device, timeout=datetime.timedelta(seconds=300))
    lib = _lib.load()
    warmup = max(args.warmup, 3)
    steps = args.steps

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- the workload ----
    full = quiet(CuboidMesh, args.spacing).lattice
    lat = full.shard(world, rank, cols_per_block=args.cols_per_block)
    n_ch = len(ELEMENTS)
    tests_per_step = full.n_points_global * H * L * n_ch
    profile = np.ascontiguousarray(TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy())
    scenes = [quiet(make_channel, el, SLITS, H, L, lattice=full) for el in ELEMENTS]
    scenes_nc = [quiet(make_channel, el, SLITS, H, L, lattice=full, run_cull=False) for el in ELEMENTS]
    tables = [quiet(LineTables, el, *LYMAN_ALPHA[el], ["EXCIT", "RECOM"]) for el in ELEMENTS]
    src = source_volume(15)  # the shipped 15 mm VMEC grid (host), resampled onto the lattice on the device
    pipe = VolumePipeline(lat, scenes, tables, profile, 2.0, chunk=args.chunk, device=device, exchange=args.exchange)
    pipe.volume.set_reff_resampled(src)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    def timed(run, n_warm, n_steps):
        """n_warm untimed + n_steps timed steps: barrier + sync on both sides, one CUDA-event interval per step, L2
        flushed (untimed) between steps; returns the max over ranks of the summed intervals in ms."""
        for _ in range(n_warm):
            flush.fill_(1)
            run()
        barrier()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_steps)]
        t0 = time.perf_counter()
        for i in range(n_steps):
            flush.fill_(i & 0xFF)
            e0[i].record()
            run()
            e1[i].record()
        barrier()
        wall = time.perf_counter() - t0
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in zip(e0, e1))), wall

    def check_overflow(what):
        bad = [ELEMENTS[i] for i, st in enumerate(pipe.global_stats()) if st["overflow"]]
        if bad:
            raise SystemExit(f"bench.py: candidate-list overflow in channel(s) {bad} during {what}; raise --candidate-fraction")

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- headline: every ray's detector test evaluated individually ----
    pipe.scenes = scenes_nc
    ms_nc, wall_nc = timed(pipe.step, warmup, steps)
    check_overflow("the headline steps")
    flux_nc = pipe.flux.cpu().numpy()[:, 0]
    # ---- the step as shipped ----
    pipe.scenes = scenes
    ms_prod, wall_prod = timed(pipe.step, warmup, steps)
    check_overflow("the production steps")
    flux_prod = pipe.flux.cpu().numpy()[:, 0]
    stats = pipe.global_stats()
    boundary = pipe.boundary.cpu().numpy()

    # ---- per-kernel times (events inside the library around every Stage-1 kernel, this rank's shard) ----
    def kernel_times(scene_list):
        """Stage 1 (+ histogram passes) of the given channels once more with the library's kernel events on: milliseconds
        of the FP32 filter kernel, the fp64 weight-only kernel and the fp64 exact kernel, and of the whole Stage 1."""
        lib.com_kernel_timing_enable(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        pipe.volume.histograms(scene_list, pipe.hist[: len(scene_list)], pipe.stats[: len(scene_list)])
        b.record()
        torch.cuda.synchronize()
        ms3, n_l = (C.c_double * 3)(), C.c_int64()
        lib.com_kernel_timing_read3(ms3, C.byref(n_l))
        lib.com_kernel_timing_enable(0)
        return {"filter_ms": ms3[0], "weight_ms": ms3[1], "exact_ms": ms3[2], "launches": int(n_l.value),
                "stage1_ms": a.elapsed_time(b)}

    kt_nc = kernel_times(scenes_nc)          # headline step, four channels
    kt_prod = kernel_times(scenes)           # production step, four channels
    kt_nc_c = kernel_times([scenes_nc[1]])   # channel C alone: the channel the ncu counter pass profiles
    kt_prod_c = kernel_times([scenes[1]])
    pipe.stage1()  # leave the pipeline's buffers as the production step left them

    # ---- the exchange + Stage 2 alone, both transports (N > 1): NCCL all-reduce + Stage 2 vs the fused peer-memory launch ----
    exchange_ms = None
    if world > 1:
        from co_monitor_b200.volume import flux_from_histograms, flux_from_peer_buffers

        def time_loop(body, reps=50):
            for _ in range(5):
                body()
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                body()
            b.record()
            barrier()
            return max_over_ranks(a.elapsed_time(b)) / reps

        tmp = pipe.packed.clone()
        tmp_hist = tmp[: n_ch * _lib.COM_MM_BINS].view(n_ch, _lib.COM_MM_BINS)
        fl_t, bd_t = torch.zeros_like(pipe.flux), torch.zeros_like(pipe.boundary)

        def nccl_body():
            tmp.copy_(pipe.packed)
            dist.all_reduce(tmp)
            flux_from_histograms(tables, pipe.profiles, tmp_hist, pipe.conc, fl_t, bd_t, pipe.meta)

        exchange_ms = {"nccl_all_reduce_then_stage2_ms": time_loop(nccl_body),
                       "what": "the step's tail alone, 50 repetitions, CUDA events, max over ranks: (a) copy + NCCL all-reduce of "
                               "the packed buffer + the Stage-2 launch, (b) device-side barrier + ONE launch that reads the "
                               "ranks' buffers over NVLink, sums them and evaluates Stage 2"}
        flux_nccl = fl_t.cpu().numpy()[:, 0].copy()
        if pipe.exchange_mode == "peer":
            def peer_body():
                pipe._symm.barrier(channel=0)
                flux_from_peer_buffers(tables, pipe.profiles, pipe._peer_ptrs[pipe._slot], _lib.N_STATS, pipe.conc, fl_t, bd_t,
                                       None, None, pipe.meta)

            exchange_ms["peer_barrier_then_fused_launch_ms"] = time_loop(peer_body)
            exchange_ms["flux_rel_diff_peer_vs_nccl"] = float(np.max(np.abs(fl_t.cpu().numpy()[:, 0] / flux_nccl - 1)))

    # ---- e2e: the host-buffer C-ABI calls, exchange included ----
    packed_h = torch.empty(n_ch * (_lib.COM_MM_BINS + _lib.N_STATS), dtype=torch.float64).pin_memory()
    packed_np = packed_h.numpy()
    hist_np = packed_np[: n_ch * _lib.COM_MM_BINS].reshape(n_ch, _lib.COM_MM_BINS)
    flux_e2e = np.zeros((n_ch, 3))
    io_bytes = {"h2d": 0, "d2h": 0}

    def e2e_step(sc):
        hist, st = pipe.volume.histograms_host(sc, reff_source=src, hist_out=hist_np)
        if world > 1:  # the path's single exchange, from and to host memory
            raw = np.array([[s[f[0]] for f in _lib.ComStats._fields_] for s in st], dtype=np.float64)
            packed_np[n_ch * _lib.COM_MM_BINS:] = raw.ravel()
            dev_buf = packed_h.to(device, non_blocking=True)
            dist.all_reduce(dev_buf)
            packed_h.copy_(dev_buf)  # synchronous D2H into pinned memory
        fl, _ = flux_from_histograms_host(tables, profile, hist_np, 0.02, device=device)
        flux_e2e[:] = fl[:, 0]

    def e2e_timed(sc, n_warm, n_steps):
        for _ in range(n_warm):
            e2e_step(sc)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            e2e_step(sc)
        barrier()
        return max_over_ranks(time.perf_counter() - t0)

    e2e_nc_s = e2e_timed(scenes_nc, 1, steps)
    flux_e2e_nc = flux_e2e.copy()
    e2e_prod_s = e2e_timed(scenes, 1, steps)
    flux_e2e_prod = flux_e2e.copy()
    exch = 2 * packed_np.nbytes if world > 1 else 0
    io_bytes["h2d"] = src[0].nbytes + profile.nbytes + hist_np.nbytes + exch // 2
    io_bytes["d2h"] = hist_np.nbytes + n_ch * C.sizeof(_lib.ComStats) + flux_e2e.nbytes + 8 * n_ch + exch // 2
    clocks = sampler.stop() if rank == 0 else None

    def teardown():
        import gc

        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        pipe.close()
        teardown()
        return

    # =========================== rank 0 only from here: sub-records, rooflines, baselines ===========================
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    issue_peak = 148 * 4 * 32 * sm_max * 1e6  # lane-instructions per second: 148 SMs x 4 schedulers x 32 lanes x f_max
    tests_c_rank = lat.n_points * H * L

    # ---- measured counters: one ncu --metrics pass per mode over channel C of this workload (counters only) ----
    cnt_nc = None if args.no_ncu else ncu_counters(args.spacing, "C", "nocull")
    cnt_prod = None if args.no_ncu else ncu_counters(args.spacing, "C", "production")
    khist = (cnt_prod or {}).get("kernels", {}).get("weight_histogram_mm_kernel")

    def kernel_record(counters, name, ms, per_unit_key=None, per_unit=None):
        """Issue-rate figures of one kernel: lane-instructions from the ncu pass (the unsharded channel C, scaled to this
        rank's shard) over this rank's live event time of the same kernel on the same channel."""
        k = (counters or {}).get("kernels", {}).get(name)
        if not k or ms <= 0:
            return None
        scale = tests_c_rank / counters["probe"]["tests"]
        lane, warp = k["smsp__thread_inst_executed.sum"] * scale, k["smsp__inst_executed.sum"] * scale
        rec = {"kernel": "com::" + name, "ms_channel_C": ms, "launches": k["launches"],
               "achieved": lane / (ms * 1e-3) / 1e12, "frac": lane / (ms * 1e-3) / issue_peak,
               "warp_issue_frac": warp * 32 / (ms * 1e-3) / issue_peak,
               "threads_per_warp_instr": k["smsp__thread_inst_executed.sum"] / k["smsp__inst_executed.sum"],
               "executed_lane_instr_per_test": k["smsp__thread_inst_executed.sum"] / counters["probe"]["tests"],
               "traffic": (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) / k["launches"]}
        if per_unit_key and per_unit:
            rec[per_unit_key] = k["smsp__thread_inst_executed.sum"] / per_unit
        return rec

    def shares(kt):
        tot = kt["stage1_ms"]
        return None if tot <= 0 else {
            "filter": kt["filter_ms"] / tot, "weight_only": kt["weight_ms"] / tot, "exact": kt["exact_ms"] / tot,
            "histogram_pass_memsets_other": 1 - (kt["filter_ms"] + kt["weight_ms"] + kt["exact_ms"]) / tot}

    n_launch = max(1, kt_nc["launches"])
    roofline = {
        "bound": "issue (FP32/ALU instruction issue; not HBM, not tensor)",
        "kernel": "com::filter_kernel_lattice (K1a, every ray's detector test evaluated: the kernel of the headline step)",
        "unit": "Tlane-inst/s", "peak": issue_peak / 1e12,
        "peak_source": f"148 SM x 4 schedulers x 32 lanes x {sm_max:.0f} MHz (MEASURED_PEAKS.json has no instruction-issue figure)",
        "kernel_ms_per_launch": kt_nc["filter_ms"] / n_launch, "launches_per_step_this_rank": int(n_launch),
        "kernel_share_of_step": kt_nc["filter_ms"] / kt_nc["stage1_ms"] if kt_nc["stage1_ms"] > 0 else None,
        "step_shares": shares(kt_nc),
        "timing": "CUDA events inside the library around every Stage-1 kernel of one headline step of this rank",
        "achieved": None, "frac": None, "traffic": None,
        "what": "achieved = lane-instructions the kernel executes (ncu smsp__thread_inst_executed.sum, measured in this run) / "
                "its event-timed duration; frac = achieved / peak.  warp_issue_frac counts every issued warp instruction as 32 "
                "lanes (issue-slot utilisation); the difference is lanes predicated off in divergent code.",
    }
    rec = kernel_record(cnt_nc, "filter_kernel_lattice", kt_nc_c["filter_ms"])
    if rec:
        rec.pop("kernel")
        roofline.update(rec)
        roofline["exact_kernel"] = kernel_record(cnt_nc, "exact_kernel", kt_nc_c["exact_ms"], "executed_lane_instr_per_candidate",
                                                 cnt_nc["probe"].get("candidates"))
        roofline["counters_command"] = cnt_nc["command"]
    else:
        # No counters in this run (--no-ncu, or ncu closed on this box): the instruction count per test is a property of
        # the kernel and the workload, not of the run - take it from the committed profile of the same kernel on the same
        # workload and say so; the time is this run's.
        roofline["counters"] = cnt_nc or {"unavailable": "--no-ncu"}
        ref_path = os.path.join(ROOT, "profiles", "r2y_bench_n1_default_run_final.json")
        try:
            with open(ref_path) as fh:
                ref_r = json.loads(fh.read().strip().splitlines()[-1])["roofline"]
            if args.spacing == 0.65 and kt_nc_c["filter_ms"] > 0 and ref_r.get("executed_lane_instr_per_test"):
                lane = ref_r["executed_lane_instr_per_test"] * tests_c_rank
                roofline.update({
                    "achieved": lane / (kt_nc_c["filter_ms"] * 1e-3) / 1e12,
                    "frac": lane / (kt_nc_c["filter_ms"] * 1e-3) / issue_peak,
                    "warp_issue_frac": lane / ref_r["threads_per_warp_instr"] * 32 / (kt_nc_c["filter_ms"] * 1e-3) / issue_peak,
                    "executed_lane_instr_per_test": ref_r["executed_lane_instr_per_test"],
                    "threads_per_warp_instr": ref_r["threads_per_warp_instr"], "ms_channel_C": kt_nc_c["filter_ms"],
                    "traffic": ref_r.get("traffic"),
                    "counters_source": "NOT measured in this run: instruction counts and DRAM bytes from "
                                       "profiles/r2y_bench_n1_default_run_final.json (ncu pass of that run, same kernel and "
                                       "workload); kernel time measured in this run"})
        except (OSError, KeyError, ValueError, IndexError):
            pass
    pc = (cnt_prod or {}).get("probe") or {}
    roofline_production = {
        "what": "the shipped step's Stage-1 kernels, same definitions as `roofline` (channel C; counters from a second ncu pass); "
                "the interval kernel's time includes expand_kernel (interval records -> per-ray list), its counters do not",
        "step_shares": shares(kt_prod), "stage1_ms_four_channels": kt_prod["stage1_ms"],
        "interval_kernel": kernel_record(cnt_prod, "interval_kernel_lattice", kt_prod_c["filter_ms"],
                                         "executed_lane_instr_per_candidate", pc.get("candidates")),
        "expand_kernel_lane_instr_per_candidate": ((cnt_prod or {}).get("kernels", {}).get("expand_kernel", {}).get(
            "smsp__thread_inst_executed.sum", 0.0) / pc["candidates"]) if pc.get("candidates") else None,
        "weight_kernel": kernel_record(cnt_prod, "weight_kernel_records", kt_prod_c["weight_ms"], "executed_lane_instr_per_ray",
                                       pc.get("certain_candidates")),
        "exact_kernel": kernel_record(cnt_prod, "exact_kernel", kt_prod_c["exact_ms"], "executed_lane_instr_per_candidate",
                                      (pc.get("candidates") or 0) - (pc.get("certain_candidates") or 0) or None),
        "counters_command": (cnt_prod or {}).get("command"),
    }
    if cnt_prod and "kernels" not in cnt_prod:
        roofline_production["counters"] = cnt_prod

    # ---- metric 2 + Stage-2 roofline: full-channel flux from device-resident weights of this run (channel N) ----
    roofline_stage2 = None
    flux_time = {"full_step_s": ms_prod / steps * 1e-3,
                 "what": "production step: Stage 1 + exchange + Stage 2 of all four channels, nothing cached"}
    if not args.no_stage2_roofline:
        el_i = 2  # N: the channel with the most visible voxels
        n = lat.n_points
        w_all = torch.empty(n, dtype=torch.float64, device=device)
        plan = scenes[el_i].plan(lat, 0, min(n, pipe.volume.chunk), device=device)
        for lo in range(0, n, pipe.volume.chunk):
            hi = min(n, lo + pipe.volume.chunk)
            plan.rebind(lo, hi).launch()
            w_all[lo:hi] = plan.weight[: hi - lo]
        del plan
        reff_mm = pipe.volume.reff_mm()
        h2 = torch.zeros((1, _lib.COM_MM_BINS), dtype=torch.float64, device=device)
        fl2 = torch.zeros((1, 1, 3), dtype=torch.float64, device=device)
        from co_monitor_b200.volume import flux_from_histograms

        cur = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        t_pass, t_all = [], []
        for _ in range(7):
            flush.fill_(3)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            h2.zero_()
            e[0].record()
            _lib.check(lib.com_weight_histogram_mm(C.c_void_p(reff_mm.data_ptr()), C.c_void_p(w_all.data_ptr()), n, None,
                                                   C.c_void_p(h2.data_ptr()), cur), "com_weight_histogram_mm")
            e[1].record()
            flux_from_histograms([tables[el_i]], pipe.profiles, h2, 0.02, fl2, None, pipe.meta)
            fl_host = fl2.cpu()
            e[2].record()
            torch.cuda.synchronize()
            t_pass.append(e[0].elapsed_time(e[1]))
            t_all.append(e[0].elapsed_time(e[2]))
        ms2 = float(np.mean(t_pass[2:]))
        peak, peak_src = hbm_peak()
        gbs = 10.0 * n / (ms2 * 1e-3) / 1e9
        rel = float(abs(fl_host[0, 0, 2].item() / flux_prod[el_i, 2] - 1)) if world == 1 else None
        roofline_stage2 = {
            "bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src,
            "kernel": "com::weight_histogram_mm_kernel", "kernel_ms_per_launch": ms2, "launch_ms": [round(t, 4) for t in t_pass],
            "voxels_per_launch": n, "algorithmic_bytes_per_voxel": 10,
            "traffic": None if not khist else (khist["dram__bytes_read.sum"] + khist["dram__bytes_write.sum"]) / khist["launches"],
            "traffic_note": "DRAM bytes per launch of this kernel in the ncu counter pass of this run (launches of one "
                            f"Stage-1 chunk, {pipe.volume.chunk} voxels = {10 * pipe.volume.chunk} algorithmic bytes)",
            "data": f"channel N of this run: real Stage-1 weights and resampled Reff bins of this rank's {n} voxels",
            "flux_rel_diff_vs_pipeline": rel,
        }
        flux_time.update({"from_cached_weights_s": float(np.mean(t_all[2:])) * 1e-3,
                          "from_cached_weights_what": f"one channel (N), {n} voxels of this rank: 10 B/voxel histogram pass + fused "
                                                      "Stage 2 + flux to the host (metric 2 of SURVEY.md 8(d))"})
        del w_all

    sub = {}
    if not args.no_subrecords:
        sub = subrecords(args, lib, device, tables, profile, flush)
    base = None
    if world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline(args.spacing)

    value = tests_per_step * steps / (ms_nc * 1e-3)
    prod_value = tests_per_step * steps / (ms_prod * 1e-3)

    def rel_diff(a, b):
        return float(np.max(np.abs(np.asarray(a) / np.asarray(b) - 1)))

    out = {
        "metric": METRIC, "value": value, "unit": "tests/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_nc / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 filter + f64 decision/weight/emissivity", "data": "synthetic",
        "config": {
            "workload": workload_name(args.spacing),
            "lattice": [full.nx, full.ny, full.nz], "voxels": full.n_points_global, "voxels_per_gpu": lat.n_points,
            "tests_per_step": tests_per_step, "shards": f"block-cyclic, {args.cols_per_block} z-columns per block",
            "launches_per_channel_per_gpu": pipe.volume.n_launches, "voxels_per_launch": pipe.volume.chunk,
            "reff": "shipped 15 mm VMEC grid resampled onto the lattice on the device (2 B/voxel, resident)",
            "l2": "256 MB buffer written between timed steps (untimed) and inputs larger than L2 (Reff bins: "
                  f"{2 * lat.n_points / 1e6:.0f} MB per GPU)",
            "timing": "sum of per-step CUDA-event intervals, barrier + synchronize on both sides, max over ranks",
            "value_is": "every ray's detector test evaluated individually (COM_FILTER_NO_RUN_CULL); `production` is the "
                        "shipped step, which skips runs of voxels the linearity of the test proves dark - every ray is decided "
                        "either way and the results are identical (flux_rel_diff)",
            "exchange": ("none (one rank)" if world == 1 else
                         f"peer memory: every rank's packed buffer ({pipe.packed.numel()} fp64) is read over NVLink by the Stage-2 "
                         "launch itself after one device-side barrier (no collective call)" if pipe.exchange_mode == "peer" else
                         f"one NCCL all-reduce(SUM) of {pipe.packed.numel()} fp64 per step ({pipe.exchange_note})"),
            "exchange_ms": exchange_ms,
            "passing_rays": {el: stats[i]["passing_rays"] for i, el in enumerate(ELEMENTS)},
            "visible_voxels": {el: stats[i]["visible_voxels"] for i, el in enumerate(ELEMENTS)},
            "fp64_candidates": {el: stats[i]["candidates"] for i, el in enumerate(ELEMENTS)},
            "certain_candidates_weight_only": {el: stats[i]["certain_candidates"] for i, el in enumerate(ELEMENTS)},
            "near_edge_rays_1e-9mm": {el: stats[i]["near_edge_rays"] for i, el in enumerate(ELEMENTS)},
            "near_edge_rays_1e-4mm": {el: stats[i]["near_band_rays"] for i, el in enumerate(ELEMENTS)},
            "near_rounding_rays_1e-9": {el: stats[i]["near_rounding_rays"] for i, el in enumerate(ELEMENTS)},
            "flux_excit_recom_total": {el: [float(x) for x in flux_prod[i]] for i, el in enumerate(ELEMENTS)},
            "total_emissivity": {el: f"{flux_prod[i, 2]:.2e}" for i, el in enumerate(ELEMENTS)},
            "reff_boundary_m": {el: float(boundary[i]) for i, el in enumerate(ELEMENTS)},
            "flux_rel_diff_value_vs_production": rel_diff(flux_nc, flux_prod),
        },
        "production": {"value": prod_value, "unit": "tests/s", "ms_per_step": ms_prod / steps, "steps": steps,
                       "flux_time_s": ms_prod / steps * 1e-3},
        "flux_time_s": flux_time,
        "e2e": {
            "value": tests_per_step * steps / e2e_nc_s, "unit": "tests/s", "ms_per_step": 1e3 * e2e_nc_s / steps,
            "h2d_bytes_per_step": int(io_bytes["h2d"]), "d2h_bytes_per_step": int(io_bytes["d2h"]),
            "ms_per_step_over_device_ms_per_step": (1e3 * e2e_nc_s / steps) / (ms_nc / steps),
            "production": {"value": tests_per_step * steps / e2e_prod_s, "unit": "tests/s", "ms_per_step": 1e3 * e2e_prod_s / steps,
                           "ms_per_step_over_device_ms_per_step": (1e3 * e2e_prod_s / steps) / (ms_prod / steps)},
            "flux_e2e_total": {el: float(flux_e2e_prod[i, 2]) for i, el in enumerate(ELEMENTS)},
            "flux_rel_diff_vs_device": max(rel_diff(flux_e2e_prod, flux_prod), rel_diff(flux_e2e_nc, flux_prod)),
            "timing": "wall clock around the K steps, barrier + synchronize on both sides, max over ranks",
            "call": "per step and rank: com_volume_histograms_host (Reff source grid up and resampled, Stage 1 of the shard "
                    "in launches + histogram passes, histograms + counters down)"
                    + (", all-reduce of the packed host buffer through pinned memory (H2D, NCCL, D2H)" if world > 1 else "")
                    + ", com_histograms_flux_host (profile + global histograms up, fused Stage 2, flux down); scene and table "
                      "handles are created once (com_scene_create / com_tables_create) and reused",
        },
        # headline step: per channel and launch the every-ray filter, the weight-only kernel, the exact kernel and the
        # histogram pass, then one Stage-2 launch (the shipped step has interval + expand kernels instead of the filter)
        "gpu_launches": int(steps * (n_ch * pipe.volume.n_launches * 4 + 1)),
        "wall_s_timed_region": wall_nc,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_production": roofline_production,
        "roofline_stage2": roofline_stage2,
        "cpu_baseline": base,
        **sub,
    }
    print(json.dumps(out), flush=True)
    pipe.close()
    teardown()


def subrecords(args, lib, device, tables, profile, flush):
    """Rank-0-only legs that keep the other BASELINE configs visible in the driver's line (none uses a collective)."""
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.engine import lattice_struct, make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import source_volume
    from co_monitor_b200.volume import VolumePipeline, flux_from_histograms

    out = {}
    n_ch = len(ELEMENTS)
    src = source_volume(15)
    n_rep = 5

    def time_steps(run, n):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        ms = []
        for i in range(n):
            flush.fill_(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.mean(ms))

    # ---- configs[1]: the golden-parity workload (10 mm lattice, 270 000 voxels x 4 channels) ----
    lat10 = quiet(CuboidMesh, 10).lattice
    sc10 = [quiet(make_channel, el, SLITS, H, L, lattice=lat10) for el in ELEMENTS]
    p10 = VolumePipeline(lat10, sc10, tables, profile, 2.0, chunk=lat10.n_points, device=device, exchange="none")
    p10.volume.set_reff_resampled(src)
    ms10 = time_steps(p10.step, 20)
    st10 = p10.global_stats()
    fl10 = p10.flux.cpu().numpy()[:, 0]
    tests10 = lat10.n_points * H * L * n_ch
    # the per-channel host-buffer calls of the drop-ins (scene / table handles kept), four channels from four threads
    from concurrent.futures import ThreadPoolExecutor

    from co_monitor_b200.geometry.reff import interpolate_reff

    reff_np = interpolate_reff(lat10, 15)
    ls = lattice_struct(lat10)
    n10 = lat10.n_points
    idx_np = [np.empty(n10, dtype=np.int64) for _ in ELEMENTS]
    w_np = [np.empty(n10) for _ in ELEMENTS]
    fl_e2e = np.zeros((n_ch, 3))

    def channel_call(i):
        st, cnt = _lib.ComStats(), C.c_uint64()
        _lib.check(lib.com_visible_weights_host_scene(sc10[i].handle, C.byref(ls), None, 0, n10, C.c_void_p(idx_np[i].ctypes.data),
                                                      C.c_void_p(w_np[i].ctypes.data), n10, C.byref(cnt), C.byref(st)),
                   "com_visible_weights_host_scene")
        m = cnt.value
        _lib.check(lib.com_emissivity_integrate_indexed_host_tables(
            tables[i].handle, _lib.as_double_ptr(profile), len(profile), 0.02, _lib.as_double_ptr(reff_np), len(reff_np),
            C.c_void_p(idx_np[i].ctypes.data), _lib.as_double_ptr(w_np[i]), m, _lib.as_double_ptr(fl_e2e[i]), None, None, None),
            "com_emissivity_integrate_indexed_host_tables")
        return m

    pool = ThreadPoolExecutor(max_workers=n_ch)

    def host_step():
        return [f.result() for f in [pool.submit(channel_call, i) for i in range(n_ch)]]

    with torch.cuda.device(device):
        for _ in range(6):
            m10 = host_step()
        t0 = time.perf_counter()
        for _ in range(20):
            host_step()
        e2e10 = (time.perf_counter() - t0) / 20
    pool.shutdown()
    out["small_grid"] = {
        "workload": "configs[1]: 10 mm lattice (270 000 voxels) x B/C/N/O, the golden-file workload",
        "tests_per_step": tests10, "ms_per_step": ms10, "value": tests10 / (ms10 * 1e-3), "unit": "tests/s",
        "passing_rays": {el: st10[i]["passing_rays"] for i, el in enumerate(ELEMENTS)},
        "visible_voxels": {el: st10[i]["visible_voxels"] for i, el in enumerate(ELEMENTS)},
        "flux_total": {el: float(fl10[i, 2]) for i, el in enumerate(ELEMENTS)},
        "e2e": {"value": tests10 / e2e10, "ms_per_step": 1e3 * e2e10,
                "h2d_bytes_per_step": int(n_ch * profile.nbytes + 16 * sum(m10)), "d2h_bytes_per_step": int(16 * sum(m10) + n_ch * 600),
                "flux_total": {el: float(fl_e2e[i, 2]) for i, el in enumerate(ELEMENTS)},
                "call": "per channel (four host threads): com_visible_weights_host_scene (visible voxels' indices + weights "
                        "down) then com_emissivity_integrate_indexed_host_tables (profile, Reff join, weights up; flux down)"},
    }
    p10.close()

    # ---- configs[3]: 1000 profile time slices from the cached weights of the 10 mm grid ----
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile

    n_slices = 1000
    profs = np.stack([TwoGaussSumProfile(NE_SWEEP, list(TE_SWEEP_LO + (TE_SWEEP_HI - TE_SWEEP_LO) * (s / (n_slices - 1)))).profiles_df[
        ["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy() for s in range(n_slices)])
    p10 = VolumePipeline(lat10, sc10, tables, profs, 2.0, chunk=lat10.n_points, device=device, exchange="none")
    p10.volume.set_reff_resampled(src)
    p10.stage1()
    ms_sw = time_steps(p10.stage2, 10)
    fl_sw = p10.flux.cpu().numpy()
    visible = sum(s["visible_voxels"] for s in p10.global_stats())
    rec = {"workload": "configs[3]: 1000 T_e slices x B/C/N/O from the cached weight histograms of the 10 mm grid",
           "sweep_1000_ms": ms_sw, "voxel_slices_per_s": visible * n_slices / (ms_sw * 1e-3), "launches": 1,
           "flux_total_slice0": {el: float(fl_sw[i, 0, 2]) for i, el in enumerate(ELEMENTS)},
           "flux_total_slice999": {el: float(fl_sw[i, 999, 2]) for i, el in enumerate(ELEMENTS)}}
    gpath = os.path.join(ROOT, "tests", "golden", "stage2_reference.npz")
    if os.path.exists(gpath):
        gold = np.load(gpath)
        rec["max_rel_err_vs_reference_slices_0_499_999"] = float(max(
            np.max(np.abs(fl_sw[ELEMENTS.index(el), s] / gold[f"sweep{s}_{el}_flux"] - 1))
            for el in "CO" for s in (0, 499, 999) if f"sweep{s}_{el}_flux" in gold))
    out["sweep_1000"] = rec
    p10.close()
    for s in sc10:
        s.close()

    # ---- explicit voxels (what the reference materialises): (P,3) fp64 + the packed float4 copy the filter streams ----
    out["explicit_voxels"] = explicit_voxels_record(lib, device, flush)

    # ---- configs[2]: the 1 mm lattice (2.7e8 voxels), whole step on this GPU ----
    if args.spacing != 1.0:
        lat1 = quiet(CuboidMesh, 1).lattice
        sc1 = [quiet(make_channel, el, SLITS, H, L, lattice=lat1) for el in ELEMENTS]
        p1 = VolumePipeline(lat1, sc1, tables, profile, 2.0, chunk=args.chunk, device=device, exchange="none")
        p1.volume.set_reff_resampled(src)
        ms1 = time_steps(p1.step, n_rep)
        st1 = p1.global_stats()
        tests1 = lat1.n_points * H * L * n_ch
        out["grid_1mm"] = {"workload": "configs[2] lattice: 1 mm (1350x800x250 = 2.7e8 voxels) x B/C/N/O, whole step on one GPU",
                           "tests_per_step": tests1, "ms_per_step": ms1, "value": tests1 / (ms1 * 1e-3), "unit": "tests/s",
                           "passing_rays": {el: st1[i]["passing_rays"] for i, el in enumerate(ELEMENTS)},
                           "flux_total": {el: float(x) for el, x in zip(ELEMENTS, p1.flux.cpu().numpy()[:, 0, 2])}}
        p1.close()
        for s in sc1:
            s.close()
    return out


def explicit_voxels_record(lib, device, flush, n_vox=1 << 25, element="C"):
    """Stage 1 on explicit voxel coordinates - the form INTEGRATION.md hands to a caller that has the reference's (P,3)
    array (mesh_calculation.py:119-159): 2^25 consecutive voxels of the 1 mm lattice through the bright region, as fp64
    rows on the device, packed once (com_voxels_pack), then com_visibility_weights_packed.  Results are compared with the
    implicit-lattice path on the same range."""
    import torch

    from co_monitor_b200 import _lib
    from co_monitor_b200.geometry.engine import make_channel
    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh

    mesh = quiet(CuboidMesh, 1)
    lat = mesh.lattice
    scene = quiet(make_channel, element, SLITS, H, L, lattice=lat)
    ix10, iy10 = BRIGHT_10MM[element]
    col = int(round(iy10 * (lat.ny - 1) / 79)) * lat.nx + int(round(ix10 * (lat.nx - 1) / 134))
    lo = max(0, min(col * lat.nz - n_vox // 2, lat.n_points - n_vox))
    pts = torch.from_numpy(mesh.points_at(np.arange(lo, lo + n_vox))).to(device)
    peak, _ = hbm_peak()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # pack: 24 B in + 16 B out per voxel
    packed = scene.pack_voxels(pts)
    t_pack = []
    for i in range(5):
        flush.fill_(i)
        a, b = ev(), ev()
        a.record()
        packed = scene.pack_voxels(pts)
        b.record()
        torch.cuda.synchronize()
        t_pack.append(a.elapsed_time(b))
    ms_pack = float(np.mean(t_pack[1:]))
    # Stage 1 on the packed voxels: whole call (filter + fp64 decision) by events, the filter kernel by the library's events
    res = scene.visibility(voxels=pts, packed=packed)
    t_all, t_filter = [], []
    for i in range(5):
        flush.fill_(i)
        lib.com_kernel_timing_enable(1)
        a, b = ev(), ev()
        a.record()
        res = scene.visibility(voxels=pts, packed=packed)
        b.record()
        torch.cuda.synchronize()
        ms3, n_l = (C.c_double * 3)(), C.c_int64()
        lib.com_kernel_timing_read3(ms3, C.byref(n_l))
        lib.com_kernel_timing_enable(0)
        t_all.append(a.elapsed_time(b))
        t_filter.append(ms3[0])
    ms_all, ms_filter = float(np.mean(t_all[1:])), float(np.mean(t_filter[1:]))
    ref = scene.visibility(lattice=lat, begin=lo, end=lo + n_vox)
    same = bool(torch.equal(ref.nrays, res.nrays)) and bool(torch.allclose(ref.weight, res.weight, rtol=1e-12, atol=0))
    groups = -(-H * L // 32)
    n_super = -(-groups // 16)  # crystal super-groups per tile: every tile is streamed once per super-group
    tests = n_vox * H * L
    rec = {
        "workload": f"{n_vox} explicit voxels (fp64 (P,3) on the device: consecutive voxels of the 1 mm lattice through the "
                    f"{element} channel's bright region) x {H * L} crystal points",
        "tests": tests, "ms": ms_all, "value": tests / (ms_all * 1e-3), "unit": "tests/s",
        "what": "whole com_visibility_weights_packed call: FP32 filter on the packed float4 voxels (512-voxel tiles moved to "
                "shared memory by cp.async.bulk, double buffered) + fp64 decision of the candidates (gathers the fp64 rows)",
        "filter_kernel_ms": ms_filter, "filter_tests_per_s": tests / (ms_filter * 1e-3),
        "filter_hbm": {"algorithmic_bytes": 16 * n_vox * n_super, "GBps": 16 * n_vox * n_super / (ms_filter * 1e-3) / 1e9,
                       "frac_of_hbm_peak": 16 * n_vox * n_super / (ms_filter * 1e-3) / 1e9 / peak,
                       "note": f"16 B per voxel per crystal super-group ({n_super} per tile); the kernel is FP32-issue bound "
                               f"({H * L} rays per 16 bytes), not HBM bound"},
        "pack_ms": ms_pack, "pack_GBps": 40 * n_vox / (ms_pack * 1e-3) / 1e9, "pack_frac_of_hbm_peak": 40 * n_vox / (ms_pack * 1e-3) / 1e9 / peak,
        "passing_rays": res.stats["passing_rays"], "candidates": res.stats["candidates"],
        "equal_to_the_lattice_path": same,
    }
    scene.close()
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spacing", type=float, default=0.65, help="lattice spacing in mm (0.65: configs[4], the headline)")
    ap.add_argument("--chunk", type=int, default=1 << 26, help="voxels per Stage-1 launch")
    ap.add_argument("--cols-per-block", type=int, default=8, help="z-columns per block of the block-cyclic shards")
    ap.add_argument("--exchange", default="auto", choices=["auto", "peer", "nccl"],
                    help="N > 1: peer = histograms read over NVLink by the Stage-2 launch itself; nccl = all-reduce; auto = peer if it can be set up")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ncu", action="store_true", help="skip the ncu counter pass (roofline instruction counts / DRAM bytes)")
    ap.add_argument("--no-stage2-roofline", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the small_grid / sweep_1000 / grid_1mm legs")
    ap.add_argument("--probe", default="", help="internal: run one channel's whole-volume Stage 1 once (ncu counter pass)")
    ap.add_argument("--probe-mode", default="production", choices=["production", "nocull"])
    args = ap.parse_args()
    if args.probe:
        run_probe(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
