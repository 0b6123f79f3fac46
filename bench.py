#!/usr/bin/env python
"""Benchmark of the co_monitor hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* = Stage 1 (visibility + geometric weight) for all four channels B/C/N/O on the default
10 mm voxel grid with 30 x 20 crystal points and 10 slits (BASELINE.json configs[1]:
4 x 270 000 x 600 = 6.48e8 voxel-detector visibility tests), followed by Stage 2 when built.
At N > 1 the grid is refined N-fold along z and its flat index range is sharded evenly over the ranks
(weak scaling: 6.48e8 tests per GPU), with one NCCL all-reduce of the per-channel weight sums.

Prints ONE JSON line (rank 0).  `value` = whole-job tests/s from CUDA-event times (inputs resident in HBM),
`e2e` = the same metric through the host-buffer C-ABI call, `roofline` = the FP32 filter kernel against the
FP32 issue peak, `cpu_baseline` = the NumPy/Qhull oracle on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
SPACING, H, L, SLITS = 10, 30, 20, 10
FLOP_PER_TEST = 247  # canonical no-early-out count, FMA = 2
METRIC = "voxel-detector visibility tests/s"
NE_MAIN = [7e13, 0, 0.37, 9.8e12, 0.5, 0.11]  # reference __main__.py:35-36
TE_MAIN = [1870, 0, 0.155, 210, 0.38, 0.07]


def nominal_fp32_tflops(sm_mhz: float, sms: int = 148) -> float:
    return sms * 128 * 2 * sm_mhz * 1e6 / 1e12


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

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
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(names, r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def refined_lattice(n_ranks: int):
    """Default CuboidMesh lattice, z-sampling refined n_ranks-fold (same cuboid, same rotation)."""
    import contextlib
    import io
    from dataclasses import replace

    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh

    with contextlib.redirect_stdout(io.StringIO()):
        lat = CuboidMesh(SPACING).lattice
    if n_ranks > 1:
        nz = lat.nz * n_ranks
        step = lat.step.copy()
        step[2] = (lat.stop[2] - lat.start[2]) / (nz - 1)
        lat = replace(lat, nz=nz, step=step)
    return lat


def cpu_baseline(sample_voxels_per_core: int = 1500):
    """Oracle (NumPy + Qhull restatement of the reference) on a bounded C-channel sample, all host cores, Stage 1 + 2."""
    import contextlib
    import io

    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import interpolate_reff

    with contextlib.redirect_stdout(io.StringIO()):
        reff_all = interpolate_reff(CuboidMesh(SPACING).lattice, 15)
    value, d = reference_step("C", sample_voxels_per_core, reff_all)
    return {"value": value, "unit": "tests/s", "cores": d["processes"], "kind": "port",
            "sample": f"C channel, voxels {d['voxels']} of the 10 mm grid x {H*L} crystal points x 22 apertures (Stage 1 oracle, "
                      f"{d['stage1_s']:.1f} s on {d['processes']} processes of {d['cores_visible']} cores) + Stage-2 oracle on the "
                      f"visible voxels of that range ({d['stage2_s']:.2f} s + {d['tables_s_prorated']:.2f} s of table building pro rata)"}


def reference_step(element, sample_voxels_per_core, reff_all):
    """One bounded sample of the workload on the CPU path: Stage 1 (oracle, all host cores) on a voxel range of one
    channel, then Stage 2 (oracle) on the visible voxels of that range.  The reference rebuilds its PEC / FA tables in
    every Emissivity call (reader.py:83-97, 245-261); that fixed cost is counted pro rata (range / whole grid)."""
    from oracle import stage1 as o1
    from oracle import stage2 as o2
    from co_monitor_b200.pipeline import LYMAN_ALPHA

    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    lo = 120000
    hi = min(270000, lo + sample_voxels_per_core * procs)
    t0 = time.perf_counter()
    w, nr, used = o1.run_range(element, lo, hi, S=SLITS, spacing=SPACING, h=H, l=L, processes=procs, chunk=500)
    t1 = time.perf_counter() - t0
    ion, wl = LYMAN_ALPHA[element]
    t0 = time.perf_counter()
    tables = o2.load_tables(element, ion, wl, ["EXCIT", "RECOM"])
    t_tables = time.perf_counter() - t0
    vis = np.flatnonzero(nr > 0)
    obs = np.column_stack((vis + lo, np.zeros((len(vis), 3)), w[vis]))
    t0 = time.perf_counter()
    flux = None
    if len(vis):
        res = o2.emissivity(element, ion, wl, 2, ["EXCIT", "RECOM"], o2.two_gauss_profile(NE_MAIN, TE_MAIN), obs, reff_all,
                            tables=tables)
        flux = res.flux["TOTAL"]
    t2 = time.perf_counter() - t0
    frac = (hi - lo) / 270000.0
    seconds = t1 + t2 + t_tables * frac
    tests = (hi - lo) * H * L
    return tests / seconds, {"element": element, "voxels": [lo, hi], "tests": tests, "stage1_s": t1, "stage2_s": t2,
                              "tables_s_prorated": t_tables * frac, "processes": used, "cores_visible": cores, "flux": flux}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores.  The reference itself needs
    dask / opt_einsum / pyvista (absent, no network), so this is the oracle port of it - the NumPy + Qhull restatement
    that reproduces the reference's shipped outputs exactly.  Each step is a bounded sample of the workload (one channel
    per step, B/C/N/O in turn)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import contextlib
    import io

    from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
    from co_monitor_b200.geometry.reff import interpolate_reff

    with contextlib.redirect_stdout(io.StringIO()):
        reff_all = interpolate_reff(CuboidMesh(SPACING).lattice, 15)
    per_core = 400 if args.steps + args.warmup > 6 else 1000
    vals, details = [], []
    for i in range(args.warmup + args.steps):
        v, d = reference_step(ELEMENTS[i % len(ELEMENTS)], per_core, reff_all)
        if i >= args.warmup:
            vals.append(v)
            details.append(d)
    value = float(np.mean(vals))
    full_step_tests = 270000 * H * L * len(ELEMENTS)  # what one step of the accelerated arm processes at N = 1
    base = {"value": value, "unit": "tests/s", "cores": details[-1]["processes"], "kind": "port",
            "sample": f"per step: one channel (B/C/N/O in turn), voxels {details[-1]['voxels']} of the 10 mm grid x {H*L} crystal "
                      f"points x 22 apertures through the NumPy/Qhull oracle on {details[-1]['processes']} processes, then the "
                      f"Stage-2 oracle on the visible voxels of that range; last step: {details[-1]}"}
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tests/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * full_step_tests / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "configs[1]: B/C/N/O, Stage 1 + Stage 2 on the default 10 mm grid (270 000 voxels x 600 crystal "
                               "points x 10 slits); each step is a bounded sample (one channel, a voxel range)",
                   "tests_per_step": full_step_tests,
                   "ms_per_step_is": "the time one full step (6.48e8 tests) would take at the sampled rate; the sampled "
                                     "steps themselves take a few seconds each"},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "tests/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))


def run_ours(args):
    import contextlib
    import io

    import torch
    import torch.distributed as dist

    from co_monitor_b200 import _lib
    from co_monitor_b200.emissivity.engine import LineTables
    from co_monitor_b200.emissivity.kinetic_profiles import TwoGaussSumProfile
    from co_monitor_b200.geometry.engine import lattice_struct, make_channel
    from co_monitor_b200.geometry.reff import interpolate_reff
    from co_monitor_b200.pipeline import LYMAN_ALPHA, ChannelPipeline

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the accelerated path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _lib.load()

    full_lattice = refined_lattice(world)
    P = full_lattice.n_points
    # block-cyclic shard of the z-columns: the visible voxels are spatially clustered, contiguous ranges would not balance
    cols_per_block = int(os.environ.get("COM_BENCH_COLS_PER_BLOCK", "8"))
    lattice = full_lattice.shard(world, rank, cols_per_block=cols_per_block)
    lo, hi = 0, lattice.n_points
    # Reff of every lattice voxel: the shipped 15 mm VMEC grid resampled onto the lattice (no VMEC service here),
    # re-ordered to this shard's local voxel order
    reff_all = interpolate_reff(full_lattice, 15)
    reff_table = torch.from_numpy(reff_all[lattice.local_to_global(np.arange(lattice.n_points))]).to(device)
    profile = TwoGaussSumProfile(NE_MAIN, TE_MAIN).profiles_df[["Reff [m]", "T_e [eV]", "n_e [m-3]"]].to_numpy()
    scenes, pipes = {}, {}
    with contextlib.redirect_stdout(io.StringIO()):
        for el in ELEMENTS:
            scenes[el] = make_channel(el, SLITS, H, L, lattice=lattice)
            ion, wl = LYMAN_ALPHA[el]
            tables = LineTables(el, ion, wl, ["EXCIT", "RECOM"])
            pipes[el] = ChannelPipeline(scenes[el], tables, lattice, reff_table, lo, hi, 2.0, device=device)
            pipes[el].set_profile(profile)
    tests_per_step_rank = (hi - lo) * H * L * len(ELEMENTS)
    flux = torch.zeros((len(ELEMENTS), 3), dtype=torch.float64, device=device)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2

    main_stream = torch.cuda.current_stream(device)
    streams = {el: torch.cuda.Stream(device=device) for el in ELEMENTS}
    fork_ev = torch.cuda.Event()
    join_ev = {el: torch.cuda.Event() for el in ELEMENTS}

    def step_sequential():
        """One stream, channel after channel (used for the per-kernel timing pass)."""
        for el in ELEMENTS:
            pipes[el].launch_geometry()
            if world > 1:
                dist.all_reduce(pipes[el].hist_mm)
        for i, el in enumerate(ELEMENTS):
            pipes[el].launch_emissivity()
            flux[i].copy_(pipes[el].flux)

    def step():
        """Stage 1 + Stage 2 for the four channels, each channel on its own stream (they are independent).  Stage 2
        works from the millimetre-bin weight histogram of the rank's voxels; at N > 1 the per-rank histograms
        (4096 fp64 bins, 32 kB) are summed with one NCCL all-reduce per channel on the channel's stream, after which
        every rank holds the whole-grid flux - the single exchange of the path."""
        cur = torch.cuda.current_stream(device)
        fork_ev.record(cur)
        for i, el in launch_order:
            with torch.cuda.stream(streams[el]):
                streams[el].wait_event(fork_ev)
                pipes[el].launch_geometry()
                if world > 1:
                    dist.all_reduce(pipes[el].hist_mm)
                pipes[el].launch_emissivity()
                flux[i].copy_(pipes[el].flux)
                join_ev[el].record(streams[el])
        for el in ELEMENTS:
            cur.wait_event(join_ev[el])

    # channels are enqueued heaviest first (by fp64 candidates, known after the first warm-up step): the filter
    # kernels fill the GPU one after the other, so the channel enqueued last leaves its exact kernel and Stage 2
    # exposed at the end of the step - that should be the cheapest one
    launch_order = list(enumerate(ELEMENTS))
    for w in range(max(args.warmup, 3)):
        flush.fill_(1)
        step()
        if w == 0:
            torch.cuda.synchronize()
            cand = {el: pipes[el].plan.stats()["candidates"] for el in ELEMENTS}
            launch_order = sorted(enumerate(ELEMENTS), key=lambda ie: -cand[ie[1]])
    torch.cuda.synchronize()
    stats = {el: pipes[el].plan.stats() for el in ELEMENTS}

    # one CUDA graph of the whole step (every launch of the path is stream-ordered, nothing syncs with the host)
    graph = None
    if not args.no_graph and (world == 1 or not args.no_graph_multi_gpu):
        try:
            graph = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(device=device)
            side.wait_stream(main_stream)
            with torch.cuda.stream(side):
                step()
            main_stream.wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                step()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:  # e.g. an NCCL build that cannot be captured: launch kernel by kernel instead
            if rank == 0:
                print(f"bench.py: CUDA graph capture failed ({type(exc).__name__}: {exc}); running without", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()
    if world > 1:  # all ranks must take the same path
        flag = torch.tensor([1 if graph is not None else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            graph = None
    run_step = graph.replay if graph is not None else step

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps (untimed) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        starts[i].record()
        run_step()
        ends[i].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_tests = tests_per_step_rank * world * args.steps
    value = total_tests / (ms_max * 1e-3)
    flux_h = flux.cpu().numpy()

    # per-kernel timing pass: same step, one stream, no graph, events around K1a / K1b inside the library
    lib.com_kernel_timing_enable(1)
    seq_start, seq_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_seq = min(args.steps, 10)
    seq_ms = 0.0
    for i in range(n_seq):
        flush.fill_(i & 0xFF)
        seq_start.record()
        step_sequential()
        seq_end.record()
        torch.cuda.synchronize()
        seq_ms += seq_start.elapsed_time(seq_end)
    f_ms, x_ms, n_launch = C.c_double(), C.c_double(), C.c_int64()
    lib.com_kernel_timing_read(C.byref(f_ms), C.byref(x_ms), C.byref(n_launch))
    lib.com_kernel_timing_enable(0)

    # ---- the same step with every ray's detector test evaluated individually (SURVEY.md 8(d) asks for the rate without
    # culling next to the headline): scenes built with COM_FILTER_NO_RUN_CULL, identical results, N = 1 only ----
    every_ray = None
    if world == 1 and not args.no_every_ray:
        saved = dict(pipes)
        with contextlib.redirect_stdout(io.StringIO()):
            for el in ELEMENTS:
                sc_nc = make_channel(el, SLITS, H, L, lattice=lattice, run_cull=False)
                scenes[el + "_nc"] = sc_nc
                pipes[el] = ChannelPipeline(sc_nc, saved[el].tables, lattice, reff_table, lo, hi, 2.0, device=device)
                pipes[el].set_profile(profile)
        for _ in range(3):
            flush.fill_(1)
            step()
        torch.cuda.synchronize()
        run_nc = step
        g_nc = None
        if graph is not None:
            try:
                g_nc = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_nc):
                    step()
                g_nc.replay()
                torch.cuda.synchronize()
                run_nc = g_nc.replay
            except Exception:  # pragma: no cover - measure without a graph then
                g_nc = None
                torch.cuda.synchronize()
        e0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        e1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            e0[i].record()
            run_nc()
            e1[i].record()
        torch.cuda.synchronize()
        ms_nc = sum(a.elapsed_time(b) for a, b in zip(e0, e1))
        same = bool(np.allclose(flux.cpu().numpy(), flux_h, rtol=1e-12))
        every_ray = {"value": tests_per_step_rank * args.steps / (ms_nc * 1e-3), "unit": "tests/s",
                     "ms_per_step": ms_nc / args.steps, "flux_equal": same,
                     "passing_rays": {el: pipes[el].plan.stats()["passing_rays"] for el in ELEMENTS}}
        del g_nc
        pipes.update(saved)

    # ---- e2e: the reference-facing host-buffer calls, per channel: com_visibility_weights_host (scene upload,
    # kernels, D2H of per-voxel weights + ray counts) then com_emissivity_integrate_host on the visible voxels ----
    n = hi - lo
    cap = n
    idx_np = {el: np.empty(cap, dtype=np.int64) for el in ELEMENTS}
    w_np = {el: np.empty(cap, dtype=np.float64) for el in ELEMENTS}
    reff_np = reff_table.cpu().numpy()
    prof_np = np.ascontiguousarray(profile)
    lat_s = lattice_struct(lattice)
    flux_e2e = np.zeros((len(ELEMENTS), 3))
    bytes_io = {"h2d": 0, "d2h": 0}

    def e2e_channel(i, el, count_bytes=False):
        """The reference-facing calls of one channel, host buffers in and out (ctypes releases the GIL inside)."""
        st, cnt = _lib.ComStats(), C.c_uint64()
        rc = lib.com_visible_weights_host(C.byref(scenes[el].desc), C.byref(lat_s), None, lo, hi,
                                          C.c_void_p(idx_np[el].ctypes.data), C.c_void_p(w_np[el].ctypes.data), cap,
                                          C.byref(cnt), C.byref(st))
        _lib.check(rc, "com_visible_weights_host")
        m = cnt.value
        # the join with the Reff table (reader.py:161-197) happens inside the call, on the visible voxels' indices
        rc = lib.com_emissivity_integrate_indexed_host(
            C.byref(pipes[el].tables.desc), _lib.as_double_ptr(prof_np), len(prof_np), 0.02, _lib.as_double_ptr(reff_np),
            len(reff_np), C.c_void_p(idx_np[el].ctypes.data), _lib.as_double_ptr(w_np[el]), m,
            _lib.as_double_ptr(flux_e2e[i]), None, None, None)
        _lib.check(rc, "com_emissivity_integrate_indexed_host")
        if count_bytes:
            k = scenes[el]._keep
            up = sum(k[x].nbytes for x in ("xyz", "nrm", "sc", "sp", "pv", "dv"))
            up += C.sizeof(_lib.ComAperture) * (2 * SLITS + 2) + prof_np.nbytes + 16 * m
            up += sum(a.nbytes for a in pipes[el].tables._keep)
            return up, 16 * m + 512 + 24
        return 0, 0

    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(max_workers=len(ELEMENTS))

    def e2e_step(count_bytes=False):
        """Four channels from four host threads: each thread has its own arena and CUDA stream inside the library."""
        futs = [pool.submit(e2e_channel, i, el, count_bytes) for i, el in enumerate(ELEMENTS)]
        for f in futs:
            up, down = f.result()
            bytes_io["h2d"] += up
            bytes_io["d2h"] += down

    e2e_step(count_bytes=True)
    for _ in range(max(args.warmup, 3) + 2):  # every pool thread grows its arena, whichever channel it gets
        e2e_step()
    if world > 1:
        dist.barrier()
    e2e_times = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        e2e_step()
        e2e_times.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None  # sampled across the timed steps, the kernel-timing pass and the e2e steps
    t = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = total_tests / float(t.item())

    def teardown():
        """Every rank leaves the same way: captured graph released first, then a barrier, then the process group."""
        nonlocal graph, run_step
        graph = run_step = None
        import gc

        gc.collect()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return

    # ---- roofline of the dominant kernel (K1a FP32 filter) ----
    launches = max(1, n_launch.value)
    filter_ms = f_ms.value / launches
    exact_ms = x_ms.value / launches
    tests_per_launch = (hi - lo) * H * L
    achieved = FLOP_PER_TEST * tests_per_launch / (filter_ms * 1e-3) / 1e12 if filter_ms > 0 else None
    measured_peak = lib.com_measure_fp32_tflops()
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    peak = nominal_fp32_tflops(sm_max)
    # Hardware-side figure next to the canonical one: lane-instructions actually executed per test, taken from the
    # committed ncu capture of this kernel on this workload (profiles/r1c_ncu_full_summary_k1.txt: warp instructions x
    # active threads per instruction / 1.62e8 tests, mean of the four channels), against the issue peak
    # 148 SM x 4 schedulers x 32 lanes x f_max.
    LANE_INSTR_PER_TEST = 11.3
    issue_peak = 148 * 4 * 32 * sm_max * 1e6
    issue_frac = (tests_per_launch / (filter_ms * 1e-3)) * LANE_INSTR_PER_TEST / issue_peak if filter_ms > 0 else None
    roofline = {
        "bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
        "frac": achieved / peak if achieved else None,
        "traffic": 1.19e5,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture in profiles/ "
                        "(119 kB: geometry tables only; the lattice kernel reads no voxel data and writes 8 B per candidate via L2)",
        "peak_source": f"nominal 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz (MEASURED_PEAKS.json has no FP32 figure); "
                       f"register-only FMA probe measured {measured_peak:.1f} TFLOP/s on this GPU in this run",
        "measured_fp32_tflops": measured_peak,
        "kernel": "com::filter_kernel_lattice (K1a)", "kernel_ms_per_launch": filter_ms, "exact_kernel_ms_per_launch": exact_ms,
        "kernel_share_of_step": (f_ms.value + x_ms.value) / seq_ms if seq_ms > 0 else None,
        "filter_kernel_share_of_step": f_ms.value / seq_ms if seq_ms > 0 else None,
        "sequential_ms_per_step": seq_ms / n_seq,
        "timing_pass": "kernel times and shares come from a separate pass of the same step on one stream without the "
                       "CUDA graph (the headline step overlaps the four channels on four streams)",
        "algorithmic_flop_per_test": FLOP_PER_TEST,
        "executed_lane_instr_per_test": LANE_INSTR_PER_TEST,
        "issue_frac": issue_frac,
        "note": "`frac` uses the canonical 247 FLOP/test of SURVEY.md 8(d) as the contract asks; it exceeds 1 because the kernel does "
                "not execute that arithmetic: mirroring the detector in the crystal makes the aperture coordinates linear "
                "forms of the voxel position, advanced along z with 3 FADD per test (DESIGN.md section 4). The hardware "
                "figure is `issue_frac`: executed lane-instructions per second over the issue peak (ncu: 70-75 % issue-slot "
                "utilisation while the kernel runs).",
    }

    # ---- second roofline: the Stage-2 voxel pass (HBM bound; SURVEY.md 8(d): 10 B per voxel per channel) ----
    # At the benchmark's 270 000 voxels per channel that pass is launch-latency bound, so it is measured here on a
    # synthetic array of 5e8 voxels (5 GB, far beyond L2; a 0.65 mm grid has 9.8e8 per channel): uint16 Reff
    # millimetres, fp64 weights of which ~89 % are zero like a real Stage-1 output.
    roofline_stage2 = None
    if not args.no_stage2_roofline:
        n2 = 500_000_000
        gen = torch.Generator(device=device).manual_seed(0)
        r2 = torch.randint(0, 600, (n2,), dtype=torch.int32, device=device, generator=gen).to(torch.int16)
        w2 = torch.rand(n2, dtype=torch.float64, device=device, generator=gen)
        w2[torch.rand(n2, device=device, generator=gen) < 0.89] = 0.0
        h2 = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device=device)
        cur = C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        times = []
        for _ in range(7):
            h2.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.com_weight_histogram_mm(C.c_void_p(r2.data_ptr()), C.c_void_p(w2.data_ptr()), n2, None,
                                                   C.c_void_p(h2.data_ptr()), cur), "com_weight_histogram_mm")
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms2 = float(np.mean(times[2:]))  # average launch duration after two warm-up launches
        want = torch.bincount(r2.to(torch.int64), weights=w2, minlength=_lib.COM_MM_BINS)[: _lib.COM_MM_BINS]
        err2 = float(((h2 - want).abs().max() / want.abs().max()).item())
        del r2, w2
        hbm_peak, src = 6650.0, "fallback of B200_PROFILING.md"
        try:
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as fh:
                hbm_peak, src = float(json.load(fh)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except (OSError, KeyError, ValueError):
            pass
        gbs = 10.0 * n2 / (ms2 * 1e-3) / 1e9
        roofline_stage2 = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                           "traffic": 5.0044e9,
                           "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this launch from the ncu --set full "
                                           "capture profiles/r1c_ncu_full_summary_hist_mm.txt (5.000 GB read = the algorithmic "
                                           "10 B x 5e8 voxels, 4.4 MB written)",
                           "peak_source": src, "kernel": "com::weight_histogram_mm_kernel",
                           "kernel_ms_per_launch": ms2, "launch_ms": [round(t, 4) for t in times],
                           "voxels_per_launch": n2, "algorithmic_bytes_per_voxel": 10,
                           "max_rel_err_vs_bincount": err2,
                           "note": "the Stage-2 voxel pass on a synthetic 5e8-voxel channel (inputs larger than L2); in the "
                                   "timed step above the same kernel runs on 270 000 voxels and is latency bound"}

    base = None
    if world == 1 and not args.no_cpu_baseline:
        base = cpu_baseline()

    out = {
        "metric": METRIC, "value": value, "unit": "tests/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 filter + f64 decision/weight/emissivity", "data": "synthetic",
        "config": {
            "workload": f"configs[1]: B/C/N/O, Stage 1 (visibility + weight) then Stage 2 (EXCIT + RECOM emissivity -> flux), "
                        f"10 mm CuboidMesh lattice {lattice.nx}x{lattice.ny}x{lattice.nz} ({P} voxels, {hi - lo} per GPU, "
                        f"block-cyclic shards of {cols_per_block} z-columns) "
                        f"x {H}x{L} crystal points x {SLITS} slits, Reff = shipped 15 mm VMEC grid resampled",
            "tests_per_step": tests_per_step_rank * world,
            "l2": "256 MB buffer written between timed steps (L2 flushed, untimed)",
            "timing": "sum of per-step CUDA-event intervals, max over ranks",
            "launch": ("one CUDA graph per step, four channels on four streams" if graph is not None else
                       "four channels on four streams, no graph"),
            "decided_how": "every (voxel, crystal point) ray is decided: by the FP32 filter (its detector test evaluated per "
                           "ray, or for a whole run of z-consecutive voxels at once when the linearity of the test proves that "
                           "none of them can pass) and, for the candidates, in fp64; nothing is pre-selected or skipped. "
                           "`every_ray_evaluated` is the same step with the run-level proof switched off",
            "every_ray_evaluated": every_ray,
            "passing_rays": {el: stats[el]["passing_rays"] for el in ELEMENTS},
            "fp64_candidates": {el: stats[el]["candidates"] for el in ELEMENTS},
            "near_edge_rays": {el: stats[el]["near_edge_rays"] for el in ELEMENTS},
            "flux_excit_recom_total": {el: [float(x) for x in flux_h[i]] for i, el in enumerate(ELEMENTS)},
            "flux_e2e_total": {el: float(flux_e2e[i, 2]) for i, el in enumerate(ELEMENTS)},
        },
        "e2e": {"value": e2e_value, "unit": "tests/s", "h2d_bytes_per_step": int(bytes_io["h2d"]),
                "d2h_bytes_per_step": int(bytes_io["d2h"]),
                "ms_per_step": 1e3 * e2e_s / args.steps, "ms_per_step_median": 1e3 * float(np.median(e2e_times)),
                "ms_per_step_max": 1e3 * float(np.max(e2e_times)),
                "call": "per channel, four channels from four host threads: com_visible_weights_host (geometry up, K1a/K1b + "
                        "compaction, visible voxels' indices and weights down) then com_emissivity_integrate_indexed_host (tables, "
                        "profile, Reff of the visible voxels and weights up, flux down)"},
        "gpu_launches": int(args.steps * len(ELEMENTS) * pipes[ELEMENTS[0]].kernels_per_launch),
        "wall_s_timed_region": wall,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_stage2": roofline_stage2,
        "cpu_baseline": base,
    }
    print(json.dumps(out), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-every-ray", action="store_true", help="skip the leg that re-times the step without run-level culling")
    ap.add_argument("--no-stage2-roofline", action="store_true", help="skip the HBM roofline leg of the Stage-2 voxel pass")
    ap.add_argument("--no-graph", action="store_true", help="launch the step kernel by kernel instead of replaying a CUDA graph")
    ap.add_argument("--no-graph-multi-gpu", action="store_true",
                    help="at N > 1 do not capture the step (NCCL all-reduces included) in a CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
