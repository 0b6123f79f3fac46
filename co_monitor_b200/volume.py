"""Whole-volume runs: per-channel line flux of a dense observation volume without per-voxel tables.

Host side of ``csrc/volume.cu`` (``ComVolume``) and of the fused Stage-2 launch ``com_histograms_flux``.  This is the
path of BASELINE.json configs 3-5 and of SURVEY.md section 8(f) rank 2 (fused Stage 1 -> Stage 2):

    Reff bins of the (shard of the) lattice on the device           (reff_VMEC.py:85-176 replaced, 2 B/voxel)
    per channel: Stage 1 launch by launch, each launch followed by the 10 B/voxel pass that folds its weights into
                 the channel's Reff-millimetre histogram             (simulation.py:50-72 -> reader.py:136-243)
    ONE all-reduce(SUM) of the packed [channels, 4096] histograms + counters over the ranks
    Stage 2 from the global histograms in one launch                 (reader.py:199-401)

``ObservationVolume`` / ``flux_from_histograms`` are the stream-ordered device forms (torch supplies memory, streams
and the NCCL collective only); the ``*_host`` forms take and return host arrays through the C ABI's host-buffer entry
points - what a binding on the reference's side would call.  No CPU fallback anywhere.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .geometry.engine import lattice_struct


def _handles(objs):
    arr = (C.c_void_p * len(objs))()
    for i, o in enumerate(objs):
        h = o.handle
        arr[i] = h.value if isinstance(h, C.c_void_p) else h
    return arr


def _stats_dicts(raw: np.ndarray) -> list[dict]:
    names = [f[0] for f in _lib.ComStats._fields_]
    return [dict(zip(names, (int(x) for x in row))) for row in raw.reshape(-1, _lib.N_STATS)]


class ObservationVolume:
    """One (shard of a) ``Lattice`` prepared for whole-volume runs."""

    def __init__(self, lattice, chunk: int = 1 << 26, candidate_fraction: float = 0.0, device=None):
        import torch

        if not torch.cuda.is_available():
            raise _lib.ComonitorError(_lib.COM_E_CUDA, "ObservationVolume", "no CUDA device: the accelerated path has no CPU fallback")
        self._lib = _lib.load()
        self.lattice = lattice
        self.device = torch.device(device if device is not None else "cuda")
        self._lat = lattice_struct(lattice)
        self.handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.com_volume_create(C.byref(self._lat), int(chunk), float(candidate_fraction),
                                                   C.byref(self.handle)), "com_volume_create")
        self.n_voxels = int(self._lib.com_volume_voxels(self.handle))
        self.chunk = int(self._lib.com_volume_chunk(self.handle))
        self.n_launches = -(-self.n_voxels // self.chunk)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.com_volume_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        import torch

        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- Reff bins ------------------------------------------------------------------------------------------
    def set_reff_resampled(self, source):
        """*source*: ``(vol (ny, nx, nz) float64 host array or CUDA tensor, its Lattice)`` - e.g. ``source_volume(15)``."""
        import torch

        vol, ls = source
        with torch.cuda.device(self.device):
            if torch.is_tensor(vol):
                vol = vol.to(device=self.device, dtype=torch.float64).contiguous()
                self._src_keep = vol
                rc = self._lib.com_volume_set_reff_resampled(self.handle, C.c_void_p(vol.data_ptr()), 0, ls.nx, ls.ny, ls.nz,
                                                             self._stream())
            else:
                vol = np.ascontiguousarray(vol, dtype=np.float64)
                self._src_keep = vol
                rc = self._lib.com_volume_set_reff_resampled(self.handle, C.c_void_p(vol.ctypes.data), 1, ls.nx, ls.ny, ls.nz,
                                                             self._stream())
        _lib.check(rc, "com_volume_set_reff_resampled")

    def set_reff_quadratic(self, model):
        import torch

        coef = (C.c_double * 10)(*[float(c) for c in model.coef])
        with torch.cuda.device(self.device):
            rc = self._lib.com_volume_set_reff_quadratic(self.handle, coef, float(model.reff_max), self._stream())
        _lib.check(rc, "com_volume_set_reff_quadratic")

    def reff_mm(self):
        """The volume's Reff millimetre bins as a CUDA uint16 tensor view (no copy)."""
        import torch

        ptr = self._lib.com_volume_reff_mm(self.handle)
        n = self.n_voxels

        class _Holder:
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<u2", "data": (int(ptr), False), "version": 3}

        return torch.as_tensor(_Holder(), device=self.device)

    # ---- Stage 1 -> histograms --------------------------------------------------------------------------------
    def histograms(self, scenes, hist=None, stats=None):
        """Stream-ordered: ``hist`` (n_channels, 4096) float64 and ``stats`` (n_channels, N_STATS) int64 CUDA tensors
        (allocated when not given; both are overwritten)."""
        import torch

        n_ch = len(scenes)
        with torch.cuda.device(self.device):
            if hist is None:
                hist = torch.empty((n_ch, _lib.COM_MM_BINS), dtype=torch.float64, device=self.device)
            if stats is None:
                stats = torch.empty((n_ch, _lib.N_STATS), dtype=torch.int64, device=self.device)
            rc = self._lib.com_volume_histograms(self.handle, n_ch, _handles(scenes), C.c_void_p(hist.data_ptr()),
                                                 C.c_void_p(stats.data_ptr()), self._stream())
        _lib.check(rc, "com_volume_histograms")
        return hist, stats

    def histograms_host(self, scenes, reff_source=None, hist_out: np.ndarray | None = None):
        """Host-buffer form: returns ``(hist (n_channels, 4096) ndarray, [stats dict per channel])``.  *reff_source*
        ``(vol host array, Lattice)`` (re)builds the Reff bins from that host grid inside the call."""
        import torch

        n_ch = len(scenes)
        hist = hist_out if hist_out is not None else np.empty((n_ch, _lib.COM_MM_BINS))
        raw = np.zeros((n_ch, _lib.N_STATS), dtype=np.uint64)
        src, dims = None, (0, 0, 0)
        if reff_source is not None:
            vol, ls = reff_source
            vol = np.ascontiguousarray(vol, dtype=np.float64)
            src, dims = C.c_void_p(vol.ctypes.data), (ls.nx, ls.ny, ls.nz)
        with torch.cuda.device(self.device):
            rc = self._lib.com_volume_histograms_host(self.handle, n_ch, _handles(scenes), src, *dims,
                                                      C.c_void_p(hist.ctypes.data), C.c_void_p(raw.ctypes.data))
        _lib.check(rc, "com_volume_histograms_host")
        return hist, _stats_dicts(raw)


def _profile_meta(profiles: np.ndarray):
    reff = profiles[0, :, 0]
    return int(np.all(np.diff(reff) >= 0)), float(reff.max())


def flux_from_histograms(tables, profiles, hist, concentration: float, flux=None, boundary=None, meta=None):
    """Stage 2 from millimetre histograms, one launch, stream-ordered.

    *tables*: one ``LineTables`` per channel; *profiles*: CUDA float64 tensor (n_slices, K, 3) [Reff, T_e, n_e] sharing
    their Reff column (*meta* = ``(sorted, reff_max)`` of that column, taken from the tensor - a sync - when not given);
    *hist*: CUDA (n_channels, 4096).  Returns ``(flux (n_channels, n_slices, 3 or T+1), boundary (n_channels,))``."""
    import torch

    lib = _lib.load()
    dev = hist.device
    n_ch = len(tables)
    n_slices, K = int(profiles.shape[0]), int(profiles.shape[1])
    stride = max(t.n_out for t in tables)
    if meta is None:
        col = profiles[0, :, 0]
        meta = (int(bool((col[1:] >= col[:-1]).all().item())), float(col.max().item()))
    with torch.cuda.device(dev):
        if flux is None:
            flux = torch.zeros((n_ch, n_slices, stride), dtype=torch.float64, device=dev)
        if boundary is None:
            boundary = torch.zeros(n_ch, dtype=torch.float64, device=dev)
        rc = lib.com_histograms_flux(n_ch, _handles(tables), C.c_void_p(profiles.data_ptr()), n_slices, K, meta[0], meta[1],
                                     float(concentration), C.c_void_p(hist.data_ptr()), C.c_void_p(flux.data_ptr()), stride,
                                     C.c_void_p(boundary.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "com_histograms_flux")
    return flux, boundary


def flux_from_peer_buffers(tables, profiles, peer_ptrs, n_counters: int, concentration: float, flux, boundary, hist_sum=None,
                           counters_sum=None, meta=None):
    """The exchange fused into Stage 2 (``com_histograms_flux_peers``): *peer_ptrs* are the device addresses of every
    rank's packed ``[n_channels * 4096 | n_channels * n_counters]`` float64 buffer as mapped into this process (peer /
    symmetric memory; on one GPU any device buffers will do - that is how the kernel is tested).  The kernel sums them
    in rank order and evaluates the flux; *hist_sum* / *counters_sum* receive the global sums.  Stream-ordered; the
    caller orders the ranks (see ``VolumePipeline``)."""
    import torch

    lib = _lib.load()
    dev = flux.device
    n_ch = len(tables)
    n_slices, K = int(profiles.shape[0]), int(profiles.shape[1])
    if meta is None:
        col = profiles[0, :, 0]
        meta = (int(bool((col[1:] >= col[:-1]).all().item())), float(col.max().item()))
    ptrs = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
    with torch.cuda.device(dev):
        rc = lib.com_histograms_flux_peers(
            n_ch, _handles(tables), C.c_void_p(profiles.data_ptr()), n_slices, K, meta[0], meta[1], float(concentration),
            len(peer_ptrs), ptrs, int(n_counters), C.c_void_p(hist_sum.data_ptr()) if hist_sum is not None else None,
            C.c_void_p(counters_sum.data_ptr()) if counters_sum is not None else None, C.c_void_p(flux.data_ptr()),
            int(flux.shape[-1]), C.c_void_p(boundary.data_ptr()) if boundary is not None else None,
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "com_histograms_flux_peers")
    return flux, boundary


def flux_from_histograms_host(tables, profiles: np.ndarray, hist: np.ndarray, concentration: float, device=None):
    """Host-buffer form of ``flux_from_histograms``: host arrays in, ``(flux (n_channels, n_slices, T+1), boundary)`` out."""
    import torch

    lib = _lib.load()
    profiles = np.ascontiguousarray(profiles, dtype=np.float64)
    if profiles.ndim == 2:
        profiles = profiles[None]
    hist = np.ascontiguousarray(hist, dtype=np.float64)
    n_ch, n_slices, K = len(tables), profiles.shape[0], profiles.shape[1]
    stride = max(t.n_out for t in tables)
    flux = np.zeros((n_ch, n_slices, stride))
    boundary = np.zeros(n_ch)
    with torch.cuda.device(torch.device(device if device is not None else "cuda")):
        rc = lib.com_histograms_flux_host(n_ch, _handles(tables), C.c_void_p(profiles.ctypes.data), n_slices, K,
                                          float(concentration), C.c_void_p(hist.ctypes.data), C.c_void_p(flux.ctypes.data),
                                          stride, C.c_void_p(boundary.ctypes.data))
    _lib.check(rc, "com_histograms_flux_host")
    return flux, boundary


def exchange_world(exchange: str, group=None) -> int:
    """Number of ranks a ``VolumePipeline`` built with *exchange* exchanges with: the process group's size, or 1 when no
    group exists or the pipeline is a rank-local side computation (``exchange="none"``) - such a pipeline must never issue
    a collective: the other ranks are not there to answer it."""
    import torch.distributed as dist

    if exchange not in ("nccl", "peer", "auto", "none"):
        raise ValueError("exchange must be 'nccl', 'peer', 'auto' or 'none'")
    if exchange == "none" or not (dist.is_available() and dist.is_initialized()):
        return 1
    return dist.get_world_size(group)


class VolumePipeline:
    """Stage 1 -> exchange -> Stage 2 of one rank's shard, device-resident and free of host synchronisation.

    ``packed`` is the single buffer the ranks exchange: ``[n_channels * 4096 histogram bins | n_channels * N_STATS
    counters]`` as float64 (counts < 2^53 are exact).  ``step()`` enqueues everything on the current stream.

    *exchange* (only matters with more than one rank):

    ``"nccl"``  one ``all_reduce(SUM)`` of ``packed`` (NCCL), then Stage 2 from the reduced histograms;
    ``"peer"``  ``packed`` lives in symmetric memory (``torch.distributed._symmetric_memory``: every rank's buffer is
                mapped into every process of the box over NVLink); after a device-side barrier on the stream the Stage-2
                kernel itself reads the peers' buffers, sums them in rank order and evaluates the flux - one launch, no
                collective call, bit-identical results on every rank.  Two buffers alternate between steps so that one
                barrier per step is enough (a rank can only be one step ahead of a peer that is still reading);
    ``"auto"``  ``"peer"`` when symmetric memory can be set up, else ``"nccl"`` (``exchange_mode`` /
                ``exchange_note`` say which and why);
    ``"none"``  this pipeline holds a whole (unsharded) lattice and exchanges nothing even though a process group exists
                (a rank-local side computation inside a multi-rank job)."""

    def __init__(self, lattice, scenes, tables, profile: np.ndarray, concentration_percent: float = 2.0, chunk: int = 1 << 26,
                 candidate_fraction: float = 0.0, device=None, group=None, exchange: str = "nccl"):
        import torch
        import torch.distributed as dist

        self.volume = ObservationVolume(lattice, chunk, candidate_fraction, device)
        dev = self.volume.device
        self.scenes, self.tables = list(scenes), list(tables)
        n_ch = len(self.scenes)
        self.n_channels = n_ch
        self.conc = concentration_percent / 100.0
        prof = np.ascontiguousarray(profile, dtype=np.float64)
        if prof.ndim == 2:
            prof = prof[None]
        self.meta = _profile_meta(prof)
        self.group = group
        self.world = exchange_world(exchange, group)
        n_packed = n_ch * (_lib.COM_MM_BINS + _lib.N_STATS)
        self.exchange_mode, self.exchange_note, self._symm, self._peer_ptrs, self._slot = "none", "one rank", None, None, 0
        if self.world > 1:
            self.exchange_mode, self.exchange_note = "nccl", "requested"
            if exchange in ("peer", "auto"):
                try:
                    import torch.distributed._symmetric_memory as symm_mem

                    with torch.cuda.device(dev):
                        buf = symm_mem.empty(2 * n_packed, dtype=torch.float64, device=dev)
                        buf.zero_()
                        self._symm = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                    self._symm_buf = buf
                    base = [int(p) for p in self._symm.buffer_ptrs]
                    self._peer_ptrs = [[b + slot * n_packed * 8 for b in base] for slot in (0, 1)]
                    self.exchange_mode, self.exchange_note = "peer", "symmetric memory over NVLink, reduction fused into Stage 2"
                except Exception as exc:  # noqa: BLE001 - any failure of the optional transport falls back to NCCL, loudly
                    if exchange == "peer":
                        raise
                    self.exchange_note = f"symmetric memory unavailable ({type(exc).__name__}: {str(exc)[:120]})"
        with torch.cuda.device(dev):
            self.profiles = torch.from_numpy(prof.copy()).to(dev)
            if self.exchange_mode == "peer":
                self._packed2 = [self._symm_buf[:n_packed], self._symm_buf[n_packed:]]
                self.global_packed = torch.zeros(n_packed, dtype=torch.float64, device=dev)
            else:
                self._packed2 = [torch.zeros(n_packed, dtype=torch.float64, device=dev)]
                self.global_packed = self._packed2[0]  # the all-reduce is in place
            self._bind(0)
            self.stats = torch.zeros((n_ch, _lib.N_STATS), dtype=torch.int64, device=dev)
            self.flux = torch.zeros((n_ch, prof.shape[0], max(t.n_out for t in self.tables)), dtype=torch.float64, device=dev)
            self.boundary = torch.zeros(n_ch, dtype=torch.float64, device=dev)

    def _bind(self, slot: int):
        n_ch = self.n_channels
        self._slot = slot
        self.packed = self._packed2[slot]
        self.hist = self.packed[: n_ch * _lib.COM_MM_BINS].view(n_ch, _lib.COM_MM_BINS)
        self.counters = self.packed[n_ch * _lib.COM_MM_BINS:].view(n_ch, _lib.N_STATS)
        g = self.global_packed
        self.global_hist = g[: n_ch * _lib.COM_MM_BINS].view(n_ch, _lib.COM_MM_BINS)
        self.global_counters = g[n_ch * _lib.COM_MM_BINS:].view(n_ch, _lib.N_STATS)

    def stage1(self):
        self.volume.histograms(self.scenes, self.hist, self.stats)
        self.counters.copy_(self.stats)  # int64 -> float64, exact below 2^53

    def exchange(self):
        """NCCL form of the path's single exchange: one all-reduce(SUM) of the packed buffer."""
        import torch.distributed as dist

        if self.exchange_mode == "nccl":
            dist.all_reduce(self.packed, group=self.group)

    def stage2(self):
        flux_from_histograms(self.tables, self.profiles, self.global_hist, self.conc, self.flux, self.boundary, self.meta)

    def exchange_stage2_fused(self):
        """Peer form: barrier on the stream, then ONE launch that reads every rank's buffer over NVLink, reduces and
        evaluates; the next step fills the other buffer."""
        self._symm.barrier(channel=0)
        flux_from_peer_buffers(self.tables, self.profiles, self._peer_ptrs[self._slot], _lib.N_STATS, self.conc, self.flux,
                               self.boundary, self.global_hist, self.global_counters, self.meta)
        self._bind(1 - self._slot)

    def step(self):
        self.stage1()
        if self.exchange_mode == "peer":
            self.exchange_stage2_fused()
        else:
            self.exchange()
            self.stage2()
        return self.flux

    def global_stats(self) -> list[dict]:
        """Counters summed over the ranks (after the exchange); ``candidate_capacity`` and ``overflow`` are sums too -
        overflow != 0 means some launch of some rank overflowed."""
        raw = self.global_counters.cpu().numpy().round().astype(np.uint64)
        return _stats_dicts(raw)

    def close(self):
        self.volume.close()
