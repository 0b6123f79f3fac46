"""Device-resident Stage 1 -> Stage 2 pipeline of one channel, free of host synchronisation.

    visibility (K1a filter + K1b exact)  ->  compaction of the visible voxels  ->  max mapped Reff
    [-> all-reduce(MAX) over ranks]      ->  emissivity integration (K2)       [-> all-reduce(SUM) of the flux]

Everything is enqueued on the current CUDA stream through the C ABI; the only values that ever leave the device
are the (T+1) flux numbers the caller reads at the end.  This is what ``bench.py`` times and what a sweep over
voxel grids would call; the file-level drop-ins (``Simulation``, ``Emissivity``) go through the same entry points
but stop in between to build the reference's DataFrames and files.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .emissivity.engine import LineTables, _ptr
from .geometry.engine import ChannelScene, VisibilityPlan
from .geometry.mesh_calculation import Lattice

#: Lyman-alpha lines of the four monitor channels (reader.py:469-474)
LYMAN_ALPHA = {"B": ("Z4", 48.6), "C": ("Z5", 33.7), "N": ("Z6", 24.8), "O": ("Z7", 19.0)}


def reff_to_mm(reff_table):
    """uint16 millimetre bins of a Reff [m] tensor (3-decimal Reff files are exact in this form; NaN -> 0xFFFF)."""
    import torch

    mm = torch.round(reff_table * 1000.0)
    mm = torch.where(torch.isnan(mm), torch.full_like(mm, 65535.0), mm.clamp(0, 65534))
    return mm.to(torch.int32).to(torch.uint16) if hasattr(torch, "uint16") else mm.to(torch.int32).to(torch.int16)


class ChannelPipeline:
    def __init__(self, scene: ChannelScene, tables: LineTables, lattice: Lattice, reff_table, begin: int, end: int,
                 concentration_percent: float = 2.0, device=None, stage2: str = "histogram"):
        """*reff_table*: CUDA float64 tensor with Reff [m] of every lattice voxel (NaN outside the LCFS).

        *stage2* = "histogram" (default): the flux is formed from the millimetre-bin weight histogram of the dense
        per-voxel weights (one 10 B/voxel pass, no compaction) - sum_v eps(k_v) w_v == sum_k H[k] eps(k);
        "per_voxel": compaction + the per-voxel kernel the ``Emissivity`` drop-in uses.
        """
        import torch

        self._lib = _lib.load()
        self.scene, self.tables = scene, tables
        self.plan: VisibilityPlan = scene.plan(lattice, begin, end, device=device)
        dev = self.plan.device
        self.device = dev
        n = end - begin
        self.n = n
        self.reff_table = reff_table
        self.conc = concentration_percent / 100.0
        with torch.cuda.device(dev):
            self.idx = torch.empty(n, dtype=torch.int64, device=dev)
            self.w_compact = torch.empty(n, dtype=torch.float64, device=dev)
            self.count = torch.zeros(1, dtype=torch.int64, device=dev)
            self.cws_bytes = self._lib.com_compact_workspace_bytes(n)
            self.cws = torch.empty(self.cws_bytes, dtype=torch.uint8, device=dev)
            self.mapped_max = torch.zeros(1, dtype=torch.float64, device=dev)
            self.flux = torch.zeros(tables.n_out, dtype=torch.float64, device=dev)
            self.ews_bytes = self._lib.com_emissivity_workspace_bytes(n)
            self.ews = torch.empty(self.ews_bytes, dtype=torch.uint8, device=dev)
            self.stage2 = stage2
            if stage2 == "histogram":
                if begin != 0 or end != reff_table.numel():
                    reff_table = reff_table[begin:end]
                self.reff_mm = reff_to_mm(reff_table).contiguous()
                self.hist_mm = torch.zeros(_lib.COM_MM_BINS, dtype=torch.float64, device=dev)
                self.hist_rows = None
                self.boundary = torch.zeros(1, dtype=torch.float64, device=dev)
        self._profile = None
        # filter, exact, then either (mm histogram, rows, sweep) or (3 x compaction, reff max, emissivity)
        self.kernels_per_launch = 2 + (3 if stage2 == "histogram" else 5)

    def set_profile(self, profile: np.ndarray):
        """(K,3) [Reff, T_e, n_e] host array -> device."""
        import torch

        prof = np.ascontiguousarray(profile, dtype=np.float64)
        self._profile = torch.from_numpy(np.array(prof, dtype=np.float64, order="C", copy=True)).to(self.device)
        self._K = len(prof)
        self._sorted = int(np.all(np.diff(prof[:, 0]) >= 0))
        self._pmax = float(prof[:, 0].max())
        if self.stage2 == "histogram":
            self.hist_rows = torch.zeros(self._K, dtype=torch.float64, device=self.device)

    def _stream(self):
        import torch

        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def launch_geometry(self):
        """Stage 1, then either this rank's millimetre-bin weight histogram (``hist_mm``; ranks all-reduce SUM it) or
        compaction and this rank's maximum mapped Reff (``mapped_max``; ranks all-reduce MAX it)."""
        s = self._stream()
        self.plan.launch()
        if self.stage2 == "histogram":
            self.hist_mm.zero_()
            rc = self._lib.com_weight_histogram_mm(_ptr(self.reff_mm), _ptr(self.plan.weight), self.n, None,
                                                   _ptr(self.hist_mm), s)
            _lib.check(rc, "com_weight_histogram_mm")
            return
        rc = self._lib.com_compact_visible(_ptr(self.plan.nrays), _ptr(self.plan.weight), self.n, self.plan.begin,
                                           _ptr(self.idx), _ptr(self.w_compact), _ptr(self.count), _ptr(self.cws),
                                           self.cws_bytes, s)
        _lib.check(rc, "com_compact_visible")
        rc = self._lib.com_reff_max(_ptr(self.reff_table), None, _ptr(self.idx), self.n, _ptr(self.count),
                                    _ptr(self.mapped_max), s)
        _lib.check(rc, "com_reff_max")

    def launch_emissivity(self):
        """Stage 2 over the compacted voxels; ``mapped_max`` may have been max-reduced over ranks in between."""
        if self._profile is None:
            raise RuntimeError("set_profile() first")
        if self.stage2 == "histogram":
            s = self._stream()
            rc = self._lib.com_histogram_rows(_ptr(self._profile), self._K, self._sorted, self._pmax, _ptr(self.hist_mm), None,
                                              _ptr(self.hist_rows), _ptr(self.boundary), s)
            _lib.check(rc, "com_histogram_rows")
            rc = self._lib.com_emissivity_sweep(self.tables.handle, _ptr(self._profile), 1, self._K, _ptr(self.hist_rows),
                                                self.conc, _ptr(self.flux), s)
            _lib.check(rc, "com_emissivity_sweep")
            return
        rc = self._lib.com_emissivity_integrate(
            self.tables.handle, _ptr(self._profile), self._K, self._sorted, self._pmax, self.conc,
            _ptr(self.reff_table), None, _ptr(self.idx), _ptr(self.w_compact), self.n, _ptr(self.count),
            _ptr(self.mapped_max), _ptr(self.flux), None, None, _ptr(self.ews), self.ews_bytes, self._stream())
        _lib.check(rc, "com_emissivity_integrate")

    def launch(self):
        self.launch_geometry()
        self.launch_emissivity()
        return self.flux

    def check(self) -> dict:
        """Statistics of the last Stage-1 launch (one host synchronisation).  Raises ``ComonitorError(COM_E_OVERFLOW)`` when
        the candidate list overflowed: the launch then accumulated nothing (the exact kernel skips an overflowed list), so
        weights, histogram and flux of that launch are zeros, not partial sums - call this once after warm-up, or size the
        plan with ``candidate_fraction`` from a measured count."""
        st = self.plan.stats()
        if st["overflow"]:
            raise _lib.ComonitorError(_lib.COM_E_OVERFLOW, "ChannelPipeline",
                                      f"candidate list overflow: {st['candidates']} candidates > capacity {st['candidate_capacity']}")
        return st
