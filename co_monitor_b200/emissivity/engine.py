"""Stage-2 engine: Python host side of the sm_100a emissivity kernels.

``LineTables`` owns the device-resident fractional-abundance / PEC tables of one spectral line
(a ``ComTables`` handle); ``integrate`` is the per-voxel pass behind ``Emissivity``;
``weight_histogram`` + ``sweep`` are the factorised form used for profile time-slice sweeps and for
grids too large to keep per-voxel outputs.  torch supplies device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np

from .. import _lib
from .fractional_abundance import FractionalAbundance
from .pec import PEC


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _device_f64(x, dev):
    """Contiguous float64 CUDA tensor from a host array or a tensor of any layout (the C ABI takes dense row-major)."""
    import torch

    if torch.is_tensor(x):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.array(x, dtype=np.float64, order="C", copy=True)).to(dev)


class LineTables:
    """FA + PEC tables of (element, ion_state, wavelength, transitions) on the device."""

    def __init__(self, element, ion_state, wavelength, transitions, input_root=None, interp_step=2000,
                 pec_interpolation: str = "reference"):
        """*pec_interpolation*: "reference" = what the reference does (linear interp2d on a 2000 x 2000 linspace table read
        at the nearest node, pec.py:153-187 + reader.py:316-341); "loglog" = bilinear in log-log space through the ADAS
        nodes at the row's own (ne, Te) - an extension, off by default, not comparable with reference output."""
        if pec_interpolation not in ("reference", "loglog"):
            raise ValueError("pec_interpolation must be 'reference' or 'loglog'")
        self.pec_interpolation = pec_interpolation
        self._lib = _lib.load()
        self.element, self.ion_state, self.wavelength = element, ion_state, wavelength
        self.transitions = list(transitions)
        if not 1 <= len(self.transitions) <= _lib.COM_MAX_TRANSITIONS:
            raise ValueError(f"1..{_lib.COM_MAX_TRANSITIONS} transitions supported")
        self.fa = FractionalAbundance(element, ion_state, input_root=input_root)
        frac = self.fa.df_interpolated_frac_ab
        self.pecs = [PEC(element, wavelength, ion_state[-1], tr, interp_step, input_root=input_root)
                     for tr in self.transitions]
        self._keep = []

        def hold(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            self._keep.append(a)
            return _lib.as_double_ptr(a)

        d = _lib.ComTablesDesc()
        d.n_fa = len(frac)
        d.n_transitions = len(self.transitions)
        d.fa_T = hold(frac.iloc[:, 0].to_numpy())
        d.fa_ion = hold(frac[ion_state].to_numpy())
        d.fa_last = hold(frac.iloc[:, -1].to_numpy())
        for t, (tr, pec) in enumerate(zip(self.transitions, self.pecs)):
            p = d.pec[t]
            p.n_ne, p.n_te, p.n_grid = pec.ne_nodes_nr, pec.te_nodes_nr, pec.interp_step
            p.use_fully_stripped = 1 if "RECOM" in tr else 0  # reader.py:355: `if "RECOM" in pec`
            p.interpolation = 1 if pec_interpolation == "loglog" else 0
            p.ne_nodes, p.te_nodes = hold(pec.ne_nodes), hold(pec.te_nodes)
            p.table = hold(pec.pec_data_array)
            p.ne_grid, p.te_grid = hold(pec.ne_grid), hold(pec.te_grid)
        self.desc = d
        self.handle = C.c_void_p()
        _lib.check(self._lib.com_tables_create(C.byref(d), C.byref(self.handle)), "com_tables_create")

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.com_tables_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def n_out(self) -> int:
        return len(self.transitions) + 1

    # ---- per-voxel pass ---------------------------------------------------------------------------------
    def integrate(self, profile: np.ndarray, concentration: float, reff, w, per_voxel: bool = True, device=None):
        """*profile* (K,3) [Reff, T_e, n_e] host array; *reff*, *w* host arrays or CUDA float64 tensors.

        Returns dict(flux (T+1,) ndarray, boundary float, valid bool ndarray|None, columns (n, 4+2T+1) ndarray|None).
        """
        import torch

        if not torch.cuda.is_available():
            raise _lib.ComonitorError(_lib.COM_E_CUDA, "LineTables.integrate",
                                      "no CUDA device: the accelerated path has no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        prof = np.ascontiguousarray(profile, dtype=np.float64)
        K = len(prof)
        sorted_flag = int(np.all(np.diff(prof[:, 0]) >= 0))
        with torch.cuda.device(dev):
            prof_d = _device_f64(prof, dev)
            reff_d = _device_f64(reff, dev)
            w_d = _device_f64(w, dev)
            n = int(reff_d.numel())
            nt = len(self.transitions)
            flux = torch.zeros(nt + 1, dtype=torch.float64, device=dev)
            cols = torch.zeros((n, 4 + 2 * nt + 1), dtype=torch.float64, device=dev) if per_voxel else None
            valid = torch.zeros(n, dtype=torch.uint8, device=dev) if per_voxel else None
            ws_bytes = self._lib.com_emissivity_workspace_bytes(n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            rc = self._lib.com_emissivity_integrate(
                self.handle, _ptr(prof_d), K, sorted_flag, float(prof[:, 0].max()), float(concentration),
                _ptr(reff_d), None, None, _ptr(w_d), n, None, None, _ptr(flux), _ptr(cols), _ptr(valid), _ptr(ws), ws_bytes,
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "com_emissivity_integrate")
            boundary = float(ws[256 + 8:256 + 16].view(torch.float64).item())
            return dict(flux=flux.cpu().numpy(), boundary=boundary,
                        valid=valid.cpu().numpy().astype(bool) if per_voxel else None,
                        columns=cols.cpu().numpy() if per_voxel else None)

    # ---- factorised form ----------------------------------------------------------------------------------
    def weight_histogram(self, profile_reff: np.ndarray, reff, w, device=None, reff_mm=None, reff_index=None,
                         n_device=None, mapped_reff_max=None):
        """H[k] = sum of w over the voxels mapped to profile row k.  Returns (hist CUDA tensor (K,), boundary float)."""
        import torch

        dev = torch.device(device if device is not None else "cuda")
        K = len(profile_reff)
        prof = np.zeros((K, 3))
        prof[:, 0] = profile_reff
        sorted_flag = int(np.all(np.diff(prof[:, 0]) >= 0))
        with torch.cuda.device(dev):
            prof_d = _device_f64(prof, dev)
            if reff is not None:
                reff = _device_f64(reff, dev)
            w_d = _device_f64(w, dev)
            n = int(w_d.numel())
            hist = torch.zeros(K, dtype=torch.float64, device=dev)
            bnd = torch.zeros(1, dtype=torch.float64, device=dev)
            ws_bytes = self._lib.com_emissivity_workspace_bytes(n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            rc = self._lib.com_weight_histogram(
                _ptr(prof_d), K, sorted_flag, float(prof[:, 0].max()), _ptr(reff), _ptr(reff_mm), _ptr(reff_index),
                _ptr(w_d), n, _ptr(n_device), _ptr(mapped_reff_max), _ptr(hist), _ptr(bnd), _ptr(ws), ws_bytes,
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "com_weight_histogram")
            return hist, float(bnd.item())

    def sweep(self, profiles, hist, concentration: float):
        """flux (n_slices, T+1) CUDA tensor for *profiles* (n_slices, K, 3) (host array or CUDA tensor)."""
        import torch

        dev = hist.device
        with torch.cuda.device(dev):
            prof_d = _device_f64(profiles, dev)  # (an np.stack of DataFrame.to_numpy() blocks is not C-ordered)
            n_slices, K = int(prof_d.shape[0]), int(prof_d.shape[1])
            if K != hist.numel():
                raise ValueError("profiles and histogram disagree on the number of profile rows")
            flux = torch.empty((n_slices, self.n_out), dtype=torch.float64, device=dev)
            rc = self._lib.com_emissivity_sweep(self.handle, _ptr(prof_d), n_slices, K, _ptr(hist), float(concentration),
                                                _ptr(flux), C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "com_emissivity_sweep")
            return flux


@lru_cache(maxsize=16)
def cached_line_tables(element, ion_state, wavelength, transitions: tuple, input_root=None,
                       pec_interpolation: str = "reference") -> LineTables:
    return LineTables(element, ion_state, wavelength, list(transitions), input_root=input_root,
                      pec_interpolation=pec_interpolation)
