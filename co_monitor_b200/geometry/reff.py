"""Effective-radius (Reff) grids without the VMEC web service.

The reference obtains Reff per voxel from a W7-X REST service (co_monitor/geometry/reff_VMEC.py:18,
85-136; no network here) and stores it as ``input_files/geometry/reff/Reff_coordinates-<s>_mm.dat``
(``idx;x [m];y [m];z [m];Reff [m]``, coordinates rounded to 4 and Reff to 3 decimals, ``NaN`` outside
the last closed flux surface; reff_VMEC.py:138-176).  The 10 mm file both entry scripts ask for is a
missing blob of the reference checkout (.MISSING_LARGE_BLOBS), only 15/20/25/30 mm grids ship.

This module provides the replacements the north-star allows:
  * ``upsample_reff``   trilinear resampling of a shipped grid onto a finer CuboidMesh lattice.  All lattices
                        span the same cuboid (linspace end points are shared), so the interpolation runs in
                        fractional lattice-index space; a voxel whose interpolation stencil touches a NaN
                        is NaN (outside the LCFS), exactly how the reference treats NaN rows (reader.py:176).
  * ``QuadraticReff``   an analytic flux-surface model fitted to a shipped grid (Reff^2 as a quadratic form of
                        position: nested elliptic cylinders), evaluable for 1e8-1e9 voxel lattices without a file.
Both run on the device (``resample_reff_device``, ``QuadraticReff.on_device``: csrc/reff.cu through
``com_reff_resample`` / ``com_reff_quadratic``) and produce the uint16 millimetre bins Stage 2 reads for any
(sharded) range of a 1e8-1e9 voxel lattice, or the 3-decimal doubles of the reference's file.  The NumPy statement the
resampling kernel is tested against is test infrastructure (oracle/reff.py); ``QuadraticReff.__call__`` is the host
form of the analytic model (used to fit and to check it).
"""
from __future__ import annotations

import contextlib
import io
from pathlib import Path

import numpy as np
import pandas as pd

from .json_reader import input_root
from .mesh_calculation import CuboidMesh

REFF_COLUMNS = ["idx", "x [m]", "y [m]", "z [m]", "Reff [m]"]


def reff_path(name_or_spacing, root: Path | None = None) -> Path:
    root = Path(root) if root is not None else input_root()
    name = name_or_spacing if isinstance(name_or_spacing, str) else f"Reff_coordinates-{name_or_spacing}_mm"
    return root / "geometry" / "reff" / f"{name}.dat"


def load_reff(name_or_spacing, root: Path | None = None) -> pd.DataFrame:
    return pd.read_csv(reff_path(name_or_spacing, root), sep=";")


def _quiet_mesh(spacing) -> CuboidMesh:
    with contextlib.redirect_stdout(io.StringIO()):
        return CuboidMesh(spacing)


def source_volume(source_spacing: int = 15, root: Path | None = None):
    """A shipped Reff grid as a ``(ny, nx, nz)`` array (z fastest: the reshape of the file's row order) and its lattice."""
    src = load_reff(source_spacing, root)
    ls = _quiet_mesh(source_spacing).lattice
    if len(src) != ls.n_points:
        raise ValueError(f"{reff_path(source_spacing, root)} has {len(src)} rows, lattice has {ls.n_points}")
    return src["Reff [m]"].to_numpy(dtype=float).reshape(ls.ny, ls.nx, ls.nz), ls


def _device_outputs(lattice, begin, end, want_reff, device):
    import torch

    if not torch.cuda.is_available():
        from .. import _lib

        raise _lib.ComonitorError(_lib.COM_E_CUDA, "Reff provider", "no CUDA device: the accelerated path has no CPU fallback")
    device = torch.device(device if device is not None else "cuda")
    end = lattice.n_points if end is None else end
    if not 0 <= begin <= end <= lattice.n_points:
        raise ValueError(f"voxel range [{begin},{end}) outside [0,{lattice.n_points})")
    with torch.cuda.device(device):
        mm = torch.empty(end - begin, dtype=torch.uint16, device=device)
        reff = torch.empty(end - begin, dtype=torch.float64, device=device) if want_reff else None
        stream = torch.cuda.current_stream().cuda_stream
    return device, end, mm, reff, stream


def resample_reff_device(target_lattice, source_spacing: int = 15, begin: int = 0, end: int | None = None,
                         want_reff: bool = False, device=None, root: Path | None = None, source=None):
    """``interpolate_reff`` on the GPU for the voxels [begin, end) of *target_lattice* (shard-local indices when the
    lattice is a shard): ``(reff_mm uint16, reff float64 | None)`` CUDA tensors.  *source*: a ``(vol, lattice)`` pair
    to resample instead of a shipped file."""
    import ctypes as C

    import torch

    from .. import _lib
    from .engine import lattice_struct

    device, end, mm, reff, stream = _device_outputs(target_lattice, begin, end, want_reff, device)
    vol, ls = source if source is not None else source_volume(source_spacing, root)
    src = torch.from_numpy(np.array(vol, dtype=np.float64, order="C", copy=True)).to(device)
    lat = lattice_struct(target_lattice)
    with torch.cuda.device(device):
        rc = _lib.load().com_reff_resample(C.byref(lat), begin, end, C.c_void_p(src.data_ptr()), ls.nx, ls.ny, ls.nz,
                                           C.c_void_p(mm.data_ptr()), C.c_void_p(reff.data_ptr()) if want_reff else None,
                                           C.c_void_p(stream))
    _lib.check(rc, "com_reff_resample")
    src.record_stream(torch.cuda.current_stream(device))
    return mm, reff


def interpolate_reff(target_lattice, source_spacing: int = 15, root: Path | None = None, device=None,
                     chunk: int = 1 << 27) -> np.ndarray:
    """Reff [m] (3 decimals, NaN outside the LCFS) on every voxel of *target_lattice*, flat index order, as a host
    array.  Computed by the device kernel (``com_reff_resample``) in ranges of *chunk* voxels; like every compute
    entry point of this package it raises without a CUDA device (the NumPy statement the kernel is tested against
    lives in oracle/reff.py, test infrastructure)."""
    n = target_lattice.n_points
    source = source_volume(source_spacing, root)
    out = np.empty(n, dtype=np.float64)
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        _, reff = resample_reff_device(target_lattice, begin=lo, end=hi, want_reff=True, device=device, source=source)
        out[lo:hi] = reff.cpu().numpy()
    return out


def upsample_reff(source_spacing: int = 15, target_spacing: int = 10, root: Path | None = None) -> pd.DataFrame:
    """Reff file contents for the *target_spacing* lattice by trilinear interpolation of the shipped
    *source_spacing* grid (same columns and rounding as reff_VMEC.py:138-165)."""
    mt = _quiet_mesh(target_spacing)
    reff = interpolate_reff(mt.lattice, source_spacing, root)
    rounded = np.round(mt.outer_cube_mesh / 1000.0, decimals=4)
    return pd.DataFrame(
        {
            "idx": np.arange(mt.lattice.n_points),
            "x [m]": rounded[:, 0],
            "y [m]": rounded[:, 1],
            "z [m]": rounded[:, 2],
            "Reff [m]": reff,
        }
    )


def write_reff_file(df: pd.DataFrame, path) -> Path:
    """Same CSV dialect as reff_VMEC.py:167-176."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    df.to_csv(path, sep=";", index=False, na_rep="NaN")
    return path


class QuadraticReff:
    """Analytic flux surfaces: ``Reff(r)^2 = (r - r0)^T A (r - r0)`` (nested elliptic cylinders), least-squares
    fitted to a shipped grid; NaN beyond ``reff_max`` (the LCFS of the fitted grid)."""

    def __init__(self, source_spacing: int = 15, root: Path | None = None):
        src = load_reff(source_spacing, root)
        pts = _quiet_mesh(source_spacing).outer_cube_mesh / 1000.0
        reff = src["Reff [m]"].to_numpy(dtype=float)
        ok = np.isfinite(reff)
        x, y, z = (pts[ok, i] for i in range(3))
        design = np.column_stack((x * x, y * y, z * z, x * y, x * z, y * z, x, y, z, np.ones_like(x)))
        self.coef, *_ = np.linalg.lstsq(design, reff[ok] ** 2, rcond=None)
        self.reff_max = float(np.nanmax(reff))

    def __call__(self, points_mm: np.ndarray) -> np.ndarray:
        p = np.asarray(points_mm, dtype=float) / 1000.0
        x, y, z = p[:, 0], p[:, 1], p[:, 2]
        c = self.coef
        q = (c[0] * x * x + c[1] * y * y + c[2] * z * z + c[3] * x * y + c[4] * x * z + c[5] * y * z
             + c[6] * x + c[7] * y + c[8] * z + c[9])
        r = np.sqrt(np.maximum(q, 0.0))
        return np.where(r <= self.reff_max, r, np.nan)

    def on_device(self, lattice, begin: int = 0, end: int | None = None, want_reff: bool = False, device=None):
        """The model on the voxels [begin, end) of *lattice*, evaluated on the GPU from the lattice closed form:
        ``(reff_mm uint16, reff float64 rounded to 3 decimals | None)`` CUDA tensors."""
        import ctypes as C

        from .. import _lib
        from .engine import lattice_struct

        device, end, mm, reff, stream = _device_outputs(lattice, begin, end, want_reff, device)
        import torch

        coef = (C.c_double * 10)(*[float(c) for c in self.coef])
        lat = lattice_struct(lattice)
        with torch.cuda.device(device):
            rc = _lib.load().com_reff_quadratic(C.byref(lat), begin, end, coef, self.reff_max, C.c_void_p(mm.data_ptr()),
                                                C.c_void_p(reff.data_ptr()) if want_reff else None, C.c_void_p(stream))
        _lib.check(rc, "com_reff_quadratic")
        return mm, reff
