"""TEST INFRASTRUCTURE - CPU statement of the Reff provider.

The reference obtains Reff per voxel from the W7-X VMEC web service (co_monitor/geometry/reff_VMEC.py:85-136) and
writes ``Reff_coordinates-<s>_mm.dat`` with Reff rounded to 3 decimals and ``NaN`` outside the last closed flux
surface (reff_VMEC.py:138-176).  There is no network here; the product replaces the service by trilinear resampling of
a shipped grid on the device (csrc/reff.cu).  This file is the NumPy statement those kernels are checked against -
it has no counterpart in the reference beyond the file format, so it pins nothing by itself: "parity unpinned",
extension only.  What IS pinned: every lattice of the reference spans the same cuboid with shared end points
(mesh_calculation.py:119-143), so index-space interpolation reproduces a shipped grid on its own lattice exactly
(tests/test_reff_provider.py).

  trilinear()        values at selected (ix, iy, iz) of a target lattice; per voxel
                         f = i * ((n_src - 1) / (n_tgt - 1)),  i0 = min(floor(f), n_src - 2),  t = f - i0
                     eight corner terms added in the order (dy, dx, dz) = 000, 001, 010, ..., each term
                     (wy * wx) * wz times the corner value; a corner that is NaN with non-zero weight makes the voxel NaN
  to_mm()            uint16 millimetre bins: rint(Reff * 1000), NaN -> 0xFFFF

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import os

import numpy as np

DEFAULT_INPUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "co_monitor_b200", "input_files")


def shipped_volume(spacing: int = 15, dims=(1350, 800, 250), root=DEFAULT_INPUT):
    """A shipped Reff grid as (ny, nx, nz) (the reshape of the file's row order: z fastest, then x, then y)."""
    import pandas as pd

    path = os.path.join(root, "geometry", "reff", f"Reff_coordinates-{spacing}_mm.dat")
    col = pd.read_csv(path, sep=";").iloc[:, 4].to_numpy(dtype=float)
    nx, ny, nz = (int(d // spacing) for d in dims)
    if len(col) != nx * ny * nz:
        raise ValueError(f"{path}: {len(col)} rows, lattice has {nx * ny * nz}")
    return col.reshape(ny, nx, nz)


def _axis(n_target: int, n_source: int, i: np.ndarray):
    f = i.astype(np.float64) * ((n_source - 1) / (n_target - 1))
    i0 = np.minimum(np.floor(f).astype(np.int64), n_source - 2)
    return i0, f - i0


def trilinear(vol: np.ndarray, target_shape, ix, iy, iz, decimals: int | None = 3) -> np.ndarray:
    """Reff at the lattice nodes (ix, iy, iz) of a target lattice of *target_shape* = (nx, ny, nz)."""
    ny_s, nx_s, nz_s = vol.shape
    nx, ny, nz = target_shape
    ix, iy, iz = (np.asarray(a, dtype=np.int64) for a in (ix, iy, iz))
    (y0, ty), (x0, tx), (z0, tz) = _axis(ny, ny_s, iy), _axis(nx, nx_s, ix), _axis(nz, nz_s, iz)
    out = np.zeros(len(ix))
    bad = np.zeros(len(ix), dtype=bool)
    for dy in (0, 1):
        wy = ty if dy else 1 - ty
        for dx in (0, 1):
            wx = tx if dx else 1 - tx
            for dz in (0, 1):
                wz = tz if dz else 1 - tz
                w = wy * wx * wz
                v = vol[y0 + dy, x0 + dx, z0 + dz]
                isnan = np.isnan(v)
                bad |= isnan & (w > 0)
                out += np.where(isnan, 0.0, v) * w
    out[bad] = np.nan
    return np.round(out, decimals) if decimals is not None else out


def on_lattice(target_shape, flat=None, source_spacing: int = 15, decimals: int | None = 3, vol=None) -> np.ndarray:
    """Reff for the GLOBAL flat indices *flat* (all voxels when None) of the lattice of *target_shape* (nx, ny, nz);
    flat = (iy * nx + ix) * nz + iz (mesh_calculation.py:119-159: meshgrid 'xy' + ravel)."""
    nx, ny, nz = target_shape
    vol = shipped_volume(source_spacing) if vol is None else vol
    flat = np.arange(nx * ny * nz, dtype=np.int64) if flat is None else np.asarray(flat, dtype=np.int64)
    iz = flat % nz
    t = flat // nz
    return trilinear(vol, target_shape, t % nx, t // nx, iz, decimals)


def to_mm(reff: np.ndarray) -> np.ndarray:
    mm = np.rint(reff * 1000.0)
    return np.where(np.isnan(mm), 65535, np.clip(mm, 0, 65534)).astype(np.uint16)


def file_frame(target_spacing, source_spacing: int = 15, dims=(1350, 800, 250)):
    """Contents of ``Reff_coordinates-<target_spacing>_mm.dat`` in the reference's format (reff_VMEC.py:138-165:
    ``idx``, coordinates in metres rounded to 4 decimals, Reff rounded to 3) with Reff resampled from a shipped grid."""
    import pandas as pd

    from . import geometry_setup as gs

    db = gs.load_db()
    shape = gs.mesh_shape(db, target_spacing, dims)
    pts = np.round(gs.voxel_mesh(db, target_spacing, dims) / 1000.0, decimals=4)
    reff = on_lattice(shape, source_spacing=source_spacing)
    return pd.DataFrame({"idx": np.arange(len(pts)), "x [m]": pts[:, 0], "y [m]": pts[:, 1], "z [m]": pts[:, 2],
                         "Reff [m]": reff})


def write_file(df, path):
    """Same CSV dialect as reff_VMEC.py:167-176."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    df.to_csv(path, sep=";", index=False, na_rep="NaN")
    return path
