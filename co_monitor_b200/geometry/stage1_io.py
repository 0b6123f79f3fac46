"""Stage-1 result files at scale.

The reference stores the visible voxels with their weights as a ``;``-separated text file written by dask
(co_monitor/geometry/simulation.py:616-630: ``ddf.to_csv("...slit_<S>*.dat", sep=";", header=True, index=False)``,
the ``*`` replaced by the partition number) and Stage 2 reads it back with pandas (emissivity/reader.py:136-159).
At the default 10 mm grid that is 3-5e4 rows; a 1 mm grid has 3-5e7 per channel, a 0.65 mm grid 1-2e8.

  * ``write_csv``      the same bytes as pandas/dask produce, formatted by the native writer on all host cores
                       (``com_write_stage1_csv``), optionally split into dask-style numbered partitions;
  * ``write_binary``   a side format for 1e8+ rows: 16 bytes per row (int64 flat index + float64 weight) after a
                       128-byte header that carries the lattice, so that the coordinates - which the text file repeats
                       per row - are regenerated bit-identically from the index on reading;
  * ``read``           either format (all partitions of a text file) -> the reference's five-column frame,
                       ordered by ``idx_sel_plas_points``.
"""
from __future__ import annotations

import ctypes as C
import glob
import struct
from pathlib import Path

import numpy as np
import pandas as pd

from .. import _lib

COLUMNS = ["idx_sel_plas_points", "plasma_x", "plasma_y", "plasma_z", "total_intensity_fraction"]
MAGIC = b"COMSTG1\0"
HEADER_BYTES = 128
_HEADER = struct.Struct("<8sIIq3q3d3d3d")  # magic, version, flags, rows, (nx, ny, nz), start, stop, step -> 8+8+8+24+72 = 120


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def write_csv(path, idx, x, y, z, w, partitions: int = 1, n_threads: int = 0) -> list[Path]:
    """Write the rows to *path*; with ``partitions > 1`` *path* must contain ``*`` (dask's placeholder) and the rows
    are split into that many contiguous, nearly equal parts ``...0.dat, ...1.dat, ...``.  Returns the files written."""
    lib = _lib.load()
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    x, y, z, w = _f64(x), _f64(y), _f64(z), _f64(w)
    n = len(idx)
    if not (len(x) == len(y) == len(z) == len(w) == n):
        raise ValueError("columns of different lengths")
    path = str(path)
    if partitions > 1 and "*" not in path:
        raise ValueError("a partitioned file name needs a '*' for the partition number")
    bounds = np.linspace(0, n, max(1, partitions) + 1).astype(np.int64)
    out = []
    for k in range(max(1, partitions)):
        lo, hi = int(bounds[k]), int(bounds[k + 1])
        name = path.replace("*", str(k)) if "*" in path else path
        Path(name).parent.mkdir(parents=True, exist_ok=True)
        rc = lib.com_write_stage1_csv(name.encode(), C.c_void_p(idx[lo:hi].ctypes.data), _lib.as_double_ptr(x[lo:hi]),
                                      _lib.as_double_ptr(y[lo:hi]), _lib.as_double_ptr(z[lo:hi]),
                                      _lib.as_double_ptr(w[lo:hi]), hi - lo, 1, n_threads)
        _lib.check(rc, "com_write_stage1_csv")
        out.append(Path(name))
    return out


def write_binary(path, lattice, idx, w) -> Path:
    """int64 index + float64 weight per row behind a header with the (unsharded) lattice."""
    if lattice.shard_count > 1:
        raise ValueError("write global indices: lattice.local_to_global(idx) with the unsharded lattice")
    idx = np.ascontiguousarray(idx, dtype="<i8")
    w = np.ascontiguousarray(w, dtype="<f8")
    if len(idx) != len(w):
        raise ValueError("columns of different lengths")
    head = _HEADER.pack(MAGIC, 1, 0, len(idx), lattice.nx, lattice.ny, lattice.nz, *map(float, lattice.start),
                        *map(float, lattice.stop), *map(float, lattice.step))
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as fh:
        fh.write(head.ljust(HEADER_BYTES, b"\0"))
        idx.tofile(fh)
        w.tofile(fh)
    return path


def is_binary(path) -> bool:
    with open(path, "rb") as fh:
        return fh.read(len(MAGIC)) == MAGIC


def read_binary(path, mesh=None, coordinates: bool = True) -> pd.DataFrame:
    """*mesh*: the ``CuboidMesh`` the file was computed on (checked against the header); built from the header's
    spacing when omitted.  ``coordinates=False`` skips the three coordinate columns."""
    with open(path, "rb") as fh:
        head = fh.read(HEADER_BYTES)
        magic, version, _flags, rows, nx, ny, nz, *geo = _HEADER.unpack(head[: _HEADER.size])
        if magic != MAGIC or version != 1:
            raise ValueError(f"{path}: not a co_monitor_b200 Stage-1 binary file")
        idx = np.fromfile(fh, dtype="<i8", count=rows)
        w = np.fromfile(fh, dtype="<f8", count=rows)
    if len(idx) != rows or len(w) != rows:
        raise ValueError(f"{path}: truncated ({len(idx)}/{len(w)} of {rows} rows)")
    frame = {"idx_sel_plas_points": idx}
    if coordinates:
        if mesh is None:
            raise ValueError("read_binary(coordinates=True) needs the CuboidMesh the file was computed on")
        lat = mesh.lattice
        same = (lat.nx, lat.ny, lat.nz) == (nx, ny, nz) and np.array_equal(
            np.concatenate((lat.start, lat.stop, lat.step)), np.array(geo))
        if not same:
            raise ValueError(f"{path}: computed on a {nx}x{ny}x{nz} lattice, not on the mesh given")
        pts = mesh.points_at(idx)
        frame.update(plasma_x=pts[:, 0].round(1), plasma_y=pts[:, 1].round(1), plasma_z=pts[:, 2].round(1))
    frame["total_intensity_fraction"] = w
    return pd.DataFrame(frame)


def read(path, mesh=None) -> pd.DataFrame:
    """Text (one file, or every partition matching a name with ``*``) or binary -> frame ordered by voxel index."""
    path = str(path)
    files = sorted(glob.glob(path), key=lambda p: (len(p), p)) if "*" in path else [path]
    if not files:
        raise FileNotFoundError(path)
    parts = [read_binary(f, mesh) if is_binary(f) else pd.read_csv(f, sep=";") for f in files]
    df = parts[0] if len(parts) == 1 else pd.concat(parts, ignore_index=True)
    if not df["idx_sel_plas_points"].is_monotonic_increasing:
        df = df.sort_values("idx_sel_plas_points", kind="stable", ignore_index=True)
    return df
