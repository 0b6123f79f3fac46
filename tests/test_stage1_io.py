"""Stage-1 result files: the native text writer must produce the bytes pandas/dask
produce (simulation.py:616-630) - pinned on the reference's four shipped files - and the binary side format must
round-trip to the same frame.  Host code only: no GPU."""
import contextlib
import ctypes as C
import io
import random
import struct

import numpy as np
import pandas as pd
import pytest

from co_monitor_b200 import _lib
from co_monitor_b200.geometry import stage1_io
from co_monitor_b200.geometry.json_reader import input_root
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh


def _shipped(el):
    return (input_root() / "geometry" / "observed_plasma_volume" / el
            / f"{el}_plasma_coordinates-10_mm_spacing-height_30-length_20-slit_100.dat")


def _columns(path):
    """Columns of a Stage-1 text file with Python's correctly rounded float parser."""
    rows = [line.split(";") for line in path.read_text().splitlines()[1:]]
    idx = np.array([int(r[0]) for r in rows], dtype=np.int64)
    cols = [np.array([float(r[k]) for r in rows]) for k in range(1, 5)]
    return idx, cols


def test_float_formatting_is_python_repr():
    lib = _lib.load()
    buf = C.create_string_buffer(48)
    rnd = random.Random(0)
    special = [0.0, -0.0, 1.0, -1.5, 0.1, 1e16, 1e15, 9999999999999998.0, 1e-4, 9.999e-5, 1e-5, 1.2345678901234568e17,
               1e22, 1e100, 1e-100, 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 2.188959843266816e-06,
               -1234.5, 100.0, 123456.0, float("inf"), -float("inf"), float("nan")]
    values = special + [struct.unpack("d", struct.pack("Q", rnd.getrandbits(64)))[0] for _ in range(20000)]
    values += [rnd.uniform(-3000, 3000) for _ in range(10000)] + [round(rnd.uniform(-3000, 3000), 1) for _ in range(10000)]
    values += [rnd.uniform(0, 1) * 10.0 ** rnd.randint(-12, -3) for _ in range(10000)]
    for v in values:
        n = lib.com_format_float_repr(v, buf, 48)
        assert n > 0 and buf.value.decode() == repr(v), (repr(v), buf.value)
    assert lib.com_format_float_repr(1.0, buf, 2) == _lib.COM_E_BADARG


@pytest.mark.parametrize("el", ["B", "C", "N", "O"])
def test_writer_reproduces_the_shipped_reference_files_byte_for_byte(el, tmp_path):
    src = _shipped(el)
    idx, (x, y, z, w) = _columns(src)
    out = stage1_io.write_csv(tmp_path / "again.dat", idx, x, y, z, w)[0]
    assert out.read_bytes() == src.read_bytes()
    # and the reader gives back what pandas reads from the shipped file, in canonical (voxel index) order -
    # dask's group-by left the shipped rows in an arbitrary one
    want = pd.read_csv(src, sep=";").sort_values("idx_sel_plas_points", kind="stable", ignore_index=True)
    pd.testing.assert_frame_equal(stage1_io.read(out), want)


def test_writer_equals_pandas_to_csv(tmp_path):
    rng = np.random.default_rng(1)
    n = 200_001  # more than one formatting block per thread, ragged tail
    idx = np.sort(rng.choice(10**9, n, replace=False)).astype(np.int64)
    xyz = np.round(rng.uniform(-4000, 4000, (n, 3)), 1)
    w = rng.uniform(0, 1, n) * 10.0 ** rng.integers(-14, -3, n)
    w[:5] = [0.0, 1e-300, 1.0, 1e16, 123.0]
    frame = pd.DataFrame({"idx_sel_plas_points": idx, "plasma_x": xyz[:, 0], "plasma_y": xyz[:, 1], "plasma_z": xyz[:, 2],
                          "total_intensity_fraction": w})
    want = tmp_path / "pandas.dat"
    frame.to_csv(want, sep=";", header=True, index=False)
    for threads in (1, 0):
        got = stage1_io.write_csv(tmp_path / f"native{threads}.dat", idx, xyz[:, 0], xyz[:, 1], xyz[:, 2], w, n_threads=threads)[0]
        assert got.read_bytes() == want.read_bytes()
    # dask-style partitions: contiguous parts, each with a header; together the same rows
    parts = stage1_io.write_csv(tmp_path / "part-*.dat", idx, xyz[:, 0], xyz[:, 1], xyz[:, 2], w, partitions=3)
    assert [p.name for p in parts] == ["part-0.dat", "part-1.dat", "part-2.dat"]
    pd.testing.assert_frame_equal(stage1_io.read(tmp_path / "part-*.dat"), frame)
    with pytest.raises(ValueError):
        stage1_io.write_csv(tmp_path / "nostar.dat", idx, xyz[:, 0], xyz[:, 1], xyz[:, 2], w, partitions=2)
    empty = stage1_io.write_csv(tmp_path / "empty.dat", idx[:0], w[:0], w[:0], w[:0], w[:0])[0]
    assert empty.read_text() == ";".join(stage1_io.COLUMNS) + "\n"
    # a path that cannot be opened for writing (a directory) is reported, not ignored
    assert _lib.load().com_write_stage1_csv(str(tmp_path).encode(), None, None, None, None, None, 0, 1, 1) == _lib.COM_E_IO


def test_binary_side_format_round_trips_to_the_text_frame(tmp_path):
    with contextlib.redirect_stdout(io.StringIO()):
        mesh = CuboidMesh(10)
        other = CuboidMesh(15)
    ref = pd.read_csv(_shipped("N"), sep=";").sort_values("idx_sel_plas_points", kind="stable", ignore_index=True)
    path = stage1_io.write_binary(tmp_path / "N.bin", mesh.lattice, ref["idx_sel_plas_points"], ref["total_intensity_fraction"])
    assert path.stat().st_size == stage1_io.HEADER_BYTES + 16 * len(ref)  # vs ~55 bytes per row of text
    assert stage1_io.is_binary(path) and not stage1_io.is_binary(_shipped("N"))
    back = stage1_io.read(path, mesh=mesh)
    pd.testing.assert_frame_equal(back, ref)  # coordinates regenerated from the index: the same 1-decimal values
    only = stage1_io.read_binary(path, coordinates=False)
    assert list(only.columns) == ["idx_sel_plas_points", "total_intensity_fraction"]
    with pytest.raises(ValueError):
        stage1_io.read(path, mesh=other)
    with pytest.raises(ValueError):
        stage1_io.write_binary(tmp_path / "x.bin", mesh.lattice.shard(2, 0), ref["idx_sel_plas_points"], ref["total_intensity_fraction"])
    path.write_bytes(path.read_bytes()[:-8])
    with pytest.raises(ValueError):
        stage1_io.read(path, mesh=mesh)
