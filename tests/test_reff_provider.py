"""Reff provider: the oracle's NumPy resampling (oracle/reff.py) against the shipped
grids on CPU, the CUDA kernels (com_reff_resample / com_reff_quadratic) against the oracle on the GPU - bit-exact for
the resampling."""
import contextlib
import io

import numpy as np
import pytest

from co_monitor_b200.geometry import reff as R
from co_monitor_b200.geometry.mesh_calculation import CuboidMesh
from oracle import reff as oreff


def _mesh(spacing):
    with contextlib.redirect_stdout(io.StringIO()):
        return CuboidMesh(spacing)


def _mm(reff):
    mm = np.rint(reff * 1000.0)
    return np.where(np.isnan(mm), 65535, np.clip(mm, 0, 65534)).astype(np.uint16)


def _shape(lat):
    return lat.nx, lat.ny, lat.nz


def test_oracle_resampling_is_consistent(reff10):
    lat = _mesh(10).lattice
    whole = oreff.on_lattice(_shape(lat))
    np.testing.assert_array_equal(whole, reff10)  # the 10 mm table the Stage-2 fixtures were generated with
    rng = np.random.default_rng(0)
    flat = np.concatenate((rng.integers(0, lat.n_points, 20000), [0, lat.n_points - 1]))
    np.testing.assert_array_equal(oreff.on_lattice(_shape(lat), flat), whole[flat])
    # resampling a shipped grid onto its own lattice returns it (the lattices share their end points)
    for spacing in (15, 30):
        l2 = _mesh(spacing).lattice
        vol = oreff.shipped_volume(spacing)
        np.testing.assert_array_equal(oreff.on_lattice(_shape(l2), source_spacing=spacing), np.round(vol.ravel(), 3))
    # the product's host view of a shipped grid is the oracle's
    np.testing.assert_array_equal(R.source_volume(15)[0], oreff.shipped_volume(15))


def test_provider_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from co_monitor_b200._lib import ComonitorError

    with pytest.raises(ComonitorError):
        R.resample_reff_device(_mesh(10).lattice)


@pytest.mark.gpu
def test_resample_kernel_is_bit_identical_to_the_oracle(reff10):
    import torch

    lat = _mesh(10).lattice
    mm, reff = R.resample_reff_device(lat, want_reff=True)
    np.testing.assert_array_equal(reff.cpu().numpy(), oreff.on_lattice(_shape(lat)))  # NaN positions included
    np.testing.assert_array_equal(R.interpolate_reff(lat), reff10)  # the host-array form of the same kernel
    np.testing.assert_array_equal(mm.cpu().numpy(), _mm(reff10))
    # the pipeline's own conversion of the float table gives the same bins
    from co_monitor_b200.pipeline import reff_to_mm

    assert torch.equal(reff_to_mm(torch.from_numpy(reff10.copy()).cuda()).view(torch.int16), mm.view(torch.int16))
    # sub-range and block-cyclic shard
    lo, hi = 123457, 201111
    mm2, _ = R.resample_reff_device(lat, begin=lo, end=hi)
    np.testing.assert_array_equal(mm2.cpu().numpy(), _mm(reff10[lo:hi]))
    for index in range(3):
        sh = lat.shard(3, index, cols_per_block=8)
        mm3, _ = R.resample_reff_device(sh)
        glob = sh.local_to_global(np.arange(sh.n_points))
        np.testing.assert_array_equal(mm3.cpu().numpy(), _mm(reff10[glob]))
    empty, _ = R.resample_reff_device(lat, begin=5, end=5)
    assert empty.numel() == 0


@pytest.mark.gpu
def test_resample_kernel_on_fine_lattice_slabs():
    """1 mm and 0.65 mm lattices (2.7e8 / 9.8e8 voxels): slabs at three positions against the pointwise NumPy form."""
    for spacing in (1, 0.65):
        lat = _mesh(spacing).lattice
        src = R.source_volume(15)
        for frac in (0.11, 0.5, 0.83):
            lo = int(lat.n_points * frac)
            hi = lo + 150_000
            mm, reff = R.resample_reff_device(lat, begin=lo, end=hi, want_reff=True, source=src)
            want = oreff.on_lattice(_shape(lat), np.arange(lo, hi))
            np.testing.assert_array_equal(reff.cpu().numpy(), want)
            np.testing.assert_array_equal(mm.cpu().numpy(), _mm(want))
        sh = lat.shard(8, 5)
        lo = sh.n_points // 2
        mm, _ = R.resample_reff_device(sh, begin=lo, end=lo + 100_000, source=src)
        np.testing.assert_array_equal(mm.cpu().numpy(), _mm(oreff.on_lattice(_shape(sh), sh.local_to_global(np.arange(lo, lo + 100_000)))))


@pytest.mark.gpu
def test_quadratic_model_kernel_matches_numpy():
    mesh = _mesh(10)
    model = R.QuadraticReff(15)
    want = np.round(model(mesh.outer_cube_mesh), 3)
    mm, reff = model.on_device(mesh.lattice, want_reff=True)
    got = reff.cpu().numpy()
    # the kernel takes the voxel position from the lattice closed form, NumPy from a BLAS product: identical up to
    # the last bit of the position, so only a value sitting on a rounding boundary may move by one millimetre
    both = np.isfinite(got) & np.isfinite(want)
    assert (np.isfinite(got) != np.isfinite(want)).sum() <= 2
    diff = np.abs(got[both] - want[both])
    assert diff.max() <= 1.0000001e-3 and (diff > 0).sum() <= 2
    assert both.sum() > 100_000
    np.testing.assert_array_equal(mm.cpu().numpy()[both], _mm(got[both]))
    sh = mesh.lattice.shard(2, 1)
    mm2, _ = model.on_device(sh)
    glob = sh.local_to_global(np.arange(sh.n_points))
    np.testing.assert_array_equal(mm2.cpu().numpy(), mm.cpu().numpy()[glob])
