"""Plasma voxel lattice (host side, fp64).

Drop-in for ``co_monitor.geometry.mesh_calculation.CuboidMesh``
(reference: co_monitor/geometry/mesh_calculation.py:9-159).  The reference
materialises the whole ``(P, 3)`` array (four temporaries deep), which cannot be
done for the 1e8-1e9 voxel grids of the large configurations; here the mesh is a
*lattice descriptor* and the dense array is built only when somebody asks for
``outer_cube_mesh``.  The closed form is

    flat index  i = (iy * nx + ix) * nz + iz            (meshgrid 'xy' + C ravel)
    point       p = ((x[ix], y[iy], z[iz]) - c) . R + c + (50, 25, 0)

with ``x = linspace(cx - Lx/2, cx + Lx/2, int(Lx // s))`` etc. and ``R`` the
1 rad rotation about z (mesh_calculation.py:128-155).  ``points_at`` evaluates it
with the same NumPy operations as the reference, so the coordinates (and their
``.round(1)`` in the Stage-1 file) are bit-identical.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .json_reader import read_json_file
from .rotation_matrix import rotation_matrix

#: translation applied after the rotation (mesh_calculation.py:153-155)
MESH_SHIFT = np.array([50.0, 25.0, 0.0])


@dataclass(frozen=True)
class Lattice:
    """Implicit description of the rotated voxel lattice, consumed by the CUDA path."""

    nx: int
    ny: int
    nz: int
    start: np.ndarray  # (3,) first linspace sample per axis
    stop: np.ndarray  # (3,) last linspace sample per axis (set exactly, as numpy does)
    step: np.ndarray  # (3,) (stop - start) / (n - 1)
    centre: np.ndarray  # (3,) plasma central point c
    rotation: np.ndarray  # (3, 3) R, row-vector convention
    shift: np.ndarray  # (3,)
    # block-cyclic shard of the z-columns (multi-GPU): every flat index is then local to the shard
    shard_cols_per_block: int = 0
    shard_count: int = 1
    shard_index: int = 0

    @property
    def n_points_global(self) -> int:
        return self.nx * self.ny * self.nz

    @property
    def n_points(self) -> int:
        """Voxels of this lattice, or of this shard of it."""
        cols = self.nx * self.ny
        if self.shard_count <= 1 or self.shard_cols_per_block <= 0:
            return cols * self.nz
        cpb = self.shard_cols_per_block
        blocks = -(-cols // cpb)
        mine = len(range(self.shard_index, blocks, self.shard_count))
        local_cols = mine * cpb
        if mine and (blocks - 1) % self.shard_count == self.shard_index:
            local_cols -= blocks * cpb - cols
        return local_cols * self.nz

    def shard(self, count: int, index: int, cols_per_block: int = 8) -> "Lattice":
        """Shard *index* of *count*: z-columns dealt round-robin in blocks of *cols_per_block*."""
        from dataclasses import replace

        if not 0 <= index < count:
            raise ValueError(f"shard {index} outside {count}")
        if count == 1:
            return replace(self, shard_cols_per_block=0, shard_count=1, shard_index=0)
        return replace(self, shard_cols_per_block=int(cols_per_block), shard_count=int(count), shard_index=int(index))

    def local_to_global(self, flat):
        """Global flat index of the shard-local flat index (identity when not sharded)."""
        flat = np.asarray(flat, dtype=np.int64)
        if self.shard_count <= 1 or self.shard_cols_per_block <= 0:
            return flat
        cpb = self.shard_cols_per_block
        cl, iz = flat // self.nz, flat % self.nz
        cg = ((cl // cpb) * self.shard_count + self.shard_index) * cpb + cl % cpb
        return cg * self.nz + iz

    def unravel(self, flat):
        flat = self.local_to_global(flat)
        iz = flat % self.nz
        t = flat // self.nz
        return t % self.nx, t // self.nx, iz  # ix, iy, iz

    def axis_values(self, axis: int) -> np.ndarray:
        n = (self.nx, self.ny, self.nz)[axis]
        return np.linspace(self.start[axis], self.stop[axis], n, endpoint=True)

    def bounding_corners(self) -> np.ndarray:
        """The 8 corner points of the (rotated) lattice, fp64, shape (8, 3)."""
        xs = (self.start[0], self.stop[0])
        ys = (self.start[1], self.stop[1])
        zs = (self.start[2], self.stop[2])
        raw = np.array([[x, y, z] for x in xs for y in ys for z in zs])
        return (raw - self.centre).dot(self.rotation) + self.centre + self.shift


class CuboidMesh:
    """Mesh of the cuboid enclosing the plasma volume seen by the C/O monitor."""

    def __init__(
        self,
        distance_between_points=10,
        cuboid_dimensions=(1350, 800, 250),
        plot: bool = False,
    ):
        assert 2 * distance_between_points <= min(
            cuboid_dimensions
        ), "Density of points is too low!"
        self.distance_between_points = distance_between_points
        self.x, self.y, self.z = cuboid_dimensions
        self.loaded_file = read_json_file()
        self.central_plasma_coord = self.get_coordinate()
        self.lattice = self._build_lattice()
        self._dense = None
        print(
            f"Number of points in the considered volume: {self.lattice.n_points} points."
        )
        if plot:
            self.visualization()

    # -- reference-named helpers -------------------------------------------------
    def get_coordinate(self) -> np.ndarray:
        return np.array(self.loaded_file["plasma"]["central point"])

    def calculate_cubes_edge_length(self, coord_range) -> float:
        return abs(coord_range[0] - coord_range[1])

    def calculate_interval(self, coord_range) -> int:
        return int(
            self.calculate_cubes_edge_length(coord_range)
            // self.distance_between_points
        )

    def set_cube_ranges(self) -> np.ndarray:
        half = np.array([self.x, self.y, self.z]) / 2
        c = self.central_plasma_coord
        return np.column_stack((c - half, c + half))

    # -- lattice -----------------------------------------------------------------
    def _build_lattice(self) -> Lattice:
        ranges = self.set_cube_ranges()
        counts = [self.calculate_interval(r) for r in ranges]
        start, stop = ranges[:, 0].copy(), ranges[:, 1].copy()
        step = np.array(
            [(b - a) / (n - 1) if n > 1 else 0.0 for a, b, n in zip(start, stop, counts)]
        )
        # rotation about z by 1 rad: axis = x_hat x y_hat (mesh_calculation.py:145-150)
        spin_axis = np.cross(np.array([1, 0, 0]) / 1.0, np.array([0, 1, 0]))
        return Lattice(
            nx=counts[0],
            ny=counts[1],
            nz=counts[2],
            start=start,
            stop=stop,
            step=step,
            centre=self.central_plasma_coord.astype(float),
            rotation=rotation_matrix(1, spin_axis),
            shift=MESH_SHIFT.copy(),
        )

    def points_at(self, flat_indices, threads: int | None = None) -> np.ndarray:
        """fp64 coordinates of the voxels with the given flat indices, shape (n, 3).

        Long index lists (the 1e7-1e8 visible voxels of a fine grid) are evaluated in blocks on a thread pool; every
        row goes through the same NumPy operations, so the values do not depend on the blocking
        (tests/test_host_setup.py)."""
        flat_indices = np.asarray(flat_indices, dtype=np.int64)
        block = 1 << 18
        if flat_indices.ndim == 1 and len(flat_indices) > 2 * block and threads != 1:
            import os
            from concurrent.futures import ThreadPoolExecutor

            parts = [flat_indices[i:i + block] for i in range(0, len(flat_indices), block)]
            with ThreadPoolExecutor(threads or min(16, os.cpu_count() or 1)) as pool:
                return np.concatenate(list(pool.map(lambda p: self.points_at(p, threads=1), parts)))
        lat = self.lattice
        ix, iy, iz = lat.unravel(flat_indices)
        raw = np.vstack(
            (lat.axis_values(0)[ix], lat.axis_values(1)[iy], lat.axis_values(2)[iz])
        ).T
        raw = raw - self.central_plasma_coord
        return raw.dot(lat.rotation) + self.central_plasma_coord + [50, 25, 0]

    def mesh_outer_cube(self) -> np.ndarray:
        """Dense ``(P, 3)`` mesh in the reference's point order."""
        return self.points_at(np.arange(self.lattice.n_points, dtype=np.int64))

    @property
    def outer_cube_mesh(self) -> np.ndarray:
        if self._dense is None:
            self._dense = self.mesh_outer_cube()
        return self._dense

    def visualization(self):
        raise NotImplementedError(
            "3-D rendering (pyvista) is outside the accelerated path of co_monitor_b200"
        )
