"""Stage 1 drop-in: ``Simulation`` with the reference's call signature, attributes and output file.

Replaces ``co_monitor.geometry.simulation.Simulation`` (reference:
co_monitor/geometry/simulation.py:30-630).  The host builds the same setup objects as the
reference (mesh, collimator, crystal, port, detector); the P x C ray work - reflection
(:273-325), aperture tests (:327-432), distance (:483-504), AOI (:506-536), weight and
per-voxel sum (:538-594) - runs in the sm_100a kernels behind ``com_visibility_weights``.
Nothing is evaluated on the CPU and nothing of size P x C is ever materialised unless a
caller asks for one of the reference's per-ray attributes.

Differences a caller can see (all deliberate):
  * ``ddf`` is a pandas DataFrame with a ``compute()`` shim (the reference returns a dask
    DataFrame); rows are ordered by ``idx_sel_plas_points`` (dask's group-by order is arbitrary).
  * ``cuboid_coordinates`` / ``crystal_coordinates`` are NumPy arrays, not dask arrays.
  * ``selected_intersections``, ``selected_indices``, ``distances``, ``angles_of_incident`` are
    computed on first access (one extra kernel pass that also returns the passing-ray list).
  * ``plot=True`` is accepted and ignored with a notice (pyvista rendering is out of scope).
  * ``save_to_file`` writes one partition, named as dask would name partition 0
    (``...-slit_<S>0.dat`` == the shipped ``slit_100`` files for S = 10), into
    ``<cwd>/input_files/geometry/observed_plasma_volume/<element>/`` where Stage 2 looks for it
    (the reference writes to a stale package-relative ``_Input_files`` path, simulation.py:619-625).
    The bytes are the ones pandas/dask produce, written by the native writer (stage1_io.py);
    ``save_to_file(partitions=k)`` writes k dask-style partitions, ``save_to_file(binary=True)``
    the 16-byte-per-row side format for 1e8-row results.
  * Fine grids (1 mm: 2.7e8 voxels, 0.65 mm: 9.8e8) run as several launches of ``stage1_chunk``
    voxels; nothing per-voxel survives a launch except the visible voxels' indices and weights.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path

import numpy as np
import pandas as pd

from .collimator import Collimator
from .decorators import timer
from .detector import Detector
from .dispersive_element import DispersiveElement
from .engine import DEFAULT_FILTER_EPS_MM, ChannelScene, crystal_normals
from .mesh_calculation import CuboidMesh
from .port import Port

DDF_COLUMNS = ["idx_sel_plas_points", "plasma_x", "plasma_y", "plasma_z", "total_intensity_fraction"]


class ComputedFrame(pd.DataFrame):
    """pandas DataFrame that also answers dask's ``.compute()`` (returns a plain DataFrame)."""

    @property
    def _constructor(self):
        return ComputedFrame

    def compute(self, **_):
        return pd.DataFrame(self)


@timer
def _ecrh_shields(element):
    from .radiation_shield import ecrh_shields

    return ecrh_shields(element)


class Simulation:
    def __init__(
        self,
        element,
        slits_number=10,
        distance_between_points=10,
        crystal_height_step=20,
        crystal_length_step=80,
        savetxt=False,
        plot=True,
        *,
        device=None,
        filter_eps_mm=DEFAULT_FILTER_EPS_MM,
        calibrate_apertures=True,
        output_dir=None,
        verbose=True,
        stage1_chunk=1 << 26,
        closing_side="top closing side",
        ecrh_shields=False,
        segment_check=False,
    ):
        """Keyword-only extras (none changes the reference's results at its default): *closing_side* selects the
        collimator position (the reference hard-codes the top one); *ecrh_shields* adds the two ECRH stray-radiation
        shields of the channel's chamber as circular stops (the reference only draws them, radiation_shield.py:52-101);
        *segment_check* accepts an aperture hit only on the physical part of the ray's line (the reference intersects
        infinite lines, simulation.py:143-173).  ``stats`` reports how many rays the last two stopped."""
        self._say = print if verbose else (lambda *a, **k: None)
        self._say(f"\nInitializing element {element}.")
        self.element = element
        self.slits_number = slits_number
        self.distance_between_points = distance_between_points
        self.crystal_height_step = crystal_height_step
        self.crystal_length_step = crystal_length_step
        self.plot = plot
        self._device = device
        self._output_dir = output_dir
        self.closing_side = closing_side  # the reference hard-codes the top one (simulation.py:91)
        self._rays = None
        self._selected = None

        self._init_cuboid_coordinates()
        self._init_collimator()
        self._init_dispersive_element()
        self._init_port()
        self._init_detector()

        self.scene = ChannelScene(
            self.disp_elem_coord, self.crystal_normals, self.slit_coord_crys_side, self.slit_coord_plasma_side,
            self.collim_vector_front_back, self.port_vertices_coordinates, self.port_orientation_vector,
            self.detector_vertices_coordinates, self.detector_orientation_vector,
            origin=self._crystal_central_point, area_pt=self.crystal_point_area, refl_max=self.max_reflectivity,
            aoi0=float(self.AOI), voxel_corners=self.mesh.lattice.bounding_corners(),
            filter_eps_mm=filter_eps_mm, calibrate_apertures=calibrate_apertures, segment_check=bool(segment_check),
            shields=_ecrh_shields(element) if ecrh_shields else None,
        )
        self._say("\n--- Calculating photon transmission ---")
        self._visible_idx, self._visible_w, self.stats = self._run_stage1(int(stage1_chunk))
        self._say("--- Photon transmission calculation finished ---")
        if self.stats["passing_rays"] == 0:
            # simulation.py:451-456
            sys.exit("Not enough points to perform calculations!")
        self.ddf = self.calculate_radiation_fraction()

        if savetxt:
            self.save_to_file()
        if plot:
            warnings.warn("co_monitor_b200: plot=True ignored (3-D rendering is outside the accelerated path)",
                          stacklevel=3)

    # ---- setup (host, NumPy fp64; same objects as simulation.py:77-141) ------------------------------
    def _init_cuboid_coordinates(self):
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()) as buf:
            self.mesh = CuboidMesh(self.distance_between_points)
        self._say(buf.getvalue(), end="")

    @property
    def cuboid_coordinates(self) -> np.ndarray:
        """(P, 3) fp64 voxel coordinates; built on demand (the kernels use the implicit lattice)."""
        return self.mesh.outer_cube_mesh

    def _init_collimator(self):
        self.collim = Collimator(self.element, self.closing_side, self.slits_number)
        self.collim_vector_front_back = self.collim.vector_front_back
        (
            self.collimator_spatial,
            self.slit_coord_crys_side,
            self.slit_coord_plasma_side,
        ) = self.collim.read_colim_coord()

    def _init_dispersive_element(self):
        de = DispersiveElement(self.element, self.crystal_height_step, self.crystal_length_step)
        self.disp_elem_coord = de.crystal_points
        self.AOI = de.AOI
        self.max_reflectivity = de.max_reflectivity
        self.radius_central_point = de.radius_central_point
        self.B = de.B
        self.C = de.C
        self._crystal_central_point = de.crystal_central_point
        self.crystal_point_area = self.calculate_crystal_point_area()
        self.crystal_coordinates = self.disp_elem_coord
        self.crystal_normals = crystal_normals(self.disp_elem_coord, self.radius_central_point, self.B, self.C)

    def _init_port(self):
        port = Port()
        self.port_vertices_coordinates = port.vertices_coordinates
        self.port_orientation_vector = port.orientation_vector

    def _init_detector(self):
        det = Detector(self.element)
        self.detector_vertices_coordinates = det.vertices
        self.detector_orientation_vector = det.orientation_vector

    def calculate_crystal_point_area(self) -> float:
        area = round(20 * 80 / self.crystal_height_step / self.crystal_length_step, 2)
        self._say(f"\nCrystal area per investigated point is: {area} mm^2.")
        return area

    # ---- results ---------------------------------------------------------------------------------------
    def _run_stage1(self, chunk: int):
        """K1a/K1b + compaction over the lattice in launches of at most *chunk* voxels: the visible voxels'
        ascending flat indices and weights (host arrays) and the summed statistics."""
        lattice = self.mesh.lattice
        total = lattice.n_points
        idx_parts, w_parts, stats = [], [], None
        for lo in range(0, total, max(1, chunk)):
            res = self.scene.visibility(lattice=lattice, begin=lo, end=min(total, lo + chunk), device=self._device)
            idx_d, w_d = self.scene.compact(res)
            idx_parts.append(idx_d.cpu().numpy())
            w_parts.append(w_d.cpu().numpy())
            if stats is None:
                stats = dict(res.stats)
            else:
                for key, value in res.stats.items():
                    stats[key] = max(stats[key], value) if key in ("candidate_capacity", "overflow") else stats[key] + value
            del res, idx_d, w_d
        if len(idx_parts) == 1:
            return idx_parts[0], w_parts[0], stats
        return np.concatenate(idx_parts), np.concatenate(w_parts), stats

    def calculate_radiation_fraction(self) -> ComputedFrame:
        """Visible voxels with their summed weight (simulation.py:538-594), ordered by voxel index."""
        idx, w = self._visible_idx, self._visible_w
        pts = self.mesh.points_at(idx)
        return ComputedFrame(
            {
                "idx_sel_plas_points": idx,
                "plasma_x": pts[:, 0].round(1),
                "plasma_y": pts[:, 1].round(1),
                "plasma_z": pts[:, 2].round(1),
                "total_intensity_fraction": w,
            },
            columns=DDF_COLUMNS,
        )

    def _ray_list(self) -> np.ndarray:
        if self._rays is None:
            res = self.scene.visibility(lattice=self.mesh.lattice, device=self._device, want_rays=True)
            rays = res.rays
            self._rays = rays[np.argsort(rays["ray_index"], kind="stable")]
        return self._rays

    @property
    def selected_intersections(self) -> np.ndarray:
        """(P, C) bool ray mask (simulation.py:408-432).  Dense: P*C bytes of host memory."""
        if self._selected is None:
            n_rays = self.mesh.lattice.n_points * self.scene.n_crystal
            mask = np.zeros(n_rays, dtype=bool)
            mask[self._ray_list()["ray_index"]] = True
            self._selected = mask.reshape(self.mesh.lattice.n_points, self.scene.n_crystal)
        return self._selected

    @property
    def selected_indices(self) -> pd.DataFrame:
        """Flat ray indices of the passing rays (simulation.py:434-461)."""
        rays = self._ray_list()
        return pd.DataFrame({"index": np.arange(len(rays)), "idx_sel_plas_points": rays["ray_index"] // self.scene.n_crystal})

    @property
    def plas_points_indices(self) -> pd.DataFrame:
        return self.selected_indices

    @property
    def distances(self) -> pd.DataFrame:
        """|plasma - crystal| of every passing ray in ray order (simulation.py:483-504)."""
        rays = self._ray_list()
        return pd.DataFrame({"index": np.arange(len(rays)), 0: rays["distance"]})

    @property
    def angles_of_incident(self) -> pd.DataFrame:
        """Angle of incidence of every passing ray in ray order (simulation.py:506-536)."""
        rays = self._ray_list()
        return pd.DataFrame({"index": np.arange(len(rays)), 0: rays["aoi"]})

    def output_path(self) -> Path:
        root = Path(self._output_dir) if self._output_dir is not None else (
            Path.cwd() / "input_files" / "geometry" / "observed_plasma_volume" / f"{self.element}")
        name = (f"{self.element}_plasma_coordinates-{self.distance_between_points}_mm_spacing-"
                f"height_{self.crystal_height_step}-length_{self.crystal_length_step}-slit_{self.slits_number}0.dat")
        return root / name

    def save_to_file(self, partitions: int = 1, binary: bool = False):
        """``;``-separated CSV with header, no index (simulation.py:616-630): byte for byte what
        ``ddf.to_csv(sep=";", header=True, index=False)`` writes.  *partitions* > 1: dask-style numbered parts
        (``...slit_<S>0.dat``, ``...slit_<S>1.dat``, ...); *binary*: the index + weight side format (``.bin``)."""
        from . import stage1_io

        path = self.output_path()
        if binary:
            path = stage1_io.write_binary(path.with_suffix(".bin"), self.mesh.lattice, self.ddf["idx_sel_plas_points"],
                                          self.ddf["total_intensity_fraction"])
        else:
            name = str(path) if partitions <= 1 else str(path)[: -len("0.dat")] + "*.dat"
            d = self.ddf
            files = stage1_io.write_csv(name, d["idx_sel_plas_points"], d["plasma_x"], d["plasma_y"], d["plasma_z"],
                                        d["total_intensity_fraction"], partitions=partitions)
            path = files[0]
        self._say("\nFile successfully saved!")
        return path
