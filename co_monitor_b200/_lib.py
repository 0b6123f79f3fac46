"""ctypes binding of ``libcomonitor_sm100.so`` (the C ABI in ``include/comonitor_sm100.h``).

The library is the product: there is no Python/NumPy fallback.  If it has not been built
(``python -m co_monitor_b200.csrc.build`` / ``__graft_entry__.build()``) loading raises
``ComonitorLibraryError``; compute entry points raise ``ComonitorError`` on any non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libcomonitor_sm100.so")

COM_OK, COM_E_BADARG, COM_E_CUDA, COM_E_NOMEM, COM_E_OVERFLOW, COM_E_GEOMETRY, COM_E_EMPTY, COM_E_IO = 0, -1, -2, -3, -4, -5, -6, -7
COM_MAX_EDGES = 32


class ComonitorLibraryError(ImportError):
    pass


class ComonitorError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str):
        self.code = code
        super().__init__(f"{where}: {detail} (status {code})")


c_double3 = C.c_double * 3


class ComAperture(C.Structure):
    _fields_ = [
        ("p1", c_double3),
        ("normal", c_double3),
        ("e1", c_double3),
        ("e2", c_double3),
        ("n_edges", C.c_int32),
        ("reserved", C.c_int32),
        ("edge", (C.c_double * 3) * COM_MAX_EDGES),
        ("vertex", (C.c_double * 2) * COM_MAX_EDGES),
    ]

    def edges(self) -> np.ndarray:
        return np.array([list(self.edge[k]) for k in range(self.n_edges)])

    def vertices(self) -> np.ndarray:
        return np.array([list(self.vertex[k]) for k in range(self.n_edges)])


COM_MAX_SHIELDS = 4
COM_PHYS_SEGMENT_CHECK = 1


class ComShield(C.Structure):
    _fields_ = [
        ("centre", c_double3),
        ("axis", c_double3),
        ("u", c_double3),
        ("v", c_double3),
        ("radius", C.c_double),
        ("n_sides", C.c_int32),
        ("reserved", C.c_int32),
    ]


class ComSceneDesc(C.Structure):
    _fields_ = [
        ("n_crystal", C.c_int32),
        ("n_slits", C.c_int32),
        ("crystal_xyz", C.POINTER(C.c_double)),
        ("crystal_normal", C.POINTER(C.c_double)),
        ("slit_crys_side", C.POINTER(C.c_double)),
        ("slit_plasma_side", C.POINTER(C.c_double)),
        ("slit_depth", c_double3),
        ("n_port_vertices", C.c_int32),
        ("n_detector_vertices", C.c_int32),
        ("port_vertices", C.POINTER(C.c_double)),
        ("port_depth", c_double3),
        ("detector_vertices", C.POINTER(C.c_double)),
        ("detector_depth", c_double3),
        ("origin", c_double3),
        ("area_pt", C.c_double),
        ("refl_max", C.c_double),
        ("aoi0", C.c_double),
        ("refl_sigma", C.c_double),
        ("filter_eps_mm", C.c_double),
        ("apertures_override", C.POINTER(ComAperture)),
        ("near_edge_mm", C.c_double),
        ("n_voxel_corners", C.c_int32),
        ("filter_flags", C.c_int32),
        ("voxel_corners", (C.c_double * 3) * 8),
        ("near_band_mm", C.c_double),
        ("physics_flags", C.c_int32),
        ("n_shields", C.c_int32),
        ("shields", C.POINTER(ComShield)),
    ]


class ComLattice(C.Structure):
    _fields_ = [
        ("nx", C.c_int64),
        ("ny", C.c_int64),
        ("nz", C.c_int64),
        ("start", c_double3),
        ("step", c_double3),
        ("stop", c_double3),
        ("centre", c_double3),
        ("rotation", C.c_double * 9),
        ("shift", c_double3),
        ("shard_cols_per_block", C.c_int64),
        ("shard_count", C.c_int64),
        ("shard_index", C.c_int64),
    ]


class ComStats(C.Structure):
    _fields_ = [
        ("tests", C.c_uint64),
        ("candidates", C.c_uint64),
        ("passing_rays", C.c_uint64),
        ("near_edge_rays", C.c_uint64),
        ("visible_voxels", C.c_uint64),
        ("candidate_capacity", C.c_uint64),
        ("overflow", C.c_uint64),
        ("certain_candidates", C.c_uint64),
        ("near_band_rays", C.c_uint64),
        ("shield_blocked_rays", C.c_uint64),
        ("segment_rejected_rays", C.c_uint64),
        ("near_rounding_rays", C.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {name: int(getattr(self, name)) for name, _ in self._fields_ }


#: number of 64-bit words of a device-side ComStats block
N_STATS = len(ComStats._fields_)

COM_MAX_TRANSITIONS = 4
COM_MM_BINS = 4096
COM_MAX_CHANNELS = 8


class ComPecDesc(C.Structure):
    _fields_ = [
        ("n_ne", C.c_int32),
        ("n_te", C.c_int32),
        ("n_grid", C.c_int32),
        ("use_fully_stripped", C.c_int32),
        ("interpolation", C.c_int32),
        ("reserved", C.c_int32),
        ("ne_nodes", C.POINTER(C.c_double)),
        ("te_nodes", C.POINTER(C.c_double)),
        ("table", C.POINTER(C.c_double)),
        ("ne_grid", C.POINTER(C.c_double)),
        ("te_grid", C.POINTER(C.c_double)),
    ]


class ComTablesDesc(C.Structure):
    _fields_ = [
        ("n_fa", C.c_int32),
        ("n_transitions", C.c_int32),
        ("fa_T", C.POINTER(C.c_double)),
        ("fa_ion", C.POINTER(C.c_double)),
        ("fa_last", C.POINTER(C.c_double)),
        ("pec", ComPecDesc * COM_MAX_TRANSITIONS),
    ]


RAY_RECORD_DTYPE = np.dtype(
    [("ray_index", np.int64), ("distance", np.float64), ("aoi", np.float64), ("weight", np.float64)]
)
STATS_DTYPE = np.dtype([(name, np.uint64) for name, _ in ComStats._fields_])

_lib = None


def _declare(lib):
    vp, i32, i64, u64p, dp = C.c_void_p, C.c_int32, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    lib.com_abi_version.restype = C.c_int
    lib.com_last_error.restype = C.c_char_p
    lib.com_error_string.restype = C.c_char_p
    lib.com_error_string.argtypes = [C.c_int]
    lib.com_aperture_from_slab.argtypes = [dp, i32, dp, C.POINTER(ComAperture)]
    lib.com_aperture_contains.argtypes = [C.POINTER(ComAperture), dp]
    lib.com_scene_create.argtypes = [C.POINTER(ComSceneDesc), C.POINTER(vp)]
    lib.com_scene_destroy.argtypes = [vp]
    lib.com_lattice_local_voxels.restype = C.c_int64
    lib.com_lattice_local_voxels.argtypes = [C.POINTER(ComLattice)]
    lib.com_scene_aperture_count.argtypes = [vp]
    lib.com_scene_get_aperture.argtypes = [vp, i32, C.POINTER(ComAperture)]
    lib.com_scene_set_detector_image.argtypes = [vp, vp, i32, i32, C.c_double, C.c_double, C.c_double]
    lib.com_scene_filter_tables_host.argtypes = [C.POINTER(ComSceneDesc), C.POINTER(i32), vp, vp, vp, vp, vp]
    lib.com_scene_interval_tables_host.argtypes = [C.POINTER(ComSceneDesc), vp, C.POINTER(i32), vp, vp, vp, C.POINTER(i32)]
    lib.com_scene_port_edges_host.argtypes = [C.POINTER(ComSceneDesc), vp, C.POINTER(i32)]
    lib.com_scene_detector_edges_host.argtypes = [C.POINTER(ComSceneDesc), vp, C.POINTER(i32)]
    lib.com_scene_filter_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), dp, dp]
    lib.com_visibility_workspace_bytes.restype = C.c_size_t
    lib.com_visibility_workspace_bytes.argtypes = [vp, i64, C.c_double]
    lib.com_visibility_weights.argtypes = [
        vp, C.POINTER(ComLattice), vp, i64, i64, vp, vp, vp, i64, vp, vp, vp, i32, vp, vp, C.c_size_t, vp,
    ]
    lib.com_voxels_pack.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.com_visibility_weights_packed.argtypes = [
        vp, vp, vp, i64, i64, vp, vp, vp, i64, vp, vp, vp, i32, vp, vp, C.c_size_t, vp,
    ]
    lib.com_visibility_weights_host.argtypes = [
        C.POINTER(ComSceneDesc), C.POINTER(ComLattice), vp, i64, i64, vp, vp, vp, i64, u64p, C.POINTER(ComStats),
    ]
    lib.com_visible_weights_host.argtypes = [
        C.POINTER(ComSceneDesc), C.POINTER(ComLattice), vp, i64, i64, vp, vp, i64, u64p, C.POINTER(ComStats),
    ]
    lib.com_visibility_weights_host_scene.argtypes = [vp, C.POINTER(ComLattice), vp, i64, i64, vp, vp, vp, i64, u64p,
                                                      C.POINTER(ComStats)]
    lib.com_visible_weights_host_scene.argtypes = [vp, C.POINTER(ComLattice), vp, i64, i64, vp, vp, i64, u64p,
                                                   C.POINTER(ComStats)]
    lib.com_emissivity_integrate_indexed_host_tables.argtypes = [vp, dp, i32, C.c_double, dp, i64, vp, dp, i64, dp, vp, vp, dp]
    lib.com_volume_create.argtypes = [C.POINTER(ComLattice), i64, C.c_double, C.POINTER(vp)]
    lib.com_volume_destroy.argtypes = [vp]
    lib.com_volume_voxels.restype = C.c_int64
    lib.com_volume_voxels.argtypes = [vp]
    lib.com_volume_chunk.restype = C.c_int64
    lib.com_volume_chunk.argtypes = [vp]
    lib.com_volume_reff_mm.restype = vp
    lib.com_volume_reff_mm.argtypes = [vp]
    lib.com_volume_set_reff_resampled.argtypes = [vp, vp, i32, i64, i64, i64, vp]
    lib.com_volume_set_reff_quadratic.argtypes = [vp, dp, C.c_double, vp]
    lib.com_volume_histograms.argtypes = [vp, i32, C.POINTER(vp), vp, vp, vp]
    lib.com_volume_histograms_host.argtypes = [vp, i32, C.POINTER(vp), vp, i64, i64, i64, vp, vp]
    lib.com_histograms_flux.argtypes = [i32, C.POINTER(vp), vp, i32, i32, i32, C.c_double, C.c_double, vp, vp, i32, vp, vp]
    lib.com_histograms_flux_host.argtypes = [i32, C.POINTER(vp), vp, i32, i32, C.c_double, vp, vp, i32, vp]
    lib.com_histograms_flux_peers.argtypes = [i32, C.POINTER(vp), vp, i32, i32, i32, C.c_double, C.c_double, i32, C.POINTER(vp),
                                              i32, vp, vp, vp, i32, vp, vp]
    lib.com_compact_workspace_bytes.restype = C.c_size_t
    lib.com_compact_workspace_bytes.argtypes = [i64]
    lib.com_compact_visible.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp, C.c_size_t, vp]
    lib.com_tables_create.argtypes = [C.POINTER(ComTablesDesc), C.POINTER(vp)]
    lib.com_tables_destroy.argtypes = [vp]
    lib.com_emissivity_workspace_bytes.restype = C.c_size_t
    lib.com_emissivity_workspace_bytes.argtypes = [i64]
    lib.com_emissivity_integrate.argtypes = [vp, vp, i32, i32, C.c_double, C.c_double, vp, vp, vp, vp, i64, vp,
                                             vp, vp, vp, vp, vp, C.c_size_t, vp]
    lib.com_reff_max.argtypes = [vp, vp, vp, i64, vp, vp, vp]
    lib.com_emissivity_integrate_host.argtypes = [C.POINTER(ComTablesDesc), dp, i32, C.c_double, dp, dp, i64, dp, vp, vp, dp]
    lib.com_emissivity_integrate_indexed_host.argtypes = [C.POINTER(ComTablesDesc), dp, i32, C.c_double, dp, i64, vp, dp, i64,
                                                          dp, vp, vp, dp]
    lib.com_weight_histogram.argtypes = [vp, i32, i32, C.c_double, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp,
                                         C.c_size_t, vp]
    lib.com_emissivity_sweep.argtypes = [vp, vp, i32, i32, vp, C.c_double, vp, vp]
    lib.com_weight_histogram_mm.argtypes = [vp, vp, i64, vp, vp, vp]
    lib.com_histogram_rows.argtypes = [vp, i32, i32, C.c_double, vp, vp, vp, vp, vp]
    lib.com_reff_resample.argtypes = [C.POINTER(ComLattice), i64, i64, vp, i64, i64, i64, vp, vp, vp]
    lib.com_reff_quadratic.argtypes = [C.POINTER(ComLattice), i64, i64, dp, C.c_double, vp, vp, vp]
    lib.com_scene_filter_certain_host.argtypes = [C.POINTER(ComSceneDesc), vp]
    lib.com_write_stage1_csv.argtypes = [C.c_char_p, vp, dp, dp, dp, dp, i64, i32, i32]
    lib.com_format_float_repr.argtypes = [C.c_double, C.c_char_p, i32]
    lib.com_measure_fp32_tflops.restype = C.c_double
    lib.com_kernel_timing_enable.argtypes = [C.c_int]
    lib.com_kernel_timing_read.argtypes = [dp, dp, C.POINTER(C.c_int64)]
    lib.com_kernel_timing_read3.argtypes = [dp, C.POINTER(C.c_int64)]
    return lib


#: every symbol include/comonitor_sm100.h declares (checked by tests/test_capi_symbols.py)
EXPORTED_SYMBOLS = [
    "com_abi_version", "com_last_error", "com_error_string",
    "com_aperture_from_slab", "com_aperture_contains",
    "com_scene_create", "com_scene_destroy", "com_scene_aperture_count", "com_scene_get_aperture",
    "com_scene_filter_info", "com_lattice_local_voxels", "com_scene_set_detector_image", "com_scene_filter_tables_host",
    "com_visibility_workspace_bytes", "com_visibility_weights", "com_visibility_weights_host",
    "com_voxels_pack", "com_visibility_weights_packed",
    "com_visible_weights_host", "com_host_arena_release",
    "com_compact_workspace_bytes", "com_compact_visible",
    "com_kernel_timing_enable", "com_kernel_timing_read", "com_kernel_timing_read3", "com_measure_fp32_tflops",
    "com_tables_create", "com_tables_destroy", "com_emissivity_workspace_bytes", "com_emissivity_integrate",
    "com_emissivity_integrate_host", "com_weight_histogram", "com_emissivity_sweep", "com_reff_max",
    "com_weight_histogram_mm", "com_histogram_rows", "com_reff_resample", "com_reff_quadratic",
    "com_write_stage1_csv", "com_format_float_repr", "com_scene_filter_certain_host",
    "com_emissivity_integrate_indexed_host", "com_emissivity_integrate_indexed_host_tables",
    "com_scene_interval_tables_host", "com_scene_port_edges_host", "com_scene_detector_edges_host",
    "com_visibility_weights_host_scene", "com_visible_weights_host_scene",
    "com_volume_create", "com_volume_destroy", "com_volume_voxels", "com_volume_chunk", "com_volume_reff_mm",
    "com_volume_set_reff_resampled", "com_volume_set_reff_quadratic", "com_volume_histograms",
    "com_volume_histograms_host", "com_histograms_flux", "com_histograms_flux_host", "com_histograms_flux_peers",
]


def load():
    """Load the shared library (once).  Raises ComonitorLibraryError when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ComonitorLibraryError(
                f"{LIB_PATH} is missing: build it with `python -m co_monitor_b200.csrc.build` "
                "(there is no CPU fallback for the accelerated path)"
            )
        try:
            _lib = _declare(C.CDLL(LIB_PATH))
        except OSError as exc:  # pragma: no cover
            raise ComonitorLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
    return _lib


def check(code: int, where: str) -> None:
    if code != COM_OK:
        lib = load()
        detail = lib.com_last_error().decode() or lib.com_error_string(code).decode()
        raise ComonitorError(code, where, detail)


def as_double_ptr(arr: np.ndarray):
    assert arr.dtype == np.float64 and arr.flags.c_contiguous
    return arr.ctypes.data_as(C.POINTER(C.c_double))


def aperture_from_slab(vertices: np.ndarray, offset: np.ndarray) -> ComAperture:
    """Host-only: plane + polygon of one aperture (no GPU needed)."""
    lib = load()
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    o = np.ascontiguousarray(np.asarray(offset, dtype=np.float64).reshape(3))
    ap = ComAperture()
    check(lib.com_aperture_from_slab(as_double_ptr(v), len(v), as_double_ptr(o), C.byref(ap)), "com_aperture_from_slab")
    return ap


def aperture_contains(ap: ComAperture, point) -> bool:
    p = np.ascontiguousarray(point, dtype=np.float64)
    return bool(load().com_aperture_contains(C.byref(ap), as_double_ptr(p)))
