"""Stage-1 engine: Python host side of the sm_100a visibility kernels.

Packs the reference-level geometry of one spectroscopic channel into a ``ComSceneDesc``,
owns the device-resident ``ComScene`` handle and drives ``com_visibility_weights`` on
torch-allocated device buffers (torch is used for memory, streams and NCCL only).
There is no CPU fallback: without the CUDA library or a CUDA device this raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from .. import _lib
from .apertures import build_apertures
from .mesh_calculation import Lattice

DEFAULT_FILTER_EPS_MM = 5e-3


def crystal_normals(points: np.ndarray, radius_central_point, B, C_vertex) -> np.ndarray:
    """Unit normals of the crystal cylinder at *points* (simulation.py:286-312).

    The cylinder axis runs through ``radius_central_point`` along ``B - C``; the normal is the
    unit vector from the foot of the perpendicular on that axis to the crystal point.
    """
    a = np.asarray(radius_central_point, dtype=float)
    ab = np.asarray(B, dtype=float) - np.asarray(C_vertex, dtype=float)
    b = a + ab
    ab = b - a
    foot = np.array([a + np.dot(p - a, ab) / np.dot(ab, ab) * ab for p in points])
    n = points - foot
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return n


def lattice_struct(lat: Lattice) -> _lib.ComLattice:
    out = _lib.ComLattice()
    out.nx, out.ny, out.nz = lat.nx, lat.ny, lat.nz
    for i in range(3):
        out.start[i], out.step[i], out.stop[i] = lat.start[i], lat.step[i], lat.stop[i]
        out.centre[i], out.shift[i] = lat.centre[i], lat.shift[i]
    for i, v in enumerate(np.asarray(lat.rotation, dtype=float).reshape(9)):
        out.rotation[i] = v
    out.shard_cols_per_block, out.shard_count, out.shard_index = lat.shard_cols_per_block, lat.shard_count, lat.shard_index
    return out


@dataclass
class VisibilityResult:
    """Device-resident output of one Stage-1 launch over the voxels [begin, end)."""

    begin: int
    end: int
    weight: "object"  # torch.float64 (n,)
    nrays: "object"  # torch.int32 (n,) (uint32 payload)
    stats: dict
    rays: np.ndarray | None = None  # structured host array of passing rays (optional)


class ChannelScene:
    """Device-resident beam line of one channel (crystal, collimator, port, detector)."""

    def __init__(self, crystal_points, normals, slit_crys, slit_plasma, vector_front_back, port_vertices,
                 port_ov, det_vertices, det_ov, origin, area_pt, refl_max, aoi0, refl_sigma=2.2,
                 voxel_corners=None, filter_eps_mm=DEFAULT_FILTER_EPS_MM, calibrate_apertures=True,
                 near_edge_mm=0.0, host_only=False, run_cull=True, near_band_mm=0.0, segment_check=False, shields=None,
                 intervals=True):
        """*near_edge_mm* / *near_band_mm*: widths of the two near-edge counters (defaults 1e-9 and 1e-4 mm).
        *run_cull* / *intervals*: measurement and test aids selecting the per-ray forms of the lattice filter.
        Extensions, off by default (not reference behaviour): *segment_check* accepts an aperture hit only on the
        physical part of the ray's line; *shields* is a list of ``ComShield``-like dicts (see ``ecrh_shields``) every
        ray must also pass."""
        self._lib = _lib.load()
        self._keep = {}

        def hold(name, arr):
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            self._keep[name] = arr
            return _lib.as_double_ptr(arr)

        d = _lib.ComSceneDesc()
        d.n_crystal = len(crystal_points)
        d.n_slits = len(slit_crys)
        d.crystal_xyz = hold("xyz", crystal_points)
        d.crystal_normal = hold("nrm", normals)
        d.slit_crys_side = hold("sc", slit_crys)
        d.slit_plasma_side = hold("sp", slit_plasma)
        d.n_port_vertices = len(port_vertices)
        d.n_detector_vertices = len(det_vertices)
        d.port_vertices = hold("pv", port_vertices)
        d.detector_vertices = hold("dv", det_vertices)
        pov = np.asarray(port_ov, dtype=float).reshape(3)
        dov = np.asarray(det_ov, dtype=float).reshape(3)
        vfb = np.asarray(vector_front_back, dtype=float).reshape(3)
        for i in range(3):
            d.slit_depth[i], d.port_depth[i], d.detector_depth[i] = vfb[i], pov[i], dov[i]
            d.origin[i] = float(origin[i])
        d.area_pt, d.refl_max, d.aoi0, d.refl_sigma = float(area_pt), float(refl_max), float(aoi0), float(refl_sigma)
        d.filter_eps_mm = float(filter_eps_mm)
        d.near_edge_mm = float(near_edge_mm)
        # COM_FILTER_NO_RUN_CULL (1): every ray's detector test evaluated on its own; COM_FILTER_NO_INTERVALS (2): the
        # per-ray lattice kernel with its run-level cull instead of the column-interval kernel
        d.filter_flags = (0 if run_cull else 1) | (0 if intervals else 2)
        d.near_band_mm = float(near_band_mm)
        d.physics_flags = _lib.COM_PHYS_SEGMENT_CHECK if segment_check else 0
        if shields:
            if len(shields) > _lib.COM_MAX_SHIELDS:
                raise ValueError(f"at most {_lib.COM_MAX_SHIELDS} shields per channel")
            arr_s = (_lib.ComShield * len(shields))()
            for q, sh in enumerate(shields):
                for i in range(3):
                    arr_s[q].centre[i], arr_s[q].axis[i] = float(sh["centre"][i]), float(sh["axis"][i])
                    arr_s[q].u[i], arr_s[q].v[i] = float(sh["u"][i]), float(sh["v"][i])
                arr_s[q].radius, arr_s[q].n_sides = float(sh["radius"]), int(sh.get("n_sides", 360))
            self._keep["shields"] = arr_s
            d.n_shields = len(shields)
            d.shields = C.cast(arr_s, C.POINTER(_lib.ComShield))
        self.aperture_shifts = None
        if calibrate_apertures:
            arr, self.aperture_shifts = build_apertures(
                self._keep["sc"].reshape(-1, 4, 3), self._keep["sp"].reshape(-1, 4, 3), vfb,
                self._keep["pv"], pov, self._keep["dv"], dov, calibrate=True)
            self._keep["apertures"] = arr
            d.apertures_override = C.cast(arr, C.POINTER(_lib.ComAperture))
        if voxel_corners is not None:
            vc = np.asarray(voxel_corners, dtype=float).reshape(-1, 3)
            d.n_voxel_corners = len(vc)
            for q in range(len(vc)):
                for i in range(3):
                    d.voxel_corners[q][i] = vc[q, i]
        self.desc = d
        self.n_crystal = d.n_crystal
        self.n_slits = d.n_slits
        self.handle = C.c_void_p()
        if not host_only:  # host_only: description + host-side analysis only (CPU tests of the filter tables)
            _lib.check(self._lib.com_scene_create(C.byref(d), C.byref(self.handle)), "com_scene_create")

    def filter_tables(self) -> dict:
        """The FP32 filter tables as NumPy arrays (host-only call, no GPU needed)."""
        cp = C.c_int32()
        _lib.check(self._lib.com_scene_filter_tables_host(C.byref(self.desc), C.byref(cp), None, None, None, None, None),
                   "com_scene_filter_tables_host")
        Cp = cp.value
        det = np.zeros((3, Cp, 4), np.float32)
        slit = np.zeros((Cp, 6, 4), np.float32)
        rects = np.zeros(8, np.float32)
        period = C.c_float()
        flags = (C.c_int32 * 3)()
        _lib.check(self._lib.com_scene_filter_tables_host(
            C.byref(self.desc), C.byref(cp), C.c_void_p(det.ctypes.data), C.c_void_p(slit.ctypes.data),
            C.c_void_p(rects.ctypes.data), C.byref(period), flags), "com_scene_filter_tables_host")
        cert = np.zeros(9, np.float32)
        _lib.check(self._lib.com_scene_filter_certain_host(C.byref(self.desc), C.c_void_p(cert.ctypes.data)),
                   "com_scene_filter_certain_host")
        r12 = np.zeros(12, np.float32)
        ivl = C.c_int32()
        port = np.zeros((Cp, 3, 4), np.float32)
        edges = np.zeros((2, 8, 4), np.float32)
        n_edges = np.zeros(2, np.int32)
        coff = C.c_int32()
        _lib.check(self._lib.com_scene_interval_tables_host(
            C.byref(self.desc), C.c_void_p(r12.ctypes.data), C.byref(ivl), C.c_void_p(port.ctypes.data),
            C.c_void_p(edges.ctypes.data), C.c_void_p(n_edges.ctypes.data), C.byref(coff)), "com_scene_interval_tables_host")
        pedges = np.zeros((16, 4), np.float32)
        n_pe = C.c_int32()
        _lib.check(self._lib.com_scene_port_edges_host(C.byref(self.desc), C.c_void_p(pedges.ctypes.data), C.byref(n_pe)),
                   "com_scene_port_edges_host")
        dedges = np.zeros((8, 4), np.float32)
        n_de = C.c_int32()
        _lib.check(self._lib.com_scene_detector_edges_host(C.byref(self.desc), C.c_void_p(dedges.ctypes.data), C.byref(n_de)),
                   "com_scene_detector_edges_host")
        return dict(det_forms=det, slit_forms=slit, rect_crys=rects[:4], rect_plasma=rects[4:], period=period.value,
                    port_edges=pedges[: n_pe.value], det_edges=dedges[: n_de.value],
                    det_inner=r12[:4], port_outer=r12[4:8], port_inner=r12[8:], interval_filter=bool(ivl.value), port_forms=port,
                    slit_edges=[edges[0, :n_edges[0]], edges[1, :n_edges[1]]], certain_off=bool(coff.value),
                    inner_crys=cert[:4], inner_plasma=cert[4:8], certain_mm=float(cert[8]),
                    slit_filter=bool(flags[0]), same_w=bool(flags[1]), disabled=bool(flags[2]),
                    origin=np.array(list(self.desc.origin)))

    # -- lifetime ----------------------------------------------------------------------
    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.com_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- introspection -------------------------------------------------------------------
    def filter_info(self) -> dict:
        sf, fd = C.c_int32(), C.c_int32()
        period, eps = C.c_double(), C.c_double()
        _lib.check(self._lib.com_scene_filter_info(self.handle, C.byref(sf), C.byref(fd), C.byref(period),
                                                   C.byref(eps)), "com_scene_filter_info")
        return dict(slit_filter=bool(sf.value), filter_disabled=bool(fd.value), slit_period_mm=period.value,
                    eps_mm=eps.value)

    def enable_detector_image(self, pixel_mm: float = 0.1, nx: int | None = None, ny: int | None = None, device=None):
        """Extension: accumulate every passing ray's weight into a detector image (see com_scene_set_detector_image).

        Returns the CUDA float64 tensor ``(ny, nx)``; it keeps accumulating over launches until zeroed by the caller.
        Default extent: the bounding box of the detector polygon in its own frame.
        """
        import torch

        det = self.aperture(2 * self.n_slits + 1)
        v = det.vertices()
        if nx is None:
            nx = int(np.ceil(v[:, 0].max() / pixel_mm))
        if ny is None:
            ny = int(np.ceil(v[:, 1].max() / pixel_mm))
        self.detector_image = torch.zeros((ny, nx), dtype=torch.float64, device=torch.device(device or "cuda"))
        _lib.check(self._lib.com_scene_set_detector_image(self.handle, C.c_void_p(self.detector_image.data_ptr()), nx, ny,
                                                          float(pixel_mm), 0.0, 0.0), "com_scene_set_detector_image")
        return self.detector_image

    def disable_detector_image(self):
        _lib.check(self._lib.com_scene_set_detector_image(self.handle, None, 0, 0, 0.0, 0.0, 0.0), "com_scene_set_detector_image")
        self.detector_image = None

    def aperture(self, index: int) -> _lib.ComAperture:
        ap = _lib.ComAperture()
        _lib.check(self._lib.com_scene_get_aperture(self.handle, index, C.byref(ap)), "com_scene_get_aperture")
        return ap

    # -- Stage 1 -------------------------------------------------------------------------
    def visibility(self, lattice: Lattice | None = None, voxels=None, begin: int = 0, end: int | None = None,
                   want_rays: bool = False, candidate_fraction: float = 0.0, reff_bin=None, hist=None,
                   device=None, max_retries: int = 3, packed=None) -> VisibilityResult:
        """Run K1a/K1b over the voxels [begin, end).

        *lattice* selects the implicit CuboidMesh lattice; *voxels* is a torch.float64 CUDA tensor ``(P, 3)``
        of explicit coordinates (global mm) - what the reference materialises (mesh_calculation.py:119-159); they are
        packed once into the recentred float4 array the FP32 filter streams (``com_voxels_pack``; pass the result of
        ``pack_voxels`` as *packed* to reuse it across calls).  Returns device tensors; nothing is copied to the host
        except the statistics block (and the ray list when *want_rays*).
        """
        import torch

        if not torch.cuda.is_available():
            raise _lib.ComonitorError(_lib.COM_E_CUDA, "ChannelScene.visibility",
                                      "no CUDA device: the accelerated path has no CPU fallback")
        if (lattice is None) == (voxels is None):
            raise ValueError("give exactly one of lattice= or voxels=")
        device = torch.device(device if device is not None else "cuda")
        lat_struct = None
        vox_ptr = None
        if lattice is not None:
            total = lattice.n_points
            lat_struct = lattice_struct(lattice)
        else:
            if voxels.dtype != torch.float64 or not voxels.is_cuda or voxels.dim() != 2 or voxels.shape[1] != 3:
                raise ValueError("voxels must be a CUDA float64 tensor of shape (P, 3)")
            voxels = voxels.contiguous()
            total = voxels.shape[0]
            vox_ptr = C.c_void_p(voxels.data_ptr())
            device = voxels.device
            if packed is None:
                packed = self.pack_voxels(voxels)
            if packed.dtype != torch.float32 or tuple(packed.shape) != (total, 4) or packed.device != device:
                raise ValueError("packed must be the float32 (P, 4) tensor pack_voxels returned for these voxels")
        end = total if end is None else end
        if not (0 <= begin <= end <= total):
            raise ValueError(f"voxel range [{begin},{end}) outside [0,{total})")
        n = end - begin
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream().cuda_stream
            w = torch.empty(n, dtype=torch.float64, device=device)
            nr = torch.empty(n, dtype=torch.int32, device=device)
            stats_t = torch.empty(_lib.N_STATS, dtype=torch.int64, device=device)
            frac = candidate_fraction
            rays_cap = 0
            for attempt in range(max_retries + 1):
                ws_bytes = self._lib.com_visibility_workspace_bytes(self.handle, n, frac)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
                rays_t = rays_cnt = None
                if want_rays:
                    rays_cap = rays_cap or max(1 << 16, (ws_bytes - 256) // 8)
                    rays_t = torch.empty(rays_cap * 4, dtype=torch.float64, device=device)
                    rays_cnt = torch.zeros(1, dtype=torch.int64, device=device)
                out_args = (
                    C.c_void_p(w.data_ptr()), C.c_void_p(nr.data_ptr()),
                    C.c_void_p(rays_t.data_ptr()) if want_rays else None, rays_cap,
                    C.c_void_p(rays_cnt.data_ptr()) if want_rays else None,
                    C.c_void_p(reff_bin.data_ptr()) if reff_bin is not None else None,
                    C.c_void_p(hist.data_ptr()) if hist is not None else None,
                    int(hist.numel()) if hist is not None else 0,
                    C.c_void_p(stats_t.data_ptr()), C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(stream))
                if lat_struct is not None:
                    rc = self._lib.com_visibility_weights(self.handle, C.byref(lat_struct), None, begin, end, *out_args)
                else:
                    rc = self._lib.com_visibility_weights_packed(self.handle, vox_ptr, C.c_void_p(packed.data_ptr()), begin, end,
                                                                 *out_args)
                _lib.check(rc, "com_visibility_weights")
                stats = dict(zip([f[0] for f in _lib.ComStats._fields_], stats_t.cpu().tolist()))
                if not stats["overflow"]:
                    break
                if hist is not None:
                    raise _lib.ComonitorError(_lib.COM_E_OVERFLOW, "com_visibility_weights",
                                              "candidate overflow while accumulating into a caller histogram")
                # the filter let more rays through than the list holds: size it from the count and redo
                frac = 1.05 * stats["candidates"] / max(1, stats["tests"]) + 1e-6
            else:
                raise _lib.ComonitorError(_lib.COM_E_OVERFLOW, "com_visibility_weights", "candidate list overflow")
            rays = None
            if want_rays:
                cnt = int(rays_cnt.item())
                if cnt > rays_cap:
                    raise _lib.ComonitorError(_lib.COM_E_OVERFLOW, "com_visibility_weights",
                                              f"{cnt} passing rays > ray-list capacity {rays_cap}")
                rays = rays_t[: cnt * 4].cpu().numpy().view(_lib.RAY_RECORD_DTYPE).copy()
        return VisibilityResult(begin=begin, end=end, weight=w, nrays=nr, stats=stats, rays=rays)

    def pack_voxels(self, voxels):
        """``(P, 3)`` float64 CUDA coordinates -> the ``(P, 4)`` float32 array the FP32 filter streams: (x - ox, y - oy,
        z - oz, 1) recentred on this scene's origin (``com_voxels_pack``; stream-ordered)."""
        import torch

        if voxels.dtype != torch.float64 or not voxels.is_cuda or voxels.dim() != 2 or voxels.shape[1] != 3:
            raise ValueError("voxels must be a CUDA float64 tensor of shape (P, 3)")
        voxels = voxels.contiguous()
        with torch.cuda.device(voxels.device):
            out = torch.empty((voxels.shape[0], 4), dtype=torch.float32, device=voxels.device)
            rc = self._lib.com_voxels_pack(self.handle, C.c_void_p(voxels.data_ptr()), 0, voxels.shape[0],
                                           C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "com_voxels_pack")
        return out

    def plan(self, lattice: Lattice, begin: int = 0, end: int | None = None, candidate_fraction: float = 0.0,
             device=None, **kwargs) -> "VisibilityPlan":
        """Pre-allocated, sync-free launcher for repeated Stage-1 runs over the same voxel range."""
        return VisibilityPlan(self, lattice, begin, lattice.n_points if end is None else end, candidate_fraction, device,
                              **kwargs)

    def compact(self, result: VisibilityResult):
        """Visible voxels of *result*: (flat indices int64, weights float64), device tensors, ascending."""
        import torch

        n = result.end - result.begin
        dev = result.weight.device
        with torch.cuda.device(dev):
            idx = torch.empty(n, dtype=torch.int64, device=dev)
            wc = torch.empty(n, dtype=torch.float64, device=dev)
            cnt = torch.zeros(1, dtype=torch.int64, device=dev)
            ws_bytes = self._lib.com_compact_workspace_bytes(n)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            rc = self._lib.com_compact_visible(
                C.c_void_p(result.nrays.data_ptr()), C.c_void_p(result.weight.data_ptr()), n, result.begin,
                C.c_void_p(idx.data_ptr()), C.c_void_p(wc.data_ptr()), C.c_void_p(cnt.data_ptr()),
                C.c_void_p(ws.data_ptr()), ws_bytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            _lib.check(rc, "com_compact_visible")
            m = int(cnt.item())
        return idx[:m], wc[:m]


class VisibilityPlan:
    """Device buffers + arguments of one ``com_visibility_weights`` call, reusable without host syncs."""

    def __init__(self, scene: ChannelScene, lattice: Lattice, begin: int, end: int, candidate_fraction: float, device,
                 reff_bin=None, hist=None, capacity: int | None = None):
        """*reff_bin* (uint16 CUDA tensor over the whole lattice/shard) and *hist* (float64 CUDA tensor): the exact
        kernel also accumulates every passing ray's weight into ``hist[reff_bin[voxel]]`` (bins >= len(hist) skipped).
        *capacity*: size the buffers for ranges of up to that many voxels (``rebind`` then moves the range)."""
        import torch

        self.scene, self.begin, self.end = scene, begin, end
        self.device = torch.device(device if device is not None else "cuda")
        n = max(end - begin, capacity or 0)
        self.capacity = n
        self._lat = lattice_struct(lattice)
        self.reff_bin, self.hist = reff_bin, hist
        if (reff_bin is None) != (hist is None):
            raise ValueError("reff_bin and hist go together")
        with torch.cuda.device(self.device):
            self.weight = torch.empty(n, dtype=torch.float64, device=self.device)
            self.nrays = torch.empty(n, dtype=torch.int32, device=self.device)
            self.stats_t = torch.zeros(_lib.N_STATS, dtype=torch.int64, device=self.device)
            self.ws_bytes = scene._lib.com_visibility_workspace_bytes(scene.handle, n, candidate_fraction)
            self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self._stats = self.stats_t

    def rebind(self, begin: int, end: int, stats=None) -> "VisibilityPlan":
        """Move the plan to the voxels [begin, end) (at most ``capacity`` of them); the outputs are then
        ``weight[:end-begin]`` / ``nrays[:end-begin]``.  *stats*: another 8 x int64 CUDA tensor to leave the statistics in."""
        if not 0 <= end - begin <= self.capacity:
            raise ValueError(f"range of {end - begin} voxels, plan capacity {self.capacity}")
        self.begin, self.end = begin, end
        if stats is not None:
            self._stats = stats
        return self

    def launch(self, stream=None) -> None:
        import torch

        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        s = self.scene
        rc = s._lib.com_visibility_weights(
            s.handle, C.byref(self._lat), None, self.begin, self.end, C.c_void_p(self.weight.data_ptr()),
            C.c_void_p(self.nrays.data_ptr()), None, 0, None,
            C.c_void_p(self.reff_bin.data_ptr()) if self.reff_bin is not None else None,
            C.c_void_p(self.hist.data_ptr()) if self.hist is not None else None,
            int(self.hist.numel()) if self.hist is not None else 0, C.c_void_p(self._stats.data_ptr()),
            C.c_void_p(self.ws.data_ptr()), self.ws_bytes, C.c_void_p(stream))
        _lib.check(rc, "com_visibility_weights")

    def stats(self) -> dict:
        out = dict(zip([f[0] for f in _lib.ComStats._fields_], self._stats.cpu().tolist()))
        return out


def make_channel(element: str, slits_number: int = 10, crystal_height_step: int = 20, crystal_length_step: int = 80,
                 lattice: Lattice | None = None, closing_side: str = "top closing side", **scene_kwargs) -> ChannelScene:
    """Channel scene from the reference's setup classes (what ``Simulation.__init__`` collects, :77-141).

    *closing_side*: the reference always simulates "top closing side" (simulation.py:91); the geometry database also
    holds "bottom closing side" (collimator.py:57-68), selectable here."""
    from .collimator import Collimator
    from .detector import Detector
    from .dispersive_element import DispersiveElement
    from .port import Port

    col = Collimator(element, closing_side, slits_number)
    _, crys, plas = col.read_colim_coord()
    de = DispersiveElement(element, crystal_height_step, crystal_length_step)
    port, det = Port(), Detector(element)
    normals = crystal_normals(de.crystal_points, de.radius_central_point, de.B, de.C)
    area = round(20 * 80 / crystal_height_step / crystal_length_step, 2)
    corners = lattice.bounding_corners() if lattice is not None else None
    return ChannelScene(de.crystal_points, normals, crys, plas, col.vector_front_back, port.vertices_coordinates,
                        port.orientation_vector, det.vertices, det.orientation_vector,
                        origin=de.crystal_central_point, area_pt=area, refl_max=de.max_reflectivity,
                        aoi0=float(de.AOI), voxel_corners=corners, **scene_kwargs)
