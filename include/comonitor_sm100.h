/*
 * comonitor_sm100.h - C ABI of libcomonitor_sm100.so
 *
 * B200 (sm_100a) implementation of the two data-parallel stages of tfornal/co_monitor.
 * The reference is pure Python and has no FFI layer; the entry points below are what a
 * binding of its two hot classes would call (ctypes stub in INTEGRATION.md):
 *
 *   Stage 1  co_monitor/geometry/simulation.py:30-75   Simulation.__init__
 *            (calculate_radiation_reflection :273-325, check_ray_transmission :327-432,
 *             calculate_distances :483-504, calculate_angles :506-536,
 *             calculate_radiation_fraction :538-594)                  -> com_visibility_weights*
 *   Stage 2  co_monitor/emissivity/reader.py:22-81     Emissivity.__init__
 *            (read_ne_te :199-243, find_closest_temp_idx :282-299, read_pec :316-341,
 *             calculate_intensity :343-378, calculate_total_emissivity :380-401)
 *                                                                      -> com_emissivity_integrate*
 *
 * Conventions
 *   - plain C: pointers and sizes only.  "_h" / "host" = host memory; everything else that is a
 *     buffer is DEVICE memory owned by the caller.  The library owns only what a ComScene /
 *     ComTables handle holds.
 *   - every function returns COM_OK (0) or a negative COM_E_* code; nothing throws.  A short text
 *     for the last failure on the calling thread is available from com_last_error().
 *   - work is enqueued on the caller's CUDA stream (com_stream_t == cudaStream_t); the *_host
 *     variants copy, run and synchronise by themselves.
 *   - lengths are in elements, coordinates in millimetres, fp64 unless stated.
 *   - there is no CPU fallback: without a CUDA device the compute entry points return COM_E_CUDA.
 */
#ifndef COMONITOR_SM100_H
#define COMONITOR_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COM_ABI_VERSION 4

#define COM_OK 0
#define COM_E_BADARG (-1)   /* null pointer, bad size, inconsistent description            */
#define COM_E_CUDA (-2)     /* CUDA runtime error (no device, launch failure, ...)         */
#define COM_E_NOMEM (-3)    /* host or device allocation failed / workspace too small      */
#define COM_E_OVERFLOW (-4) /* candidate / ray list capacity exceeded (see ComStats)       */
#define COM_E_GEOMETRY (-5) /* degenerate aperture (collinear plane points, empty section) */
#define COM_E_EMPTY (-6)    /* no ray passed: the reference's "Not enough points to perform
                               calculations!" (simulation.py:451-456)                      */
#define COM_E_IO (-7)       /* a result file could not be opened or written                */

typedef void* com_stream_t; /* cudaStream_t */

#define COM_MAX_EDGES 32
#define COM_FILTER_NO_RUN_CULL 1
#define COM_FILTER_NO_INTERVALS 2
#define COM_MAX_SHIELDS 4
/* ComSceneDesc.physics_flags - optional physics the reference does not compute; 0 = parity path */
#define COM_PHYS_SEGMENT_CHECK 1 /* an aperture only counts where the physical ray meets it: on the segment voxel ->
                                    crystal point for slits, port and shields, on the half-line leaving the crystal for
                                    the detector.  The reference intersects infinite lines (simulation.py:143-173) */

/* One ECRH stray-radiation shield reduced to a circular stop (extension; the reference only draws the shields:
 * radiation_shield.py:52-101 builds, per shield, a 1 mm long cylinder of `radius` around `central point` along
 * `orientation vector`, 360 vertices per rim, and takes its convex hull).  A ray passes the shield iff the point where
 * its line meets the mid-plane of that cylinder lies inside the hull's cross-section: the regular n_sides-gon with
 * vertex k at  centre + radius * (cos(2 pi k / n_sides) u + sin(2 pi k / n_sides) v). */
typedef struct ComShield {
  double centre[3]; /* point of the mid-plane on the cylinder axis                           */
  double axis[3];   /* unit normal of the plane (cylinder axis)                              */
  double u[3];      /* in-plane orthonormal basis; vertex 0 of the rim polygon lies along +u */
  double v[3];
  double radius;    /* circumradius of the rim polygon, mm                                    */
  int32_t n_sides;  /* 360 in the reference's model; 0 = a true circle                        */
  int32_t reserved;
} ComShield;

/* One aperture of the beam line reduced to "plane + convex polygon in that plane".
 * The reference tests  Delaunay(vertices +/- orientation_vector).find_simplex(X) >= 0  for the point X
 * where the ray meets the plane through the aperture's first three vertices
 * (simulation.py:175-207, 209-271).  X lies in that plane, so the test equals "X inside the
 * cross-section of the slab's convex hull with the plane" - a convex polygon. */
typedef struct ComAperture {
  double p1[3];     /* first vertex: the plane point used by the reference                    */
  double normal[3]; /* (p2-p1) x (p3-p2), NOT normalised: the |N.u| < 1e-6 rule needs it      */
  double e1[3];     /* in-plane orthonormal basis, e1 = (p2-p1)/|p2-p1|, e2 = n^ x e1         */
  double e2[3];
  int32_t n_edges;
  int32_t reserved;
  double edge[COM_MAX_EDGES][3];   /* a*x + b*y + c >= 0 inside, (a,b) unit, (x,y) = ((X-p1).e1,(X-p1).e2) */
  double vertex[COM_MAX_EDGES][2]; /* polygon vertices, counter-clockwise, same frame          */
} ComAperture;

/* Host-only helper (no CUDA): cross-section of conv(verts + offset, verts - offset) with the plane through
 * verts[0..2].  verts is (n,3) row-major, 3 <= n <= 16. */
int com_aperture_from_slab(const double* verts_h, int32_t n, const double* offset_h, ComAperture* out_h);

/* Host-only helper (no CUDA): 1 if the point (global mm) passes the aperture test in fp64, else 0.
 * Replaces Simulation.check_in_hull for one point (simulation.py:234-271). */
int com_aperture_contains(const ComAperture* ap_h, const double* point_h);

/* The beam line of one spectroscopic channel, at the level of the reference's own arrays. */
typedef struct ComSceneDesc {
  int32_t n_crystal;             /* C = crystal_height_step * crystal_length_step                     */
  int32_t n_slits;               /* S                                                                 */
  const double* crystal_xyz;     /* (C,3) DispersiveElement.crystal_points                            */
  const double* crystal_normal;  /* (C,3) unit cylinder normals (simulation.py:286-312)               */
  const double* slit_crys_side;  /* (S,4,3) Collimator.read_colim_coord()[1]                          */
  const double* slit_plasma_side;/* (S,4,3) Collimator.read_colim_coord()[2]                          */
  double slit_depth[3];          /* collimator "vector front-back" (slab half-depth of every slit)    */
  int32_t n_port_vertices;       /* 14                                                                */
  int32_t n_detector_vertices;   /* 4                                                                 */
  const double* port_vertices;   /* (n_port_vertices,3)                                               */
  double port_depth[3];          /* Port.orientation_vector                                           */
  const double* detector_vertices; /* (n_detector_vertices,3)                                         */
  double detector_depth[3];      /* Detector.orientation_vector                                       */
  double origin[3];              /* recentring offset of the FP32 filter (crystal central point)      */
  double area_pt;                /* round(1600/(h*l), 2) mm^2, simulation.py:137-139                  */
  double refl_max;               /* max reflectivity                                                  */
  double aoi0;                   /* nominal angle of incidence, degrees                               */
  double refl_sigma;             /* 2.2                                                               */
  double filter_eps_mm;          /* FP32 filter band; <=0 selects the default (5e-3 mm);
                                    >= 1e30 disables the filter: every ray is evaluated in fp64       */
  const ComAperture* apertures_override; /* NULL, or 2S+2 apertures (crys_0, plasma_0, ..., port, detector)
                                    to use instead of the exact hull sections computed from the vertex lists.
                                    Lets the host reproduce scipy/Qhull's in-simplex leeway: find_simplex accepts
                                    points up to sqrt(100*eps) = 1.5e-7 (barycentric) outside hull faces that
                                    Qhull covered with a flat simplex, i.e. up to ~1e-4 mm beyond an edge.     */
  double near_edge_mm;           /* |margin| below which a decided ray is counted in ComStats.near_edge_rays;
                                    <= 0 selects 1e-9 mm                                               */
  int32_t n_voxel_corners;       /* optional hint, 0..8: points whose convex hull contains every voxel that   */
  int32_t filter_flags;          /* will be tested (the lattice corners).  Not needed for correctness: the    */
  double voxel_corners[8][3];    /* filter accepts both nappes of the detector cone, as the reference's
                                    infinite-line intersection does (simulation.py:143-173).
                                    filter_flags: COM_FILTER_NO_RUN_CULL (1) makes the lattice kernel evaluate the
                                    detector test of every ray individually instead of skipping whole boxes, columns
                                    and z-runs that the linearity of the forms proves to miss (a measurement aid: same
                                    results).   */
  double near_band_mm;           /* second, wider near-edge counter (ComStats.near_band_rays): the width that bounds the
                                    residual disagreement between the polygon test and scipy's find_simplex on edges
                                    the calibration cannot move exactly; <= 0 selects 1e-4 mm                     */
  int32_t physics_flags;         /* COM_PHYS_* bits; 0 = what the reference computes                            */
  int32_t n_shields;             /* 0..COM_MAX_SHIELDS circular stops every ray must also pass (extension)      */
  const ComShield* shields;      /* (n_shields) HOST                                                            */
} ComSceneDesc;

typedef struct ComScene ComScene; /* opaque: device-resident tables of one channel */

/* Uploads what Simulation.__init__ collects before the ray work starts (simulation.py:77-141: _init_collimator,
 * _init_dispersive_element, _init_port, _init_detector) plus the slabs of make_spatial_object (:209-232) reduced to
 * apertures; com_scene_destroy frees it. */
int com_scene_create(const ComSceneDesc* desc_h, ComScene** out);
int com_scene_destroy(ComScene* scene);
/* number of apertures (2S+2, order: crys_0, plasma_0, crys_1, ..., port, detector - the order of the 22 in-hull
 * tests of check_ray_transmission, simulation.py:344-400) and a host copy of one */
int com_scene_aperture_count(const ComScene* scene);
int com_scene_get_aperture(const ComScene* scene, int32_t index, ComAperture* out_h);
/* Extension (not in the reference; SURVEY.md 8(d) config 3 "per-pixel detector resolution"): while pix_out != NULL every
 * com_visibility_weights call on this scene also accumulates each passing ray's weight into a detector image,
 * pix_out[py*nx + px] (DEVICE, nx*ny fp64, never cleared by the library), where (x, y) is the ray's hit on the detector
 * plane in the detector aperture's frame (origin = first detector vertex, e1 along vertex 1 -> 2), px = floor((x - x0)/
 * pixel), py = floor((y - y0)/pixel), clamped to the image, so that sum(image) == sum(weights).  NULL switches it off. */
int com_scene_set_detector_image(ComScene* scene, double* pix_out, int32_t nx, int32_t ny, double pixel_mm, double x0_mm,
                                 double y0_mm);

/* Host-only (no CUDA): the FP32 filter tables com_scene_create would upload, for inspection and CPU-side tests.
 * (No counterpart in the reference: the filter is a conservative pre-selection in front of simulation.py:327-432.)
 *   det_forms_out  [3][Cp][4] floats: forms Xs, Ys, W of the mirrored detector, value = g.p' + h with p' = p - origin
 *   slit_forms_out [Cp][6][4] floats: X', Y', W on the crystal-side slit plane, then on the plasma side
 *   rects_out      8 floats: xlo, xhi, ylo, yhi of the crystal-side and of the plasma-side slit rectangle (band included)
 *   flags_out      3 ints: collimator periodic (stage B active), both sides share W, filter disabled
 * Any output pointer may be NULL; *n_crystal_padded (Cp) is always written - call once with NULLs to size the buffers. */
int com_scene_filter_tables_host(const ComSceneDesc* desc_h, int32_t* n_crystal_padded, float* det_forms_out,
                                 float* slit_forms_out, float* rects_out, float* period_out, int32_t* flags_out);

/* Host-only (no CUDA): the "certain pass" regions of the filter.  out9 = xlo, xhi, ylo, yhi of the inner rectangle of
 * the crystal-side slit polygon, the same for the plasma side (every point of them is >= out9[8] mm inside the
 * polygon; xlo > xhi: shortcut off), then that distance, max(2e-2, 2 near_edge_mm).  A candidate the FP32 filter finds
 * inside both rectangles of one slit skips the two fp64 slit tests of the exact kernel - they could not fail
 * (check_ray_transmission, simulation.py:344-369, would find it in that slit); port and detector are still tested. */
int com_scene_filter_certain_host(const ComSceneDesc* desc_h, float* out9);

/* Host-only (no CUDA): the extra tables of the column-interval form of the lattice filter.  rects12 = xlo, xhi, ylo, yhi of
 * the detector's certain rectangle in the scaled frame of det_forms (the grown bounding rectangle is |x|, |y| <= 1), of
 * the port's bounding rectangle (band included) and of the port's certain rectangle, both in the port aperture's frame
 * (mm); *interval_filter = 1 when the geometry allows the interval form (periodic collimator with one shared normal,
 * disjoint slit rectangles, filter enabled, no COM_FILTER_NO_* flag); port_forms_out [Cp][3][4] floats: X', Y', W on
 * the port plane; slit_edges_out [2][8][4] floats: per side (crystal, plasma) the edges of slit 0's polygon as
 * (a, b, c + eps, c - certain_mm) with a x + b y + c >= 0 inside, n_slit_edges_out [2] how many are used; certain_off_out
 * 1 when no ray is vouched for (extensions on).  Any output pointer may be NULL. */
int com_scene_interval_tables_host(const ComSceneDesc* desc_h, float* rects12, int32_t* interval_filter,
                                   float* port_forms_out, float* slit_edges_out, int32_t* n_slit_edges_out,
                                   int32_t* certain_off_out);
/* The port polygon as the filters' certain test uses it: edges_out64 [16][4] = (a, b, c + eps, c - certain_mm) per edge
 * with a x + b y + c >= 0 inside (aperture frame, mm), *n_edges_out how many are used (0: the inner rectangle is used). */
int com_scene_port_edges_host(const ComSceneDesc* desc_h, float* edges_out64, int32_t* n_edges_out);
/* The detector polygon likewise, in the SCALED frame of the detector forms (|Xs|, |Ys| <= |W| is the grown bounding
 * rectangle): edges_out32 [8][4] = (A, B, C + eps, C - certain_mm) with A Xs + B Ys + C W >= 0 (times sign W) inside. */
int com_scene_detector_edges_host(const ComSceneDesc* desc_h, float* edges_out32, int32_t* n_edges_out);

/* what the FP32 filter was built with (any output pointer may be NULL; no counterpart in the reference, whose slits
 * are tested one by one, simulation.py:344-369): whether the collimator was found periodic
 * (stage B of the filter active), whether the filter is disabled, the slit pitch and the band in mm */
int com_scene_filter_info(const ComScene* scene, int32_t* slit_filter, int32_t* filter_disabled,
                          double* slit_period_mm, double* eps_mm);

/* Implicit voxel lattice == CuboidMesh (mesh_calculation.py:119-159):
 *   flat = (iy*nx + ix)*nz + iz ;  p = ((x[ix],y[iy],z[iz]) - centre) . R + centre + shift
 *   x[i] = start[0] + i*step[0]  (last sample = stop[0], as numpy.linspace does)
 * Sharding (multi-GPU): with shard_count > 1 the structure describes ONE SHARD of the lattice.  The nx*ny z-columns are
 * dealt round-robin in blocks of shard_cols_per_block columns; shard shard_index owns blocks shard_index,
 * shard_index + shard_count, ...  Every flat index the API takes or returns is then LOCAL to the shard:
 *   local column cl = flat / nz  ->  global column ((cl / cpb) * shard_count + shard_index) * cpb + cl % cpb.
 * Block-cyclic shards keep the (spatially clustered) visible voxels balanced over the ranks. */
typedef struct ComLattice {
  int64_t nx, ny, nz;
  double start[3];
  double step[3];
  double stop[3];
  double centre[3];
  double rotation[9]; /* row-major R, row-vector convention (p_row . R) */
  double shift[3];
  int64_t shard_cols_per_block; /* 0 or shard_count <= 1: not sharded */
  int64_t shard_count;
  int64_t shard_index;
} ComLattice;

/* Number of voxels of the shard described by *lattice_h* (nx*ny*nz when not sharded). */
int64_t com_lattice_local_voxels(const ComLattice* lattice_h);

typedef struct ComStats {
  uint64_t tests;           /* voxel x crystal-point rays examined                                   */
  uint64_t candidates;      /* rays the FP32 filter sent to the fp64 stage                           */
  uint64_t passing_rays;    /* rays that pass every aperture in fp64                                 */
  uint64_t near_edge_rays;  /* fp64-decided rays within near_edge_mm of an aperture edge they were tested on */
  uint64_t visible_voxels;  /* voxels with >= 1 passing ray (counted from nrays_out: 0 when that is NULL;
                               com_volume_histograms counts voxels with a non-zero weight instead)   */
  uint64_t candidate_capacity; /* capacity the launch ran with                                       */
  uint64_t overflow;        /* bit 0: candidates > capacity - the launch accumulated NOTHING (COM_E_OVERFLOW);
                               bit 1: an index guard of the fp64 kernels fired (never expected)      */
  uint64_t certain_candidates; /* candidates the FP32 filter vouched for (>= certain_mm inside every aperture):
                                  only their weight is computed in fp64 (lattice kernels only)               */
  uint64_t near_band_rays;  /* like near_edge_rays at the wider near_band_mm (default 1e-4 mm)                */
  uint64_t shield_blocked_rays;   /* rays through every reference aperture but stopped by an ECRH shield (extension) */
  uint64_t segment_rejected_rays; /* rays the infinite-line test accepts but COM_PHYS_SEGMENT_CHECK rejects (extension) */
  uint64_t near_rounding_rays;    /* passing rays whose angle deg * 100 lies within 1e-9 of a half-integer: the only rays
                                     whose round(deg, 2) (simulation.py:527) could differ from the reference's by one step */
} ComStats;

/* One passing ray (optional output): what the reference keeps per selected ray
 * (calculate_indices :434-461, calculate_distances :483-504, calculate_angles :506-536). */
typedef struct ComRayRecord {
  int64_t ray_index; /* voxel_flat_index * C + crystal_index */
  double distance;   /* |plasma - crystal|, mm              */
  double aoi;        /* (180 - round(deg, 2)) / 2           */
  double weight;     /* contribution to total_intensity_fraction */
} ComRayRecord;

/* Bytes of scratch needed by com_visibility_weights for n voxels against this scene: 256 B of counters, the candidate
 * list (rays the FP32 filter could not reject, 8 B each; the reference materialises all P x C rays instead,
 * simulation.py:316-321) and the column-interval filter's records (16 B each, as many slots as the candidate list).
 * candidate_fraction <= 0 selects the default (2 % of the rays, at least 1 Mi entries). */
size_t com_visibility_workspace_bytes(const ComScene* scene, int64_t n_voxels, double candidate_fraction);

/* Stage 1 on the voxels [vox_begin, vox_end).  Replaces, per (voxel, crystal point) ray, calculate_radiation_reflection
 * (simulation.py:273-325), find_intersection_points / line_plane_intersection (:143-207), check_in_hull (:234-271),
 * check_ray_transmission (:327-432), calculate_indices (:434-481), calculate_distances (:483-504), calculate_angles
 * (:506-536) and the per-ray weight and per-voxel sum of calculate_radiation_fraction (:538-577).
 *   lattice_h != NULL : voxel coordinates come from the implicit lattice (flat indices are lattice indices);
 *   vox_xyz   != NULL : (n_total,3) fp64 DEVICE array in global mm, indexed by the same flat index.
 * Outputs are indexed by (flat - vox_begin):
 *   w_out     (n) fp64   sum of the passing rays' weights (0 for invisible voxels) - overwritten
 *   nrays_out (n) u32    number of passing rays per voxel (visibility mask = nrays > 0) - overwritten, nullable
 *   rays_out  optional list of passing rays, unordered; *rays_count_out (device u64) receives how many
 *             were produced (may exceed rays_capacity: then only the first rays_capacity are valid)
 *   reff_bin / hist_out optional fused Reff-bin histogram: hist_out[reff_bin[flat]] += weight
 *             (reff_bin indexed by absolute flat index; bins >= n_bins are skipped) - accumulated, not cleared
 *   stats_out device ComStats - overwritten
 * Stream-ordered; returns after enqueueing. */
int com_visibility_weights(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz,
                           int64_t vox_begin, int64_t vox_end, double* w_out, uint32_t* nrays_out,
                           ComRayRecord* rays_out, int64_t rays_capacity, uint64_t* rays_count_out,
                           const uint16_t* reff_bin, double* hist_out, int32_t n_bins, ComStats* stats_out,
                           void* workspace, size_t workspace_bytes, com_stream_t stream);

/* Explicit voxels the way the kernels want them (north-star: "voxel coordinates streamed from HBM as coalesced float4"):
 * the reference materialises (P,3) fp64 coordinates (mesh_calculation.py:119-159); the fp64 decision kernel keeps reading
 * those (only for the ~0.5 % of rays that reach it), the FP32 filter reads a packed copy - one 16-byte element
 * (x - ox, y - oy, z - oz, 1) per voxel, recentred on the scene origin - whose 512-voxel tiles are moved into shared memory
 * by bulk asynchronous copies (cp.async.bulk, double buffered).
 *   com_voxels_pack                 vox_packed_out[v] for v in [vox_begin, vox_end) from vox_xyz (DEVICE, indexed by v);
 *                                   vox_packed_out: DEVICE, 16 x P bytes, 16-byte aligned
 *   com_visibility_weights_packed   com_visibility_weights for explicit voxels with the packed copy supplied (results are
 *                                   identical to passing vox_xyz alone: the same FP32 values enter the filter) */
int com_voxels_pack(const ComScene* scene, const double* vox_xyz, int64_t vox_begin, int64_t vox_end, void* vox_packed_out,
                    com_stream_t stream);
int com_visibility_weights_packed(const ComScene* scene, const double* vox_xyz, const void* vox_packed, int64_t vox_begin,
                                  int64_t vox_end, double* w_out, uint32_t* nrays_out, ComRayRecord* rays_out,
                                  int64_t rays_capacity, uint64_t* rays_count_out, const uint16_t* reff_bin,
                                  double* hist_out, int32_t n_bins, ComStats* stats_out, void* workspace,
                                  size_t workspace_bytes, com_stream_t stream);

/* Host-buffer variant (the reference-facing call: everything Simulation.__init__ does after its _init_* methods,
 * simulation.py:50-72): host in, host out, device work inside; synchronous.
 * vox_xyz_h may be NULL when lattice_h is given.  rays_out_h may be NULL.  Returns COM_E_EMPTY when no
 * ray passes (outputs are still valid: all zero). */
int com_visibility_weights_host(const ComSceneDesc* desc_h, const ComLattice* lattice_h, const double* vox_xyz_h,
                                int64_t vox_begin, int64_t vox_end, double* w_out_h, uint32_t* nrays_out_h,
                                ComRayRecord* rays_out_h, int64_t rays_capacity, uint64_t* rays_count_out_h,
                                ComStats* stats_out_h);

/* Same, but only the visible voxels come back (what Simulation.ddf holds, simulation.py:579-594): ascending flat
 * indices idx_out_h[i] and their weights w_out_h[i], i < *count_out_h <= capacity (COM_E_OVERFLOW beyond). */
int com_visible_weights_host(const ComSceneDesc* desc_h, const ComLattice* lattice_h, const double* vox_xyz_h,
                             int64_t vox_begin, int64_t vox_end, int64_t* idx_out_h, double* w_out_h,
                             int64_t capacity, uint64_t* count_out_h, ComStats* stats_out_h);

/* The same two calls for a caller that keeps the ComScene handle (com_scene_create) across calls - e.g. one Simulation
 * per voxel range, or a sweep over lattices: the scene is neither rebuilt nor uploaded again, the compacted result is
 * written straight into pinned host memory by the device, and the call waits once (blocking, not spinning). */
int com_visibility_weights_host_scene(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz_h,
                                      int64_t vox_begin, int64_t vox_end, double* w_out_h, uint32_t* nrays_out_h,
                                      ComRayRecord* rays_out_h, int64_t rays_capacity, uint64_t* rays_count_out_h,
                                      ComStats* stats_out_h);
int com_visible_weights_host_scene(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz_h,
                                   int64_t vox_begin, int64_t vox_end, int64_t* idx_out_h, double* w_out_h,
                                   int64_t capacity, uint64_t* count_out_h, ComStats* stats_out_h);

/* (No counterpart in the reference.)  The *_host calls keep one grow-only device block and one pinned staging block per calling thread between calls
 * (cudaMalloc and cudaFree synchronise the device) and run on the per-thread CUDA stream: they may be called
 * concurrently from several host threads (e.g. one channel per thread).  This frees the calling thread's blocks. */
int com_host_arena_release(void);

/* Compaction of the visible voxels (nrays > 0) - the group-by of calculate_radiation_fraction (simulation.py:564-577) -
 * ascending flat index: idx_out[i] = vox_begin + position,
 * w_compact[i] = w[position].  *count_out (device u64).  idx_out / w_compact need room for n entries. */
int com_compact_visible(const uint32_t* nrays, const double* w, int64_t n, int64_t vox_begin, int64_t* idx_out,
                        double* w_compact, uint64_t* count_out, void* workspace, size_t workspace_bytes,
                        com_stream_t stream);
size_t com_compact_workspace_bytes(int64_t n);

/* ------------------------------------------------------------------------------------------------------------
 * Stage 2: line-emissivity integration (reader.py:199-401).
 * All tables are fp64 host arrays copied to the device by com_tables_create. */
#define COM_MAX_TRANSITIONS 4
#define COM_PEC_REFERENCE 0
#define COM_PEC_LOGLOG 1

typedef struct ComPecDesc {
  int32_t n_ne, n_te;        /* ADAS block size (24 x 24; B: 8 x 11), pec.py:122-151                           */
  int32_t n_grid;            /* interp_step = 2000, pec.py:169-174                                            */
  int32_t use_fully_stripped;/* 1 for RECOM (multiplies the last FA column), 0 otherwise (reader.py:355-369)  */
  int32_t interpolation;     /* COM_PEC_REFERENCE (0): linear interp2d on the n_grid x n_grid linspace table, read at the
                                nearest grid node (pec.py:153-187, reader.py:316-341) - the parity path.
                                COM_PEC_LOGLOG (1), extension off by default: bilinear in (log ne, log Te) of log PEC
                                through the ADAS nodes at the row's own (ne, Te), clamped to the node range            */
  int32_t reserved;
  const double* ne_nodes;    /* (n_ne)                                                                        */
  const double* te_nodes;    /* (n_te)                                                                        */
  const double* table;       /* (n_ne, n_te) row-major PEC values                                             */
  const double* ne_grid;     /* (n_grid) linspace(ne_nodes[0], ne_nodes[-1], n_grid): the nearest-index axis  */
  const double* te_grid;     /* (n_grid)                                                                      */
} ComPecDesc;

typedef struct ComTablesDesc {
  int32_t n_fa;              /* rows of the interpolated fractional-abundance table (19 995)                  */
  int32_t n_transitions;     /* 1..COM_MAX_TRANSITIONS                                                        */
  const double* fa_T;        /* (n_fa) arange(T[0], T[-1], 1), fractional_abundance.py:72                     */
  const double* fa_ion;      /* (n_fa) column of the requested ion, already / 100                             */
  const double* fa_last;     /* (n_fa) last column ("Fully stripped")                                         */
  ComPecDesc pec[COM_MAX_TRANSITIONS];
} ComTablesDesc;

typedef struct ComTables ComTables; /* opaque: device-resident Stage-2 tables of one line */

/* Uploads the tables Emissivity.__init__ builds before its per-voxel loops: FractionalAbundance.df_interpolated_frac_ab
 * (fractional_abundance.py:59-91; reader.py:245-261) and, per transition, the ADAS block of PEC (pec.py:122-151) with
 * its two linspace grids (pec.py:169-174) instead of the dense 2000 x 2000 table of _get_pec_data (reader.py:83-97). */
int com_tables_create(const ComTablesDesc* desc_h, ComTables** out);
int com_tables_destroy(ComTables* tables);

/* scratch of com_emissivity_integrate / com_weight_histogram for n voxels (block partials of the flux sums that
 * calculate_total_emissivity takes with DataFrame.sum(), reader.py:380-401) */
size_t com_emissivity_workspace_bytes(int64_t n);

/* Per-voxel pass == Emissivity.read_ne_te (reader.py:199-243), find_closest_temp_idx / assign_temp_accodring_to_indexes
 * (:282-314), read_pec (:316-341), calculate_intensity (:343-378), calculate_total_emissivity (:380-401) for n voxels.
 *   profile        DEVICE (K,3) row-major [Reff, T_e, n_e] (the reference's profiles_df column order)
 *   profile_sorted 1 when the Reff column is non-decreasing (binary search); 0 = literal linear argmin
 *   reff / reff_mm Reff in metres (NaN = outside the LCFS), or as uint16 millimetres (0xFFFF = NaN); exactly one
 *   reff_index     NULL: voxel i uses reff[i].  Else (n) int64: voxel i uses reff[reff_index[i]] - lets the compacted
 *                  visible-voxel list of com_compact_visible address a whole-lattice Reff table
 *   w              per-voxel total_intensity_fraction
 *   n / n_device   number of voxels; n_device (DEVICE u64, nullable) overrides n with a count produced on the device
 *                  (n is then only the upper bound) so that Stage 1 -> compaction -> Stage 2 needs no host sync
 *   mapped_reff_max NULL: the maximum mapped Reff is computed here and the cut is min(that, profile_reff_max)
 *                  (reader.py:210-217).  Else DEVICE (1) double holding the maximum already reduced by the caller -
 *                  sharded runs take com_reff_max per rank and all-reduce(MAX) it, all on the device
 *   flux_out       DEVICE (n_transitions + 1): sum of Emissivity_<t> * w, then of Emissivity_TOTAL * w
 *   per_voxel_out  DEVICE (n, 4 + 2*n_transitions + 1) or NULL: n_e, T_e, FA_ion, FA_stripped, pec_t..., Emissivity_t...,
 *                  Emissivity_TOTAL (rows of dropped voxels are left untouched)
 *   valid_out      DEVICE (n) bytes or NULL: 1 for voxels the reference keeps (finite Reff < boundary) */
int com_emissivity_integrate(const ComTables* tables, const double* profile, int32_t K, int32_t profile_sorted,
                             double profile_reff_max, double concentration, const double* reff,
                             const uint16_t* reff_mm, const int64_t* reff_index, const double* w, int64_t n,
                             const uint64_t* n_device, const double* mapped_reff_max, double* flux_out,
                             double* per_voxel_out, uint8_t* valid_out, void* workspace, size_t workspace_bytes,
                             com_stream_t stream);

/* Maximum finite Reff of the addressed voxels into max_out (DEVICE, 1 double); 0 when there is none: the
 * "Max reff from mapped vmec" of read_ne_te (reader.py:210-212). */
int com_reff_max(const double* reff, const uint16_t* reff_mm, const int64_t* reff_index, int64_t n,
                 const uint64_t* n_device, double* max_out, com_stream_t stream);

/* Host-buffer variant (the reference-facing call: what Emissivity.__init__ does after get_plas_coords_with_rad_frac,
 * reader.py:68-79): everything host, synchronous. */
int com_emissivity_integrate_host(const ComTablesDesc* tables_h, const double* profile_h, int32_t K,
                                  double concentration, const double* reff_h, const double* w_h, int64_t n,
                                  double* flux_out_h, double* per_voxel_out_h, uint8_t* valid_out_h,
                                  double* boundary_out_h);

/* Factorised form for profile sweeps (the loop over time slices of co_monitor/emissivity/main.py:39-60, one Emissivity
 * per slice) and very large grids: every quantity of reader.py:199-378 depends on the voxel only through its nearest
 * profile row, so the flux (reader.py:380-401) depends on the voxels only through
 * H[k] = sum of w over the voxels whose nearest profile row is k.
 *   com_weight_histogram  one pass over the voxels (10 B/voxel with uint16 Reff), hist_out DEVICE (K) overwritten,
 *                         boundary_out DEVICE (1) or NULL
 *   com_emissivity_sweep  flux_out DEVICE (n_slices, n_transitions + 1) for profiles DEVICE (n_slices, K, 3) that share
 *                         the Reff column the histogram was built with */
int com_weight_histogram(const double* profile, int32_t K, int32_t profile_sorted, double profile_reff_max,
                         const double* reff, const uint16_t* reff_mm, const int64_t* reff_index, const double* w,
                         int64_t n, const uint64_t* n_device, const double* mapped_reff_max, double* hist_out,
                         double* boundary_out, void* workspace, size_t workspace_bytes, com_stream_t stream);
/* The two halves of the uint16 fast path (same reference lines as com_weight_histogram: the Reff cut of read_ne_te,
 * reader.py:210-217 and 240-242, and its nearest-row lookup, :220-238), for pipelines that keep the millimetre histogram (e.g. to all-reduce it
 * over ranks: per-rank histograms simply add, and the Reff cut is then taken from the global one):
 *   com_weight_histogram_mm  hist_mm[min(reff_mm[v], COM_MM_BINS-1)] += w[v] for voxels with w != 0 and reff_mm != 0xFFFF;
 *                            hist_mm DEVICE (COM_MM_BINS) fp64, ACCUMULATED (clear it yourself); 16-byte aligned inputs
 *   com_histogram_rows       boundary = min(mapped max, profile_reff_max) with mapped max = largest non-empty bin / 1000
 *                            (or *mapped_reff_max), then hist_rows_out[row(b/1000)] += hist_mm[b] for b/1000 < boundary;
 *                            hist_rows_out DEVICE (K) overwritten, boundary_out DEVICE (1) or NULL */
#define COM_MM_BINS 4096
int com_weight_histogram_mm(const uint16_t* reff_mm, const double* w, int64_t n, const uint64_t* n_device,
                            double* hist_mm, com_stream_t stream);
int com_histogram_rows(const double* profile, int32_t K, int32_t profile_sorted, double profile_reff_max,
                       const double* hist_mm, const double* mapped_reff_max, double* hist_rows_out,
                       double* boundary_out, com_stream_t stream);
int com_emissivity_sweep(const ComTables* tables, const double* profiles, int32_t n_slices, int32_t K,
                         const double* hist, double concentration, double* flux_out, com_stream_t stream);

/* Same, with the Reff join of Emissivity.get_plas_coords_with_rad_frac (reader.py:161-197) done inside: voxel i has
 * Reff reff_table_h[reff_index_h[i]] (reff_table_h: the n_table values of the Reff file, NaN outside the LCFS;
 * reff_index_h: idx_sel_plas_points of the Stage-1 table).  NaN rows are dropped as the reference drops them. */
int com_emissivity_integrate_indexed_host(const ComTablesDesc* tables_h, const double* profile_h, int32_t K,
                                          double concentration, const double* reff_table_h, int64_t n_table,
                                          const int64_t* reff_index_h, const double* w_h, int64_t n, double* flux_out_h,
                                          double* per_voxel_out_h, uint8_t* valid_out_h, double* boundary_out_h);

/* Same as com_emissivity_integrate_indexed_host for a caller that keeps the ComTables handle (com_tables_create) across
 * calls - the loop over time slices of co_monitor/emissivity/main.py:39-60 builds one Emissivity per slice with the same
 * tables: they are then neither rebuilt nor uploaded again. */
int com_emissivity_integrate_indexed_host_tables(const ComTables* tables, const double* profile_h, int32_t K,
                                                 double concentration, const double* reff_table_h, int64_t n_table,
                                                 const int64_t* reff_index_h, const double* w_h, int64_t n,
                                                 double* flux_out_h, double* per_voxel_out_h, uint8_t* valid_out_h,
                                                 double* boundary_out_h);

/* ---- whole-volume runs (fused Stage 1 -> Stage 2; BASELINE config 5) ----------------
 * The reference writes a per-voxel table (Simulation.save_to_file, simulation.py:616-630) that Emissivity reads back
 * (reader.py:136-197).  For a lattice of 1e8-1e9 voxels that table is never built here: a ComVolume holds the Reff
 * millimetre bins of the voxels of one (shard of a) lattice on the device (2 B per voxel) and runs Stage 1 launch by
 * launch, each launch followed by the 10 B/voxel pass that folds its weights into the channel's millimetre histogram
 * (the only thing reader.py:199-401 needs of the voxels).  Per-rank histograms add: ranks all-reduce(SUM) the packed
 * (n_channels, COM_MM_BINS) buffer once, then every rank evaluates Stage 2 from the global histograms.
 *   chunk_voxels        voxels per Stage-1 launch (<= 0: 2^26), rounded up to a multiple of 8
 *   candidate_fraction  capacity of the per-launch candidate list as a fraction of the launch's rays (<= 0: 2 %) */
#define COM_MAX_CHANNELS 8
typedef struct ComVolume ComVolume;
int com_volume_create(const ComLattice* lattice_h, int64_t chunk_voxels, double candidate_fraction, ComVolume** out);
int com_volume_destroy(ComVolume* volume);
int64_t com_volume_voxels(const ComVolume* volume);
int64_t com_volume_chunk(const ComVolume* volume);
const uint16_t* com_volume_reff_mm(const ComVolume* volume); /* DEVICE (voxels) uint16, 0xFFFF outside the LCFS */
/* Reff bins of the volume's voxels (replaces reff_VMEC.py:85-136 like com_reff_resample / com_reff_quadratic).
 * src_reff: (src_ny, src_nx, src_nz) fp64 grid, HOST when src_is_host != 0 (copied up on the stream) else DEVICE. */
int com_volume_set_reff_resampled(ComVolume* volume, const double* src_reff, int32_t src_is_host, int64_t src_nx,
                                  int64_t src_ny, int64_t src_nz, com_stream_t stream);
int com_volume_set_reff_quadratic(ComVolume* volume, const double* coef_h, double reff_max, com_stream_t stream);
/* Stage 1 of every voxel of the volume for n_channels scenes (what n_channels Simulation runs would compute,
 * simulation.py:50-72) reduced to  hist_mm[c][mm] = sum of total_intensity_fraction over the voxels with that Reff
 * millimetre  (DEVICE (n_channels, COM_MM_BINS) fp64, overwritten) and the counters summed over the launches
 * (stats DEVICE (n_channels) ComStats, overwritten).  Stream-ordered, no host synchronisation. */
int com_volume_histograms(ComVolume* volume, int32_t n_channels, const ComScene* const* scenes, double* hist_mm,
                          ComStats* stats, com_stream_t stream);
/* Host-buffer form: hist_mm_h / stats_h are HOST arrays; src_reff_h != NULL also (re)builds the Reff bins from that
 * host grid first.  One device wait.  COM_E_OVERFLOW when a launch overflowed its candidate list. */
int com_volume_histograms_host(ComVolume* volume, int32_t n_channels, const ComScene* const* scenes,
                               const double* src_reff_h, int64_t src_nx, int64_t src_ny, int64_t src_nz,
                               double* hist_mm_h, ComStats* stats_h);
/* Stage 2 from (global) millimetre histograms in one launch - what n_channels x n_slices Emissivity runs would return
 * as total_emissivity (reader.py:199-401): per channel the Reff cut at min(largest non-empty bin / 1000,
 * profile_reff_max) (:210-217), bins -> nearest profile rows (:220-238), rows -> emissivities (:282-378), flux sums
 * (:380-401).  profiles (n_slices, K, 3) [Reff, T_e, n_e] share their Reff column; flux_out[(c * n_slices + s) *
 * flux_stride + t], t < transitions(c) + 1 (EXCIT.., then TOTAL); boundary_out (n_channels) nullable. */
int com_histograms_flux(int32_t n_channels, const ComTables* const* tables, const double* profiles, int32_t n_slices,
                        int32_t K, int32_t profile_sorted, double profile_reff_max, double concentration,
                        const double* hist_mm, double* flux_out, int32_t flux_stride, double* boundary_out,
                        com_stream_t stream);
int com_histograms_flux_host(int32_t n_channels, const ComTables* const* tables, const double* profiles_h,
                             int32_t n_slices, int32_t K, double concentration, const double* hist_mm_h,
                             double* flux_out_h, int32_t flux_stride, double* boundary_out_h);
/* The path's single exchange FUSED into Stage 2 (N > 1 ranks on one NVLink/NVSwitch box; no reference counterpart - the
 * reference is one process): instead of all-reducing the histograms and then calling com_histograms_flux, every rank
 * keeps its packed buffer [n_channels x 4096 bins | n_channels x n_counters], fp64, in memory its peers can read
 * (CUDA peer / symmetric memory) and the Stage-2 kernel itself loads the n_peers buffers over NVLink, sums them in rank
 * order (bit-identical sums on every rank) in shared memory and carries on as com_histograms_flux does.
 * peers_h: HOST array of n_peers DEVICE pointers (rank r's buffer as mapped into this process; entry `own rank` is the
 * local one).  hist_sum_out (n_channels, 4096) / counters_sum_out (n_channels, n_counters): nullable, the global sums
 * (hist_sum_out is required for more than 4 slices: the peers are then summed once and the sweep runs from the copy).
 * The CALLER orders the ranks: every rank's buffer must be complete and visible before this launch reads it, and must
 * not be overwritten until every peer's launch has finished (a device-side barrier on the stream before the call and
 * double buffering - see co_monitor_b200/volume.py). */
int com_histograms_flux_peers(int32_t n_channels, const ComTables* const* tables, const double* profiles,
                              int32_t n_slices, int32_t K, int32_t profile_sorted, double profile_reff_max,
                              double concentration, int32_t n_peers, const double* const* peers_h, int32_t n_counters,
                              double* hist_sum_out, double* counters_sum_out, double* flux_out, int32_t flux_stride,
                              double* boundary_out, com_stream_t stream);

/* ---- Reff provider ------------------------------------------------------------
 * Replaces the VMEC web-service step (co_monitor/geometry/reff_VMEC.py:85-136) and its 3-decimal file
 * (reff_VMEC.py:138-176) for lattices that have no shipped file: Reff of the voxels [vox_begin, vox_end) of the
 * (sharded) target lattice, as uint16 millimetre bins reff_mm_out[v - vox_begin] = rint(Reff * 1000), 0xFFFF outside
 * the last closed flux surface, and optionally (reff_out, nullable) as the 3-decimal doubles the file would hold.
 *   com_reff_resample   trilinear resampling of a coarser grid src_reff (DEVICE fp64, (src_ny, src_nx, src_nz) row-major
 *                       = the reference's flat order; NaN outside the LCFS) spanning the same cuboid; a stencil that
 *                       touches NaN with non-zero weight gives NaN.  Bit-identical to
 *                       co_monitor_b200.geometry.reff.interpolate_reff.
 *   com_reff_quadratic  analytic nested flux surfaces: Reff^2 = c0 x^2 + c1 y^2 + c2 z^2 + c3 xy + c4 xz + c5 yz +
 *                       c6 x + c7 y + c8 z + c9 (position in metres; coef_h HOST, 10 doubles), NaN beyond reff_max. */
int com_reff_resample(const ComLattice* target_h, int64_t vox_begin, int64_t vox_end, const double* src_reff,
                      int64_t src_nx, int64_t src_ny, int64_t src_nz, uint16_t* reff_mm_out, double* reff_out,
                      com_stream_t stream);
int com_reff_quadratic(const ComLattice* target_h, int64_t vox_begin, int64_t vox_end, const double* coef_h,
                       double reff_max, uint16_t* reff_mm_out, double* reff_out, com_stream_t stream);

/* ---- Stage-1 result file -------------------------------------------------------
 * The reference writes its result with to_csv(sep=";", header=True, index=False) (simulation.py:616-630):
 *   idx_sel_plas_points;plasma_x;plasma_y;plasma_z;total_intensity_fraction
 * integers in decimal, floats in Python's shortest round-trip repr.  com_write_stage1_csv writes the same bytes from
 * HOST arrays (n rows), formatting on n_threads host threads (0: all cores); 1e7-1e8 rows in seconds.
 * com_format_float_repr formats one value (returns its length; out needs >= 32 bytes) - the rule the writer uses. */
int com_write_stage1_csv(const char* path, const int64_t* idx, const double* x, const double* y, const double* z,
                         const double* w, int64_t n, int32_t write_header, int32_t n_threads);
int com_format_float_repr(double value, char* out, int32_t out_bytes);

/* Measurement aid (no counterpart in the reference; not thread-safe, one stream at a time): when enabled, every com_visibility_weights call
 * records CUDA events around its two kernels on the launching stream; com_kernel_timing_read waits for the
 * last call and returns the accumulated milliseconds of the FP32 filter kernel and of the fp64 kernels (weight-only + exact)
 * and the number of calls since the last enable. */
int com_kernel_timing_enable(int on);
int com_kernel_timing_read(double* filter_ms, double* exact_ms, int64_t* launches);
/* the same with the fp64 stage split: ms3 = {filter kernel, weight-only kernel, exact kernel} */
int com_kernel_timing_read3(double* ms3, int64_t* launches);

/* FP32 FMA throughput of the current device measured by a register-only probe kernel, TFLOP/s (FMA = 2);
 * the denominator bench.py reports next to the nominal 148 x 128 x 2 x f_max.  < 0 on error. */
double com_measure_fp32_tflops(void);

/* Error text of the last failure on the calling thread / of a status code.  The reference raises Python exceptions and
 * calls sys.exit("Not enough points to perform calculations!") (simulation.py:451-456): COM_E_EMPTY carries that text. */
const char* com_last_error(void);
const char* com_error_string(int code);
int com_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* COMONITOR_SM100_H */
