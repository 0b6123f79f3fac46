// Internal layout of a channel's device-resident tables (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <vector>

#include "../../include/comonitor_sm100.h"

namespace com {

// Bounding rectangle of a slit polygon in its aperture frame, band included.
struct SlitRect {
  float xlo, xhi, ylo, yhi;
};

// Compact fp64 aperture for the exact kernel: plane, in-plane basis, and a slice of the edge pool.
struct DevAperture {
  double p1[3];
  double normal[3];  // raw (p2-p1) x (p3-p2)
  double e1[3];
  double e2[3];
  int n_edges;
  int edge_offset;  // first (a,b,c) triple in DeviceScene::edges
};

// Everything the kernels need, passed by value as a __grid_constant__ parameter.
struct DeviceScene {
  int n_crystal;
  int n_crystal_padded;  // multiple of 32
  int n_slits;
  int n_apertures;       // 2S + 2: crys_0, plasma_0, crys_1, ..., port, detector
  int n_edges_total;
  int filter_disabled;   // 1: every ray goes to the fp64 stage
  int no_run_cull;       // 1: the lattice kernel tests every ray individually (measurement aid)
  int slit_filter;       // 1: the collimator is periodic and stage B of the filter tests it
  int slit_same_w;       // 1: both collimator planes share their normal, so one W form serves both sides
  float slit_period;     // slit pitch along e1 (2 |vector_top_bottom|)
  float slit_inv_period;
  SlitRect rect_crys, rect_plasma;
  // "certain pass" regions of the FP32 filter: rectangles inside the slit polygons whose every point is at least
  // certain_mm inside.  A ray the filter finds in them on both sides of one slit skips the fp64 slit tests of the
  // exact kernel (it could not fail them); xlo > xhi switches the shortcut off.
  SlitRect inner_crys, inner_plasma;
  float certain_mm;
  double slit_period_d;  // 0 when the collimator is not periodic (exact kernel then tries every slit)
  double slit_xlo_d, slit_xhi_d;  // extent along e1 of the crystal-side polygon of slit 0 (which neighbours a hit can be in)
  double origin[3];
  double area_pt, refl_max, aoi0, refl_sigma;
  double near_edge_mm;
  double near_band_mm;           // second near-edge counter (>= near_edge_mm)
  int physics_flags;             // COM_PHYS_* (extensions; 0 on the parity path)
  int n_shields;
  ComShield shields[COM_MAX_SHIELDS];  // circular stops (extension), by value: a few hundred bytes
  const double* crystal_xyz;     // (C,3) global
  const double* crystal_normal;  // (C,3)
  const DevAperture* apertures;  // (2S+2)
  const double* edges;           // (n_edges_total,3)
  // FP32 filter tables: linear forms in the recentred voxel position p' = p - origin, value = g.p' + h
  const float4* det_forms;       // [3][Cp]: Xs, Ys, W of the mirrored detector; passes iff max(|Xs|,|Ys|) <= |W|
  // optional detector image (extension, off unless com_scene_set_detector_image was called): weight of every passing
  // ray accumulated into the pixel its detector hit falls in, pixels of pix_mm in the detector aperture frame (e1, e2)
  double* pix_out;
  int pix_nx, pix_ny;
  double pix_mm, pix_x0, pix_y0;
  const float4* slit_forms;      // [Cp][6]: X', Y', W on the crystal-side slit plane, then on the plasma side
  // Column-interval filter (lattice kernel): along a z-column every aperture test is a linear inequality in the voxel
  // index, so the passing voxels of a (column, crystal point) pair form a few intervals.
  int interval_filter;           // 1: the geometry allows it (periodic collimator, disjoint slit rectangles, ...)
  int certain_off;               // 1: no ray is vouched for (extensions that add fp64 tests of their own)
  SlitRect det_inner;            // inner rectangle of the detector polygon in the SCALED frame of det_forms (|x|,|y| <= 1
                                 // is the grown bounding rectangle); xlo > xhi: no certain region
  SlitRect port_outer, port_inner;  // bounding / inner rectangle of the port polygon in its aperture frame (mm)
  // edges of slit 0's polygon on the crystal side [0] and on the plasma side [1]: a x + b y + c >= 0 inside, (a, b) unit;
  // .z = c + eps (outer: candidates), .w = c - certain_mm (inner: certain).  Slit j: x -> x - j * period.
  float4 slit_edge[2][8];
  int n_slit_edge[2];
  // edges of the port polygon in its aperture frame, same layout (.w = c - certain_mm: a point inside all of them is at
  // least certain_mm inside the polygon).  n_port_edge = 0 (more than 16 edges): the certain test falls back to the
  // inner rectangle port_inner, which vouches for fewer rays.
  float4 port_edge[16];
  int n_port_edge;
  // edges of the detector polygon in the SCALED frame of det_forms: (A, B, C + eps, C - certain_mm) with
  // A Xs + B Ys + C W >= 0 (times the sign of W) inside.  A skewed detector section loses a tenth of its area to the inner
  // rectangle det_inner (channel O); the edges lose nothing.  n_det_edge = 0: det_inner is used.
  float4 det_edge[8];
  int n_det_edge;
  const float4* port_forms;      // [Cp][3]: X', Y', W on the port plane
  // reflectivity profile at every value the rounded angle can take: refl_table[m] = refl_max * exp(-(aoi - aoi0)^2 /
  // (2 sigma^2)) with aoi = (180 - m / 100) / 2, m = rint(deg * 100) in 0..18000 (simulation.py:520-531, 564-577)
  const double* refl_table;
};

// Host-side result of analysing a ComSceneDesc (pure CPU, no CUDA).
struct HostScene {
  DeviceScene dev{};                  // pointers unset
  std::vector<ComAperture> apertures; // 2S + 2
  std::vector<DevAperture> dev_apertures;
  std::vector<double> edges;
  std::vector<float> det_forms;       // 4 * 3 * Cp
  std::vector<float> slit_forms;      // 4 * 6 * Cp
  std::vector<float> port_forms;      // 4 * 3 * Cp
  std::vector<double> refl_table;     // 18001
  std::vector<double> crystal_xyz, crystal_normal;
};

// Builds apertures + filter tables.  Returns COM_OK or COM_E_*.
int build_host_scene(const ComSceneDesc& d, HostScene& out);

void set_error(const char* fmt, ...);

// Grow-only device block + pinned staging block kept per host thread by the synchronous *_host entry points, so that a
// reference-facing call does not pay cudaMalloc/cudaFree (both synchronise the device) every time.  Each thread
// works on its own per-thread CUDA stream and its own blocks, so calls from different threads overlap on the device
// and need no lock; the blocks are freed when the thread exits (or by com_host_arena_release).
// reserve() returns COM_OK or COM_E_*.
struct HostArena {
  static int reserve(size_t device_bytes, size_t pinned_bytes, char** device_out, char** pinned_out);
  static void release_all();  // frees the calling thread's blocks (com_host_arena_release)
};

}  // namespace com

struct ComScene {
  com::DeviceScene dev{};
  std::vector<ComAperture> apertures_h;
  void* d_blob = nullptr;  // single device allocation holding all tables
  double filter_eps_mm = 0;
  int device = -1;  // device holding d_blob (-1: tables live in a caller-managed arena)
};
