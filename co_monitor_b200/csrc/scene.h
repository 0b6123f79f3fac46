// Internal layout of a channel's device-resident tables (not part of the C ABI).
#pragma once
#include <cstdint>
#include <vector>

#include "../../include/comonitor_sm100.h"

namespace com {

constexpr int kMaxFilterEdges = 32;   // FP32 polygon of one filter plane (port is the largest: 14)
constexpr int kMaxConeEdges = 8;      // stage-A detector cone faces per crystal point

// One aperture plane for the FP32 filter, in the recentred frame (global - origin).
struct FilterPlane {
  float n[3];   // unit normal
  float q0[3];  // plane point (aperture p1)
  float e1[3];
  float e2[3];
  int n_edges;
  float eps;    // band in mm at this plane
  float edge[kMaxFilterEdges][3];  // a*x + b*y + c >= -eps passes; (a,b) unit
};

// Everything the kernels need, passed by value as a __grid_constant__ parameter.
struct DeviceScene {
  int n_crystal;
  int n_crystal_padded;  // multiple of 32
  int n_slits;
  int n_apertures;       // 2S + 2
  int cone_edges;        // faces per crystal point in cone_forms (<= kMaxConeEdges)
  int single_nappe;      // 1: only the forward nappe of the detector cone can be hit inside voxel_bounds
  int filter_disabled;   // 1: every ray goes to the fp64 stage
  float cone_eps;        // stage-A band, mm of distance from the cone face at the voxel
  float slit_period;     // 2 |vector_top_bottom| along e1
  float slit_inv_period;
  double origin[3];
  double area_pt, refl_max, aoi0, refl_sigma;
  double near_edge_mm;
  const double* crystal_xyz;     // (C,3) global
  const double* crystal_normal;  // (C,3)
  const ComAperture* apertures;  // (2S+2)
  const float4* cone_forms;      // [cone_edges][n_crystal_padded]: (gx,gy,gz,h), E = g.p' + h
  const float4* crystal_f;       // [n_crystal_padded] recentred crystal points, w unused
  FilterPlane slit_crys, slit_plasma, port;
};

// Host-side result of analysing a ComSceneDesc (pure CPU, no CUDA).
struct HostScene {
  DeviceScene dev{};                  // pointers unset
  std::vector<ComAperture> apertures; // 2S + 2
  std::vector<float> cone_forms;      // 4 * cone_edges * n_crystal_padded
  std::vector<float> crystal_f;       // 4 * n_crystal_padded
  std::vector<double> crystal_xyz, crystal_normal;
};

// Builds apertures + filter tables.  Returns COM_OK or COM_E_*.
int build_host_scene(const ComSceneDesc& d, HostScene& out);

void set_error(const char* fmt, ...);

}  // namespace com

struct ComScene {
  com::DeviceScene dev{};
  std::vector<ComAperture> apertures_h;
  void* d_blob = nullptr;  // single device allocation holding all tables
  int device = 0;
};
