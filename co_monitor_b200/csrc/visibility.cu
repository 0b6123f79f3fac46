// Stage 1 on sm_100a: visibility + geometric weight of every (voxel, crystal point) ray.
//
//   K1a  filter_kernel   FP32, all P*C rays.  One warp = 32 crystal points (their detector-cone faces live
//                        in registers), looping over a tile of voxels broadcast from shared memory.  A ray
//                        survives stage A when the voxel lies inside the crystal point's mirrored-detector
//                        cone (3 FFMA per cone face, nothing else); survivors are compacted through a per-warp
//                        shared-memory queue (ballot + popc) and, 32 at a time, pushed through the periodic
//                        collimator (both sides) and the port polygon (stage B).  Every comparison carries a
//                        band eps, so the filter can only over-accept.  Survivors go to a global candidate list.
//   K1b  exact_kernel    fp64, candidates only (~0.5 % of the rays).  Restates the reference arithmetic ray by
//                        ray (simulation.py:143-173, 273-325, 344-428, 483-536, 564-577): per-slit plane hits,
//                        exact polygon tests, distance, AOI with round(deg, 2), weight; accumulates per voxel.
//
// The visibility mask is therefore decided in fp64 for every ray the reference could accept; the FP32 pass
// only discards rays that miss an aperture by more than eps.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "scene.h"

namespace com {

constexpr int kTileVoxels = 512;
constexpr int kMaxWarps = 16;
constexpr int kUnroll = 4;
constexpr int kQueueCap = 32 * kUnroll + 32;

struct VoxelSource {
  int use_lattice;
  ComLattice lat;
  const double* xyz;
};

struct K1Params {
  DeviceScene sc;
  VoxelSource vox;
  long long vox_begin;
  long long n_vox;
  int n_tiles;
  int n_super;       // crystal super-groups (warps_per_cta * 32 points each)
  int warps_per_cta;
  uint2* cand;       // (voxel_local, crystal) pairs
  unsigned long long cand_capacity;
  unsigned long long* counters;  // [0] candidates, [1] next work item
  double* w_out;
  unsigned int* nrays_out;
  ComRayRecord* rays_out;
  long long rays_capacity;
  unsigned long long* rays_count;
  const uint16_t* reff_bin;
  double* hist_out;
  int n_bins;
  ComStats* stats;
};

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 ld3(const double* p) { return {p[0], p[1], p[2]}; }

// CuboidMesh closed form (mesh_calculation.py:128-155); mul/add kept separate like numpy.linspace.
__device__ __forceinline__ D3 voxel_position(const VoxelSource& v, long long flat) {
  if (!v.use_lattice) return ld3(v.xyz + 3 * flat);
  const ComLattice& L = v.lat;
  long long iz = flat % L.nz, t = flat / L.nz;
  long long ix = t % L.nx, iy = t / L.nx;
  double x = ix == L.nx - 1 ? L.stop[0] : __dadd_rn(__dmul_rn((double)ix, L.step[0]), L.start[0]);
  double y = iy == L.ny - 1 ? L.stop[1] : __dadd_rn(__dmul_rn((double)iy, L.step[1]), L.start[1]);
  double z = iz == L.nz - 1 ? L.stop[2] : __dadd_rn(__dmul_rn((double)iz, L.step[2]), L.start[2]);
  double mx = x - L.centre[0], my = y - L.centre[1], mz = z - L.centre[2];
  D3 q;
  q.x = mx * L.rotation[0] + my * L.rotation[3] + mz * L.rotation[6];
  q.y = mx * L.rotation[1] + my * L.rotation[4] + mz * L.rotation[7];
  q.z = mx * L.rotation[2] + my * L.rotation[5] + mz * L.rotation[8];
  return {__dadd_rn(__dadd_rn(q.x, L.centre[0]), L.shift[0]), __dadd_rn(__dadd_rn(q.y, L.centre[1]), L.shift[1]),
          __dadd_rn(__dadd_rn(q.z, L.centre[2]), L.shift[2])};
}

// ------------------------------------------------------------------------------------------------------
// K1a stage B: collimator (periodic, both sides) and port, FP32 with band.
struct F3 {
  float x, y, z;
};
__device__ __forceinline__ float dotf(const float* a, F3 b) { return fmaf(a[0], b.x, fmaf(a[1], b.y, a[2] * b.z)); }

// Homogeneous in-plane hit of the line c + s*u with the plane: x = X/W, y = Y/W.
__device__ __forceinline__ void plane_hit_h(const FilterPlane& pl, F3 c, F3 u, float& X, float& Y, float& W) {
  F3 r = {c.x - pl.q0[0], c.y - pl.q0[1], c.z - pl.q0[2]};
  float nd = dotf(pl.n, u), a = -dotf(pl.n, r);
  X = fmaf(dotf(pl.e1, r), nd, a * dotf(pl.e1, u));
  Y = fmaf(dotf(pl.e2, r), nd, a * dotf(pl.e2, u));
  W = nd;
}

__device__ __forceinline__ bool poly_pass(const FilterPlane& pl, float x, float y) {
  float m = 1e30f;
  for (int k = 0; k < pl.n_edges; ++k) m = fminf(m, fmaf(pl.edge[k][0], x, fmaf(pl.edge[k][1], y, pl.edge[k][2])));
  return m >= -pl.eps;
}

__device__ __forceinline__ bool stage_b(const DeviceScene& sc, F3 p, F3 c) {
  F3 u = {p.x - c.x, p.y - c.y, p.z - c.z};
  float X, Y, W;
  // port: homogeneous test, no division
  plane_hit_h(sc.port, c, u, X, Y, W);
  {
    float s = W < 0.f ? -1.f : 1.f, aw = fabsf(W), m = 1e30f;
    for (int k = 0; k < sc.port.n_edges; ++k) {
      float e = fmaf(sc.port.edge[k][0], X, fmaf(sc.port.edge[k][1], Y, sc.port.edge[k][2] * W));
      m = fminf(m, s * e);
    }
    if (!(m >= -sc.port.eps * aw) && aw > 1e-3f) return false;
  }
  if (sc.slit_crys.n_edges == 0) return true;  // non-periodic collimator: decided in fp64 only
  float Xp, Yp, Wp;
  plane_hit_h(sc.slit_crys, c, u, X, Y, W);
  plane_hit_h(sc.slit_plasma, c, u, Xp, Yp, Wp);
  if (fabsf(W) < 1e-3f || fabsf(Wp) < 1e-3f) return true;  // grazing: let fp64 decide
  float x = X / W, y = Y / W, xp = Xp / Wp, yp = Yp / Wp;
  float kf = floorf(x * sc.slit_inv_period);
  bool ok = false;
#pragma unroll
  for (int dk = 0; dk < 2; ++dk) {
    float k = kf + (float)dk;
    if (k >= 0.f && k < (float)sc.n_slits) {
      float sh = k * sc.slit_period;
      ok |= poly_pass(sc.slit_crys, x - sh, y) && poly_pass(sc.slit_plasma, xp - sh, yp);
    }
  }
  return ok;
}

// ------------------------------------------------------------------------------------------------------
template <int NE, bool SINGLE>
__global__ void __launch_bounds__(kMaxWarps * 32) filter_kernel(const __grid_constant__ K1Params P) {
  __shared__ float4 s_tile[kTileVoxels];
  __shared__ unsigned int s_queue[kMaxWarps][kQueueCap];
  __shared__ int s_item;

  const DeviceScene& sc = P.sc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Cp = sc.n_crystal_padded;
  const float eps = sc.cone_eps;
  const bool all_pass = sc.filter_disabled != 0;
  const int n_items = P.n_tiles * P.n_super;
  unsigned int* queue = s_queue[warp];

  int cur_super = -1;
  float4 g[NE];
  int ci = 0;
  bool warp_active = false;

  for (;;) {
    __syncthreads();  // previous tile fully consumed, s_item free
    if (threadIdx.x == 0) s_item = (int)atomicAdd(&P.counters[1], 1ull);
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int tile = item / P.n_super, sup = item - tile * P.n_super;
    const long long tile_base = (long long)tile * kTileVoxels;
    const int nv = (int)min((long long)kTileVoxels, P.n_vox - tile_base);

    // stage the tile: fp64 position -> recentred FP32
    for (int i = threadIdx.x; i < nv; i += blockDim.x) {
      D3 p = voxel_position(P.vox, P.vox_begin + tile_base + i);
      s_tile[i] = make_float4((float)(p.x - sc.origin[0]), (float)(p.y - sc.origin[1]), (float)(p.z - sc.origin[2]), 0.f);
    }
    if (sup != cur_super) {
      cur_super = sup;
      ci = (sup * P.warps_per_cta + warp) * 32 + lane;
      warp_active = (sup * P.warps_per_cta + warp) * 32 < Cp;
      if (warp_active) {
#pragma unroll
        for (int j = 0; j < NE; ++j) g[j] = __ldg(&sc.cone_forms[(size_t)j * Cp + ci]);
      }
    }
    __syncthreads();
    if (!warp_active) continue;

    const float4 cf = __ldg(&sc.crystal_f[ci]);
    const F3 c = {cf.x, cf.y, cf.z};
    int qn = 0;  // warp-uniform queue fill

    auto drain = [&](int take) {  // stage B on `take` (<= 32) queued rays, then append survivors
      __syncwarp();
      const unsigned int ent = lane < take ? queue[qn - take + lane] : 0u;
      const int src = ent & 31;
      F3 cc;  // crystal point of the queued ray lives in lane `src` of this warp
      cc.x = __shfl_sync(0xffffffffu, c.x, src);
      cc.y = __shfl_sync(0xffffffffu, c.y, src);
      cc.z = __shfl_sync(0xffffffffu, c.z, src);
      bool keep = false;
      if (lane < take) {
        const float4 pv = s_tile[ent >> 5];
        keep = all_pass || stage_b(sc, F3{pv.x, pv.y, pv.z}, cc);
      }
      qn -= take;
      const unsigned int b = __ballot_sync(0xffffffffu, keep);
      if (b) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&P.counters[0], (unsigned long long)__popc(b));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (keep) {
          const unsigned long long pos = base + __popc(b & ((1u << lane) - 1u));
          if (pos < P.cand_capacity)
            P.cand[pos] = make_uint2((unsigned int)(tile_base + (ent >> 5)), (unsigned int)(ci - lane + src));
        }
      }
      __syncwarp();
    };

    for (int v0 = 0; v0 < nv; v0 += kUnroll) {
#pragma unroll
      for (int i = 0; i < kUnroll; ++i) {
        const int v = v0 + i;
        const float4 p = s_tile[v < nv ? v : nv - 1];
        float m = 1e30f, M = -1e30f;
#pragma unroll
        for (int j = 0; j < NE; ++j) {
          const float e = fmaf(g[j].x, p.x, fmaf(g[j].y, p.y, fmaf(g[j].z, p.z, g[j].w)));
          m = fminf(m, e);
          if (!SINGLE) M = fmaxf(M, e);
        }
        bool pass = m >= -eps;
        if (!SINGLE) pass = pass || (M <= eps);
        pass = (pass || all_pass) && v < nv && ci < sc.n_crystal;
        const unsigned int b = __ballot_sync(0xffffffffu, pass);
        if (b) {
          if (pass) queue[qn + __popc(b & ((1u << lane) - 1u))] = ((unsigned int)v << 5) | (unsigned int)lane;
          qn += __popc(b);
        }
      }
      while (qn >= 32) drain(32);
    }
    if (qn > 0) drain(qn);
  }
}

// ------------------------------------------------------------------------------------------------------
// K1b: the reference arithmetic, one candidate ray per thread.
struct HitResult {
  bool pass;
  double margin;
};

// simulation.py:143-173 (+175-207): line through `dst` with direction src - dst against the aperture plane,
// then the polygon test that replaces Delaunay.find_simplex (geometry_host.cu).
__device__ __forceinline__ HitResult aperture_hit(const ComAperture& ap, D3 src, D3 dst) {
  D3 N = ld3(ap.normal), u = src - dst;
  double ndotu = dot(N, u);
  if (fabs(ndotu) < 1e-6) return {false, -1e300};  // the reference writes NaN -> never inside
  D3 w = dst - ld3(ap.p1);
  double si = -(dot(N, w) / ndotu);
  D3 r = w + si * u;  // X - p1
  double x = dot(r, ld3(ap.e1)), y = dot(r, ld3(ap.e2));
  double m = 1e300;
  for (int k = 0; k < ap.n_edges; ++k) m = fmin(m, ap.edge[k][0] * x + ap.edge[k][1] * y + ap.edge[k][2]);
  return {m >= 0.0, m};
}

__global__ void __launch_bounds__(128) exact_kernel(const __grid_constant__ K1Params P) {
  const DeviceScene& sc = P.sc;
  const unsigned long long n = min(P.counters[0], P.cand_capacity);
  const int S = sc.n_slits;
  unsigned int n_pass = 0, n_edge = 0, n_vis = 0;
  const double kNear = sc.near_edge_mm;

  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 e = P.cand[i];
    const long long vloc = e.x;
    const int ci = (int)e.y;
    const D3 p = voxel_position(P.vox, P.vox_begin + vloc);
    const D3 c = ld3(sc.crystal_xyz + 3 * ci), nrm = ld3(sc.crystal_normal + 3 * ci);
    // simulation.py:309-321
    const D3 d = c - p;
    const D3 r = d - (2.0 * dot(d, nrm)) * nrm;
    const D3 reflected = c + r, plasma = c - d, crystal = p + d;

    bool near_edge = false;
    HitResult h = aperture_hit(sc.apertures[2 * S + 1], reflected, crystal);  // detector :387-400
    near_edge |= fabs(h.margin) < kNear;
    bool ok = h.pass;
    if (ok) {
      h = aperture_hit(sc.apertures[2 * S], plasma, crystal);  // port :372-384
      near_edge |= fabs(h.margin) < kNear;
      ok = h.pass;
    }
    if (ok) {
      ok = false;
      for (int k = 0; k < S && !ok; ++k) {  // :344-369, same slit on both sides
        HitResult a = aperture_hit(sc.apertures[2 * k], plasma, crystal);
        near_edge |= fabs(a.margin) < kNear;
        if (!a.pass) continue;
        HitResult b = aperture_hit(sc.apertures[2 * k + 1], plasma, crystal);
        near_edge |= fabs(b.margin) < kNear;
        ok = b.pass;
      }
    }
    n_edge += near_edge ? 1u : 0u;
    if (!ok) continue;

    // simulation.py:494-497, 516-531, 564-577
    const D3 v = plasma - crystal, rc = reflected - crystal;
    const double vv = dot(v, v), dist = sqrt(vv);
    const double cosang = dot(v, rc) / (sqrt(vv) * sqrt(dot(rc, rc)));
    const double deg = acos(cosang) * (180.0 / 3.141592653589793);
    const double aoi = (180.0 - rint(deg * 100.0) / 100.0) / 2.0;
    const double fraction = sc.area_pt / (4.0 * 3.141592653589793 * (dist * dist));
    const double da = aoi - sc.aoi0;
    const double prof = sc.refl_max * exp(-(da * da) / (2.0 * (sc.refl_sigma * sc.refl_sigma)));
    const double w = fraction * prof;

    ++n_pass;
    atomicAdd(&P.w_out[vloc], w);
    if (P.nrays_out) {
      if (atomicAdd(&P.nrays_out[vloc], 1u) == 0u) ++n_vis;
    }
    if (P.hist_out) {
      const int bin = P.reff_bin[P.vox_begin + vloc];
      if (bin < P.n_bins) atomicAdd(&P.hist_out[bin], w);
    }
    if (P.rays_out) {
      const unsigned long long pos = atomicAdd(P.rays_count, 1ull);
      if ((long long)pos < P.rays_capacity)
        P.rays_out[pos] = ComRayRecord{(P.vox_begin + vloc) * (long long)sc.n_crystal + ci, dist, aoi, w};
    }
  }

  // block-level stats
  __shared__ unsigned int s_acc[3];
  if (threadIdx.x < 3) s_acc[threadIdx.x] = 0;
  __syncthreads();
  for (int off = 16; off; off >>= 1) {
    n_pass += __shfl_down_sync(0xffffffffu, n_pass, off);
    n_edge += __shfl_down_sync(0xffffffffu, n_edge, off);
    n_vis += __shfl_down_sync(0xffffffffu, n_vis, off);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_acc[0], n_pass);
    atomicAdd(&s_acc[1], n_edge);
    atomicAdd(&s_acc[2], n_vis);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd((unsigned long long*)&P.stats->passing_rays, (unsigned long long)s_acc[0]);
    atomicAdd((unsigned long long*)&P.stats->near_edge_rays, (unsigned long long)s_acc[1]);
    atomicAdd((unsigned long long*)&P.stats->visible_voxels, (unsigned long long)s_acc[2]);
    if (blockIdx.x == 0) {
      const unsigned long long total = P.counters[0];
      P.stats->candidates = total;
      P.stats->candidate_capacity = P.cand_capacity;
      P.stats->overflow = total > P.cand_capacity ? 1 : 0;
      P.stats->tests = (unsigned long long)P.n_vox * (unsigned long long)sc.n_crystal;
    }
  }
}

// ------------------------------------------------------------------------------------------------------
static int pick_warps(int n_groups) {  // warps per CTA wasting the fewest idle warps; ties -> larger CTA
  if (n_groups <= kMaxWarps) return n_groups;
  int best = kMaxWarps, best_waste = 1 << 30;
  for (int w = 6; w <= kMaxWarps; ++w) {
    int waste = (n_groups + w - 1) / w * w - n_groups;
    if (waste <= best_waste) best = w, best_waste = waste;
  }
  return best;
}

template <int NE>
static cudaError_t launch_filter(const K1Params& P, int grid, int threads, cudaStream_t s) {
  if (P.sc.single_nappe)
    filter_kernel<NE, true><<<grid, threads, 0, s>>>(P);
  else
    filter_kernel<NE, false><<<grid, threads, 0, s>>>(P);
  return cudaGetLastError();
}

template <int NE>
static int filter_occupancy(bool single, int threads) {
  int n = 0;
  if (single)
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, filter_kernel<NE, true>, threads, 0);
  else
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, filter_kernel<NE, false>, threads, 0);
  return n;
}

size_t visibility_workspace_bytes(const DeviceScene& sc, long long n_vox, double frac) {
  if (sc.filter_disabled) frac = 1.0;
  if (frac <= 0) frac = 0.02;
  double rays = (double)n_vox * sc.n_crystal;
  unsigned long long cap = (unsigned long long)(rays * frac) + 1;
  if (!sc.filter_disabled) cap = std::max(cap, 1ull << 20);
  return 256 + cap * sizeof(uint2);
}

int visibility_launch(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz, long long vox_begin,
                      long long vox_end, double* w_out, unsigned int* nrays_out, ComRayRecord* rays_out,
                      long long rays_capacity, unsigned long long* rays_count_out, const uint16_t* reff_bin,
                      double* hist_out, int n_bins, ComStats* stats_out, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  if (!scene || (!lattice_h && !vox_xyz) || vox_end < vox_begin || !w_out || !stats_out || !workspace ||
      (rays_out && !rays_count_out) || (hist_out && !reff_bin)) {
    set_error("com_visibility_weights: bad arguments");
    return COM_E_BADARG;
  }
  const long long n_vox = vox_end - vox_begin;
  if (n_vox >= (1ll << 32)) {
    set_error("com_visibility_weights: at most 2^32-1 voxels per call (got %lld); split the range", n_vox);
    return COM_E_BADARG;
  }
  if (workspace_bytes < 256 + sizeof(uint2)) {
    set_error("workspace too small");
    return COM_E_NOMEM;
  }
  K1Params P{};
  P.sc = scene->dev;
  P.vox.use_lattice = lattice_h ? 1 : 0;
  if (lattice_h) {
    P.vox.lat = *lattice_h;
    if (vox_end > lattice_h->nx * lattice_h->ny * lattice_h->nz || vox_begin < 0) {
      set_error("voxel range [%lld,%lld) outside the lattice", vox_begin, vox_end);
      return COM_E_BADARG;
    }
  }
  P.vox.xyz = vox_xyz;
  P.vox_begin = vox_begin;
  P.n_vox = n_vox;
  P.n_tiles = (int)((n_vox + kTileVoxels - 1) / kTileVoxels);
  const int n_groups = P.sc.n_crystal_padded / 32;
  P.warps_per_cta = pick_warps(n_groups);
  P.n_super = (n_groups + P.warps_per_cta - 1) / P.warps_per_cta;
  P.counters = reinterpret_cast<unsigned long long*>(workspace);
  P.cand = reinterpret_cast<uint2*>(static_cast<char*>(workspace) + 256);
  P.cand_capacity = (workspace_bytes - 256) / sizeof(uint2);
  P.w_out = w_out;
  P.nrays_out = nrays_out;
  P.rays_out = rays_out;
  P.rays_capacity = rays_capacity;
  P.rays_count = rays_count_out;
  P.reff_bin = reff_bin;
  P.hist_out = hist_out;
  P.n_bins = n_bins;
  P.stats = stats_out;

  cudaError_t e;
#define COM_CUDA(x)                                                        \
  if ((e = (x)) != cudaSuccess) {                                          \
    set_error("%s failed: %s", #x, cudaGetErrorString(e));                 \
    return COM_E_CUDA;                                                     \
  }
  COM_CUDA(cudaMemsetAsync(workspace, 0, 256, stream));
  COM_CUDA(cudaMemsetAsync(stats_out, 0, sizeof(ComStats), stream));
  COM_CUDA(cudaMemsetAsync(w_out, 0, sizeof(double) * (size_t)n_vox, stream));
  if (nrays_out) COM_CUDA(cudaMemsetAsync(nrays_out, 0, sizeof(unsigned int) * (size_t)n_vox, stream));
  if (rays_count_out) COM_CUDA(cudaMemsetAsync(rays_count_out, 0, sizeof(unsigned long long), stream));
  if (n_vox == 0) return COM_OK;

  int dev = 0, sms = 0;
  COM_CUDA(cudaGetDevice(&dev));
  COM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int threads = P.warps_per_cta * 32;
  const int ne = P.sc.cone_edges;
  const bool single = P.sc.single_nappe != 0;
  int occ = ne <= 4 ? filter_occupancy<4>(single, threads)
            : ne <= 6 ? filter_occupancy<6>(single, threads)
                      : filter_occupancy<8>(single, threads);
  if (occ < 1) occ = 1;
  const long long items = (long long)P.n_tiles * P.n_super;
  const int grid = (int)std::min<long long>(items, (long long)sms * occ);
  // cone_forms is padded to the template width by scene creation (see capi.cu)
  e = ne <= 4 ? launch_filter<4>(P, grid, threads, stream)
      : ne <= 6 ? launch_filter<6>(P, grid, threads, stream)
                : launch_filter<8>(P, grid, threads, stream);
  if (e != cudaSuccess) {
    set_error("filter_kernel launch failed: %s", cudaGetErrorString(e));
    return COM_E_CUDA;
  }
  exact_kernel<<<sms * 8, 128, 0, stream>>>(P);
  COM_CUDA(cudaGetLastError());
#undef COM_CUDA
  return COM_OK;
}

}  // namespace com
