// Stage 1 on sm_100a: visibility + geometric weight of every (voxel, crystal point) ray.
//
// FP32 front ends (find the candidate rays; can only over-accept - bounding shapes grown by eps = 5e-3 mm):
//   interval_kernel_lattice   the shipped filter for implicit lattices.  Along a z-column every aperture test is a linear
//                             inequality in the voxel index, so the passing voxels of a (column, crystal point) pair are a
//                             few intervals; one 16-byte record per interval leaves the kernel, with the part of it that
//                             is >= certain_mm inside every polygon (slit, detector and port edges) marked "certain".
//   expand_kernel             interval records -> per-ray list of the uncertain rays (block scan + binary search).
//   filter_kernel_lattice     every ray's detector test evaluated: forward differences along z-runs (3 FADD per ray), with
//                             or without the run-level cull; survivors compacted through a per-warp shared-memory queue
//                             and tested against the periodic collimator (stage B); certain rays classified on the
//                             compacted survivors.  `value` of bench.py, and geometries the interval form does not cover.
//   filter_kernel             explicit voxels: 512-voxel tiles in shared memory, moved there by cp.async.bulk from the
//                             packed float4 copy (voxel_pack_kernel) or converted from the fp64 rows.
// fp64 back ends (simulation.py:143-173, 273-325, 344-428, 483-536, 564-577):
//   weight_kernel_records     certain rays straight from the interval records: distance, angle, weight only.
//   weight_kernel             the same from the per-ray list (front region) of filter_kernel_lattice.
//   exact_kernel              every other candidate: the reference arithmetic ray by ray - plane hits, polygon tests for
//                             port, slit and detector, distance, AOI with round(deg, 2), weight; optional extensions.
//   count_visible_kernel      visible voxels = voxels with a ray count > 0.
//
// The visibility mask is therefore decided in fp64 for every ray within eps of any aperture edge; the FP32 pass only
// discards rays that miss an aperture by more than eps, and only vouches for rays that clear every edge by certain_mm.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>

#include "lattice.cuh"
#include "scene.h"

namespace com {

constexpr int kTileVoxels = 512;       // voxels per work item, explicit-voxel kernel
constexpr int kSegmentMax = 64;        // longest run of z-consecutive voxels advanced by forward differences
constexpr int kMaxWarps = 16;
constexpr int kUnroll = 8;
constexpr int kQueueCap = 32 * kUnroll + 32;

struct VoxelSource {
  int use_lattice;
  ComLattice lat;
  const double* xyz;   // explicit voxels: (P,3) fp64, global mm (what the reference materialises, mesh_calculation.py:119-159)
  const float4* f4;    // null, or the same voxels packed for the filter: (x - ox, y - oy, z - oz, 1) FP32, one 16-byte
                       // element per voxel (com_voxels_pack); tiles of it are moved into shared memory by bulk copies
};

struct K1Params {
  DeviceScene sc;
  VoxelSource vox;
  long long vox_begin;
  long long n_vox;
  int n_tiles;
  int n_super;       // crystal super-groups (warps_per_cta * 32 points each)
  int warps_per_cta;
  int tile_voxels;
  int segs_per_item;  // lattice kernel: segments per work item
  float dz[3];       // lattice kernel: recentred position step between z-consecutive voxels (FP32)
  uint2* cand;       // (voxel_local, crystal) pairs
  unsigned long long cand_capacity;
  unsigned long long* counters;  // [0] slots used at the front (certain), [1] next work item, [2] slots used at the back,
                                 // [5] interval records reserved, [6] rays in all interval records, [7] certain rays the
                                 // weight-only kernel took straight from the records
  double* w_out;
  unsigned int* nrays_out;
  ComRayRecord* rays_out;
  long long rays_capacity;
  unsigned long long* rays_count;
  const uint16_t* reff_bin;
  double* hist_out;
  int n_bins;
  unsigned long long magic_nz, magic_nx, magic_cpb;  // weight kernel: multipliers replacing divisions by nz, nx, cols per block
  uint4* recs;            // column-interval kernel: one 16-byte record per interval (layout at the kernel's emit)
  unsigned long long rec_capacity;
  int records_to_weight;  // 1: the weight-only kernel reads the certain parts straight from the records (expand_kernel then
                          // lists the uncertain rays only, counters[0] stays 0 and counters[7] counts the certain rays)
  double dz_d[3];         // position step between z-consecutive voxels, fp64
  int front_weight_only;  // 1: the front region of the list was produced by the column-interval kernel and is handled by
                          // weight_kernel (the exact kernel then takes the back region only)
  int accumulate_stats;  // 1: add this launch's counters to *stats instead of overwriting them (whole-volume runs)
  ComStats* stats;
};

// Voxel source: the implicit lattice (lattice.cuh) or explicit fp64 coordinates.
__device__ __forceinline__ D3 voxel_position(const VoxelSource& v, long long flat) {
  if (!v.use_lattice) return ld3(v.xyz + 3 * flat);
  return lattice_position(v.lat, flat);
}

// ------------------------------------------------------------------------------------------------------
// K1a stage B: periodic collimator, both sides, FP32.  x = X'/W, y = Y'/W are the in-plane hit coordinates
// relative to slit 0; slit k is slit 0 moved by k*T along x.  Bounding rectangles grown by eps: superset.
__device__ __forceinline__ float form(const float4 g, const float4 p) {
  return fmaf(g.x, p.x, fmaf(g.y, p.y, fmaf(g.z, p.z, g.w)));
}

__device__ __forceinline__ float rcp_approx(float x) {  // MUFU.RCP, ~1 ulp: ample for a test with a 5e-3 mm band
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

constexpr int kSlitUnknown = 15;  // candidate records carry the slit the FP32 filter saw in 4 bits; 15 = not known
constexpr int kSlitCertain = 0x100;  // stage_b: the ray is inside the inner rectangle of that slit on both sides
constexpr unsigned int kCandCertain = 1u << 27;  // candidate record: the slit tests cannot fail (the exact kernel skips them)
constexpr unsigned int kCandCertainAll = 1u << 26;  // ... nor can the port and detector tests: only the weight is left
constexpr unsigned int kCandCrystalMask = kCandCertainAll - 1u;

// Returns -1 when the ray misses every slit (band included), else the index of the slit it goes through on both
// sides, or kSlitUnknown when that is ambiguous / not evaluated.  The exact kernel tries the hinted slit first.
__device__ __forceinline__ bool ratio_inside(float s, float n, float w, float lo_b, float hi_b) {
  return s * fmaf(-lo_b, w, n) >= 0.f && s * fmaf(hi_b, w, -n) >= 0.f;
}

__device__ __forceinline__ int stage_b(const DeviceScene& sc, const float4 p, int cj) {
  if (!sc.slit_filter) return kSlitUnknown;  // non-periodic collimator: decided in fp64 only
  const float4* f = sc.slit_forms + 6 * cj;  // six forms of one crystal point are contiguous (96 B)
  const float Xc = form(__ldg(f), p), Yc = form(__ldg(f + 1), p), Wc = form(__ldg(f + 2), p);
  const float Xp = form(__ldg(f + 3), p), Yp = form(__ldg(f + 4), p);
  const float Wp = sc.slit_same_w ? Wc : form(__ldg(f + 5), p);
  if (fminf(fabsf(Wc), fabsf(Wp)) < 1e-3f) return kSlitUnknown;  // grazing ray: let fp64 decide
  const float ic = rcp_approx(Wc), ip = sc.slit_same_w ? ic : rcp_approx(Wp);
  const float xc = Xc * ic, yc = Yc * ic, xp = Xp * ip, yp = Yp * ip;
  const bool y_ok = yc >= sc.rect_crys.ylo && yc <= sc.rect_crys.yhi && yp >= sc.rect_plasma.ylo && yp <= sc.rect_plasma.yhi;
  // slit index from the crystal side; the polygon of slit k sticks out below k*T by |xlo| < T, so k and k+1 are tried
  const float kf = floorf(xc * sc.slit_inv_period);
  int code = -1;
  bool inner = false;
#pragma unroll
  for (int dk = 0; dk < 2; ++dk) {
    const float k = kf + (float)dk;
    const float a = fmaf(-k, sc.slit_period, xc), b = fmaf(-k, sc.slit_period, xp);
    const bool ok = k >= 0.f && k < (float)sc.n_slits && a >= sc.rect_crys.xlo && a <= sc.rect_crys.xhi &&
                    b >= sc.rect_plasma.xlo && b <= sc.rect_plasma.xhi;
    if (ok) {
      inner = code < 0 && a >= sc.inner_crys.xlo && a <= sc.inner_crys.xhi && b >= sc.inner_plasma.xlo && b <= sc.inner_plasma.xhi;
      code = code < 0 ? min((int)k, kSlitUnknown) : kSlitUnknown;
    }
  }
  if (!y_ok) return -1;
  inner = inner && code != kSlitUnknown && yc >= sc.inner_crys.ylo && yc <= sc.inner_crys.yhi &&
          yp >= sc.inner_plasma.ylo && yp <= sc.inner_plasma.yhi;
  return inner ? (code | kSlitCertain) : code;
}

// Detector (scaled frame of det_forms) and port (edge by edge, inner offsets) for a ray stage B found well inside its
// slit: when both hold too, no aperture test of the fp64 stage can fail and only the weight is left.
__device__ __forceinline__ bool certain_everywhere(const DeviceScene& sc, const float4 p, int cj) {
  const int Cp = sc.n_crystal_padded;
  const float xs = form(__ldg(&sc.det_forms[cj]), p), ys = form(__ldg(&sc.det_forms[Cp + cj]), p);
  const float wd = form(__ldg(&sc.det_forms[2 * Cp + cj]), p);
  const float sd = wd < 0.f ? -1.f : 1.f;
  if (!(fabsf(wd) > 1e-3f)) return false;
  if (sc.n_det_edge > 0) {
    float md = 0.f;
    for (int q = 0; q < sc.n_det_edge; ++q) {
      const float4 ed = sc.det_edge[q];
      md = fminf(md, sd * fmaf(ed.x, xs, fmaf(ed.y, ys, ed.w * wd)));
    }
    if (md < 0.f) return false;
  } else if (!(ratio_inside(sd, xs, wd, sc.det_inner.xlo, sc.det_inner.xhi) &&
               ratio_inside(sd, ys, wd, sc.det_inner.ylo, sc.det_inner.yhi))) {
    return false;
  }
  const float4* pf = sc.port_forms + 3 * cj;
  const float Xq = form(__ldg(pf), p), Yq = form(__ldg(pf + 1), p), Wq = form(__ldg(pf + 2), p);
  const float sp = Wq < 0.f ? -1.f : 1.f;
  float m = fabsf(Wq) - 1e-3f;  // > 0 required
  for (int q = 0; q < sc.n_port_edge; ++q) {
    const float4 ed = sc.port_edge[q];
    m = fminf(m, sp * fmaf(ed.x, Xq, fmaf(ed.y, Yq, ed.w * Wq)));
  }
  return m >= 0.f;
}

// ------------------------------------------------------------------------------------------------------
// Explicit voxels, packed for the filter: fp64 (P,3) -> recentred FP32 float4.  One pass, 24 B in + 16 B out per voxel.
__global__ void __launch_bounds__(256) voxel_pack_kernel(const double* __restrict__ xyz, long long begin, long long end,
                                                         double ox, double oy, double oz, float4* __restrict__ out) {
  for (long long i = begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (long long)gridDim.x * blockDim.x) {
    const D3 p = ld3(xyz + 3 * i);
    out[i] = make_float4((float)(p.x - ox), (float)(p.y - oy), (float)(p.z - oz), 1.f);
  }
}

// Bulk asynchronous copy global -> shared (the TMA unit's 1-D form, cp.async.bulk) completing on an mbarrier.
__device__ __forceinline__ unsigned int smem_addr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  unsigned int done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

__global__ void __launch_bounds__(kMaxWarps * 32) filter_kernel(const __grid_constant__ K1Params P) {
  __shared__ __align__(128) float4 s_tiles[2][kTileVoxels];  // [1] only with packed voxels (double buffering)
  __shared__ unsigned int s_queue[kMaxWarps][kQueueCap];
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ int s_item, s_next;

  const DeviceScene& sc = P.sc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Cp = sc.n_crystal_padded;
  const int n_items = P.n_tiles * P.n_super;
  unsigned int* queue = s_queue[warp];
  bool all_pass = false;  // filter disabled: every real ray is a candidate (padding lanes never are)

  int cur_super = -1;
  float4 gx, gy, gw;  // detector forms of this lane's crystal point
  int ci = 0;
  bool warp_active = false;

  // Packed voxels: the tile of the NEXT work item is already on its way (one elected thread, one bulk copy of
  // 16 B x nv into the other buffer, completion counted on an mbarrier) while this one is evaluated.
  const bool packed = P.vox.f4 != nullptr;
  unsigned int iter = 0;  // block-uniform: iteration i uses buffer i & 1, whose barrier is then in its (i >> 1)-th phase
  auto issue = [&](int it, unsigned int b) {  // thread 0 only
    const int t = it / P.n_super;
    const long long tb = (long long)t * kTileVoxels;
    const int nvt = (int)min((long long)kTileVoxels, P.n_vox - tb);
    bulk_load(s_tiles[b], P.vox.f4 + P.vox_begin + tb, (unsigned int)nvt * 16u, &s_bar[b]);
  };
  if (packed) {
    if (threadIdx.x == 0) {
      mbar_init(&s_bar[0], 1), mbar_init(&s_bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      s_next = (int)atomicAdd(&P.counters[1], 1ull);
      if (s_next < n_items) issue(s_next, 0);
    }
  }

  for (;;) {
    const unsigned int buf = packed ? (iter & 1u) : 0u, par = (iter >> 1) & 1u;
    ++iter;
    __syncthreads();  // previous tile fully consumed, s_item free
    if (threadIdx.x == 0) {
      if (packed) {
        s_item = s_next;
        if (s_item < n_items) {
          s_next = (int)atomicAdd(&P.counters[1], 1ull);
          if (s_next < n_items) issue(s_next, buf ^ 1u);
        }
      } else {
        s_item = (int)atomicAdd(&P.counters[1], 1ull);
      }
    }
    __syncthreads();
    const int item = s_item;
    if (item >= n_items) break;
    const int tile = item / P.n_super, sup = item - tile * P.n_super;
    const long long tile_base = (long long)tile * kTileVoxels;
    const int nv = (int)min((long long)kTileVoxels, P.n_vox - tile_base);
    const float4* s_tile = s_tiles[buf];

    if (packed) {
      mbar_wait(&s_bar[buf], par);
    } else {
      // stage the tile: fp64 position -> recentred FP32 (w = 1 so that a form is one dot4-like FFMA chain)
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        D3 p = voxel_position(P.vox, P.vox_begin + tile_base + i);
        s_tiles[0][i] = make_float4((float)(p.x - sc.origin[0]), (float)(p.y - sc.origin[1]), (float)(p.z - sc.origin[2]), 1.f);
      }
    }
    if (sup != cur_super) {
      cur_super = sup;
      ci = (sup * P.warps_per_cta + warp) * 32 + lane;
      warp_active = (sup * P.warps_per_cta + warp) * 32 < Cp;
      all_pass = sc.filter_disabled != 0 && ci < sc.n_crystal;
      if (warp_active) {
        gx = __ldg(&sc.det_forms[ci]);
        gy = __ldg(&sc.det_forms[Cp + ci]);
        gw = __ldg(&sc.det_forms[2 * Cp + ci]);
      }
    }
    __syncthreads();
    if (!warp_active) continue;

    int qn = 0;  // warp-uniform queue fill

    auto drain = [&](int take) {  // stage B on `take` (<= 32) queued rays, then append survivors
      __syncwarp();
      bool keep = false;
      unsigned int ent = 0;
      int slit = -1;
      if (lane < take) {
        ent = queue[qn - take + lane];
        slit = sc.filter_disabled != 0 ? kSlitUnknown : stage_b(sc, s_tile[ent >> 5], ci - lane + (int)(ent & 31));
        keep = slit >= 0;
        slit &= 0xF;
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
            P.cand[pos] = make_uint2((unsigned int)(tile_base + (ent >> 5)),
                                     (unsigned int)(ci - lane + (ent & 31)) | ((unsigned int)slit << 28));
        }
      }
      __syncwarp();
    };
    // margin of one ray against the mirrored detector rectangle: >= 0 passes (band folded into the forms)
    auto margin = [&](int v) {
      const float4 p = s_tile[v];
      const float xs = form(gx, p), ys = form(gy, p), w = form(gw, p);
      return fabsf(w) - fmaxf(fabsf(xs), fabsf(ys));
    };
    auto push = [&](int v, bool pass) {
      const unsigned int b = __ballot_sync(0xffffffffu, pass);
      if (b) {
        if (pass) queue[qn + __popc(b & ((1u << lane) - 1u))] = ((unsigned int)v << 5) | (unsigned int)lane;
        qn += __popc(b);
      }
    };

    int v0 = 0;
#pragma unroll 1
    for (; v0 + kUnroll <= nv; v0 += kUnroll) {
      float t[kUnroll];
#pragma unroll
      for (int i = 0; i < kUnroll; ++i) t[i] = margin(v0 + i);
      float tm = t[0];
#pragma unroll
      for (int i = 1; i < kUnroll; ++i) tm = fmaxf(tm, t[i]);
      if (__any_sync(0xffffffffu, tm >= 0.f || all_pass)) {
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) push(v0 + i, t[i] >= 0.f || all_pass);
        while (qn >= 32) drain(32);
      }
    }
    for (; v0 < nv; ++v0) {
      push(v0, margin(v0) >= 0.f || all_pass);
      while (qn >= 32) drain(32);
    }
    if (qn > 0) drain(qn);
  }
}

// ------------------------------------------------------------------------------------------------------
// K1a for the implicit CuboidMesh lattice.  z is the fastest lattice axis and the three detector forms are linear
// in the voxel position, so along a run of z-consecutive voxels each form advances by a per-lane constant
// (G . dz): 3 FADD per ray instead of 9 FFMA.  Runs ("segments") are re-based from an fp64-evaluated position
// every kSegmentMax voxels at most, which keeps the accumulated FP32 rounding (<= 64 ulp of |W| ~ 0.02) far
// inside the band folded into the forms (>= 0.4 in the same units).
//
// Warps are independent: a work item is (a few consecutive segments) x (32 crystal points), fetched by one warp
// with an atomic; there is no block-level barrier, so a warp that has to run stage B more often delays nobody.
// Only the segment bases are staged (per warp, in shared memory); stage B rebuilds a queued voxel's position as
// base + k * dz.
constexpr int kLatticeWarps = 8;
constexpr int kSegsPerItemMax = 16;

__device__ __forceinline__ D3 lattice_position_u32(const ComLattice& L, unsigned int flat) {
  const unsigned int nz = (unsigned int)L.nz, nx = (unsigned int)L.nx;
  const unsigned int tl = flat / nz, iz = flat - tl * nz;
  const unsigned int t = (unsigned int)global_column(L, tl), iy = t / nx, ix = t - iy * nx;
  return lattice_point(L, ix, iy, iz);
}

__global__ void __launch_bounds__(kLatticeWarps * 32) filter_kernel_lattice(const __grid_constant__ K1Params P) {
  __shared__ float4 s_base[kLatticeWarps][kSegsPerItemMax];  // segment base positions (recentred FP32, w = 1)
  __shared__ int2 s_seg[kLatticeWarps][kSegsPerItemMax];     // (first launch-local voxel, length)
  __shared__ unsigned int s_queue[kLatticeWarps][kQueueCap];
  __shared__ uint2 s_out[kLatticeWarps][64];                 // survivors of stage B waiting for a 32-wide flush
  __shared__ float4 s_outp[kLatticeWarps][64];               // ... and their recentred positions

  const DeviceScene& sc = P.sc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Cp = sc.n_crystal_padded, n_cgroups = Cp >> 5;
  const unsigned int nz = (unsigned int)P.vox.lat.nz;
  const unsigned int cpc = (nz + kSegmentMax - 1) / kSegmentMax;  // segments (chunks) per z-column
  const unsigned int flat_lo = (unsigned int)P.vox_begin, flat_hi = (unsigned int)(P.vox_begin + P.n_vox);
  // absolute segment index of a flat voxel index: column * cpc + chunk
  const unsigned int colA = flat_lo / nz, gA = colA * cpc + (flat_lo - colA * nz) / kSegmentMax;
  unsigned int* queue = s_queue[warp];
  uint2* outq = s_out[warp];
  float4* outp = s_outp[warp];
  const unsigned int lt_mask = (1u << lane) - 1u;
  int on = 0;  // warp-uniform number of pending survivors
  // which candidates fill the list from the front: the ones that need no aperture test at all when the weight-only
  // kernel takes the front region, else the ones that need no slit test
  const unsigned int front_flag = P.front_weight_only ? kCandCertainAll : kCandCertain;
  // append `count` (<= 32) pending survivors to the global candidate list: one atomic, one coalesced store
  // certain candidates grow from the front of the list, the others from its back: the exact kernel then runs warps of
  // one kind (a warp with a single uncertain lane would execute the tests the certain lanes skip anyway)
  auto flush = [&](int count) {
    __syncwarp();
    uint2 mine = outq[min(lane, 63)];
    if (P.front_weight_only && lane < count && (mine.y & kCandCertain) != 0u &&
        certain_everywhere(sc, outp[lane], (int)(mine.y & kCandCrystalMask)))
      mine.y |= kCandCertainAll;  // decided here, 32 compacted survivors at a time, not in the sparse drain
    const bool is_c = lane < count && (mine.y & front_flag) != 0u;
    const unsigned int bc = __ballot_sync(0xffffffffu, is_c);
    const int nc = __popc(bc), nu = count - nc;
    unsigned long long posc = 0, posu = 0;
    if (lane == 0) {
      if (nc) posc = atomicAdd(&P.counters[0], (unsigned long long)nc);
      if (nu) posu = atomicAdd(&P.counters[2], (unsigned long long)nu);
    }
    posc = __shfl_sync(0xffffffffu, posc, 0);
    posu = __shfl_sync(0xffffffffu, posu, 0);
    if (lane < count) {
      if (is_c) {
        const unsigned long long pos = posc + __popc(bc & lt_mask);
        if (pos < P.cand_capacity) P.cand[pos] = mine;
      } else {
        const unsigned long long pos = posu + (lane - __popc(bc & lt_mask));
        if (pos < P.cand_capacity) P.cand[P.cand_capacity - 1ull - pos] = mine;
      }
    }
    const uint2 carry = outq[min(lane + count, 63)];
    const float4 carryp = outp[min(lane + count, 63)];
    __syncwarp();
    if (lane + count < on) outq[lane] = carry, outp[lane] = carryp;
    on -= count;
    __syncwarp();
  };
  float4* base = s_base[warp];
  int2* segs = s_seg[warp];
  const long long n_items = (long long)P.n_tiles * n_cgroups;
  const bool run_cull = sc.filter_disabled == 0 && sc.no_run_cull == 0;  // skip z-runs proven to miss the detector

  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(&P.counters[1], 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((long long)item >= n_items) break;
    const unsigned int sg = (unsigned int)(item / n_cgroups);
    const int cg = (int)(item - (unsigned long long)sg * n_cgroups);
    const int ci = cg * 32 + lane;
    const bool all_pass = sc.filter_disabled != 0 && ci < sc.n_crystal;

    // stage this item's segments: lane l describes segment gA + sg * segs_per_item + l, clipped to the launch range
    __syncwarp();
    int nseg = 0;
    {
      const unsigned int g = gA + sg * (unsigned int)P.segs_per_item + (unsigned int)lane;
      const unsigned int col = g / cpc, iz0 = (g - col * cpc) * kSegmentMax;
      const unsigned long long a64 = (unsigned long long)col * nz + iz0;
      bool live = lane < P.segs_per_item && a64 < flat_hi;
      unsigned int a = 0, b = 0;
      if (live) {
        a = max((unsigned int)a64, flat_lo);
        b = (unsigned int)min((unsigned long long)flat_hi, a64 + min(nz - iz0, (unsigned int)kSegmentMax));
        live = b > a;
      }
      if (live) {
        const D3 p = lattice_position_u32(P.vox.lat, a);
        base[lane] = make_float4((float)(p.x - sc.origin[0]), (float)(p.y - sc.origin[1]), (float)(p.z - sc.origin[2]), 1.f);
        segs[lane] = make_int2((int)(a - flat_lo), (int)(b - a));
      }
      nseg = __popc(__ballot_sync(0xffffffffu, live));  // live segments are a prefix of the lanes
    }
    const float4 gx = __ldg(&sc.det_forms[ci]), gy = __ldg(&sc.det_forms[Cp + ci]), gw = __ldg(&sc.det_forms[2 * Cp + ci]);
    const float dxs = fmaf(gx.x, P.dz[0], fmaf(gx.y, P.dz[1], gx.z * P.dz[2]));
    const float dys = fmaf(gy.x, P.dz[0], fmaf(gy.y, P.dz[1], gy.z * P.dz[2]));
    const float dws = fmaf(gw.x, P.dz[0], fmaf(gw.y, P.dz[1], gw.z * P.dz[2]));
    __syncwarp();

    int qn = 0;
    auto drain = [&](int take) {
      __syncwarp();
      bool keep = false, certain = false;
      unsigned int ent = 0;
      int vloc = 0, slit = -1;
      float4 pos = make_float4(0.f, 0.f, 0.f, 1.f);
      if (lane < take) {
        ent = queue[qn - take + lane];
        const int seg = ent >> 11, k = (ent >> 5) & 63;
        const float4 b = base[seg];
        const float kf = (float)k;
        const float4 p = make_float4(fmaf(kf, P.dz[0], b.x), fmaf(kf, P.dz[1], b.y), fmaf(kf, P.dz[2], b.z), 1.f);
        vloc = segs[seg].x + k;
        const int cj = cg * 32 + (int)(ent & 31);
        slit = sc.filter_disabled != 0 ? kSlitUnknown : stage_b(sc, p, cj);
        keep = slit >= 0;
        certain = keep && (slit & kSlitCertain) != 0;  // well inside its slit on both sides
        slit &= 0xF;
        pos = p;
      }
      qn -= take;
      const unsigned int bal = __ballot_sync(0xffffffffu, keep);
      if (bal) {
        if (keep) {
          const int at = on + __popc(bal & lt_mask);
          outq[at] = make_uint2((unsigned int)vloc, (unsigned int)(cg * 32 + (ent & 31)) | (certain ? kCandCertain : 0u) |
                                                        ((unsigned int)slit << 28));
          outp[at] = pos;
        }
        on += __popc(bal);
        if (on >= 32) flush(32);
      }
      __syncwarp();
    };
    auto push = [&](int seg, int k, bool pass) {  // branch-free: a predicated store and two popcounts
      const unsigned int b = __ballot_sync(0xffffffffu, pass);
      if (pass) queue[qn + __popc(b & lt_mask)] = ((unsigned int)seg << 11) | ((unsigned int)k << 5) | (unsigned int)lane;
      qn += __popc(b);
    };

#pragma unroll 1
    for (int seg = 0; seg < nseg; ++seg) {
      const float4 p0 = base[seg];
      const int len = segs[seg].y;
      float xs = form(gx, p0), ys = form(gy, p0), w = form(gw, p0);
      // Hierarchical cull.  xs - w and xs + w are linear along the run, and a ray passes the X test iff they differ in
      // sign (|xs| <= |w|).  If they agree in sign at both ends of a run - |xs| > |w| at both ends, with xs of the same
      // sign - they agree everywhere in between, so no voxel of the run passes; likewise for Y.  First the whole
      // segment, then every group of kUnroll voxels; only the groups that survive are evaluated per voxel.  A passing
      // ray has a margin >= the band (0.4 in these units), the end values are good to 0.02: a run holding one is kept.
      auto culled = [&](float xa, float ya, float wa, float xb, float yb, float wb) {
        return (fabsf(xa) > fabsf(wa) && fabsf(xb) > fabsf(wb) && xa * xb > 0.f) ||
               (fabsf(ya) > fabsf(wa) && fabsf(yb) > fabsf(wb) && ya * yb > 0.f);
      };
      if (run_cull) {
        const float n1 = (float)(len - 1);
        if (__all_sync(0xffffffffu, culled(xs, ys, w, fmaf(n1, dxs, xs), fmaf(n1, dys, ys), fmaf(n1, dws, w)))) continue;
      }
      int k = 0;
#pragma unroll 1
      for (; k + kUnroll <= len; k += kUnroll) {
        if (run_cull) {
          constexpr float kLast = (float)(kUnroll - 1), kStep = (float)kUnroll;
          if (__all_sync(0xffffffffu, culled(xs, ys, w, fmaf(kLast, dxs, xs), fmaf(kLast, dys, ys), fmaf(kLast, dws, w)))) {
            xs = fmaf(kStep, dxs, xs), ys = fmaf(kStep, dys, ys), w = fmaf(kStep, dws, w);
            continue;
          }
        }
        float t[kUnroll];
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
          t[i] = fabsf(w) - fmaxf(fabsf(xs), fabsf(ys));
          xs += dxs, ys += dys, w += dws;
        }
        float tm = t[0];
#pragma unroll
        for (int i = 1; i < kUnroll; ++i) tm = fmaxf(tm, t[i]);
        if (__any_sync(0xffffffffu, tm >= 0.f || all_pass)) {
#pragma unroll
          for (int i = 0; i < kUnroll; ++i) push(seg, k + i, t[i] >= 0.f || all_pass);
          while (qn >= 32) drain(32);
        }
      }
      if (k < len && run_cull) {  // the ragged tail is one more run
        const float n1 = (float)(len - 1 - k);
        if (__all_sync(0xffffffffu, culled(xs, ys, w, fmaf(n1, dxs, xs), fmaf(n1, dys, ys), fmaf(n1, dws, w)))) continue;
      }
      for (; k < len; ++k) {
        const float t = fabsf(w) - fmaxf(fabsf(xs), fabsf(ys));
        xs += dxs, ys += dys, w += dws;
        push(seg, k, t >= 0.f || all_pass);
        while (qn >= 32) drain(32);
      }
    }
    if (qn > 0) drain(qn);
  }
  if (on > 0) flush(on);
}

// ------------------------------------------------------------------------------------------------------
// K1a, column-interval form (the production lattice kernel).
//
// Along a z-column the recentred voxel position is p0 + k * dz, so every linear form is f0 + k * df and every aperture
// test of the filter - a bound on a ratio of two forms whose denominator keeps its sign - is a linear inequality in
// the voxel index k.  The voxels of a (column, crystal point) pair that can reach the detector are therefore one
// interval [d_lo, d_hi] (four inequalities), the voxels that also go through slit j on both sides another interval
// (eight more), and only the slits the column can reach are visited (the slit coordinate at the two ends of the
// detector interval names them).  What costs ~15 instructions per ray when every ray is evaluated costs ~150 per
// (column, crystal point) pair here, whatever the column length.
//
// Two nested sets of bounds are intersected: the OUTER ones (bounding rectangles grown by eps - the same superset the
// per-ray kernel accepts) give the candidates, the INNER ones (rectangles at least `certain_mm` inside every polygon
// edge; for the slit length and the port they are checked at the two ends of the interval, which by convexity
// covers what lies between) give the rays whose aperture tests cannot fail: the fp64 stage only computes their
// weight.  Candidates outside the inner interval - the ends of each interval - are decided in fp64 as before.
// A column on which a denominator changes sign or comes close to zero is handed to the fp64 stage whole.
constexpr int kColsPerItem = 32;

// intersect [lo, hi] with { k : c0 + k * dc >= 0 }
__device__ __forceinline__ void clip_halfline(float& lo, float& hi, float c0, float dc) {
  const float t = -c0 * rcp_approx(dc);
  if (dc > 0.f) lo = fmaxf(lo, t);
  else if (dc < 0.f) hi = fminf(hi, t);
  else if (c0 < 0.f) lo = 1.f, hi = 0.f;
}

// bounds lo_b <= N/W <= hi_b with sign(W) = s, as two half-lines in k (N, W linear in k)
__device__ __forceinline__ void clip_ratio(float& lo, float& hi, float s, float n0, float dn, float w0, float dw, float lo_b,
                                           float hi_b) {
  clip_halfline(lo, hi, s * fmaf(-lo_b, w0, n0), s * fmaf(-lo_b, dw, dn));
  clip_halfline(lo, hi, s * fmaf(hi_b, w0, -n0), s * fmaf(hi_b, dw, -dn));
}

__device__ __forceinline__ float dz_dot(const float4 g, const float* dz) {
  return fmaf(g.x, dz[0], fmaf(g.y, dz[1], g.z * dz[2]));
}

__global__ void __launch_bounds__(kLatticeWarps * 32, 3) interval_kernel_lattice(const __grid_constant__ K1Params P) {
  __shared__ float4 s_base[kLatticeWarps][kColsPerItem];  // column base positions (recentred FP32, w = 1)
  __shared__ uint4 s_out[kLatticeWarps][64];             // interval records waiting for a 32-wide flush

  const DeviceScene& sc = P.sc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Cp = sc.n_crystal_padded, n_cgroups = Cp >> 5;
  const unsigned int nz = (unsigned int)P.vox.lat.nz;
  const unsigned int flat_lo = (unsigned int)P.vox_begin, flat_hi = (unsigned int)(P.vox_begin + P.n_vox);
  const unsigned int colA = flat_lo / nz, colB = (flat_hi - 1u) / nz;
  uint4* outq = s_out[warp];
  float4* base = s_base[warp];
  const unsigned int lt_mask = (1u << lane) - 1u;
  int on = 0;  // warp-uniform number of pending records
  // One 16-byte record per interval (not one entry per ray: the intervals of a warp's 32 crystal points rarely line up,
  // a per-ray emit loop runs with a handful of lanes): .x launch-local index of the interval's first voxel, .y crystal
  // index | slit hint << 28, .z length | (offset of the certain part) << 16, .w length of the certain part (0: none).
  // expand_kernel turns the records into the per-ray list the fp64 kernels read, with every lane busy.
  unsigned long long n_rays = 0;  // per lane: rays in the records it emitted (sizes the workspace after an overflow)
  auto flush = [&](int count) {
    __syncwarp();
    const uint4 mine = outq[min(lane, 63)];
    unsigned long long pos = 0;
    if (lane == 0) pos = atomicAdd(&P.counters[5], (unsigned long long)count);
    pos = __shfl_sync(0xffffffffu, pos, 0);
    if (lane < count && pos + lane < P.rec_capacity) P.recs[pos + lane] = mine;
    const uint4 carry = outq[min(lane + count, 63)];
    __syncwarp();
    if (lane + count < on) outq[lane] = carry;
    on -= count;
    __syncwarp();
  };
  // every lane offers the interval [a, b] of its column (certain where ci_lo <= k <= ci_hi)
  auto emit = [&](unsigned int col, int a, int b, int ci_lo, int ci_hi, int ci, int slit) {
    const bool have = a <= b;
    const unsigned int bal = __ballot_sync(0xffffffffu, have);
    if (have) {
      const unsigned int v0 = col * nz - flat_lo;  // launch-local index of the column's voxel 0 (wraps for a clipped first column)
      const bool cert = ci_lo <= ci_hi;
      const unsigned int len = (unsigned int)(b - a + 1), c_off = cert ? (unsigned int)(ci_lo - a) : 0u,
                         c_len = cert ? (unsigned int)(ci_hi - ci_lo + 1) : 0u;
      outq[on + __popc(bal & lt_mask)] = make_uint4(v0 + (unsigned int)a, (unsigned int)ci | ((unsigned int)slit << 28),
                                                    len | (c_off << 16), c_len);
      n_rays += len;
    }
    on += __popc(bal);
    if (on >= 32) flush(32);
  };

  const float nz1 = (float)(nz - 1u);
  const float T = sc.slit_period, invT = sc.slit_inv_period;
  const long long n_items = (long long)P.n_tiles * n_cgroups;
  for (;;) {
    unsigned long long item = 0;
    if (lane == 0) item = atomicAdd(&P.counters[1], 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if ((long long)item >= n_items) break;
    const unsigned int cb = (unsigned int)(item / n_cgroups);
    const int cg = (int)(item - (unsigned long long)cb * n_cgroups);
    const int ci = cg * 32 + lane;
    const bool real = ci < sc.n_crystal;

    __syncwarp();
    const unsigned int col_first = colA + cb * kColsPerItem;
    {
      const unsigned int col = col_first + (unsigned int)lane;
      if (col <= colB) {
        const D3 p = lattice_position_u32(P.vox.lat, col * nz);
        base[lane] = make_float4((float)(p.x - sc.origin[0]), (float)(p.y - sc.origin[1]), (float)(p.z - sc.origin[2]), 1.f);
      }
    }
    const int ncols = (int)min((unsigned int)kColsPerItem, colB - col_first + 1u);
    // per-lane constants of this crystal point: detector and slit forms with their z-derivatives
    const float4 gx = __ldg(&sc.det_forms[ci]), gy = __ldg(&sc.det_forms[Cp + ci]), gw = __ldg(&sc.det_forms[2 * Cp + ci]);
    const float dxs = dz_dot(gx, P.dz), dys = dz_dot(gy, P.dz), dws = dz_dot(gw, P.dz);
    const float4* sf = sc.slit_forms + 6 * ci;
    const float4 fXc = __ldg(sf), fYc = __ldg(sf + 1), fWc = __ldg(sf + 2), fXp = __ldg(sf + 3), fYp = __ldg(sf + 4);
    const float dXc = dz_dot(fXc, P.dz), dYc = dz_dot(fYc, P.dz), dWc = dz_dot(fWc, P.dz), dXp = dz_dot(fXp, P.dz),
                dYp = dz_dot(fYp, P.dz);
    __syncwarp();

#pragma unroll 1
    for (int j = 0; j < ncols; ++j) {
      const unsigned int col = col_first + (unsigned int)j;
      const float4 p0 = base[j];
      // launch range inside this column
      const int k_first = col == colA ? (int)(flat_lo - colA * nz) : 0;
      const int k_last = col == colB ? (int)(flat_hi - 1u - colB * nz) : (int)nz - 1;
      // ---- mirrored detector: one interval ----
      const float xs0 = form(gx, p0), ys0 = form(gy, p0), w0 = form(gw, p0);
      const float w1 = fmaf(nz1, dws, w0);
      bool weird = real && (!(w0 * w1 > 0.f) || fminf(fabsf(w0), fabsf(w1)) < 1e-3f);
      const float s = w0 < 0.f ? -1.f : 1.f;
      float lo = (float)k_first, hi = (float)k_last;
      clip_ratio(lo, hi, s, xs0, dxs, w0, dws, -1.f, 1.f);
      clip_ratio(lo, hi, s, ys0, dys, w0, dws, -1.f, 1.f);
      int d_lo = (int)ceilf(lo - 1e-3f), d_hi = (int)floorf(hi + 1e-3f);
      d_lo = max(d_lo, k_first), d_hi = min(d_hi, k_last);
      if (!real || !(lo <= hi + 2e-3f)) d_lo = 1, d_hi = 0;  // (NaN bounds compare false: empty)
      if (weird) d_lo = k_first, d_hi = k_last;
      if (!__any_sync(0xffffffffu, d_lo <= d_hi)) continue;
      // certain part of the detector interval
      // (inner bounds start unbounded: only thresholds that really clip are moved inwards by 1e-3 voxel below)
      float li = -1e9f, hi_i = 1e9f;
      if (sc.n_det_edge > 0) {  // the detector polygon edge by edge, inner offsets
        for (int q = 0; q < sc.n_det_edge; ++q) {
          const float4 ed = sc.det_edge[q];
          clip_halfline(li, hi_i, s * fmaf(ed.x, xs0, fmaf(ed.y, ys0, ed.w * w0)), s * fmaf(ed.x, dxs, fmaf(ed.y, dys, ed.w * dws)));
        }
      } else {
        clip_ratio(li, hi_i, s, xs0, dxs, w0, dws, sc.det_inner.xlo, sc.det_inner.xhi);
        clip_ratio(li, hi_i, s, ys0, dys, w0, dws, sc.det_inner.ylo, sc.det_inner.yhi);
      }
      if (!(sc.det_inner.xlo <= sc.det_inner.xhi)) li = 1.f, hi_i = 0.f;

      // ---- collimator: which slits can the detector interval reach ----
      const float Xc0 = form(fXc, p0), Yc0 = form(fYc, p0), Wc0 = form(fWc, p0), Xp0 = form(fXp, p0), Yp0 = form(fYp, p0);
      const float Wc1 = fmaf(nz1, dWc, Wc0);
      if (d_lo <= d_hi && (!(Wc0 * Wc1 > 0.f) || fminf(fabsf(Wc0), fabsf(Wc1)) < 1e-3f)) weird = true;
      const bool active = d_lo <= d_hi && !weird;
      const float sq = Wc0 < 0.f ? -1.f : 1.f;
      int j_lo = 1, j_hi = 0;
      if (active) {
        const float ka = (float)d_lo, kb = (float)d_hi;
        const float xa = fmaf(ka, dXc, Xc0) * rcp_approx(fmaf(ka, dWc, Wc0));
        const float xb = fmaf(kb, dXc, Xc0) * rcp_approx(fmaf(kb, dWc, Wc0));
        j_lo = max(0, (int)ceilf((fminf(xa, xb) - sc.rect_crys.xhi) * invT - 1e-3f));
        j_hi = min(sc.n_slits - 1, (int)floorf((fmaxf(xa, xb) - sc.rect_crys.xlo) * invT + 1e-3f));
      }
      if (weird) {  // a denominator changes sign or vanishes on this column: every voxel of it goes to the fp64 stage
        j_lo = 0, j_hi = 0;
      }
      const int slit_trips = __reduce_max_sync(0xffffffffu, j_hi - j_lo + 1);
      // port forms of this crystal point at the column base (only needed where something survives)
      float Xq0 = 0.f, Yq0 = 0.f, Wq0 = 1.f, dXq = 0.f, dYq = 0.f, dWq = 0.f;
      bool port_loaded = false;
      for (int t = 0; t < slit_trips; ++t) {
        const int js = j_lo + t;
        int a = 1, b = 0, ci_lo = 1, ci_hi = 0, hint = kSlitUnknown;
        if (weird && t == 0) {
          a = d_lo, b = d_hi;
        } else if (active && js <= j_hi) {
          // the polygon of slit js on both sides, edge by edge: s (a X + b Y + (c - a js T) W) >= 0, outer and inner offsets
          const float off = (float)js * T;
          float lo_j = (float)d_lo, hi_j = (float)d_hi, l2 = li, h2 = hi_i;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const float X0 = side ? Xp0 : Xc0, Y0 = side ? Yp0 : Yc0, dX = side ? dXp : dXc, dY = side ? dYp : dYc;
            for (int q = 0; q < sc.n_slit_edge[side]; ++q) {
              const float4 ed = sc.slit_edge[side][q];
              const float base0 = fmaf(ed.x, X0, ed.y * Y0), based = fmaf(ed.x, dX, ed.y * dY);
              const float co = fmaf(-ed.x, off, ed.z), cin = fmaf(-ed.x, off, ed.w);
              clip_halfline(lo_j, hi_j, sq * fmaf(co, Wc0, base0), sq * fmaf(co, dWc, based));
              clip_halfline(l2, h2, sq * fmaf(cin, Wc0, base0), sq * fmaf(cin, dWc, based));
            }
          }
          if (lo_j <= hi_j + 2e-3f) {
            a = max(d_lo, (int)ceilf(lo_j - 1e-3f)), b = min(d_hi, (int)floorf(hi_j + 1e-3f));
            hint = min(js, kSlitUnknown);
          }
          if (a <= b && !sc.certain_off) {
            // certain part: inner offsets of the slit edges and of the detector rectangle; the port at both ends
            if (l2 <= h2) ci_lo = max(a, (int)ceilf(fmaxf(l2, -1e6f) + 1e-3f)), ci_hi = min(b, (int)floorf(fminf(h2, 1e6f) - 1e-3f));
            if (ci_lo <= ci_hi) {
              if (!port_loaded) {
                const float4* pf = sc.port_forms + 3 * ci;
                const float4 fXq = __ldg(pf), fYq = __ldg(pf + 1), fWq = __ldg(pf + 2);
                Xq0 = form(fXq, p0), Yq0 = form(fYq, p0), Wq0 = form(fWq, p0);
                dXq = dz_dot(fXq, P.dz), dYq = dz_dot(fYq, P.dz), dWq = dz_dot(fWq, P.dz);
                port_loaded = true;
              }
              // the port denominator must keep its sign (and stay away from zero) between the two ends
              const float wqa = fmaf((float)ci_lo, dWq, Wq0), wqb = fmaf((float)ci_hi, dWq, Wq0);
              bool ok = sc.port_inner.xlo <= sc.port_inner.xhi && wqa * wqb > 0.f && fminf(fabsf(wqa), fabsf(wqb)) > 1e-3f;
              const float sp = wqa < 0.f ? -1.f : 1.f;
              bool in_rect = ok;  // both ends inside the inner rectangle: nothing more to do (convexity)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float k = (float)(e ? ci_hi : ci_lo);
                const float wq = e ? wqb : wqa;
                in_rect = in_rect && ratio_inside(sp, fmaf(k, dXq, Xq0), wq, sc.port_inner.xlo, sc.port_inner.xhi) &&
                          ratio_inside(sp, fmaf(k, dYq, Yq0), wq, sc.port_inner.ylo, sc.port_inner.yhi);
              }
              if (ok && !in_rect && sc.n_port_edge > 0) {
                // the port polygon edge by edge, inner offsets: each is one more linear inequality in k
                float l3 = -1e9f, h3 = 1e9f;
                for (int q = 0; q < sc.n_port_edge; ++q) {
                  const float4 ed = sc.port_edge[q];
                  clip_halfline(l3, h3, sp * fmaf(ed.x, Xq0, fmaf(ed.y, Yq0, ed.w * Wq0)),
                                sp * fmaf(ed.x, dXq, fmaf(ed.y, dYq, ed.w * dWq)));
                }
                if (l3 <= h3)
                  ci_lo = max(ci_lo, (int)ceilf(fmaxf(l3, -1e6f) + 1e-3f)), ci_hi = min(ci_hi, (int)floorf(fminf(h3, 1e6f) - 1e-3f));
                else
                  ok = false;
              } else {
                ok = ok && in_rect;
              }
              if (!ok) ci_lo = 1, ci_hi = 0;
            }
          }
        }
        if (__any_sync(0xffffffffu, a <= b)) emit(col, a, b, ci_lo, ci_hi, ci, hint);
      }
    }
  }
  if (on > 0) flush(on);
  n_rays = __reduce_add_sync(0xffffffffu, (unsigned int)min(n_rays, 0xFFFFFFFFull >> 5));  // per lane < 2^27 in any launch
  if (lane == 0 && n_rays) atomicAdd(&P.counters[6], n_rays);
}

// ------------------------------------------------------------------------------------------------------
// Interval records -> the per-ray candidate list.  A block takes 256 records, scans the lengths of their certain and
// uncertain parts, reserves the two ranges of the list with one atomic each (certain rays fill the list from the front -
// the weight-only kernel's region - the others from its back) and writes the rays with all lanes busy: ray r of the
// block belongs to the record found by a binary search in the scanned lengths.
constexpr int kExpandThreads = 256;

__global__ void __launch_bounds__(kExpandThreads) expand_kernel(const __grid_constant__ K1Params P) {
  __shared__ uint4 s_rec[kExpandThreads];
  __shared__ unsigned int s_c[kExpandThreads + 1], s_u[kExpandThreads + 1];  // exclusive scans of the two lengths
  __shared__ unsigned int s_warp[2][kExpandThreads / 32];
  __shared__ unsigned long long s_pos[2];
  const unsigned long long reserved = P.counters[5];
  if (reserved > P.rec_capacity) {
    // The record list overflowed: make the ray list report it (exact_kernel checks), with a candidate count that sizes a
    // workspace in which both lists fit - callers re-run with candidate_fraction = that count / tests.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const unsigned long long by_recs = (reserved * P.cand_capacity + P.rec_capacity - 1ull) / max(P.rec_capacity, 1ull);
      P.counters[0] = max(max(P.counters[6], by_recs), P.cand_capacity) + 1ull;
      P.counters[2] = 0ull;
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned long long first = (unsigned long long)blockIdx.x * kExpandThreads; first < reserved;
       first += (unsigned long long)gridDim.x * kExpandThreads) {
    const unsigned long long i = first + threadIdx.x;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (i < reserved) rec = P.recs[i];
    const unsigned int len = rec.z & 0xFFFFu, c_len = rec.w, u_len = len - c_len;
    // block-wide exclusive scans of c_len and u_len
    unsigned int ic = c_len, iu = u_len;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int a = __shfl_up_sync(0xffffffffu, ic, off), b = __shfl_up_sync(0xffffffffu, iu, off);
      if (lane >= off) ic += a, iu += b;
    }
    if (lane == 31) s_warp[0][warp] = ic, s_warp[1][warp] = iu;
    s_rec[threadIdx.x] = rec;
    __syncthreads();
    unsigned int bc = 0, bu = 0;
    for (int q = 0; q < warp; ++q) bc += s_warp[0][q], bu += s_warp[1][q];
    s_c[threadIdx.x + 1] = bc + ic, s_u[threadIdx.x + 1] = bu + iu;
    if (threadIdx.x == 0) s_c[0] = 0u, s_u[0] = 0u;
    __syncthreads();
    const unsigned int Tc = s_c[kExpandThreads], Tu = s_u[kExpandThreads];
    const bool list_certain = P.records_to_weight == 0;
    if (threadIdx.x == 0) {
      s_pos[0] = (Tc && list_certain) ? atomicAdd(&P.counters[0], (unsigned long long)Tc) : 0ull;
      s_pos[1] = Tu ? atomicAdd(&P.counters[2], (unsigned long long)Tu) : 0ull;
    }
    __syncthreads();
    const unsigned long long pc = s_pos[0], pu = s_pos[1];
    auto owner = [](const unsigned int* scan, unsigned int r) {  // largest i with scan[i] <= r
      int lo = 0, hi = kExpandThreads - 1;
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (scan[mid] <= r) lo = mid; else hi = mid - 1;
      }
      return lo;
    };
    for (unsigned int r = threadIdx.x; list_certain && r < Tc; r += kExpandThreads) {
      const int q = owner(s_c, r);
      const uint4 e = s_rec[q];
      const unsigned long long pos = pc + r;
      if (pos < P.cand_capacity)
        P.cand[pos] = make_uint2(e.x + (e.z >> 16) + (r - s_c[q]), e.y | kCandCertain | kCandCertainAll);
    }
    for (unsigned int r = threadIdx.x; r < Tu; r += kExpandThreads) {
      const int q = owner(s_u, r);
      const uint4 e = s_rec[q];
      const unsigned int j = r - s_u[q], c_off = e.z >> 16;
      const unsigned long long pos = pu + r;
      if (pos < P.cand_capacity) P.cand[P.cand_capacity - 1ull - pos] = make_uint2(e.x + (j < c_off ? j : j + e.w), e.y);
    }
    __syncthreads();  // s_rec / scans are rewritten by the next batch
  }
}

// ------------------------------------------------------------------------------------------------------
// K1b: the reference arithmetic, one candidate ray per thread (grid-stride over the candidate list).
struct HitResult {
  bool pass;
  double margin;
  double x, y;  // in-plane coordinates along e1, e2 (x picks the slit; both address the detector image)
  double s;     // line parameter of the hit: X = dst + s (src - dst)
};

// simulation.py:143-173 (+175-207): line through `dst` with direction src - dst against the aperture plane,
// then the polygon test that replaces Delaunay.find_simplex (geometry_host.cu / apertures.py).
__device__ __forceinline__ HitResult aperture_hit(const DevAperture& ap, const double* __restrict__ edges,
                                                  const float* __restrict__ edges_f, float tau, D3 src, D3 dst) {
  const D3 N = ld3(ap.normal), u = src - dst;
  const double ndotu = dot(N, u);
  if (fabs(ndotu) < 1e-6) return {false, -1e300, 0.0, 0.0, 0.0};  // the reference writes NaN -> never inside
  const D3 w = dst - ld3(ap.p1);
  const double si = -(dot(N, w) / ndotu);
  const D3 r = w + si * u;  // X - p1
  const double x = dot(r, ld3(ap.e1)), y = dot(r, ld3(ap.e2));
  // The hit point is fp64 with the reference's arithmetic; only its distance to the polygon edges is first taken
  // in FP32 (unit edge normals, |x|,|y| < 1e3 mm: good to ~1e-4 mm).  A point further than tau (>= 2e-3 mm) from
  // every edge line is decided; anything closer is decided by the fp64 margins below.
  {
    const float xf = (float)x, yf = (float)y;
    const float* ef = edges_f + 3 * ap.edge_offset;
    float mf = 3.0e38f;
    for (int k = 0; k < ap.n_edges; ++k) mf = fminf(mf, fmaf(ef[3 * k], xf, fmaf(ef[3 * k + 1], yf, ef[3 * k + 2])));
    if (fabsf(mf) > tau) return {mf > 0.f, (double)mf, x, y, si};
  }
  double m = 1e300;
  const double* e = edges + 3 * ap.edge_offset;
  for (int k = 0; k < ap.n_edges; ++k) m = fmin(m, e[3 * k] * x + e[3 * k + 1] * y + e[3 * k + 2]);
  return {m >= 0.0, m, x, y, si};
}

// Extension (off on the parity path): one ECRH shield as a circular stop - the line src -> dst against the shield's
// mid-plane, inside the regular n-gon rim (ComShield).  margin > 0 inside.  A line parallel to the plane never goes
// through the stop.
__device__ __forceinline__ HitResult shield_hit(const ComShield& sh, D3 src, D3 dst) {
  const D3 a = ld3(sh.axis), u = src - dst;
  const double adotu = dot(a, u);
  if (fabs(adotu) < 1e-12) return {false, -1e300, 0.0, 0.0, 0.0};
  const D3 w = dst - ld3(sh.centre);
  const double si = -(dot(a, w) / adotu);
  const D3 r = w + si * u;
  const double x = dot(r, ld3(sh.u)), y = dot(r, ld3(sh.v));
  const double rr = sqrt(x * x + y * y);
  if (sh.n_sides <= 0) return {sh.radius - rr >= 0.0, sh.radius - rr, x, y, si};
  const double step = 6.283185307179586 / (double)sh.n_sides;
  const double inr = sh.radius * cos(0.5 * step);  // inradius of the rim polygon
  double m = inr - rr;                              // >= 0: inside whatever the direction
  if (m < 0.0 || m < 1e-2) {
    // between inradius and circumradius (or close): the edge of the hit's own sector and its two neighbours decide
    double phi = atan2(y, x);
    if (phi < 0.0) phi += 6.283185307179586;
    const int k0 = (int)floor(phi / step);
    m = 1e300;
    for (int dk = -1; dk <= 1; ++dk) {
      const double th = ((double)(k0 + dk) + 0.5) * step;
      m = fmin(m, inr - (x * cos(th) + y * sin(th)));
    }
  }
  return {m >= 0.0, m, x, y, si};
}

constexpr int kExactThreads = 256;

__global__ void __launch_bounds__(kExactThreads) exact_kernel(const __grid_constant__ K1Params P) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const DeviceScene& sc = P.sc;
  // front region (certain first, or everything from the explicit-voxel kernel) and back region (uncertain)
  const unsigned long long total = P.counters[0] + P.counters[2];
  const bool overflow = total > P.cand_capacity;  // the two regions collided: nothing of this launch can be trusted
  const unsigned long long n_front = overflow ? 0ull : P.counters[0];
  const unsigned long long n = overflow ? 0ull : total;
  const unsigned long long i_first = P.front_weight_only ? n_front : 0ull;  // weight_kernel owns the front region
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long direct = P.records_to_weight ? P.counters[7] : 0ull;  // certain rays that never were listed
    const unsigned long long written = total + direct;
    const unsigned long long certain = (sc.filter_disabled || P.tile_voxels != 0) ? 0ull : min(P.counters[0], total) + direct;
    const unsigned long long tests = (unsigned long long)P.n_vox * (unsigned long long)sc.n_crystal;
    if (P.accumulate_stats) {  // launches of one stream run one after the other: plain read-modify-write
      P.stats->candidates += written, P.stats->certain_candidates += certain, P.stats->tests += tests;
      P.stats->candidate_capacity = max((unsigned long long)P.stats->candidate_capacity, P.cand_capacity);
      if (overflow) atomicOr((unsigned long long*)&P.stats->overflow, 1ull);
    } else {
      P.stats->candidates = written, P.stats->certain_candidates = certain, P.stats->tests = tests;
      P.stats->candidate_capacity = P.cand_capacity;
      if (overflow) atomicOr((unsigned long long*)&P.stats->overflow, 1ull);  // stats were zeroed before the launch
    }
  }
  if (i_first + (unsigned long long)blockIdx.x * kExactThreads >= n) return;
  // apertures + edge pool -> shared memory (a few kB)
  DevAperture* s_ap = reinterpret_cast<DevAperture*>(s_raw);
  double* s_edges = reinterpret_cast<double*>(s_raw + sizeof(DevAperture) * sc.n_apertures);
  float* s_edges_f = reinterpret_cast<float*>(s_edges + 3 * sc.n_edges_total);
  {
    const double* src = reinterpret_cast<const double*>(sc.apertures);
    double* dst = reinterpret_cast<double*>(s_ap);
    const int n_ap_d = (int)(sizeof(DevAperture) / sizeof(double)) * sc.n_apertures;
    for (int i = threadIdx.x; i < n_ap_d; i += kExactThreads) dst[i] = src[i];
    for (int i = threadIdx.x; i < 3 * sc.n_edges_total; i += kExactThreads) {
      const double v = sc.edges[i];
      s_edges[i] = v;
      s_edges_f[i] = (float)v;
    }
  }
  __syncthreads();

  const int S = sc.n_slits;
  unsigned int n_pass = 0, n_edge = 0, n_band = 0, n_shield = 0, n_seg = 0, n_round = 0;
  const double kNear = sc.near_edge_mm, kBand = sc.near_band_mm;
  const bool seg_check = (sc.physics_flags & COM_PHYS_SEGMENT_CHECK) != 0;
  // FP32 edge distances decide beyond this; twice the near-edge radii so that the near-edge counts stay exact
  const float tau = fmaxf(2.0e-3f, 2.f * (float)sc.near_band_mm);
  for (unsigned long long i = i_first + (unsigned long long)blockIdx.x * kExactThreads + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * kExactThreads) {
    const uint2 e = i < n_front ? P.cand[i] : P.cand[P.cand_capacity - 1ull - (i - n_front)];
    const long long vloc = e.x;
    const int ci = (int)(e.y & kCandCrystalMask), slit_hint = (int)(e.y >> 28);
    if (vloc >= P.n_vox || ci >= sc.n_crystal) {  // never true for a list the filters wrote: a guard, not a code path
      atomicOr((unsigned long long*)&P.stats->overflow, 2ull);  // (no compute-sanitizer on this pool; tests require 0)
      continue;
    }
    const bool certain = (e.y & kCandCertain) != 0u;  // K1a found the ray >= certain_mm inside one slit on both sides
    // ... and inside port and detector: nothing is left to test (the detector image still needs the detector hit)
    const bool certain_all = (e.y & kCandCertainAll) != 0u && sc.pix_out == nullptr;
    const D3 p = voxel_position(P.vox, P.vox_begin + vloc);
    const D3 c = ld3(sc.crystal_xyz + 3 * ci), nrm = ld3(sc.crystal_normal + 3 * ci);
    // simulation.py:309-321
    const D3 d = c - p;
    const D3 r = d - (2.0 * dot(d, nrm)) * nrm;
    const D3 reflected = c + r, plasma = c - d, crystal = p + d;

    double near = 1e300;      // smallest |margin| over the apertures this ray was tested on
    bool seg_ok = true;       // every accepted hit lies on the physical part of its line (extension)
    auto note = [&](const HitResult& h) { near = fmin(near, fabs(h.margin)); };
    HitResult h{true, 1e300, 0.0, 0.0, 0.5};
    if (!certain_all) {
      h = aperture_hit(s_ap[2 * S], s_edges, s_edges_f, tau, plasma, crystal);  // port :372-384
      note(h);
    }
    bool ok = h.pass;
    seg_ok = seg_ok && h.s >= 0.0 && h.s <= 1.0;
    if (ok && !certain) {  // collimator :344-369, the same slit on both sides
      auto both_sides = [&](int k) {
        const HitResult a = aperture_hit(s_ap[2 * k], s_edges, s_edges_f, tau, plasma, crystal);
        note(a);
        if (!a.pass) return false;
        const HitResult b = aperture_hit(s_ap[2 * k + 1], s_edges, s_edges_f, tau, plasma, crystal);
        note(b);
        if (b.pass) seg_ok = seg_ok && a.s >= 0.0 && a.s <= 1.0 && b.s >= 0.0 && b.s <= 1.0;
        return b.pass;
      };
      // the slit the FP32 filter saw is tried first: "any slit passes" is all the reference asks (:408-411)
      ok = slit_hint < S && slit_hint != kSlitUnknown && both_sides(slit_hint);
      if (!ok) {
        int k_first = -1, k_lo = 0, k_hi = S - 1;
        if (sc.slit_period_d > 0.0) {
          // slit k is slit 0 moved by k*T along e1, so the hit on slit 0's plane names the only slits that can contain
          // it: k0 = floor(x / T), k0 + 1 when the hit is within slit 0's reach below a period boundary (its polygon
          // starts at xlo < 0), k0 - 1 when it is within its reach above one (xhi > T)
          const HitResult h0 = aperture_hit(s_ap[0], s_edges, s_edges_f, tau, plasma, crystal);
          if (h0.margin == -1e300) {
            k_hi = -1;  // parallel to the collimator plane: NaN in the reference
          } else {
            const double T = sc.slit_period_d, q = h0.x / T;
            const int k0 = (int)floor(q);
            const double off = (q - floor(q)) * T;  // position inside the period, [0, T)
            k_first = min(max(k0, 0), S - 1);
            k_lo = off <= sc.slit_xhi_d - T + 1e-3 ? max(0, k0 - 1) : max(0, k0);
            k_hi = off >= T + sc.slit_xlo_d - 1e-3 ? min(S - 1, k0 + 1) : min(S - 1, k0);
          }
        }
        if (k_hi >= 0) {
          if (k_first >= 0) ok = k_first != slit_hint && both_sides(k_first);
          for (int k = k_lo; k <= k_hi && !ok; ++k)
            if (k != k_first && k != slit_hint) ok = both_sides(k);
        }
      }
    }
    if (ok && !certain_all) {
      h = aperture_hit(s_ap[2 * S + 1], s_edges, s_edges_f, tau, reflected, crystal);  // detector :387-400
      note(h);
      ok = h.pass;
      seg_ok = seg_ok && h.s >= 0.0;
    }
    n_edge += near < kNear ? 1u : 0u;
    n_band += near < kBand ? 1u : 0u;
    if (ok && seg_check && !seg_ok) {  // extension: the reference's infinite lines accept it, the physical ray does not
      ++n_seg;
      ok = false;
    }
    if (ok && sc.n_shields > 0) {  // extension: every ECRH shield of the channel is a circular stop
      bool through = true;
      for (int q = 0; q < sc.n_shields && through; ++q) {
        const HitResult g = shield_hit(sc.shields[q], plasma, crystal);
        if (fabs(g.margin) < kBand) ++n_band;
        through = g.pass && (!seg_check || (g.s >= 0.0 && g.s <= 1.0));
      }
      if (!through) {
        ++n_shield;
        ok = false;
      }
    }
    if (ok) {
      // simulation.py:494-497, 516-531, 564-577
      const D3 v = plasma - crystal, rc = reflected - crystal;
      const double vv = dot(v, v), dist = sqrt(vv);
      const double cosang = dot(v, rc) / (sqrt(vv) * sqrt(dot(rc, rc)));
      const double deg = acos(cosang) * (180.0 / 3.141592653589793);
      const double aoi = (180.0 - rint(deg * 100.0) / 100.0) / 2.0;
      if (fabs(fabs(deg * 100.0 - rint(deg * 100.0)) - 0.5) < 1e-9) ++n_round;
      const double fraction = sc.area_pt / (4.0 * 3.141592653589793 * (dist * dist));
      const double da = aoi - sc.aoi0;
      const double prof = sc.refl_max * exp(-(da * da) / (2.0 * (sc.refl_sigma * sc.refl_sigma)));
      const double w = fraction * prof;

      ++n_pass;
      atomicAdd(&P.w_out[vloc], w);
      if (P.nrays_out) {
        atomicAdd(&P.nrays_out[vloc], 1u);  // result unused: a reduction, nobody waits for the round trip to L2
      }
      if (P.hist_out) {
        const int bin = P.reff_bin[P.vox_begin + vloc];
        if (bin < P.n_bins) atomicAdd(&P.hist_out[bin], w);
      }
      if (sc.pix_out) {  // h still holds the detector hit
        const int px = min(max((int)floor((h.x - sc.pix_x0) / sc.pix_mm), 0), sc.pix_nx - 1);
        const int py = min(max((int)floor((h.y - sc.pix_y0) / sc.pix_mm), 0), sc.pix_ny - 1);
        atomicAdd(&sc.pix_out[(size_t)py * sc.pix_nx + px], w);
      }
      if (P.rays_out) {
        const unsigned long long pos = atomicAdd(P.rays_count, 1ull);
        if ((long long)pos < P.rays_capacity)
          P.rays_out[pos] = ComRayRecord{(P.vox_begin + vloc) * (long long)sc.n_crystal + ci, dist, aoi, w};
      }
    }
  }

  // block-level stats
  constexpr int kStats = 6;  // visible voxels are counted by count_visible_kernel afterwards
  __shared__ unsigned int s_acc[kStats];
  if (threadIdx.x < kStats) s_acc[threadIdx.x] = 0;
  __syncthreads();
  unsigned int mine[kStats] = {n_pass, n_edge, n_band, n_shield, n_seg, n_round};
#pragma unroll
  for (int q = 0; q < kStats; ++q) {
    unsigned int v = mine[q];
    for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&s_acc[q], v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* out[kStats] = {(unsigned long long*)&P.stats->passing_rays, (unsigned long long*)&P.stats->near_edge_rays,
                                       (unsigned long long*)&P.stats->near_band_rays,
                                       (unsigned long long*)&P.stats->shield_blocked_rays,
                                       (unsigned long long*)&P.stats->segment_rejected_rays,
                                       (unsigned long long*)&P.stats->near_rounding_rays};
    for (int q = 0; q < kStats; ++q)
      if (s_acc[q]) atomicAdd(out[q], (unsigned long long)s_acc[q]);
  }
}

// ------------------------------------------------------------------------------------------------------
// K1w: the weight of a ray no aperture test can fail (the front region of the column-interval kernel's list): the
// reference's distance, angle of incidence with round(deg, 2) and weight (simulation.py:309-321, 494-497, 516-531,
// 564-577), nothing else.  The reflectivity profile is read from a table over the 18001 values the rounded angle can
// take.
constexpr int kWeightThreads = 256;

// n / d for 32-bit n by one 64-bit high multiply: m = floor((2^64 - 1) / d) + 1 (exact for every n < 2^32, d >= 2)
__device__ __forceinline__ unsigned int fast_div(unsigned int n, unsigned long long m, unsigned int d) {
  return d == 1u ? n : (unsigned int)__umul64hi((unsigned long long)n, m);
}

// distance^2, rounded angle (as m = rint(100 deg)) and weight of one ray from d = c - p and the crystal normal
__device__ __forceinline__ double certain_ray_weight(const DeviceScene& sc, const double* __restrict__ table, const D3 d,
                                                     const D3 nrm, double& dd, double& m, unsigned int& n_round) {
  dd = dot(d, d);
  const double dn = dot(d, nrm);
  const double rdd = 1.0 / dd;  // one reciprocal serves the cosine and the solid angle (each within an ulp or two of the
  const double cosang = 2.0 * (dn * dn) * rdd - 1.0;  // quotient the exact kernel forms)
  // m = rint(100 deg(acos(cos))) without the fp64 arc cosine: an FP32 estimate (good to a few 1e-3 of a unit), then the
  // decision in fp64 against the cosines of the two rounding boundaries, cos_bound[m + 1] < cos <= cos_bound[m].  Equal to
  // rounding the fp64 angle except within ~1e-12 of a boundary; rays within 1e-9 of one are counted, as before
  // (a boundary distance of 1e-9 units is at most 1e-9 * pi / 18000 in the cosine).
  const double* __restrict__ cos_bound = table + 18001;
  bool bracketed = false;
  if (cosang >= -1.0 && cosang <= 1.0) {
    int mi = __float2int_rn(acosf((float)cosang) * 5729.5779513082f);
    mi = min(max(mi, 0), 18000);
    double hi = cos_bound[mi], lo = cos_bound[mi + 1];
#pragma unroll 1
    for (int it = 0; it < 4 && !(lo < cosang && cosang <= hi); ++it) {
      mi = cosang > hi ? max(mi - 1, 0) : min(mi + 1, 18000);
      hi = cos_bound[mi], lo = cos_bound[mi + 1];
    }
    if (lo < cosang && cosang <= hi) {
      bracketed = true;
      m = (double)mi;
      if (fmin(hi - cosang, cosang - lo) < 1e-9 * (3.141592653589793 / 18000.0)) ++n_round;
    }
  }
  if (!bracketed) {  // cosine rounded beyond +-1, or an angle at the ends of the table: the formula itself
    const double deg100 = acos(cosang) * (180.0 / 3.141592653589793) * 100.0;
    m = rint(deg100);
    if (fabs(fabs(deg100 - m) - 0.5) < 1e-9) ++n_round;
  }
  const double fraction = (sc.area_pt / (4.0 * 3.141592653589793)) * rdd;
  double prof;
  if (m >= 0.0 && m <= 18000.0) {
    prof = table[(int)m];
  } else {  // NaN angle (cosine rounded beyond 1): what the formula gives
    const double aoi = (180.0 - m / 100.0) / 2.0, da = aoi - sc.aoi0;
    prof = sc.refl_max * exp(-(da * da) / (2.0 * (sc.refl_sigma * sc.refl_sigma)));
  }
  return fraction * prof;
}

__global__ void __launch_bounds__(kWeightThreads) weight_kernel(const __grid_constant__ K1Params P) {
  const DeviceScene& sc = P.sc;
  const ComLattice& L = P.vox.lat;
  const unsigned long long total = P.counters[0] + P.counters[2];
  const unsigned long long n = total > P.cand_capacity ? 0ull : P.counters[0];  // overflow: the exact kernel reports it
  unsigned int n_pass = 0, n_round = 0;
  const double* __restrict__ table = sc.refl_table;
  const unsigned int nz = (unsigned int)L.nz, nx = (unsigned int)L.nx, cpb = L.shard_cols_per_block > 0 ? (unsigned int)L.shard_cols_per_block : 1u;
  const bool sharded = L.shard_count > 1 && L.shard_cols_per_block > 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * kWeightThreads + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * kWeightThreads) {
    const uint2 e = P.cand[i];
    const unsigned int vloc = e.x;
    const int ci = (int)(e.y & kCandCrystalMask);
    if ((long long)vloc >= P.n_vox || ci >= sc.n_crystal) {  // guard (see exact_kernel)
      atomicOr((unsigned long long*)&P.stats->overflow, 2ull);
      continue;
    }
    // voxel position (lattice.cuh: lattice_point), the index arithmetic by multiplication
    const unsigned int f = (unsigned int)P.vox_begin + vloc;
    unsigned int col = fast_div(f, P.magic_nz, nz);
    const unsigned int iz = f - col * nz;
    if (sharded) {
      const unsigned int lb = fast_div(col, P.magic_cpb, cpb);
      col = (lb * (unsigned int)L.shard_count + (unsigned int)L.shard_index) * cpb + (col - lb * cpb);
    }
    const unsigned int iy = fast_div(col, P.magic_nx, nx), ix = col - iy * nx;
    const D3 p = lattice_point(L, ix, iy, iz);
    const D3 c = ld3(sc.crystal_xyz + 3 * ci), nrm = ld3(sc.crystal_normal + 3 * ci);
    // The reference forms plasma = c - d, crystal = p + d, reflected = c + r with d = c - p, r = d - 2 (d.n) n and then
    // v = plasma - crystal (= -d), rc = reflected - crystal (= r), dist = |v|, cos = v.rc / (|v| |rc|)
    // (simulation.py:309-321, 494-497, 516-520).  |r| = |d| and d.r = d.d - 2 (d.n)^2, so
    //     dist^2 = d.d,   cos = 2 (d.n)^2 / (d.d) - 1
    // - the same numbers to a few ulp (the exact kernel keeps the reference's own operation order for every ray it
    // decides).  The rounded angle can differ from the reference's only when deg * 100 is within rounding distance of a
    // half-integer; rays within 1e-9 of one are counted.
    const D3 d = c - p;
    double dd, m;
    const double w = certain_ray_weight(sc, table, d, nrm, dd, m, n_round);
    ++n_pass;
    atomicAdd(&P.w_out[vloc], w);
    if (P.nrays_out) {
      atomicAdd(&P.nrays_out[vloc], 1u);  // result unused: a reduction, nobody waits for the round trip to L2
    }
    if (P.hist_out) {
      const int bin = P.reff_bin[P.vox_begin + vloc];
      if (bin < P.n_bins) atomicAdd(&P.hist_out[bin], w);
    }
    if (P.rays_out) {
      const unsigned long long pos = atomicAdd(P.rays_count, 1ull);
      if ((long long)pos < P.rays_capacity)
        P.rays_out[pos] = ComRayRecord{(P.vox_begin + (long long)vloc) * (long long)sc.n_crystal + ci, sqrt(dd),
                                       (180.0 - m / 100.0) / 2.0, w};
    }
  }
  __shared__ unsigned int s_acc[2];
  if (threadIdx.x < 2) s_acc[threadIdx.x] = 0;
  __syncthreads();
  for (int off = 16; off; off >>= 1) {
    n_pass += __shfl_down_sync(0xffffffffu, n_pass, off);
    n_round += __shfl_down_sync(0xffffffffu, n_round, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (n_pass) atomicAdd(&s_acc[0], n_pass);
    if (n_round) atomicAdd(&s_acc[1], n_round);
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // (visible voxels are counted by count_visible_kernel afterwards)
    if (s_acc[0]) atomicAdd((unsigned long long*)&P.stats->passing_rays, (unsigned long long)s_acc[0]);
    if (s_acc[1]) atomicAdd((unsigned long long*)&P.stats->near_rounding_rays, (unsigned long long)s_acc[1]);
  }
}

// ------------------------------------------------------------------------------------------------------
// K1w on the interval records themselves: no per-ray list is written or read for the certain rays.  A WARP takes 32
// consecutive records; the lane that loaded record q prepares what its rays share - d0 = c - p of the first certain
// voxel (the lattice formula once per record) and the crystal normal - in the warp's slice of shared memory; ray r of
// the warp finds its record by a 5-step binary search over the lanes' scanned lengths (shuffles) and is d0 - j dz,
// j voxels further along the column.  Warp-synchronous: no block barrier in the loop.
__global__ void __launch_bounds__(kWeightThreads) weight_kernel_records(const __grid_constant__ K1Params P) {
  __shared__ unsigned int s_v[kWeightThreads], s_ci[kWeightThreads];
  __shared__ double s_d0[kWeightThreads][3], s_n[kWeightThreads][3];
  __shared__ unsigned int s_acc[2];
  const DeviceScene& sc = P.sc;
  const ComLattice& L = P.vox.lat;
  const unsigned long long reserved = P.counters[5];
  if (reserved > P.rec_capacity) return;  // record overflow: expand_kernel reports it, nothing may be accumulated
  const double* __restrict__ table = sc.refl_table;
  const unsigned int nz = (unsigned int)L.nz, nx = (unsigned int)L.nx, cpb = L.shard_cols_per_block > 0 ? (unsigned int)L.shard_cols_per_block : 1u;
  const bool sharded = L.shard_count > 1 && L.shard_cols_per_block > 0;
  const int lane = threadIdx.x & 31, wbase = (int)threadIdx.x - lane;  // wbase: first shared-memory slot of this warp
  const D3 dz = {P.dz_d[0], P.dz_d[1], P.dz_d[2]};
  unsigned int n_pass = 0, n_round = 0;
  if (threadIdx.x < 2) s_acc[threadIdx.x] = 0;
  const unsigned long long n_warps = (unsigned long long)gridDim.x * (kWeightThreads / 32);
  for (unsigned long long first = ((unsigned long long)blockIdx.x * (kWeightThreads / 32) + (threadIdx.x >> 5)) * 32ull;
       first < reserved; first += n_warps * 32ull) {
    const unsigned long long i = first + lane;
    uint4 rec = make_uint4(0u, 0u, 0u, 0u);
    if (i < reserved) rec = P.recs[i];
    const unsigned int c_len = rec.w;
    unsigned int ic = c_len;  // inclusive scan over the warp's 32 records
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int a = __shfl_up_sync(0xffffffffu, ic, off);
      if (lane >= off) ic += a;
    }
    const unsigned int ex = ic - c_len;  // first ray of this lane's record
    const unsigned int Tw = __shfl_sync(0xffffffffu, ic, 31);
    __syncwarp();  // the previous batch's rays are done with the shared slots
    if (c_len) {  // what the rays of this record share
      const unsigned int vloc = rec.x + (rec.z >> 16);
      const int ci = (int)(rec.y & kCandCrystalMask);
      const unsigned int f = (unsigned int)P.vox_begin + vloc;
      unsigned int col = fast_div(f, P.magic_nz, nz);
      const unsigned int iz = f - col * nz;
      if (sharded) {
        const unsigned int lb = fast_div(col, P.magic_cpb, cpb);
        col = (lb * (unsigned int)L.shard_count + (unsigned int)L.shard_index) * cpb + (col - lb * cpb);
      }
      const unsigned int iy = fast_div(col, P.magic_nx, nx), ix = col - iy * nx;
      const D3 p = lattice_point(L, ix, iy, iz);
      const D3 c = ld3(sc.crystal_xyz + 3 * ci), nrm = ld3(sc.crystal_normal + 3 * ci);
      s_d0[threadIdx.x][0] = c.x - p.x, s_d0[threadIdx.x][1] = c.y - p.y, s_d0[threadIdx.x][2] = c.z - p.z;
      s_n[threadIdx.x][0] = nrm.x, s_n[threadIdx.x][1] = nrm.y, s_n[threadIdx.x][2] = nrm.z;
      s_v[threadIdx.x] = vloc, s_ci[threadIdx.x] = (unsigned int)ci;
    }
    __syncwarp();
    for (unsigned int r0 = 0; r0 < Tw; r0 += 32u) {  // warp-uniform trip count: the shuffles need every lane
      const unsigned int r = r0 + (unsigned int)lane;
      int lo = 0, hi = 31;  // largest q with ex[q] <= r
#pragma unroll
      for (int step = 0; step < 5; ++step) {
        const int mid = (lo + hi + 1) >> 1;
        const unsigned int e = __shfl_sync(0xffffffffu, ex, mid);
        if (e <= r) lo = mid; else hi = mid - 1;
      }
      const unsigned int ex_lo = __shfl_sync(0xffffffffu, ex, lo);  // before any lane leaves: every lane is a source
      if (r >= Tw) continue;
      const int q = wbase + lo;
      const unsigned int j = r - ex_lo;
      const unsigned int vloc = s_v[q] + j;
      if ((long long)vloc >= P.n_vox || (int)s_ci[q] >= sc.n_crystal) {  // guard (see exact_kernel)
        atomicOr((unsigned long long*)&P.stats->overflow, 2ull);
        continue;
      }
      const double jf = (double)j;
      // d = c - (p + j dz): one rounding per component on top of the record's own d0
      const D3 d = {fma(-jf, dz.x, s_d0[q][0]), fma(-jf, dz.y, s_d0[q][1]), fma(-jf, dz.z, s_d0[q][2])};
      const D3 nrm = {s_n[q][0], s_n[q][1], s_n[q][2]};
      double dd, m;
      const double w = certain_ray_weight(sc, table, d, nrm, dd, m, n_round);
      ++n_pass;
      atomicAdd(&P.w_out[vloc], w);
      if (P.nrays_out) atomicAdd(&P.nrays_out[vloc], 1u);
      if (P.hist_out) {
        const int bin = P.reff_bin[P.vox_begin + vloc];
        if (bin < P.n_bins) atomicAdd(&P.hist_out[bin], w);
      }
      if (P.rays_out) {
        const unsigned long long pos = atomicAdd(P.rays_count, 1ull);
        if ((long long)pos < P.rays_capacity)
          P.rays_out[pos] = ComRayRecord{(P.vox_begin + (long long)vloc) * (long long)sc.n_crystal + (long long)s_ci[q], sqrt(dd),
                                         (180.0 - m / 100.0) / 2.0, w};
      }
    }
  }
  for (int off = 16; off; off >>= 1) {
    n_pass += __shfl_down_sync(0xffffffffu, n_pass, off);
    n_round += __shfl_down_sync(0xffffffffu, n_round, off);
  }
  __syncthreads();
  if (lane == 0) {
    if (n_pass) atomicAdd(&s_acc[0], n_pass);
    if (n_round) atomicAdd(&s_acc[1], n_round);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (s_acc[0]) {
      atomicAdd((unsigned long long*)&P.stats->passing_rays, (unsigned long long)s_acc[0]);
      atomicAdd(&P.counters[7], (unsigned long long)s_acc[0]);
    }
    if (s_acc[1]) atomicAdd((unsigned long long*)&P.stats->near_rounding_rays, (unsigned long long)s_acc[1]);
  }
}

// ------------------------------------------------------------------------------------------------------
// Visible voxels of a launch = voxels with a ray count > 0, counted once both fp64 kernels are done (4 B per voxel read;
// the kernels above add to the counts without waiting for the old value).
__global__ void __launch_bounds__(256) count_visible_kernel(const unsigned int* __restrict__ nrays, long long n, ComStats* stats) {
  unsigned int c = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n; i += 4 * stride) {  // four independent coalesced loads in flight
    const unsigned int a = __ldcs(nrays + i), b = __ldcs(nrays + i + stride), d = __ldcs(nrays + i + 2 * stride),
                       e = __ldcs(nrays + i + 3 * stride);
    c += (a != 0u) + (b != 0u) + (d != 0u) + (e != 0u);
  }
  for (; i < n; i += stride) c += __ldcs(nrays + i) != 0u;
  for (int off = 16; off; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
  __shared__ unsigned int s_c;
  if (threadIdx.x == 0) s_c = 0u;
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_c, c);
  __syncthreads();
  if (threadIdx.x == 0 && s_c) atomicAdd((unsigned long long*)&stats->visible_voxels, (unsigned long long)s_c);
}

// ------------------------------------------------------------------------------------------------------
// Optional per-kernel timing (bench.py's roofline leg): events on the launching stream around K1a / K1b.
static thread_local bool g_accumulate_stats = false;
void visibility_accumulate_stats(bool on) { g_accumulate_stats = on; }

static bool g_timing_on = false;
static cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // before K1a | after K1a | after K1w | after K1b
static double g_ms_accum[3] = {0, 0, 0};                            // filter, weight-only, exact
static long long g_ms_launches = 0;
static bool g_ev_pending = false;

static void timing_collect() {
  if (!g_ev_pending) return;
  cudaEventSynchronize(g_ev[3]);
  float a = 0, b = 0, c = 0;
  cudaEventElapsedTime(&a, g_ev[0], g_ev[1]);
  cudaEventElapsedTime(&b, g_ev[1], g_ev[2]);
  cudaEventElapsedTime(&c, g_ev[2], g_ev[3]);
  g_ms_accum[0] += a, g_ms_accum[1] += b, g_ms_accum[2] += c, ++g_ms_launches;
  g_ev_pending = false;
}

void kernel_timing_enable(int on) {
  timing_collect();
  g_timing_on = on != 0;
  if (g_timing_on && !g_ev[0])
    for (auto& e : g_ev) cudaEventCreate(&e);
  g_ms_accum[0] = g_ms_accum[1] = g_ms_accum[2] = 0;
  g_ms_launches = 0;
}

void kernel_timing_read(double* filter_ms, double* exact_ms, long long* launches) {  // exact_ms: both fp64 kernels
  timing_collect();
  if (filter_ms) *filter_ms = g_ms_accum[0];
  if (exact_ms) *exact_ms = g_ms_accum[1] + g_ms_accum[2];
  if (launches) *launches = g_ms_launches;
}

void kernel_timing_read3(double* ms3, long long* launches) {  // filter, weight-only, exact
  timing_collect();
  if (ms3) ms3[0] = g_ms_accum[0], ms3[1] = g_ms_accum[1], ms3[2] = g_ms_accum[2];
  if (launches) *launches = g_ms_launches;
}

long long lattice_local_voxels(const ComLattice& L) {
  const long long cols = L.nx * L.ny;
  if (L.shard_count <= 1 || L.shard_cols_per_block <= 0) return cols * L.nz;
  const long long cpb = L.shard_cols_per_block, blocks = (cols + cpb - 1) / cpb;
  if (L.shard_index < 0 || L.shard_index >= L.shard_count) return 0;
  const long long mine = blocks > L.shard_index ? (blocks - L.shard_index + L.shard_count - 1) / L.shard_count : 0;
  long long local_cols = mine * cpb;
  // the last global block may be partial; it belongs to shard (blocks - 1) % shard_count
  if (mine > 0 && (blocks - 1) % L.shard_count == L.shard_index) local_cols -= blocks * cpb - cols;
  return local_cols * L.nz;
}

static int pick_warps(int n_groups) {  // warps per CTA wasting the fewest idle warps; ties -> larger CTA
  if (n_groups <= kMaxWarps) return n_groups;
  int best = kMaxWarps, best_waste = 1 << 30;
  for (int w = 6; w <= kMaxWarps; ++w) {
    int waste = (n_groups + w - 1) / w * w - n_groups;
    if (waste <= best_waste) best = w, best_waste = waste;
  }
  return best;
}

size_t visibility_workspace_bytes(const DeviceScene& sc, long long n_vox, double frac) {
  if (sc.filter_disabled) frac = 1.0;
  if (frac <= 0) frac = 0.02;
  double rays = (double)n_vox * sc.n_crystal;
  unsigned long long cap = (unsigned long long)(rays * frac) + 1;
  if (!sc.filter_disabled) cap = std::max(cap, 1ull << 20);
  // [256 B counters][ray list: cap x 8 B][interval records: cap x 16 B].  An interval holds at least one ray, so a
  // record list as long as the ray list cannot overflow before the ray list does (coarse lattices: ~1.3 rays per interval).
  return 256 + cap * (sizeof(uint2) + sizeof(uint4));
}

// the split of a workspace of `bytes` bytes into ray list and record list (the inverse of the sizing above)
static void split_workspace(size_t bytes, unsigned long long& cap, unsigned long long& rec_cap) {
  const size_t body = bytes - 256;
  cap = std::max<size_t>(1, body / (sizeof(uint2) + sizeof(uint4)));
  rec_cap = (body - cap * sizeof(uint2)) / sizeof(uint4);
}

int voxels_pack(const ComScene* scene, const double* vox_xyz, long long vox_begin, long long vox_end, float4* out,
                cudaStream_t stream) {
  if (!scene || !vox_xyz || !out || vox_end < vox_begin || vox_begin < 0) {
    set_error("com_voxels_pack: bad arguments");
    return COM_E_BADARG;
  }
  if (vox_end == vox_begin) return COM_OK;
  if ((reinterpret_cast<uintptr_t>(out) & 15u) != 0) {
    set_error("com_voxels_pack: the packed array must be 16-byte aligned");
    return COM_E_BADARG;
  }
  const long long n = vox_end - vox_begin;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148ll * 16);
  voxel_pack_kernel<<<blocks, 256, 0, stream>>>(vox_xyz, vox_begin, vox_end, scene->dev.origin[0], scene->dev.origin[1],
                                                scene->dev.origin[2], out);
  if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
    set_error("voxel_pack_kernel failed: %s", cudaGetErrorString(e));
    return COM_E_CUDA;
  }
  return COM_OK;
}

int visibility_launch_packed(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz, const float4* vox_f4,
                             long long vox_begin, long long vox_end, double* w_out, unsigned int* nrays_out,
                             ComRayRecord* rays_out, long long rays_capacity, unsigned long long* rays_count_out,
                             const uint16_t* reff_bin, double* hist_out, int n_bins, ComStats* stats_out, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream);

int visibility_launch(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz, long long vox_begin,
                      long long vox_end, double* w_out, unsigned int* nrays_out, ComRayRecord* rays_out,
                      long long rays_capacity, unsigned long long* rays_count_out, const uint16_t* reff_bin,
                      double* hist_out, int n_bins, ComStats* stats_out, void* workspace, size_t workspace_bytes,
                      cudaStream_t stream) {
  return visibility_launch_packed(scene, lattice_h, vox_xyz, nullptr, vox_begin, vox_end, w_out, nrays_out, rays_out,
                                  rays_capacity, rays_count_out, reff_bin, hist_out, n_bins, stats_out, workspace,
                                  workspace_bytes, stream);
}

int visibility_launch_packed(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz, const float4* vox_f4,
                             long long vox_begin, long long vox_end, double* w_out, unsigned int* nrays_out,
                             ComRayRecord* rays_out, long long rays_capacity, unsigned long long* rays_count_out,
                             const uint16_t* reff_bin, double* hist_out, int n_bins, ComStats* stats_out, void* workspace,
                             size_t workspace_bytes, cudaStream_t stream) {
  // an empty range has nothing to write: the output pointers may then be null (empty device tensors are)
  const bool empty = vox_end == vox_begin;
  if (!scene || (!lattice_h && !vox_xyz && !empty) || vox_end < vox_begin || (!w_out && !empty) || !stats_out ||
      !workspace || (rays_out && !rays_count_out) || (hist_out && !reff_bin)) {
    set_error("com_visibility_weights: bad arguments");
    return COM_E_BADARG;
  }
  const long long n_vox = vox_end - vox_begin;
  if (n_vox >= (1ll << 32)) {
    set_error("com_visibility_weights: at most 2^32-1 voxels per call (got %lld); split the range", n_vox);
    return COM_E_BADARG;
  }
  if (workspace_bytes < 256 + 64) {
    set_error("workspace too small");
    return COM_E_NOMEM;
  }
  K1Params P{};
  P.sc = scene->dev;
  P.vox.use_lattice = lattice_h ? 1 : 0;
  if (lattice_h) {
    P.vox.lat = *lattice_h;
    if (vox_end > lattice_local_voxels(*lattice_h) || vox_begin < 0) {
      set_error("voxel range [%lld,%lld) outside the lattice (shard)", vox_begin, vox_end);
      return COM_E_BADARG;
    }
  }
  P.vox.xyz = vox_xyz;
  P.vox.f4 = lattice_h ? nullptr : vox_f4;
  if (P.vox.f4 && (reinterpret_cast<uintptr_t>(P.vox.f4) & 15u) != 0) {
    set_error("com_visibility_weights_packed: the packed voxel array must be 16-byte aligned");
    return COM_E_BADARG;
  }
  P.vox_begin = vox_begin;
  P.n_vox = n_vox;
  // the lattice kernel indexes voxels with 32 bits
  const bool lattice_kernel = lattice_h != nullptr && lattice_h->nz >= 1 &&
                              lattice_h->nx * lattice_h->ny * lattice_h->nz < (1ll << 32);
  const int n_groups = P.sc.n_crystal_padded / 32;
  const bool interval_kernel = lattice_kernel && P.sc.interval_filter != 0 && lattice_h->nz < 65536;  // 16-bit lengths
  if (interval_kernel) {
    for (int i = 0; i < 3; ++i) P.dz[i] = (float)(lattice_h->step[2] * lattice_h->rotation[6 + i]);
    const long long nz = lattice_h->nz;
    const long long n_cols = n_vox > 0 ? (vox_end - 1) / nz - vox_begin / nz + 1 : 0;
    P.n_tiles = (int)((n_cols + kColsPerItem - 1) / kColsPerItem);  // column blocks
    P.tile_voxels = 0;
    P.warps_per_cta = kLatticeWarps;
    P.n_super = n_groups;
  } else if (lattice_kernel) {
    // recentred position step along z: step_z * (row 2 of R), mesh_calculation.py:150
    for (int i = 0; i < 3; ++i) P.dz[i] = (float)(lattice_h->step[2] * lattice_h->rotation[6 + i]);
    const long long nz = lattice_h->nz, cpc = (nz + kSegmentMax - 1) / kSegmentMax;
    const long long seg_len = (nz + cpc - 1) / cpc;  // typical segment length
    P.segs_per_item = (int)std::max<long long>(1, std::min<long long>(kSegsPerItemMax, 256 / seg_len));
    const long long colA = vox_begin / nz, gA = colA * cpc + (vox_begin - colA * nz) / kSegmentMax;
    const long long colB = (vox_end - 1) / nz, gB = colB * cpc + ((vox_end - 1) - colB * nz) / kSegmentMax;
    const long long n_segments = n_vox > 0 ? gB - gA + 1 : 0;
    P.n_tiles = (int)((n_segments + P.segs_per_item - 1) / P.segs_per_item);  // segment groups
    P.tile_voxels = 0;
    P.warps_per_cta = kLatticeWarps;
    P.n_super = n_groups;
  } else {
    P.tile_voxels = kTileVoxels;
    P.n_tiles = (int)((n_vox + P.tile_voxels - 1) / P.tile_voxels);
    P.warps_per_cta = pick_warps(n_groups);
    P.n_super = (n_groups + P.warps_per_cta - 1) / P.warps_per_cta;
  }
  P.counters = reinterpret_cast<unsigned long long*>(workspace);
  P.cand = reinterpret_cast<uint2*>(static_cast<char*>(workspace) + 256);
  split_workspace(workspace_bytes, P.cand_capacity, P.rec_capacity);
  P.recs = reinterpret_cast<uint4*>(static_cast<char*>(workspace) + 256 + ((P.cand_capacity * sizeof(uint2) + 15) / 16) * 16);
  if (reinterpret_cast<char*>(P.recs + P.rec_capacity) > static_cast<char*>(workspace) + workspace_bytes) --P.rec_capacity;
  P.w_out = w_out;
  P.nrays_out = nrays_out;
  P.rays_out = rays_out;
  P.rays_capacity = rays_capacity;
  P.rays_count = rays_count_out;
  P.reff_bin = reff_bin;
  P.hist_out = hist_out;
  P.n_bins = n_bins;
  P.stats = stats_out;
  P.accumulate_stats = g_accumulate_stats ? 1 : 0;

  cudaError_t e;
#define COM_CUDA(x)                                                        \
  if ((e = (x)) != cudaSuccess) {                                          \
    set_error("%s failed: %s", #x, cudaGetErrorString(e));                 \
    return COM_E_CUDA;                                                     \
  }
  int dev = 0, sms = 0;
  COM_CUDA(cudaGetDevice(&dev));
  if (scene->device >= 0 && scene->device != dev) {
    set_error("com_visibility_weights: the scene lives on device %d, the current device is %d", scene->device, dev);
    return COM_E_BADARG;
  }
  COM_CUDA(cudaMemsetAsync(workspace, 0, 256, stream));
  if (!g_accumulate_stats) COM_CUDA(cudaMemsetAsync(stats_out, 0, sizeof(ComStats), stream));
  if (n_vox) COM_CUDA(cudaMemsetAsync(w_out, 0, sizeof(double) * (size_t)n_vox, stream));
  if (nrays_out && n_vox) COM_CUDA(cudaMemsetAsync(nrays_out, 0, sizeof(unsigned int) * (size_t)n_vox, stream));
  if (rays_count_out) COM_CUDA(cudaMemsetAsync(rays_count_out, 0, sizeof(unsigned long long), stream));
  if (n_vox == 0) return COM_OK;

  COM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int threads = P.warps_per_cta * 32;
  int occ = 0;
  if (interval_kernel) {
    COM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, interval_kernel_lattice, threads, 0));
  } else if (lattice_kernel) {
    COM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_kernel_lattice, threads, 0));
  } else {
    COM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, filter_kernel, threads, 0));
  }
  if (occ < 1) occ = 1;
  const long long items = (long long)P.n_tiles * P.n_super;  // lattice kernel: warp-level items
  const long long ctas_wanted = lattice_kernel ? (items + kLatticeWarps - 1) / kLatticeWarps : items;
  const int grid = (int)std::max<long long>(1, std::min<long long>(ctas_wanted, (long long)sms * occ));
  // the weight-only kernel takes the front region of the list when the filter can vouch for whole rays: the interval
  // kernel always, the per-ray lattice kernel when the scene has certain regions for every aperture (the detector image
  // needs the fp64 detector hit: everything then goes through the exact kernel)
  P.front_weight_only = (P.sc.pix_out == nullptr && !P.sc.filter_disabled &&
                         (interval_kernel || (lattice_kernel && !P.sc.certain_off && P.sc.n_port_edge > 0 &&
                                              P.sc.det_inner.xlo <= P.sc.det_inner.xhi && P.sc.inner_crys.xlo <= P.sc.inner_crys.xhi)))
                            ? 1 : 0;
  if (g_timing_on) {
    timing_collect();
    cudaEventRecord(g_ev[0], stream);
  }
  P.records_to_weight = (interval_kernel && P.front_weight_only) ? 1 : 0;
  if (lattice_h)
    for (int i = 0; i < 3; ++i) P.dz_d[i] = lattice_h->step[2] * lattice_h->rotation[6 + i];
  if (P.front_weight_only) {
    auto magic = [](long long d) { return d >= 2 ? ~0ull / (unsigned long long)d + 1ull : 0ull; };
    P.magic_nz = magic(lattice_h->nz), P.magic_nx = magic(lattice_h->nx);
    P.magic_cpb = magic(std::max<long long>(1, lattice_h->shard_cols_per_block));
  }
  if (interval_kernel) {
    interval_kernel_lattice<<<grid, threads, 0, stream>>>(P);
    // interval records -> per-ray list (grid-stride over batches of 256 records; the count is only known on the device)
    const unsigned long long batches = (std::min<unsigned long long>(P.rec_capacity, (unsigned long long)n_vox * P.sc.n_crystal) +
                                        kExpandThreads - 1) / kExpandThreads;
    expand_kernel<<<(unsigned int)std::max<unsigned long long>(1, std::min<unsigned long long>(batches, (unsigned long long)sms * 8)),
                    kExpandThreads, 0, stream>>>(P);
  } else if (lattice_kernel) {
    filter_kernel_lattice<<<grid, threads, 0, stream>>>(P);
  }
  else
    filter_kernel<<<grid, threads, 0, stream>>>(P);
  COM_CUDA(cudaGetLastError());
  if (g_timing_on) cudaEventRecord(g_ev[1], stream);
  {
    // grid-stride over the candidate list: at most a few blocks per SM, whatever the list's capacity (a grid sized by
    // the capacity costs ~4 ns per empty block: 13 ms per launch of 6.7e7 voxels at the default 2 % capacity)
    const size_t smem = sizeof(DevAperture) * P.sc.n_apertures + (sizeof(double) + sizeof(float)) * 3 * P.sc.n_edges_total;
    int occ_x = 0;
    COM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_x, exact_kernel, kExactThreads, smem));
    if (occ_x < 1) occ_x = 1;
    const unsigned long long slots = std::min<unsigned long long>(P.cand_capacity, (unsigned long long)n_vox * P.sc.n_crystal);
    const unsigned int blocks = (unsigned int)std::max<unsigned long long>(
        1, std::min<unsigned long long>((slots + kExactThreads - 1) / kExactThreads, (unsigned long long)sms * occ_x * 4ull));
    if (P.records_to_weight) {
      int occ_w = 0;
      COM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_w, weight_kernel_records, kWeightThreads, 0));
      if (occ_w < 1) occ_w = 1;
      const unsigned long long batches = (std::min<unsigned long long>(P.rec_capacity, slots) + kWeightThreads - 1) / kWeightThreads;
      weight_kernel_records<<<(unsigned int)std::max<unsigned long long>(1, std::min<unsigned long long>(batches, (unsigned long long)sms * occ_w * 2ull)),
                              kWeightThreads, 0, stream>>>(P);
    } else if (P.front_weight_only) {
      int occ_w = 0;
      COM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_w, weight_kernel, kWeightThreads, 0));
      if (occ_w < 1) occ_w = 1;
      const unsigned int wblocks = (unsigned int)std::max<unsigned long long>(
          1, std::min<unsigned long long>((slots + kWeightThreads - 1) / kWeightThreads, (unsigned long long)sms * occ_w * 4ull));
      weight_kernel<<<wblocks, kWeightThreads, 0, stream>>>(P);
    }
    if (g_timing_on) cudaEventRecord(g_ev[2], stream);
    exact_kernel<<<blocks, kExactThreads, smem, stream>>>(P);
    if (nrays_out)
      count_visible_kernel<<<(unsigned int)std::max<long long>(1, std::min<long long>((n_vox + 1023) / 1024, (long long)sms * 8)),
                             256, 0, stream>>>(nrays_out, n_vox, stats_out);
  }
  COM_CUDA(cudaGetLastError());
  if (g_timing_on) {
    cudaEventRecord(g_ev[3], stream);
    g_ev_pending = true;
  }
#undef COM_CUDA
  return COM_OK;
}

}  // namespace com
