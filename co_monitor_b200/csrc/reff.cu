// Reff provider on the device: effective radius of every voxel of an implicit (possibly sharded) lattice, as the
// uint16 millimetre bins Stage 2 reads, without a file and without the VMEC web service.
//
// The reference asks a W7-X REST service for Reff per voxel and stores the answer rounded to three decimals
// (co_monitor/geometry/reff_VMEC.py:85-136, 138-176); there is no network here and the 10 mm file is a missing blob.
// Two replacements:
//   reff_resample_kernel   trilinear resampling of a shipped coarse grid onto a finer lattice spanning the same cuboid
//                          (fractional lattice-index space; a stencil that touches NaN with non-zero weight is NaN),
//                          the same operations in the same order as co_monitor_b200/geometry/reff.py::interpolate_reff
//   reff_quadratic_kernel  analytic nested flux surfaces Reff^2 = quadratic form of the position (QuadraticReff)
// Both are one pass writing 2 B (+8 B optional) per voxel: HBM-write bound; the coarse grid (< 1 MB) stays in L1/L2.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "lattice.cuh"
#include "scene.h"

namespace com {

constexpr int kReffThreads = 256;

struct ReffResampleParams {
  ComLattice lat;
  long long begin, n;
  const double* src;  // (sny, snx, snz) row-major, z fastest: the reshape of the reference's flat order
  long long snx, sny, snz;
  double scale[3];  // (n_s - 1) / (n_t - 1) per axis x, y, z
  uint16_t* mm_out;
  double* reff_out;
};

__device__ __forceinline__ uint16_t mm_bin(double reff_m, double* rounded) {
  // np.round(x, 3) == rint(x * 1000) / 1000 (reff_VMEC.py:157); the bin is the integer in the middle
  if (isnan(reff_m)) {
    if (rounded) *rounded = reff_m;
    return 0xFFFFu;
  }
  double k = rint(__dmul_rn(reff_m, 1000.0));
  if (rounded) *rounded = k / 1000.0;
  k = fmin(fmax(k, 0.0), 65534.0);
  return (uint16_t)k;
}

__device__ __forceinline__ void axis_stencil(long long i, double scale, long long n_s, long long& i0, double& frac) {
  const double f = __dmul_rn((double)i, scale);
  i0 = (long long)floor(f);
  if (i0 > n_s - 2) i0 = n_s - 2;
  frac = __dadd_rn(f, -(double)i0);
}

__global__ void __launch_bounds__(kReffThreads) reff_resample_kernel(const __grid_constant__ ReffResampleParams P) {
  const long long v = (long long)blockIdx.x * kReffThreads + threadIdx.x;
  if (v >= P.n) return;
  long long ix, iy, iz;
  lattice_unravel(P.lat, P.begin + v, ix, iy, iz);
  long long x0, y0, z0;
  double fx, fy, fz;
  axis_stencil(ix, P.scale[0], P.snx, x0, fx);
  axis_stencil(iy, P.scale[1], P.sny, y0, fy);
  axis_stencil(iz, P.scale[2], P.snz, z0, fz);
  const double wy[2] = {__dadd_rn(1.0, -fy), fy}, wx[2] = {__dadd_rn(1.0, -fx), fx}, wz[2] = {__dadd_rn(1.0, -fz), fz};
  double acc = 0.0;
  bool bad = false;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dz = 0; dz < 2; ++dz) {
        const double w = __dmul_rn(__dmul_rn(wy[dy], wx[dx]), wz[dz]);
        const double s = __ldg(P.src + ((y0 + dy) * P.snx + (x0 + dx)) * P.snz + (z0 + dz));
        const bool nan = isnan(s);
        bad |= nan && w > 0.0;
        acc = __dadd_rn(acc, __dmul_rn(nan ? 0.0 : s, w));
      }
  double rounded;
  const uint16_t bin = mm_bin(bad ? nan("") : acc, &rounded);
  P.mm_out[v] = bin;
  if (P.reff_out) P.reff_out[v] = rounded;
}

struct ReffQuadraticParams {
  ComLattice lat;
  long long begin, n;
  double c[10];
  double reff_max;
  uint16_t* mm_out;
  double* reff_out;
};

__global__ void __launch_bounds__(kReffThreads) reff_quadratic_kernel(const __grid_constant__ ReffQuadraticParams P) {
  const long long v = (long long)blockIdx.x * kReffThreads + threadIdx.x;
  if (v >= P.n) return;
  const D3 p = lattice_position(P.lat, P.begin + v);
  const double x = p.x / 1000.0, y = p.y / 1000.0, z = p.z / 1000.0;
  // left-to-right like the NumPy expression of QuadraticReff.__call__
  const double* c = P.c;
  double q = __dmul_rn(__dmul_rn(c[0], x), x);
  q = __dadd_rn(q, __dmul_rn(__dmul_rn(c[1], y), y));
  q = __dadd_rn(q, __dmul_rn(__dmul_rn(c[2], z), z));
  q = __dadd_rn(q, __dmul_rn(__dmul_rn(c[3], x), y));
  q = __dadd_rn(q, __dmul_rn(__dmul_rn(c[4], x), z));
  q = __dadd_rn(q, __dmul_rn(__dmul_rn(c[5], y), z));
  q = __dadd_rn(q, __dmul_rn(c[6], x));
  q = __dadd_rn(q, __dmul_rn(c[7], y));
  q = __dadd_rn(q, __dmul_rn(c[8], z));
  q = __dadd_rn(q, c[9]);
  const double r = sqrt(fmax(q, 0.0));
  double rounded;
  const uint16_t bin = mm_bin(r <= P.reff_max ? r : nan(""), &rounded);
  P.mm_out[v] = bin;
  if (P.reff_out) P.reff_out[v] = rounded;
}

static int check_range(const char* who, const ComLattice* L, long long begin, long long end, const void* out) {
  if (!L || (!out && end != begin) || end < begin || begin < 0) {  // empty device tensors have null pointers
    set_error("%s: bad arguments", who);
    return COM_E_BADARG;
  }
  if (end > lattice_local_voxels(*L)) {
    set_error("%s: voxel range [%lld,%lld) outside the lattice (shard)", who, begin, end);
    return COM_E_BADARG;
  }
  if (end - begin > (long long)0x7fffffff * kReffThreads) {
    set_error("%s: range too long for one launch", who);
    return COM_E_BADARG;
  }
  return COM_OK;
}

}  // namespace com

extern "C" {

int com_reff_resample(const ComLattice* target_h, int64_t vox_begin, int64_t vox_end, const double* src_reff,
                      int64_t src_nx, int64_t src_ny, int64_t src_nz, uint16_t* reff_mm_out, double* reff_out,
                      com_stream_t stream) {
  using namespace com;
  if (int rc = check_range("com_reff_resample", target_h, vox_begin, vox_end, reff_mm_out)) return rc;
  if (!src_reff || src_nx < 2 || src_ny < 2 || src_nz < 2 || target_h->nx < 2 || target_h->ny < 2 || target_h->nz < 2) {
    set_error("com_reff_resample: source and target lattices need at least two samples per axis");
    return COM_E_BADARG;
  }
  ReffResampleParams P{};
  P.lat = *target_h;
  P.begin = vox_begin;
  P.n = vox_end - vox_begin;
  P.src = src_reff;
  P.snx = src_nx, P.sny = src_ny, P.snz = src_nz;
  P.scale[0] = (double)(src_nx - 1) / (double)(target_h->nx - 1);
  P.scale[1] = (double)(src_ny - 1) / (double)(target_h->ny - 1);
  P.scale[2] = (double)(src_nz - 1) / (double)(target_h->nz - 1);
  P.mm_out = reff_mm_out;
  P.reff_out = reff_out;
  if (P.n == 0) return COM_OK;
  const unsigned int blocks = (unsigned int)((P.n + kReffThreads - 1) / kReffThreads);
  reff_resample_kernel<<<blocks, kReffThreads, 0, static_cast<cudaStream_t>(stream)>>>(P);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("reff_resample_kernel launch failed: %s", cudaGetErrorString(e));
    return COM_E_CUDA;
  }
  return COM_OK;
}

int com_reff_quadratic(const ComLattice* target_h, int64_t vox_begin, int64_t vox_end, const double* coef_h,
                       double reff_max, uint16_t* reff_mm_out, double* reff_out, com_stream_t stream) {
  using namespace com;
  if (int rc = check_range("com_reff_quadratic", target_h, vox_begin, vox_end, reff_mm_out)) return rc;
  if (!coef_h) {
    set_error("com_reff_quadratic: null coefficients");
    return COM_E_BADARG;
  }
  ReffQuadraticParams P{};
  P.lat = *target_h;
  P.begin = vox_begin;
  P.n = vox_end - vox_begin;
  for (int i = 0; i < 10; ++i) P.c[i] = coef_h[i];
  P.reff_max = reff_max;
  P.mm_out = reff_mm_out;
  P.reff_out = reff_out;
  if (P.n == 0) return COM_OK;
  const unsigned int blocks = (unsigned int)((P.n + kReffThreads - 1) / kReffThreads);
  reff_quadratic_kernel<<<blocks, kReffThreads, 0, static_cast<cudaStream_t>(stream)>>>(P);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("reff_quadratic_kernel launch failed: %s", cudaGetErrorString(e));
    return COM_E_CUDA;
  }
  return COM_OK;
}

}  // extern "C"
