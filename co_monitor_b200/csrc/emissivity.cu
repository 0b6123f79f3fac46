// Stage 2 on sm_100a: line-emissivity integration.
//
// Reference semantics (co_monitor/emissivity/reader.py):
//   :199-243  nearest profile row per voxel: argmin_k |Reff_k - Reff_v| (first minimum), boundary cut Reff >= b dropped,
//             b = min(max mapped Reff, max profile Reff); NaN Reff dropped (:176)
//   :282-314  nearest fractional-abundance temperature on arange(5, 20000, 1)
//   :316-341  per transition: nearest index on the 2000-point linspace ne grid and Te grid, PEC looked up there
//             (pec.py:153-187: the grid values are the bilinear interpolant of the ADAS nodes)
//   :343-378  eps_t = FA * pec_t * ne^2 * c   (FA = "Fully stripped" column for RECOM, the ion's column otherwise)
//   :380-401  flux_t = sum_v eps_t(v) * w_v,   TOTAL = sum_v (sum_t eps_t(v)) * w_v
//
// Kernels
//   K2a reff_max_kernel        max of the finite mapped Reff (for the boundary)
//   K2  emissivity_kernel      one voxel per thread: three exact nearest-neighbour searches (binary search + the
//                              reference's first-minimum tie rule), on-the-fly bilinear PEC, fp64 products, block
//                              reduction of the flux partials; the last block to finish adds the partials in block
//                              order, so the result does not depend on scheduling.
//   K2h weight_histogram_kernel  voxel pass for large grids / profile sweeps: w_v accumulated per profile row
//                              - the only per-voxel work a sweep needs
//       weight_histogram_mm_kernel + histogram_rows_kernel  the same in two halves for uint16 millimetre Reff (10 B/voxel
//                              of HBM traffic): weights binned by millimetre in shared memory, then the Reff cut and the
//                              millimetre -> profile-row mapping on the 4096-bin histogram (which ranks can all-reduce)
//   K2s sweep_kernel           flux[s][t] = sum_k H[k] * eps_t(s, k) for a batch of profile time slices
//   Extension (off by default): log-log bilinear PEC interpolation, pec_loglog.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "scene.h"

namespace com {

constexpr int kMaxTransitions = 4;

struct DevPec {
  int n_ne, n_te, n_grid, use_stripped, loglog;
  const double* ne_nodes;
  const double* te_nodes;
  const double* table;  // (n_ne, n_te) row-major
  const double* ne_grid;
  const double* te_grid;
};

struct DevTables {
  int n_fa, n_trans;
  const double* fa_T;
  const double* fa_ion;
  const double* fa_last;
  DevPec pec[kMaxTransitions];
};

// argmin_k |t[k*STRIDE] - v| with numpy's first-minimum rule, for a non-decreasing table.
// All tables of this path are (nearly) uniform grids - arange(0, 0.539, 0.001), arange(5, 20000, 1), linspace(.., 2000)
// - so the search starts from the index a uniform grid would give and walks to the minimum of |t - v|, which is
// unimodal along a sorted table; a few steps at most.  If the walk does not settle (non-uniform table) it falls back
// to a binary search.  Either way the result is exactly numpy's argmin.
template <int STRIDE>
__device__ __forceinline__ int nearest_first_sorted(const double* __restrict__ t, int n, double v) {
  if (n > 4) {
    const double t0 = t[0], t1 = t[(size_t)(n - 1) * STRIDE];
    if (t1 > t0) {
      const double g = (v - t0) * ((double)(n - 1) / (t1 - t0));
      int i = g <= 0.0 ? 0 : (g >= (double)(n - 1) ? n - 1 : (int)(g + 0.5));
      double d = fabs(t[(size_t)i * STRIDE] - v);
      int moves = 0;
      bool settled = true;
      while (i + 1 < n) {  // towards larger values while not farther
        const double dn = fabs(t[(size_t)(i + 1) * STRIDE] - v);
        if (!(dn <= d)) break;
        if (++moves > 6) { settled = false; break; }
        ++i, d = dn;
      }
      while (settled && i > 0) {  // back to the first index attaining the minimum
        const double dp = fabs(t[(size_t)(i - 1) * STRIDE] - v);
        if (!(dp <= d)) break;
        if (++moves > 12) { settled = false; break; }
        --i, d = dp;
      }
      if (settled) return i;
    }
  }
  int lo = 0, hi = n;  // lower_bound: first index with t[i] >= v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (t[(size_t)mid * STRIDE] < v) lo = mid + 1; else hi = mid;
  }
  if (lo == 0) return 0;
  int left = lo - 1;
  const double dl = fabs(t[(size_t)left * STRIDE] - v);
  if (lo < n) {
    const double dr = fabs(t[(size_t)lo * STRIDE] - v);
    if (dr < dl) return lo;
  }
  // first index that attains dl: equal table values, or differences that round to the same double
  while (left > 0 && fabs(t[(size_t)(left - 1) * STRIDE] - v) == dl) --left;
  return left;
}

// Same result for an arbitrary table (the reference never sorts): linear scan.
template <int STRIDE>
__device__ __forceinline__ int nearest_first_scan(const double* __restrict__ t, int n, double v) {
  int best = 0;
  double bd = fabs(t[0] - v);
  for (int k = 1; k < n; ++k) {
    const double d = fabs(t[(size_t)k * STRIDE] - v);
    if (d < bd) bd = d, best = k;
  }
  return best;
}

// Degree-1 tensor spline through the ADAS nodes at (x, y), evaluated in FITPACK's order (fpbspl + fpbisp).
__device__ __forceinline__ double pec_bilinear(const DevPec& p, double x, double y) {
  auto cell = [](const double* nodes, int n, double q) {
    int lo = 0, hi = n;  // upper_bound
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (nodes[mid] <= q) lo = mid + 1; else hi = mid;
    }
    return min(max(lo - 1, 0), n - 2);
  };
  const int i = cell(p.ne_nodes, p.n_ne, x), j = cell(p.te_nodes, p.n_te, y);
  const double x0 = p.ne_nodes[i], x1 = p.ne_nodes[i + 1], y0 = p.te_nodes[j], y1 = p.te_nodes[j + 1];
  const double fx = 1.0 / (x1 - x0), fy = 1.0 / (y1 - y0);
  const double hx0 = __dmul_rn(fx, x1 - x), hx1 = __dmul_rn(fx, x - x0);
  const double hy0 = __dmul_rn(fy, y1 - y), hy1 = __dmul_rn(fy, y - y0);
  const double* c = p.table + (size_t)i * p.n_te + j;
  double sp = __dmul_rn(__dmul_rn(c[0], hx0), hy0);
  sp = __dadd_rn(sp, __dmul_rn(__dmul_rn(c[1], hx0), hy1));
  sp = __dadd_rn(sp, __dmul_rn(__dmul_rn(c[p.n_te], hx1), hy0));
  sp = __dadd_rn(sp, __dmul_rn(__dmul_rn(c[p.n_te + 1], hx1), hy1));
  return sp;
}

// Extension (COM_PEC_LOGLOG, not what the reference computes): bilinear interpolation of log PEC in (log ne, log Te)
// through the ADAS nodes at the row's own (ne, Te), arguments clamped to the node range.  A cell with a non-positive
// PEC value falls back to the linear form.
__device__ __forceinline__ double pec_loglog(const DevPec& p, double ne, double te) {
  const double x = fmin(fmax(ne, p.ne_nodes[0]), p.ne_nodes[p.n_ne - 1]);
  const double y = fmin(fmax(te, p.te_nodes[0]), p.te_nodes[p.n_te - 1]);
  auto cell = [](const double* nodes, int n, double q) {
    int lo = 0, hi = n;  // upper_bound
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (nodes[mid] <= q) lo = mid + 1; else hi = mid;
    }
    return min(max(lo - 1, 0), n - 2);
  };
  const int i = cell(p.ne_nodes, p.n_ne, x), j = cell(p.te_nodes, p.n_te, y);
  const double* c = p.table + (size_t)i * p.n_te + j;
  const double c00 = c[0], c01 = c[1], c10 = c[p.n_te], c11 = c[p.n_te + 1];
  if (!(c00 > 0.0 && c01 > 0.0 && c10 > 0.0 && c11 > 0.0 && p.ne_nodes[i] > 0.0 && p.te_nodes[j] > 0.0))
    return pec_bilinear(p, x, y);
  const double lx0 = log(p.ne_nodes[i]), lx1 = log(p.ne_nodes[i + 1]), ly0 = log(p.te_nodes[j]), ly1 = log(p.te_nodes[j + 1]);
  const double u = (log(x) - lx0) / (lx1 - lx0), v = (log(y) - ly0) / (ly1 - ly0);
  const double L = (1.0 - u) * (1.0 - v) * log(c00) + (1.0 - u) * v * log(c01) + u * (1.0 - v) * log(c10) + u * v * log(c11);
  return exp(L);
}

struct RowEval {
  double fa_ion, fa_last;
  double pec[kMaxTransitions];
  double emis[kMaxTransitions];
  double total;
};

// Everything that depends only on (ne, Te): reader.py:282-378 for one profile row.
__device__ __forceinline__ void eval_row(const DevTables& T, double ne, double te, double conc, RowEval& r) {
  const int j = nearest_first_sorted<1>(T.fa_T, T.n_fa, te);
  r.fa_ion = T.fa_ion[j];
  r.fa_last = T.fa_last[j];
  r.total = 0.0;
  const double ne2 = __dmul_rn(ne, ne);
  for (int t = 0; t < T.n_trans; ++t) {
    const DevPec& p = T.pec[t];
    if (p.loglog) {
      r.pec[t] = pec_loglog(p, ne, te);
    } else {
      const int ie = nearest_first_sorted<1>(p.ne_grid, p.n_grid, ne);
      const int it = nearest_first_sorted<1>(p.te_grid, p.n_grid, te);
      r.pec[t] = pec_bilinear(p, p.ne_grid[ie], p.te_grid[it]);
    }
    const double fa = p.use_stripped ? r.fa_last : r.fa_ion;
    r.emis[t] = __dmul_rn(__dmul_rn(__dmul_rn(fa, r.pec[t]), ne2), conc);
    r.total = t == 0 ? r.emis[0] : __dadd_rn(r.total, r.emis[t]);
  }
}

struct K2Params {
  DevTables tab;
  const double* profile;  // (K,3): Reff, T_e, n_e
  int K;
  int profile_sorted;
  double profile_reff_max;
  double conc;
  const double* reff;       // (n) or null
  const uint16_t* reff_mm;  // (n) or null: round(Reff * 1000), 0xFFFF = NaN
  const long long* reff_index;  // null, or (n): the Reff of voxel i is reff[reff_index[i]]
  const double* w;
  long long n;
  const unsigned long long* n_device;  // null, or device count overriding n (n is then the upper bound)
  const double* mapped_max;  // null, or device scalar: max mapped Reff already reduced (over ranks) by the caller
  double* flux_out;          // (n_trans + 1)
  double* per_voxel;         // (n, 4 + 2*n_trans + 1) or null
  unsigned char* valid_out;  // (n) or null
  double* hist_out;          // (K) or null (histogram kernels)
  double* boundary_out;      // null, or where histogram_rows_kernel leaves the Reff cut it applied
  double* scratch;           // [0] max mapped reff, [1] boundary, partials from +8
  unsigned int* done_counter;
};

__device__ __forceinline__ double voxel_reff(const K2Params& P, long long i) {
  const long long j = P.reff_index ? P.reff_index[i] : i;
  if (P.reff) return P.reff[j];
  const uint16_t b = P.reff_mm[j];
  return b == 0xFFFF ? nan("") : (double)b / 1000.0;
}

__device__ __forceinline__ long long voxel_count(const K2Params& P) {
  return P.n_device ? (long long)min((unsigned long long)P.n, *P.n_device) : P.n;
}

__device__ __forceinline__ double cut_boundary(const K2Params& P) {  // reader.py:210-217
  return fmin(P.mapped_max ? *P.mapped_max : P.scratch[0], P.profile_reff_max);
}

__device__ __forceinline__ int profile_row(const K2Params& P, double reff) {
  return P.profile_sorted ? nearest_first_sorted<3>(P.profile, P.K, reff) : nearest_first_scan<3>(P.profile, P.K, reff);
}

__global__ void reff_max_kernel(const K2Params P) {
  double m = -1.0;
  const long long n = voxel_count(P);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double r = voxel_reff(P, i);
    if (r == r) m = fmax(m, r);
  }
  for (int off = 16; off; off >>= 1) m = fmax(m, __shfl_down_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m >= 0.0)  // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long*>(&P.scratch[0]), (unsigned long long)__double_as_longlong(m));
}

constexpr int kK2Threads = 256;

// kRowTable: everything a voxel's columns hold depends on its profile row alone (reader.py:282-378 work on n_e and T_e
// of the row), so the block first evaluates every row once into shared memory - the profile, the fractional-abundance
// grid and the PEC nodes are then touched K times per block instead of once per voxel - and the voxel loop is a row
// search, a shared-memory gather and the weighted sums.  The same eval_row on the same inputs: identical numbers.
// Used when the K x (4 + 2 T + 1) table fits in shared memory (K <= ~1300 rows); longer profiles evaluate per voxel.
template <bool kRowTable>
__global__ void __launch_bounds__(kK2Threads) emissivity_kernel(const K2Params P) {
  extern __shared__ double s_rowtab[];  // kRowTable: (K, ncol) = n_e, T_e, FA ion, FA stripped, pec[T], emis[T], total
  const int nt = P.tab.n_trans, nout = nt + 1, ncol = 4 + 2 * nt + 1;
  const double boundary = cut_boundary(P);
  const long long n = voxel_count(P);
  double acc[kMaxTransitions + 1];
#pragma unroll
  for (int t = 0; t <= kMaxTransitions; ++t) acc[t] = 0.0;
  if (kRowTable) {
    for (int k = threadIdx.x; k < P.K; k += kK2Threads) {
      const double te = P.profile[3 * k + 1], ne = P.profile[3 * k + 2];
      RowEval r;
      eval_row(P.tab, ne, te, P.conc, r);
      double* o = s_rowtab + (size_t)k * ncol;
      o[0] = ne, o[1] = te, o[2] = r.fa_ion, o[3] = r.fa_last;
      for (int t = 0; t < nt; ++t) o[4 + t] = r.pec[t], o[4 + nt + t] = r.emis[t];
      o[4 + 2 * nt] = r.total;
    }
    __syncthreads();
  }

  for (long long i = blockIdx.x * (long long)kK2Threads + threadIdx.x; i < n; i += (long long)gridDim.x * kK2Threads) {
    const double reff = voxel_reff(P, i);
    const bool valid = (reff == reff) && !(reff >= boundary);  // dropna (:176), boundary cut (:240-242)
    if (P.valid_out) P.valid_out[i] = valid ? 1 : 0;
    if (!valid) continue;
    const int k = profile_row(P, reff);
    const double w = P.w[i];
    if (kRowTable) {
      const double* r = s_rowtab + (size_t)k * ncol;
#pragma unroll
      for (int t = 0; t < kMaxTransitions; ++t)
        if (t < nt) acc[t] += r[4 + nt + t] * w;
      acc[kMaxTransitions] += r[4 + 2 * nt] * w;
      if (P.per_voxel) {
        double* o = P.per_voxel + (size_t)i * ncol;
        for (int c = 0; c < ncol; ++c) o[c] = r[c];
      }
      continue;
    }
    const double te = P.profile[3 * k + 1], ne = P.profile[3 * k + 2];
    RowEval r;
    eval_row(P.tab, ne, te, P.conc, r);
#pragma unroll
    for (int t = 0; t < kMaxTransitions; ++t)
      if (t < nt) acc[t] += r.emis[t] * w;
    acc[kMaxTransitions] += r.total * w;
    if (P.per_voxel) {
      double* o = P.per_voxel + (size_t)i * ncol;
      o[0] = ne, o[1] = te, o[2] = r.fa_ion, o[3] = r.fa_last;
      for (int t = 0; t < nt; ++t) o[4 + t] = r.pec[t], o[4 + nt + t] = r.emis[t];
      o[4 + 2 * nt] = r.total;
    }
  }

  // block reduction, then the last block adds the per-block partials in block order (deterministic)
  __shared__ double s_red[kK2Threads / 32][kMaxTransitions + 1];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t <= kMaxTransitions; ++t) {
    double v = acc[t];
    for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[warp][t] = v;
  }
  __syncthreads();
  double* partials = P.scratch + 8;
  if (threadIdx.x <= kMaxTransitions) {
    double v = 0.0;
    for (int wq = 0; wq < kK2Threads / 32; ++wq) v += s_red[wq][threadIdx.x];
    partials[(size_t)blockIdx.x * (kMaxTransitions + 1) + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(P.done_counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last && threadIdx.x < nout) {
    const int slot = threadIdx.x == nt ? kMaxTransitions : threadIdx.x;  // TOTAL is kept in the last slot
    double v = 0.0;
    for (unsigned int b = 0; b < gridDim.x; ++b) v += partials[(size_t)b * (kMaxTransitions + 1) + slot];
    P.flux_out[threadIdx.x] = v;
    if (threadIdx.x == 0) P.scratch[1] = boundary;
  }
}

// Voxel pass of the factorised form: H[k] += w_v for the profile row k the voxel maps to.
// Shared-memory histogram per block (K <= 12288 rows = 96 kB), flushed with one fp64 atomic per non-empty bin.
__global__ void __launch_bounds__(kK2Threads) weight_histogram_kernel(const K2Params P) {
  extern __shared__ double s_hist[];
  for (int k = threadIdx.x; k < P.K; k += kK2Threads) s_hist[k] = 0.0;
  __syncthreads();
  const double boundary = cut_boundary(P);
  const long long n = voxel_count(P);
  for (long long i = blockIdx.x * (long long)kK2Threads + threadIdx.x; i < n; i += (long long)gridDim.x * kK2Threads) {
    const double reff = voxel_reff(P, i);
    const double w = P.w[i];
    if ((reff == reff) && !(reff >= boundary) && w != 0.0) atomicAdd(&s_hist[profile_row(P, reff)], w);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < P.K; k += kK2Threads)
    if (s_hist[k] != 0.0) atomicAdd(&P.hist_out[k], s_hist[k]);
  if (blockIdx.x == 0 && threadIdx.x == 0) P.scratch[1] = boundary;
}

// Fast voxel pass for uint16 millimetre bins (the form large grids keep their Reff in): every voxel is read once,
// 10 B/voxel, with 16-byte vector loads; no search in the hot loop - the weights are binned by the raw millimetre
// value in shared memory and the boundary cut and the bin -> profile-row mapping are applied afterwards on the
// <= 4096-bin histogram (histogram_rows_kernel).  Only voxels with w != 0 count: they are the visible ones
// (every passing ray adds a strictly positive weight), which is the set the reference takes the maximum over.
constexpr int kMmBins = 4096;

__device__ __forceinline__ void bin_one(double* s_h, unsigned int b, double w, unsigned int& n_nonzero) {
  n_nonzero += w != 0.0;
  if (w != 0.0 && b != 0xFFFFu) atomicAdd(&s_h[min(b, (unsigned int)(kMmBins - 1))], w);
}

__global__ void __launch_bounds__(kK2Threads) weight_histogram_mm_kernel(const uint16_t* __restrict__ reff_mm,
                                                                       const double* __restrict__ w, long long n_max,
                                                                       const unsigned long long* n_device,
                                                                       double* __restrict__ hist_mm,
                                                                       unsigned long long* nonzero_out) {
  __shared__ double s_h[kMmBins];
  unsigned int nzc = 0;  // voxels with a non-zero weight seen by this thread (= visible voxels, see above)
  for (int k = threadIdx.x; k < kMmBins; k += kK2Threads) s_h[k] = 0.0;
  __syncthreads();
  const long long n = n_device ? (long long)min((unsigned long long)n_max, *n_device) : n_max;
  const long long n8 = n >> 3;
  const uint4* r4 = reinterpret_cast<const uint4*>(reff_mm);
  const double2* w2 = reinterpret_cast<const double2*>(w);
  const long long stride = (long long)gridDim.x * kK2Threads;
  for (long long i = blockIdx.x * (long long)kK2Threads + threadIdx.x; i < n8; i += 2 * stride) {
    const long long j = i + stride;
    const bool second = j < n8;
    // all loads of the two 8-voxel groups are issued before any is consumed
    const uint4 ra = __ldcs(r4 + i);
    const double2 a0 = __ldcs(w2 + 4 * i), a1 = __ldcs(w2 + 4 * i + 1), a2 = __ldcs(w2 + 4 * i + 2), a3 = __ldcs(w2 + 4 * i + 3);
    uint4 rb = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
    double2 b0 = make_double2(0, 0), b1 = b0, b2 = b0, b3 = b0;
    if (second) {
      rb = __ldcs(r4 + j);
      b0 = __ldcs(w2 + 4 * j), b1 = __ldcs(w2 + 4 * j + 1), b2 = __ldcs(w2 + 4 * j + 2), b3 = __ldcs(w2 + 4 * j + 3);
    }
    bin_one(s_h, ra.x & 0xFFFFu, a0.x, nzc), bin_one(s_h, ra.x >> 16, a0.y, nzc);
    bin_one(s_h, ra.y & 0xFFFFu, a1.x, nzc), bin_one(s_h, ra.y >> 16, a1.y, nzc);
    bin_one(s_h, ra.z & 0xFFFFu, a2.x, nzc), bin_one(s_h, ra.z >> 16, a2.y, nzc);
    bin_one(s_h, ra.w & 0xFFFFu, a3.x, nzc), bin_one(s_h, ra.w >> 16, a3.y, nzc);
    bin_one(s_h, rb.x & 0xFFFFu, b0.x, nzc), bin_one(s_h, rb.x >> 16, b0.y, nzc);
    bin_one(s_h, rb.y & 0xFFFFu, b1.x, nzc), bin_one(s_h, rb.y >> 16, b1.y, nzc);
    bin_one(s_h, rb.z & 0xFFFFu, b2.x, nzc), bin_one(s_h, rb.z >> 16, b2.y, nzc);
    bin_one(s_h, rb.w & 0xFFFFu, b3.x, nzc), bin_one(s_h, rb.w >> 16, b3.y, nzc);
  }
  if (blockIdx.x == 0)  // the < 8 voxels left over
    for (long long i = (n8 << 3) + threadIdx.x; i < n; i += kK2Threads) bin_one(s_h, reff_mm[i], w[i], nzc);
  __syncthreads();
  for (int k = threadIdx.x; k < kMmBins; k += kK2Threads)
    if (s_h[k] != 0.0) atomicAdd(&hist_mm[k], s_h[k]);
  if (nonzero_out) {
    for (int off = 16; off; off >>= 1) nzc += __shfl_down_sync(0xffffffffu, nzc, off);
    if ((threadIdx.x & 31) == 0 && nzc) atomicAdd(nonzero_out, (unsigned long long)nzc);
  }
}

// Boundary cut and millimetre bin -> profile row on the small histogram (one block).
__global__ void __launch_bounds__(kK2Threads) histogram_rows_kernel(const K2Params P, const double* __restrict__ hist_mm) {
  __shared__ int s_max;
  if (threadIdx.x == 0) s_max = -1;
  __syncthreads();
  int m = -1;
  for (int b = threadIdx.x; b < kMmBins; b += kK2Threads)
    if (hist_mm[b] != 0.0) m = b;
  atomicMax(&s_max, m);
  __syncthreads();
  const double mapped = P.mapped_max ? *P.mapped_max : (s_max >= 0 ? (double)s_max / 1000.0 : 0.0);
  const double boundary = fmin(mapped, P.profile_reff_max);
  for (int b = threadIdx.x; b < kMmBins; b += kK2Threads) {
    const double h = hist_mm[b], reff = (double)b / 1000.0;
    if (h != 0.0 && !(reff >= boundary)) atomicAdd(&P.hist_out[profile_row(P, reff)], h);
  }
  if (threadIdx.x == 0) {
    if (P.scratch) P.scratch[0] = mapped, P.scratch[1] = boundary;
    if (P.boundary_out) *P.boundary_out = boundary;
  }
}

struct SweepParams {
  DevTables tab;
  const double* profiles;  // (n_slices, K, 3)
  const double* hist;      // (K)
  int K, n_slices;
  double conc;
  double* flux_out;        // (n_slices, n_trans + 1)
};

// One block per time slice: every thread evaluates profile rows (only those with weight) and the block reduces in a
// fixed order.  THREADS = 256 for large batches; 1024 for a few slices, where the dependent table look-ups of a row
// are the critical path and one row per thread is the shortest chain.
template <int THREADS>
__global__ void __launch_bounds__(THREADS) sweep_kernel(const SweepParams P) {
  const int s = blockIdx.x, nt = P.tab.n_trans;
  const double* prof = P.profiles + (size_t)s * P.K * 3;
  double acc[kMaxTransitions + 1];
#pragma unroll
  for (int t = 0; t <= kMaxTransitions; ++t) acc[t] = 0.0;
  for (int k = threadIdx.x; k < P.K; k += THREADS) {
    const double h = P.hist[k];
    if (h == 0.0) continue;
    RowEval r;
    eval_row(P.tab, prof[3 * k + 2], prof[3 * k + 1], P.conc, r);
#pragma unroll
    for (int t = 0; t < kMaxTransitions; ++t)
      if (t < nt) acc[t] += r.emis[t] * h;
    acc[kMaxTransitions] += r.total * h;
  }
  __shared__ double s_red[THREADS / 32][kMaxTransitions + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t <= kMaxTransitions; ++t) {
    double v = acc[t];
    for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x <= nt) {
    const int slot = threadIdx.x == nt ? kMaxTransitions : threadIdx.x;
    double v = 0.0;
    for (int wq = 0; wq < THREADS / 32; ++wq) v += s_red[wq][slot];
    P.flux_out[(size_t)s * (nt + 1) + threadIdx.x] = v;
  }
}

// Stage 2 of a whole-volume run in ONE launch: block (channel, slice) applies the Reff cut to the channel's
// millimetre histogram (largest non-empty bin, reader.py:210-217), maps the bins to profile rows in shared memory
// (:220-238), evaluates the rows that carry weight (:282-378) and reduces the flux (:380-401) in a fixed order.
// Replaces histogram_rows_kernel<<<1,256>>> + sweep_kernel<<<1,1024>>> + the fill of the row histogram per channel.
constexpr int kMaxChannels = 8;
constexpr int kFluxThreads = 1024;

constexpr int kMaxPeers = 16;

struct FluxParams {
  DevTables tab[kMaxChannels];
  const double* profiles;  // (n_slices, K, 3)
  const double* hist_mm;   // (n_channels, kMmBins); unused when n_peers > 0
  // The exchange fused into Stage 2 (n_peers > 0): rank r's packed buffer [n_channels x kMmBins | n_channels x
  // n_counters] as mapped into this process (NVLink peer / symmetric memory).  The block sums the peers' bins in rank
  // order - every rank gets bit-identical sums - into shared memory and carries on from there.
  const double* peers[kMaxPeers];
  int n_peers, n_channels, n_counters;
  double* hist_sum_out;      // null, or (n_channels, kMmBins): the global histograms (written by the slice-0 blocks)
  double* counters_sum_out;  // null, or (n_channels, n_counters)
  int K, n_slices, profile_sorted;
  double profile_reff_max, conc;
  double* flux_out;      // (n_channels, n_slices, n_trans + 1)
  double* boundary_out;  // (n_channels) or null
  int n_out;             // row stride of flux_out per (channel, slice): max over channels of n_trans + 1
};

__global__ void __launch_bounds__(kFluxThreads) histogram_flux_kernel(const __grid_constant__ FluxParams P) {
  extern __shared__ double s_rows[];  // (K), then (kMmBins) summed bins when the peers are read here
  __shared__ int s_max;
  __shared__ double s_red[kFluxThreads / 32][kMaxTransitions + 1];
  const int ch = blockIdx.x, s = blockIdx.y;
  const DevTables& T = P.tab[ch];
  const int nt = T.n_trans;
  const double* hist_mm = P.hist_mm + (size_t)ch * kMmBins;
  const double* prof = P.profiles + (size_t)s * P.K * 3;
  if (threadIdx.x == 0) s_max = -1;
  for (int k = threadIdx.x; k < P.K; k += kFluxThreads) s_rows[k] = 0.0;
  if (P.n_peers > 0) {
    double* s_hist = s_rows + P.K;
    constexpr int kPer = kMmBins / kFluxThreads;  // bins per thread: all loads of a peer in flight together
    double h[kPer];
#pragma unroll
    for (int i = 0; i < kPer; ++i) h[i] = 0.0;
    for (int r = 0; r < P.n_peers; ++r) {
      const double* src = P.peers[r] + (size_t)ch * kMmBins;
      double v[kPer];
#pragma unroll
      for (int i = 0; i < kPer; ++i) v[i] = __ldcg(src + threadIdx.x + i * kFluxThreads);
#pragma unroll
      for (int i = 0; i < kPer; ++i) h[i] += v[i];
    }
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const int b = threadIdx.x + i * kFluxThreads;
      s_hist[b] = h[i];
      if (s == 0 && P.hist_sum_out) P.hist_sum_out[(size_t)ch * kMmBins + b] = h[i];
    }
    if (s == 0 && P.counters_sum_out && threadIdx.x < P.n_counters) {
      double c = 0.0;
      const size_t off = (size_t)P.n_channels * kMmBins + (size_t)ch * P.n_counters + threadIdx.x;
      for (int r = 0; r < P.n_peers; ++r) c += __ldcg(P.peers[r] + off);
      P.counters_sum_out[(size_t)ch * P.n_counters + threadIdx.x] = c;
    }
    hist_mm = s_hist;
  }
  __syncthreads();
  int m = -1;
  for (int b = threadIdx.x; b < kMmBins; b += kFluxThreads)
    if (hist_mm[b] != 0.0) m = b;
  if (m >= 0) atomicMax(&s_max, m);
  __syncthreads();
  const double mapped = s_max >= 0 ? (double)s_max / 1000.0 : 0.0;
  const double boundary = fmin(mapped, P.profile_reff_max);
  K2Params R{};  // profile_row() reads these
  R.profile = prof, R.K = P.K, R.profile_sorted = P.profile_sorted;
  for (int b = threadIdx.x; b < kMmBins; b += kFluxThreads) {
    const double h = hist_mm[b], reff = (double)b / 1000.0;
    if (h != 0.0 && !(reff >= boundary)) atomicAdd(&s_rows[profile_row(R, reff)], h);
  }
  __syncthreads();
  double acc[kMaxTransitions + 1];
#pragma unroll
  for (int t = 0; t <= kMaxTransitions; ++t) acc[t] = 0.0;
  for (int k = threadIdx.x; k < P.K; k += kFluxThreads) {
    const double h = s_rows[k];
    if (h == 0.0) continue;
    RowEval r;
    eval_row(T, prof[3 * k + 2], prof[3 * k + 1], P.conc, r);
#pragma unroll
    for (int t = 0; t < kMaxTransitions; ++t)
      if (t < nt) acc[t] += r.emis[t] * h;
    acc[kMaxTransitions] += r.total * h;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int t = 0; t <= kMaxTransitions; ++t) {
    double v = acc[t];
    for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) s_red[warp][t] = v;
  }
  __syncthreads();
  if (threadIdx.x <= nt) {
    const int slot = threadIdx.x == nt ? kMaxTransitions : threadIdx.x;
    double v = 0.0;
    for (int wq = 0; wq < kFluxThreads / 32; ++wq) v += s_red[wq][slot];
    P.flux_out[((size_t)ch * P.n_slices + s) * P.n_out + threadIdx.x] = v;
  }
  if (threadIdx.x == 0 && s == 0 && P.boundary_out) P.boundary_out[ch] = boundary;
}

}  // namespace com

// ---------------------------------------------------------------------------------------------------------
struct ComTables {
  com::DevTables dev{};
  void* d_blob = nullptr;
  int n_trans = 0;
};

#define COM_CUDA(x)                                                            \
  do {                                                                         \
    cudaError_t e__ = (x);                                                     \
    if (e__ != cudaSuccess) {                                                  \
      com::set_error("%s failed: %s", #x, cudaGetErrorString(e__));            \
      return COM_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool non_decreasing(const double* a, int n, int stride = 1) {
  for (int i = 1; i < n; ++i)
    if (!(a[(size_t)i * stride] >= a[(size_t)(i - 1) * stride])) return false;
  return true;
}
}  // namespace

// com_weight_histogram_mm, also adding the number of voxels with a non-zero weight to *nonzero_out (device; null: not counted)
namespace com {
int weight_histogram_mm_count(const uint16_t* reff_mm, const double* w, long long n, const unsigned long long* n_device,
                              double* hist_mm, unsigned long long* nonzero_out, cudaStream_t stream);
}
int com::weight_histogram_mm_count(const uint16_t* reff_mm, const double* w, long long n, const unsigned long long* n_device,
                                   double* hist_mm, unsigned long long* nonzero_out, cudaStream_t stream) {
  if (!reff_mm || !w || !hist_mm || n < 0 || (reinterpret_cast<uintptr_t>(reff_mm) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) {
    com::set_error("com_weight_histogram_mm: null or not 16-byte aligned arguments");
    return COM_E_BADARG;
  }
  if (n == 0) return COM_OK;
  const int blocks = (int)std::min<long long>((n / 8 + com::kK2Threads - 1) / com::kK2Threads + 1, 148 * 6);
  com::weight_histogram_mm_kernel<<<blocks, com::kK2Threads, 0, stream>>>(reff_mm, w, n, n_device, hist_mm, nonzero_out);
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}

extern "C" {

// Validates the description and lays all tables out in one fp64 blob; `dev` gets pointers relative to a null base
// (offsets), rebased by bind_tables once the blob has a device address.
static int build_tables_blob(const ComTablesDesc* d, std::vector<double>& blob, com::DevTables& dev) {
  if (!d || d->n_fa < 1 || !d->fa_T || !d->fa_ion || !d->fa_last || d->n_transitions < 1 ||
      d->n_transitions > com::kMaxTransitions) {
    com::set_error("Stage-2 tables: bad arguments (1..%d transitions)", com::kMaxTransitions);
    return COM_E_BADARG;
  }
  if (!non_decreasing(d->fa_T, d->n_fa)) {
    com::set_error("Stage-2 tables: fractional-abundance temperature grid must be non-decreasing");
    return COM_E_BADARG;
  }
  blob.clear();
  auto push = [&](const double* p, size_t n) {
    size_t off = blob.size();
    blob.insert(blob.end(), p, p + n);
    while (blob.size() % 32) blob.push_back(0.0);
    return reinterpret_cast<const double*>(off * sizeof(double));
  };
  dev = com::DevTables{};
  dev.n_fa = d->n_fa;
  dev.n_trans = d->n_transitions;
  dev.fa_T = push(d->fa_T, d->n_fa), dev.fa_ion = push(d->fa_ion, d->n_fa), dev.fa_last = push(d->fa_last, d->n_fa);
  for (int t = 0; t < d->n_transitions; ++t) {
    const ComPecDesc& p = d->pec[t];
    if (p.n_ne < 2 || p.n_te < 2 || p.n_grid < 1 || !p.ne_nodes || !p.te_nodes || !p.table || !p.ne_grid || !p.te_grid ||
        !non_decreasing(p.ne_nodes, p.n_ne) || !non_decreasing(p.te_nodes, p.n_te) ||
        !non_decreasing(p.ne_grid, p.n_grid) || !non_decreasing(p.te_grid, p.n_grid)) {
      com::set_error("Stage-2 tables: PEC block %d incomplete or not sorted", t);
      return COM_E_BADARG;
    }
    if (p.interpolation != COM_PEC_REFERENCE && p.interpolation != COM_PEC_LOGLOG) {
      com::set_error("Stage-2 tables: PEC block %d asks for unknown interpolation %d", t, p.interpolation);
      return COM_E_BADARG;
    }
    dev.pec[t] = com::DevPec{p.n_ne, p.n_te, p.n_grid, p.use_fully_stripped, p.interpolation, push(p.ne_nodes, p.n_ne),
                             push(p.te_nodes, p.n_te), push(p.table, (size_t)p.n_ne * p.n_te),
                             push(p.ne_grid, p.n_grid), push(p.te_grid, p.n_grid)};
  }
  return COM_OK;
}

static void bind_tables(com::DevTables& dev, const void* base) {
  auto fix = [&](const double*& p) {
    p = reinterpret_cast<const double*>(static_cast<const char*>(base) + reinterpret_cast<size_t>(p));
  };
  fix(dev.fa_T), fix(dev.fa_ion), fix(dev.fa_last);
  for (int t = 0; t < dev.n_trans; ++t) {
    fix(dev.pec[t].ne_nodes), fix(dev.pec[t].te_nodes), fix(dev.pec[t].table);
    fix(dev.pec[t].ne_grid), fix(dev.pec[t].te_grid);
  }
}

int com_tables_create(const ComTablesDesc* d, ComTables** out) {
  if (!out) {
    com::set_error("com_tables_create: null argument");
    return COM_E_BADARG;
  }
  std::vector<double> blob;
  com::DevTables dev;
  int rc = build_tables_blob(d, blob, dev);
  if (rc) return rc;
  ComTables* T = new (std::nothrow) ComTables();
  if (!T) return COM_E_NOMEM;
  cudaError_t e = cudaMalloc(&T->d_blob, blob.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(T->d_blob, blob.data(), blob.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    com::set_error("com_tables_create: %s", cudaGetErrorString(e));
    if (T->d_blob) cudaFree(T->d_blob);
    delete T;
    return COM_E_CUDA;
  }
  bind_tables(dev, T->d_blob);
  T->dev = dev;
  T->n_trans = d->n_transitions;
  *out = T;
  return COM_OK;
}

int com_tables_destroy(ComTables* t) {
  if (!t) return COM_OK;
  if (t->d_blob) cudaFree(t->d_blob);
  delete t;
  return COM_OK;
}

size_t com_emissivity_workspace_bytes(int64_t n) {
  long long blocks = std::min<long long>(std::max<long long>(1, (n + com::kK2Threads - 1) / com::kK2Threads), 148 * 8);  // upper bound of any launch below
  return 256 + align_up(sizeof(double) * (8 + (size_t)blocks * (com::kMaxTransitions + 1)), 256) +
         sizeof(double) * com::kMmBins;  // + the millimetre-bin histogram of the fast voxel pass
}

static int fill_params(com::K2Params& P, const ComTables* tables, const double* profile, int32_t K, int32_t sorted,
                       double profile_reff_max, double concentration, const double* reff, const uint16_t* reff_mm,
                       const int64_t* reff_index, const double* w, int64_t n, const uint64_t* n_device,
                       const double* mapped_max, void* workspace, size_t workspace_bytes) {
  if (!tables || !profile || K < 1 || (!reff && !reff_mm) || !w || n < 0 || !workspace ||
      workspace_bytes < com_emissivity_workspace_bytes(n)) {
    com::set_error("com_emissivity_*: bad arguments or workspace too small");
    return COM_E_BADARG;
  }
  P.tab = tables->dev;
  P.profile = profile, P.K = K, P.profile_sorted = sorted, P.profile_reff_max = profile_reff_max;
  P.conc = concentration;
  P.reff = reff, P.reff_mm = reff_mm, P.w = w, P.n = n;
  P.reff_index = reinterpret_cast<const long long*>(reff_index);
  P.n_device = reinterpret_cast<const unsigned long long*>(n_device);
  P.mapped_max = mapped_max;
  P.done_counter = reinterpret_cast<unsigned int*>(workspace);
  P.scratch = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  return COM_OK;
}

int com_emissivity_integrate(const ComTables* tables, const double* profile, int32_t K, int32_t profile_sorted,
                             double profile_reff_max, double concentration, const double* reff,
                             const uint16_t* reff_mm, const int64_t* reff_index, const double* w, int64_t n,
                             const uint64_t* n_device, const double* mapped_reff_max, double* flux_out,
                             double* per_voxel_out, uint8_t* valid_out, void* workspace, size_t workspace_bytes,
                             com_stream_t stream) {
  com::K2Params P{};
  int rc = fill_params(P, tables, profile, K, profile_sorted, profile_reff_max, concentration, reff, reff_mm,
                       reff_index, w, n, n_device, mapped_reff_max, workspace, workspace_bytes);
  if (rc) return rc;
  if (!flux_out) {
    com::set_error("com_emissivity_integrate: flux_out is null");
    return COM_E_BADARG;
  }
  P.flux_out = flux_out, P.per_voxel = per_voxel_out, P.valid_out = valid_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  COM_CUDA(cudaMemsetAsync(workspace, 0, 256 + 64, s));
  COM_CUDA(cudaMemsetAsync(flux_out, 0, sizeof(double) * (tables->n_trans + 1), s));
  if (n == 0) return COM_OK;
  // row table in shared memory when it fits next to three resident blocks per SM; the rows are evaluated once per
  // block, so small inputs get few blocks (a block's set-up is K row evaluations)
  const size_t rowtab = sizeof(double) * (size_t)K * (4 + 2 * tables->n_trans + 1);
  const bool use_rowtab = rowtab <= 72 * 1024 && n >= 4 * (long long)K;
  const int per_sm = use_rowtab ? 3 : 2;
  const int blocks = (int)std::min<long long>((n + com::kK2Threads - 1) / com::kK2Threads, 148 * per_sm);
  if (!mapped_reff_max) com::reff_max_kernel<<<blocks, com::kK2Threads, 0, s>>>(P);
  if (use_rowtab) {
    const int rblocks = (int)std::min<long long>(blocks, std::max<long long>(1, n / (4 * (long long)K)));
    COM_CUDA(cudaFuncSetAttribute(com::emissivity_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rowtab));
    com::emissivity_kernel<true><<<rblocks, com::kK2Threads, rowtab, s>>>(P);
  } else {
    com::emissivity_kernel<false><<<blocks, com::kK2Threads, 0, s>>>(P);
  }
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}

int com_weight_histogram(const double* profile, int32_t K, int32_t profile_sorted, double profile_reff_max,
                         const double* reff, const uint16_t* reff_mm, const int64_t* reff_index, const double* w,
                         int64_t n, const uint64_t* n_device, const double* mapped_reff_max, double* hist_out,
                         double* boundary_out, void* workspace, size_t workspace_bytes, com_stream_t stream) {
  com::K2Params P{};
  ComTables dummy;
  int rc = fill_params(P, &dummy, profile, K, profile_sorted, profile_reff_max, 0.0, reff, reff_mm, reff_index, w, n,
                       n_device, mapped_reff_max, workspace, workspace_bytes);
  if (rc) return rc;
  if (!hist_out || K > 12288) {
    com::set_error("com_weight_histogram: hist_out null or more than 12288 profile rows");
    return COM_E_BADARG;
  }
  P.hist_out = hist_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  COM_CUDA(cudaMemsetAsync(workspace, 0, 256 + 64, s));
  COM_CUDA(cudaMemsetAsync(hist_out, 0, sizeof(double) * K, s));
  if (n > 0 && reff_mm && !reff && !reff_index && (reinterpret_cast<uintptr_t>(reff_mm) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(w) & 15) == 0) {
    // fast path: one 10 B/voxel pass binned by millimetre, then the cut and row mapping on 4096 bins
    double* hist_mm = reinterpret_cast<double*>(static_cast<char*>(workspace) + workspace_bytes - sizeof(double) * com::kMmBins);
    COM_CUDA(cudaMemsetAsync(hist_mm, 0, sizeof(double) * com::kMmBins, s));
    const int blocks = (int)std::min<long long>((n / 8 + com::kK2Threads - 1) / com::kK2Threads + 1, 148 * 6);
    com::weight_histogram_mm_kernel<<<blocks, com::kK2Threads, 0, s>>>(
        reff_mm, w, n, reinterpret_cast<const unsigned long long*>(n_device), hist_mm, nullptr);
    com::histogram_rows_kernel<<<1, com::kK2Threads, 0, s>>>(P, hist_mm);
    COM_CUDA(cudaGetLastError());
  } else if (n > 0) {
    const size_t smem = sizeof(double) * K;
    // per device and cheap: set it on every call rather than remember it per process
    COM_CUDA(cudaFuncSetAttribute(com::weight_histogram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    const int blocks = (int)std::min<long long>((n + com::kK2Threads - 1) / com::kK2Threads, 148 * 4);
    if (!mapped_reff_max) com::reff_max_kernel<<<blocks, com::kK2Threads, 0, s>>>(P);
    com::weight_histogram_kernel<<<blocks, com::kK2Threads, smem, s>>>(P);
    COM_CUDA(cudaGetLastError());
  }
  if (boundary_out) COM_CUDA(cudaMemcpyAsync(boundary_out, P.scratch + 1, sizeof(double), cudaMemcpyDeviceToDevice, s));
  return COM_OK;
}

int com_weight_histogram_mm(const uint16_t* reff_mm, const double* w, int64_t n, const uint64_t* n_device,
                            double* hist_mm, com_stream_t stream) {
  return com::weight_histogram_mm_count(reff_mm, w, n, reinterpret_cast<const unsigned long long*>(n_device), hist_mm, nullptr,
                                        static_cast<cudaStream_t>(stream));
}


int com_histogram_rows(const double* profile, int32_t K, int32_t profile_sorted, double profile_reff_max,
                       const double* hist_mm, const double* mapped_reff_max, double* hist_rows_out,
                       double* boundary_out, com_stream_t stream) {
  if (!profile || K < 1 || !hist_mm || !hist_rows_out) {
    com::set_error("com_histogram_rows: bad arguments");
    return COM_E_BADARG;
  }
  com::K2Params P{};
  P.profile = profile, P.K = K, P.profile_sorted = profile_sorted, P.profile_reff_max = profile_reff_max;
  P.mapped_max = mapped_reff_max;
  P.hist_out = hist_rows_out;
  P.boundary_out = boundary_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  COM_CUDA(cudaMemsetAsync(hist_rows_out, 0, sizeof(double) * K, s));
  com::histogram_rows_kernel<<<1, com::kK2Threads, 0, s>>>(P, hist_mm);
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}

int com_reff_max(const double* reff, const uint16_t* reff_mm, const int64_t* reff_index, int64_t n,
                 const uint64_t* n_device, double* max_out, com_stream_t stream) {
  if ((!reff && !reff_mm) || n < 0 || !max_out) {
    com::set_error("com_reff_max: bad arguments");
    return COM_E_BADARG;
  }
  com::K2Params P{};
  P.reff = reff, P.reff_mm = reff_mm, P.n = n;
  P.reff_index = reinterpret_cast<const long long*>(reff_index);
  P.n_device = reinterpret_cast<const unsigned long long*>(n_device);
  P.scratch = max_out;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  COM_CUDA(cudaMemsetAsync(max_out, 0, sizeof(double), s));
  if (n > 0) {
    const int blocks = (int)std::min<long long>((n + com::kK2Threads - 1) / com::kK2Threads, 148 * 8);
    com::reff_max_kernel<<<blocks, com::kK2Threads, 0, s>>>(P);
    COM_CUDA(cudaGetLastError());
  }
  return COM_OK;
}

int com_emissivity_sweep(const ComTables* tables, const double* profiles, int32_t n_slices, int32_t K,
                         const double* hist, double concentration, double* flux_out, com_stream_t stream) {
  if (!tables || !profiles || !hist || !flux_out || n_slices < 0 || K < 1) {
    com::set_error("com_emissivity_sweep: bad arguments");
    return COM_E_BADARG;
  }
  if (n_slices == 0) return COM_OK;
  com::SweepParams P{tables->dev, profiles, hist, K, n_slices, concentration, flux_out};
  if (n_slices <= 32)
    com::sweep_kernel<1024><<<n_slices, 1024, 0, static_cast<cudaStream_t>(stream)>>>(P);
  else
    com::sweep_kernel<com::kK2Threads><<<n_slices, com::kK2Threads, 0, static_cast<cudaStream_t>(stream)>>>(P);
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}

int com_histograms_flux(int32_t n_channels, const ComTables* const* tables, const double* profiles, int32_t n_slices,
                        int32_t K, int32_t profile_sorted, double profile_reff_max, double concentration,
                        const double* hist_mm, double* flux_out, int32_t flux_stride, double* boundary_out,
                        com_stream_t stream) {
  if (n_channels < 1 || n_channels > com::kMaxChannels || !tables || !profiles || n_slices < 1 || K < 1 || !hist_mm ||
      !flux_out) {
    com::set_error("com_histograms_flux: bad arguments (1..%d channels)", com::kMaxChannels);
    return COM_E_BADARG;
  }
  com::FluxParams P{};
  int n_out = 0;
  for (int c = 0; c < n_channels; ++c) {
    if (!tables[c]) {
      com::set_error("com_histograms_flux: null tables handle for channel %d", c);
      return COM_E_BADARG;
    }
    P.tab[c] = tables[c]->dev;
    n_out = std::max(n_out, tables[c]->n_trans + 1);
  }
  if (flux_stride < n_out) {
    com::set_error("com_histograms_flux: flux_stride %d < transitions + 1 = %d", flux_stride, n_out);
    return COM_E_BADARG;
  }
  const size_t smem = sizeof(double) * (size_t)K;
  if (smem > 160 * 1024) {
    com::set_error("com_histograms_flux: at most %d profile rows", 160 * 1024 / 8);
    return COM_E_BADARG;
  }
  P.profiles = profiles, P.hist_mm = hist_mm, P.K = K, P.n_slices = n_slices, P.profile_sorted = profile_sorted;
  P.profile_reff_max = profile_reff_max, P.conc = concentration;
  P.flux_out = flux_out, P.boundary_out = boundary_out, P.n_out = flux_stride;
  if (smem > 40 * 1024)
    COM_CUDA(cudaFuncSetAttribute(com::histogram_flux_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  com::histogram_flux_kernel<<<dim3(n_channels, n_slices), com::kFluxThreads, smem, static_cast<cudaStream_t>(stream)>>>(P);
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}

int com_histograms_flux_peers(int32_t n_channels, const ComTables* const* tables, const double* profiles,
                              int32_t n_slices, int32_t K, int32_t profile_sorted, double profile_reff_max,
                              double concentration, int32_t n_peers, const double* const* peers_h, int32_t n_counters,
                              double* hist_sum_out, double* counters_sum_out, double* flux_out, int32_t flux_stride,
                              double* boundary_out, com_stream_t stream) {
  if (n_channels < 1 || n_channels > com::kMaxChannels || !tables || !profiles || n_slices < 1 || K < 1 || n_peers < 1 ||
      n_peers > com::kMaxPeers || !peers_h || n_counters < 0 || n_counters > com::kFluxThreads || !flux_out) {
    com::set_error("com_histograms_flux_peers: bad arguments (1..%d channels, 1..%d peers)", com::kMaxChannels, com::kMaxPeers);
    return COM_E_BADARG;
  }
  // many slices: sum the peers once (a one-slice launch that only keeps the sums), then sweep from the local copy
  const bool two_phase = n_slices > 4;
  if (two_phase && !hist_sum_out) {
    com::set_error("com_histograms_flux_peers: more than 4 slices need hist_sum_out (the peers are summed once)");
    return COM_E_BADARG;
  }
  com::FluxParams P{};
  int n_out = 0;
  for (int c = 0; c < n_channels; ++c) {
    if (!tables[c]) {
      com::set_error("com_histograms_flux_peers: null tables handle for channel %d", c);
      return COM_E_BADARG;
    }
    P.tab[c] = tables[c]->dev;
    n_out = std::max(n_out, tables[c]->n_trans + 1);
  }
  if (flux_stride < n_out) {
    com::set_error("com_histograms_flux_peers: flux_stride %d < transitions + 1 = %d", flux_stride, n_out);
    return COM_E_BADARG;
  }
  for (int r = 0; r < n_peers; ++r) {
    if (!peers_h[r]) {
      com::set_error("com_histograms_flux_peers: null buffer for rank %d", r);
      return COM_E_BADARG;
    }
    P.peers[r] = peers_h[r];
  }
  const size_t smem = sizeof(double) * ((size_t)K + com::kMmBins);
  if (smem > 160 * 1024) {
    com::set_error("com_histograms_flux_peers: at most %d profile rows", 160 * 1024 / 8 - com::kMmBins);
    return COM_E_BADARG;
  }
  P.profiles = profiles, P.hist_mm = nullptr, P.K = K, P.n_slices = two_phase ? 1 : n_slices, P.profile_sorted = profile_sorted;
  P.profile_reff_max = profile_reff_max, P.conc = concentration;
  P.flux_out = flux_out, P.boundary_out = boundary_out, P.n_out = flux_stride;
  P.n_peers = n_peers, P.n_channels = n_channels, P.n_counters = n_counters;
  P.hist_sum_out = hist_sum_out, P.counters_sum_out = counters_sum_out;
  COM_CUDA(cudaFuncSetAttribute(com::histogram_flux_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  com::histogram_flux_kernel<<<dim3(n_channels, P.n_slices), com::kFluxThreads, smem, static_cast<cudaStream_t>(stream)>>>(P);
  COM_CUDA(cudaGetLastError());
  if (two_phase)
    return com_histograms_flux(n_channels, tables, profiles, n_slices, K, profile_sorted, profile_reff_max, concentration,
                               hist_sum_out, flux_out, flux_stride, boundary_out, stream);
  return COM_OK;
}

}  // extern "C"

// reff_h (n values), or reff_table_h + reff_index_h: voxel i has Reff reff_table_h[reff_index_h[i]] - the join of
// Emissivity.get_plas_coords_with_rad_frac (reader.py:161-197), done while the inputs are staged for the upload.
namespace com {
int host_call_wait(cudaStream_t s);  // volume.cu
}

static int emissivity_host(const ComTables* tables_handle, const ComTablesDesc* tables_h, const double* profile_h, int32_t K,
                           double concentration,
                           const double* reff_h, const double* reff_table_h, int64_t n_table, const int64_t* reff_index_h,
                           const double* w_h, int64_t n, double* flux_out_h, double* per_voxel_out_h, uint8_t* valid_out_h,
                           double* boundary_out_h) {
  if ((!tables_h && !tables_handle) || !profile_h || (!reff_h && !(reff_table_h && reff_index_h && n_table > 0)) || !w_h ||
      !flux_out_h || n < 0 || K < 1) {
    com::set_error("com_emissivity_integrate_host: bad arguments");
    return COM_E_BADARG;
  }
  if (!reff_h)
    for (int64_t i = 0; i < n; ++i)
      if (reff_index_h[i] < 0 || reff_index_h[i] >= n_table) {
        com::set_error("com_emissivity_integrate_indexed_host: index %lld of voxel %lld outside the Reff table (%lld rows)",
                       (long long)reff_index_h[i], (long long)i, (long long)n_table);
        return COM_E_BADARG;
      }
  std::vector<double> blob;
  ComTables T;  // description given instead of a handle: tables live in the host-call arena, nothing to free
  int rc = COM_OK;
  if (tables_handle) {
    T.dev = tables_handle->dev, T.n_trans = tables_handle->n_trans;
  } else {
    rc = build_tables_blob(tables_h, blob, T.dev);
    if (rc) return rc;
    T.n_trans = tables_h->n_transitions;
  }
  const int nt = T.n_trans, ncol = 4 + 2 * nt + 1;
  const int sorted = non_decreasing(profile_h, K, 3) ? 1 : 0;
  double pmax = profile_h[0];
  for (int k = 1; k < K; ++k) pmax = std::max(pmax, profile_h[3 * (size_t)k]);
  const size_t ws = com_emissivity_workspace_bytes(n);
  const size_t b_tab = align_up(blob.size() * sizeof(double), 256);
  const size_t b_prof = align_up(sizeof(double) * 3 * K, 256), b_n = align_up(sizeof(double) * (size_t)n, 256);
  const size_t b_pv = per_voxel_out_h ? align_up(sizeof(double) * (size_t)n * ncol, 256) : 0;
  const size_t b_valid = valid_out_h ? align_up((size_t)n, 256) : 0;
  const size_t up = b_tab + b_prof + 2 * b_n;  // staged through pinned memory in one copy
  char *d = nullptr, *pin = nullptr;
  rc = com::HostArena::reserve(up + b_pv + b_valid + 256 + ws, up + 256, &d, &pin);
  if (rc) return rc;
  std::memcpy(pin, blob.data(), blob.size() * sizeof(double));
  std::memcpy(pin + b_tab, profile_h, sizeof(double) * 3 * K);
  if (reff_h) {
    std::memcpy(pin + b_tab + b_prof, reff_h, sizeof(double) * (size_t)n);
  } else {
    double* dst = reinterpret_cast<double*>(pin + b_tab + b_prof);
    for (int64_t i = 0; i < n; ++i) dst[i] = reff_table_h[reff_index_h[i]];
  }
  std::memcpy(pin + b_tab + b_prof + b_n, w_h, sizeof(double) * (size_t)n);
  cudaStream_t s = cudaStreamPerThread;
  COM_CUDA(cudaMemcpyAsync(d, pin, up, cudaMemcpyHostToDevice, s));
  if (!tables_handle) bind_tables(T.dev, d);
  double* d_prof = reinterpret_cast<double*>(d + b_tab);
  double* d_reff = reinterpret_cast<double*>(d + b_tab + b_prof);
  double* d_w = reinterpret_cast<double*>(d + b_tab + b_prof + b_n);
  double* d_pv = b_pv ? reinterpret_cast<double*>(d + up) : nullptr;
  uint8_t* d_valid = b_valid ? reinterpret_cast<uint8_t*>(d + up + b_pv) : nullptr;
  double* d_flux = reinterpret_cast<double*>(d + up + b_pv + b_valid);
  char* d_ws = d + up + b_pv + b_valid + 256;
  rc = com_emissivity_integrate(&T, d_prof, K, sorted, pmax, concentration, d_reff, nullptr, nullptr, d_w, n, nullptr,
                                nullptr, d_flux, d_pv, d_valid, d_ws, ws, s);
  if (rc) return rc;
  char* pin_out = pin + up;
  COM_CUDA(cudaMemcpyAsync(pin_out, d_flux, sizeof(double) * (nt + 1), cudaMemcpyDeviceToHost, s));
  COM_CUDA(cudaMemcpyAsync(pin_out + 128, d_ws + 256 + 8, sizeof(double), cudaMemcpyDeviceToHost, s));
  if (d_pv) COM_CUDA(cudaMemcpyAsync(per_voxel_out_h, d_pv, sizeof(double) * (size_t)n * ncol, cudaMemcpyDeviceToHost, s));
  if (d_valid) COM_CUDA(cudaMemcpyAsync(valid_out_h, d_valid, (size_t)n, cudaMemcpyDeviceToHost, s));
  rc = com::host_call_wait(s);
  if (rc) return rc;
  std::memcpy(flux_out_h, pin_out, sizeof(double) * (nt + 1));
  if (boundary_out_h) std::memcpy(boundary_out_h, pin_out + 128, sizeof(double));
  return COM_OK;
}

extern "C" {

int com_emissivity_integrate_host(const ComTablesDesc* tables_h, const double* profile_h, int32_t K,
                                  double concentration, const double* reff_h, const double* w_h, int64_t n,
                                  double* flux_out_h, double* per_voxel_out_h, uint8_t* valid_out_h,
                                  double* boundary_out_h) {
  if (!reff_h) {
    com::set_error("com_emissivity_integrate_host: bad arguments");
    return COM_E_BADARG;
  }
  return emissivity_host(nullptr, tables_h, profile_h, K, concentration, reff_h, nullptr, 0, nullptr, w_h, n, flux_out_h,
                         per_voxel_out_h, valid_out_h, boundary_out_h);
}

int com_emissivity_integrate_indexed_host(const ComTablesDesc* tables_h, const double* profile_h, int32_t K,
                                          double concentration, const double* reff_table_h, int64_t n_table,
                                          const int64_t* reff_index_h, const double* w_h, int64_t n, double* flux_out_h,
                                          double* per_voxel_out_h, uint8_t* valid_out_h, double* boundary_out_h) {
  if (!reff_table_h || !reff_index_h) {
    com::set_error("com_emissivity_integrate_indexed_host: bad arguments");
    return COM_E_BADARG;
  }
  return emissivity_host(nullptr, tables_h, profile_h, K, concentration, nullptr, reff_table_h, n_table, reff_index_h, w_h, n,
                         flux_out_h, per_voxel_out_h, valid_out_h, boundary_out_h);
}

int com_emissivity_integrate_indexed_host_tables(const ComTables* tables, const double* profile_h, int32_t K,
                                                 double concentration, const double* reff_table_h, int64_t n_table,
                                                 const int64_t* reff_index_h, const double* w_h, int64_t n,
                                                 double* flux_out_h, double* per_voxel_out_h, uint8_t* valid_out_h,
                                                 double* boundary_out_h) {
  if (!tables || !reff_table_h || !reff_index_h) {
    com::set_error("com_emissivity_integrate_indexed_host_tables: bad arguments");
    return COM_E_BADARG;
  }
  return emissivity_host(tables, nullptr, profile_h, K, concentration, nullptr, reff_table_h, n_table, reff_index_h, w_h, n,
                         flux_out_h, per_voxel_out_h, valid_out_h, boundary_out_h);
}

}  // extern "C"
