// Whole-volume runs (BASELINE config 5 and the fused Stage 1 -> Stage 2 row of SURVEY.md section 8(f)): a (shard of a)
// lattice too large to keep per-voxel results is processed launch by launch, every launch followed by the 10 B/voxel
// pass that folds its weights into the channel's Reff-millimetre histogram.  What leaves the device per channel is
// that 4096-bin histogram (which ranks all-reduce) and the counters; Stage 2 then runs from the global histogram.
//
// Reference shape of the same computation: Simulation.__init__ (simulation.py:50-72) writes the per-voxel table,
// Emissivity.__init__ (reader.py:68-79) reads it back, joins Reff (:161-197), cuts at the Reff boundary (:210-217,
// 240-242) and sums (:380-401).  The sums depend on the voxels only through sum of w per Reff millimetre.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

#include "scene.h"

namespace com {
int visibility_launch(const ComScene*, const ComLattice*, const double*, long long, long long, double*, unsigned int*,
                      ComRayRecord*, long long, unsigned long long*, const uint16_t*, double*, int, ComStats*, void*,
                      size_t, cudaStream_t);
size_t visibility_workspace_bytes(const DeviceScene&, long long, double);
long long lattice_local_voxels(const ComLattice&);
void visibility_accumulate_stats(bool on);
int weight_histogram_mm_count(const uint16_t*, const double*, long long, const unsigned long long*, double*,
                              unsigned long long*, cudaStream_t);
}  // namespace com

#define COM_CUDA(x)                                                            \
  do {                                                                         \
    cudaError_t e__ = (x);                                                     \
    if (e__ != cudaSuccess) {                                                  \
      com::set_error("%s failed: %s", #x, cudaGetErrorString(e__));            \
      return COM_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

struct ComVolume {
  ComLattice lat{};
  long long n = 0;      // voxels of this (shard of the) lattice
  long long chunk = 0;  // voxels per Stage-1 launch (multiple of 8: the histogram pass loads 16 bytes at a time)
  double cand_fraction = 0;
  int device = 0;
  uint16_t* d_reff_mm = nullptr;  // (n)
  double* d_w = nullptr;          // (chunk)
  void* d_ws = nullptr;           // candidate list of one launch, grown on demand
  size_t ws_bytes = 0;
  double* d_src = nullptr;        // Reff source grid
  size_t src_bytes = 0;
  bool reff_set = false;
};

namespace {
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// thread-local blocking event: the *_host calls wait on it instead of spinning in cudaStreamSynchronize
int blocking_wait(cudaStream_t s) {
  thread_local cudaEvent_t ev = nullptr;
  thread_local int ev_device = -1;
  int dev = 0;
  COM_CUDA(cudaGetDevice(&dev));
  if (!ev || ev_device != dev) {
    if (ev) cudaEventDestroy(ev);
    COM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming));
    ev_device = dev;
  }
  COM_CUDA(cudaEventRecord(ev, s));
  COM_CUDA(cudaEventSynchronize(ev));
  return COM_OK;
}
}  // namespace

namespace com {
int host_call_wait(cudaStream_t s) { return blocking_wait(s); }
}  // namespace com

extern "C" {

int com_volume_create(const ComLattice* lattice_h, int64_t chunk_voxels, double candidate_fraction, ComVolume** out) {
  if (!lattice_h || !out || chunk_voxels < 0) {
    com::set_error("com_volume_create: bad arguments");
    return COM_E_BADARG;
  }
  *out = nullptr;
  const long long n = com::lattice_local_voxels(*lattice_h);
  if (n <= 0) {
    com::set_error("com_volume_create: the lattice (shard) has no voxels");
    return COM_E_BADARG;
  }
  ComVolume* v = new (std::nothrow) ComVolume();
  if (!v) return COM_E_NOMEM;
  v->lat = *lattice_h;
  v->n = n;
  long long chunk = chunk_voxels > 0 ? chunk_voxels : (1ll << 26);
  chunk = std::min<long long>((chunk + 7) / 8 * 8, (n + 7) / 8 * 8);
  v->chunk = chunk;
  v->cand_fraction = candidate_fraction;
  cudaError_t e = cudaGetDevice(&v->device);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&v->d_reff_mm), align_up(sizeof(uint16_t) * (size_t)n, 256));
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&v->d_w), sizeof(double) * (size_t)chunk);
  if (e != cudaSuccess) {
    com::set_error("com_volume_create: %s", cudaGetErrorString(e));
    cudaGetLastError();
    com_volume_destroy(v);
    return e == cudaErrorMemoryAllocation ? COM_E_NOMEM : COM_E_CUDA;
  }
  *out = v;
  return COM_OK;
}

int com_volume_destroy(ComVolume* v) {
  if (!v) return COM_OK;
  if (v->d_reff_mm) cudaFree(v->d_reff_mm);
  if (v->d_w) cudaFree(v->d_w);
  if (v->d_ws) cudaFree(v->d_ws);
  if (v->d_src) cudaFree(v->d_src);
  delete v;
  return COM_OK;
}

int64_t com_volume_voxels(const ComVolume* v) { return v ? v->n : 0; }
int64_t com_volume_chunk(const ComVolume* v) { return v ? v->chunk : 0; }
const uint16_t* com_volume_reff_mm(const ComVolume* v) { return v ? v->d_reff_mm : nullptr; }

int com_volume_set_reff_resampled(ComVolume* v, const double* src_reff, int32_t src_is_host, int64_t src_nx,
                                  int64_t src_ny, int64_t src_nz, com_stream_t stream) {
  if (!v || !src_reff || src_nx < 2 || src_ny < 2 || src_nz < 2) {
    com::set_error("com_volume_set_reff_resampled: bad arguments");
    return COM_E_BADARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const double* d_src = src_reff;
  if (src_is_host) {
    const size_t bytes = sizeof(double) * (size_t)(src_nx * src_ny * src_nz);
    if (bytes > v->src_bytes) {
      if (v->d_src) COM_CUDA(cudaFree(v->d_src));
      v->d_src = nullptr, v->src_bytes = 0;
      COM_CUDA(cudaMalloc(reinterpret_cast<void**>(&v->d_src), bytes));
      v->src_bytes = bytes;
    }
    COM_CUDA(cudaMemcpyAsync(v->d_src, src_reff, bytes, cudaMemcpyHostToDevice, s));
    d_src = v->d_src;
  }
  // ranges of < 2^31 * 256 voxels per launch: one launch is enough for any lattice that fits a GPU
  int rc = com_reff_resample(&v->lat, 0, v->n, d_src, src_nx, src_ny, src_nz, v->d_reff_mm, nullptr, stream);
  if (rc) return rc;
  v->reff_set = true;
  return COM_OK;
}

int com_volume_set_reff_quadratic(ComVolume* v, const double* coef_h, double reff_max, com_stream_t stream) {
  if (!v || !coef_h) {
    com::set_error("com_volume_set_reff_quadratic: bad arguments");
    return COM_E_BADARG;
  }
  int rc = com_reff_quadratic(&v->lat, 0, v->n, coef_h, reff_max, v->d_reff_mm, nullptr, stream);
  if (rc) return rc;
  v->reff_set = true;
  return COM_OK;
}

int com_volume_histograms(ComVolume* v, int32_t n_channels, const ComScene* const* scenes, double* hist_mm,
                          ComStats* stats, com_stream_t stream) {
  if (!v || n_channels < 1 || !scenes || !hist_mm || !stats) {
    com::set_error("com_volume_histograms: bad arguments");
    return COM_E_BADARG;
  }
  if (!v->reff_set) {
    com::set_error("com_volume_histograms: the volume has no Reff bins yet (com_volume_set_reff_*)");
    return COM_E_BADARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // candidate list sized for the largest scene
  size_t need = 0;
  for (int c = 0; c < n_channels; ++c) {
    if (!scenes[c]) {
      com::set_error("com_volume_histograms: null scene handle for channel %d", c);
      return COM_E_BADARG;
    }
    need = std::max(need, com::visibility_workspace_bytes(scenes[c]->dev, v->chunk, v->cand_fraction));
  }
  if (need > v->ws_bytes) {
    if (v->d_ws) COM_CUDA(cudaFree(v->d_ws));
    v->d_ws = nullptr, v->ws_bytes = 0;
    cudaError_t e = cudaMalloc(&v->d_ws, need);
    if (e != cudaSuccess) {
      com::set_error("com_volume_histograms: cudaMalloc(%zu) for the candidate list: %s", need, cudaGetErrorString(e));
      cudaGetLastError();
      return COM_E_NOMEM;
    }
    v->ws_bytes = need;
  }
  COM_CUDA(cudaMemsetAsync(hist_mm, 0, sizeof(double) * COM_MM_BINS * (size_t)n_channels, s));
  COM_CUDA(cudaMemsetAsync(stats, 0, sizeof(ComStats) * (size_t)n_channels, s));
  com::visibility_accumulate_stats(true);  // the launches add to stats[c] instead of overwriting it
  int rc = COM_OK;
  for (int c = 0; c < n_channels && rc == COM_OK; ++c) {
    for (long long lo = 0; lo < v->n && rc == COM_OK; lo += v->chunk) {
      const long long hi = std::min(v->n, lo + v->chunk);
      // No per-voxel ray counts here: a voxel is visible iff its weight is non-zero (every passing ray adds a strictly
      // positive weight - the set the Stage-2 histogram uses anyway), so the 10 B/voxel pass counts the visible voxels
      // and Stage 1 issues one atomic per ray instead of two.
      rc = com::visibility_launch(scenes[c], &v->lat, nullptr, lo, hi, v->d_w, nullptr, nullptr, 0, nullptr, nullptr,
                                  nullptr, 0, &stats[c], v->d_ws, v->ws_bytes, s);
      if (rc == COM_OK)
        rc = com::weight_histogram_mm_count(v->d_reff_mm + lo, v->d_w, hi - lo, nullptr, hist_mm + (size_t)c * COM_MM_BINS,
                                            reinterpret_cast<unsigned long long*>(&stats[c].visible_voxels), s);
    }
  }
  com::visibility_accumulate_stats(false);
  return rc;
}

int com_volume_histograms_host(ComVolume* v, int32_t n_channels, const ComScene* const* scenes, const double* src_reff_h,
                               int64_t src_nx, int64_t src_ny, int64_t src_nz, double* hist_mm_h, ComStats* stats_h) {
  if (!v || n_channels < 1 || !hist_mm_h || !stats_h) {
    com::set_error("com_volume_histograms_host: bad arguments");
    return COM_E_BADARG;
  }
  cudaStream_t s = cudaStreamPerThread;
  int rc = COM_OK;
  if (src_reff_h) {
    rc = com_volume_set_reff_resampled(v, src_reff_h, 1, src_nx, src_ny, src_nz, s);
    if (rc) return rc;
  }
  const size_t b_hist = sizeof(double) * COM_MM_BINS * (size_t)n_channels;
  const size_t b_stats = align_up(sizeof(ComStats) * (size_t)n_channels, 256);
  char *d = nullptr, *pin = nullptr;
  rc = com::HostArena::reserve(b_hist + b_stats, b_hist + b_stats, &d, &pin);
  if (rc) return rc;
  double* d_hist = reinterpret_cast<double*>(d);
  ComStats* d_stats = reinterpret_cast<ComStats*>(d + b_hist);
  rc = com_volume_histograms(v, n_channels, scenes, d_hist, d_stats, s);
  if (rc) return rc;
  COM_CUDA(cudaMemcpyAsync(pin, d, b_hist + b_stats, cudaMemcpyDeviceToHost, s));  // one copy, one wait
  rc = com::host_call_wait(s);
  if (rc) return rc;
  std::memcpy(hist_mm_h, pin, b_hist);
  std::memcpy(stats_h, pin + b_hist, sizeof(ComStats) * (size_t)n_channels);
  for (int c = 0; c < n_channels; ++c)
    if (stats_h[c].overflow) {
      com::set_error("candidate list overflow in channel %d: raise candidate_fraction of com_volume_create", c);
      return COM_E_OVERFLOW;
    }
  return COM_OK;
}

int com_histograms_flux_host(int32_t n_channels, const ComTables* const* tables, const double* profiles_h,
                             int32_t n_slices, int32_t K, double concentration, const double* hist_mm_h,
                             double* flux_out_h, int32_t flux_stride, double* boundary_out_h) {
  if (n_channels < 1 || !tables || !profiles_h || n_slices < 1 || K < 1 || !hist_mm_h || !flux_out_h || flux_stride < 2) {
    com::set_error("com_histograms_flux_host: bad arguments");
    return COM_E_BADARG;
  }
  int sorted = 1;
  double pmax = profiles_h[0];
  for (int k = 0; k < K; ++k) {
    pmax = std::max(pmax, profiles_h[3 * (size_t)k]);
    if (k && !(profiles_h[3 * (size_t)k] >= profiles_h[3 * (size_t)(k - 1)])) sorted = 0;
  }
  const size_t b_prof = align_up(sizeof(double) * 3 * (size_t)K * n_slices, 256);
  const size_t b_hist = sizeof(double) * COM_MM_BINS * (size_t)n_channels;
  const size_t b_flux = align_up(sizeof(double) * (size_t)n_channels * n_slices * flux_stride, 256);
  const size_t b_bnd = align_up(sizeof(double) * (size_t)n_channels, 256);
  char *d = nullptr, *pin = nullptr;
  int rc = com::HostArena::reserve(b_prof + b_hist + b_flux + b_bnd, b_prof + b_hist + b_flux + b_bnd, &d, &pin);
  if (rc) return rc;
  cudaStream_t s = cudaStreamPerThread;
  std::memcpy(pin, profiles_h, sizeof(double) * 3 * (size_t)K * n_slices);
  std::memcpy(pin + b_prof, hist_mm_h, b_hist);
  COM_CUDA(cudaMemcpyAsync(d, pin, b_prof + b_hist, cudaMemcpyHostToDevice, s));
  double* d_flux = reinterpret_cast<double*>(d + b_prof + b_hist);
  double* d_bnd = reinterpret_cast<double*>(d + b_prof + b_hist + b_flux);
  COM_CUDA(cudaMemsetAsync(d_flux, 0, b_flux + b_bnd, s));
  rc = com_histograms_flux(n_channels, tables, reinterpret_cast<const double*>(d), n_slices, K, sorted, pmax, concentration,
                           reinterpret_cast<const double*>(d + b_prof), d_flux, flux_stride, d_bnd, s);
  if (rc) return rc;
  COM_CUDA(cudaMemcpyAsync(pin + b_prof + b_hist, d_flux, b_flux + b_bnd, cudaMemcpyDeviceToHost, s));
  rc = com::host_call_wait(s);
  if (rc) return rc;
  std::memcpy(flux_out_h, pin + b_prof + b_hist, sizeof(double) * (size_t)n_channels * n_slices * flux_stride);
  if (boundary_out_h) std::memcpy(boundary_out_h, pin + b_prof + b_hist + b_flux, sizeof(double) * (size_t)n_channels);
  return COM_OK;
}

}  // extern "C"
