// C ABI glue: scene lifetime, stream-ordered entry points, host-buffer variants, visible-voxel compaction.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "scene.h"

namespace com {
int visibility_launch(const ComScene*, const ComLattice*, const double*, long long, long long, double*, unsigned int*,
                      ComRayRecord*, long long, unsigned long long*, const uint16_t*, double*, int, ComStats*, void*,
                      size_t, cudaStream_t);
size_t visibility_workspace_bytes(const DeviceScene&, long long, double);
void kernel_timing_enable(int on);
void kernel_timing_read(double*, double*, long long*);
}  // namespace com

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

// ---- visible-voxel compaction: count per 1024-block, scan the block counts, scatter -------------------
constexpr int kCompactBlock = 1024;

__global__ void compact_count(const unsigned int* nrays, long long n, unsigned int* block_counts) {
  long long i = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
  int flag = (i < n && nrays[i] > 0) ? 1 : 0;
  int total = __syncthreads_count(flag);
  if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

__global__ void compact_scan(unsigned int* block_counts, long long n_blocks, unsigned long long* offsets,
                             unsigned long long* count_out) {
  // single CTA, sequential over chunks of blockDim.x; n_blocks <= ~1e6, negligible next to Stage 1
  __shared__ unsigned long long s_scan[1024];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long base = 0; base < n_blocks; base += blockDim.x) {
    long long i = base + threadIdx.x;
    unsigned long long v = i < n_blocks ? block_counts[i] : 0;
    s_scan[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
      unsigned long long t = threadIdx.x >= off ? s_scan[threadIdx.x - off] : 0;
      __syncthreads();
      s_scan[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_blocks) offsets[i] = s_carry + s_scan[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry += s_scan[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *count_out = s_carry;
}

__global__ void compact_scatter(const unsigned int* nrays, const double* w, long long n, long long vox_begin,
                                const unsigned long long* offsets, long long* idx_out, double* w_out) {
  __shared__ unsigned int s_warp[32];
  long long i = (long long)blockIdx.x * kCompactBlock + threadIdx.x;
  bool flag = i < n && nrays[i] > 0;
  unsigned int b = __ballot_sync(0xffffffffu, flag);
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_warp[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    unsigned int v = s_warp[lane], x = v;
    for (int off = 1; off < 32; off <<= 1) {
      unsigned int t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    s_warp[lane] = x - v;
  }
  __syncthreads();
  if (flag) {
    unsigned long long pos = offsets[blockIdx.x] + s_warp[warp] + __popc(b & ((1u << lane) - 1u));
    idx_out[pos] = vox_begin + i;
    if (w_out) w_out[pos] = w[i];
  }
}

// FP32 FMA issue-rate probe: 8 independent chains per thread, register operands only.
__global__ void __launch_bounds__(256) fma_probe(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fmaf(x0, a, b), x1 = fmaf(x1, a, b), x2 = fmaf(x2, a, b), x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b), x5 = fmaf(x5, a, b), x6 = fmaf(x6, a, b), x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

}  // namespace

extern "C" {

/* Measured FP32 FMA throughput of the current device in TFLOP/s (FMA = 2 FLOP); < 0 on error. */
double com_measure_fp32_tflops(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return -1.0;
  const int blocks = sms * 8, threads = 256, iters = 4096;
  float* d = nullptr;
  if (cudaMalloc(&d, sizeof(float) * blocks * threads) != cudaSuccess) return -1.0;
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    fma_probe<<<blocks, threads>>>(d, iters, 1.0000001f, 1e-7f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(a), cudaEventDestroy(b);
  cudaFree(d);
  return cudaGetLastError() == cudaSuccess ? best : -1.0;
}

int com_scene_create(const ComSceneDesc* desc_h, ComScene** out) {
  if (!desc_h || !out) {
    com::set_error("com_scene_create: null argument");
    return COM_E_BADARG;
  }
  *out = nullptr;
  com::HostScene hs;
  int rc = com::build_host_scene(*desc_h, hs);
  if (rc) return rc;
  ComScene* sc = new (std::nothrow) ComScene();
  if (!sc) return COM_E_NOMEM;
  sc->dev = hs.dev;
  sc->apertures_h = hs.apertures;
  sc->filter_eps_mm = desc_h->filter_eps_mm > 0 ? desc_h->filter_eps_mm : 5e-3;

  // one device allocation, sections 256-byte aligned
  struct Section {
    const void* src;
    size_t bytes;
    size_t offset;
  } sec[] = {
      {hs.crystal_xyz.data(), hs.crystal_xyz.size() * sizeof(double), 0},
      {hs.crystal_normal.data(), hs.crystal_normal.size() * sizeof(double), 0},
      {hs.dev_apertures.data(), hs.dev_apertures.size() * sizeof(com::DevAperture), 0},
      {hs.edges.data(), hs.edges.size() * sizeof(double), 0},
      {hs.det_forms.data(), hs.det_forms.size() * sizeof(float), 0},
      {hs.slit_forms.data(), hs.slit_forms.size() * sizeof(float), 0},
  };
  size_t total = 0;
  for (auto& x : sec) {
    x.offset = total;
    total += align_up(x.bytes, 256);
  }
  std::vector<char> blob(total, 0);
  for (auto& x : sec) std::memcpy(blob.data() + x.offset, x.src, x.bytes);

  cudaError_t e = cudaGetDevice(&sc->device);
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_blob, blob.size());
  if (e == cudaSuccess) e = cudaMemcpy(sc->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    com::set_error("com_scene_create: %s", cudaGetErrorString(e));
    if (sc->d_blob) cudaFree(sc->d_blob);
    delete sc;
    return COM_E_CUDA;
  }
  char* d = static_cast<char*>(sc->d_blob);
  sc->dev.crystal_xyz = reinterpret_cast<const double*>(d + sec[0].offset);
  sc->dev.crystal_normal = reinterpret_cast<const double*>(d + sec[1].offset);
  sc->dev.apertures = reinterpret_cast<const com::DevAperture*>(d + sec[2].offset);
  sc->dev.edges = reinterpret_cast<const double*>(d + sec[3].offset);
  sc->dev.det_forms = reinterpret_cast<const float4*>(d + sec[4].offset);
  sc->dev.slit_forms = reinterpret_cast<const float4*>(d + sec[5].offset);
  *out = sc;
  return COM_OK;
}

int com_scene_destroy(ComScene* scene) {
  if (!scene) return COM_OK;
  if (scene->d_blob) cudaFree(scene->d_blob);
  delete scene;
  return COM_OK;
}

int com_scene_aperture_count(const ComScene* scene) { return scene ? (int)scene->apertures_h.size() : COM_E_BADARG; }

int com_scene_get_aperture(const ComScene* scene, int32_t index, ComAperture* out_h) {
  if (!scene || !out_h || index < 0 || index >= (int)scene->apertures_h.size()) return COM_E_BADARG;
  *out_h = scene->apertures_h[index];
  return COM_OK;
}

int com_scene_filter_info(const ComScene* scene, int32_t* slit_filter, int32_t* filter_disabled, double* slit_period_mm,
                          double* eps_mm) {
  if (!scene) return COM_E_BADARG;
  if (slit_filter) *slit_filter = scene->dev.slit_filter;
  if (filter_disabled) *filter_disabled = scene->dev.filter_disabled;
  if (slit_period_mm) *slit_period_mm = scene->dev.slit_period_d;
  if (eps_mm) *eps_mm = scene->filter_eps_mm;
  return COM_OK;
}

size_t com_visibility_workspace_bytes(const ComScene* scene, int64_t n_voxels, double candidate_fraction) {
  if (!scene || n_voxels < 0) return 0;
  return com::visibility_workspace_bytes(scene->dev, n_voxels, candidate_fraction);
}

int com_visibility_weights(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz,
                           int64_t vox_begin, int64_t vox_end, double* w_out, uint32_t* nrays_out,
                           ComRayRecord* rays_out, int64_t rays_capacity, uint64_t* rays_count_out,
                           const uint16_t* reff_bin, double* hist_out, int32_t n_bins, ComStats* stats_out,
                           void* workspace, size_t workspace_bytes, com_stream_t stream) {
  return com::visibility_launch(scene, lattice_h, vox_xyz, vox_begin, vox_end, w_out, nrays_out, rays_out,
                                rays_capacity, reinterpret_cast<unsigned long long*>(rays_count_out), reff_bin,
                                hist_out, n_bins, stats_out, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

int com_visibility_weights_host(const ComSceneDesc* desc_h, const ComLattice* lattice_h, const double* vox_xyz_h,
                                int64_t vox_begin, int64_t vox_end, double* w_out_h, uint32_t* nrays_out_h,
                                ComRayRecord* rays_out_h, int64_t rays_capacity, uint64_t* rays_count_out_h,
                                ComStats* stats_out_h) {
  if (!desc_h || (!lattice_h && !vox_xyz_h) || !w_out_h || !stats_out_h || vox_end < vox_begin ||
      (rays_out_h && !rays_count_out_h)) {
    com::set_error("com_visibility_weights_host: bad arguments");
    return COM_E_BADARG;
  }
  ComScene* scene = nullptr;
  int rc = com_scene_create(desc_h, &scene);
  if (rc) return rc;
  const int64_t n = vox_end - vox_begin;
  const size_t ws_bytes = com_visibility_workspace_bytes(scene, n, 0.0);
  char* d_all = nullptr;
  const size_t b_w = align_up(sizeof(double) * (size_t)n, 256), b_n = align_up(sizeof(uint32_t) * (size_t)n, 256);
  const size_t b_vox = (!lattice_h) ? align_up(sizeof(double) * 3 * (size_t)n, 256) : 0;
  const size_t b_rays = rays_out_h ? align_up(sizeof(ComRayRecord) * (size_t)rays_capacity, 256) : 0;
  const size_t b_misc = 512;
  cudaError_t e = cudaMalloc(&d_all, b_w + b_n + b_vox + b_rays + b_misc + ws_bytes);
  if (e != cudaSuccess) {
    com::set_error("com_visibility_weights_host: cudaMalloc: %s", cudaGetErrorString(e));
    com_scene_destroy(scene);
    return e == cudaErrorMemoryAllocation ? COM_E_NOMEM : COM_E_CUDA;
  }
  double* d_w = reinterpret_cast<double*>(d_all);
  uint32_t* d_n = reinterpret_cast<uint32_t*>(d_all + b_w);
  double* d_vox = b_vox ? reinterpret_cast<double*>(d_all + b_w + b_n) : nullptr;
  ComRayRecord* d_rays = b_rays ? reinterpret_cast<ComRayRecord*>(d_all + b_w + b_n + b_vox) : nullptr;
  ComStats* d_stats = reinterpret_cast<ComStats*>(d_all + b_w + b_n + b_vox + b_rays);
  uint64_t* d_rcount = reinterpret_cast<uint64_t*>(d_all + b_w + b_n + b_vox + b_rays + 256);
  void* d_ws = d_all + b_w + b_n + b_vox + b_rays + b_misc;
  auto fail = [&](int code) {
    cudaFree(d_all);
    com_scene_destroy(scene);
    return code;
  };
  // explicit voxels: the caller passes rows [vox_begin, vox_end) only; shift so that flat indexing works
  const double* d_vox_base = nullptr;
  if (d_vox) {
    e = cudaMemcpy(d_vox, vox_xyz_h, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      com::set_error("H2D voxels: %s", cudaGetErrorString(e));
      return fail(COM_E_CUDA);
    }
    d_vox_base = d_vox - 3 * vox_begin;
  }
  rc = com_visibility_weights(scene, lattice_h, d_vox_base, vox_begin, vox_end, d_w, nrays_out_h ? d_n : d_n, d_rays,
                              rays_capacity, d_rays ? d_rcount : nullptr, nullptr, nullptr, 0, d_stats, d_ws, ws_bytes,
                              nullptr);
  if (rc) return fail(rc);
  e = cudaMemcpy(w_out_h, d_w, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && nrays_out_h) e = cudaMemcpy(nrays_out_h, d_n, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(stats_out_h, d_stats, sizeof(ComStats), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && d_rays) {
    e = cudaMemcpy(rays_count_out_h, d_rcount, sizeof(uint64_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) {
      size_t m = (size_t)std::min<uint64_t>(*rays_count_out_h, (uint64_t)rays_capacity);
      e = cudaMemcpy(rays_out_h, d_rays, sizeof(ComRayRecord) * m, cudaMemcpyDeviceToHost);
    }
  }
  if (e != cudaSuccess) {
    com::set_error("com_visibility_weights_host: %s", cudaGetErrorString(e));
    return fail(COM_E_CUDA);
  }
  cudaFree(d_all);
  com_scene_destroy(scene);
  if (stats_out_h->overflow) {
    com::set_error("candidate list overflow: %llu candidates > capacity %llu",
                   (unsigned long long)stats_out_h->candidates, (unsigned long long)stats_out_h->candidate_capacity);
    return COM_E_OVERFLOW;
  }
  if (stats_out_h->passing_rays == 0) {
    com::set_error("Not enough points to perform calculations!");
    return COM_E_EMPTY;
  }
  return COM_OK;
}

int com_kernel_timing_enable(int on) {
  com::kernel_timing_enable(on);
  return COM_OK;
}

int com_kernel_timing_read(double* filter_ms, double* exact_ms, int64_t* launches) {
  long long n = 0;
  com::kernel_timing_read(filter_ms, exact_ms, &n);
  if (launches) *launches = n;
  return COM_OK;
}

size_t com_compact_workspace_bytes(int64_t n) {
  size_t blocks = (size_t)((n + kCompactBlock - 1) / kCompactBlock);
  return align_up(blocks * sizeof(unsigned int), 256) + align_up(blocks * sizeof(unsigned long long), 256) + 256;
}

int com_compact_visible(const uint32_t* nrays, const double* w, int64_t n, int64_t vox_begin, int64_t* idx_out,
                        double* w_compact, uint64_t* count_out, void* workspace, size_t workspace_bytes,
                        com_stream_t stream) {
  if (!nrays || !idx_out || !count_out || n < 0 || !workspace || workspace_bytes < com_compact_workspace_bytes(n) ||
      (w_compact && !w)) {
    com::set_error("com_compact_visible: bad arguments");
    return COM_E_BADARG;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    COM_CUDA(cudaMemsetAsync(count_out, 0, sizeof(uint64_t), s));
    return COM_OK;
  }
  const long long blocks = (n + kCompactBlock - 1) / kCompactBlock;
  unsigned int* counts = reinterpret_cast<unsigned int*>(workspace);
  unsigned long long* offsets = reinterpret_cast<unsigned long long*>(
      static_cast<char*>(workspace) + align_up((size_t)blocks * sizeof(unsigned int), 256));
  compact_count<<<(unsigned int)blocks, kCompactBlock, 0, s>>>(nrays, n, counts);
  compact_scan<<<1, 1024, 0, s>>>(counts, blocks, offsets, reinterpret_cast<unsigned long long*>(count_out));
  compact_scatter<<<(unsigned int)blocks, kCompactBlock, 0, s>>>(nrays, w, n, vox_begin, offsets,
                                                                 reinterpret_cast<long long*>(idx_out), w_compact);
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}

}  // extern "C"
