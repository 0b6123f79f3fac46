// C ABI glue: scene lifetime, stream-ordered entry points, host-buffer variants, visible-voxel compaction.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "scene.h"

namespace com {
int voxels_pack(const ComScene*, const double*, long long, long long, float4*, cudaStream_t);
int visibility_launch_packed(const ComScene*, const ComLattice*, const double*, const float4*, long long, long long, double*,
                             unsigned int*, ComRayRecord*, long long, unsigned long long*, const uint16_t*, double*, int,
                             ComStats*, void*, size_t, cudaStream_t);
int visibility_launch(const ComScene*, const ComLattice*, const double*, long long, long long, double*, unsigned int*,
                      ComRayRecord*, long long, unsigned long long*, const uint16_t*, double*, int, ComStats*, void*,
                      size_t, cudaStream_t);
size_t visibility_workspace_bytes(const DeviceScene&, long long, double);
long long lattice_local_voxels(const ComLattice&);
void kernel_timing_enable(int on);
void kernel_timing_read(double*, double*, long long*);
void kernel_timing_read3(double*, long long*);
int host_call_wait(cudaStream_t s);  // blocking (non-spinning) wait for the stream, volume.cu
}  // namespace com

#define COM_CUDA(x)                                                            \
  do {                                                                         \
    cudaError_t e__ = (x);                                                     \
    if (e__ != cudaSuccess) {                                                  \
      com::set_error("%s failed: %s", #x, cudaGetErrorString(e__));            \
      return COM_E_CUDA;                                                       \
    }                                                                          \
  } while (0)

namespace com {
namespace {
// One arena per host thread: the *_host entry points may be called concurrently from several threads (one channel
// per thread), each on its own per-thread CUDA stream.
struct ThreadArena {
  char* dev = nullptr;
  size_t dev_cap = 0;
  int device = -1;
  char* pinned = nullptr;
  size_t pinned_cap = 0;
  void release() {
    // at process exit the CUDA context may already be gone: the calls then fail harmlessly
    if (dev) cudaFree(dev);
    if (pinned) cudaFreeHost(pinned);
    dev = pinned = nullptr;
    dev_cap = pinned_cap = 0;
    device = -1;
    cudaGetLastError();
  }
  ~ThreadArena() { release(); }  // a worker thread that exits gives its blocks back
};
thread_local ThreadArena g_arena;
#define g_arena_dev g_arena.dev
#define g_arena_dev_cap g_arena.dev_cap
#define g_arena_device g_arena.device
#define g_arena_pinned g_arena.pinned
#define g_arena_pinned_cap g_arena.pinned_cap
}  // namespace

void HostArena::release_all() { g_arena.release(); }

int HostArena::reserve(size_t device_bytes, size_t pinned_bytes, char** device_out, char** pinned_out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return COM_E_CUDA;
  }
  if (dev != g_arena_device) {
    release_all();
    g_arena_device = dev;
  }
  if (device_bytes > g_arena_dev_cap) {
    if (g_arena_dev) cudaFree(g_arena_dev);
    g_arena_dev = nullptr, g_arena_dev_cap = 0;
    const size_t want = device_bytes + device_bytes / 4;
    e = cudaMalloc(reinterpret_cast<void**>(&g_arena_dev), want);
    if (e != cudaSuccess) {
      set_error("host-call arena: cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
      cudaGetLastError();
      return e == cudaErrorMemoryAllocation ? COM_E_NOMEM : COM_E_CUDA;
    }
    g_arena_dev_cap = want;
  }
  if (pinned_bytes > g_arena_pinned_cap) {
    if (g_arena_pinned) cudaFreeHost(g_arena_pinned);
    g_arena_pinned = nullptr, g_arena_pinned_cap = 0;
    const size_t want = pinned_bytes + pinned_bytes / 4;
    e = cudaMallocHost(reinterpret_cast<void**>(&g_arena_pinned), want);
    if (e != cudaSuccess) {
      set_error("host-call arena: cudaMallocHost(%zu): %s", want, cudaGetErrorString(e));
      cudaGetLastError();
      return COM_E_NOMEM;
    }
    g_arena_pinned_cap = want;
  }
  if (device_out) *device_out = g_arena_dev;
  if (pinned_out) *pinned_out = g_arena_pinned;
  return COM_OK;
}
}  // namespace com

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Scene tables as one blob; offsets[i] are the section starts (256-byte aligned).
struct SceneBlob {
  std::vector<char> bytes;
  size_t offset[8];
};

void make_scene_blob(const com::HostScene& hs, SceneBlob& b) {
  const void* src[8] = {hs.crystal_xyz.data(), hs.crystal_normal.data(), hs.dev_apertures.data(), hs.edges.data(),
                        hs.det_forms.data(),   hs.slit_forms.data(),     hs.port_forms.data(),    hs.refl_table.data()};
  const size_t bytes[8] = {hs.crystal_xyz.size() * sizeof(double),       hs.crystal_normal.size() * sizeof(double),
                           hs.dev_apertures.size() * sizeof(com::DevAperture), hs.edges.size() * sizeof(double),
                           hs.det_forms.size() * sizeof(float),          hs.slit_forms.size() * sizeof(float),
                           hs.port_forms.size() * sizeof(float),         hs.refl_table.size() * sizeof(double)};
  size_t total = 0;
  for (int i = 0; i < 8; ++i) {
    b.offset[i] = total;
    total += align_up(bytes[i], 256);
  }
  b.bytes.assign(total, 0);
  for (int i = 0; i < 8; ++i) std::memcpy(b.bytes.data() + b.offset[i], src[i], bytes[i]);
}

void bind_scene(com::DeviceScene& dev, char* d, const SceneBlob& b) {
  dev.crystal_xyz = reinterpret_cast<const double*>(d + b.offset[0]);
  dev.crystal_normal = reinterpret_cast<const double*>(d + b.offset[1]);
  dev.apertures = reinterpret_cast<const com::DevAperture*>(d + b.offset[2]);
  dev.edges = reinterpret_cast<const double*>(d + b.offset[3]);
  dev.det_forms = reinterpret_cast<const float4*>(d + b.offset[4]);
  dev.slit_forms = reinterpret_cast<const float4*>(d + b.offset[5]);
  dev.port_forms = reinterpret_cast<const float4*>(d + b.offset[6]);
  dev.refl_table = reinterpret_cast<const double*>(d + b.offset[7]);
}

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
                                const unsigned long long* offsets, long long* idx_out, double* w_out,
                                unsigned long long capacity) {
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
    if (pos < capacity) {
      idx_out[pos] = vox_begin + i;
      if (w_out) w_out[pos] = w[i];
    }
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

  SceneBlob blob;
  make_scene_blob(hs, blob);
  cudaError_t e = cudaGetDevice(&sc->device);
  if (e == cudaSuccess) e = cudaMalloc(&sc->d_blob, blob.bytes.size());
  if (e == cudaSuccess) e = cudaMemcpy(sc->d_blob, blob.bytes.data(), blob.bytes.size(), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    com::set_error("com_scene_create: %s", cudaGetErrorString(e));
    if (sc->d_blob) cudaFree(sc->d_blob);
    delete sc;
    return COM_E_CUDA;
  }
  bind_scene(sc->dev, static_cast<char*>(sc->d_blob), blob);
  *out = sc;
  return COM_OK;
}

int64_t com_lattice_local_voxels(const ComLattice* lattice_h) {
  return lattice_h ? com::lattice_local_voxels(*lattice_h) : 0;
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

int com_scene_set_detector_image(ComScene* scene, double* pix_out, int32_t nx, int32_t ny, double pixel_mm, double x0_mm,
                                 double y0_mm) {
  if (!scene || (pix_out && (nx < 1 || ny < 1 || !(pixel_mm > 0)))) {
    com::set_error("com_scene_set_detector_image: bad arguments");
    return COM_E_BADARG;
  }
  scene->dev.pix_out = pix_out;
  scene->dev.pix_nx = nx, scene->dev.pix_ny = ny;
  scene->dev.pix_mm = pixel_mm, scene->dev.pix_x0 = x0_mm, scene->dev.pix_y0 = y0_mm;
  return COM_OK;
}

int com_scene_filter_tables_host(const ComSceneDesc* desc_h, int32_t* n_crystal_padded, float* det_forms_out,
                                 float* slit_forms_out, float* rects_out, float* period_out, int32_t* flags_out) {
  if (!desc_h || !n_crystal_padded) {
    com::set_error("com_scene_filter_tables_host: null argument");
    return COM_E_BADARG;
  }
  com::HostScene hs;
  int rc = com::build_host_scene(*desc_h, hs);
  if (rc) return rc;
  *n_crystal_padded = hs.dev.n_crystal_padded;
  if (det_forms_out) std::memcpy(det_forms_out, hs.det_forms.data(), hs.det_forms.size() * sizeof(float));
  if (slit_forms_out) std::memcpy(slit_forms_out, hs.slit_forms.data(), hs.slit_forms.size() * sizeof(float));
  if (rects_out) {
    const com::SlitRect r[2] = {hs.dev.rect_crys, hs.dev.rect_plasma};
    std::memcpy(rects_out, r, sizeof(r));
  }
  if (period_out) *period_out = hs.dev.slit_period;
  if (flags_out) flags_out[0] = hs.dev.slit_filter, flags_out[1] = hs.dev.slit_same_w, flags_out[2] = hs.dev.filter_disabled;
  return COM_OK;
}

int com_scene_filter_certain_host(const ComSceneDesc* desc_h, float* out9) {
  if (!desc_h || !out9) {
    com::set_error("com_scene_filter_certain_host: null argument");
    return COM_E_BADARG;
  }
  com::HostScene hs;
  int rc = com::build_host_scene(*desc_h, hs);
  if (rc) return rc;
  const com::SlitRect r[2] = {hs.dev.inner_crys, hs.dev.inner_plasma};
  std::memcpy(out9, r, sizeof(r));
  out9[8] = hs.dev.certain_mm;
  return COM_OK;
}

int com_scene_interval_tables_host(const ComSceneDesc* desc_h, float* rects12, int32_t* interval_filter,
                                   float* port_forms_out, float* slit_edges_out, int32_t* n_slit_edges_out,
                                   int32_t* certain_off_out) {
  if (!desc_h) {
    com::set_error("com_scene_interval_tables_host: null argument");
    return COM_E_BADARG;
  }
  com::HostScene hs;
  int rc = com::build_host_scene(*desc_h, hs);
  if (rc) return rc;
  if (rects12) {
    const com::SlitRect r[3] = {hs.dev.det_inner, hs.dev.port_outer, hs.dev.port_inner};
    std::memcpy(rects12, r, sizeof(r));
  }
  if (interval_filter) *interval_filter = hs.dev.interval_filter;
  if (port_forms_out) std::memcpy(port_forms_out, hs.port_forms.data(), hs.port_forms.size() * sizeof(float));
  if (slit_edges_out) std::memcpy(slit_edges_out, hs.dev.slit_edge, sizeof(hs.dev.slit_edge));
  if (n_slit_edges_out) n_slit_edges_out[0] = hs.dev.n_slit_edge[0], n_slit_edges_out[1] = hs.dev.n_slit_edge[1];
  if (certain_off_out) *certain_off_out = hs.dev.certain_off;
  return COM_OK;
}

int com_scene_port_edges_host(const ComSceneDesc* desc_h, float* edges_out64, int32_t* n_edges_out) {
  if (!desc_h || !edges_out64 || !n_edges_out) {
    com::set_error("com_scene_port_edges_host: null argument");
    return COM_E_BADARG;
  }
  com::HostScene hs;
  int rc = com::build_host_scene(*desc_h, hs);
  if (rc) return rc;
  std::memcpy(edges_out64, hs.dev.port_edge, sizeof(hs.dev.port_edge));
  *n_edges_out = hs.dev.n_port_edge;
  return COM_OK;
}

int com_scene_detector_edges_host(const ComSceneDesc* desc_h, float* edges_out32, int32_t* n_edges_out) {
  if (!desc_h || !edges_out32 || !n_edges_out) {
    com::set_error("com_scene_detector_edges_host: null argument");
    return COM_E_BADARG;
  }
  com::HostScene hs;
  int rc = com::build_host_scene(*desc_h, hs);
  if (rc) return rc;
  std::memcpy(edges_out32, hs.dev.det_edge, sizeof(hs.dev.det_edge));
  *n_edges_out = hs.dev.n_det_edge;
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

int com_voxels_pack(const ComScene* scene, const double* vox_xyz, int64_t vox_begin, int64_t vox_end, void* vox_packed_out,
                    com_stream_t stream) {
  return com::voxels_pack(scene, vox_xyz, vox_begin, vox_end, reinterpret_cast<float4*>(vox_packed_out),
                          static_cast<cudaStream_t>(stream));
}

int com_visibility_weights_packed(const ComScene* scene, const double* vox_xyz, const void* vox_packed, int64_t vox_begin,
                                  int64_t vox_end, double* w_out, uint32_t* nrays_out, ComRayRecord* rays_out,
                                  int64_t rays_capacity, uint64_t* rays_count_out, const uint16_t* reff_bin,
                                  double* hist_out, int32_t n_bins, ComStats* stats_out, void* workspace,
                                  size_t workspace_bytes, com_stream_t stream) {
  if (!vox_xyz || !vox_packed) {
    com::set_error("com_visibility_weights_packed: both voxel arrays are needed (fp64 for the decision, packed for the filter)");
    return COM_E_BADARG;
  }
  return com::visibility_launch_packed(scene, nullptr, vox_xyz, reinterpret_cast<const float4*>(vox_packed), vox_begin, vox_end,
                                       w_out, nrays_out, rays_out, rays_capacity,
                                       reinterpret_cast<unsigned long long*>(rays_count_out), reff_bin, hist_out, n_bins,
                                       stats_out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
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

// Shared body of the host-buffer Stage-1 calls.  dense: per-voxel arrays come back; otherwise the visible voxels are
// compacted on the device straight into pinned host memory (zero-copy stores) and only (flat index, weight) pairs
// come back.  scene_h != NULL: the caller's scene handle is used as it is (nothing is rebuilt or uploaded).
static int compact_visible_capped(const uint32_t* nrays, const double* w, int64_t n, int64_t vox_begin, int64_t* idx_out,
                                  double* w_compact, int64_t capacity, uint64_t* count_out, void* workspace,
                                  size_t workspace_bytes, cudaStream_t s);

static int visibility_host_impl(const ComScene* scene_h, const ComSceneDesc* desc_h, const ComLattice* lattice_h,
                                const double* vox_xyz_h, int64_t vox_begin, int64_t vox_end, bool dense, double* w_out_h,
                                uint32_t* nrays_out_h, int64_t* idx_out_h, int64_t visible_capacity,
                                uint64_t* visible_count_h, ComRayRecord* rays_out_h, int64_t rays_capacity,
                                uint64_t* rays_count_out_h, ComStats* stats_out_h) {
  com::HostScene hs;
  SceneBlob blob;
  ComScene local;  // description given instead of a handle: tables live in the arena, nothing to free
  const ComScene* scene = scene_h;
  if (!scene) {
    int rc0 = com::build_host_scene(*desc_h, hs);
    if (rc0) return rc0;
    make_scene_blob(hs, blob);
    local.dev = hs.dev;
    local.filter_eps_mm = desc_h->filter_eps_mm;
    scene = &local;
  }
  int rc = COM_OK;
  const int64_t n = vox_end - vox_begin;
  const int64_t cap = dense ? 0 : std::min<int64_t>(visible_capacity, n);
  const size_t ws_bytes = com::visibility_workspace_bytes(scene->dev, n, 0.0);
  const size_t cws_bytes = dense ? 0 : com_compact_workspace_bytes(n);
  const size_t b_scene = align_up(blob.bytes.size(), 256);
  const size_t b_w = align_up(sizeof(double) * (size_t)n, 256), b_n = align_up(sizeof(uint32_t) * (size_t)n, 256);
  const size_t b_vox = (!lattice_h) ? align_up(sizeof(double) * 3 * (size_t)n, 256) : 0;
  const size_t b_vox4 = (!lattice_h) ? align_up(16 * (size_t)n, 256) : 0;  // the same voxels packed for the filter
  const size_t b_rays = rays_out_h ? align_up(sizeof(ComRayRecord) * (size_t)rays_capacity, 256) : 0;
  const size_t b_misc = 512;
  const size_t b_cap = align_up(sizeof(int64_t) * (size_t)cap, 256);  // per compacted array, in pinned memory
  const size_t total = b_scene + b_w + b_n + b_vox + b_vox4 + b_rays + b_misc + align_up(cws_bytes, 256) + ws_bytes;
  char *d = nullptr, *pin = nullptr;
  rc = com::HostArena::reserve(total, b_scene + 512 + 2 * b_cap, &d, &pin);
  if (rc) return rc;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    char* p = d + o;
    o += bytes;
    return p;
  };
  char* d_scene = take(b_scene);
  double* d_w = reinterpret_cast<double*>(take(b_w));
  uint32_t* d_n = reinterpret_cast<uint32_t*>(take(b_n));
  double* d_vox = b_vox ? reinterpret_cast<double*>(take(b_vox)) : nullptr;
  float4* d_vox4 = b_vox4 ? reinterpret_cast<float4*>(take(b_vox4)) : nullptr;
  ComRayRecord* d_rays = b_rays ? reinterpret_cast<ComRayRecord*>(take(b_rays)) : nullptr;
  char* d_misc = take(b_misc);
  ComStats* d_stats = reinterpret_cast<ComStats*>(d_misc);
  uint64_t* d_rcount = reinterpret_cast<uint64_t*>(d_misc + 256);
  uint64_t* d_vcount = reinterpret_cast<uint64_t*>(d_misc + 264);
  void* d_cws = take(align_up(cws_bytes, 256));
  void* d_ws = take(ws_bytes);
  char* pin_misc = pin + b_scene;
  int64_t* pin_idx = reinterpret_cast<int64_t*>(pin + b_scene + 512);
  double* pin_w = reinterpret_cast<double*>(pin + b_scene + 512 + b_cap);

  cudaStream_t s = cudaStreamPerThread;  // concurrent host threads overlap on the device
  if (!scene_h) {
    std::memcpy(pin, blob.bytes.data(), blob.bytes.size());
    COM_CUDA(cudaMemcpyAsync(d_scene, pin, blob.bytes.size(), cudaMemcpyHostToDevice, s));
    bind_scene(local.dev, d_scene, blob);
  }
  const double* d_vox_base = nullptr;
  const float4* d_vox4_base = nullptr;
  if (d_vox) {  // explicit voxels: the caller passes rows [vox_begin, vox_end) only; shift so that flat indexing works
    COM_CUDA(cudaMemcpyAsync(d_vox, vox_xyz_h, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, s));
    d_vox_base = d_vox - 3 * vox_begin;
    rc = com::voxels_pack(scene, d_vox_base, vox_begin, vox_end, d_vox4 - vox_begin, s);
    if (rc) return rc;
    d_vox4_base = d_vox4 - vox_begin;
  }
  rc = com::visibility_launch_packed(scene, lattice_h, d_vox_base, d_vox4_base, vox_begin, vox_end, d_w, d_n, d_rays,
                                     rays_capacity, d_rays ? reinterpret_cast<unsigned long long*>(d_rcount) : nullptr,
                                     nullptr, nullptr, 0, d_stats, d_ws, ws_bytes, s);
  if (rc) return rc;
  if (dense) {
    COM_CUDA(cudaMemcpyAsync(w_out_h, d_w, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    if (nrays_out_h) COM_CUDA(cudaMemcpyAsync(nrays_out_h, d_n, sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToHost, s));
  } else {
    // pinned host memory is addressable from the device: the scatter kernel stores the compacted pairs there
    rc = compact_visible_capped(d_n, d_w, n, vox_begin, pin_idx, pin_w, cap, d_vcount, d_cws, cws_bytes, s);
    if (rc) return rc;
  }
  COM_CUDA(cudaMemcpyAsync(pin_misc, d_misc, 512, cudaMemcpyDeviceToHost, s));
  rc = com::host_call_wait(s);  // the only wait of the call unless a ray list was asked for
  if (rc) return rc;
  std::memcpy(stats_out_h, pin_misc, sizeof(ComStats));
  if (!dense) {
    uint64_t count = 0;
    std::memcpy(&count, pin_misc + 264, sizeof(count));
    *visible_count_h = count;
    const size_t m = (size_t)std::min<uint64_t>(count, (uint64_t)cap);
    if (m) {
      std::memcpy(idx_out_h, pin_idx, sizeof(int64_t) * m);
      std::memcpy(w_out_h, pin_w, sizeof(double) * m);
    }
  }
  if (d_rays) {
    uint64_t count = 0;
    std::memcpy(&count, pin_misc + 256, sizeof(count));
    *rays_count_out_h = count;
    const size_t m = (size_t)std::min<uint64_t>(count, (uint64_t)rays_capacity);
    if (m) {
      COM_CUDA(cudaMemcpyAsync(rays_out_h, d_rays, sizeof(ComRayRecord) * m, cudaMemcpyDeviceToHost, s));
      rc = com::host_call_wait(s);
      if (rc) return rc;
    }
  }
  if (stats_out_h->overflow) {
    com::set_error("candidate list overflow: %llu candidates > capacity %llu",
                   (unsigned long long)stats_out_h->candidates, (unsigned long long)stats_out_h->candidate_capacity);
    return COM_E_OVERFLOW;
  }
  if (!dense && *visible_count_h > (uint64_t)visible_capacity) {
    com::set_error("%llu visible voxels > capacity %lld", (unsigned long long)*visible_count_h, (long long)visible_capacity);
    return COM_E_OVERFLOW;
  }
  if (stats_out_h->passing_rays == 0) {
    com::set_error("Not enough points to perform calculations!");
    return COM_E_EMPTY;
  }
  return COM_OK;
}

static bool dense_args_ok(const void* scene_or_desc, const ComLattice* lattice_h, const double* vox_xyz_h, int64_t vox_begin,
                          int64_t vox_end, const double* w_out_h, const ComRayRecord* rays_out_h,
                          const uint64_t* rays_count_out_h, const ComStats* stats_out_h) {
  return scene_or_desc && (lattice_h || vox_xyz_h) && w_out_h && stats_out_h && vox_end >= vox_begin &&
         !(rays_out_h && !rays_count_out_h);
}

static bool sparse_args_ok(const void* scene_or_desc, const ComLattice* lattice_h, const double* vox_xyz_h, int64_t vox_begin,
                           int64_t vox_end, const int64_t* idx_out_h, const double* w_out_h, int64_t capacity,
                           const uint64_t* count_out_h, const ComStats* stats_out_h) {
  return scene_or_desc && (lattice_h || vox_xyz_h) && idx_out_h && w_out_h && count_out_h && stats_out_h &&
         vox_end >= vox_begin && capacity >= 0;
}

int com_visibility_weights_host(const ComSceneDesc* desc_h, const ComLattice* lattice_h, const double* vox_xyz_h,
                                int64_t vox_begin, int64_t vox_end, double* w_out_h, uint32_t* nrays_out_h,
                                ComRayRecord* rays_out_h, int64_t rays_capacity, uint64_t* rays_count_out_h,
                                ComStats* stats_out_h) {
  if (!dense_args_ok(desc_h, lattice_h, vox_xyz_h, vox_begin, vox_end, w_out_h, rays_out_h, rays_count_out_h, stats_out_h)) {
    com::set_error("com_visibility_weights_host: bad arguments");
    return COM_E_BADARG;
  }
  return visibility_host_impl(nullptr, desc_h, lattice_h, vox_xyz_h, vox_begin, vox_end, true, w_out_h, nrays_out_h, nullptr,
                              0, nullptr, rays_out_h, rays_capacity, rays_count_out_h, stats_out_h);
}

int com_visibility_weights_host_scene(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz_h,
                                      int64_t vox_begin, int64_t vox_end, double* w_out_h, uint32_t* nrays_out_h,
                                      ComRayRecord* rays_out_h, int64_t rays_capacity, uint64_t* rays_count_out_h,
                                      ComStats* stats_out_h) {
  if (!dense_args_ok(scene, lattice_h, vox_xyz_h, vox_begin, vox_end, w_out_h, rays_out_h, rays_count_out_h, stats_out_h)) {
    com::set_error("com_visibility_weights_host_scene: bad arguments");
    return COM_E_BADARG;
  }
  return visibility_host_impl(scene, nullptr, lattice_h, vox_xyz_h, vox_begin, vox_end, true, w_out_h, nrays_out_h, nullptr,
                              0, nullptr, rays_out_h, rays_capacity, rays_count_out_h, stats_out_h);
}

int com_visible_weights_host(const ComSceneDesc* desc_h, const ComLattice* lattice_h, const double* vox_xyz_h,
                             int64_t vox_begin, int64_t vox_end, int64_t* idx_out_h, double* w_out_h,
                             int64_t capacity, uint64_t* count_out_h, ComStats* stats_out_h) {
  if (!sparse_args_ok(desc_h, lattice_h, vox_xyz_h, vox_begin, vox_end, idx_out_h, w_out_h, capacity, count_out_h, stats_out_h)) {
    com::set_error("com_visible_weights_host: bad arguments");
    return COM_E_BADARG;
  }
  return visibility_host_impl(nullptr, desc_h, lattice_h, vox_xyz_h, vox_begin, vox_end, false, w_out_h, nullptr, idx_out_h,
                              capacity, count_out_h, nullptr, 0, nullptr, stats_out_h);
}

int com_visible_weights_host_scene(const ComScene* scene, const ComLattice* lattice_h, const double* vox_xyz_h,
                                   int64_t vox_begin, int64_t vox_end, int64_t* idx_out_h, double* w_out_h,
                                   int64_t capacity, uint64_t* count_out_h, ComStats* stats_out_h) {
  if (!sparse_args_ok(scene, lattice_h, vox_xyz_h, vox_begin, vox_end, idx_out_h, w_out_h, capacity, count_out_h, stats_out_h)) {
    com::set_error("com_visible_weights_host_scene: bad arguments");
    return COM_E_BADARG;
  }
  return visibility_host_impl(scene, nullptr, lattice_h, vox_xyz_h, vox_begin, vox_end, false, w_out_h, nullptr, idx_out_h,
                              capacity, count_out_h, nullptr, 0, nullptr, stats_out_h);
}

int com_host_arena_release(void) {
  com::HostArena::release_all();
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

int com_kernel_timing_read3(double* ms3, int64_t* launches) {
  long long n = 0;
  com::kernel_timing_read3(ms3, &n);
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
  return compact_visible_capped(nrays, w, n, vox_begin, idx_out, w_compact, n, count_out, workspace, workspace_bytes,
                                static_cast<cudaStream_t>(stream));
}

}  // extern "C"

static int compact_visible_capped(const uint32_t* nrays, const double* w, int64_t n, int64_t vox_begin, int64_t* idx_out,
                                  double* w_compact, int64_t capacity, uint64_t* count_out, void* workspace,
                                  size_t workspace_bytes, cudaStream_t s) {
  if (!nrays || (!idx_out && capacity > 0) || !count_out || n < 0 || !workspace ||
      workspace_bytes < com_compact_workspace_bytes(n) || (w_compact && !w)) {
    com::set_error("com_compact_visible: bad arguments");
    return COM_E_BADARG;
  }
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
                                                                 reinterpret_cast<long long*>(idx_out), w_compact,
                                                                 (unsigned long long)capacity);
  COM_CUDA(cudaGetLastError());
  return COM_OK;
}
