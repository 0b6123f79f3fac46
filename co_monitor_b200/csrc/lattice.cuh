// Implicit CuboidMesh lattice on the device: flat voxel index -> (ix, iy, iz) -> fp64 position.
// Reference: co_monitor/geometry/mesh_calculation.py:119-159 (linspace per axis, meshgrid 'xy' + C ravel with z
// fastest, rotation about z, shift).  Shared by the visibility kernels and the Reff provider.
#pragma once

#include <cuda_runtime.h>

#include "../../include/comonitor_sm100.h"

namespace com {

struct D3 {
  double x, y, z;
};
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ D3 operator*(double s, D3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ double dot(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ D3 ld3(const double* p) { return {p[0], p[1], p[2]}; }

// Local column of a block-cyclic shard -> global column (identity when the lattice is not sharded).
__device__ __forceinline__ long long global_column(const ComLattice& L, long long cl) {
  if (L.shard_count <= 1 || L.shard_cols_per_block <= 0) return cl;
  const long long lb = cl / L.shard_cols_per_block, o = cl - lb * L.shard_cols_per_block;
  return (lb * L.shard_count + L.shard_index) * L.shard_cols_per_block + o;
}

// (shard-local) flat index -> global lattice indices
__device__ __forceinline__ void lattice_unravel(const ComLattice& L, long long flat, long long& ix, long long& iy,
                                                long long& iz) {
  if (flat < (1ll << 32) && L.shard_count <= 1) {  // 32-bit divisions are several times cheaper
    const unsigned int f = (unsigned int)flat, nz = (unsigned int)L.nz, nx = (unsigned int)L.nx;
    const unsigned int t = f / nz, y = t / nx;
    iz = f - t * nz, iy = y, ix = t - y * nx;
  } else {
    iz = flat % L.nz;
    const long long t = global_column(L, flat / L.nz);
    ix = t % L.nx, iy = t / L.nx;
  }
}

// CuboidMesh closed form (mesh_calculation.py:128-155); mul/add kept separate like numpy.linspace.
__device__ __forceinline__ D3 lattice_point(const ComLattice& L, long long ix, long long iy, long long iz) {
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

__device__ __forceinline__ D3 lattice_position(const ComLattice& L, long long flat) {
  long long ix, iy, iz;
  lattice_unravel(L, flat, ix, iy, iz);
  return lattice_point(L, ix, iy, iz);
}

// number of voxels of the (sharded) lattice; defined in visibility.cu
long long lattice_local_voxels(const ComLattice& L);

}  // namespace com
