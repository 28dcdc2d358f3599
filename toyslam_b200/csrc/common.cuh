// common.cuh — shared device/host types of libndt_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdio.h>

// -DNDTB200_CHECKED: device-side bounds / protocol assertions in the hand-synchronised kernels (compute-sanitizer is
// closed on this GPU pool: the checked build + the parity suite is the memory-safety evidence, profiles/r02_checked_build.md)
#ifdef NDTB200_CHECKED
#define NDT_CHECK(cond)                                                                                          \
  do {                                                                                                           \
    if (!(cond)) {                                                                                               \
      printf("NDT_CHECK failed: %s at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, blockIdx.x, threadIdx.x); \
      __trap();                                                                                                  \
    }                                                                                                            \
  } while (0)
#else
#define NDT_CHECK(cond) do { } while (0)
#endif

namespace ndtb200 {

// ---------------------------------------------------------------------------------------------
// HBM layout of the target map (replaces std::map<size_t, Leaf>, voxel_grid_covariance_omp.h:98-201)
// ---------------------------------------------------------------------------------------------
// One 64-byte record per occupied voxel, sorted by voxel key (x fastest), so neighbouring cells
// along x are neighbouring records.  The fp64 mean (map-scale coordinates need more than fp32,
// SURVEY §7 hard part 2) is stored as an exact hi + lo pair of fp32 (mean = hi + lo to 2^-48), so the
// derivative pass forms x' - mean without touching the fp64 / conversion pipes:
// (x' - hi) - lo equals fl32(double(x') - mean) up to one ulp of the (sub-voxel-sized) result.
// icov is fp32 because the derivative pass casts it to fp32 anyway (ndt_omp_impl.hpp:493-494).
// Three 16-byte loads fetch the 48 bytes the hot loop needs.
struct __align__(16) VoxelRecord {
  float mean_hi[3];  // fl32(Leaf::mean_)
  float mean_lo[3];  // fl32(Leaf::mean_ - mean_hi)
  float icov[6];     // Leaf::icov_ as fp32: c00 c01 c02 c11 c12 c22
  int32_t key;       // linear voxel index (voxel_grid_covariance_omp_impl.hpp:223)
  int32_t count;     // Leaf::nr_points; -1 = rejected leaf (…_impl.hpp:337-341, 360-364)
  int32_t pad[2];
};
static_assert(sizeof(VoxelRecord) == 64, "VoxelRecord must be 64 bytes");

__host__ __device__ __forceinline__ double record_mean(const VoxelRecord& r, int a) {
  return static_cast<double>(r.mean_hi[a]) + static_cast<double>(r.mean_lo[a]);
}

// Open-addressing hash over the VALID voxels only (count >= min_points, not rejected): the only ones a
// lookup may return (…_impl.hpp:395).  Slot = {key (low 32), record index (high 32)}, linear probing,
// capacity = power of two >= 4 * n_valid (load <= 0.25).
typedef unsigned long long HashSlot;
#define NDTB200_HASH_EMPTY 0xFFFFFFFFFFFFFFFFull

__host__ __device__ __forceinline__ uint32_t hash_key(uint32_t key, int shift) {
  return (key * 0x9E3779B1u) >> shift;
}

struct GridDesc {  // pcl::VoxelGrid members used by the path
  int32_t min_b[3], max_b[3], div_b[3], mul[3];
  float leaf[3], inv_leaf[3];
  float min_p[3], max_p[3];
  long long ncell;     // dx*dy*dz from the int64 overflow guard expression
  int32_t overflow;    // 1 => guard tripped (…_impl.hpp:79-84)
  int32_t n_finite;    // points that passed the finite check
};

struct MapView {  // what the kernels need to probe the map
  const VoxelRecord* records;
  const double* icov64;  // [n_voxels][6] fp64 inverse covariance for the fp64 Hessian-only pass
  const HashSlot* hash;
  const int32_t* dense;  // direct-mapped cell table [dx*dy*dz] -> record index or -1 (nullptr: use the hash)
  // KDTREE mode (radius search over the voxel centroids, voxel_grid_covariance_omp.h:476-505): cell table over ALL
  // occupied voxels, the fp32 centroid of every voxel, squared radius
  const int32_t* cell_all;
  const float4* centroids;
  float kd_r2;
  uint32_t hash_mask;
  int32_t hash_shift;
  int32_t min_b[3], max_b[3], mul[3];
  float leaf[3];
  float inv_leaf[3];  // fl32(1 / leaf): fast path of lookup_cell only (the reference's lookup DIVIDES, Q8)
  int32_t min_points;
  unsigned long long n_cells;  // dx*dy*dz (bounds of `dense` / `cell_all`)
  uint32_t n_records;          // occupied voxels (bounds of `records` / `icov64` / `centroids`)
  uint32_t pad;
};

// key -> record index of a VALID voxel, or -1.  Two interchangeable indexes over the same records:
//   * dense: one 4-byte load from a direct-mapped table over the grid's dx*dy*dz cells (x-neighbours share a 32-byte
//     sector) — a single round trip, no compare, no collision chain; used whenever the table fits the memory budget;
//   * hash: open addressing over the valid voxels, for huge sparse grids (up to the int32 key guard, Q9).
__device__ __forceinline__ int map_find(const MapView& m, int32_t key) {
  NDT_CHECK(key >= 0 && static_cast<unsigned long long>(key) < m.n_cells);
  if (m.dense != nullptr) return __ldg(m.dense + key);
  uint32_t h = hash_key(static_cast<uint32_t>(key), m.hash_shift);
  while (true) {
    HashSlot s = __ldg(m.hash + h);
    if (static_cast<uint32_t>(s) == static_cast<uint32_t>(key) && s != NDTB200_HASH_EMPTY)
      return static_cast<int>(s >> 32);
    if (s == NDTB200_HASH_EMPTY) return -1;
    h = (h + 1) & m.hash_mask;
  }
}

// pcl::transformPointCloud, PCL 1.10 order: m0*x + (m1*y + (m2*z + m3)), un-fused fp32.
// T is row-major 3x4.
__device__ __forceinline__ void transform_point(const float* T, float x, float y, float z, float& ox, float& oy,
                                                float& oz) {
  ox = __fadd_rn(__fmul_rn(T[0], x), __fadd_rn(__fmul_rn(T[1], y), __fadd_rn(__fmul_rn(T[2], z), T[3])));
  oy = __fadd_rn(__fmul_rn(T[4], x), __fadd_rn(__fmul_rn(T[5], y), __fadd_rn(__fmul_rn(T[6], z), T[7])));
  oz = __fadd_rn(__fmul_rn(T[8], x), __fadd_rn(__fmul_rn(T[9], y), __fadd_rn(__fmul_rn(T[10], z), T[11])));
}

// getNeighborhoodAtPoint's cell index (voxel_grid_covariance_omp_impl.hpp:379-381): int(floor(x / leaf)), an fp32
// DIVISION (quirk Q8 — the build multiplies by the inverse instead).  A correctly rounded division costs ~35
// instructions; here the product q = x * fl(1/leaf) decides whenever it can:
//     |q - x/leaf| <= 2^-23 |x/leaf| (two roundings) and |fl(x/leaf) - x/leaf| <= 2^-24 |x/leaf|,
// so if q is further than 2^-20 |q| from every integer, fl(x/leaf) lies strictly inside the same unit interval and has
// the same floor.  Otherwise (q within ~1e-6 |q| of a cell face, q == 0, huge or NaN) the exact division decides —
// bit-identical cells, ~0.1 % of the points take the slow path at scan-scale coordinates.
__device__ __forceinline__ int lookup_cell(float x, float leaf, float inv_leaf) {
  const float q = __fmul_rn(x, inv_leaf);
  const float f = floorf(q);
  const float d = fminf(__fsub_rn(q, f), __fsub_rn(__fadd_rn(f, 1.0f), q));  // distance to the nearest integer
  if (d > __fmul_rn(fabsf(q), 9.5367431640625e-07f)) return static_cast<int>(f);
  return static_cast<int>(floorf(__fdiv_rn(x, leaf)));
}

// ---------------------------------------------------------------------------------------------
// block-wide helpers (256-thread blocks)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive scan of one value per thread over the block; `total` receives the block sum.
// smem must hold (blockDim.x / 32) uint32.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  uint32_t inc = warp_inclusive_scan(v);
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < nwarp) ? smem[lane] : 0u;
    uint32_t winc = warp_inclusive_scan(w);
    if (lane < nwarp) smem[lane] = winc - w;
    if (lane == nwarp - 1) smem[nwarp] = winc;
  }
  __syncthreads();
  uint32_t r = inc - v + smem[warp];
  total = smem[nwarp];
  __syncthreads();
  return r;
}

}  // namespace ndtb200
