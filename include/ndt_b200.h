/* ndt_b200.h — C ABI of libndt_b200.so, the B200-native NDT registration hot path.
 *
 * The reference (ToySLAM's vendored ndt_omp) has no FFI: its boundary is the public method set of
 * the C++ class pclomp::NormalDistributionsTransform (ndt_omp/include/pclomp/ndt_omp.h:70-502).
 * Each entry point below names the reference method(s) it replaces.  The header-only C++ shim
 * include/pclomp_b200/ndt_b200.hpp forwards the reference's method names to these calls, so a
 * caller such as ndt_omp/apps/align.cpp or lidar_subscriber/src/ndt_rosbag_mapping_node.cpp
 * only changes its include and namespace.
 *
 * Conventions: plain C types only; every call returns an ndtb200_status (0 = ok); no exceptions
 * cross the ABI; the caller owns all host buffers; the library owns all device memory and one
 * CUDA stream per handle.  A handle is single-threaded (as one reference object is: it keeps
 * mutable per-evaluation tables, ndt_omp.h:473-488); distinct handles are independent.
 * Point buffers are arrays of structs whose first 12 bytes are x,y,z as fp32 and whose size is
 * `stride_bytes` (16 for pcl::PointXYZ, 32 for PointXYZI / PointXYZRGB — the three instantiations
 * in ndt_omp/src/pclomp/ndt_omp.cpp:4-6).  4x4 matrices are column-major fp32 (Eigen::Matrix4f).
 *
 * There is NO CPU fallback: every compute entry point fails with NDTB200_ERR_NO_DEVICE when no
 * CUDA device is usable.
 */
#ifndef NDT_B200_H_
#define NDT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ndtb200_handle ndtb200_handle;

typedef enum ndtb200_status {
  NDTB200_OK = 0,
  NDTB200_ERR_NO_INPUT = 1,      /* no target / source set (voxel_grid_covariance_omp_impl.hpp:54-60) */
  NDTB200_ERR_GRID_OVERFLOW = 2, /* int32 voxel index would overflow (…_impl.hpp:75-84): map left empty */
  NDTB200_ERR_CUDA = 3,          /* a CUDA call failed; see ndtb200_last_error */
  NDTB200_ERR_INVALID = 4,       /* bad argument */
  NDTB200_ERR_NO_DEVICE = 5      /* no usable CUDA device (the library never falls back to the CPU) */
} ndtb200_status;

/* pclomp::NeighborSearchMethod (ndt_omp.h:52-57), same values. */
enum { NDTB200_KDTREE = 0, NDTB200_DIRECT26 = 1, NDTB200_DIRECT7 = 2, NDTB200_DIRECT1 = 3 };

/* Setters of the reference object, as one struct.  Defaults = ndt_omp_impl.hpp:46-76 and
 * voxel_grid_covariance_omp.h:208-211. */
typedef struct ndtb200_params {
  float resolution;         /* setResolution            (1.0)  */
  double step_size;         /* setStepSize              (0.1)  */
  double outlier_ratio;     /* setOutlierRatio          (0.55) */
  double trans_eps;         /* setTransformationEpsilon (0.1)  */
  int32_t max_iterations;   /* setMaximumIterations     (35)   */
  int32_t search_method;    /* setNeighborhoodSearchMethod (DIRECT7) */
  int32_t min_points_per_voxel; /* VoxelGridCovariance::setMinPointPerVoxel (6) */
  double eig_ratio;         /* setCovEigValueInflationRatio (0.01) */
} ndtb200_params;

/* What the reference object exposes after align(). */
typedef struct ndtb200_result {
  float final_transformation[16]; /* getFinalTransformation(): pose of the LAST line-search trial */
  float last_increment[16];       /* pcl::Registration::getLastIncrementalTransformation() */
  int32_t converged;              /* hasConverged() */
  int32_t iterations;             /* getFinalNumIteration() */
  double trans_probability;       /* getTransformationProbability() */
  double final_pose[6];           /* x,y,z,roll,pitch,yaw of the last evaluated trial (fp64) */
  double final_score;             /* score of the last derivative evaluation */
  int32_t n_evaluations;          /* computeDerivatives calls in this align() */
  int32_t n_hessian_passes;       /* computeHessian calls in this align() */
  int64_t n_hits;                 /* (point, voxel) pairs summed over all evaluations */
} ndtb200_result;

typedef struct ndtb200_map_info {
  int32_t min_b[3], max_b[3], div_b[3]; /* VoxelGrid::min_b_/max_b_/div_b_ */
  int64_t n_points;  /* target points handed in */
  int64_t n_voxels;  /* occupied voxels (leaves_.size()) */
  int64_t n_valid;   /* voxels a lookup can return (count >= min_points and spectrum/inverse ok) */
  int64_t hash_capacity;
} ndtb200_map_info;

/* ---- lifetime ------------------------------------------------------------------------------ */
int ndtb200_create(ndtb200_handle** out, int device);        /* NormalDistributionsTransform()        */
int ndtb200_destroy(ndtb200_handle* h);                      /* ~NormalDistributionsTransform()       */
int ndtb200_clone(const ndtb200_handle* src, ndtb200_handle** out); /* copy-construction
                                                                (ndt_omp_mapping_node.cpp:151-169)   */
const char* ndtb200_last_error(const ndtb200_handle* h);
int ndtb200_device_count(void);

/* ---- parameters ---------------------------------------------------------------------------- */
int ndtb200_default_params(ndtb200_params* p);
/* Stores the parameters.  Like setResolution (ndt_omp.h:132-142) a CHANGED resolution rebuilds the
 * target map only if a source is already set; other fields never trigger work. */
int ndtb200_set_params(ndtb200_handle* h, const ndtb200_params* p);
int ndtb200_get_params(const ndtb200_handle* h, ndtb200_params* p);

/* ---- inputs -------------------------------------------------------------------------------- */
/* setInputTarget (ndt_omp.h:122-127): copies the cloud to the device and builds the voxel map now
 * (VoxelGridCovariance::applyFilter, voxel_grid_covariance_omp_impl.hpp:48-370). */
int ndtb200_set_target(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, int is_dense);
/* setInputSource (pcl::Registration). */
int ndtb200_set_source(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes);
/* Same, for clouds already resident in device memory (16-byte float4 records, this device). */
int ndtb200_set_target_device(ndtb200_handle* h, const void* d_points_xyzw, size_t n, int is_dense);
int ndtb200_set_source_device(ndtb200_handle* h, const void* d_points_xyzw, size_t n);
/* setInputTarget WITHOUT a copy: the handle keeps the caller's device pointer (as the reference keeps the caller's cloud
 * by shared pointer).  The buffer must stay alive and unchanged until the next set_target* call on this handle. */
int ndtb200_set_target_device_view(ndtb200_handle* h, const void* d_points_xyzw, size_t n, int is_dense);

/* ---- registration -------------------------------------------------------------------------- */
/* align(output[, guess]) (pcl::Registration::align -> computeTransformation, ndt_omp_impl.hpp:80-171).
 * guess: column-major 4x4 or NULL (= identity).  out_points: NULL or a buffer of n_source records of
 * out_stride_bytes each; x,y,z receive the source transformed by the final pose, the 4th float is 1. */
int ndtb200_align(ndtb200_handle* h, const float* guess, void* out_points, size_t out_stride_bytes);
/* The two halves of ndtb200_align for callers that keep everything on the device: enqueue the whole
 * Newton / More-Thuente solve on the handle's stream without any host synchronisation ... */
int ndtb200_align_async(ndtb200_handle* h, const float* guess);
/* ... and wait for it + fetch the result block (one small D2H copy). */
int ndtb200_sync(ndtb200_handle* h);
int ndtb200_get_result(ndtb200_handle* h, ndtb200_result* out);
/* Throughput mode (not in the reference): the solve kernel of this handle uses small CTAs (one per SM), so the solves
 * of up to four handles — launched on their own streams before anyone waits — are co-resident on every SM and cover
 * each other's barrier / Newton-step latency.  Default off: one solve owns the whole GPU (lowest latency). */
int ndtb200_set_throughput_mode(ndtb200_handle* h, int on);
/* Batched scan pairs (the ndt_rosbag_mapping_node sequence, ndt_rosbag_mapping_node.cpp:99-144, when the pairs are
 * independent): n handles, each with its own target map and source already set, are aligned together in throughput
 * mode.  guesses: n column-major 4x4 matrices or NULL (identity); out_points: n host buffers for the aligned clouds
 * (out_stride_bytes each record) or NULL; results: n entries or NULL.  Returns after all solves have finished; the
 * first non-OK status is reported. */
int ndtb200_align_batch(ndtb200_handle* const* hs, int n, const float* guesses, void* const* out_points,
                        size_t out_stride_bytes, ndtb200_result* results);
/* Batched scan-to-scan odometry (BASELINE configs[2]; the per-scan body of ndt_rosbag_mapping_node.cpp:99-144 for
 * many independent pairs): pair k = (targets[k], sources[k]), host clouds of n_targets[k] / n_sources[k] points with
 * stride_bytes per record.  lanes[l] (one handle) is driven by its own HOST THREAD inside this call and takes the pairs
 * l, l + n_lanes, ...: setInputTarget (upload + voxel-map build), setInputSource (upload), align(guess_k), result —
 * so uploads, builds and solves of different pairs overlap on the device without any caller-side threading.
 * guesses16: n_pairs column-major 4x4 or NULL (identity); results: n_pairs entries.  Returns the first non-OK status. */
int ndtb200_run_pairs(ndtb200_handle* const* lanes, int n_lanes, const void* const* targets, const size_t* n_targets,
                      const void* const* sources, const size_t* n_sources, size_t stride_bytes, const float* guesses16,
                      int n_pairs, ndtb200_result* results);
/* Enqueue-only half (every solve, output cloud and result copy goes onto its handle's stream; no host wait):
 * finish each handle with ndtb200_sync.  out_points: n host buffers (or NULL) as in ndtb200_align. */
int ndtb200_align_batch_async(ndtb200_handle* const* hs, int n, const float* guesses, void* const* out_points,
                              size_t out_stride_bytes);
/* getFitnessScore(max_range) (pcl::Registration): mean squared distance of T*source to its exact
 * nearest raw target point. */
int ndtb200_fitness_score(ndtb200_handle* h, double max_range, double* out);
/* The two sums behind it (sum of the accepted squared distances, their count): with the source sharded over several
 * GPUs every rank calls this on its slice and the ranks all-reduce the pair (SURVEY 8e "getFitnessScore"). */
int ndtb200_fitness_sums(ndtb200_handle* h, double max_range, double* sum_sq_dist, int64_t* n_accepted);
/* calculateScore(cloud) (ndt_omp_impl.hpp:935-983). */
int ndtb200_calculate_score(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, double* out);
/* calculateScore of MANY already transformed clouds against the current map in one call (loop-closure screening:
 * one upload, one launch).  Cloud c = points[cloud_offsets[c] .. cloud_offsets[c + 1]); out[c] = its score. */
int ndtb200_calculate_score_batch(ndtb200_handle* h, const void* points, const size_t* cloud_offsets, int n_clouds,
                                  size_t stride_bytes, double* out);
/* The same screening for ONE cloud under MANY candidate poses: out[k] = calculateScore(transformPointCloud(source,
 * poses[k])) for the handle's current source (nothing is uploaded but the poses; column-major 4x4 fp32 each). */
int ndtb200_score_poses(ndtb200_handle* h, const float* poses16, int n_poses, double* out);

/* ---- multi-GPU source sharding (one process + one handle per GPU of an NVSwitch node; not in the reference) ----
 * Every rank holds the full target map and a contiguous slice of the source; each derivative evaluation exchanges the
 * 29 partial sums by direct stores into every peer's mailbox over NVLink inside the persistent kernel, summed in rank
 * order, so all ranks take the identical Newton step.  Like a collective: every rank must issue the same sequence of
 * align / eval calls.  Usage: export -> exchange the 64-byte handles out of band (e.g. torch.distributed) -> attach. */
#define NDTB200_COMM_HANDLE_BYTES 64
int ndtb200_comm_export(ndtb200_handle* h, void* handle_out64);
int ndtb200_comm_attach(ndtb200_handle* h, int rank, int world, const void* all_handles, int64_t n_source_total);
int ndtb200_comm_detach(ndtb200_handle* h);
/* The same source-sharded solve with `world` (2..8) ranks EMULATED inside one cooperative launch on this handle's GPU:
 * the CTAs of the grid are divided evenly between the ranks, every rank works on the contiguous source range it would
 * own on its own GPU (ranges start on 32-point group boundaries) with its own partial rows, barrier words, result
 * block and mailbox, and the ranks exchange their 29 per-evaluation sums through the same tagged mailbox stores and
 * polls as the multi-GPU path.  results[r] = rank r's own result block (all ranks must report identical bits).
 * For boxes with fewer GPUs than ranks: separate launches that wait on one another are not guaranteed to be
 * co-resident on one device, one cooperative launch is. */
int ndtb200_align_emulated_ranks(ndtb200_handle* h, int world, const float* guess, ndtb200_result* results);

/* ---- scan pre-processing: pcl::VoxelGrid centroid downsample (callers: ndt_omp/apps/align.cpp:57-69,
 * ndt_rosbag_mapping_node.cpp:108-118,153-160, ndt_omp_node.cpp:87-95) -------------------------------------------
 * One fp32 centroid per occupied leaf-sized cell, cells in ascending index order, points of a cell added in input
 * order: the same arithmetic as pcl::VoxelGrid, bit-identical output.  Non-finite points are skipped.
 * out_points: host buffer of out_capacity records (NULL = count only); n_out receives the number of centroids.
 * NDTB200_ERR_GRID_OVERFLOW = "Leaf size is too small for the input dataset" (PCL then passes the cloud through). */
int ndtb200_voxelgrid_filter(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, float leaf,
                             void* out_points, size_t out_capacity, size_t out_stride_bytes, int64_t* n_out);
/* pcl::VoxelGrid::setLeafSize(lx, ly, lz) with a non-cubic leaf. */
int ndtb200_voxelgrid_filter3(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, const float leaf[3],
                              void* out_points, size_t out_capacity, size_t out_stride_bytes, int64_t* n_out);
/* Same with device-resident float4 input / output (this device). */
int ndtb200_voxelgrid_filter_device(ndtb200_handle* h, const void* d_points_xyzw, size_t n, float leaf, void* d_out_xyzw,
                                    size_t out_capacity, int64_t* n_out);

/* ---- the mapping-node loop as one device-resident pipeline (lidar_subscriber/src/ndt_rosbag_mapping_node.cpp:42-161) --
 * push_scan(k) = downsample_cloud (VoxelGrid voxel_leaf) -> perform_registration(previous filtered scan, this one,
 * guess = previous transform; Identity if not converged) -> pose = pose * transform -> update_global_map (transform by
 * the pose, append, VoxelGrid map_voxel).  The first scan only initialises the map.  Clouds stay in HBM between
 * steps; the map of scan k is built while scan k is still being aligned against the map of scan k-1.
 * ndt_params NULL = the node's defaults (eps 0.01, 64 iterations, step 0.1, resolution 1.0, DIRECT7). */
typedef struct ndtb200_mapper ndtb200_mapper;
typedef struct ndtb200_mapper_step {
  float transform[16];   /* column-major: this step's registration result (perform_registration's return value) */
  float pose[16];        /* column-major: accumulated pose after this step */
  double fitness;        /* getFitnessScore() of the registration (0 for the first scan or when disabled) */
  int32_t converged, iterations, n_evaluations, pad;
  int64_t n_filtered;    /* points of the downsampled scan */
  int64_t n_map;         /* points of the global map after this step */
} ndtb200_mapper_step;
int ndtb200_mapper_create(ndtb200_mapper** out, int device, const ndtb200_params* ndt_params, float voxel_leaf, float map_voxel,
                          int compute_fitness);
int ndtb200_mapper_destroy(ndtb200_mapper* m);
int ndtb200_mapper_push_scan(ndtb200_mapper* m, const void* points, size_t n, size_t stride_bytes, ndtb200_mapper_step* out);
int ndtb200_mapper_get_map(ndtb200_mapper* m, void* out_points, size_t capacity, size_t out_stride_bytes, int64_t* n_out);
const char* ndtb200_mapper_last_error(const ndtb200_mapper* m);
int64_t ndtb200_mapper_launch_count(const ndtb200_mapper* m);

/* ---- multi-GPU target-map build (not in the reference; SURVEY 8e "target-map build") ----------------------------
 * The cloud is split by contiguous point ranges, one per rank.  (1) ndtb200_cloud_bounds: bounding box + finite count
 * of this rank's slice (device pointer, float4 records); the caller all-reduces min / max / count over the ranks.
 * (2) ndtb200_build_partials: keys with the COMMON grid, sort, per-voxel {key, count, sum x, sum x x^T} of the slice;
 * (3) the caller all-gathers the partial arrays (ndtb200_copy_partials copies them into caller-owned device buffers:
 * int32 keys, uint32 counts, 9 fp64 moments per voxel) and concatenates them in rank order;
 * (4) ndtb200_build_from_partials merges them (stable sort by key, sums in rank order: identical bits on every
 * rank), finalises every voxel and builds the index: every rank ends with the full map.  Keys and counts are exact;
 * moments differ from a single-GPU build only by the fp64 summation order.  getFitnessScore is not available on a
 * handle whose map was merged (it holds only its slice of the raw target). */
int ndtb200_cloud_bounds(ndtb200_handle* h, const void* d_points_xyzw, size_t n, int is_dense, float out_min[3],
                         float out_max[3], int64_t* n_finite);
int ndtb200_build_partials(ndtb200_handle* h, const float global_min[3], const float global_max[3], int64_t* n_partials);
int ndtb200_copy_partials(ndtb200_handle* h, void* d_keys, void* d_counts, void* d_moments);
int ndtb200_build_from_partials(ndtb200_handle* h, const float global_min[3], const float global_max[3],
                                int64_t n_finite_total, const void* d_keys, const void* d_counts, const void* d_moments,
                                size_t n_total);

/* Owner-partitioned variant of steps (3)-(4) (SURVEY 8e row 3 as designed): instead of gathering every partial on every
 * rank, (3a) ndtb200_partials_split tells where this rank's (key-sorted) partials cross the owners' key boundaries
 * (upper_keys[r] = first key NOT owned by rank r, world-1 ascending values; offsets_out[world+1]), (3b) the caller
 * moves each range to its owner (all-to-all), (4a) ndtb200_merge_partials merges + finalises the received partials
 * (rank order) into finished records without building an index, (4b) ndtb200_copy_records exports them (64-byte
 * records + 6 fp64 inverse-covariance entries per voxel), the caller all-gathers them in rank (= key) order, and (4c)
 * ndtb200_set_map_from_records installs the full map and builds the voxel index on every rank. */
int ndtb200_partials_split(ndtb200_handle* h, const int32_t* upper_keys, int world, int64_t* offsets_out);
int ndtb200_merge_partials(ndtb200_handle* h, const float global_min[3], const float global_max[3], int64_t n_finite_total,
                           const void* d_keys, const void* d_counts, const void* d_moments, size_t n_total, int64_t* n_merged);
int ndtb200_copy_records(ndtb200_handle* h, void* d_records_out, void* d_icov64_out);
int ndtb200_set_map_from_records(ndtb200_handle* h, const float global_min[3], const float global_max[3],
                                 int64_t n_finite_total, const void* d_records, const void* d_icov64, size_t n_total);

/* ---- parity / inspection (stage dumps; used by tests, not by callers) ------------------------ */
int ndtb200_get_map_info(const ndtb200_handle* h, ndtb200_map_info* out);
/* voxel key of every target point, input order (-1 = skipped non-finite point). */
int ndtb200_dump_point_keys(ndtb200_handle* h, int32_t* keys);
/* all occupied voxels in ascending key order: counts (-1 = rejected leaf), mean[3], cov[9], icov[9]
 * (row-major fp64), inflated flag.  Any pointer may be NULL. */
int ndtb200_dump_voxels(ndtb200_handle* h, int32_t* keys, int32_t* counts, double* mean, double* cov,
                        double* icov, int32_t* inflated);
/* One computeDerivatives(p) (ndt_omp_impl.hpp:179-285) at pose p.  T: matrix to transform the source
 * with, or NULL (= built from p like computeStepLengthMT does).  out43 = score, gradient[6],
 * hessian[36] row-major; n_hits may be NULL. */
int ndtb200_eval_derivatives(ndtb200_handle* h, const double p[6], const float* T, int compute_hessian,
                             double out43[43], int64_t* n_hits);
/* One computeHessian (ndt_omp_impl.hpp:540-645) at pose p (fp64 path, fp64 angle tables). */
int ndtb200_eval_hessian(ndtb200_handle* h, const double p[6], const float* T, double out36[36]);
/* getNeighborhoodAtPoint{,7,1} (voxel_grid_covariance_omp_impl.hpp:373-442) for n query points:
 * out_keys[n][26], -1 padded, in the reference's offset order. */
int ndtb200_lookup(ndtb200_handle* h, const void* points, size_t n, size_t stride_bytes, int search_method,
                   int32_t* out_keys);

/* static convertTransform (ndt_omp.h:216-233): x = [x, y, z, roll, pitch, yaw] -> Translation * AngleAxis(roll, X) *
 * AngleAxis(pitch, Y) * AngleAxis(yaw, Z) as a column-major fp32 4x4.  Pure host arithmetic (no device, no handle). */
int ndtb200_pose_to_matrix(const double x[6], float out16[16]);

/* The pose vector computeTransformation starts from (ndt_omp_impl.hpp:103-111): translation and
 * rotation().eulerAngles(0,1,2) (Eigen 3.3 conventions, polar factor of the linear part) of a column-major 4x4 guess.
 * Pure host arithmetic: needs no device and no handle. */
int ndtb200_debug_guess_to_pose(const float guess[16], double p_out[6]);
/* The on-device Newton solve H * delta = -g (JacobiSVD(H).solve(-g), ndt_omp_impl.hpp:127-129) exactly as the persistent
 * kernel runs it (one warp; definite elimination / pivoted elimination / SVD pseudo-inverse).  H row-major 6x6;
 * path_out (may be NULL): 0 = definite fast path, 1 = pivoted / pseudo-inverse path. */
int ndtb200_debug_newton_solve(ndtb200_handle* h, const double H[36], const double g[6], double delta_out[6], int* path_out);

/* The on-device Newton / More-Thuente trace of the last align: one entry per evaluation
 * (kind 0 = derivatives+Hessian, 1 = derivatives only, 2 = Hessian only; pose; step length; score).
 * n_out receives the number of evaluations; at most `cap` entries are written. */
int ndtb200_get_trace(ndtb200_handle* h, int32_t* kinds, double* x6, double* a_t, double* score, int cap, int* n_out);

/* Profiling: CTA-0 timeline of the last solve, 4 stamps per evaluation in ns since the first evaluation
 * started: evaluation start, local work done, reduced totals available, next step decided. */
int ndtb200_get_timeline(ndtb200_handle* h, double* t4, int cap, int* n_out);

/* ---- plumbing for benchmarks ----------------------------------------------------------------- */
/* The handle's cudaStream_t (as void*), so a caller can record CUDA events on it. */
void* ndtb200_stream(ndtb200_handle* h);
/* Kernels launched by this handle since creation (or since the last reset). */
int64_t ndtb200_launch_count(const ndtb200_handle* h);
void ndtb200_reset_launch_count(ndtb200_handle* h);
/* Milliseconds the last ndtb200_align_async solve kernel took (CUDA events on the handle stream). */
int ndtb200_last_align_ms(ndtb200_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* NDT_B200_H_ */
