/*
 * plslam_b200 -- C ABI of the B200-native ICP-odometry hot path of pyLiDAR-SLAM.
 *
 * Drop-in boundary: these are the entry points the Python
 * plug-in classes of pylidar_slam_b200 (mirrors of the reference's OdometryAlgorithm /
 * LocalMap / RigidAlignment / Filter interfaces) bind through ctypes.  Every function
 * cites the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - Every call returns an int status (PLS_OK == 0).  pls_last_error(ctx) gives text.
 *  - Data pointers may be HOST or DEVICE pointers; the library classifies each pointer
 *    with cudaPointerGetAttributes and stages host memory itself (pinned host memory is
 *    copied asynchronously).  Results are written back to wherever the out-pointer lives.
 *    A call returns after its results are visible to the caller.
 *  - Point clouds are row-major [N,3]; projection maps are planar [C,H,W] (the reference's
 *    layouts); poses are row-major 4x4; pose parameters are (tx,ty,tz,ex,ey,ez), Euler xyz,
 *    R = Rz(ez) Ry(ey) Rx(ex) (slam/common/rotation.py:144-150).
 *  - float means IEEE binary32, the hot path's arithmetic type (icp_odometry.py:353-354).
 *  - Not re-entrant per context: one host thread drives a context (as one Python thread
 *    drives the reference algorithm).
 */
#ifndef PLSLAM_B200_H
#define PLSLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PLS_API __attribute__((visibility("default")))
#else
#define PLS_API
#endif

typedef struct pls_context pls_context;

/* Status codes.  PLS_E_SINGULAR <-> RuntimeError("Invalid Jacobian in Gauss Newton
 * minimization") (slam/common/optimization.py:334-336); PLS_W_TINY_RESIDUAL <-> the
 * logging.warning + early return at optimization.py:323-327. */
enum {
    PLS_OK = 0,
    PLS_E_INVALID = 1,       /* bad argument / shape  (AssertionError in the reference) */
    PLS_E_CUDA = 2,          /* CUDA runtime failure */
    PLS_E_SINGULAR = 3,
    PLS_W_TINY_RESIDUAL = 4,
    PLS_E_STATE = 5,         /* call order (e.g. search before any map update) */
    PLS_E_COMM = 6           /* NCCL failure */
};

/* Robust weighting schemes, slam/common/optimization.py:211-226 (_LS_SCHEME). */
enum {
    PLS_SCHEME_DEFAULT = 0,
    PLS_SCHEME_LEAST_SQUARE = 1,
    PLS_SCHEME_HUBER = 2,
    PLS_SCHEME_EXP = 3,
    PLS_SCHEME_NEIGHBORHOOD = 4,
    PLS_SCHEME_GEMAN_MCCLURE = 5,
    PLS_SCHEME_SQUARE_GEMAN_MCCLURE = 6,
    PLS_SCHEME_CAUCHY = 7
};

/* LOCAL_MAP registry, slam/odometry/local_map.py:437-445. */
enum { PLS_MAP_KDTREE = 0, PLS_MAP_PROJECTIVE = 1 };

/* The three input layouts of ICPFrameToModel._read_input, icp_odometry.py:319-358. */
enum {
    PLS_INPUT_NDARRAY = 0,    /* np.ndarray [N,3]: queries = the raw points              */
    PLS_INPUT_TENSOR = 1,     /* torch [N,3]: queries = non-null pixels of its vertex map */
    PLS_INPUT_VERTEX_MAP = 2, /* torch [1,3,H,W]: used as the vertex map directly         */
    /* float64 clouds (e.g. the output of the Distortion filter): the reference projects them in float64 and
     * rounds the vertex map / the points to float32 afterwards (icp_odometry.py:331-352); `data` is double [n,3] */
    PLS_INPUT_NDARRAY_F64 = 3,
    PLS_INPUT_TENSOR_F64 = 4
};

/* Optional residency hints, OR-ed into the `layout` argument of pls_process_frame by a caller that knows where `data`
 * lives (no reference counterpart: the reference's tensors carry their device).  Without a hint the pointer is classified. */
enum { PLS_PTR_DEVICE = 0x100, PLS_PTR_HOST = 0x200 };

/* Configuration = SphericalProjector (projection.py:439-450) + ICPFrameToModelConfig
 * (icp_odometry.py:27-64) + local-map configs (local_map.py:83-88,244-251) +
 * GaussNewtonPointToPlaneConfig.gauss_newton_config (alignment.py:69-77). */
typedef struct pls_config {
    int32_t height, width;          /* projector image size */
    float up_fov_deg, down_fov_deg; /* projector vertical field of view */
    int32_t local_map_type;         /* PLS_MAP_* */
    int32_t local_map_size;         /* frames kept (20) */
    int32_t num_neighbors_normals;  /* kd map: k for normals (10) */
    int32_t normals_kernel_size;    /* projective map: box size (5) */
    int32_t scheme;                 /* PLS_SCHEME_* */
    float sigma;                    /* scheme parameter */
    int32_t gn_max_iters;           /* Gauss-Newton iterations per alignment (1) */
    float gn_norm_stop;             /* GN stop on |dx| (1e-3) */
    int32_t max_num_alignments;     /* ICP iterations per frame */
    float threshold_delta_pose;     /* ICP stop on |delta| (1e-4) */
    float threshold_trans;          /* key-frame policy, metres (0.1) */
    float threshold_rot;            /* key-frame policy, degrees (0.3) */
    int32_t device;                 /* CUDA device ordinal */
    void* stream;                   /* cudaStream_t to run on, or NULL for a private stream */
} pls_config;

/* ---- lifetime ------------------------------------------------------------------- */
PLS_API int pls_config_default(pls_config* cfg);
PLS_API int pls_create(const pls_config* cfg, pls_context** out);
PLS_API int pls_destroy(pls_context* ctx);
PLS_API const char* pls_last_error(pls_context* ctx);
PLS_API const char* pls_version(void);
/* cudaStreamSynchronize on the context's stream. */
PLS_API int pls_synchronize(pls_context* ctx);
/* Orders the work already enqueued on `other_stream` (a cudaStream_t, e.g. PyTorch's current stream that produced a
 * CUDA tensor about to be passed in) before everything this context enqueues afterwards; no host synchronisation. */
PLS_API int pls_wait_stream(pls_context* ctx, void* other_stream);

/* ---- a1: voxel-grid subsample ----------------------------------------------------
 * voxelise + voxel_hashing  (slam/common/pointcloud.py:13-23,40-79): int64 voxel
 * coordinates round_half_even(p / voxel) computed in float64 and the signed 64-bit hash
 * 73856093 x + 19349669 y + 83492791 z.  is_f64 selects float64 input points. */
PLS_API int pls_voxel_hash(pls_context* ctx, const void* xyz, int is_f64, int64_t n, double voxel,
                   int64_t* coords_out /* [n,3] or NULL */, int64_t* hashes_out /* [n] */);
/* voxelise with one voxel length per axis (pointcloud.py:55-79: voxel_x, voxel_y, voxel_z). */
PLS_API int pls_voxel_hash_xyz(pls_context* ctx, const void* xyz, int is_f64, int64_t n, double voxel_x, double voxel_y,
                       double voxel_z, int64_t* coords_out /* [n,3] or NULL */, int64_t* hashes_out /* [n] or NULL */);
/* grid_sample / GridSample.filter (pointcloud.py:170-195, preprocessing.py:213-226):
 * one point per distinct hash (its first occurrence), ordered by ascending hash.
 * out_xyz [n,3] (same dtype as the input), out_idx [n] int64; *out_count = S. */
PLS_API int pls_grid_sample(pls_context* ctx, const void* xyz, int is_f64, int64_t n, double voxel,
                    void* out_xyz, int64_t* out_idx, int64_t* out_count);

/* The same subsample for HOST callers without an extra copy (GridSample.filter hands numpy arrays to the next
 * filter): the gather kernel writes samples and indices straight into the context's pinned, device-mapped staging and
 * keeps a device-resident copy.  *out_xyz_host / *out_idx_host point into the staging (S rows, valid until the next
 * grid-sample call on this context); *out_xyz_dev (optional) is the device copy, which pls_process_frame accepts with
 * PLS_PTR_DEVICE -- the samples then never travel host -> device again.  One stream synchronisation per call.
 * Caller-owned staging: when *out_xyz_host and *out_idx_host are non-NULL ON ENTRY they name pinned, device-mapped
 * buffers from pls_pinned_alloc with room for n rows each, and the kernel writes there instead -- the caller (the
 * GridSample filter) then hands those very buffers on as arrays, no host-side copy at all. */
PLS_API int pls_grid_sample_staged(pls_context* ctx, const void* xyz, int is_f64, int64_t n, double voxel,
                           const void** out_xyz_host, const int64_t** out_idx_host, const void** out_xyz_dev,
                           int64_t* out_count);
/* Page-locked, device-mapped host memory for results a kernel writes straight into (see pls_grid_sample_staged). */
PLS_API int pls_pinned_alloc(int64_t num_bytes, void** out_ptr);
PLS_API int pls_pinned_free(void* ptr);
/* 64-bit fingerprint of a host buffer (256 evenly spread 8-byte words + its length): lets a caller check cheaply that
 * an array it handed out earlier still holds what the device-resident copy holds. */
PLS_API int pls_host_fingerprint(const void* host_ptr, int64_t num_bytes, uint64_t* out);

/* ---- a2/a3: spherical projection + closest-wins z-buffer --------------------------
 * SphericalProjector.project_pointcloud (projection.py:11-73,452-484): float pixel
 * coordinates, rows_cols_out [n,2] = (row, col). */
PLS_API int pls_project_pixels(pls_context* ctx, const float* xyz, int64_t n, int height, int width,
                       float up_fov_deg, float down_fov_deg, float* rows_cols_out);
/* Projector.build_projection_map (projection.py:331-418) for B stacked clouds:
 * xyz [B,n,3]; channels [B,n,C] or NULL (then C = 3 and the xyz are scattered);
 * out [B,C,H,W].  Per pixel the closest point (lowest index on exact ties) survives. */
PLS_API int pls_build_projection_map(pls_context* ctx, const float* xyz, const float* channels,
                             int batch, int64_t n, int num_channels, int height, int width,
                             float up_fov_deg, float down_fov_deg, float* out);
/* The same with the reference's `default_value` (projection.py:333,378-391): pixels no point lands on hold it. */
PLS_API int pls_build_projection_map_filled(pls_context* ctx, const float* xyz, const float* channels,
                                    int batch, int64_t n, int num_channels, int height, int width,
                                    float up_fov_deg, float down_fov_deg, float default_value, float* out);

/* ---- a4: box-filter normal map  (slam/common/geometry.py:240-295) ----------------- */
PLS_API int pls_normal_map(pls_context* ctx, const float* vertex_map /* [B,3,H,W] */, int batch,
                   int height, int width, int kernel_size, float* out /* [B,3,H,W] */);

/* ---- a6: projective association  (slam/common/geometry.py:397-439) ----------------
 * tgt [3,H,W]; ref [K,3,H,W]; fields [K,C,H,W] or NULL; out_nb [3,H,W]; out_fields [C,H,W]. */
PLS_API int pls_compute_neighbors(pls_context* ctx, const float* tgt, const float* ref, const float* fields,
                          int num_ref, int num_field_channels, int height, int width,
                          float* out_nb, float* out_fields);

/* ---- a16: pose algebra  (slam/common/pose.py:120-144,188-207) --------------------- */
PLS_API int pls_build_pose_matrix(pls_context* ctx, const float* params /* [B,6] */, int batch, float* out /* [B,4,4] */);
PLS_API int pls_from_pose_matrix(pls_context* ctx, const float* mats /* [B,4,4] */, int batch, float* out /* [B,6] */);

/* ---- a11-a15: point-to-plane Gauss-Newton alignment --------------------------------
 * GaussNewtonPointToPlaneAlignment.align (alignment.py:91-127) -> GaussNewton.compute
 * (optimization.py:296-344) with PointToPlaneCost closures (optimization.py:356-435).
 * ref/tgt/nrm are [n,3] (batch 1); is_f64 selects float64 data and arithmetic.
 * x0 [6] or NULL (zeros).  Outputs (same dtype as the inputs): out_dT [16], out_x [6],
 * out_loss [n] = (w r)^2 or NULL.  Returns PLS_E_SINGULAR / PLS_W_TINY_RESIDUAL like the
 * reference raises / warns. */
PLS_API int pls_align_p2plane(pls_context* ctx, const void* ref, const void* tgt, const void* nrm, int64_t n,
                      int is_f64, int scheme, double sigma, int max_iters, double norm_stop,
                      const void* x0, void* out_dT, void* out_x, void* out_loss);

/* ---- a5/a7/a8/a9: local maps (LocalMap ABC, slam/odometry/local_map.py:31-79) -------
 * One local map lives in a context; its type is cfg.local_map_type. */
PLS_API int pls_map_init(pls_context* ctx); /* LocalMap.init */
/* KdTreeLocalMap.update (local_map.py:302-362): rel_pose [16]; new points [n,3] or NULL.
 * The map is moved by inverse(rel_pose), the new frame appended, the oldest frame dropped
 * beyond local_map_size, the search index rebuilt and the normal cache cleared. */
PLS_API int pls_kdmap_update_points(pls_context* ctx, const float* rel_pose, const float* points, int64_t n);
/* Same, inserting the pixels of a vertex map with |p| > 0.01 (local_map.py:320-324). */
PLS_API int pls_kdmap_update_vertex_map(pls_context* ctx, const float* rel_pose, const float* vertex_map,
                                int height, int width);
PLS_API int pls_kdmap_size(pls_context* ctx, int64_t* num_points);
/* Debug counters of the kd search (enabled by the environment variable PLS_KD_STATS=1), 16 u64. */
PLS_API int pls_kdmap_stats(pls_context* ctx, unsigned long long* out16);
PLS_API int pls_kdmap_points(pls_context* ctx, float* out /* [M,3], insertion order */);
/* KdTreeLocalMap.nearest_neighbor_search (local_map.py:372-422): exact 1-NN; normals from
 * the 10 nearest map neighbours of the matched map point (smallest-eigenvalue direction),
 * cached per map point until the next update.  out_idx [n] (int64, insertion order) or NULL. */
PLS_API int pls_kdmap_nn_search(pls_context* ctx, const float* queries, int64_t n,
                        float* out_neighbors, float* out_normals, int64_t* out_idx);
/* ProjectiveLocalMap.update (local_map.py:126-202): rel_pose [16]; vertex_map [3,H,W] or NULL. */
PLS_API int pls_projmap_update(pls_context* ctx, const float* rel_pose, const float* vertex_map);
PLS_API int pls_projmap_num_frames(pls_context* ctx, int* num_frames);
/* The re-projected model maps _model_vmap/_model_nmap, each [K,3,H,W]. */
PLS_API int pls_projmap_model(pls_context* ctx, float* out_vmap, float* out_nmap);
/* ProjectiveLocalMap.nearest_neighbor_search (local_map.py:205-235): queries [n,3];
 * outputs [Nc,3] each (capacity H*W rows), row-major pixel order; *out_count = Nc. */
PLS_API int pls_projmap_nn_search(pls_context* ctx, const float* queries, int64_t n,
                          float* out_neighbors, float* out_normals, float* out_targets,
                          int64_t* out_count);

/* ---- a17/a18: the odometry ----------------------------------------------------------
 * ICPFrameToModel.init (icp_odometry.py:128-145). */
PLS_API int pls_odometry_init(pls_context* ctx);
/* ICPFrameToModel.register_new_frame (icp_odometry.py:248-299): points [n,3], T0 [16].
 * out_T [16], out_params [6], out_losses [max_num_alignments] (unused entries NaN),
 * *out_iters = iterations executed (= len(losses)). */
PLS_API int pls_register_frame(pls_context* ctx, const float* points, int64_t n, const float* T0,
                       float* out_T, float* out_params, float* out_losses, int* out_iters);
/* ICPFrameToModel.do_process_next_frame (icp_odometry.py:157-246): `data` is [n,3] points
 * (PLS_INPUT_NDARRAY / PLS_INPUT_TENSOR: float; the _F64 variants: double) or a float [3,H,W] vertex map
 * (PLS_INPUT_VERTEX_MAP, n ignored).  init_pose [16] or NULL (identity).  On frame 0 the map is initialised and
 * *out_has_pose = 0 (the reference writes no "odometry_pose" then).  out_info (optional,
 * 12 doubles): iterations, final loss, queries used, map points, grid samples, NaN rows dropped,
 * status, key-frame inserted, first non-null pixel x/y/z (vertex-map layout), 1 if the correspondences were
 * sharded over the ranks of a multi-GPU communicator (0: every rank ran the whole frame). */
PLS_API int pls_process_frame(pls_context* ctx, const void* data, int layout, int64_t n,
                      const float* init_pose, float* out_pose, float* out_params,
                      int* out_has_pose, double* out_info);
/* Fused preprocessing + odometry for the shipped pipeline (grid_sample.yaml): GridSample
 * (voxel) -> ToTensor -> process_frame(PLS_INPUT_TENSOR or _NDARRAY) with no host hop. */
PLS_API int pls_process_frame_grid_sample(pls_context* ctx, const float* raw_points, int64_t n, double voxel,
                                  int layout, const float* init_pose, float* out_pose,
                                  float* out_params, int* out_has_pose, double* out_info);

/* ---- the rows either side of the path -----------------------------------
 * Distortion.filter (slam/preprocessing.py:148-191): de-skew of a frame with the estimated relative motion.
 * xyz [n,3] float32|float64; timestamps [n] float32|float64 (alpha is formed in their dtype, like numpy does);
 * rel_pose [16] float32|float64 (the data_dict's init_rpose); out [n,3] FLOAT64 = Slerp(I -> R)(alpha_i) p_i +
 * alpha_i t.  Constant timestamps give alpha = 0; a NaN timestamp makes every output NaN (np.max/np.min). */
PLS_API int pls_distort(pls_context* ctx, const void* xyz, int xyz_is_f64, const void* timestamps, int ts_is_f64,
                int64_t n, const void* rel_pose, int pose_is_f64, double* out);
/* Voxelization.filter (slam/preprocessing.py:71-97) = voxelise + voxel_hashing + voxel_normal_distribution
 * (slam/common/pointcloud.py:54-79,40-51,83-167).  coords_out [n,3] / hashes_out [n] as pls_voxel_hash (nullable);
 * per distinct hash, in ascending hash order: sizes_out [V] int64, means_out [V,3], covs_out [V,3,3] = the scatter
 * matrix sum (x - mean)(x - mean)^T (not divided by the count), both in the dtype of the input; ids_out [n] int64 =
 * voxel rank of every point.  Per-voxel outputs need capacity n rows; *out_count = V. */
PLS_API int pls_voxel_statistics(pls_context* ctx, const void* xyz, int is_f64, int64_t n, double voxel,
                         int64_t* coords_out, int64_t* hashes_out, int64_t* sizes_out, void* means_out,
                         void* covs_out, int64_t* ids_out, int64_t* out_count);
/* GaussNewtonPointToPointAlignment.align (slam/odometry/alignment.py:144-189) -> GaussNewton.compute
 * (optimization.py:296-344) with PointToPointCost closures (optimization.py:458-541): r = |R p + t - q| and the
 * reference's Jacobian as written, J[k] = (dT/dx_k p~) . (R p + t - q).  Arguments as pls_align_p2plane. */
PLS_API int pls_align_p2point(pls_context* ctx, const void* ref, const void* tgt, int64_t n, int is_f64, int scheme,
                      double sigma, int max_iters, double norm_stop, const void* x0, void* out_dT, void* out_x,
                      void* out_loss);
/* weighted_procrustes, numpy path (slam/common/registration.py:15-76): rigid T (float64 [16]) with
 * T * tgt ~ ref.  tgt / ref [n,3] and weights [n] (nullable) share one dtype; the weights only enter the centroids. */
PLS_API int pls_weighted_procrustes(pls_context* ctx, const void* tgt, const void* ref, const void* weights, int64_t n,
                            int is_f64, double* out_T);

/* _PointToPlaneLossModule.point_to_plane_loss (slam/training/loss_modules.py:51-104), forward AND backward: the
 * unsupervised point-to-plane training loss of a batch of (target, reference) vertex-map pairs and its gradient with
 * respect to the predicted poses, as the reference's autograd computes it (values flow through the z-buffer scatter to
 * every point written to a pixel; rounded pixel coordinates carry no gradient).  vm_target / vm_reference /
 * nm_reference [B,3,H,W]; pose_mats [B,16] or NULL (then built from pose_params [B,6], Euler xyz); up/down fov of the
 * projector; scheme / sigma = least_square_scheme.  out_loss [1] = mean_b(sum C(|r|)^2 / sum mask); optional
 * out_loss_per_batch [B], out_grad_mats [B,16] (d loss / d pose matrix, last row 0), out_grad_params [B,6]. */
PLS_API int pls_p2plane_loss(pls_context* ctx, const float* vm_target, const float* vm_reference,
                     const float* nm_reference, const float* pose_mats, const float* pose_params, int batch,
                     int height, int width, float up_fov_deg, float down_fov_deg, int scheme, float sigma,
                     float* out_loss, float* out_loss_per_batch, float* out_grad_mats, float* out_grad_params);

/* ---- the rows either side of the path, rank 4: dataset -> vertex-map ingestion and pose chains --------------------------
 * KITTIOdometrySequence.correct_scan (slam/dataset/kitti_dataset.py:200-231): the HDL-64 intrinsic correction, every point
 * rotated by 0.205 degrees about normalise(p x e_z).  scan [n, stride] float32 rows (x, y, z[, reflectance]), stride 3 or 4;
 * out_xyz [n,3] FLOAT64 (the reference's result dtype).  Points on the vertical axis come out NaN, as in the reference. */
PLS_API int pls_kitti_correct_scan(pls_context* ctx, const float* scan, int64_t n, int stride, double* out_xyz);
/* KITTIOdometrySequence.__getitem__ (kitti_dataset.py:233-249): optional rectification, then the spherical projection +
 * closest-wins z-buffer in float64 (Projector.build_projection_map on the float64 cloud).  out_xyz [n,3] float64 (the
 * `numpy_pc` entry; may be NULL), out_vertex_map [3,H,W] float64 (the `vertex_map` entry). */
PLS_API int pls_ingest_scan(pls_context* ctx, const float* scan, int64_t n, int stride, int correct, int height, int width,
                    float up_fov_deg, float down_fov_deg, double* out_xyz, double* out_vertex_map);
/* compute_relative_poses (slam/eval/eval_odometry.py:80-83): out[i] = inv(poses[i-1]) @ poses[i], out[0] = poses[0];
 * poses / out [n,4,4] float32 or float64 (is_f64), computed in that precision. */
PLS_API int pls_relative_poses(pls_context* ctx, const void* poses, int64_t n, int is_f64, void* out);
/* compute_absolute_poses (eval_odometry.py:86-96): out[0] = rel[0], out[i+1] = out[i] @ rel[i+1] (sequential product). */
PLS_API int pls_absolute_poses(pls_context* ctx, const void* relative_poses, int64_t n, int is_f64, void* out);

/* ---- multi-GPU: per-iteration allreduce of the normal-equation accumulators ---------
 * (no reference counterpart: SURVEY.md section 8e).  Every rank holds the whole local map
 * (kd) or its band of image rows (projective) and a shard of the queries; after
 * pls_comm_init the 30 accumulators are summed across ranks once per ICP iteration.
 * nccl_unique_id: the 128-byte ncclUniqueId created on rank 0 and broadcast by the host
 * program (torch.distributed); nccl_library: path of libnccl.so.2 to dlopen. */
PLS_API int pls_comm_init(pls_context* ctx, int num_ranks, int rank, const void* nccl_unique_id,
                  const char* nccl_library);
PLS_API int pls_comm_unique_id(const char* nccl_library, void* out_id_128_bytes);
/* One-shot peer-to-peer mode (NVLink / NVSwitch peer memory, no NCCL on the data path): the all-reduce is
 * fused into the solve kernel -- every rank stores its 30 partial sums and a sequence flag directly into each
 * peer's exchange slots and sums the slots in rank order.  Step 1: each rank exports the 64-byte CUDA IPC
 * handle of its exchange buffer; the host program all-gathers the handles; step 2 maps them. */
PLS_API int pls_comm_p2p_handle(pls_context* ctx, int num_ranks, void* out_handle_64_bytes);
PLS_API int pls_comm_p2p_init(pls_context* ctx, int num_ranks, int rank, const void* all_handles /* [num_ranks][64] */);
PLS_API int pls_comm_destroy(pls_context* ctx);
/* Sharding threshold: a frame's correspondences are split over the ranks only if every rank gets at least this many work
 * items (queries of the kd map / pixels of the projective map); below it every rank runs the whole frame itself and no
 * exchange takes place (identical inputs + deterministic kernels = identical poses).  Default 24576, a third of it for
 * kd maps of 2 M points and more, where a query costs several times as much (environment
 * variable PLS_SHARD_MIN); a negative value restores the default.  Process-wide; every rank must set the same value. */
PLS_API int pls_set_shard_min(int64_t work_items_per_rank);
/* 1 if the last ICP of this context split its correspondences over the ranks, else 0. */
PLS_API int pls_last_sharded(pls_context* ctx, int* out);

/* ---- measurement ---------------------------------------------------------------------
 * CUDA-event timing of one kernel family inside the library's own launches.
 * which: 0 = kd correspondence+reduction kernel, 1 = projective correspondence+reduction
 * kernel, 2 = model rebuild, 3 = index build, 4 = grid sample, 5 = Gauss-Newton solve.
 * pls_profile_read returns the accumulated device milliseconds, launch count and the
 * algorithmic bytes the launches moved (DESIGN.md states the per-unit figures). */
/* Number of CUDA kernels this library has launched in this process (all contexts). */
PLS_API int pls_launch_count(int64_t* out);
PLS_API int pls_profile_enable(pls_context* ctx, int which, int enable);
PLS_API int pls_profile_read(pls_context* ctx, int which, double* ms_total, int64_t* launches,
                     double* algorithmic_bytes, int reset);

#ifdef __cplusplus
}
#endif
#endif /* PLSLAM_B200_H */
