// Internal declarations shared by the translation units of libplslam_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <deque>
#include <string>
#include <vector>

#include "../../include/plslam_b200.h"

namespace pls {

constexpr int kNumSMs = 148;  // B200
constexpr int NACC = 30;      // 21 JtJ upper + 6 Jtr + sum (w r)^2 + sum r^2 + count
constexpr int kMaxAlign = 128;
constexpr int kProfileSlots = 16;  // 0-5: kernel families of the bench line; 6-15: single kernels (development)
constexpr size_t kScalarOffset = 2048;  // FrameResult first, then u32 scalar slots

struct Error {
    int code;
    std::string msg;
};

extern long long g_kernel_launches;

#define PLS_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            throw pls::Error{PLS_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) +     \
                                             " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"}; \
        }                                                                                         \
    } while (0)

// every kernel launch of the library is followed by this check, which also counts it
#define PLS_CHECK_LAUNCH()            \
    do {                              \
        ++pls::g_kernel_launches;     \
        PLS_CUDA(cudaGetLastError()); \
    } while (0)

#define PLS_REQUIRE(cond, text)                                   \
    do {                                                          \
        if (!(cond)) throw pls::Error{PLS_E_INVALID, (text)};     \
    } while (0)

// Growable device buffer.  Growth synchronises the stream (only before steady state).
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
    void reserve(size_t bytes, cudaStream_t s, bool keep = false);
    void reserve_exact(size_t bytes, cudaStream_t s, bool keep = false);  // no head-room: the caller planned the size
    void release();
};

// Pinned host buffer.
struct HBuf {
    void* p = nullptr;
    size_t cap = 0;
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
    void reserve(size_t bytes);
    void release();
    void* device_ptr() const;  // the device-side alias of the mapped allocation
};

struct ProfileSlot {
    bool enabled = false;
    std::vector<cudaEvent_t> pool;  // pairs
    size_t used = 0;
    double ms = 0.0;
    int64_t launches = 0;
    double bytes = 0.0;
};

// ---- radix sort / scan scratch ---------------------------------------------------------
struct SortScratch {
    DBuf keys_alt, vals_alt;  // ping-pong partners of the caller's arrays
    DBuf hist;                // [8][256] u32 histograms, then exclusive bases
    DBuf status;              // [passes][tiles][256] look-back words + tile counters
    DBuf plan;                // device-side SortPlan
};

struct ScanScratch {
    DBuf status;  // look-back words + tile counter + total
};

// select_device.cuh: epoch-tagged look-back words of the single-pass selection (one scratch per stream)
struct SelectScratch {
    DBuf status;
    uint32_t epoch = 0;
};

// ---- the kd local map: point storage + the cell-pyramid search index (kdmap_device.cuh) ---------------------
struct KdMap {
    // insertion-ordered storage, float4 (x,y,z,unused); ping-pong for the per-frame move
    DBuf store[2];
    int cur = 0;
    int64_t count = 0;                // points in store[cur]
    std::deque<int64_t> frame_counts; // per inserted frame
    // search index
    DBuf morton, order;     // u64 cell keys, u32 original index (sorted)
    DBuf sorted;            // float4 (x,y,z, bitcast original index), level-0 cell order
    DBuf normals;           // float4 (nx,ny,nz, state word) in sorted order; states carry the build generation
    DBuf bbox;              // 6 ordered-int words
    bool bbox_clean = false; // the header kernel of the last build left the box empty
    DBuf grid_hdr;          // KdGridHeader (quantisation + cell levels)
    DBuf cells;             // cell hash tables of all levels (generation-stamped entries, never cleared)
    DBuf stats;             // optional debug counters (PLS_KD_STATS=1)
    size_t table_offset[16] = {};
    uint32_t table_mask[16] = {};
    uint32_t gen = 0;       // generation of the current index build (0 = never built)
    int64_t indexed = 0;    // points covered by the index
    bool valid = false;
    int64_t cap_points = 0; // every per-point array holds this many points (kd_reserve_capacity)
    int64_t max_frame = 0;  // largest frame inserted so far (sizes the steady state: local_map_size frames)
};

struct ProjMap {
    int K = 0;              // frames held
    int head = 0;           // ring slot of the oldest frame
    DBuf vmaps, nmaps;      // [Kmax+1][3][H][W] ring; logical frame k lives in slot (head + k) % (Kmax+1)
    DBuf poses;             // device copy of [Kmax][16] poses (frame -> newest frame)
    std::vector<float> host_poses;  // [K][16]
    DBuf model_v, model_n;  // [Kmax][3][H][W] re-projected model
    DBuf zbuf;              // [Kmax][H][W] u64
    bool valid = false;
    bool zbuf_clean = false; // the query z-buffer (tmp[3]) is known to be all-empty
    int64_t built_lo = 0, built_hi = 0;  // pixel range the model currently covers (a rank of a sharded job builds its share)
};

struct FrameResult {        // mirrored to pinned host memory at the end of a frame
    float T[16];
    float params[6];
    float losses[kMaxAlign];
    int iters;
    int status;
    int done;
    int pad;
    double last_sums[NACC];
    long long counts[8];    // [0]=samples S, [1]=queries, [2]=valid rows of the frame's points
    float first_pt[4];      // first non-null pixel (vertex-map input quirk)
};

struct Comm;  // NCCL glue (comm.cu)

static_assert(sizeof(FrameResult) <= kScalarOffset, "FrameResult must fit before the scalar slots");

}  // namespace pls

struct pls_context {
    pls_config cfg;
    cudaStream_t stream = nullptr;      // the stream kernels are currently enqueued on
    cudaStream_t stream_main = nullptr; // caller-visible stream (== stream outside a map-update scope)
    cudaStream_t stream_map = nullptr;  // local-map update stream (overlaps the next frame's preprocessing)
    cudaEvent_t ev_map_done = nullptr;  // recorded on stream_map after every asynchronous map update
    cudaEvent_t ev_inputs = nullptr;    // pls_wait_stream: orders the caller's stream before this context's work
    bool map_pending = false;
    bool own_stream = false;
    std::string err;

    // staging for host<->device argument traffic
    pls::DBuf stage_in[4], stage_out[6];
    pls::HBuf pinned;       // FrameResult + small scalars
    pls::DBuf scalars;      // device scalars (counts, flags, FrameResult)

    pls::SortScratch sort;
    pls::SortScratch sort_map;          // radix-sort scratch of the map-update stream
    pls::ScanScratch scan;
    pls::SelectScratch sel[2];          // [0] main stream, [1] map-update stream
    pls::DBuf input_zbuf;               // z-buffer of the frame-input stage (kept all-empty between frames by its consumers)
    bool input_zbuf_clean = false;

    // generic scratch
    pls::DBuf tmp[8];
    // scratch of the stateless filters / alignments either side of the path (filters.cu, registration.cu, voxel
    // statistics): kept apart from tmp[] so that they never alias buffers of an in-flight frame
    pls::DBuf next_buf[8];

    // odometry state (icp_odometry.py:100-121)
    pls::KdMap kd;
    pls::ProjMap pm;
    int frame_index = 0;
    int last_icp_iters = 0;             // iterations the previous frame's ICP executed (sizes the up-front enqueue)
    bool last_sharded = false;          // the last ICP actually split its correspondences over the ranks
    int sample_pointcloud = 0;          // _sample_pointcloud
    float delta_since_update[16];       // _delta_since_map_update
    pls::DBuf frame_vmap_buf[2];        // [3][H][W] of the current frame (double-buffered: the previous one may
    pls::DBuf frame_pts_buf[2];         // still be read by the map-update stream); float4 packed valid points
    int frame_slot = 0;
    // the local-map update of the last frame, decided but not yet enqueued: the next call enqueues it once its own
    // first kernels are in flight (flush_map_update), so its ~25 us of launches leave the critical path of both calls
    bool upd_pending = false, upd_insert = false;
    float upd_T[16];
    int upd_slot = 0;
    long long upd_count = 0;
    pls::DBuf queries;                  // float4 queries P0 (owned storage)
    const float4* query_ptr = nullptr;  // the queries of the current frame (may alias frame_pts)
    pls::DBuf nn_prev;                  // previous-iteration match per query
    pls::DBuf kd_worklist;              // map points whose normal the current iteration has to compute; queued queries
    pls::DBuf kd_nn_state;              // per query: position at its last full search + runner-up bound (float4)
    pls::DBuf partials;                 // [blocks][NACC] doubles
    pls::DBuf gs_keys, gs_vals, gs_out_xyz, gs_out_idx;
    uint32_t gs_seq = 0;                // stamp of the last compact-key grid sample (overflow detection)
    pls::HBuf gs_host_xyz, gs_host_idx; // pinned + mapped staging the grid sample's gather writes directly (host callers)
    int64_t last_query_count = 0;

    pls::Comm* comm = nullptr;
    void* p2p_pending_xchg = nullptr;   // exchange buffer exported by pls_comm_p2p_handle, adopted by pls_comm_p2p_init
    pls::ProfileSlot prof[pls::kProfileSlots];
};

namespace pls {
void flush_map_update(pls_context* ctx);  // odometry.cu: enqueues the deferred local-map update of the last frame, if any
}

#define PLS_API_BEGIN(ctx)                               \
    if (!(ctx)) return PLS_E_INVALID;                    \
    try {                                                \
        cudaSetDevice((ctx)->cfg.device);                \
        pls::flush_map_update(ctx);

/* entry points of the per-frame path: they enqueue the pending map update themselves, behind their own first kernels */
#define PLS_API_BEGIN_FRAME(ctx)                         \
    if (!(ctx)) return PLS_E_INVALID;                    \
    try {                                                \
        cudaSetDevice((ctx)->cfg.device);

#define PLS_API_END(ctx)                                 \
    }                                                    \
    catch (const pls::Error& e) {                        \
        (ctx)->err = e.msg;                              \
        return e.code;                                   \
    }                                                    \
    catch (const std::exception& e) {                    \
        (ctx)->err = e.what();                           \
        return PLS_E_INVALID;                            \
    }                                                    \
    return PLS_OK;

namespace pls {

// device scalar slots (u32) living behind the FrameResult in ctx->scalars
enum { SC_GS_COUNT = 0, SC_QUERY_COUNT = 1, SC_NAN_COUNT = 2, SC_INSERT_COUNT = 3, SC_PROJ_NC = 4,
       SC_SPARE0 = 5, SC_SPARE1 = 6,
       SC_GS_OVERFLOW = 7,
       SC_KD_COUNTERS = 8,      // four u64 counters of the kd search kernels (slots 8..15)
       SC_KD_LISTS = 16,        // per-iteration work-list counters of the kd search (kdmap.cu: KDL_*), 8 words
       SC_NUM = 32 };

// ---- pointer classification + staging ---------------------------------------------------
bool is_device_ptr(const void* p);           // cached per address
bool is_device_ptr_uncached(const void* p);
// Returns a device pointer holding `bytes` of `p` (copying through `stage` if p is host).
const void* to_device(pls_context* ctx, const void* p, size_t bytes, DBuf& stage);
// Returns a device pointer results may be written to; if `p` is host, it is `stage` and
// finish_out() copies back.
struct OutArg {
    void* host = nullptr;
    void* dev = nullptr;
    size_t bytes = 0;
};
OutArg out_arg(pls_context* ctx, void* p, size_t bytes, DBuf& stage);
void finish_out(pls_context* ctx, const OutArg& o, size_t bytes_used = (size_t)-1);

inline uint32_t* scalar_u32(pls_context* ctx, int i) {
    return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ctx->scalars.p) + kScalarOffset) + i;
}
inline FrameResult* frame_result_dev(pls_context* ctx) { return reinterpret_cast<FrameResult*>(ctx->scalars.p); }
inline FrameResult* frame_result_host(pls_context* ctx) { return reinterpret_cast<FrameResult*>(ctx->pinned.p); }

// The asynchronous map-update scope: kernels enqueued inside run on stream_map with its own sort scratch.
void map_stream_begin(pls_context* ctx);
void map_stream_end(pls_context* ctx);
// Makes the main stream wait for the last asynchronous map update (no-op if none is pending).
void map_stream_wait(pls_context* ctx);
// Host-blocking: both streams idle.
void sync_all(pls_context* ctx);

// ---- profiling ------------------------------------------------------------------------------
struct ProfileScope {
    pls_context* ctx;
    int which;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    // count = false: the launch may be a device-side no-op (ICP already converged); the caller
    // credits launches/bytes afterwards with profile_credit() once the executed count is known
    ProfileScope(pls_context* c, int w, double bytes, bool count = true);
    ~ProfileScope();
};
void profile_collect(pls_context* ctx, int which);
inline void profile_credit(pls_context* ctx, int which, int64_t launches, double bytes) {
    if (!ctx->prof[which].enabled) return;
    ctx->prof[which].launches += launches;
    ctx->prof[which].bytes += bytes;
}

// ---- primitives (sort.cu / scan.cu) ------------------------------------------------------------
// Stable LSD radix sort of (u64 key, u32 value) pairs on bits [0, 8*num_passes).
// keys/vals are sorted in place from the caller's view: on return *keys_out/*vals_out point at
// the arrays holding the result (either the inputs or the scratch partners) -- they are device
// pointers stored in DEVICE memory (the plan), and also returned on the host when the number of
// executed passes is statically known (no skipping), which is how it is used here.
// cap_n >= n: the scratch is sized for cap_n elements (callers whose n grows towards a known bound pass the bound).
void radix_sort_pairs(pls_context* ctx, uint64_t* keys, uint32_t* vals, int64_t n, int num_passes,
                      uint64_t** keys_out, uint32_t** vals_out, int64_t cap_n = 0);

// Exclusive scan / stream compaction with a single-pass decoupled look-back.
// flags[i] in {0,1}; pos_out[i] = number of set flags before i; *total_dev = number set.
void exclusive_scan_flags(pls_context* ctx, const uint8_t* flags, int64_t n, uint32_t* pos_out,
                          uint32_t* total_dev);

// ---- device-side math shared by kernels ---------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// float <-> order-preserving int (for atomicMin/Max on floats)
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) {
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}
#endif

// ---- module entry points used across translation units -----------------------------------------
// projection.cu
void launch_projection(pls_context* ctx, const float* xyz, const float* channels, int batch, int64_t n,
                       int C, int H, int W, float up, float down, float* out, unsigned long long* zbuf, float fill = 0.f);
// float64 cloud [n,3] -> float32 vertex map [3,H,W]: pixel math, range comparison and validity in float64, values
// rounded to float32 at the end (what the reference does with a float64 input, icp_odometry.py:331-352)
void launch_projection_f64(pls_context* ctx, const double* xyz, int64_t n, int H, int W, float up, float down, float* out,
                           unsigned long long* zbuf);
// z-buffer pass only (no resolve): one 64-bit atomicMin of (range bits << 32 | point index) per point; n_dev (nullable)
// is a device-side point count overriding n.  The buffer must hold ~0 in every pixel beforehand.
void launch_zbuf_points(pls_context* ctx, const float* xyz, int64_t n, const uint32_t* n_dev, int H, int W, float up, float down,
                        unsigned long long* zbuf);
// normal_map.cu
void launch_normal_map(pls_context* ctx, const float* vmap, int batch, int H, int W, int ksize, float* out);
// gn.cu
struct GnParams {
    int scheme;
    double sigma;
};
// kdmap.cu
void kdmap_reset(pls_context* ctx);
// raw rows [n,3] or a vertex map [3,H,W] are packed (NaN rows / near-null pixels dropped) and
// inserted; known_count >= 0 skips the host sync that otherwise reads the packed count.
void kdmap_update(pls_context* ctx, const float* rel_pose_host, const float* pts_dev, int64_t n,
                  const float* vmap_dev, int H, int W, int64_t known_count);
// insertion of already packed float4 points (nullable) whose count the host knows
void kdmap_update_packed(pls_context* ctx, const float* rel_pose_host, const float4* fresh_dev, int64_t num_new,
                         bool has_new);
// ICP iteration `it` of the frame over ctx->query_ptr; returns the number of partial rows written
// it == 0: no previous matches; fuse_threshold >= 0: finish the iteration (sum + solve + pose update) in the last
// block of the reduction kernel (*solved tells).
int kdmap_icp_iteration(pls_context* ctx, int64_t query_bound, int rank, int num_ranks, int it, float fuse_threshold,
                        bool* solved);
// pack [n,3] rows without NaN into float4 (stable); count -> *count_dev (u32)
void pack_valid_rows(pls_context* ctx, const float* pts_dev, int64_t n, float4* out, uint32_t* count_dev);
void pack_valid_rows_f64(pls_context* ctx, const double* pts_dev, int64_t n, float4* out, uint32_t* count_dev);
// pack the non-null pixels (any channel != 0) of a [3,H,W] map into float4, row-major order
void pack_nonnull_pixels(pls_context* ctx, const float* vmap_dev, int64_t hw, float4* out, uint32_t* count_dev);
// grid_sample.cu
// compact = true sorts on 40-bit keys (5 radix passes instead of 8): exact whenever every hash lies in [-2^39, 2^39),
// i.e. voxel coordinates up to ~3 000 000 in magnitude; otherwise the kernel stamps `gs_seq` into the device scalar
// SC_GS_OVERFLOW and the caller -- which reads the sample count back anyway -- repeats the call with compact = false
// (grid_sample_overflowed() tells).
template <typename T>
void grid_sample_device(pls_context* ctx, const T* xyz_dev, int64_t n, double voxel, T* out_xyz_dev,
                        long long* out_idx_dev, bool compact = true, T* host_xyz = nullptr, long long* host_idx = nullptr);
// Reads SC_GS_COUNT (and the overflow stamp) back: one 32-byte copy + one stream sync.  Returns the sample count;
// *overflowed tells whether the last compact grid sample has to be repeated with full keys.
uint32_t grid_sample_read_count(pls_context* ctx, bool* overflowed);
// projmap.cu
void projmap_reset(pls_context* ctx);
// odometry.cu: would an ICP iteration over `work` items be split across the ranks (the rule of enqueue_icp_iterations)?
bool icp_shards(pls_context* ctx, int64_t work);
// comm.cu
int comm_rank(pls_context* ctx);
int comm_size(pls_context* ctx);
void projmap_update(pls_context* ctx, const float* rel_pose_host, const float* vmap_dev);
// odometry.cu
void odometry_reset(pls_context* ctx);
// comm.cu
void comm_allreduce_sums(pls_context* ctx, double* sums_dev);
void comm_free(pls_context* ctx);

}  // namespace pls
