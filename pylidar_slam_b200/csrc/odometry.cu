// The frame-to-model ICP odometry state machine (replaces ICPFrameToModel,
// slam/odometry/icp_odometry.py:72-380) driven from the host with NO per-iteration host sync:
// every ICP iteration is {correspondence+reduction kernel, [allreduce], icp_step_kernel}; the
// convergence test (icp_odometry.py:292), the Gauss-Newton guards (optimization.py:323-336) and
// the pose composition with its Euler round trip (icp_odometry.py:296-297) run on the device
// and latch a `done` flag that turns the remaining launches of the frame into no-ops.
// One device->host copy of the FrameResult ends the frame.
#include <stdlib.h>

#include <chrono>

#include "internal.cuh"
#include "icp_device.cuh"
#include "pose_device.cuh"
#include "projection_device.cuh"
#include "select_device.cuh"

namespace pls {

// projmap.cu
int projmap_icp_iteration(pls_context* ctx, int64_t query_bound, int rank, int num_ranks);
// comm.cu
int comm_rank(pls_context* ctx);
int comm_size(pls_context* ctx);
bool comm_is_p2p(pls_context* ctx);
void* comm_p2p_peers(pls_context* ctx);
unsigned long long* comm_p2p_seq(pls_context* ctx);
double* comm_allreduce_buffer(pls_context* ctx);

namespace {

int64_t g_shard_min_override = -1;  // pls_set_shard_min
int64_t shard_min_default() {
    static const int64_t v = getenv("PLS_SHARD_MIN") ? atoll(getenv("PLS_SHARD_MIN")) : 24576;
    return v;
}

__global__ void frame_begin_kernel(FrameResult* fr, const float* T0 /*16, device or null*/, int max_iters,
                                   uint32_t* worklist_counts /*2*/) {
    int t = threadIdx.x;
    if (t < 16) worklist_counts[t] = 0;  // SC_KD_COUNTERS + SC_KD_LISTS: the kd search's counters and work lists
    if (t < 16) fr->T[t] = T0 ? T0[t] : ((t % 5 == 0) ? 1.f : 0.f);
    if (t < 6) fr->params[t] = 0.f;
    if (t < kMaxAlign) fr->losses[t] = __int_as_float(0x7fc00000);
    if (t == 0) {
        fr->iters = 0;
        fr->status = 0;
        fr->done = 0;
        fr->pad = 0;  // last-block ticket of the fused correspondence+solve kernels
    }
    if (t < NACC) fr->last_sums[t] = 0.0;
}


// NCCL mode, before the all-reduce: this rank's sums go to the exchange buffer `out` (not into the FrameResult: the
// all-reduce runs in place, and launches enqueued after convergence must leave the frame's result alone).
__global__ void __launch_bounds__(256) reduce_partials_kernel(const FrameResult* fr, const double* __restrict__ partials,
                                                              int num_blocks, double* __restrict__ out) {
    if (fr->done) return;
    __shared__ double sums[NACC];
    sum_partials_256(partials, num_blocks, sums);
    __syncthreads();
    if (threadIdx.x < NACC) out[threadIdx.x] = sums[threadIdx.x];
}

// K7: normal-equation solve + ICP bookkeeping (256 threads).  num_blocks == 0: `reduced` holds the (all-reduced) sums.
__global__ void __launch_bounds__(256) icp_step_kernel(FrameResult* fr, const double* __restrict__ partials,
                                                       int num_blocks, const double* __restrict__ reduced, float threshold_delta) {
    if (fr->done) return;
    __shared__ double sums[NACC];
    if (num_blocks > 0) {
        sum_partials_256(partials, num_blocks, sums);
    } else if (threadIdx.x < NACC) {
        sums[threadIdx.x] = reduced[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < NACC) fr->last_sums[threadIdx.x] = sums[threadIdx.x];
    if (threadIdx.x != 0) return;
    icp_solve_and_update(fr, sums, threshold_delta);
}

// K9 fused: block-partial sum + ONE-SHOT all-reduce over NVLink peer memory + solve, in one kernel.
// Every rank owns an exchange buffer mapped into all peers (CUDA IPC): slot[parity][r] is written by rank r.
// A rank stores its 30 sums into slot[parity][me] of EVERY peer (plain stores to peer-mapped addresses), fences
// system-wide, then stores the launch's sequence number; it then spins on its LOCAL slots until all peers'
// sequence numbers arrived and adds the slots in rank order -- the same order on every rank, so all ranks
// obtain bit-identical sums and hence bit-identical poses without any broadcast.  Two parities make the
// overwrite of a slot wait for a full further round.  The spin is bounded: a peer that never shows up turns
// into PLS_E_COMM instead of a hang.
struct P2PSlot {
    double sums[NACC];
    unsigned long long seq;
    unsigned long long pad;
};
static_assert(sizeof(P2PSlot) == 256, "P2PSlot must match comm.cu's kP2PSlotBytes");
__global__ void __launch_bounds__(256)
icp_step_p2p_kernel(FrameResult* fr, const double* __restrict__ partials, int num_blocks, float threshold_delta,
                    P2PSlot* const* __restrict__ peers, int world, int rank, unsigned long long* seq_counter) {
    if (fr->done) return;
    __shared__ double sums[NACC];
    __shared__ int timed_out;
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) {
        timed_out = 0;
        s_seq = ++(*seq_counter);  // this exchange's round
    }
    sum_partials_256(partials, num_blocks, sums);
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int parity = (int)(seq & 1ull);
    if (threadIdx.x < NACC)
        for (int r = 0; r < world; ++r) peers[r][parity * world + rank].sums[threadIdx.x] = sums[threadIdx.x];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        volatile unsigned long long* flag = &peers[threadIdx.x][parity * world + rank].seq;
        *flag = seq;
    }
    if (threadIdx.x < world) {
        volatile unsigned long long* mine = &peers[rank][parity * world + threadIdx.x].seq;
        const long long t0 = clock64();
        while (*mine < seq) {
            if (clock64() - t0 > 6000000000ll) {  // ~3 s: a peer is gone
                timed_out = 1;
                break;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (timed_out) {
        if (threadIdx.x == 0) {
            fr->status = PLS_E_COMM;
            fr->done = 1;
        }
        return;
    }
    if (threadIdx.x < NACC) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcv(&peers[rank][parity * world + r].sums[threadIdx.x]);
        sums[threadIdx.x] = s;
        fr->last_sums[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    icp_solve_and_update(fr, sums, threshold_delta);
}

__global__ void scrub_vertex_map_kernel(const float* __restrict__ in, int64_t hw, float* __restrict__ out) {
    // modify_nan_pmap (utils.py:187-196): a pixel with any NaN channel becomes 0
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        float x = in[i], y = in[hw + i], z = in[2 * hw + i];
        bool bad = !(x == x) || !(y == y) || !(z == z);
        out[i] = bad ? 0.f : x;
        out[hw + i] = bad ? 0.f : y;
        out[2 * hw + i] = bad ? 0.f : z;
    }
}

__global__ void first_point_kernel(const float4* __restrict__ packed, const uint32_t* __restrict__ count,
                                   float4* __restrict__ out, uint32_t* __restrict__ out_count, FrameResult* fr) {
    // reference quirk (icp_odometry.py:342-344,356-358): with a vertex-map input `_tgt_pc` keeps
    // only the first non-null pixel
    if (threadIdx.x == 0) {
        uint32_t c = *count;
        if (c > 0) {
            out[0] = packed[0];
            fr->first_pt[0] = packed[0].x; fr->first_pt[1] = packed[0].y; fr->first_pt[2] = packed[0].z;
        }
        *out_count = c > 0 ? 1u : 0u;
    }
}

// _read_input + sample_points of a float32 point layout on the kd map after frame 0, in one selection pass over the
// input rows (select_device.cuh): the NaN-free rows (utils.py:169-184) are packed as float4 -- the points the map
// update inserts -- and, when the queries are the non-null pixels of the cloud's vertex map (icp_odometry.py:303-305),
// the rows that won their pixel of the z-buffer (closest point, lowest index on ties: projection.py:393-415) are packed
// as the queries.  The vertex map itself is never materialised: its non-null pixels ARE the winners.  (The queries come
// out in input order rather than pixel order; the reduction over them is order-independent up to fp64 rounding.)
// Each winner also resets its pixel, which leaves the z-buffer empty for the next frame.
struct FrameInputSelect {
    const float* pts;
    ProjConst pc;
    unsigned long long* zbuf;   // null: no query selection (queries = the valid rows)
    float4* frame_pts;
    float4* queries;
    struct State {
        float x, y, z;
        int pix;
    };
    __device__ __forceinline__ uint32_t flags(int64_t i, State& s) const {
        s.x = pts[3 * i];
        s.y = pts[3 * i + 1];
        s.z = pts[3 * i + 2];
        s.pix = -1;
        const bool valid = s.x == s.x && s.y == s.y && s.z == s.z;
        if (!valid) return 0u;
        uint32_t f = 1u;
        if (zbuf) {
            int pix;
            float r;
            if (project_to_pixel(s.x, s.y, s.z, pc, pix, r) &&
                zbuf[pix] == (((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(uint32_t)i)) {
                s.pix = pix;
                f |= 2u;
            }
        }
        return f;
    }
    __device__ __forceinline__ void emit(int64_t, int which, uint32_t pos, const State& s) const {
        if (which == 0) {
            frame_pts[pos] = make_float4(s.x, s.y, s.z, 0.f);
        } else {
            queries[pos] = make_float4(s.x, s.y, s.z, 0.f);
            zbuf[s.pix] = ~0ull;
        }
    }
};

}  // namespace

bool icp_shards(pls_context* ctx, int64_t work) {
    const int size = comm_size(ctx);
    int64_t shard_min = g_shard_min_override >= 0 ? g_shard_min_override : shard_min_default();
    // a query against a multi-million-point kd map walks cold cell tables and misses L2: a third of the usual share
    // already outweighs the exchange (BASELINE config 4: 131 k queries, 5 M points, still split at 8 ranks)
    if (g_shard_min_override < 0 && ctx->cfg.local_map_type == PLS_MAP_KDTREE && ctx->kd.indexed >= 2000000) shard_min /= 3;
    return size > 1 && work / size >= shard_min;
}

namespace {

inline int grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = 8 * kNumSMs;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

uint32_t* count_slot(pls_context* ctx, int i) { return reinterpret_cast<uint32_t*>(&frame_result_dev(ctx)->counts[i]); }

// Enqueues ICP iterations [first, last) (icp_odometry.py:274-297).
int enqueue_icp_iterations(pls_context* ctx, int64_t query_bound, int first, int last) {
    cudaStream_t st = ctx->stream;
    FrameResult* fr = frame_result_dev(ctx);
    int rank = comm_rank(ctx), size = comm_size(ctx);
    // Sharding pays only when a rank's share of the correspondences outweighs the exchange it buys: below
    // PLS_SHARD_MIN work items (queries / pixels) per rank every rank runs the whole iteration itself -- same inputs,
    // same deterministic kernels, hence the same bits on every rank, and no exchange at all.  (The bound is a host
    // value every rank computes alike, so the ranks always take the same branch.)
    if (size > 1) {
        const int64_t work = ctx->cfg.local_map_type == PLS_MAP_KDTREE ? query_bound : (int64_t)ctx->cfg.height * ctx->cfg.width;
        if (!icp_shards(ctx, work)) {
            rank = 0;
            size = 1;
        }
    }
    ctx->last_sharded = size > 1;
    int last_blocks = 0;
    for (int it = first; it < last; ++it) {
        int blocks;
        bool solved = false;  // the kd kernels finish the iteration themselves on a single GPU
        if (ctx->cfg.local_map_type == PLS_MAP_KDTREE)
            blocks = kdmap_icp_iteration(ctx, query_bound, rank, size, it, size == 1 ? ctx->cfg.threshold_delta_pose : -1.f, &solved);
        else
            blocks = projmap_icp_iteration(ctx, query_bound, rank, size);
        last_blocks = blocks;
        if (solved) continue;
        if (size > 1 && comm_is_p2p(ctx)) {
            icp_step_p2p_kernel<<<1, 256, 0, st>>>(fr, ctx->partials.as<double>(), blocks, ctx->cfg.threshold_delta_pose,
                                                   (P2PSlot* const*)comm_p2p_peers(ctx), size, rank, comm_p2p_seq(ctx));
            PLS_CHECK_LAUNCH();
            continue;
        }
        const double* reduced = nullptr;
        if (size > 1) {
            double* buf = comm_allreduce_buffer(ctx);
            reduce_partials_kernel<<<1, 256, 0, st>>>(fr, ctx->partials.as<double>(), blocks, buf);
            PLS_CHECK_LAUNCH();
            comm_allreduce_sums(ctx, buf);
            reduced = buf;
            blocks = 0;
        }
        icp_step_kernel<<<1, 256, 0, st>>>(fr, ctx->partials.as<double>(), blocks, reduced, ctx->cfg.threshold_delta_pose);
        PLS_CHECK_LAUNCH();
    }
    return last_blocks;
}

// PLS_HOST_TRACE=1: host-side time of the phases of a frame (enqueue up to the ICP, the wait for the pose, the
// map-update enqueue), averaged and printed every 64 frames -- a development aid for the end-to-end path.
struct HostTrace {
    bool on = getenv("PLS_HOST_TRACE") != nullptr;
    double acc[4] = {0, 0, 0, 0};
    int frames = 0, extra_rounds = 0;
    std::chrono::steady_clock::time_point t;
    void start() { if (on) t = std::chrono::steady_clock::now(); }
    void lap(int k) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        acc[k] += std::chrono::duration<double, std::micro>(now - t).count();
        t = now;
    }
    void end_frame() {
        if (!on || ++frames < 64) return;
        fprintf(stderr, "[plslam_b200 host trace] per frame: enqueue input+ICP %.1f us, wait for the pose %.1f us, "
                        "map-update enqueue %.1f us, rest %.1f us; %d of %d frames needed a second round of ICP launches\n",
                acc[0] / frames, acc[1] / frames, acc[2] / frames, acc[3] / frames, extra_rounds, frames);
        frames = 0;
        extra_rounds = 0;
        acc[0] = acc[1] = acc[2] = acc[3] = 0;
    }
};
HostTrace g_trace;

// The ICP loop (icp_odometry.py:248-299) over ctx->query_ptr / counts[1]: iterations are enqueued without
// host syncs and turn into no-ops once the device-side `done` flag latches.  To avoid paying for
// max_num_alignments launches when ICP converges in 2-3, only `previous frame's count + 1` iterations are
// enqueued up front; the rare frame that needs more continues after the result fetch (same arithmetic,
// one extra sync).  Returns the block count of the correspondence kernel.
int run_icp(pls_context* ctx, const float* T0_dev, int64_t query_bound) {
    cudaStream_t st = ctx->stream;
    FrameResult* fr = frame_result_dev(ctx);
    frame_begin_kernel<<<1, kMaxAlign, 0, st>>>(fr, T0_dev, ctx->cfg.max_num_alignments, scalar_u32(ctx, SC_KD_COUNTERS));
    PLS_CHECK_LAUNCH();
    if (query_bound < 1) query_bound = 1;
    ctx->pm.zbuf_clean = false;  // tmp[3] may have been used by the frame's own projection
    ctx->nn_prev.reserve((size_t)query_bound * sizeof(int), st);  // previous matches: ignored by iteration 0
    const int max_it = ctx->cfg.max_num_alignments;
    static const bool all_upfront = getenv("PLS_ICP_UPFRONT_ALL") != nullptr;
    int upfront = (ctx->last_icp_iters > 0 && !all_upfront) ? ctx->last_icp_iters + 1 : max_it;
    if (upfront > max_it) upfront = max_it;
    int blocks = enqueue_icp_iterations(ctx, query_bound, 0, upfront);
    int enq = upfront;
    g_trace.lap(0);
    while (enq < max_it) {
        // continue only if the device has not latched `done` (checked on the host: rare path)
        int flags[3];
        PLS_CUDA(cudaMemcpyAsync(flags, &fr->iters, sizeof(flags), cudaMemcpyDeviceToHost, st));
        PLS_CUDA(cudaStreamSynchronize(st));
        if (flags[2] /*done*/) break;
        g_trace.extra_rounds += 1;
        const int more = (max_it - enq) < 4 ? (max_it - enq) : 4;
        blocks = enqueue_icp_iterations(ctx, query_bound, enq, enq + more);
        enq += more;
    }
    return blocks;
}

void fetch_result(pls_context* ctx) {
    // the FrameResult and the u32 / u64 scalar slots behind it in one copy
    PLS_CUDA(cudaMemcpyAsync(ctx->pinned.p, ctx->scalars.p, kScalarOffset + SC_NUM * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             ctx->stream));
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
}

// Algorithmic bytes of the frame's executed ICP iterations, by SURVEY.md 8d's formulas.
// kd map:  K5a exact 1-NN    per query 12 B (query) + 27 * 8 B (cell-range lookups) + 4 B (match), plus 16 B per
//                            candidate actually tested (counted by the kernel);
//          K5b/c normals     per computed normal 16 B (point) + 16 B (normal written), plus 16 B per candidate tested;
//          K6 reduction      per correspondence 36 B, plus one 30-double partial row per block.
// (bench.py also reports the lower bound SURVEY names: 44 B per query + 16 B per touched map point.)
void credit_icp_profile(pls_context* ctx, const FrameResult* h, int blocks) {
    if (ctx->cfg.local_map_type == PLS_MAP_KDTREE) {
        const unsigned long long* kc = reinterpret_cast<const unsigned long long*>(
            reinterpret_cast<const char*>(h) + kScalarOffset + SC_KD_COUNTERS * sizeof(uint32_t));
        // a rank of a sharded frame handles its share of the queries (the counters already are per rank)
        const double iters = (double)h->iters, nq = (double)h->counts[1] / (ctx->last_sharded ? (double)comm_size(ctx) : 1.0);
        const double bytes = iters * nq * (12.0 + 27.0 * 8.0 + 4.0) + 16.0 * (double)kc[0]      // K5a
                             + 32.0 * (double)kc[2] + 16.0 * (double)kc[1]                      // K5b/c
                             + iters * (nq * 36.0 + (double)blocks * NACC * 8.0);               // K6
        profile_credit(ctx, 0, h->iters, bytes);
    } else {
        // projective map: HW*12*(K+1) (target + K candidate vertex maps, each read once)
        // + N_c*12 (winner normals) + one partial row per block
        // a rank of a sharded frame streams its share of the tiles
        const double share = ctx->last_sharded ? 1.0 / (double)comm_size(ctx) : 1.0;
        const double hw = (double)ctx->cfg.height * ctx->cfg.width * share;
        profile_credit(ctx, 1, h->iters, (double)h->iters * (hw * 12.0 * (ctx->pm.K + 1) + h->last_sums[29] * share * 12.0 +
                                                              (double)blocks * NACC * 8.0));
    }
}

void raise_status(pls_context* ctx, int status) {
    if (status == PLS_E_COMM) throw pls::Error{PLS_E_COMM, "peer-to-peer all-reduce timed out waiting for a peer rank"};
    if (status == PLS_E_SINGULAR) throw pls::Error{PLS_E_SINGULAR, "Invalid Jacobian in Gauss Newton minimization"};
}

// __update_map (icp_odometry.py:360-380), host side: key-frame policy on the accumulated motion.
bool keyframe_decision(pls_context* ctx, const float* T) {
    float nd[16], prm[6];
    mat4_mul(ctx->delta_since_update, T, nd);
    from_pose(nd, prm);
    const float tn = sqrtf(prm[0] * prm[0] + prm[1] * prm[1] + prm[2] * prm[2]);
    const float rn = sqrtf(prm[3] * prm[3] + prm[4] * prm[4] + prm[5] * prm[5]);
    const bool insert = tn > ctx->cfg.threshold_trans || rn * 180.0f / 3.14159265358979323846f > ctx->cfg.threshold_rot;
    if (insert) {
        for (int i = 0; i < 16; ++i) ctx->delta_since_update[i] = (i % 5 == 0) ? 1.f : 0.f;
    } else {
        memcpy(ctx->delta_since_update, nd, sizeof(nd));
    }
    return insert;
}

}  // namespace

// The deferred local-map update (ICPFrameToModel.__update_map, icp_odometry.py:360-380) of the last frame: move / append /
// evict + index rebuild, or the projective model rebuild, on the map stream.
void flush_map_update(pls_context* ctx) {
    if (!ctx->upd_pending) return;
    ctx->upd_pending = false;
    const bool kd = ctx->cfg.local_map_type == PLS_MAP_KDTREE;
    DBuf& frame_vmap = ctx->frame_vmap_buf[ctx->upd_slot];
    DBuf& frame_pts = ctx->frame_pts_buf[ctx->upd_slot];
    map_stream_begin(ctx);
    try {
        if (kd) {
            if (ctx->upd_insert) kdmap_update_packed(ctx, ctx->upd_T, frame_pts.as<float4>(), (int64_t)ctx->upd_count, true);
            else kdmap_update_packed(ctx, ctx->upd_T, nullptr, 0, false);
        } else {
            projmap_update(ctx, ctx->upd_T, ctx->upd_insert ? frame_vmap.as<float>() : nullptr);
        }
    } catch (...) {
        map_stream_end(ctx);
        throw;
    }
    map_stream_end(ctx);
}

namespace {

void process_frame_device(pls_context* ctx, const void* data_void, int layout, int64_t n, const float* init_pose,
                          float* out_pose, float* out_params, int* out_has_pose, double* out_info) {
    cudaStream_t st = ctx->stream;
    g_trace.start();
    // float64 point layouts: same flow, the cloud is rounded to float32 for the queries / map insertion while the frame's
    // own vertex map is projected in float64 (icp_odometry.py:331-352)
    const bool is64 = layout == PLS_INPUT_NDARRAY_F64 || layout == PLS_INPUT_TENSOR_F64;
    if (layout == PLS_INPUT_NDARRAY_F64) layout = PLS_INPUT_NDARRAY;
    if (layout == PLS_INPUT_TENSOR_F64) layout = PLS_INPUT_TENSOR;
    const float* data_dev = is64 ? nullptr : (const float*)data_void;
    const double* data64 = is64 ? (const double*)data_void : nullptr;
    const int H = ctx->cfg.height, W = ctx->cfg.width;
    const int64_t hw = (int64_t)H * W;
    const bool kd = ctx->cfg.local_map_type == PLS_MAP_KDTREE;
    FrameResult* fr = frame_result_dev(ctx);
    PLS_CUDA(cudaMemsetAsync(fr->counts, 0, sizeof(fr->counts), st));
    if (layout == PLS_INPUT_NDARRAY) ctx->sample_pointcloud = 1;  // icp_odometry.py:330
    const bool first = ctx->frame_index == 0;
    // the previous frame's buffers may still feed the asynchronous map update: use the other pair
    ctx->frame_slot ^= 1;
    DBuf& frame_vmap = ctx->frame_vmap_buf[ctx->frame_slot];
    DBuf& frame_pts = ctx->frame_pts_buf[ctx->frame_slot];

    // ---- _read_input (icp_odometry.py:319-358)
    frame_vmap.reserve((size_t)3 * hw * sizeof(float), st);
    int64_t pts_bound = 0;
    bool fused_input = false;
    if (layout == PLS_INPUT_VERTEX_MAP) {
        scrub_vertex_map_kernel<<<grid_for(hw), 256, 0, st>>>(data_dev, hw, frame_vmap.as<float>());
        PLS_CHECK_LAUNCH();
        ctx->tmp[5].reserve((size_t)hw * sizeof(float4), st);
        pack_nonnull_pixels(ctx, frame_vmap.as<float>(), hw, ctx->tmp[5].as<float4>(), count_slot(ctx, 1));
        frame_pts.reserve(sizeof(float4) * 4, st);
        first_point_kernel<<<1, 32, 0, st>>>(ctx->tmp[5].as<float4>(), count_slot(ctx, 1), frame_pts.as<float4>(),
                                              count_slot(ctx, 2), fr);
        PLS_CHECK_LAUNCH();
        pts_bound = 1;
    } else {
        PLS_REQUIRE(n > 0, "process_frame: empty point cloud");
        frame_pts.reserve((size_t)n * sizeof(float4), st);
        pts_bound = n;
        // the shipped pipelines (float32 points, kd map, any frame but the first): one z-buffer pass and one selection
        fused_input = !is64 && kd && !first && n <= SEL_MAX_N && n < (1ll << 32);
        if (fused_input) {
            const bool pixel_queries = !ctx->sample_pointcloud;
            FrameInputSelect op;
            op.pts = data_dev;
            op.pc = make_proj_const(H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg);
            op.zbuf = nullptr;
            op.frame_pts = frame_pts.as<float4>();
            op.queries = nullptr;
            if (pixel_queries) {
                if (ctx->input_zbuf.cap < (size_t)hw * sizeof(unsigned long long)) ctx->input_zbuf_clean = false;
                ctx->input_zbuf.reserve((size_t)hw * sizeof(unsigned long long), st);
                if (!ctx->input_zbuf_clean)
                    PLS_CUDA(cudaMemsetAsync(ctx->input_zbuf.p, 0xff, (size_t)hw * sizeof(unsigned long long), st));
                ctx->input_zbuf_clean = false;  // dirty until the selection below (whose winners reset their pixels) is enqueued
                ctx->queries.reserve((size_t)(n < hw ? n : hw) * sizeof(float4), st);
                launch_zbuf_points(ctx, data_dev, n, nullptr, H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg,
                                   ctx->input_zbuf.as<unsigned long long>());
                op.zbuf = ctx->input_zbuf.as<unsigned long long>();
                op.queries = ctx->queries.as<float4>();
            }
            select_launch(ctx, op, n, nullptr, count_slot(ctx, 2), pixel_queries ? count_slot(ctx, 1) : nullptr);
            if (pixel_queries) ctx->input_zbuf_clean = true;
        } else if (is64) {
            pack_valid_rows_f64(ctx, data64, n, frame_pts.as<float4>(), count_slot(ctx, 2));
        } else {
            pack_valid_rows(ctx, data_dev, n, frame_pts.as<float4>(), count_slot(ctx, 2));
        }
        // the vertex map of the points is needed on frame 0 (map initialisation), as the query
        // source when _sample_pointcloud is False, and by the projective map's update
        if (!fused_input && (first || !ctx->sample_pointcloud || !kd)) {
            ctx->tmp[3].reserve((size_t)hw * sizeof(unsigned long long), st);
            if (is64)
                launch_projection_f64(ctx, data64, n, H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg, frame_vmap.as<float>(),
                                      ctx->tmp[3].as<unsigned long long>());
            else
                launch_projection(ctx, data_dev, nullptr, 1, n, 3, H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg,
                                  frame_vmap.as<float>(), ctx->tmp[3].as<unsigned long long>());
        }
    }

    float eye[16];
    for (int i = 0; i < 16; ++i) eye[i] = (i % 5 == 0) ? 1.f : 0.f;

    if (first) {
        // icp_odometry.py:171-181: the first frame only initialises the map, via its vertex map
        flush_map_update(ctx);
        map_stream_wait(ctx);
        if (kd) kdmap_update(ctx, eye, nullptr, 0, frame_vmap.as<float>(), H, W, -1);
        else projmap_update(ctx, eye, frame_vmap.as<float>());
        ctx->frame_index = 1;
        if (out_has_pose) *out_has_pose = 0;
        if (out_pose) memcpy(out_pose, eye, sizeof(eye));
        if (out_params) memset(out_params, 0, 6 * sizeof(float));
        fetch_result(ctx);
        if (out_info) {
            FrameResult* h = frame_result_host(ctx);
            for (int i = 0; i < 12; ++i) out_info[i] = 0.0;
            out_info[3] = (double)ctx->kd.count;
            out_info[5] = (double)(pts_bound - (int64_t)h->counts[2]);
        }
        return;
    }

    // ---- sample_points (icp_odometry.py:301-308)
    int64_t query_bound;
    if (ctx->sample_pointcloud && layout != PLS_INPUT_VERTEX_MAP) {
        // queries = the (NaN-free) input points themselves
        ctx->query_ptr = frame_pts.as<float4>();
        PLS_CUDA(cudaMemcpyAsync(count_slot(ctx, 1), count_slot(ctx, 2), sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        query_bound = n;
    } else if (layout == PLS_INPUT_VERTEX_MAP) {
        ctx->query_ptr = ctx->tmp[5].as<float4>();
        query_bound = hw;
    } else if (fused_input) {
        ctx->query_ptr = ctx->queries.as<float4>();  // selected together with the valid rows above
        query_bound = n < hw ? n : hw;
    } else {
        ctx->queries.reserve((size_t)hw * sizeof(float4), st);
        pack_nonnull_pixels(ctx, frame_vmap.as<float>(), hw, ctx->queries.as<float4>(), count_slot(ctx, 1));
        ctx->query_ptr = ctx->queries.as<float4>();
        query_bound = n < hw ? n : hw;
    }

    // ---- register_new_frame
    const float* T0_dev = nullptr;
    if (init_pose) {
        ctx->tmp[6].reserve(16 * sizeof(float), st);
        PLS_CUDA(cudaMemcpyAsync(ctx->tmp[6].p, init_pose, 16 * sizeof(float), cudaMemcpyHostToDevice, st));
        T0_dev = ctx->tmp[6].as<float>();
    }
    flush_map_update(ctx);  // (already enqueued by the grid-sample call of this frame, if there was one)
    map_stream_wait(ctx);   // the ICP below reads the local map the previous frame's update is still building
    const int icp_blocks = run_icp(ctx, T0_dev, query_bound);
    fetch_result(ctx);
    g_trace.lap(1);
    FrameResult* h = frame_result_host(ctx);
    ctx->last_icp_iters = h->iters;
    credit_icp_profile(ctx, h, icp_blocks);
    raise_status(ctx, h->status);

    // ---- __update_map: decided now, enqueued (on the map stream, beside the NEXT frame's preprocessing) by the next call
    const bool insert = keyframe_decision(ctx, h->T);
    ctx->upd_pending = true;
    ctx->upd_insert = insert;
    memcpy(ctx->upd_T, h->T, sizeof(ctx->upd_T));
    ctx->upd_slot = ctx->frame_slot;
    ctx->upd_count = (long long)h->counts[2];
    static const bool eager = getenv("PLS_MAP_UPDATE_EAGER") != nullptr;  // A/B: enqueue before returning, as before
    if (eager) flush_map_update(ctx);
    g_trace.lap(2);
    ctx->frame_index += 1;
    if (out_pose) memcpy(out_pose, h->T, 16 * sizeof(float));
    if (out_params) memcpy(out_params, h->params, 6 * sizeof(float));
    if (out_has_pose) *out_has_pose = 1;
    if (out_info) {
        out_info[0] = (double)h->iters;
        out_info[1] = h->iters > 0 ? (double)h->losses[h->iters - 1] : 0.0;
        out_info[2] = (double)h->counts[1];
        out_info[3] = (double)ctx->kd.count;
        out_info[4] = (double)h->counts[0];
        out_info[5] = (double)(pts_bound - (int64_t)h->counts[2]);
        out_info[6] = (double)h->status;
        out_info[7] = insert ? 1.0 : 0.0;
        out_info[8] = h->first_pt[0]; out_info[9] = h->first_pt[1]; out_info[10] = h->first_pt[2];
        out_info[11] = ctx->last_sharded ? 1.0 : 0.0;  // the correspondences were split over the ranks
    }
    g_trace.lap(3);
    g_trace.end_frame();
}

}  // namespace

void odometry_reset(pls_context* ctx) {
    kdmap_reset(ctx);
    projmap_reset(ctx);
    ctx->frame_index = 0;
    ctx->last_icp_iters = 0;
    // _sample_pointcloud is set in the reference's constructor only: ICPFrameToModel.init() keeps it
    // (icp_odometry.py:105,128-137), so a re-initialised sequence samples like the last frame of the previous one
    for (int i = 0; i < 16; ++i) ctx->delta_since_update[i] = (i % 5 == 0) ? 1.f : 0.f;
}

}  // namespace pls

using namespace pls;

extern "C" {

int pls_map_init(pls_context* ctx) {
    PLS_API_BEGIN(ctx)
    sync_all(ctx);
    kdmap_reset(ctx);
    projmap_reset(ctx);
    PLS_API_END(ctx)
}

int pls_set_shard_min(int64_t work_items_per_rank) {
    g_shard_min_override = work_items_per_rank;
    return PLS_OK;
}

int pls_last_sharded(pls_context* ctx, int* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(out, "pls_last_sharded: null output");
    *out = ctx->last_sharded ? 1 : 0;
    PLS_API_END(ctx)
}

int pls_odometry_init(pls_context* ctx) {
    PLS_API_BEGIN(ctx)
    sync_all(ctx);
    odometry_reset(ctx);
    PLS_API_END(ctx)
}

int pls_register_frame(pls_context* ctx, const float* points, int64_t n, const float* T0, float* out_T,
                       float* out_params, float* out_losses, int* out_iters) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(points && n > 0, "pls_register_frame: points must be [n,3] with n > 0");
    PLS_REQUIRE(ctx->cfg.gn_max_iters == 1, "fused ICP path supports gauss_newton_config.max_iters == 1");
    cudaStream_t st = ctx->stream;
    map_stream_wait(ctx);
    const float* d = (const float*)to_device(ctx, points, (size_t)n * 3 * sizeof(float), ctx->stage_in[0]);
    FrameResult* fr = frame_result_dev(ctx);
    PLS_CUDA(cudaMemsetAsync(fr->counts, 0, sizeof(fr->counts), st));
    ctx->queries.reserve((size_t)n * sizeof(float4), st);
    pack_valid_rows(ctx, d, n, ctx->queries.as<float4>(), count_slot(ctx, 1));
    ctx->query_ptr = ctx->queries.as<float4>();
    const float* T0_dev = nullptr;
    if (T0) {
        ctx->tmp[6].reserve(16 * sizeof(float), st);
        PLS_CUDA(cudaMemcpyAsync(ctx->tmp[6].p, T0, 16 * sizeof(float),
                                 is_device_ptr(T0) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
        T0_dev = ctx->tmp[6].as<float>();
    }
    const int icp_blocks = run_icp(ctx, T0_dev, n);
    fetch_result(ctx);
    FrameResult* h = frame_result_host(ctx);
    credit_icp_profile(ctx, h, icp_blocks);
    auto put = [&](void* dst, const void* src, size_t bytes) {
        if (!dst) return;
        if (is_device_ptr(dst)) PLS_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
        else memcpy(dst, src, bytes);
    };
    put(out_T, h->T, 16 * sizeof(float));
    put(out_params, h->params, 6 * sizeof(float));
    put(out_losses, h->losses, (size_t)ctx->cfg.max_num_alignments * sizeof(float));
    if (out_iters) *out_iters = h->iters;
    raise_status(ctx, h->status);
    PLS_API_END(ctx)
}

int pls_process_frame(pls_context* ctx, const void* data, int layout, int64_t n, const float* init_pose,
                      float* out_pose, float* out_params, int* out_has_pose, double* out_info) {
    PLS_API_BEGIN(ctx)  // (enqueues the last frame's map update first, unless this frame's grid-sample call already did)
    PLS_REQUIRE(data, "pls_process_frame: null data");
    // optional residency hint in the high bits: the caller knows where `data` lives (a device-resident grid-sample
    // result handed over by pls_grid_sample_staged, a CUDA tensor, a numpy array) and saves the classification
    const int hint = layout & (PLS_PTR_DEVICE | PLS_PTR_HOST);
    layout &= ~(PLS_PTR_DEVICE | PLS_PTR_HOST);
    PLS_REQUIRE(layout >= PLS_INPUT_NDARRAY && layout <= PLS_INPUT_TENSOR_F64, "pls_process_frame: unknown layout");
    PLS_REQUIRE(ctx->cfg.gn_max_iters == 1, "fused ICP path supports gauss_newton_config.max_iters == 1");
    const bool is64 = layout == PLS_INPUT_NDARRAY_F64 || layout == PLS_INPUT_TENSOR_F64;
    const size_t bytes = layout == PLS_INPUT_VERTEX_MAP ? (size_t)3 * ctx->cfg.height * ctx->cfg.width * sizeof(float)
                                                        : (size_t)n * 3 * (is64 ? sizeof(double) : sizeof(float));
    const void* d = data;
    if (hint != PLS_PTR_DEVICE) {
        if (hint == PLS_PTR_HOST) {
            ctx->stage_in[0].reserve(bytes, ctx->stream);
            PLS_CUDA(cudaMemcpyAsync(ctx->stage_in[0].p, data, bytes, cudaMemcpyHostToDevice, ctx->stream));
            d = ctx->stage_in[0].p;
        } else {
            d = to_device(ctx, data, bytes, ctx->stage_in[0]);
        }
    }
    process_frame_device(ctx, d, layout, n, init_pose, out_pose, out_params, out_has_pose, out_info);
    PLS_API_END(ctx)
}

int pls_process_frame_grid_sample(pls_context* ctx, const float* raw_points, int64_t n, double voxel, int layout,
                                  const float* init_pose, float* out_pose, float* out_params, int* out_has_pose,
                                  double* out_info) {
    // (the pending map update is enqueued first here: with no host gap between the frames the next ICP would
    // otherwise wait for an index build that started a subsample's worth of launches later)
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(raw_points && n > 0 && voxel > 0.0, "pls_process_frame_grid_sample: bad arguments");
    PLS_REQUIRE(layout == PLS_INPUT_NDARRAY || layout == PLS_INPUT_TENSOR, "grid-sampled input is a point layout");
    PLS_REQUIRE(ctx->cfg.gn_max_iters == 1, "fused ICP path supports gauss_newton_config.max_iters == 1");
    const float* d = (const float*)to_device(ctx, raw_points, (size_t)n * 3 * sizeof(float), ctx->stage_in[0]);
    ctx->gs_out_xyz.reserve((size_t)n * 3 * sizeof(float), ctx->stream);
    // the sample count is needed on the host to size the point-layout frame: one small sync (it also carries the
    // overflow stamp of the 40-bit sort keys; a frame whose hashes exceed them is re-sampled on the raw keys)
    uint32_t S = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        grid_sample_device<float>(ctx, d, n, voxel, ctx->gs_out_xyz.as<float>(), nullptr, attempt == 0);
        bool overflowed = false;
        S = grid_sample_read_count(ctx, &overflowed);
        if (!(attempt == 0 && overflowed)) break;
    }
    process_frame_device(ctx, ctx->gs_out_xyz.as<float>(), layout, (int64_t)S, init_pose, out_pose, out_params,
                         out_has_pose, out_info);
    if (out_info) out_info[4] = (double)S;
    PLS_API_END(ctx)
}

}  // extern "C"
