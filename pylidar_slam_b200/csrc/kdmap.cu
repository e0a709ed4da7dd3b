// K5/K8 -- the kd local map on the GPU (replaces KdTreeLocalMap, slam/odometry/local_map.py:254-427).
//
// Per update (local_map.py:302-369): the whole map is moved by inverse(relative_pose) in float32, the new frame's
// points are appended, the oldest frame is dropped beyond local_map_size, the search index is rebuilt and the normal
// cache invalidated -- exactly the reference's life cycle, with the pykdtree build replaced by a cell-pyramid build:
//   kd_move_append_kernel   move + append + bounding box (block reduce, ordered-int atomics)
//   kd_grid_header_kernel   quantisation of the new bounding box
//   kd_cell_key_kernel      Morton id of every point's level-0 cell
//   radix sort (4 passes)   primitives.cu (stable: insertion order inside a cell)
//   kd_finalize_kernel      sorted float4 copy + the hashed cell tables of all levels
// Tables and cached normals carry the build's generation number: nothing is cleared between frames.
// Search (local_map.py:372-422): whole warps per query, see kdmap_device.cuh; one ICP iteration = three launches
// (kd_nn_warp_kernel, kd_normals_warp_kernel, kd_residual_kernel which also runs the solve in its last block).
#include <stdlib.h>

#include "gn_device.cuh"
#include "internal.cuh"
#include "icp_device.cuh"
#include "kdmap_device.cuh"
#include "pose_device.cuh"

namespace pls {

namespace {

inline int grid_for(int64_t n, int threads, int cap_blocks) {
    int64_t b = (n + threads - 1) / threads;
    return (int)(b < 1 ? 1 : (b > cap_blocks ? cap_blocks : b));
}

__global__ void kd_bbox_init_kernel(int* bbox) {
    if (threadIdx.x < 3) bbox[threadIdx.x] = 0x7fffffff;        // min (ordered int)
    else if (threadIdx.x < 6) bbox[threadIdx.x] = (int)0x80000000;  // max
}

struct Rigid {
    float R[9];
    float t[3];
};

// dst[k] = R src[k + skip] + t for k < kept;  dst[kept + j] = fresh[j] for j < num_new (from a
// device count);  bounding box of everything written.
__global__ void __launch_bounds__(256)
kd_move_append_kernel(const float4* __restrict__ src, int64_t skip, int64_t kept, Rigid X,
                      const float4* __restrict__ fresh, const uint32_t* __restrict__ num_new_dev, int64_t num_new_cap,
                      float4* __restrict__ dst, int* __restrict__ bbox) {
    const int64_t num_new = num_new_dev ? (int64_t)*num_new_dev : num_new_cap;
    const int64_t total = kept + num_new;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
        float4 p;
        if (k < kept) {
            float4 s = src[k + skip];
            p.x = X.R[0] * s.x + X.R[1] * s.y + X.R[2] * s.z + X.t[0];
            p.y = X.R[3] * s.x + X.R[4] * s.y + X.R[5] * s.z + X.t[1];
            p.z = X.R[6] * s.x + X.R[7] * s.y + X.R[8] * s.z + X.t[2];
            p.w = 0.f;
        } else {
            p = fresh[k - kept];
        }
        dst[k] = p;
        mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
        mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    __shared__ float smn[8][3], smx[8][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { smn[warp][a] = mn[a]; smx[warp][a] = mx[a]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {  // one atomic pair per axis per block
        const int a = threadIdx.x;
        float lo = smn[0][a], hi = smx[0][a];
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, smn[w][a]); hi = fmaxf(hi, smx[w][a]); }
        if (lo != FLT_MAX) atomicMin(&bbox[a], float_to_ordered(lo));
        if (hi != -FLT_MAX) atomicMax(&bbox[3 + a], float_to_ordered(hi));
    }
}

// [n,3] raw points -> float4, dropping rows containing NaN (utils.py:169-184); flags only.
// T = double: the reference rounds a float64 cloud to float32 first (`_tgt_pc = pc_data.to(torch.float32)`,
// icp_odometry.py:352) and removes NaN rows afterwards; NaN survives the rounding, so the order does not matter.
template <typename T>
__global__ void kd_valid_rows_kernel(const T* __restrict__ pts, int64_t n, uint8_t* __restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        flags[i] = (x == x && y == y && z == z) ? 1 : 0;
    }
}
template <typename T>
__global__ void kd_pack_rows_kernel(const T* __restrict__ pts, int64_t n, const uint8_t* __restrict__ flags,
                                    const uint32_t* __restrict__ pos, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) out[pos[i]] = make_float4((float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2], 0.f);
}
// vertex map [3,H,W] -> pixels with |p| > 0.01 and no NaN (local_map.py:320-328)
__global__ void kd_valid_pixels_kernel(const float* __restrict__ vmap, int64_t hw, uint8_t* __restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        float x = vmap[i], y = vmap[hw + i], z = vmap[2 * hw + i];
        float nrm = sqrtf(x * x + y * y + z * z);
        flags[i] = (nrm > 0.01f) ? 1 : 0;  // NaN compares false
    }
}
__global__ void kd_pack_pixels_kernel(const float* __restrict__ vmap, int64_t hw, const uint8_t* __restrict__ flags,
                                      const uint32_t* __restrict__ pos, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x)
        if (flags[i]) out[pos[i]] = make_float4(vmap[i], vmap[hw + i], vmap[2 * hw + i], 0.f);
}

// Quantisation of the map.  A level-0 cell is always 2^KD_MIN_B0 = 8 quantisation units, so its id (10 bits per axis,
// Morton-interleaved) fits 30 bits and the sort needs four 8-bit passes whatever the extent; the unit is chosen so
// that the cell side equals the target -- or, for maps wider than 1024 cells, so that the 13-bit range just covers the
// extent (coarser cells).  The coarsest level (top = 10) is a single cell.
__global__ void kd_grid_header_kernel(int* __restrict__ bbox, KdGridHeader* __restrict__ hdr, float cell_target) {
    if (threadIdx.x != 0) return;
    const float mnx = ordered_to_float(bbox[0]), mny = ordered_to_float(bbox[1]), mnz = ordered_to_float(bbox[2]);
    const float ex = ordered_to_float(bbox[3]) - mnx, ey = ordered_to_float(bbox[4]) - mny,
                ez = ordered_to_float(bbox[5]) - mnz;
    const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-6f));
    const int b0 = KD_MIN_B0;
    const float scale = fminf((float)(1 << b0) / cell_target, (float)KD_COORD_MAX / ext);
    hdr->mn[0] = mnx; hdr->mn[1] = mny; hdr->mn[2] = mnz;
    hdr->scale = scale;
    hdr->b0 = b0;
    hdr->cell0 = (float)(1 << b0) / scale;
    hdr->top = KD_COORD_BITS - b0;
    for (int l = 0; l < KD_MAX_LEVELS; ++l) hdr->overflow[l] = 0;
    // leave the box empty for the next update (saves that update an init launch)
    bbox[0] = bbox[1] = bbox[2] = 0x7fffffff;
    bbox[3] = bbox[4] = bbox[5] = (int)0x80000000;
}

// Sort key of a map point = the Morton id of its level-0 cell (<= 30 bits); the order inside a cell is the
// insertion order (stable sort), nothing finer is needed: every level's cell is a prefix of this id.
__global__ void kd_cell_key_kernel(const float4* __restrict__ pts, int64_t n, const KdGridHeader* __restrict__ hdr,
                                   uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const float mnx = hdr->mn[0], mny = hdr->mn[1], mnz = hdr->mn[2];
    const float scale = hdr->scale;
    const int b0 = hdr->b0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const uint32_t qx = (uint32_t)fminf(fmaxf((p.x - mnx) * scale, 0.f), (float)KD_COORD_MAX);
        const uint32_t qy = (uint32_t)fminf(fmaxf((p.y - mny) * scale, 0.f), (float)KD_COORD_MAX);
        const uint32_t qz = (uint32_t)fminf(fmaxf((p.z - mnz) * scale, 0.f), (float)KD_COORD_MAX);
        keys[i] = (uint64_t)kd_cell_id(qx >> b0, qy >> b0, qz >> b0);
        vals[i] = (uint32_t)i;
    }
}

struct CellTables {
    uint4* table[KD_MAX_LEVELS];
    uint32_t mask[KD_MAX_LEVELS];
};

// Finds or claims the slot of cell `id` in this generation's table.  Slots of older generations are free.
__device__ __forceinline__ int cell_claim(uint4* table, uint32_t mask, uint32_t id, uint32_t gen) {
    const unsigned long long want = (unsigned long long)id | ((unsigned long long)gen << 32);
    uint32_t h = kd_hash(id) & mask;
    for (int probe = 0; probe < 64; ++probe) {
        unsigned long long* word = reinterpret_cast<unsigned long long*>(&table[h]);
        unsigned long long old = *reinterpret_cast<volatile unsigned long long*>(word);
        while (true) {
            if (old == want) return (int)h;
            if ((uint32_t)(old >> 32) == gen) break;  // another cell of this generation lives here
            const unsigned long long prev = atomicCAS(word, old, want);
            if (prev == old) return (int)h;
            old = prev;
        }
        h = (h + 1) & mask;
    }
    return -1;
}

// One pass over the sorted order finishes the index: the Morton-ordered float4 copy of the points and the cell
// tables of ALL levels -- element i is the head of a level-l cell run if its id prefix differs from element i-1's,
// i.e. for every level up to (highest differing bit) / 3, and the tail likewise against element i+1; heads store
// `first`, tails store `last` into the slot they find-or-claim.
__global__ void __launch_bounds__(256)
kd_finalize_kernel(const float4* __restrict__ pts, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ order,
                   int64_t n, CellTables T, KdGridHeader* hdr, uint32_t gen, float4* __restrict__ sorted) {
    const int top = hdr->top;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t src = order[i];
        float4 p = pts[src];
        p.w = __uint_as_float(src);
        sorted[i] = p;
        const uint32_t k = (uint32_t)keys[i];
        // number of levels (from 0) at which this element starts / ends a run
        int nh = top + 1, nt = top + 1;
        if (i > 0) {
            const uint32_t d = k ^ (uint32_t)keys[i - 1];
            nh = d ? min((31 - __clz((int)d)) / 3 + 1, top + 1) : 0;
        }
        if (i + 1 < n) {
            const uint32_t d = k ^ (uint32_t)keys[i + 1];
            nt = d ? min((31 - __clz((int)d)) / 3 + 1, top + 1) : 0;
        }
        const int nl = max(nh, nt);
        for (int l = 0; l < nl; ++l) {
            const int slot = cell_claim(T.table[l], T.mask[l], k >> (3 * l), gen);
            if (slot < 0) {
                hdr->overflow[l] = 1;
            } else {
                if (l < nh) T.table[l][slot].z = (uint32_t)i;
                if (l < nt) T.table[l][slot].w = (uint32_t)i;
            }
        }
    }
}

__global__ void kd_export_kernel(const float4* __restrict__ pts, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = pts[i];
        out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
    }
}

// ---- one ICP iteration on the kd map (icp_odometry.py:275-284 + alignment.py:91-127 at x0 = 0) -----------------------
//   p = T p0; q = NN(p); n = normal(q); r = n.(p - q); J = [n, p x n]; w; reduce        -- three launches:
//   kd_nn_verify_kernel     (iterations after a frame's first) a thread per query proves that its previous match is
//                           still the nearest neighbour; the unproven ones are queued
//   kd_nn_warp_kernel       transform, exact 1-NN -> match[qi] for every query (first iteration) or the queued ones: a
//                           warp per query over the cell pyramid; the first to match a map point whose normal is not
//                           cached claims it (CAS on the state word) and queues it (warp-aggregated appends)
//   kd_normals_warp_kernel  exact (k+1)-NN of every queued map point, second moments, lane-parallel eigen-solves
//   kd_residual_kernel      a thread per query: r, J, robust weight, the 30 fp64 accumulators -> block partials; the
//                           last block sums them in fixed order and runs the solve / stop test / pose update
constexpr int KD_THREADS = 256;
constexpr int KD_WARPS = KD_THREADS / 32;

// counters of the search kernels (u64, behind the u32 scalar slots): candidates tested by the 1-NN searches, by the
// k-NN searches, normals computed -- the inputs of SURVEY 8d's algorithmic-bytes formulas
enum { KDC_NN_CAND = 0, KDC_KNN_CAND = 1, KDC_NORMALS = 2 };
// per-iteration work-list counters (u32 words at SC_KD_LISTS), one pair per list, indexed by the iteration's parity:
// the first kernel of iteration `it` zeroes the words of parity (it + 1) & 1 -- consumed by the previous iteration,
// filled by the next -- so no list is ever reset by a separate launch
enum { KDL_PENDING = 0, KDL_HARD_NN = 2, KDL_WORDS = 4 };

// Appends this block's entries (collected in shared memory by any of its threads) to a global list: one atomic per block.
__device__ __forceinline__ void block_flush_list(const int* s_list, int n, int* __restrict__ list, uint32_t* count, int* s_base) {
    if (n == 0) return;  // block-uniform
    if (threadIdx.x == 0) *s_base = (int)atomicAdd(count, (uint32_t)n);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) list[*s_base + i] = s_list[i];
}

// Claims the normal of map point `pos` for computation if it is neither cached nor already claimed.
__device__ __forceinline__ bool claim_normal(const KdIndex& ix, int pos) {
    const uint32_t claimed = kd_normal_claimed(ix.gen), valid = kd_normal_valid(ix.gen);
    uint32_t* w = reinterpret_cast<uint32_t*>(&ix.normals[pos].w);
    const uint32_t cur = __ldcg(w);
    return cur != valid && cur != claimed && atomicCAS(w, cur, claimed) == cur;
}

// 1-NN of ICP iterations after a frame's first: VERIFY instead of searching.  The full search stored, per query, where
// the query stood (its transformed position) and a lower bound of the distance to every map point other than its match.
// If the query has moved by eps since then and its match is now at distance d, every other point is still at least
// (bound - eps) away, so d + eps < bound proves the match unchanged -- one point load and a dozen flops per query.
// The few queries that cannot be proven are queued for the full search.
__global__ void __launch_bounds__(KD_THREADS)
kd_nn_verify_kernel(KdIndex ix, const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev, int64_t q_begin,
                    int64_t q_stride, const float* __restrict__ T, const int* __restrict__ done, const int* __restrict__ match,
                    const float4* __restrict__ nn_state, int* __restrict__ hard, uint32_t* lists, int parity) {
    if (done && *done) return;
    __shared__ float sT[12];
    __shared__ int s_hard[KD_THREADS];
    __shared__ int s_nh, s_base;
    if (threadIdx.x < 12) sT[threadIdx.x] = T[threadIdx.x];
    if (threadIdx.x == 0) {
        s_nh = 0;
        if (blockIdx.x == 0)
            for (int l = 0; l < KDL_WORDS; l += 2) lists[l + (parity ^ 1)] = 0;
    }
    __syncthreads();
    const int64_t nq = (int64_t)*nq_dev;
    const int64_t qi = q_begin + ((int64_t)blockIdx.x * KD_THREADS + threadIdx.x) * q_stride;
    if (qi < nq) {
        const float4 p0 = queries[qi];
        const float px = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
        const float py = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
        const float pz = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
        const int pos = match[qi];
        bool proven = false;
        if (pos >= 0) {
            const float4 s = nn_state[qi];
            const float d = sqrtf(dist2_point(px, py, pz, __ldg(ix.sorted + pos)));
            const float eps = sqrtf(dist2_point(px, py, pz, s));
            proven = (d + eps) * 1.00001f + 1e-6f < sqrtf(s.w);
        }
        if (!proven) s_hard[atomicAdd(&s_nh, 1)] = (int)qi;
    }
    __syncthreads();
    block_flush_list(s_hard, s_nh, hard, lists + KDL_HARD_NN + parity, &s_base);
}

// 1-NN, full search: a warp per query over the cell pyramid (warp_nearest).  hard == nullptr: every query of this
// rank's shard (a frame's first iteration); else the queued ones, seeded with their previous match.  The first warp to
// match a map point whose normal is not cached claims it (CAS on the state word) and queues it; claims are made by all
// lanes at once after 32 queries.  Each query's position and runner-up bound are kept for the later iterations' checks.
__global__ void __launch_bounds__(KD_THREADS)
kd_nn_warp_kernel(KdIndex ix, const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev, int64_t q_begin,
                  int64_t q_stride, const int* __restrict__ hard, uint32_t* lists, int parity, const float* __restrict__ T,
                  const int* __restrict__ done, int* __restrict__ match, float4* __restrict__ nn_state, int want_normals,
                  int* __restrict__ pending, unsigned long long* __restrict__ counters) {
    if (done && *done) return;
    int n;
    if (hard) {
        n = (int)lists[KDL_HARD_NN + parity];
    } else {
        const int64_t nq = (int64_t)*nq_dev;
        n = nq > q_begin ? (int)((nq - q_begin + q_stride - 1) / q_stride) : 0;
        if (blockIdx.x == 0 && threadIdx.x == 0)  // first kernel of the iteration: recycle the other parity's lists
            for (int l = 0; l < KDL_WORDS; l += 2) lists[l + (parity ^ 1)] = 0;
    }
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * KD_WARPS + (threadIdx.x >> 5);
    const int total_warps = gridDim.x * KD_WARPS;
    if (warp_global >= n) return;
    float t[12];
#pragma unroll
    for (int a = 0; a < 12; ++a) t[a] = T[a];
    const KdGridLocal g = kd_load_grid(ix);
    uint32_t* pending_count = lists + KDL_PENDING + parity;
    int my_pos = -1, held = 0, cand = 0;
    auto flush_claims = [&]() {
        const bool mine = want_normals && my_pos >= 0 && claim_normal(ix, my_pos);
        const unsigned m = __ballot_sync(FULL, mine);
        if (m) {
            int base = 0;
            if (lane == 0) base = (int)atomicAdd(pending_count, (uint32_t)__popc(m));
            base = __shfl_sync(FULL, base, 0);
            if (mine) pending[base + __popc(m & ((1u << lane) - 1u))] = my_pos;
        }
        my_pos = -1;
        held = 0;
    };
    // the next query's data is fetched while this one is searched
    int s = warp_global;
    int64_t qi = hard ? (int64_t)hard[s] : q_begin + (int64_t)s * q_stride;
    float4 p0 = queries[qi];
    int hint = hard ? match[qi] : -1;
    while (true) {
        const int sn = s + total_warps;
        int64_t qn = qi;
        float4 pn = p0;
        int hn = -1;
        if (sn < n) {
            qn = hard ? (int64_t)hard[sn] : q_begin + (int64_t)sn * q_stride;
            pn = queries[qn];
            if (hard) hn = match[qn];
        }
        const float px = p0.x * t[0] + p0.y * t[1] + p0.z * t[2] + t[3];
        const float py = p0.x * t[4] + p0.y * t[5] + p0.z * t[6] + t[7];
        const float pz = p0.x * t[8] + p0.y * t[9] + p0.z * t[10] + t[11];
        float second;
        const int pos = warp_nearest(ix, g, px, py, pz, hint, lane, &cand, &second);
        if (lane == 0) {
            match[qi] = pos;
            if (nn_state) nn_state[qi] = make_float4(px, py, pz, second);
        }
        if (lane == held) my_pos = pos;
        if (++held == 32) flush_claims();
        if (sn >= n) break;
        s = sn;
        qi = qn;
        p0 = pn;
        hint = hn;
    }
    flush_claims();
    if (counters && lane == 0 && cand) atomicAdd(counters + KDC_NN_CAND, (unsigned long long)cand);
}

// Normals: a warp per queued map point, exact (k+1)-NN over the cell pyramid (warp_knn), second moments; the
// eigen-solves are deferred and run lane-parallel (each lane one point) so that no warp idles behind a serial solve.
__global__ void __launch_bounds__(KD_THREADS)
kd_normals_warp_kernel(KdIndex ix, int k_normals, const int* __restrict__ worklist, const uint32_t* __restrict__ wl_count,
                       const int* __restrict__ done, unsigned long long* __restrict__ counters) {
    if (done && *done) return;
    const int n = (int)*wl_count;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * KD_WARPS + (threadIdx.x >> 5);
    const int total_warps = gridDim.x * KD_WARPS;
    if (warp_global >= n) return;
    const KdGridLocal g = kd_load_grid(ix);
    const float valid = __uint_as_float(kd_normal_valid(ix.gen));
    float mycov[6];
    int mypos = -1, held = 0, cand = 0, done_here = 0;
    int e = warp_global;
    int pos = worklist[e];
    float4 c = __ldg(ix.sorted + pos);
    while (true) {
        // the next point is fetched while this one is searched
        const int en = e + total_warps;
        int posn = 0;
        float4 cn = c;
        if (en < n) {
            posn = worklist[en];
            cn = __ldg(ix.sorted + posn);
        }
        float nd;
        int ni;
        const int found = warp_knn(ix, g, c.x, c.y, c.z, k_normals + 1, lane, nd, ni, &cand);
        float cov[6];
        warp_second_moments(ix, c, k_normals, found, ni, lane, cov);
        if (lane == held) {
#pragma unroll
            for (int a = 0; a < 6; ++a) mycov[a] = cov[a];
            mypos = pos;
        }
        ++done_here;
        if (++held == 32) {  // 32 moments collected: every lane solves its own
            float nn[3];
            smallest_eigenvector(mycov, nn);
            __stcg(ix.normals + mypos, make_float4(nn[0], nn[1], nn[2], valid));
            held = 0;
            mypos = -1;
        }
        if (en >= n) break;
        e = en;
        pos = posn;
        c = cn;
    }
    if (mypos >= 0) {
        float nn[3];
        smallest_eigenvector(mycov, nn);
        __stcg(ix.normals + mypos, make_float4(nn[0], nn[1], nn[2], valid));
    }
    if (counters && lane == 0) {
        atomicAdd(counters + KDC_KNN_CAND, (unsigned long long)cand);
        atomicAdd(counters + KDC_NORMALS, (unsigned long long)done_here);
    }
}

constexpr int KD_RES_THREADS = 256;

__global__ void __launch_bounds__(KD_RES_THREADS)
kd_residual_kernel(KdIndex ix, const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev, int64_t q_begin,
                   int64_t q_stride, FrameResult* fr, int scheme, float sigma,
                   const int* __restrict__ match, double* __restrict__ partials, float fuse_threshold) {
    if (fr->done) return;
    __shared__ float sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = fr->T[threadIdx.x];
    __syncthreads();
    const int64_t nq = (int64_t)*nq_dev;
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t qi = q_begin + s * q_stride;
        if (qi >= nq) break;
        const float4 p0 = queries[qi];
        float p[3];
        p[0] = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
        p[1] = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
        p[2] = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
        const int pos = match[qi];
        if (pos < 0) continue;
        const float4 qq = __ldg(ix.sorted + pos);
        const float4 nv = __ldcg(ix.normals + pos);
        float q[3] = {qq.x, qq.y, qq.z};
        float nn[3] = {nv.x, nv.y, nv.z};
        float J[6];
        const float r = p2plane_residual_jacobian_identity(p, q, nn, J);
        const float w = ls_weight<float>(scheme, sigma, r, p, q);
        accumulate_normal_equations<float>(acc, J, w, r * w, r);
    }
    block_reduce_store<KD_RES_THREADS>(acc, partials + (size_t)blockIdx.x * NACC);
    if (fuse_threshold >= 0.f) icp_finish_in_last_block(fr, partials, fuse_threshold);
}

// Fine-grained API: [n,3] rows -> float4 queries (no row is dropped: outputs stay aligned with the inputs)
__global__ void kd_rows_to_float4_kernel(const float* __restrict__ rows, int64_t n, float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = make_float4(rows[3 * i], rows[3 * i + 1], rows[3 * i + 2], 0.f);
}

__global__ void kd_search_export_kernel(KdIndex ix, const int* __restrict__ match, int64_t n, float* __restrict__ out_nb,
                                        float* __restrict__ out_nrm, long long* __restrict__ out_idx) {
    const float nan = __int_as_float(0x7fc00000);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int pos = match[i];
        float4 q = make_float4(nan, nan, nan, 0.f), nv = make_float4(nan, nan, nan, 0.f);
        long long idx = -1;
        if (pos >= 0) {
            q = __ldg(ix.sorted + pos);
            idx = (long long)__float_as_uint(q.w);
            if (out_nrm) nv = __ldcg(ix.normals + pos);
        }
        out_nb[3 * i] = q.x; out_nb[3 * i + 1] = q.y; out_nb[3 * i + 2] = q.z;
        if (out_idx) out_idx[i] = idx;
        if (out_nrm) { out_nrm[3 * i] = nv.x; out_nrm[3 * i + 1] = nv.y; out_nrm[3 * i + 2] = nv.z; }
    }
}

size_t cell_table_bytes(int64_t M, uint32_t* masks, size_t* offsets) {
    size_t off = 0;
    for (int l = 0; l < KD_MAX_LEVELS; ++l) {
        uint64_t want = (uint64_t)(1.5 * (double)M) >> l;
        uint32_t sz = 64;
        while (sz < want) sz <<= 1;
        if (masks) masks[l] = sz - 1;
        if (offsets) offsets[l] = off;
        off += (size_t)sz * sizeof(uint4);
    }
    return off;
}

// Sizes every per-point array of the map at once.  The map grows frame by frame until local_map_size frames are
// held, then oscillates around that size: reserving the steady state (local_map_size + 1 frames of the largest
// frame seen, 30 % head-room) on the first insertion means NO further allocation -- and none of the stream
// synchronisations an allocation implies -- while the map fills up, i.e. inside any timed region that starts after
// the first frame.  `need` beyond the plan (a denser frame later on) re-plans with 25 % head-room.
// Tables and normal states are generation-stamped and never cleared per build, so fresh memory is zeroed here once
// (generation 0 is never used).
void kd_reserve_capacity(pls_context* ctx, int64_t need) {
    KdMap& kd = ctx->kd;
    if (need <= kd.cap_points) return;
    cudaStream_t st = ctx->stream;
    int64_t steady = (int64_t)(1.3 * (double)kd.max_frame * (double)(ctx->cfg.local_map_size + 1));
    if (steady > need + (8ll << 20)) steady = need + (8ll << 20);  // a multi-million-point insertion plans 8 M ahead at most
    int64_t cap = need + need / 4 + 64;
    if (cap < steady) cap = steady;
    const size_t C = (size_t)cap;
    kd.store[kd.cur].reserve_exact(C * sizeof(float4), st, true);  // the live points survive
    kd.store[kd.cur ^ 1].reserve_exact(C * sizeof(float4), st);
    kd.morton.reserve_exact(C * sizeof(uint64_t), st);
    kd.order.reserve_exact(C * sizeof(uint32_t), st);
    kd.sorted.reserve_exact(C * sizeof(float4), st);
    kd.normals.reserve_exact(C * sizeof(float4), st);
    const size_t table_bytes = cell_table_bytes(cap, nullptr, nullptr);
    kd.cells.reserve_exact(table_bytes, st);
    PLS_CUDA(cudaMemsetAsync(kd.normals.p, 0, C * sizeof(float4), st));
    PLS_CUDA(cudaMemsetAsync(kd.cells.p, 0, table_bytes, st));
    kd.cap_points = cap;
}

KdIndex make_index(pls_context* ctx) {
    KdIndex ix;
    ix.sorted = ctx->kd.sorted.as<float4>();
    ix.normals = ctx->kd.normals.as<float4>();
    ix.M = (int)ctx->kd.indexed;
    ix.gen = ctx->kd.gen;
    ix.grid = ctx->kd.grid_hdr.as<KdGridHeader>();
    for (int l = 0; l < KD_MAX_LEVELS; ++l) {
        ix.table[l] = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(ctx->kd.cells.p) + ctx->kd.table_offset[l]);
        ix.mask[l] = ctx->kd.table_mask[l];
    }
    ix.stats = ctx->kd.stats.p ? ctx->kd.stats.as<unsigned long long>() : nullptr;
    return ix;
}

// The per-frame index build (stands in for the KDTree rebuild of local_map.py:365-369): header, cell keys, four
// radix passes, one finishing pass.  Nothing is cleared: tables and normal states carry the build's generation.
void build_index(pls_context* ctx) {
    KdMap& kd = ctx->kd;
    cudaStream_t st = ctx->stream;
    const int64_t M = kd.count;
    kd.indexed = M;
    kd.valid = M > 0;
    if (M <= 0) return;
    ProfileScope ps(ctx, 3, (double)M * 32.0);
    PLS_REQUIRE(M < (1ll << 30), "kd map: too many points");
    kd_reserve_capacity(ctx, M);
    kd.gen += 1;
    if (kd.gen >= 0x7ffffff0u) {  // state words are 2 gen (+1): restart the generations on clean memory
        PLS_CUDA(cudaMemsetAsync(kd.normals.p, 0, kd.normals.cap, st));
        PLS_CUDA(cudaMemsetAsync(kd.cells.p, 0, kd.cells.cap, st));
        kd.gen = 1;
    }
    const float4* pts = kd.store[kd.cur].as<float4>();
    kd.grid_hdr.reserve(sizeof(KdGridHeader), st);
    static const float cell_target = getenv("PLS_KD_CELL") ? (float)atof(getenv("PLS_KD_CELL")) : KD_CELL_TARGET;
    kd_grid_header_kernel<<<1, 32, 0, st>>>(kd.bbox.as<int>(), kd.grid_hdr.as<KdGridHeader>(), cell_target);
    PLS_CHECK_LAUNCH();
    kd.bbox_clean = true;
    kd_cell_key_kernel<<<grid_for(M, 256, 8 * kNumSMs), 256, 0, st>>>(pts, M, kd.grid_hdr.as<KdGridHeader>(),
                                                                       kd.morton.as<uint64_t>(), kd.order.as<uint32_t>());
    PLS_CHECK_LAUNCH();
    uint64_t* sk;
    uint32_t* sv;
    radix_sort_pairs(ctx, kd.morton.as<uint64_t>(), kd.order.as<uint32_t>(), M, 4, &sk, &sv, kd.cap_points);
    // table geometry follows the capacity, not M: it only changes when the buffers are re-planned
    cell_table_bytes(kd.cap_points, kd.table_mask, kd.table_offset);
    CellTables T;
    for (int l = 0; l < KD_MAX_LEVELS; ++l) {
        T.table[l] = reinterpret_cast<uint4*>(reinterpret_cast<char*>(kd.cells.p) + kd.table_offset[l]);
        T.mask[l] = kd.table_mask[l];
    }
    kd_finalize_kernel<<<grid_for(M, 256, 8 * kNumSMs), 256, 0, st>>>(pts, sk, sv, M, T, kd.grid_hdr.as<KdGridHeader>(),
                                                                       kd.gen, kd.sorted.as<float4>());
    PLS_CHECK_LAUNCH();
}

}  // namespace

void kdmap_reset(pls_context* ctx) {
    if (getenv("PLS_KD_STATS") && !ctx->kd.stats.p) {
        ctx->kd.stats.reserve(16 * sizeof(unsigned long long), ctx->stream);
        cudaMemsetAsync(ctx->kd.stats.p, 0, 16 * sizeof(unsigned long long), ctx->stream);
    }
    ctx->kd.count = 0;
    ctx->kd.cur = 0;
    ctx->kd.frame_counts.clear();
    ctx->kd.indexed = 0;
    ctx->kd.valid = false;
    ctx->kd.bbox_clean = false;
    ctx->kd.max_frame = 0;   // the buffers (cap_points) and the generation counter are kept: a re-initialised
                             // sequence reuses them
}

template <typename T>
static void pack_valid_rows_impl(pls_context* ctx, const T* pts_dev, int64_t n, float4* out, uint32_t* count_dev) {
    cudaStream_t st = ctx->stream;
    if (n <= 0) {
        PLS_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(uint32_t), st));
        return;
    }
    ctx->tmp[1].reserve((size_t)n, st);
    ctx->tmp[2].reserve((size_t)n * sizeof(uint32_t), st);
    const int g = grid_for(n, 256, 8 * kNumSMs);
    kd_valid_rows_kernel<T><<<g, 256, 0, st>>>(pts_dev, n, ctx->tmp[1].as<uint8_t>());
    PLS_CHECK_LAUNCH();
    exclusive_scan_flags(ctx, ctx->tmp[1].as<uint8_t>(), n, ctx->tmp[2].as<uint32_t>(), count_dev);
    kd_pack_rows_kernel<T><<<g, 256, 0, st>>>(pts_dev, n, ctx->tmp[1].as<uint8_t>(), ctx->tmp[2].as<uint32_t>(), out);
    PLS_CHECK_LAUNCH();
}

void pack_valid_rows(pls_context* ctx, const float* pts_dev, int64_t n, float4* out, uint32_t* count_dev) {
    pack_valid_rows_impl<float>(ctx, pts_dev, n, out, count_dev);
}
void pack_valid_rows_f64(pls_context* ctx, const double* pts_dev, int64_t n, float4* out, uint32_t* count_dev) {
    pack_valid_rows_impl<double>(ctx, pts_dev, n, out, count_dev);
}

namespace {
__global__ void nonnull_pixels_kernel(const float* __restrict__ vmap, int64_t hw, uint8_t* __restrict__ flags) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        float x = vmap[i], y = vmap[hw + i], z = vmap[2 * hw + i];
        // points[points.norm(dim=-1) > 0]  (icp_odometry.py:303-305)
        flags[i] = (sqrtf(x * x + y * y + z * z) > 0.0f) ? 1 : 0;
    }
}
}  // namespace

void pack_nonnull_pixels(pls_context* ctx, const float* vmap_dev, int64_t hw, float4* out, uint32_t* count_dev) {
    cudaStream_t st = ctx->stream;
    ctx->tmp[1].reserve((size_t)hw, st);
    ctx->tmp[2].reserve((size_t)hw * sizeof(uint32_t), st);
    const int g = grid_for(hw, 256, 8 * kNumSMs);
    nonnull_pixels_kernel<<<g, 256, 0, st>>>(vmap_dev, hw, ctx->tmp[1].as<uint8_t>());
    PLS_CHECK_LAUNCH();
    exclusive_scan_flags(ctx, ctx->tmp[1].as<uint8_t>(), hw, ctx->tmp[2].as<uint32_t>(), count_dev);
    kd_pack_pixels_kernel<<<g, 256, 0, st>>>(vmap_dev, hw, ctx->tmp[1].as<uint8_t>(), ctx->tmp[2].as<uint32_t>(), out);
    PLS_CHECK_LAUNCH();
}

// Move the map by inverse(rel_pose), append `num_new` packed points, evict, rebuild the index
// (local_map.py:330-369).
void kdmap_update_packed(pls_context* ctx, const float* rel_pose_host, const float4* fresh_dev, int64_t num_new,
                         bool has_new) {
    KdMap& kd = ctx->kd;
    cudaStream_t st = ctx->stream;
    Rigid X;
    int64_t skip = 0;
    const bool first = kd.frame_counts.empty() && kd.count == 0;
    if (first) {
        for (int i = 0; i < 9; ++i) X.R[i] = (i % 4 == 0) ? 1.f : 0.f;
        X.t[0] = X.t[1] = X.t[2] = 0.f;
        kd.frame_counts.push_back(num_new);
    } else {
        float inv[16];
        rigid_inverse(rel_pose_host, inv);
        X.R[0] = inv[0]; X.R[1] = inv[1]; X.R[2] = inv[2];
        X.R[3] = inv[4]; X.R[4] = inv[5]; X.R[5] = inv[6];
        X.R[6] = inv[8]; X.R[7] = inv[9]; X.R[8] = inv[10];
        X.t[0] = inv[3]; X.t[1] = inv[7]; X.t[2] = inv[11];
        if (has_new) {
            kd.frame_counts.push_back(num_new);
            if ((int)kd.frame_counts.size() > ctx->cfg.local_map_size) {
                skip = kd.frame_counts.front();
                kd.frame_counts.pop_front();
            }
        }
    }
    if (!has_new) num_new = 0;
    const int64_t kept = kd.count - skip;
    const int64_t total = kept + num_new;
    if (num_new > kd.max_frame) kd.max_frame = num_new;
    kd_reserve_capacity(ctx, total > 0 ? total : 1);
    const int dst = kd.cur ^ 1;
    kd.bbox.reserve(8 * sizeof(int), st);
    if (!kd.bbox_clean) {
        kd_bbox_init_kernel<<<1, 32, 0, st>>>(kd.bbox.as<int>());
        PLS_CHECK_LAUNCH();
    }
    kd.bbox_clean = false;
    if (total > 0) {
        kd_move_append_kernel<<<grid_for(total, 256, 8 * kNumSMs), 256, 0, st>>>(
            kd.store[kd.cur].as<float4>(), skip, kept, X, fresh_dev, nullptr, num_new, kd.store[dst].as<float4>(),
            kd.bbox.as<int>());
        PLS_CHECK_LAUNCH();
    }
    kd.cur = dst;
    kd.count = total;
    build_index(ctx);
}

void kdmap_update(pls_context* ctx, const float* rel_pose_host, const float* pts_dev, int64_t n,
                  const float* vmap_dev, int H, int W, int64_t known_count) {
    cudaStream_t st = ctx->stream;
    const bool has_new = (pts_dev != nullptr) || (vmap_dev != nullptr);
    const int64_t cap_new = pts_dev ? n : (vmap_dev ? (int64_t)H * W : 0);
    int64_t num_new = 0;
    if (has_new && cap_new > 0) {
        ctx->tmp[4].reserve((size_t)cap_new * sizeof(float4), st);
        uint32_t* cnt = scalar_u32(ctx, SC_INSERT_COUNT);
        if (pts_dev) {
            pack_valid_rows(ctx, pts_dev, cap_new, ctx->tmp[4].as<float4>(), cnt);
        } else {
            ctx->tmp[1].reserve((size_t)cap_new, st);
            ctx->tmp[2].reserve((size_t)cap_new * sizeof(uint32_t), st);
            const int g = grid_for(cap_new, 256, 8 * kNumSMs);
            kd_valid_pixels_kernel<<<g, 256, 0, st>>>(vmap_dev, cap_new, ctx->tmp[1].as<uint8_t>());
            PLS_CHECK_LAUNCH();
            exclusive_scan_flags(ctx, ctx->tmp[1].as<uint8_t>(), cap_new, ctx->tmp[2].as<uint32_t>(), cnt);
            kd_pack_pixels_kernel<<<g, 256, 0, st>>>(vmap_dev, cap_new, ctx->tmp[1].as<uint8_t>(),
                                                      ctx->tmp[2].as<uint32_t>(), ctx->tmp[4].as<float4>());
            PLS_CHECK_LAUNCH();
        }
        if (known_count >= 0) {
            num_new = known_count;
        } else {
            uint32_t c = 0;
            PLS_CUDA(cudaMemcpyAsync(&c, cnt, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
            PLS_CUDA(cudaStreamSynchronize(st));
            num_new = c;
        }
    }
    kdmap_update_packed(ctx, rel_pose_host, ctx->tmp[4].as<float4>(), num_new, has_new);
}

// Launch geometry of the search kernels: about ONE resident wave of warps (a second, partial wave would wait for the
// first to drain), each warp looping over its share.
static int resident_blocks(const void* kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, KD_THREADS, 0) != cudaSuccess || per_sm < 1) {
        cudaGetLastError();
        per_sm = 2;
    }
    return per_sm * kNumSMs;
}

static unsigned long long* kd_counters(pls_context* ctx) {
    return reinterpret_cast<unsigned long long*>(scalar_u32(ctx, SC_KD_COUNTERS));
}

// The search of one ICP iteration (or of one fine-grained API call).  first: every query is searched; later iterations
// first verify the previous matches and search only the unproven ones.
static void launch_search(pls_context* ctx, const KdIndex& ix, const float4* queries, const uint32_t* nq_dev, int64_t mine,
                          int rank, int num_ranks, const float* T, const int* done, int* match, bool first, bool normals,
                          int parity) {
    cudaStream_t st = ctx->stream;
    const size_t slots = (size_t)mine + 64;
    ctx->kd_worklist.reserve(2 * slots * sizeof(int), st);
    ctx->kd_nn_state.reserve(slots * (size_t)num_ranks * sizeof(float4), st);  // indexed by query, not by shard slot
    int* pending = ctx->kd_worklist.as<int>();
    int* hard_nn = pending + slots;
    float4* nn_state = ctx->kd_nn_state.as<float4>();
    uint32_t* lists = scalar_u32(ctx, SC_KD_LISTS);
    unsigned long long* counters = kd_counters(ctx);
    const int tblocks = (int)((mine + KD_THREADS - 1) / KD_THREADS);
    static const int resident_nn = resident_blocks((const void*)kd_nn_warp_kernel);
    static const int resident_kn = resident_blocks((const void*)kd_normals_warp_kernel);
    const int wblocks = (int)((mine + KD_WARPS - 1) / KD_WARPS);
    if (!first) {
        ProfileScope p6(ctx, 6, 0.0);
        kd_nn_verify_kernel<<<tblocks, KD_THREADS, 0, st>>>(ix, queries, nq_dev, (int64_t)rank, (int64_t)num_ranks, T, done, match,
                                                            nn_state, hard_nn, lists, parity);
        PLS_CHECK_LAUNCH();
    }
    {
        ProfileScope p7(ctx, 7, 0.0);
        kd_nn_warp_kernel<<<wblocks < resident_nn ? wblocks : resident_nn, KD_THREADS, 0, st>>>(
            ix, queries, nq_dev, (int64_t)rank, (int64_t)num_ranks, first ? nullptr : hard_nn, lists, parity, T, done, match, nn_state,
            normals ? 1 : 0, pending, counters);
        PLS_CHECK_LAUNCH();
    }
    if (!normals) return;
    ProfileScope p9(ctx, 9, 0.0);
    kd_normals_warp_kernel<<<wblocks < resident_kn ? wblocks : resident_kn, KD_THREADS, 0, st>>>(
        ix, ctx->cfg.num_neighbors_normals, pending, lists + KDL_PENDING + parity, done, counters);
    PLS_CHECK_LAUNCH();
}

// One ICP iteration over the device-resident queries (float4 in ctx->query_ptr, count in the FrameResult); writes
// block partials to ctx->partials and returns the block count.
int kdmap_icp_iteration(pls_context* ctx, int64_t query_bound, int rank, int num_ranks, int it, float fuse_threshold,
                        bool* solved) {
    PLS_REQUIRE(ctx->kd.valid, "kd map: search before any update");
    cudaStream_t st = ctx->stream;
    const int64_t mine = (query_bound + num_ranks - 1) / num_ranks;
    FrameResult* fr = frame_result_dev(ctx);
    const uint32_t* nq_dev = reinterpret_cast<const uint32_t*>(&fr->counts[1]);
    // credited per executed iteration by the caller (the launch is a no-op once ICP converged)
    ProfileScope ps(ctx, 0, 0.0, false);
    const KdIndex ix = make_index(ctx);
    launch_search(ctx, ix, ctx->query_ptr, nq_dev, mine, rank, num_ranks, fr->T, &fr->done, ctx->nn_prev.as<int>(), it == 0, true,
                  it & 1);
    const int blocks = grid_for(mine, KD_RES_THREADS, 8 * kNumSMs);
    ctx->partials.reserve((size_t)blocks * NACC * sizeof(double), st);
    ProfileScope p10(ctx, 10, 0.0);
    kd_residual_kernel<<<blocks, KD_RES_THREADS, 0, st>>>(ix, ctx->query_ptr, nq_dev, (int64_t)rank, (int64_t)num_ranks, fr,
                                                          ctx->cfg.scheme, ctx->cfg.sigma, ctx->nn_prev.as<int>(),
                                                          ctx->partials.as<double>(), fuse_threshold);
    PLS_CHECK_LAUNCH();
    *solved = fuse_threshold >= 0.f;
    return blocks;
}

}  // namespace pls

using namespace pls;

extern "C" {

int pls_kdmap_update_points(pls_context* ctx, const float* rel_pose, const float* points, int64_t n) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(rel_pose, "pls_kdmap_update_points: rel_pose required");
    PLS_REQUIRE(ctx->cfg.local_map_type == PLS_MAP_KDTREE, "context holds a projective map");
    float rel[16];
    if (is_device_ptr(rel_pose)) PLS_CUDA(cudaMemcpy(rel, rel_pose, sizeof(rel), cudaMemcpyDeviceToHost));
    else memcpy(rel, rel_pose, sizeof(rel));
    const float* d = (points && n > 0) ? (const float*)to_device(ctx, points, (size_t)n * 3 * sizeof(float), ctx->stage_in[0]) : nullptr;
    kdmap_update(ctx, rel, d, d ? n : 0, nullptr, 0, 0, -1);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_kdmap_update_vertex_map(pls_context* ctx, const float* rel_pose, const float* vertex_map, int height, int width) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(rel_pose && vertex_map && height > 0 && width > 0, "pls_kdmap_update_vertex_map: bad arguments");
    PLS_REQUIRE(ctx->cfg.local_map_type == PLS_MAP_KDTREE, "context holds a projective map");
    float rel[16];
    if (is_device_ptr(rel_pose)) PLS_CUDA(cudaMemcpy(rel, rel_pose, sizeof(rel), cudaMemcpyDeviceToHost));
    else memcpy(rel, rel_pose, sizeof(rel));
    const float* d = (const float*)to_device(ctx, vertex_map, (size_t)3 * height * width * sizeof(float), ctx->stage_in[0]);
    kdmap_update(ctx, rel, nullptr, 0, d, height, width, -1);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_kdmap_stats(pls_context* ctx, unsigned long long* out16) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(out16, "pls_kdmap_stats: null output");
    memset(out16, 0, 16 * sizeof(unsigned long long));
    if (ctx->kd.stats.p) {
        PLS_CUDA(cudaMemcpyAsync(out16, ctx->kd.stats.p, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    PLS_API_END(ctx)
}

int pls_kdmap_size(pls_context* ctx, int64_t* num_points) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(num_points, "pls_kdmap_size: null output");
    *num_points = ctx->kd.count;
    PLS_API_END(ctx)
}

int pls_kdmap_points(pls_context* ctx, float* out) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(out, "pls_kdmap_points: null output");
    const int64_t M = ctx->kd.count;
    if (M > 0) {
        OutArg o = out_arg(ctx, out, (size_t)M * 3 * sizeof(float), ctx->stage_out[0]);
        kd_export_kernel<<<grid_for(M, 256, 8 * kNumSMs), 256, 0, ctx->stream>>>(ctx->kd.store[ctx->kd.cur].as<float4>(), M,
                                                                                 (float*)o.dev);
        PLS_CHECK_LAUNCH();
        finish_out(ctx, o);
        PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    PLS_API_END(ctx)
}

int pls_kdmap_nn_search(pls_context* ctx, const float* queries, int64_t n, float* out_neighbors, float* out_normals,
                        int64_t* out_idx) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(queries && out_neighbors && n > 0, "pls_kdmap_nn_search: bad arguments");
    if (!ctx->kd.valid) throw pls::Error{PLS_E_STATE, "pls_kdmap_nn_search: the map is empty"};
    const float* d = (const float*)to_device(ctx, queries, (size_t)n * 3 * sizeof(float), ctx->stage_in[0]);
    OutArg onb = out_arg(ctx, out_neighbors, (size_t)n * 3 * sizeof(float), ctx->stage_out[0]);
    OutArg onr = out_arg(ctx, out_normals, (size_t)n * 3 * sizeof(float), ctx->stage_out[1]);
    OutArg oix = out_arg(ctx, out_idx, (size_t)n * sizeof(int64_t), ctx->stage_out[2]);
    // the same warp-cooperative kernels as the ICP loop, with an identity transform and no previous matches
    cudaStream_t st = ctx->stream;
    ctx->queries.reserve((size_t)n * sizeof(float4), st);
    ctx->nn_prev.reserve((size_t)n * sizeof(int), st);
    ctx->tmp[6].reserve(16 * sizeof(float) + 16, st);
    static const float eye12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    PLS_CUDA(cudaMemcpyAsync(ctx->tmp[6].p, eye12, sizeof(eye12), cudaMemcpyHostToDevice, st));
    uint32_t* nq = scalar_u32(ctx, SC_QUERY_COUNT);
    const uint32_t nq_host[3] = {(uint32_t)n, 0u, 0u};
    PLS_CUDA(cudaMemcpyAsync(nq, nq_host, sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    PLS_CUDA(cudaMemsetAsync(scalar_u32(ctx, SC_KD_LISTS), 0, 8 * sizeof(uint32_t), st));
    kd_rows_to_float4_kernel<<<grid_for(n, 256, 8 * kNumSMs), 256, 0, st>>>(d, n, ctx->queries.as<float4>());
    PLS_CHECK_LAUNCH();
    const KdIndex ix = make_index(ctx);
    launch_search(ctx, ix, ctx->queries.as<float4>(), nq, n, 0, 1, ctx->tmp[6].as<float>(), nullptr, ctx->nn_prev.as<int>(), true,
                  out_normals != nullptr, 0);
    kd_search_export_kernel<<<grid_for(n, 256, 8 * kNumSMs), 256, 0, st>>>(ix, ctx->nn_prev.as<int>(), n, (float*)onb.dev,
                                                                            (float*)onr.dev, (long long*)oix.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, onb);
    finish_out(ctx, onr);
    finish_out(ctx, oix);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

}  // extern "C"
