// Exact nearest-neighbour search over the kd local map: a linear BVH (Karras 2012) over the
// Morton-sorted map points, traversed with a per-thread stack.  Device side only.
//
// Index layout (all in HBM, L2-resident for the BASELINE map sizes):
//   sorted[M]   float4  map points in Morton order; .w = bit-cast insertion index
//   nodes[M-1]  64 B    internal node i covers sorted[first..last], split after `gamma`:
//                       left child = [first..gamma], right child = [gamma+1..last]; the node
//                       stores BOTH child boxes, so one 64-byte fetch decides both descents.
//                       Child ids are implicit (Karras): internal(left) = gamma,
//                       internal(right) = gamma + 1.  Ranges of <= LEAF points are scanned
//                       linearly (contiguous float4 loads) instead of being descended.
//   normals[M]  float4  lazily computed unit normal of each map point, .w = 1 once valid
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace pls {

#ifndef PLS_KD_LEAF
#define PLS_KD_LEAF 8
#endif
constexpr int KD_LEAF = PLS_KD_LEAF;  // treelet size: subtrees of <= KD_LEAF points are scanned linearly
constexpr int KD_STACK = 96;
constexpr int KD_KMAX = 32;  // k + 1 <= 32

struct BvhNode {
    float lmin[3], lmax[3], rmin[3], rmax[3];
    int first, gamma, last, pad;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 bytes");

__device__ __forceinline__ float dist2_point(float x, float y, float z, const float4& p) {
    float dx = x - p.x, dy = y - p.y, dz = z - p.z;
    return dx * dx + dy * dy + dz * dz;
}

// Multi-level cell tables over the SAME Morton-sorted array: at level l a cell is the key prefix
// key >> 3 (b0 + l) (cubic cell of side cell0 * 2^l metres) and owns a contiguous range of `sorted`.
// Open-addressing hash tables (linear probing) map cell id -> [start, end].  A query probes the 27
// cells around it (independent loads) and scans their ranges (contiguous float4 reads): memory-level
// parallelism instead of the BVH's dependent node chain.  The result is exact whenever the k-th best
// squared distance is below (cell - margin)^2, because every point that close lies inside the 27-block;
// otherwise the next coarser level is tried and finally the BVH.
constexpr int KD_LEVELS = 1;
constexpr int KD_COORD_BITS = 13;                      // Morton bits per axis
constexpr int KD_COORD_MAX = (1 << KD_COORD_BITS) - 1;
constexpr float KD_CELL_TARGET = 0.16f;   // default level-0 cell side in [0.16, 0.32) m (PLS_KD_CELL overrides)
constexpr float KD_CELL_MARGIN = 2e-3f;   // quantisation slack, metres

struct KdGridHeader {
    float mn[3];
    float scale;      // Morton units per metre
    int b0;           // bits dropped per axis at level 0
    float cell0;      // level-0 cell side, metres
    int overflow[KD_LEVELS];
};

struct KdIndex {
    const float4* sorted;
    const float4* nodes;  // 4 float4 per node
    float4* normals;
    int M;
    const KdGridHeader* grid;
    const uint4* table[KD_LEVELS];  // {id+1 lo, id+1 hi, start, end}
    uint32_t mask[KD_LEVELS];
    unsigned long long* stats;      // optional debug counters (PLS_KD_STATS=1), else null
};
// stats slots: 0 nn queries, 1 nn exact@L0, 2 nn exact@L1, 3 nn exact@L2, 4 nn bvh, 5 nn candidates,
//              6 knn queries, 7 knn exact@L0, 8 @L1, 9 @L2, 10 knn bvh, 11 knn candidates
__device__ __forceinline__ void kd_stat(const KdIndex& ix, int slot, unsigned long long v = 1ull) {
    if (ix.stats) atomicAdd(ix.stats + slot, v);
}

__device__ __forceinline__ uint64_t kd_spread3(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}
__device__ __forceinline__ uint32_t kd_hash(uint64_t id) {
    uint64_t h = id * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(h >> 32);
}

// Looks up cell `id` at one level; returns false if the cell is empty.
__device__ __forceinline__ bool kd_cell_lookup(const uint4* __restrict__ table, uint32_t mask, uint64_t id, int& start,
                                               int& end) {
    const uint32_t lo = (uint32_t)(id + 1), hi = (uint32_t)((id + 1) >> 32);
    uint32_t h = kd_hash(id) & mask;
    for (int probe = 0; probe < 64; ++probe) {
        const uint4 e = __ldg(table + h);
        if (e.x == lo && e.y == hi) {
            start = (int)e.z;
            end = (int)e.w;
            return true;
        }
        if (e.x == 0u && e.y == 0u) return false;
        h = (h + 1) & mask;
    }
    return false;
}

// Grid search of one level.  Calls visit(i, d2) for every point of the 27-block; returns the squared
// exactness radius (cell - margin)^2 of that level, or -1 if the level is unusable.
// Phase 1 probes the 27 cells (nine independent table loads per z-slab) and records the non-empty ranges;
// phase 2 walks the lane's ranges in ONE flattened loop, so a warp runs max_lane(sum of counts)
// iterations instead of sum_cells(max_lane(count)) -- the nested form diverged ~7x.
template <typename Visit>
__device__ __forceinline__ float kd_grid_scan(const KdIndex& ix, int level, float x, float y, float z, Visit visit) {
    const KdGridHeader* g = ix.grid;
    if (g->overflow[level]) return -1.f;
    const int b = g->b0 + level;
    const float fx = (x - g->mn[0]) * g->scale, fy = (y - g->mn[1]) * g->scale, fz = (z - g->mn[2]) * g->scale;
    // the same truncating quantisation as kd_morton_kernel for in-range points; floor for the rest
    const int cx = ((int)floorf(fx)) >> b, cy = ((int)floorf(fy)) >> b, cz = ((int)floorf(fz)) >> b;
    const int cmax = KD_COORD_MAX >> b;
    const uint4* __restrict__ table = ix.table[level];
    const uint32_t mask = ix.mask[level];
    uint64_t sx[3];
    bool okx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int xx = cx + d - 1;
        okx[d] = xx >= 0 && xx <= cmax;
        sx[d] = okx[d] ? kd_spread3((uint64_t)xx) : 0ull;
    }
    int rs[27], re[27];
    int nr = 0;
#pragma unroll 1
    for (int dz = -1; dz <= 1; ++dz) {
        const int zz = cz + dz;
        if (zz < 0 || zz > cmax) continue;
        const uint64_t kz = kd_spread3((uint64_t)zz) << 2;
        uint64_t id[9];
        uint32_t hh[9];
        uint4 ent[9];
        bool ok[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const int yy = cy + j / 3 - 1;
            ok[j] = okx[j % 3] && yy >= 0 && yy <= cmax;
            id[j] = kz | (kd_spread3((uint64_t)(ok[j] ? yy : 0)) << 1) | sx[j % 3];
            hh[j] = kd_hash(id[j]) & mask;
        }
#pragma unroll
        for (int j = 0; j < 9; ++j) ent[j] = ok[j] ? __ldg(table + hh[j]) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            if (!ok[j]) continue;
            const uint32_t lo = (uint32_t)(id[j] + 1), hi = (uint32_t)((id[j] + 1) >> 32);
            uint4 e = ent[j];
            bool hit = e.x == lo && e.y == hi;
            if (!hit && (e.x | e.y) != 0u) {  // collision: keep probing
                uint32_t h = hh[j];
                for (int probe = 0; probe < 64 && !hit; ++probe) {
                    h = (h + 1) & mask;
                    e = __ldg(table + h);
                    hit = e.x == lo && e.y == hi;
                    if ((e.x | e.y) == 0u) break;
                }
            }
            if (hit) {
                rs[nr] = (int)e.z;
                re[nr] = (int)e.w;
                ++nr;
            }
        }
    }
    // one candidate per iteration and a single visit site: the lanes of a warp stay converged while they have
    // candidates left (the earlier "advance range / continue" form let them drift out of phase: ncu showed
    // ~4 active lanes on the visit body)
    int total = 0;
    for (int r = 0; r < nr; ++r) total += re[r] - rs[r] + 1;
    int j = 0, i = 0, end = -1;
    for (int t = 0; t < total; ++t) {
        if (i > end) {
            i = rs[j];
            end = re[j];
            ++j;
        }
        visit(i, dist2_point(x, y, z, __ldg(ix.sorted + i)));
        ++i;
    }
    const float cell = g->cell0 * (float)(1 << level) - KD_CELL_MARGIN;
    return cell > 0.f ? cell * cell : -1.f;
}

__device__ __forceinline__ float dist2_box(float x, float y, float z, float mnx, float mny, float mnz, float mxx,
                                           float mxy, float mxz) {
    float dx = fmaxf(fmaxf(mnx - x, x - mxx), 0.f);
    float dy = fmaxf(fmaxf(mny - y, y - mxy), 0.f);
    float dz = fmaxf(fmaxf(mnz - z, z - mxz), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

// Exact 1-NN.  `hint` (a sorted position or -1) only seeds the pruning bound.
__device__ __forceinline__ int kd_nearest(const KdIndex& ix, float x, float y, float z, int hint, float* best_out) {
    float best = FLT_MAX;
    int best_i = -1;
    if (hint >= 0 && hint < ix.M) {
        best = dist2_point(x, y, z, __ldg(ix.sorted + hint));
        best_i = hint;
    }
    if (ix.M <= KD_LEAF) {
        for (int i = 0; i < ix.M; ++i) {
            float d = dist2_point(x, y, z, __ldg(ix.sorted + i));
            if (d < best) { best = d; best_i = i; }
        }
        if (best_out) *best_out = best;
        return best_i;
    }
    int stack_n[KD_STACK];
    float stack_d[KD_STACK];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4 a = __ldg(ix.nodes + 4 * (size_t)node);
        const float4 b = __ldg(ix.nodes + 4 * (size_t)node + 1);
        const float4 c = __ldg(ix.nodes + 4 * (size_t)node + 2);
        const float4 dd = __ldg(ix.nodes + 4 * (size_t)node + 3);
        const int first = __float_as_int(dd.x), gamma = __float_as_int(dd.y), last = __float_as_int(dd.z);
        float dl = dist2_box(x, y, z, a.x, a.y, a.z, a.w, b.x, b.y);
        float dr = dist2_box(x, y, z, b.z, b.w, c.x, c.y, c.z, c.w);
        const bool lleaf = (gamma - first + 1) <= KD_LEAF;
        const bool rleaf = (last - gamma) <= KD_LEAF;
        if (lleaf && dl < best) {
            for (int i = first; i <= gamma; ++i) {
                float d = dist2_point(x, y, z, __ldg(ix.sorted + i));
                if (d < best) { best = d; best_i = i; }
            }
        }
        if (rleaf && dr < best) {
            for (int i = gamma + 1; i <= last; ++i) {
                float d = dist2_point(x, y, z, __ldg(ix.sorted + i));
                if (d < best) { best = d; best_i = i; }
            }
        }
        const bool cl = !lleaf && dl < best;
        const bool cr = !rleaf && dr < best;
        if (cl && cr) {
            if (dl <= dr) {
                if (sp < KD_STACK) { stack_n[sp] = gamma + 1; stack_d[sp] = dr; ++sp; }
                node = gamma;
            } else {
                if (sp < KD_STACK) { stack_n[sp] = gamma; stack_d[sp] = dl; ++sp; }
                node = gamma + 1;
            }
            continue;
        }
        if (cl) { node = gamma; continue; }
        if (cr) { node = gamma + 1; continue; }
        // pop
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stack_d[sp] < best) { node = stack_n[sp]; found = true; break; }
        }
        if (!found) break;
    }
    if (best_out) *best_out = best;
    return best_i;
}

// Exact 1-NN, grid first: levels 0..KD_LEVELS-1, then the BVH seeded with the best candidate so far.
__device__ __forceinline__ int kd_nearest_fast(const KdIndex& ix, float x, float y, float z, int hint, int* used_bvh) {
    float best = FLT_MAX;
    int best_i = -1;
    if (hint >= 0 && hint < ix.M) {
        best = dist2_point(x, y, z, __ldg(ix.sorted + hint));
        best_i = hint;
    }
    kd_stat(ix, 0);
    if (ix.M > KD_LEAF) {
        for (int level = 0; level < KD_LEVELS; ++level) {
            int cand = 0;
            const float r2 = kd_grid_scan(ix, level, x, y, z, [&](int i, float d) {
                ++cand;
                if (d < best) { best = d; best_i = i; }
            });
            kd_stat(ix, 5, cand);
            if (r2 > 0.f && best <= r2) { kd_stat(ix, 1 + level); return best_i; }
        }
    }
    kd_stat(ix, 4);
    if (used_bvh) *used_bvh = 1;
    return kd_nearest(ix, x, y, z, best_i, nullptr);
}

// Sorted insertion into an ascending (d, i) list of capacity k.
__device__ __forceinline__ void knn_insert(float* d, int* idx, int k, int& count, float dn, int in) {
    if (count == k && dn >= d[k - 1]) return;
    int pos = count < k ? count : k - 1;
    while (pos > 0 && d[pos - 1] > dn) {
        d[pos] = d[pos - 1];
        idx[pos] = idx[pos - 1];
        --pos;
    }
    d[pos] = dn;
    idx[pos] = in;
    if (count < k) ++count;
}

// Exact k-NN (k <= KD_KMAX): fills d[]/idx[] ascending, returns the number found (min(k, M)).
// `count` entries of d[]/idx[] may already hold candidates (seeds): they bound the search from the start.
__device__ __forceinline__ int kd_knn(const KdIndex& ix, float x, float y, float z, int k, float* d, int* idx,
                                      int count = 0) {
    if (ix.M <= KD_LEAF) {
        for (int i = 0; i < ix.M; ++i) knn_insert(d, idx, k, count, dist2_point(x, y, z, __ldg(ix.sorted + i)), i);
        return count;
    }
    int stack_n[KD_STACK];
    float stack_d[KD_STACK];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4 a = __ldg(ix.nodes + 4 * (size_t)node);
        const float4 b = __ldg(ix.nodes + 4 * (size_t)node + 1);
        const float4 c = __ldg(ix.nodes + 4 * (size_t)node + 2);
        const float4 dd = __ldg(ix.nodes + 4 * (size_t)node + 3);
        const int first = __float_as_int(dd.x), gamma = __float_as_int(dd.y), last = __float_as_int(dd.z);
        float dl = dist2_box(x, y, z, a.x, a.y, a.z, a.w, b.x, b.y);
        float dr = dist2_box(x, y, z, b.z, b.w, c.x, c.y, c.z, c.w);
        const bool lleaf = (gamma - first + 1) <= KD_LEAF;
        const bool rleaf = (last - gamma) <= KD_LEAF;
        float worst = count < k ? FLT_MAX : d[k - 1];
        if (lleaf && dl < worst) {
            for (int i = first; i <= gamma; ++i)
                knn_insert(d, idx, k, count, dist2_point(x, y, z, __ldg(ix.sorted + i)), i);
            worst = count < k ? FLT_MAX : d[k - 1];
        }
        if (rleaf && dr < worst) {
            for (int i = gamma + 1; i <= last; ++i)
                knn_insert(d, idx, k, count, dist2_point(x, y, z, __ldg(ix.sorted + i)), i);
            worst = count < k ? FLT_MAX : d[k - 1];
        }
        const bool cl = !lleaf && dl < worst;
        const bool cr = !rleaf && dr < worst;
        if (cl && cr) {
            if (dl <= dr) {
                if (sp < KD_STACK) { stack_n[sp] = gamma + 1; stack_d[sp] = dr; ++sp; }
                node = gamma;
            } else {
                if (sp < KD_STACK) { stack_n[sp] = gamma; stack_d[sp] = dl; ++sp; }
                node = gamma + 1;
            }
            continue;
        }
        if (cl) { node = gamma; continue; }
        if (cr) { node = gamma + 1; continue; }
        bool found = false;
        while (sp > 0) {
            --sp;
            float w2 = count < k ? FLT_MAX : d[k - 1];
            if (stack_d[sp] < w2) { node = stack_n[sp]; found = true; break; }
        }
        if (!found) break;
    }
    return count;
}

// Register-resident ascending list of the K best (distance, index) pairs: fully unrolled, branch-free
// bubble insertion -- no local memory, no per-lane loops (the local-memory insertion sort ran with ~2.5
// active lanes per instruction and dominated the first version of the correspondence kernel).
template <int K>
struct KBest {
    float d[K];
    int i[K];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int j = 0; j < K; ++j) { d[j] = FLT_MAX; i[j] = -1; }
    }
    __device__ __forceinline__ bool full() const { return i[K - 1] >= 0; }
    __device__ __forceinline__ void bubble() {
#pragma unroll
        for (int j = K - 1; j > 0; --j) {
            const bool sw = d[j] < d[j - 1];
            const float td = sw ? d[j - 1] : d[j];
            const int ti = sw ? i[j - 1] : i[j];
            d[j - 1] = sw ? d[j] : d[j - 1];
            i[j - 1] = sw ? i[j] : i[j - 1];
            d[j] = td;
            i[j] = ti;
        }
    }
    // strict: a candidate equal to the current K-th distance does not displace it
    __device__ __forceinline__ void insert(float dn, int in) {
        if (dn < d[K - 1]) {
            d[K - 1] = dn;
            i[K - 1] = in;
            bubble();
        }
    }
    // the same result without a branch: for scans where nearly every candidate enters the list (a divergent
    // insert ran with ~4 of 32 lanes active and was 37 % of the normals kernel's instructions)
    __device__ __forceinline__ void insert_uniform(float dn, int in) {
        const bool take = dn < d[K - 1];
        d[K - 1] = take ? dn : d[K - 1];
        i[K - 1] = take ? in : i[K - 1];
        bubble();
    }
};

// Exact K-NN over the BVH into a fresh list, pruned from the start by `bound2` (an upper bound of the
// K-th squared distance, inclusive; FLT_MAX if none).  Every point is visited at most once.
template <int K>
__device__ __forceinline__ void kd_knn_bounded(const KdIndex& ix, float x, float y, float z, float bound2, KBest<K>& L) {
    L.reset();
    if (ix.M <= KD_LEAF) {
        for (int i = 0; i < ix.M; ++i) L.insert(dist2_point(x, y, z, __ldg(ix.sorted + i)), i);
        return;
    }
    // until the list is full, the bound (nudged up one ulp so that equality passes the strict tests) prunes
    const float open_bound = bound2 < FLT_MAX ? __int_as_float(__float_as_int(bound2) + 1) : FLT_MAX;
    int stack_n[KD_STACK];
    float stack_d[KD_STACK];
    int sp = 0;
    int node = 0;
    while (true) {
        const float4 a = __ldg(ix.nodes + 4 * (size_t)node);
        const float4 b = __ldg(ix.nodes + 4 * (size_t)node + 1);
        const float4 c = __ldg(ix.nodes + 4 * (size_t)node + 2);
        const float4 dd = __ldg(ix.nodes + 4 * (size_t)node + 3);
        const int first = __float_as_int(dd.x), gamma = __float_as_int(dd.y), last = __float_as_int(dd.z);
        const float dl = dist2_box(x, y, z, a.x, a.y, a.z, a.w, b.x, b.y);
        const float dr = dist2_box(x, y, z, b.z, b.w, c.x, c.y, c.z, c.w);
        const bool lleaf = (gamma - first + 1) <= KD_LEAF;
        const bool rleaf = (last - gamma) <= KD_LEAF;
        float worst = L.full() ? L.d[K - 1] : open_bound;
        if (lleaf && dl < worst) {
            for (int i = first; i <= gamma; ++i) {
                const float dp = dist2_point(x, y, z, __ldg(ix.sorted + i));
                if (dp < open_bound) L.insert(dp, i);
            }
            worst = L.full() ? L.d[K - 1] : open_bound;
        }
        if (rleaf && dr < worst) {
            for (int i = gamma + 1; i <= last; ++i) {
                const float dp = dist2_point(x, y, z, __ldg(ix.sorted + i));
                if (dp < open_bound) L.insert(dp, i);
            }
            worst = L.full() ? L.d[K - 1] : open_bound;
        }
        const bool cl = !lleaf && dl < worst;
        const bool cr = !rleaf && dr < worst;
        if (cl && cr) {
            if (dl <= dr) {
                if (sp < KD_STACK) { stack_n[sp] = gamma + 1; stack_d[sp] = dr; ++sp; }
                node = gamma;
            } else {
                if (sp < KD_STACK) { stack_n[sp] = gamma; stack_d[sp] = dl; ++sp; }
                node = gamma + 1;
            }
            continue;
        }
        if (cl) { node = gamma; continue; }
        if (cr) { node = gamma + 1; continue; }
        bool found = false;
        while (sp > 0) {
            --sp;
            const float w2 = L.full() ? L.d[K - 1] : open_bound;
            if (stack_d[sp] < w2) { node = stack_n[sp]; found = true; break; }
        }
        if (!found) break;
    }
}

// Eigenvector of the smallest eigenvalue of a symmetric 3x3 matrix (cyclic Jacobi, fp64).
__device__ __forceinline__ void smallest_eigenvector(const float* c /*xx,xy,xz,yy,yz,zz*/, float* n) {
    double A[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-18 * diag || off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0;
            const int q = pq == 0 ? 1 : 2;
            double apq = A[p][q];
            if (apq == 0.0) continue;
            double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double akp = A[k][p], akq = A[k][q];
                A[k][p] = cs * akp - sn * akq;
                A[k][q] = sn * akp + cs * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double apk = A[p][k], aqk = A[q][k];
                A[p][k] = cs * apk - sn * aqk;
                A[q][k] = sn * apk + cs * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = cs * vkp - sn * vkq;
                V[k][q] = sn * vkp + cs * vkq;
            }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    double nx = V[0][m], ny = V[1][m], nz = V[2][m];
    double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    n[0] = (float)(nx * inv);
    n[1] = (float)(ny * inv);
    n[2] = (float)(nz * inv);
}

// Second moments about the point itself of its k nearest OTHER map points (entry 0 of the (k+1)-NN list is
// the point itself), float32 sequential sums in ascending-distance order, then the smallest-eigenvalue
// direction (slam/odometry/local_map.py:397-422).
template <typename GetIdx>
__device__ __forceinline__ void kd_normal_from_neighbours(const KdIndex& ix, const float4& c, int k, int found,
                                                          GetIdx get_idx, float* n) {
    float sxx = 0.f, sxy = 0.f, sxz = 0.f, syy = 0.f, syz = 0.f, szz = 0.f;
    for (int j = 1; j < found; ++j) {
        const float4 q = __ldg(ix.sorted + get_idx(j));
        float dx = __fsub_rn(q.x, c.x), dy = __fsub_rn(q.y, c.y), dz = __fsub_rn(q.z, c.z);
        sxx = __fadd_rn(sxx, __fmul_rn(dx, dx));
        sxy = __fadd_rn(sxy, __fmul_rn(dx, dy));
        sxz = __fadd_rn(sxz, __fmul_rn(dx, dz));
        syy = __fadd_rn(syy, __fmul_rn(dy, dy));
        syz = __fadd_rn(syz, __fmul_rn(dy, dz));
        szz = __fadd_rn(szz, __fmul_rn(dz, dz));
    }
    const float kk = (float)k;
    float cov[6] = {__fdiv_rn(sxx, kk), __fdiv_rn(sxy, kk), __fdiv_rn(sxz, kk),
                    __fdiv_rn(syy, kk), __fdiv_rn(syz, kk), __fdiv_rn(szz, kk)};
    smallest_eigenvector(cov, n);
}

// Fast path for the default k = 10: grid block into a register list; if that is not provably exact, a fresh
// BVH pass bounded by the grid's K-th distance.
__device__ __forceinline__ void kd_point_normal_k10(const KdIndex& ix, int pos, float* n) {
    constexpr int K = 11;
    const float4 c = __ldg(ix.sorted + pos);
    KBest<K> L;
    L.reset();
    bool exact = false;
    kd_stat(ix, 6);
    if (ix.M > KD_LEAF) {
        const float r2 = kd_grid_scan(ix, 0, c.x, c.y, c.z, [&](int i, float dd) { L.insert_uniform(dd, i); });
        exact = r2 > 0.f && L.full() && L.d[K - 1] <= r2;
        if (exact) kd_stat(ix, 7);
    }
    if (!exact) {
        kd_stat(ix, 10);
        float bound = L.full() ? L.d[K - 1] : FLT_MAX;
        if (!L.full() && ix.M >= K) {
            // fewer than K points in the whole block (a sparse region): the farthest of K consecutive points
            // in Morton order bounds the K-NN radius, so the BVH walk below starts pruned
            const int lo = min(max(pos - K / 2, 0), ix.M - K);
            float far = 0.f;
            for (int i = lo; i < lo + K; ++i) far = fmaxf(far, dist2_point(c.x, c.y, c.z, __ldg(ix.sorted + i)));
            bound = far;
        }
        kd_knn_bounded<K>(ix, c.x, c.y, c.z, bound, L);
    }
    int found = 0;
#pragma unroll
    for (int j = 0; j < K; ++j) found += (L.i[j] >= 0) ? 1 : 0;
    // copy the indices out through a small switch-free accessor (keeps the list in registers)
    int idx[K];
#pragma unroll
    for (int j = 0; j < K; ++j) idx[j] = L.i[j];
    kd_normal_from_neighbours(ix, c, 10, found, [&](int j) { return idx[j]; }, n);
}

// Generic k (3..31): local-memory list over the BVH.
__device__ __forceinline__ void kd_point_normal(const KdIndex& ix, int pos, int k, float* n) {
    if (k == 10) {
        kd_point_normal_k10(ix, pos, n);
        return;
    }
    const float4 c = __ldg(ix.sorted + pos);
    float d[KD_KMAX];
    int idx[KD_KMAX];
    const int found = kd_knn(ix, c.x, c.y, c.z, k + 1, d, idx);
    kd_normal_from_neighbours(ix, c, k, found, [&](int j) { return idx[j]; }, n);
}

// Cached normal of map point `pos`; computes and publishes it on first use.  The 16-byte
// store carries normal and valid flag together, so a concurrent reader sees either the old
// (flag 0) or the complete new value; concurrent writers store identical bits.
__device__ __forceinline__ void kd_cached_normal(const KdIndex& ix, int pos, int k, float* n) {
    float4 v = __ldcg(ix.normals + pos);
    if (v.w == 0.f) {
        kd_point_normal(ix, pos, k, n);
        __stcg(ix.normals + pos, make_float4(n[0], n[1], n[2], 1.f));
    } else {
        n[0] = v.x; n[1] = v.y; n[2] = v.z;
    }
}

}  // namespace pls
