// Warp-cooperative exact nearest-neighbour search on the kd local map (one warp per query).
//
// The per-thread BVH walk is a chain of dependent L2/HBM loads executed with ~10 of 32 lanes active; with
// only ~1000 warps in flight it is latency-bound (ncu: 24 % of warp time parked at EXIT, long-scoreboard
// stalls on every distance evaluation).  Here the 32 lanes of a warp work on ONE query:
//   * the (2R+1)^3 level-0 cells around the query are probed 32 at a time (one independent table load per
//     lane), their point ranges are flattened with a warp prefix sum, and lanes read consecutive float4
//     points (coalesced);
//   * 1-NN: per-lane minimum + warp arg-min;  k-NN: per-lane register k-best lists (branch-free bubble
//     insert) merged by K rounds of warp arg-min;
//   * the search is exact once the k-th squared distance is below (R cell - margin)^2, because every point
//     that close lies inside the block; otherwise the next shell (R+1) is probed -- cheap exactly where it
//     is needed (sparse regions have few points per shell) -- up to KD_MAX_RING, then lane 0 finishes on
//     the BVH with the current bound.
#pragma once
#include "kdmap_device.cuh"

namespace pls {

constexpr int KD_MAX_RING = 4;
constexpr unsigned FULL = 0xffffffffu;

struct WarpGrid {
    float x, y, z;
    int cx, cy, cz, cmax;
    const uint4* table;
    uint32_t mask;
    float cell;
    bool usable;
};

__device__ __forceinline__ WarpGrid warp_grid_setup(const KdIndex& ix, float x, float y, float z) {
    WarpGrid g;
    const KdGridHeader* h = ix.grid;
    g.x = x; g.y = y; g.z = z;
    const int b = h->b0;
    g.cx = ((int)floorf((x - h->mn[0]) * h->scale)) >> b;
    g.cy = ((int)floorf((y - h->mn[1]) * h->scale)) >> b;
    g.cz = ((int)floorf((z - h->mn[2]) * h->scale)) >> b;
    g.cmax = 65535 >> b;
    g.table = ix.table[0];
    g.mask = ix.mask[0];
    g.cell = h->cell0;
    g.usable = h->overflow[0] == 0 && ix.M > KD_LEAF;
    return g;
}

// Visits (per lane) every map point in the cells at Chebyshev distance `ring` from the query's cell
// (ring == 0 together with 1 on the first call: `from_ring` = 0 scans the whole 3x3x3 block).
template <typename Visit>
__device__ __forceinline__ void warp_scan_ring(const KdIndex& ix, const WarpGrid& g, int from_ring, int ring, Visit visit) {
    const int lane = threadIdx.x & 31;
    const int side = 2 * ring + 1;
    const int ncell = side * side * side;
    for (int base = 0; base < ncell; base += 32) {
        const int c = base + lane;
        int start = 0, cnt = 0;
        if (c < ncell) {
            const int dz = c / (side * side) - ring;
            const int rem = c % (side * side);
            const int dy = rem / side - ring, dx = rem % side - ring;
            const int cheb = max(max(abs(dx), abs(dy)), abs(dz));
            const int xx = g.cx + dx, yy = g.cy + dy, zz = g.cz + dz;
            if (cheb >= from_ring && xx >= 0 && xx <= g.cmax && yy >= 0 && yy <= g.cmax && zz >= 0 && zz <= g.cmax) {
                const uint64_t id = kd_spread3((uint64_t)xx) | (kd_spread3((uint64_t)yy) << 1) | (kd_spread3((uint64_t)zz) << 2);
                int s, e;
                if (kd_cell_lookup(g.table, g.mask, id, s, e)) {
                    start = s;
                    cnt = e - s + 1;
                }
            }
        }
        // flatten the ranges of this batch: exclusive prefix of the counts over the lanes
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(FULL, incl, 31);
        const int excl = incl - cnt;
        for (int tb = 0; tb < total; tb += 32) {
            const int t = tb + lane;
            // largest j with excl_j <= t
            int lo = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const int cand = lo + step;
                const int v = __shfl_sync(FULL, excl, cand & 31);
                if (cand < 32 && v <= t) lo = cand;
            }
            const int sj = __shfl_sync(FULL, start, lo);
            const int ej = __shfl_sync(FULL, excl, lo);
            if (t < total) {
                const int i = sj + (t - ej);
                visit(i, dist2_point(g.x, g.y, g.z, __ldg(ix.sorted + i)));
            }
        }
    }
}

__device__ __forceinline__ void warp_argmin(float& d, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(FULL, d, o);
        const int oi = __shfl_xor_sync(FULL, i, o);
        if (od < d || (od == d && (unsigned)oi < (unsigned)i)) {
            d = od;
            i = oi;
        }
    }
}

// Exact 1-NN of (x,y,z); the result (sorted position) is returned in every lane.
__device__ __forceinline__ int warp_nearest(const KdIndex& ix, float x, float y, float z, int hint) {
    const int lane = threadIdx.x & 31;
    float best = FLT_MAX;
    int best_i = -1;
    if (hint >= 0 && hint < ix.M && lane == 0) {
        best = dist2_point(x, y, z, __ldg(ix.sorted + hint));
        best_i = hint;
    }
    kd_stat(ix, 0, lane == 0 ? 1ull : 0ull);
    const WarpGrid g = warp_grid_setup(ix, x, y, z);
    if (g.usable) {
        for (int ring = 1; ring <= KD_MAX_RING; ++ring) {
            warp_scan_ring(ix, g, ring == 1 ? 0 : ring, ring, [&](int i, float d) {
                if (d < best) { best = d; best_i = i; }
            });
            float gd = best;
            int gi = best_i;
            warp_argmin(gd, gi);
            const float r = ring * g.cell - KD_CELL_MARGIN;
            if (gi >= 0 && gd <= r * r) {
                kd_stat(ix, ring <= 3 ? ring : 3, lane == 0 ? 1ull : 0ull);
                return gi;
            }
        }
    }
    // last resort: lane 0 walks the BVH seeded with the best candidate so far
    float gd = best;
    int gi = best_i;
    warp_argmin(gd, gi);
    kd_stat(ix, 4, lane == 0 ? 1ull : 0ull);
    int res = 0;
    if (lane == 0) res = kd_nearest(ix, x, y, z, gi, nullptr);
    return __shfl_sync(FULL, res, 0);
}

// Exact (k+1 = 11)-NN of map point `pos`, ascending by distance, returned in every lane (idx[0] is the
// point itself or a duplicate at distance 0).
__device__ __forceinline__ int warp_knn11(const KdIndex& ix, int pos, int* idx_out /*[11]*/) {
    constexpr int K = 11;
    const int lane = threadIdx.x & 31;
    const float4 c = __ldg(ix.sorted + pos);
    KBest<K> L;
    L.reset();
    kd_stat(ix, 6, lane == 0 ? 1ull : 0ull);
    const WarpGrid g = warp_grid_setup(ix, c.x, c.y, c.z);
    float md[K];
    int mi[K];
    bool exact = false;
    if (g.usable) {
        for (int ring = 1; ring <= KD_MAX_RING && !exact; ++ring) {
            warp_scan_ring(ix, g, ring == 1 ? 0 : ring, ring, [&](int i, float d) { L.insert(d, i); });
            // merge the 32 per-lane lists: K rounds of warp arg-min over the heads of (copies of) the lists
            KBest<K> C = L;
#pragma unroll
            for (int r = 0; r < K; ++r) {
                float hd = C.d[0];
                int hi = C.i[0];
                float wd = hd;
                int wi = hi;
                warp_argmin(wd, wi);
                md[r] = wd;
                mi[r] = wi;
                if (hd == wd && hi == wi && wi >= 0) {  // the winning lane pops its head
#pragma unroll
                    for (int j = 0; j + 1 < K; ++j) { C.d[j] = C.d[j + 1]; C.i[j] = C.i[j + 1]; }
                    C.d[K - 1] = FLT_MAX;
                    C.i[K - 1] = -1;
                }
            }
            const float rad = ring * g.cell - KD_CELL_MARGIN;
            exact = mi[K - 1] >= 0 && md[K - 1] <= rad * rad;
            if (exact) kd_stat(ix, 6 + (ring <= 3 ? ring : 3), lane == 0 ? 1ull : 0ull);
        }
    }
    if (!exact) {
        kd_stat(ix, 10, lane == 0 ? 1ull : 0ull);
        const float bound = (g.usable && mi[K - 1] >= 0) ? md[K - 1] : FLT_MAX;
        KBest<K> B;
        B.reset();
        if (lane == 0) kd_knn_bounded<K>(ix, c.x, c.y, c.z, bound, B);
#pragma unroll
        for (int r = 0; r < K; ++r) {
            md[r] = __shfl_sync(FULL, B.d[r], 0);
            mi[r] = __shfl_sync(FULL, B.i[r], 0);
        }
    }
    int found = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        idx_out[r] = mi[r];
        found += mi[r] >= 0 ? 1 : 0;
    }
    return found;
}

// Unit normal of map point `pos` from its 10 nearest other points (local_map.py:397-422): lanes 1..10 load
// one neighbour each; the float32 second moments are summed in ascending-distance order (as numpy's mean
// over the neighbour axis does); every lane runs the (uniform) eigen-solve and returns the same normal.
__device__ __forceinline__ void warp_point_normal_k10(const KdIndex& ix, int pos, float* n) {
    const int lane = threadIdx.x & 31;
    int idx[11];
    const int found = warp_knn11(ix, pos, idx);
    const float4 c = __ldg(ix.sorted + pos);
    int my = -1;
#pragma unroll
    for (int j = 1; j < 11; ++j)
        if (lane == j) my = idx[j];
    float pxx = 0.f, pxy = 0.f, pxz = 0.f, pyy = 0.f, pyz = 0.f, pzz = 0.f;
    if (my >= 0) {
        const float4 q = __ldg(ix.sorted + my);
        const float dx = __fsub_rn(q.x, c.x), dy = __fsub_rn(q.y, c.y), dz = __fsub_rn(q.z, c.z);
        pxx = __fmul_rn(dx, dx); pxy = __fmul_rn(dx, dy); pxz = __fmul_rn(dx, dz);
        pyy = __fmul_rn(dy, dy); pyz = __fmul_rn(dy, dz); pzz = __fmul_rn(dz, dz);
    }
    float sxx = 0.f, sxy = 0.f, sxz = 0.f, syy = 0.f, syz = 0.f, szz = 0.f;
    for (int j = 1; j < 11; ++j) {
        if (j >= found) break;
        sxx = __fadd_rn(sxx, __shfl_sync(FULL, pxx, j));
        sxy = __fadd_rn(sxy, __shfl_sync(FULL, pxy, j));
        sxz = __fadd_rn(sxz, __shfl_sync(FULL, pxz, j));
        syy = __fadd_rn(syy, __shfl_sync(FULL, pyy, j));
        syz = __fadd_rn(syz, __shfl_sync(FULL, pyz, j));
        szz = __fadd_rn(szz, __shfl_sync(FULL, pzz, j));
    }
    float cov[6] = {__fdiv_rn(sxx, 10.f), __fdiv_rn(sxy, 10.f), __fdiv_rn(sxz, 10.f),
                    __fdiv_rn(syy, 10.f), __fdiv_rn(syz, 10.f), __fdiv_rn(szz, 10.f)};
    smallest_eigenvector(cov, n);
}

}  // namespace pls
