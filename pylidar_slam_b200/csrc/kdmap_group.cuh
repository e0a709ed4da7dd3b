// Sub-warp cooperative exact search on the kd local map: G lanes (2, 4 or 8) share ONE query.
//
// Why sub-warps.  A thread-per-query search is a long dependent chain (27 table probes, ~100 candidate
// distances, a register k-best insert per candidate) and the frame only has ~32 k queries: 1 000 warps on
// 148 SMs, ~7 resident warps per SM where 64 fit -- the kernel is latency-bound with 9x occupancy headroom.
// A full warp per query (tried, see git history) overshoots: 32x more warps than fit, several waves, and a
// 32-way merge.  G = 4 or 8 lanes per query shorten every chain G-fold while the grid still fits in one wave.
//
//   * the 27 level-0 cells around the query are dealt round-robin to the G lanes: each lane probes its
//     <= ceil(27/G) cells (independent table loads) and scans their point ranges (contiguous float4 loads);
//   * 1-NN: per-lane minimum + group arg-min (log2 G shuffle steps);
//   * k-NN: per-lane register k-best lists merged by K rounds of group arg-min over the list heads;
//   * exactness is the same argument as kd_grid_scan's: every point closer than (cell - margin) lies inside
//     the 27-block; otherwise lane 0 of the group finishes on the BVH with the bound found so far.
#pragma once
#include "kdmap_device.cuh"

namespace pls {

template <int G>
struct LaneGroup {
    unsigned mask;  // the lanes of this group inside the warp
    int sub;        // rank of this lane inside the group
    __device__ __forceinline__ LaneGroup() {
        const int lane = threadIdx.x & 31;
        sub = lane & (G - 1);
        if constexpr (G == 32) mask = 0xffffffffu;
        else mask = ((1u << G) - 1u) << (lane - sub);
    }
    template <typename T>
    __device__ __forceinline__ T bcast(T v, int src = 0) const { return __shfl_sync(mask, v, src, G); }
    // arg-min of (d, i) over the group (ties: smaller index); the result lands in every lane
    __device__ __forceinline__ void argmin(float& d, int& i) const {
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(mask, d, o, G);
            const int oi = __shfl_xor_sync(mask, i, o, G);
            if (od < d || (od == d && (unsigned)oi < (unsigned)i)) {
                d = od;
                i = oi;
            }
        }
    }
};

// Calls visit(i, d2) for every map point in THIS LANE's share of the 27 level-0 cells around (x, y, z).
// Returns the squared exactness radius of the block, or -1 if the grid is unusable.
//   which = SCAN_ALL     all 27 cells
//   which = SCAN_CENTRE  only the query's own cell (one lane works)
//   which = SCAN_RING    the 26 others, skipping every cell whose nearest face is farther than sqrt(prune2)
//                        (a lower bound of the distance to any point binned into it, minus the quantisation
//                        slack) -- with a good bound from the centre cell or the previous iteration's match,
//                        a query probes 1-3 cells instead of 27
enum { SCAN_ALL = 0, SCAN_CENTRE = 1, SCAN_RING = 2 };
template <int G, typename Visit>
__device__ __forceinline__ float group_scan_block(const KdIndex& ix, int sub, float x, float y, float z, int which,
                                                  float prune2, Visit visit) {
    const KdGridHeader* g = ix.grid;
    if (g->overflow[0] || ix.M <= KD_LEAF) return -1.f;
    const int b = g->b0;
    const float fx = (x - g->mn[0]) * g->scale, fy = (y - g->mn[1]) * g->scale, fz = (z - g->mn[2]) * g->scale;
    const int cx = ((int)floorf(fx)) >> b, cy = ((int)floorf(fy)) >> b, cz = ((int)floorf(fz)) >> b;
    const int cmax = KD_COORD_MAX >> b;
    // metres from the query to the lower / upper face of its own cell, per axis, shrunk by the slack
    const float inv = 1.f / g->scale, side = (float)(1 << b);
    const float lox = fmaxf((fx - (float)cx * side) * inv - KD_CELL_MARGIN, 0.f);
    const float loy = fmaxf((fy - (float)cy * side) * inv - KD_CELL_MARGIN, 0.f);
    const float loz = fmaxf((fz - (float)cz * side) * inv - KD_CELL_MARGIN, 0.f);
    const float hix = fmaxf(((float)(cx + 1) * side - fx) * inv - KD_CELL_MARGIN, 0.f);
    const float hiy = fmaxf(((float)(cy + 1) * side - fy) * inv - KD_CELL_MARGIN, 0.f);
    const float hiz = fmaxf(((float)(cz + 1) * side - fz) * inv - KD_CELL_MARGIN, 0.f);
    const uint4* __restrict__ table = ix.table[0];
    const uint32_t mask = ix.mask[0];
    constexpr int NC = (27 + G - 1) / G;      // cells per lane
    constexpr int CH = NC > 7 ? 7 : NC;        // probed in batches of CH independent loads
    // Morton-spread coordinates of the three cell columns per axis (0 where out of range)
    uint64_t sx[3], sy[3], sz[3];
    bool vx[3], vy[3], vz[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int xx = cx + d - 1, yy = cy + d - 1, zz = cz + d - 1;
        vx[d] = xx >= 0 && xx <= cmax;
        vy[d] = yy >= 0 && yy <= cmax;
        vz[d] = zz >= 0 && zz <= cmax;
        sx[d] = vx[d] ? kd_spread3((uint64_t)xx) : 0ull;
        sy[d] = vy[d] ? kd_spread3((uint64_t)yy) << 1 : 0ull;
        sz[d] = vz[d] ? kd_spread3((uint64_t)zz) << 2 : 0ull;
    }
    int rs[NC], re[NC];
    int nr = 0, total = 0;
#pragma unroll 1
    for (int t0 = 0; t0 < NC; t0 += CH) {
        uint64_t id[CH];
        uint32_t hh[CH];
        uint4 ent[CH];
        bool ok[CH];
#pragma unroll
        for (int t = 0; t < CH; ++t) {
            const int c = sub + (t0 + t) * G;
            const int dz = c / 9, rem = c - dz * 9, dy = rem / 3, dx = rem - dy * 3;
            const bool okx = dx == 0 ? vx[0] : (dx == 1 ? vx[1] : vx[2]);
            const bool oky = dy == 0 ? vy[0] : (dy == 1 ? vy[1] : vy[2]);
            const bool okz = dz == 0 ? vz[0] : (dz == 1 ? vz[1] : vz[2]);
            const uint64_t kx = dx == 0 ? sx[0] : (dx == 1 ? sx[1] : sx[2]);
            const uint64_t ky = dy == 0 ? sy[0] : (dy == 1 ? sy[1] : sy[2]);
            const uint64_t kz = dz == 0 ? sz[0] : (dz == 1 ? sz[1] : sz[2]);
            ok[t] = (t0 + t) < NC && c < 27 && okx && oky && okz;
            if (which == SCAN_CENTRE) ok[t] = ok[t] && c == 13;
            if (which == SCAN_RING) {
                const float ax = dx == 0 ? lox : (dx == 1 ? 0.f : hix);
                const float ay = dy == 0 ? loy : (dy == 1 ? 0.f : hiy);
                const float az = dz == 0 ? loz : (dz == 1 ? 0.f : hiz);
                ok[t] = ok[t] && c != 13 && (ax * ax + ay * ay + az * az) < prune2;
            }
            id[t] = kx | ky | kz;
            hh[t] = kd_hash(id[t]) & mask;
        }
#pragma unroll
        for (int t = 0; t < CH; ++t) ent[t] = ok[t] ? __ldg(table + hh[t]) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int t = 0; t < CH; ++t) {
            if (!ok[t]) continue;
            const uint32_t lo = (uint32_t)(id[t] + 1), hi = (uint32_t)((id[t] + 1) >> 32);
            uint4 e = ent[t];
            bool hit = e.x == lo && e.y == hi;
            if (!hit && (e.x | e.y) != 0u) {  // collision: keep probing
                uint32_t h = hh[t];
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
                total += (int)e.w - (int)e.z + 1;
                ++nr;
            }
        }
    }
    // one candidate per iteration, a single visit site: lanes stay converged while they have candidates left
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
    const float cell = g->cell0 - KD_CELL_MARGIN;
    return cell > 0.f ? cell * cell : -1.f;
}

// Exact 1-NN of (x, y, z); every lane of the group returns the same sorted position.
template <int G>
__device__ __forceinline__ int group_nearest(const KdIndex& ix, const LaneGroup<G>& lg, float x, float y, float z, int hint) {
    float best = FLT_MAX;
    int best_i = -1;
    if (lg.sub == 0) {
        kd_stat(ix, 0);
        if (hint >= 0 && hint < ix.M) {
            best = dist2_point(x, y, z, __ldg(ix.sorted + hint));
            best_i = hint;
        }
    }
    int cand = 0;
    auto keep_min = [&](int i, float d) {
        ++cand;
        if (d < best) { best = d; best_i = i; }
    };
    // stage 1: the previous match (lane 0) and the query's own cell give a bound; stage 2: only the cells
    // of the block that can hold something closer
    group_scan_block<G>(ix, lg.sub, x, y, z, SCAN_CENTRE, FLT_MAX, keep_min);
    lg.argmin(best, best_i);
    const float r2 = group_scan_block<G>(ix, lg.sub, x, y, z, SCAN_RING, best, keep_min);
    if (ix.stats) kd_stat(ix, 5, cand);
    lg.argmin(best, best_i);
    if (r2 > 0.f && best_i >= 0 && best <= r2) {
        if (lg.sub == 0) kd_stat(ix, 1);
        return best_i;
    }
    int res = 0;
    if (lg.sub == 0) {
        kd_stat(ix, 4);
        res = kd_nearest(ix, x, y, z, best_i, nullptr);
    }
    return lg.bcast(res);
}

// Exact 11-NN (the point itself + its 10 nearest others) of map point `pos`, ascending by distance, in every
// lane of the group; returns the number found.
template <int G>
__device__ __forceinline__ int group_knn11(const KdIndex& ix, const LaneGroup<G>& lg, const float4& c, int pos, int* idx_out) {
    constexpr int K = 11;
    KBest<K> L;
    L.reset();
    if (lg.sub == 0) kd_stat(ix, 6);
    int cand = 0;
    const float r2 = group_scan_block<G>(ix, lg.sub, c.x, c.y, c.z, SCAN_ALL, FLT_MAX, [&](int i, float d) {
        ++cand;
        L.insert_uniform(d, i);
    });
    if (ix.stats) kd_stat(ix, 11, cand);
    // merge the G ascending per-lane lists: K rounds of group arg-min over the heads; the winner pops its head
    float md[K];
    int mi[K];
#pragma unroll
    for (int r = 0; r < K; ++r) {
        const float hd = L.d[0];
        const int hi = L.i[0];
        float wd = hd;
        int wi = hi;
        lg.argmin(wd, wi);
        md[r] = wd;
        mi[r] = wi;
        if (hi == wi && wi >= 0) {  // cells are disjoint across lanes: an index lives in exactly one list
#pragma unroll
            for (int j = 0; j + 1 < K; ++j) { L.d[j] = L.d[j + 1]; L.i[j] = L.i[j + 1]; }
            L.d[K - 1] = FLT_MAX;
            L.i[K - 1] = -1;
        }
    }
    const bool exact = r2 > 0.f && mi[K - 1] >= 0 && md[K - 1] <= r2;
    if (exact) {
        if (lg.sub == 0) kd_stat(ix, 7);
    } else {
        KBest<K> B;
        B.reset();
        if (lg.sub == 0) {
            kd_stat(ix, 10);
            float bound = mi[K - 1] >= 0 ? md[K - 1] : FLT_MAX;
            if (mi[K - 1] < 0 && ix.M >= K) {
                // fewer than K points in the whole block (a sparse region): the farthest of K consecutive
                // points in Morton order bounds the K-NN radius, so the BVH walk starts pruned
                const int lo = min(max(pos - K / 2, 0), ix.M - K);
                float far = 0.f;
                for (int i = lo; i < lo + K; ++i) far = fmaxf(far, dist2_point(c.x, c.y, c.z, __ldg(ix.sorted + i)));
                bound = far;
            }
            kd_knn_bounded<K>(ix, c.x, c.y, c.z, bound, B);
        }
#pragma unroll
        for (int r = 0; r < K; ++r) mi[r] = lg.bcast(B.i[r]);
    }
    int found = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        idx_out[r] = mi[r];
        found += mi[r] >= 0 ? 1 : 0;
    }
    return found;
}

// Unit normal of map point `pos` from its 10 nearest other points (local_map.py:397-422), in every lane.
template <int G>
__device__ __forceinline__ void group_point_normal_k10(const KdIndex& ix, const LaneGroup<G>& lg, int pos, float* n) {
    const float4 c = __ldg(ix.sorted + pos);
    int idx[11];
    const int found = group_knn11<G>(ix, lg, c, pos, idx);
    // every lane runs the same (uniform) accumulation and eigen-solve: the loads broadcast inside the warp
    kd_normal_from_neighbours(ix, c, 10, found, [&](int j) { return idx[j]; }, n);
}

}  // namespace pls
