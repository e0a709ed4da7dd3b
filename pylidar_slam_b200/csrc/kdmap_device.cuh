// Exact nearest-neighbour search over the kd local map (replaces the pykdtree queries of
// slam/odometry/local_map.py:385,405): a pyramid of hashed cell tables over ONE array of map points sorted by
// the Morton code of their level-0 cell, searched by WHOLE WARPS.  Device side only.
//
// Index layout (HBM; L2-resident at the BASELINE map sizes):
//   sorted[M]   float4  map points ordered by level-0 cell id (insertion order inside a cell); .w = bit-cast
//                       insertion index
//   normals[M]  float4  lazily computed unit normal of each map point; .w carries a state word:
//                       2 gen = claimed (queued for computation), 2 gen + 1 = valid, anything else = stale
//   table[l]    uint4   open-addressing hash table of level l: {cell id, gen, first, last}.  A level-l cell is the
//                       Morton prefix id0 >> 3 l (cube of side cell0 * 2^l) and owns a CONTIGUOUS range of `sorted`,
//                       so one sorted array serves every level.  Entries of older generations count as empty: no
//                       table is ever cleared.  Level `top` is a single cell holding the whole map.
//
// Search.  A group of 8 lanes owns a query (four searches per warp).  The lanes share the 27 probes of the 3x3x3 block
// around the query (independent table loads = one L2 round trip); the candidate ranges are flattened with a group scan,
// their indices staged in a small shared-memory list and read G at a time, lane t taking candidate t -- the lanes of a
// group read consecutive float4 of a cell as one segment instead of scattered sectors (the thread-per-cell scans of
// round 1 spent their time in L1 wavefronts).  Every point closer than (cell side - margin) lies inside the block, so a
// best (or k-th best) distance below that is exact; otherwise the next coarser level is scanned (cells pruned by their
// box distance), up to the level that holds everything.  No tree, no stack, no divergent descent.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "eigen_device.cuh"

namespace pls {

constexpr int KD_COORD_BITS = 13;                       // quantisation bits per axis
constexpr int KD_COORD_MAX = (1 << KD_COORD_BITS) - 1;
constexpr int KD_MIN_B0 = 3;                            // level-0 cell ids then fit 30 bits (4 radix passes)
constexpr int KD_MAX_LEVELS = KD_COORD_BITS - KD_MIN_B0 + 1;  // levels 0 .. top, top <= 10
constexpr float KD_CELL_TARGET = 0.20f;   // default level-0 cell side, metres (PLS_KD_CELL overrides)
constexpr float KD_CELL_MARGIN = 2e-3f;   // quantisation slack, metres
constexpr int KD_KMAX = 32;               // k + 1 <= 32
constexpr unsigned FULL = 0xffffffffu;

struct KdGridHeader {
    float mn[3];
    float scale;      // quantisation units per metre
    int b0;           // bits dropped per axis at level 0
    float cell0;      // level-0 cell side, metres
    int top;          // coarsest level = KD_COORD_BITS - b0 (one cell)
    int overflow[KD_MAX_LEVELS];
};

struct KdIndex {
    const float4* sorted;
    float4* normals;
    int M;
    uint32_t gen;                      // generation of this build (tables, normal states)
    const KdGridHeader* grid;
    const uint4* table[KD_MAX_LEVELS];
    uint32_t mask[KD_MAX_LEVELS];
    unsigned long long* stats;         // optional debug counters (PLS_KD_STATS=1), else null
};
// stats slots: 0 nn queries, 1 nn exact at level 0, 2 nn needing coarser levels, 3 nn candidates,
//              4 knn queries, 5 knn exact at level 0, 6 knn needing coarser levels, 7 knn candidates
__device__ __forceinline__ void kd_stat(const KdIndex& ix, int slot, unsigned long long v = 1ull) {
    if (ix.stats) atomicAdd(ix.stats + slot, v);
}

__device__ __forceinline__ uint32_t kd_spread10(uint32_t x) {  // 10 bits -> every third bit
    x &= 0x3ffu;
    x = (x | (x << 16)) & 0x030000ffu;
    x = (x | (x << 8)) & 0x0300f00fu;
    x = (x | (x << 4)) & 0x030c30c3u;
    x = (x | (x << 2)) & 0x09249249u;
    return x;
}
__device__ __forceinline__ uint32_t kd_cell_id(uint32_t cx, uint32_t cy, uint32_t cz) {
    return kd_spread10(cx) | (kd_spread10(cy) << 1) | (kd_spread10(cz) << 2);
}
__device__ __forceinline__ uint32_t kd_hash(uint32_t id) { return (id * 0x9E3779B1u) ^ (id >> 15); }

__device__ __forceinline__ float dist2_point(float x, float y, float z, const float4& p) {
    const float dx = x - p.x, dy = y - p.y, dz = z - p.z;
    return dx * dx + dy * dy + dz * dz;
}

// The grid header in registers (warp-uniform values).
struct KdGridLocal {
    float mnx, mny, mnz, scale, inv_scale;
    int b0, top;
};
__device__ __forceinline__ KdGridLocal kd_load_grid(const KdIndex& ix) {
    KdGridLocal g;
    g.mnx = __ldg(&ix.grid->mn[0]); g.mny = __ldg(&ix.grid->mn[1]); g.mnz = __ldg(&ix.grid->mn[2]);
    g.scale = __ldg(&ix.grid->scale);
    g.inv_scale = 1.f / g.scale;
    g.b0 = __ldg(&ix.grid->b0);
    g.top = __ldg(&ix.grid->top);
    return g;
}

// ---- lane groups ---------------------------------------------------------------------------------------------------
// G lanes (8 here) share ONE query; a warp runs 32 / G searches side by side.  Everything a search needs across lanes --
// probing the 27 cells, walking the candidates, the arg-min / selection rounds -- costs the same number of warp
// instructions whatever G is, so a narrow group divides the per-query instruction count by 32 / G, while the group is
// still wide enough to read a cell's points as one coalesced segment and to finish a selection in few rounds.
// (Round 1 used 4 lanes with a private k-best list per lane -- ~70 instructions per candidate; a full warp per query,
// tried first in round 2, replicates the bookkeeping 32 times for ~25 candidates.)
template <int G>
struct LaneGroup {
    int lane, sub;
    unsigned mask;
    __device__ __forceinline__ LaneGroup() {
        lane = threadIdx.x & 31;
        sub = lane & (G - 1);
        mask = G == 32 ? FULL : (((1u << G) - 1u) << (lane - sub));
    }
    template <typename T>
    __device__ __forceinline__ T bcast(T v, int src) const { return __shfl_sync(mask, v, src, G); }
    __device__ __forceinline__ unsigned min_u32(unsigned v) const { return __reduce_min_sync(mask, v); }
    // arg-min of (d, i) over the group with two REDUX (d >= 0: the float's bit pattern orders like its value; ties:
    // smaller index); (FLT_MAX, -1) if no lane holds a candidate
    __device__ __forceinline__ void argmin(float& d, int& i) const {
        const unsigned bits = __float_as_uint(d);
        const unsigned m = min_u32(bits);
        const unsigned wi = min_u32(bits == m ? (unsigned)i : 0xffffffffu);
        d = __uint_as_float(m);
        i = (int)wi;
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
};

constexpr int KD_LIST = 256;  // candidate indices a group stages in shared memory per window

// The 27 cells of the 3x3x3 block (at `level`, around the query) are dealt to the G lanes, NC = ceil(27 / G) each
// (cell c = sub + k G): start[k], count[k] = the cell's point range (count 0: outside the grid / empty), box2[k] = the
// squared distance from the query to the cell's box, shrunk by the quantisation slack (a lower bound for every point
// binned into it): the caller drops cells farther than its current best.  The probes do not depend on any bound: all
// table loads of a lane are issued back to back.
// Returns the squared exactness radius of the block (FLT_MAX at the top level: the block then holds every point;
// negative if the level is unusable).  A query outside the grid is clamped to the border cell: all points lie on one
// side of it along that axis, so the block still holds everything within one cell side.
template <int G>
__device__ __forceinline__ float group_probe_block(const KdIndex& ix, const KdGridLocal& g, int level, float x, float y, float z,
                                                   int sub, int* start, int* count, float* box2) {
    constexpr int NC = (27 + G - 1) / G;
    const int b = g.b0 + level;
    const int cmax = KD_COORD_MAX >> b;
    const float side_u = (float)(1 << b);                      // cell side in quantisation units
    const float fx = fminf(fmaxf((x - g.mnx) * g.scale, -1.0e6f), 1.0e6f);
    const float fy = fminf(fmaxf((y - g.mny) * g.scale, -1.0e6f), 1.0e6f);
    const float fz = fminf(fmaxf((z - g.mnz) * g.scale, -1.0e6f), 1.0e6f);
    const int cx = min(max(((int)floorf(fx)) >> b, 0), cmax);
    const int cy = min(max(((int)floorf(fy)) >> b, 0), cmax);
    const int cz = min(max(((int)floorf(fz)) >> b, 0), cmax);
    const bool usable = level >= g.top || !__ldg(&ix.grid->overflow[level]);
    const uint4* __restrict__ table = ix.table[level];
    const uint32_t mask = ix.mask[level];
    uint32_t id[NC], hh[NC];
    uint4 ent[NC];
    bool ok[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int c = sub + k * G;
        const int dz = c / 9, rem = c - dz * 9, dy = rem / 3, dx = rem - dy * 3;
        const int xx = cx + dx - 1, yy = cy + dy - 1, zz = cz + dz - 1;
        ok[k] = usable && c < 27 && xx >= 0 && xx <= cmax && yy >= 0 && yy <= cmax && zz >= 0 && zz <= cmax;
        id[k] = kd_cell_id((uint32_t)xx, (uint32_t)yy, (uint32_t)zz);
        hh[k] = kd_hash(id[k]) & mask;
        // metres from the query to the cell's box
        const float lox = ((float)xx * side_u - fx), hix = (fx - (float)(xx + 1) * side_u);
        const float loy = ((float)yy * side_u - fy), hiy = (fy - (float)(yy + 1) * side_u);
        const float loz = ((float)zz * side_u - fz), hiz = (fz - (float)(zz + 1) * side_u);
        const float ax = fmaxf(fmaxf(lox, hix) * g.inv_scale - KD_CELL_MARGIN, 0.f);
        const float ay = fmaxf(fmaxf(loy, hiy) * g.inv_scale - KD_CELL_MARGIN, 0.f);
        const float az = fmaxf(fmaxf(loz, hiz) * g.inv_scale - KD_CELL_MARGIN, 0.f);
        box2[k] = ok[k] ? ax * ax + ay * ay + az * az : FLT_MAX;
    }
#pragma unroll
    for (int k = 0; k < NC; ++k) ent[k] = ok[k] ? __ldg(table + hh[k]) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        start[k] = 0;
        count[k] = 0;
        if (!ok[k]) continue;
        uint4 e = ent[k];
        uint32_t h = hh[k];
        for (int probe = 0; probe < 64; ++probe) {
            if (e.y != ix.gen) break;              // empty (or stale generation): the cell holds no point
            if (e.x == id[k]) {
                start[k] = (int)e.z;
                count[k] = (int)e.w - (int)e.z + 1;
                break;
            }
            h = (h + 1) & mask;
            e = __ldg(table + h);
        }
    }
    if (level >= g.top) return FLT_MAX;
    if (!usable) return -1.f;  // this level's table was too small: nothing scanned, nothing proven
    const float cell = side_u * g.inv_scale - KD_CELL_MARGIN;
    return cell > 0.f ? cell * cell : -1.f;
}

// The group's candidates as one flat sequence: lane `sub` owns positions [base, base + own) of it.  Returns the total.
template <int G>
__device__ __forceinline__ int group_flatten(const LaneGroup<G>& lg, const int* count, int& base) {
    constexpr int NC = (27 + G - 1) / G;
    int own = 0;
#pragma unroll
    for (int k = 0; k < NC; ++k) own += count[k];
    int incl = own;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const int v = __shfl_up_sync(lg.mask, incl, o, G);
        if (lg.sub >= o) incl += v;
    }
    base = incl - own;
    return lg.bcast(incl, G - 1);
}

// Stages the point indices of flat positions [w0, w0 + KD_LIST) in the group's shared-memory list: every lane writes
// the part of its own cells' ranges that falls into the window; afterwards lane t reads list[t], list[t + G], ... so
// that the lanes of a group load consecutive float4 of a cell as one segment.
template <int G>
__device__ __forceinline__ void group_stage(const LaneGroup<G>& lg, const int* start, const int* count, int base, int w0,
                                            int* __restrict__ list) {
    constexpr int NC = (27 + G - 1) / G;
    lg.sync();  // the previous window has been consumed
    int p = base;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const int lo = max(p, w0), hi = min(p + count[k], w0 + KD_LIST);
        for (int q = lo; q < hi; ++q) list[q - w0] = start[k] + (q - p);
        p += count[k];
    }
    lg.sync();
}

// Exact 1-NN of (x, y, z) by a lane group; every lane of the group returns the same sorted position (-1 if the map is
// empty).  `hint` (a sorted position or -1, group-uniform) only seeds the pruning bound; its load overlaps the level-0
// probes.  *cand_out (optional) accumulates the number of candidates tested.
// *second_out receives a lower bound of the squared distance from the query to every map point OTHER than the winner
// (the runner-up inside the scanned block, the boxes of the cells that were pruned, the block's exactness radius):
// as long as the query moves by less than the gap between the two, the winner stays the nearest neighbour -- the next
// ICP iterations verify that instead of searching again (kd_nn_verify_kernel).
template <int G>
__device__ __forceinline__ int group_nearest(const KdIndex& ix, const KdGridLocal& g, const LaneGroup<G>& lg, float x, float y,
                                             float z, int hint, int* __restrict__ list, int* cand_out, float* second_out) {
    constexpr int NC = (27 + G - 1) / G;
    float best = FLT_MAX, second = FLT_MAX;
    int best_i = -1;
    float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_hint = hint >= 0 && hint < ix.M;
    if (has_hint) hp = __ldg(ix.sorted + hint);
    if (lg.sub == 0) kd_stat(ix, 0);
    for (int level = 0; level <= g.top; ++level) {
        int start[NC], count[NC];
        float box2[NC];
        const float r2 = group_probe_block<G>(ix, g, level, x, y, z, lg.sub, start, count, box2);
        if (level == 0 && has_hint) {
            best = dist2_point(x, y, z, hp);
            best_i = hint;
        }
        float l2 = FLT_MAX;          // this lane's bound for points other than its own best
#pragma unroll
        for (int k = 0; k < NC; ++k) {
            if (box2[k] > best) {    // nothing in that cell can beat (or tie) the bound ...
                if (count[k] > 0) l2 = fminf(l2, box2[k]);  // ... but its points may be the runner-up
                count[k] = 0;
            }
        }
        int base;
        const int total = group_flatten<G>(lg, count, base);
        float ld = best;
        int li = best_i;
        for (int w0 = 0; w0 < total; w0 += KD_LIST) {
            group_stage<G>(lg, start, count, base, w0, list);
            const int wn = min(total - w0, KD_LIST);
            for (int t = lg.sub; t < wn; t += G) {
                const int idx = list[t];
                const float d = dist2_point(x, y, z, __ldg(ix.sorted + idx));
                if (d < ld || (d == ld && (unsigned)idx < (unsigned)li)) {
                    l2 = fminf(l2, ld);
                    ld = d;
                    li = idx;
                } else if (idx != li) {
                    l2 = fminf(l2, d);
                }
            }
        }
        float wd = ld;
        int wi = li;
        lg.argmin(wd, wi);
        if (li != wi) l2 = fminf(l2, ld);  // this lane's best lost: it is a runner-up candidate
        second = __uint_as_float(lg.min_u32(__float_as_uint(l2)));
        if (r2 >= 0.f) second = fminf(second, r2);  // anything outside the block is at least that far
        best = wd;
        best_i = wi;
        if (cand_out) *cand_out += total;
        if (ix.stats && lg.sub == 0) {
            kd_stat(ix, 3, (unsigned long long)total);
            if (level == 0) kd_stat(ix, (best_i >= 0 && best <= r2) ? 1 : 2);
        }
        if (best_i >= 0 && best <= r2) break;
    }
    if (second_out) *second_out = second;
    return best_i;
}

// Exact K-NN (K <= 32, group-uniform) of (x, y, z) by a lane group.  The r-th nearest point (ascending by distance,
// then index) ends up in lane r % G, slot r / G of (out_d, out_i) (KD_KEEP slots per lane); returns the number found
// (min(K, M)).
//
// Selection.  The candidates of a block are taken in chunks of G * R (R register slots per lane, lane `sub` holding
// list positions sub, sub + G, ...); K rounds of {lane-local minimum over its slots, group arg-min (two REDUX), the
// owner retires that slot} extract the K smallest of the chunk plus the K kept so far (which re-enter through the
// lanes' keep slots).  A 27-cell block of the BASELINE maps holds 30-80 points: one chunk.
constexpr int KD_KNN_SLOTS = 8;
// KEEP = result slots per lane: ceil(Kmax / G) for the largest K the instantiation serves.
template <int G, int KEEP>
__device__ __forceinline__ int group_knn(const KdIndex& ix, const KdGridLocal& g, const LaneGroup<G>& lg, float x, float y, float z,
                                         int K, int* __restrict__ list, float* out_d, int* out_i, int* cand_out) {
    constexpr int NC = (27 + G - 1) / G;
    constexpr int R = KD_KNN_SLOTS;
    float keep_d[KEEP];
    int keep_i[KEEP];
    int found = 0;
    if (lg.sub == 0) kd_stat(ix, 4);
    float bound = FLT_MAX;    // K-th distance of the previous (finer) level: an upper bound for this one
    for (int level = 0; level <= g.top; ++level) {
        int start[NC], count[NC];
        float box2[NC];
        const float r2 = group_probe_block<G>(ix, g, level, x, y, z, lg.sub, start, count, box2);
#pragma unroll
        for (int k = 0; k < NC; ++k)
            if (box2[k] > bound) count[k] = 0;
        int base;
        const int total = group_flatten<G>(lg, count, base);
        if (cand_out) *cand_out += total;
        if (ix.stats && lg.sub == 0) kd_stat(ix, 7, (unsigned long long)total);
#pragma unroll
        for (int c = 0; c < KEEP; ++c) {  // this level's block is a superset of the previous one: select afresh
            keep_d[c] = FLT_MAX;
            keep_i[c] = -1;
        }
        for (int w0 = 0; w0 < total || w0 == 0; w0 += KD_LIST) {
            group_stage<G>(lg, start, count, base, w0, list);
            const int wn = min(total - w0, KD_LIST);
            for (int chunk = 0; chunk < wn || chunk == 0; chunk += G * R) {
                float sd[R + KEEP];
                int si[R + KEEP];
#pragma unroll
                for (int s = 0; s < R; ++s) {
                    const int t = chunk + s * G + lg.sub;
                    const bool active = t < wn;
                    const int idx = active ? list[t] : -1;
                    sd[s] = active ? dist2_point(x, y, z, __ldg(ix.sorted + idx)) : FLT_MAX;
                    si[s] = idx;
                }
#pragma unroll
                for (int c = 0; c < KEEP; ++c) {  // the K kept so far compete again
                    sd[R + c] = keep_d[c];
                    si[R + c] = keep_i[c];
                    keep_d[c] = FLT_MAX;
                    keep_i[c] = -1;
                }
                for (int r = 0; r < K; ++r) {
                    float ld = sd[0];
                    int li = si[0];
#pragma unroll
                    for (int s = 1; s < R + KEEP; ++s) {
                        const bool lt = sd[s] < ld || (sd[s] == ld && (unsigned)si[s] < (unsigned)li);
                        ld = lt ? sd[s] : ld;
                        li = lt ? si[s] : li;
                    }
                    float wd = ld;
                    int wi = li;
                    lg.argmin(wd, wi);
                    if (wi < 0) break;  // nothing left
                    if (li == wi) {     // indices are unique: exactly one lane retires a slot
#pragma unroll
                        for (int s = 0; s < R + KEEP; ++s) {
                            const bool hit = si[s] == wi;
                            sd[s] = hit ? FLT_MAX : sd[s];
                            si[s] = hit ? -1 : si[s];
                        }
                    }
                    if (lg.sub == (r & (G - 1))) {
#pragma unroll
                        for (int c = 0; c < KEEP; ++c)
                            if (c == r / G) {
                                keep_d[c] = wd;
                                keep_i[c] = wi;
                            }
                    }
                }
            }
        }
        // found = number of kept entries over the group; K-th distance = entry K - 1
        int mine = 0;
#pragma unroll
        for (int c = 0; c < KEEP; ++c) mine += keep_i[c] >= 0 ? 1 : 0;
        found = __reduce_add_sync(lg.mask, mine);
        float kth = FLT_MAX;
#pragma unroll
        for (int c = 0; c < KEEP; ++c)
            if (c == (K - 1) / G) kth = keep_d[c];
        kth = lg.bcast(kth, (K - 1) & (G - 1));
        const bool exact = (found == K && kth <= r2) || level >= g.top;
        if (ix.stats && lg.sub == 0 && level == 0) kd_stat(ix, exact ? 5 : 6);
        if (exact) break;
        if (found == K) bound = kth;
    }
#pragma unroll
    for (int c = 0; c < KEEP; ++c) {
        out_d[c] = keep_d[c];
        out_i[c] = keep_i[c];
    }
    return found;
}

// Second moments about map point c of its k nearest OTHER map points (entry 0 of the (k+1)-NN list is the point
// itself): float32 sums taken sequentially in ascending-distance order and divided by k, as numpy's
// `.mean(axis=1)` forms them (slam/odometry/local_map.py:411-413).  Neighbour j lives in lane j % G, slot j / G
// (group_knn's layout, KEEP slots per lane); every lane of the group returns the same six moments.
template <int G, int KEEP>
__device__ __forceinline__ void group_second_moments(const KdIndex& ix, const LaneGroup<G>& lg, const float4& c, int k, int found,
                                                     const int* nb_i, float* cov) {
    float dx[KEEP], dy[KEEP], dz[KEEP];
#pragma unroll
    for (int s = 0; s < KEEP; ++s) {
        dx[s] = dy[s] = dz[s] = 0.f;
        const int j = s * G + lg.sub;
        if (j >= 1 && j < found && nb_i[s] >= 0) {
            const float4 q = __ldg(ix.sorted + nb_i[s]);
            dx[s] = __fsub_rn(q.x, c.x);
            dy[s] = __fsub_rn(q.y, c.y);
            dz[s] = __fsub_rn(q.z, c.z);
        }
    }
    float sxx = 0.f, sxy = 0.f, sxz = 0.f, syy = 0.f, syz = 0.f, szz = 0.f;
    for (int j = 1; j < found; ++j) {
        float vx = 0.f, vy = 0.f, vz = 0.f;
#pragma unroll
        for (int s = 0; s < KEEP; ++s)
            if (s == j / G) { vx = dx[s]; vy = dy[s]; vz = dz[s]; }
        const int src = j & (G - 1);
        const float bx = lg.bcast(vx, src), by = lg.bcast(vy, src), bz = lg.bcast(vz, src);
        sxx = __fadd_rn(sxx, __fmul_rn(bx, bx));
        sxy = __fadd_rn(sxy, __fmul_rn(bx, by));
        sxz = __fadd_rn(sxz, __fmul_rn(bx, bz));
        syy = __fadd_rn(syy, __fmul_rn(by, by));
        syz = __fadd_rn(syz, __fmul_rn(by, bz));
        szz = __fadd_rn(szz, __fmul_rn(bz, bz));
    }
    const float kk = (float)k;
    cov[0] = __fdiv_rn(sxx, kk); cov[1] = __fdiv_rn(sxy, kk); cov[2] = __fdiv_rn(sxz, kk);
    cov[3] = __fdiv_rn(syy, kk); cov[4] = __fdiv_rn(syz, kk); cov[5] = __fdiv_rn(szz, kk);
}

// State words of the normal cache.
__device__ __forceinline__ uint32_t kd_normal_claimed(uint32_t gen) { return 2u * gen; }
__device__ __forceinline__ uint32_t kd_normal_valid(uint32_t gen) { return 2u * gen + 1u; }

}  // namespace pls
