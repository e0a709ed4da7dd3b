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
// Search.  A warp owns a query.  Lanes 0..26 each probe one cell of the 3x3x3 block around the query (27 independent
// table loads = one L2 round trip); the candidate ranges are flattened with a warp scan and read 32 at a time, lane t
// taking candidate t -- consecutive lanes read consecutive float4 of a range, so a round is a handful of full 128-byte
// lines instead of 32 scattered sectors (the thread-per-cell scans of round 1 spent their time in L1 wavefronts).
// Every point closer than (cell side - margin) lies inside the block, so a best (or k-th best) distance below that is
// exact; otherwise the next coarser level is scanned (cells pruned by their box distance), up to the level that holds
// everything.  No tree, no stack, no divergent descent.
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

// The cell of the 3x3x3 block (at `level`, around the query) owned by this lane: its point range [start, start+count)
// or count = 0 (lanes >= 27, cells outside the grid, empty cells).  `box2` receives the squared distance from the
// query to the cell's box, shrunk by the quantisation slack (a lower bound for every point binned into it): the caller
// drops cells farther than its current best.  The probe itself does not depend on any bound, so it can be in flight
// together with the load of the previous match.
// Returns the squared exactness radius of the block -- cell side + the query's clearance inside its own cell, minus the
// quantisation slack -- (FLT_MAX at the top level: the block then holds every point;
// negative if the level is unusable).  A query outside the grid is clamped to the border cell: all points lie on one
// side of it along that axis, so the block still holds everything within one cell side.
__device__ __forceinline__ float warp_probe_block(const KdIndex& ix, const KdGridLocal& g, int level, float x, float y,
                                                  float z, int lane, int& start, int& count, float& box2) {
    start = 0;
    count = 0;
    box2 = FLT_MAX;
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
    if (lane < 27 && usable) {
        const int dz = lane / 9, rem = lane - dz * 9, dy = rem / 3, dx = rem - dy * 3;
        const int xx = cx + dx - 1, yy = cy + dy - 1, zz = cz + dz - 1;
        if (xx >= 0 && xx <= cmax && yy >= 0 && yy <= cmax && zz >= 0 && zz <= cmax) {
            const uint32_t id = kd_cell_id((uint32_t)xx, (uint32_t)yy, (uint32_t)zz);
            const uint4* __restrict__ table = ix.table[level];
            const uint32_t mask = ix.mask[level];
            uint32_t h = kd_hash(id) & mask;
            uint4 e = __ldg(table + h);
            // box distance while the probe is in flight
            const float lox = ((float)xx * side_u - fx), hix = (fx - (float)(xx + 1) * side_u);
            const float loy = ((float)yy * side_u - fy), hiy = (fy - (float)(yy + 1) * side_u);
            const float loz = ((float)zz * side_u - fz), hiz = (fz - (float)(zz + 1) * side_u);
            const float ax = fmaxf(fmaxf(lox, hix) * g.inv_scale - KD_CELL_MARGIN, 0.f);
            const float ay = fmaxf(fmaxf(loy, hiy) * g.inv_scale - KD_CELL_MARGIN, 0.f);
            const float az = fmaxf(fmaxf(loz, hiz) * g.inv_scale - KD_CELL_MARGIN, 0.f);
            box2 = ax * ax + ay * ay + az * az;
            for (int probe = 0; probe < 64; ++probe) {
                if (e.y != ix.gen) break;              // empty (or stale generation): the cell holds no point
                if (e.x == id) {
                    start = (int)e.z;
                    count = (int)e.w - (int)e.z + 1;
                    break;
                }
                h = (h + 1) & mask;
                e = __ldg(table + h);
            }
        }
    }
    if (level >= g.top) return FLT_MAX;
    if (!usable) return -1.f;  // this level's table was too small: nothing scanned, nothing proven
    // every point outside the block is at least one cell side PLUS the query's distance to the nearest face of its
    // own cell away (zero for a query clamped in from outside the grid)
    const float ox = fminf(fx - (float)cx * side_u, (float)(cx + 1) * side_u - fx);
    const float oy = fminf(fy - (float)cy * side_u, (float)(cy + 1) * side_u - fy);
    const float oz = fminf(fz - (float)cz * side_u, (float)(cz + 1) * side_u - fz);
    const float own = fmaxf(fminf(fminf(ox, oy), oz), 0.f);
    const float cell = (side_u + own) * g.inv_scale - KD_CELL_MARGIN;
    return cell > 0.f ? cell * cell : -1.f;
}


// arg-min of (d, i) over the warp with two REDUX instructions (d >= 0, so the float's bit pattern orders like its
// value; ties: smaller index).  The result lands in every lane; (FLT_MAX, -1) if no lane holds a candidate.
__device__ __forceinline__ void warp_argmin(float& d, int& i) {
    const unsigned bits = __float_as_uint(d);
    const unsigned m = __reduce_min_sync(FULL, bits);
    const unsigned wi = __reduce_min_sync(FULL, bits == m ? (unsigned)i : 0xffffffffu);
    d = __uint_as_float(m);
    i = (int)wi;
}

// Inclusive warp scan of the per-lane range sizes; returns the total.
__device__ __forceinline__ int warp_scan_counts(int count, int lane, int& incl) {
    incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    return __shfl_sync(FULL, incl, 31);
}

// Index (into `sorted`) of the t-th candidate of the concatenated ranges: the owner cell is the number of lanes whose
// inclusive prefix is <= t (the prefixes are non-decreasing), found by a 5-step search over lane registers.
__device__ __forceinline__ int warp_candidate(int incl, int adj /* = start - exclusive prefix */, int t) {
    int c = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const int v = __shfl_sync(FULL, incl, c + s - 1);
        if (v <= t) c += s;
    }
    return __shfl_sync(FULL, adj, c) + t;
}

// Exact 1-NN of (x, y, z) by the whole warp; every lane returns the same sorted position (-1 if the map is empty).
// `hint` (a sorted position or -1, warp-uniform) only seeds the pruning bound; its load overlaps the level-0 probes.
// *cand_out (optional) accumulates the number of candidates tested.
// *second_out receives a lower bound of the squared distance from the query to every map point OTHER than the winner
// (the runner-up inside the scanned block, the boxes of the cells that were pruned, the block's exactness radius):
// as long as the query moves by less than the gap between the two, the winner stays the nearest neighbour -- the next
// ICP iterations verify that instead of searching again (kd_nn_verify_kernel).
__device__ __forceinline__ int warp_nearest(const KdIndex& ix, const KdGridLocal& g, float x, float y, float z, int hint,
                                            int lane, int* cand_out, float* second_out) {
    float best = FLT_MAX, second = FLT_MAX;
    int best_i = -1;
    float4 hp = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool has_hint = hint >= 0 && hint < ix.M;
    if (has_hint) hp = __ldg(ix.sorted + hint);
    if (lane == 0) kd_stat(ix, 0);
    for (int level = 0; level <= g.top; ++level) {
        int start, count;
        float box2;
        const float r2 = warp_probe_block(ix, g, level, x, y, z, lane, start, count, box2);
        if (level == 0 && has_hint) {
            best = dist2_point(x, y, z, hp);
            best_i = hint;
        }
        float l2 = FLT_MAX;          // this lane's bound for points other than its own best
        if (box2 > best) {           // nothing in that cell can beat (or tie) the bound ...
            if (count > 0) l2 = box2;  // ... but its points may be the runner-up
            count = 0;
        }
        int incl;
        const int total = warp_scan_counts(count, lane, incl);
        const int adj = start - (incl - count);
        float ld = best;
        int li = best_i;
        for (int base = 0; base < total; base += 32) {
            const int t = base + lane;
            const bool active = t < total;
            const int idx = warp_candidate(incl, adj, active ? t : total - 1);
            if (active) {
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
        warp_argmin(wd, wi);
        if (li != wi) l2 = fminf(l2, ld);  // this lane's best lost: it is a runner-up candidate
        const float inner = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(l2)));  // runner-up inside the block
        second = r2 >= 0.f ? fminf(inner, r2) : inner;  // anything outside the block is at least r2 away
        best = wd;
        best_i = wi;
        if (cand_out) *cand_out += total;
        if (ix.stats && lane == 0) {
            kd_stat(ix, 3, (unsigned long long)total);
            if (level == 0) kd_stat(ix, (best_i >= 0 && best <= r2) ? 1 : 2);
        }
        if (best_i >= 0 && best <= r2) break;
    }
    if (second_out) *second_out = second;
    return best_i;
}

// One chunk of the K-NN selection: R fresh candidates per lane (t = chunk + s * 32 + lane) plus the lane's entry of the
// list kept so far compete; on return lane r < K holds the r-th smallest of them.
//   * every lane sorts its R + 1 entries ascending (a small compare-exchange network, no communication);
//   * K rounds: the warp's minimum over the lane heads (two REDUX), the owner pops its head (a register shift).
template <int R>
__device__ __forceinline__ void knn_select_chunk(const KdIndex& ix, float x, float y, float z, int K, int lane, int incl, int adj,
                                                 int total, int chunk, float& keep_d, int& keep_i) {
    float sd[R + 1];
    int si[R + 1];
#pragma unroll
    for (int s = 0; s < R; ++s) {
        const int t = chunk + s * 32 + lane;
        const bool active = t < total;
        const int idx = warp_candidate(incl, adj, active ? t : total - 1);
        sd[s] = active ? dist2_point(x, y, z, __ldg(ix.sorted + idx)) : FLT_MAX;
        si[s] = active ? idx : -1;
    }
    sd[R] = keep_d;
    si[R] = keep_i;
    // insertion network: after pass p the first p + 2 entries are ordered
#pragma unroll
    for (int p = 1; p <= R; ++p) {
#pragma unroll
        for (int q = p; q >= 1; --q) {
            const bool sw = sd[q] < sd[q - 1] || (sd[q] == sd[q - 1] && (unsigned)si[q] < (unsigned)si[q - 1]);
            const float td = sw ? sd[q - 1] : sd[q];
            const int ti = sw ? si[q - 1] : si[q];
            sd[q - 1] = sw ? sd[q] : sd[q - 1];
            si[q - 1] = sw ? si[q] : si[q - 1];
            sd[q] = td;
            si[q] = ti;
        }
    }
    float nd = FLT_MAX;
    int ni = -1;
    for (int r = 0; r < K; ++r) {
        float wd = sd[0];
        int wi = si[0];
        const int mine = wi;
        warp_argmin(wd, wi);
        if (wi < 0) break;  // nothing left anywhere
        if (mine == wi) {   // indices are unique: exactly one lane pops
#pragma unroll
            for (int s = 0; s < R; ++s) {
                sd[s] = sd[s + 1];
                si[s] = si[s + 1];
            }
            sd[R] = FLT_MAX;
            si[R] = -1;
        }
        if (lane == r) {
            nd = wd;
            ni = wi;
        }
    }
    keep_d = nd;
    keep_i = ni;
}

// Exact K-NN (K <= 32, warp-uniform) of (x, y, z): on return lane r < found holds the r-th nearest point
// (out_d, out_i), ascending by (distance, index); returns the number found (min(K, M)).
// The candidates of a block are taken in chunks of 32 * R, R = 2, 4 or 8 register slots per lane by block size
// (a 27-cell block of the BASELINE maps holds 30-80 points: one chunk of R = 2 or 4).
__device__ __forceinline__ int warp_knn(const KdIndex& ix, const KdGridLocal& g, float x, float y, float z, int K, int lane,
                                        float& out_d, int& out_i, int* cand_out) {
    float keep_d = FLT_MAX;   // lane r: r-th best so far (carried between chunks / the result)
    int keep_i = -1;
    int found = 0;
    if (lane == 0) kd_stat(ix, 4);
    float bound = FLT_MAX;    // K-th distance of the previous (finer) level: an upper bound for this one
    for (int level = 0; level <= g.top; ++level) {
        int start, count;
        float box2;
        const float r2 = warp_probe_block(ix, g, level, x, y, z, lane, start, count, box2);
        if (box2 > bound) count = 0;
        int incl;
        const int total = warp_scan_counts(count, lane, incl);
        const int adj = start - (incl - count);
        if (cand_out) *cand_out += total;
        if (ix.stats && lane == 0) kd_stat(ix, 7, (unsigned long long)total);
        keep_d = FLT_MAX;     // this level's block is a superset of the previous one: select afresh
        keep_i = -1;
        if (total <= 64) {
            knn_select_chunk<2>(ix, x, y, z, K, lane, incl, adj, total, 0, keep_d, keep_i);
        } else if (total <= 128) {
            knn_select_chunk<4>(ix, x, y, z, K, lane, incl, adj, total, 0, keep_d, keep_i);
        } else {
            for (int chunk = 0; chunk < total; chunk += 256)
                knn_select_chunk<8>(ix, x, y, z, K, lane, incl, adj, total, chunk, keep_d, keep_i);
        }
        found = __popc(__ballot_sync(FULL, keep_i >= 0));
        const float kth = __shfl_sync(FULL, keep_d, K - 1);
        const bool exact = (found == K && kth <= r2) || level >= g.top;
        if (ix.stats && lane == 0 && level == 0) kd_stat(ix, exact ? 5 : 6);
        if (exact) break;
        if (found == K) bound = kth;
    }
    out_d = keep_d;
    out_i = keep_i;
    return found;
}

// Second moments about map point c of its k nearest OTHER map points (entry 0 of the (k+1)-NN list is the point
// itself): float32 sums taken sequentially in ascending-distance order and divided by k, as numpy's
// `.mean(axis=1)` forms them (slam/odometry/local_map.py:411-413).  Lane j holds neighbour j (nb_i); every lane
// returns the same six moments.
__device__ __forceinline__ void warp_second_moments(const KdIndex& ix, const float4& c, int k, int found, int nb_i, int lane,
                                                    float* cov) {
    float dx = 0.f, dy = 0.f, dz = 0.f;
    if (lane >= 1 && lane < found) {
        const float4 q = __ldg(ix.sorted + nb_i);
        dx = __fsub_rn(q.x, c.x);
        dy = __fsub_rn(q.y, c.y);
        dz = __fsub_rn(q.z, c.z);
    }
    float sxx = 0.f, sxy = 0.f, sxz = 0.f, syy = 0.f, syz = 0.f, szz = 0.f;
    for (int j = 1; j < found; ++j) {
        const float bx = __shfl_sync(FULL, dx, j), by = __shfl_sync(FULL, dy, j), bz = __shfl_sync(FULL, dz, j);
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
