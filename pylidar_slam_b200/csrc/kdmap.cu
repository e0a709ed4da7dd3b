// K5/K8 -- the kd local map on the GPU (replaces KdTreeLocalMap, slam/odometry/local_map.py:254-427).
//
// Per update (local_map.py:302-369): the whole map is moved by inverse(relative_pose) in float32,
// the new frame's points are appended, the oldest frame is dropped beyond local_map_size, the
// search index is rebuilt and the normal cache cleared -- exactly the reference's life cycle,
// with the pykdtree build replaced by an LBVH build:
//   kd_move_append_kernel   move + append + bounding box (block reduce, ordered-int atomics)
//   kd_morton_kernel        48-bit Morton keys (16 bits/axis, cubic cells)
//   radix sort (6 passes)   primitives.cu
//   kd_gather_kernel        Morton-ordered float4 copy of the points
//   kd_hierarchy_kernel     Karras 2012 radix-tree topology, one thread per internal node
//   kd_boxes_kernel         bottom-up child boxes, second arrival at a node does the work
// Search (local_map.py:372-422): kd_search_kernel (fine-grained API) and kd_icp_iter_kernel
// (one launch per ICP iteration: transform, exact 1-NN, lazy 10-NN normals, point-to-plane
// residual/Jacobian/weight and the block-reduced normal equations).
#include <stdlib.h>

#include "gn_device.cuh"
#include "internal.cuh"
#include "icp_device.cuh"
#include "kdmap_device.cuh"
#include "kdmap_group.cuh"
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

__device__ __forceinline__ uint64_t spread3(uint64_t x) {
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// Quantisation of the map: 256 Morton units per metre (3.9 mm) when the map extent allows it (<= 32 m),
// else the KD_COORD_BITS-bit range is stretched over the extent (13 bits per axis = 39-bit keys = five 8-bit
// sort passes; the unit stays far below the cell side, and points sharing a unit are ordered by index).  Level-0 cells are 2^b0 units with a side in
// [KD_CELL_TARGET, 2 KD_CELL_TARGET).
__global__ void kd_grid_header_kernel(int* __restrict__ bbox, KdGridHeader* __restrict__ hdr, float cell_target) {
    if (threadIdx.x != 0) return;
    const float mnx = ordered_to_float(bbox[0]), mny = ordered_to_float(bbox[1]), mnz = ordered_to_float(bbox[2]);
    const float ex = ordered_to_float(bbox[3]) - mnx, ey = ordered_to_float(bbox[4]) - mny,
                ez = ordered_to_float(bbox[5]) - mnz;
    const float ext = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-6f));
    const float scale = fminf(256.0f, (float)KD_COORD_MAX / ext);
    int b0 = 0;
    while (b0 < 12 && (float)(1 << b0) < cell_target * scale) ++b0;
    hdr->mn[0] = mnx; hdr->mn[1] = mny; hdr->mn[2] = mnz;
    hdr->scale = scale;
    hdr->b0 = b0;
    hdr->cell0 = (float)(1 << b0) / scale;
    for (int l = 0; l < KD_LEVELS; ++l) hdr->overflow[l] = 0;
    // leave the box empty for the next update (saves that update an init launch)
    bbox[0] = bbox[1] = bbox[2] = 0x7fffffff;
    bbox[3] = bbox[4] = bbox[5] = (int)0x80000000;
}

__global__ void kd_morton_kernel(const float4* __restrict__ pts, int64_t n, const KdGridHeader* __restrict__ hdr,
                                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const float mnx = hdr->mn[0], mny = hdr->mn[1], mnz = hdr->mn[2];
    const float scale = hdr->scale;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = pts[i];
        uint32_t qx = (uint32_t)fminf(fmaxf((p.x - mnx) * scale, 0.f), (float)KD_COORD_MAX);
        uint32_t qy = (uint32_t)fminf(fmaxf((p.y - mny) * scale, 0.f), (float)KD_COORD_MAX);
        uint32_t qz = (uint32_t)fminf(fmaxf((p.z - mnz) * scale, 0.f), (float)KD_COORD_MAX);
        keys[i] = spread3(qx) | (spread3(qy) << 1) | (spread3(qz) << 2);
        vals[i] = (uint32_t)i;
    }
}

// Cell tables of all levels in one pass over the sorted keys: thread i is the head of a cell run at
// level l if its prefix differs from key[i-1]'s, the tail if it differs from key[i+1]'s; heads store
// `start`, tails store `end` into the slot they find-or-claim (64-bit CAS on the id words).
struct CellTables {
    uint4* table[KD_LEVELS];
    uint32_t mask[KD_LEVELS];
};

__device__ __forceinline__ int cell_slot(uint4* table, uint32_t mask, uint64_t id) {
    uint32_t h = kd_hash(id) & mask;
    const unsigned long long want = id + 1;
    for (int probe = 0; probe < 64; ++probe) {
        unsigned long long* word = reinterpret_cast<unsigned long long*>(&table[h]);
        const unsigned long long old = atomicCAS(word, 0ull, want);
        if (old == 0ull || old == want) return (int)h;
        h = (h + 1) & mask;
    }
    return -1;
}

__global__ void kd_cells_kernel(const uint64_t* __restrict__ keys, int64_t n, CellTables T, KdGridHeader* hdr) {
    const int b0 = hdr->b0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        const uint64_t kp = i > 0 ? keys[i - 1] : ~0ull;
        const uint64_t kn = i + 1 < n ? keys[i + 1] : ~0ull;
#pragma unroll
        for (int l = 0; l < KD_LEVELS; ++l) {
            const int sh = 3 * (b0 + l);
            const uint64_t id = k >> sh;
            const bool head = (i == 0) || ((kp >> sh) != id);
            const bool tail = (i + 1 == n) || ((kn >> sh) != id);
            if (head || tail) {
                const int slot = cell_slot(T.table[l], T.mask[l], id);
                if (slot < 0) {
                    hdr->overflow[l] = 1;
                } else {
                    if (head) T.table[l][slot].z = (uint32_t)i;
                    if (tail) T.table[l][slot].w = (uint32_t)i;
                }
            }
        }
    }
}

__global__ void kd_gather_kernel(const float4* __restrict__ pts, const uint32_t* __restrict__ order, int64_t n,
                                 float4* __restrict__ sorted, uint32_t* __restrict__ inv_order) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t src = order[i];
        float4 p = pts[src];
        p.w = __uint_as_float(src);
        sorted[i] = p;
        inv_order[src] = (uint32_t)i;
    }
}

__device__ __forceinline__ int delta_fn(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

// Karras 2012: internal node i in [0, n-2]
__global__ void kd_hierarchy_kernel(const uint64_t* __restrict__ keys, int n, int4* __restrict__ ranges,
                                    int* __restrict__ parent /* [n-1 internal][n leaves] */) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n - 1; i += gridDim.x * blockDim.x) {
        const int d = (delta_fn(keys, n, i, i + 1) - delta_fn(keys, n, i, i - 1)) >= 0 ? 1 : -1;
        const int dmin = delta_fn(keys, n, i, i - d);
        int lmax = 2;
        while (delta_fn(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
        int l = 0;
        for (int t = lmax >> 1; t >= 1; t >>= 1)
            if (delta_fn(keys, n, i, i + (l + t) * d) > dmin) l += t;
        const int j = i + l * d;
        const int dnode = delta_fn(keys, n, i, j);
        int s = 0;
        int t = l;
        do {
            t = (t + 1) >> 1;
            if (delta_fn(keys, n, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        const int gamma = i + s * d + min(d, 0);
        const int first = min(i, j), last = max(i, j);
        ranges[i] = make_int4(first, gamma, last, 0);
        if (first == gamma) parent[(n - 1) + gamma] = i; else parent[gamma] = i;
        if (last == gamma + 1) parent[(n - 1) + gamma + 1] = i; else parent[gamma + 1] = i;
        if (i == 0) parent[0] = -1;
    }
}

// Box of a child of a "big" node: a treelet (<= KD_LEAF points, scanned directly) or a completed big node.
__device__ __forceinline__ void child_box(const float4* __restrict__ sorted, const float4* nodes, int lo, int hi,
                                          int internal_id, float* mn, float* mx) {
    if (hi - lo + 1 <= KD_LEAF) {
        float4 p = sorted[lo];
        mn[0] = mx[0] = p.x; mn[1] = mx[1] = p.y; mn[2] = mx[2] = p.z;
        for (int i = lo + 1; i <= hi; ++i) {
            p = sorted[i];
            mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
            mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
        }
    } else {
        const float4 a = __ldcg(nodes + 4 * (size_t)internal_id);
        const float4 b = __ldcg(nodes + 4 * (size_t)internal_id + 1);
        const float4 c = __ldcg(nodes + 4 * (size_t)internal_id + 2);
        mn[0] = fminf(a.x, b.z); mn[1] = fminf(a.y, b.w); mn[2] = fminf(a.z, c.x);
        mx[0] = fmaxf(a.w, c.y); mx[1] = fmaxf(b.x, c.z); mx[2] = fmaxf(b.y, c.w);
    }
}

// Bottom-up child boxes of the BIG nodes (> KD_LEAF points).  Subtrees of <= KD_LEAF points ("treelets")
// are never descended by the search (they are scanned linearly), so the pass starts at the treelet roots:
// thread t < n-1 is internal node t, thread t >= n-1 is point leaf t-(n-1); a thread whose node is small
// while its parent is big "arrives" at the parent; the second arrival at a node computes its child boxes
// and continues upward.  Chains are ~log2(KD_LEAF) levels shorter than a per-point pass.
__global__ void kd_boxes_kernel(const float4* __restrict__ sorted, int n, const int4* __restrict__ ranges,
                                const int* __restrict__ parent, int* __restrict__ visit, float4* nodes) {
    const int total = 2 * n - 1;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
        int cur;
        if (t < n - 1) {
            const int4 rg = ranges[t];
            if (rg.z - rg.x + 1 > KD_LEAF || t == 0) continue;  // big node (completed by its children) or root
            cur = parent[t];
        } else {
            cur = parent[t];  // point leaf
        }
        {
            const int4 pr = ranges[cur];
            if (pr.z - pr.x + 1 <= KD_LEAF) continue;  // inside a treelet: nothing to do
        }
        bool wrote = false;  // a thread arriving from a treelet has published nothing yet: no fence needed
        while (cur >= 0) {
            if (wrote) __threadfence();
            wrote = true;
            if (atomicAdd(&visit[cur], 1) == 0) break;  // first arrival: the sibling subtree is not done
            const int4 rg = ranges[cur];
            float lmn[3], lmx[3], rmn[3], rmx[3];
            child_box(sorted, nodes, rg.x, rg.y, rg.y, lmn, lmx);
            child_box(sorted, nodes, rg.y + 1, rg.z, rg.y + 1, rmn, rmx);
            float4* o = nodes + 4 * (size_t)cur;
            __stcg(o + 0, make_float4(lmn[0], lmn[1], lmn[2], lmx[0]));
            __stcg(o + 1, make_float4(lmx[1], lmx[2], rmn[0], rmn[1]));
            __stcg(o + 2, make_float4(rmn[2], rmx[0], rmx[1], rmx[2]));
            __stcg(o + 3, make_float4(__int_as_float(rg.x), __int_as_float(rg.y), __int_as_float(rg.z), 0.f));
            cur = parent[cur];
        }
    }
}

// Fine-grained search: queries [n,3] -> neighbour points, normals, insertion indices.
__global__ void __launch_bounds__(128)
kd_search_kernel(KdIndex ix, int k_normals, const float* __restrict__ queries, int64_t n, float* __restrict__ out_nb,
                 float* __restrict__ out_nrm, long long* __restrict__ out_idx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float x = queries[3 * i], y = queries[3 * i + 1], z = queries[3 * i + 2];
        int pos = kd_nearest_fast(ix, x, y, z, -1, nullptr);
        float4 q = __ldg(ix.sorted + pos);
        out_nb[3 * i] = q.x; out_nb[3 * i + 1] = q.y; out_nb[3 * i + 2] = q.z;
        if (out_idx) out_idx[i] = (long long)__float_as_uint(q.w);
        if (out_nrm) {
            float nn[3];
            kd_cached_normal(ix, pos, k_normals, nn);
            out_nrm[3 * i] = nn[0]; out_nrm[3 * i + 1] = nn[1]; out_nrm[3 * i + 2] = nn[2];
        }
    }
}

constexpr int KD_ITER_THREADS = 128;

// One ICP iteration on the kd map (icp_odometry.py:275-284 + alignment.py:91-127 at x0 = 0):
//   p = T p0; q = NN(p); n = normal(q); r = n.(p - q); J = [n, p x n]; w; reduce.
__global__ void __launch_bounds__(KD_ITER_THREADS)
kd_icp_iter_kernel(KdIndex ix, int k_normals, const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev,
                   int64_t q_begin, int64_t q_stride, const FrameResult* __restrict__ fr, int scheme, float sigma,
                   int* __restrict__ nn_prev, double* __restrict__ partials) {
    if (fr->done) return;
    __shared__ float sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = fr->T[threadIdx.x];
    __syncthreads();
    const int64_t nq = (int64_t)*nq_dev;
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    // q_begin/q_stride shard the queries across ranks (multi-GPU): rank r takes r, r+R, ...
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;; s += (int64_t)gridDim.x * blockDim.x) {
        const int64_t qi = q_begin + s * q_stride;
        if (qi >= nq) break;
        const float4 p0 = queries[qi];
        float p[3];
        p[0] = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
        p[1] = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
        p[2] = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
        const int pos = kd_nearest_fast(ix, p[0], p[1], p[2], nn_prev[qi], nullptr);
        nn_prev[qi] = pos;
        const float4 qq = __ldg(ix.sorted + pos);
        float q[3] = {qq.x, qq.y, qq.z};
        float nn[3];
        kd_cached_normal(ix, pos, k_normals, nn);
        float J[6];
        float r = p2plane_residual_jacobian_identity(p, q, nn, J);
        float w = ls_weight<float>(scheme, sigma, r, p, q);
        accumulate_normal_equations<float>(acc, J, w, r * w, r);
    }
    block_reduce_store<KD_ITER_THREADS>(acc, partials + (size_t)blockIdx.x * NACC);
}

__global__ void kd_export_kernel(const float4* __restrict__ pts, int64_t n, float* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float4 p = pts[i];
        out[3 * i] = p.x; out[3 * i + 1] = p.y; out[3 * i + 2] = p.z;
    }
}

constexpr int KD_GROUP_THREADS = 128;
constexpr int KD_NGROUP_LATER_DEFAULT = 0;  // 0 = same group width in every iteration; 32 = a warp per pending normal later

// The same ICP iteration as kd_icp_iter_kernel, split in three launches so that the two searches can run
// with G lanes per query (kdmap_group.cuh) at full residency and the reduction stays thread-per-query:
//   kd_nn_group_kernel      p = T p0, exact 1-NN -> nn_prev[qi]
//   kd_normals_group_kernel lazily cached 10-NN normal of every matched map point that has none yet
//   kd_residual_kernel      r, J = [n, p x n], robust weight, the 30 fp64 accumulators -> block partials
template <int G>
__global__ void __launch_bounds__(KD_GROUP_THREADS, 8)
kd_nn_group_kernel(KdIndex ix, const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev, int64_t q_begin,
                   int64_t q_stride, const FrameResult* __restrict__ fr, int* __restrict__ nn_prev, int first) {
    if (fr->done) return;
    __shared__ float sT[12];
    if (threadIdx.x < 12) sT[threadIdx.x] = fr->T[threadIdx.x];
    __syncthreads();
    const LaneGroup<G> lg;
    const int64_t nq = (int64_t)*nq_dev;
    const int64_t groups_total = (int64_t)gridDim.x * (KD_GROUP_THREADS / G);
    for (int64_t s = ((int64_t)blockIdx.x * KD_GROUP_THREADS + threadIdx.x) / G;; s += groups_total) {
        const int64_t qi = q_begin + s * q_stride;
        if (qi >= nq) break;
        const float4 p0 = queries[qi];
        const float px = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
        const float py = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
        const float pz = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
        const int hint = (lg.sub == 0 && !first) ? nn_prev[qi] : -1;
        const int pos = group_nearest<G>(ix, lg, px, py, pz, hint);
        if (lg.sub == 0) nn_prev[qi] = pos;
    }
}

template <int G>
__global__ void __launch_bounds__(KD_GROUP_THREADS)
kd_normals_group_kernel(KdIndex ix, int k_normals, const uint32_t* __restrict__ nq_dev, int64_t q_begin, int64_t q_stride,
                        const FrameResult* __restrict__ fr, const int* __restrict__ nn_prev) {
    if (fr->done) return;
    const LaneGroup<G> lg;
    const int64_t nq = (int64_t)*nq_dev;
    const int64_t groups_total = (int64_t)gridDim.x * (KD_GROUP_THREADS / G);
    for (int64_t s = ((int64_t)blockIdx.x * KD_GROUP_THREADS + threadIdx.x) / G;; s += groups_total) {
        const int64_t qi = q_begin + s * q_stride;
        if (qi >= nq) break;
        const int pos = nn_prev[qi];
        // one lane decides for the group (another group may publish the same normal concurrently)
        const float valid = lg.bcast(lg.sub == 0 ? __ldcg(ix.normals + pos).w : 0.f);
        if (valid != 0.f) continue;
        float nn[3];
        if (k_normals == 10) {
            group_point_normal_k10<G>(ix, lg, pos, nn);
        } else if (lg.sub == 0) {
            kd_point_normal(ix, pos, k_normals, nn);
        }
        if (lg.sub == 0) __stcg(ix.normals + pos, make_float4(nn[0], nn[1], nn[2], 1.f));
    }
}

constexpr int KD_RES_THREADS = 256;

__global__ void __launch_bounds__(KD_RES_THREADS)
kd_residual_kernel(KdIndex ix, const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev, int64_t q_begin,
                   int64_t q_stride, FrameResult* fr, int scheme, float sigma,
                   const int* __restrict__ nn_prev, double* __restrict__ partials, float fuse_threshold) {
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
        const int pos = nn_prev[qi];
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

template <int G>
int launch_group_iteration(pls_context* ctx, const KdIndex& ix, int64_t mine, const uint32_t* nq_dev, int rank, int num_ranks,
                           bool first, float fuse_threshold) {
    cudaStream_t st = ctx->stream;
    const int gblocks = grid_for(mine * G, KD_GROUP_THREADS, 16 * kNumSMs);
    int* nn_prev = ctx->nn_prev.as<int>();
    FrameResult* fr = frame_result_dev(ctx);
    kd_nn_group_kernel<G><<<gblocks, KD_GROUP_THREADS, 0, st>>>(ix, ctx->query_ptr, nq_dev, (int64_t)rank,
                                                                (int64_t)num_ranks, fr, nn_prev, first ? 1 : 0);
    PLS_CHECK_LAUNCH();
    // lanes per pending normal (default = the search's G; PLS_KD_NGROUP=2|4|8 decouples the two for A/B runs: later
    // iterations only compute a few hundred new normals, where a wider group shortens the straggler chains)
    static const int ngroup = getenv("PLS_KD_NGROUP") ? atoi(getenv("PLS_KD_NGROUP")) : G;
    // Iterations >= 2 of a frame only meet a few hundred new matches: the kernel is then a tail of a few groups
    // walking long dependent 10-NN chains while the grid has drained (profiles/r1_kd_traffic.json: 7-9 of 32 lanes
    // active, 33-83 us).  PLS_KD_NGROUP_LATER=32 hands every pending normal of those iterations a full warp (one of
    // the 27 cells per lane); the first iteration, where all ~32 k matches need a normal and throughput matters,
    // keeps the narrow groups.
    static const int later = getenv("PLS_KD_NGROUP_LATER") ? atoi(getenv("PLS_KD_NGROUP_LATER")) : KD_NGROUP_LATER_DEFAULT;
    if (!first && later == 32) {
        kd_normals_group_kernel<32><<<grid_for(mine * 32, KD_GROUP_THREADS, 16 * kNumSMs), KD_GROUP_THREADS, 0, st>>>(
            ix, ctx->cfg.num_neighbors_normals, nq_dev, (int64_t)rank, (int64_t)num_ranks, fr, nn_prev);
    } else if (ngroup == 8 && G != 8) {
        kd_normals_group_kernel<8><<<grid_for(mine * 8, KD_GROUP_THREADS, 16 * kNumSMs), KD_GROUP_THREADS, 0, st>>>(
            ix, ctx->cfg.num_neighbors_normals, nq_dev, (int64_t)rank, (int64_t)num_ranks, fr, nn_prev);
    } else if (ngroup == 2 && G != 2) {
        kd_normals_group_kernel<2><<<grid_for(mine * 2, KD_GROUP_THREADS, 16 * kNumSMs), KD_GROUP_THREADS, 0, st>>>(
            ix, ctx->cfg.num_neighbors_normals, nq_dev, (int64_t)rank, (int64_t)num_ranks, fr, nn_prev);
    } else if (ngroup == 4 && G != 4) {
        kd_normals_group_kernel<4><<<grid_for(mine * 4, KD_GROUP_THREADS, 16 * kNumSMs), KD_GROUP_THREADS, 0, st>>>(
            ix, ctx->cfg.num_neighbors_normals, nq_dev, (int64_t)rank, (int64_t)num_ranks, fr, nn_prev);
    } else {
        kd_normals_group_kernel<G><<<gblocks, KD_GROUP_THREADS, 0, st>>>(ix, ctx->cfg.num_neighbors_normals, nq_dev,
                                                                         (int64_t)rank, (int64_t)num_ranks, fr, nn_prev);
    }
    PLS_CHECK_LAUNCH();
    const int blocks = grid_for(mine, KD_RES_THREADS, 8 * kNumSMs);
    ctx->partials.reserve((size_t)blocks * NACC * sizeof(double), st);
    kd_residual_kernel<<<blocks, KD_RES_THREADS, 0, st>>>(ix, ctx->query_ptr, nq_dev, (int64_t)rank, (int64_t)num_ranks, fr,
                                                          ctx->cfg.scheme, ctx->cfg.sigma, nn_prev,
                                                          ctx->partials.as<double>(), fuse_threshold);
    PLS_CHECK_LAUNCH();
    return blocks;
}

KdIndex make_index(pls_context* ctx) {
    KdIndex ix;
    ix.sorted = ctx->kd.sorted.as<float4>();
    ix.nodes = ctx->kd.nodes.as<float4>();
    ix.normals = ctx->kd.normals.as<float4>();
    ix.M = (int)ctx->kd.indexed;
    ix.grid = ctx->kd.grid_hdr.as<KdGridHeader>();
    for (int l = 0; l < KD_LEVELS; ++l) {
        ix.table[l] = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(ctx->kd.cells.p) + ctx->kd.table_offset[l]);
        ix.mask[l] = ctx->kd.table_mask[l];
    }
    ix.stats = ctx->kd.stats.p ? ctx->kd.stats.as<unsigned long long>() : nullptr;
    return ix;
}

size_t cell_table_bytes(int64_t M, uint32_t* masks, size_t* offsets) {
    size_t off = 0;
    for (int l = 0; l < KD_LEVELS; ++l) {
        uint64_t want = (uint64_t)(1.5 * (double)M) >> l;
        uint32_t sz = 1024;
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
void kd_reserve_capacity(pls_context* ctx, int64_t need) {
    KdMap& kd = ctx->kd;
    if (need <= kd.cap_points) return;
    cudaStream_t st = ctx->stream;
    const int64_t steady = (int64_t)(1.3 * (double)kd.max_frame * (double)(ctx->cfg.local_map_size + 1));
    int64_t cap = need + need / 4 + 64;
    if (cap < steady) cap = steady;
    const size_t C = (size_t)cap;
    kd.store[kd.cur].reserve_exact(C * sizeof(float4), st, true);  // the live points survive
    kd.store[kd.cur ^ 1].reserve_exact(C * sizeof(float4), st);
    kd.morton.reserve_exact(C * sizeof(uint64_t), st);
    kd.order.reserve_exact(C * sizeof(uint32_t), st);
    kd.sorted.reserve_exact(C * sizeof(float4), st);
    kd.normals.reserve_exact(C * sizeof(float4), st);
    kd.inv_order.reserve_exact(C * sizeof(uint32_t), st);
    kd.nodes.reserve_exact(C * 64, st);
    kd.parent.reserve_exact(2 * C * sizeof(int), st);
    kd.visit.reserve_exact(C * sizeof(int) + C * sizeof(int4) + 16, st);
    kd.cells.reserve_exact(cell_table_bytes(cap, nullptr, nullptr), st);
    kd.cap_points = cap;
}

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
    const float4* pts = kd.store[kd.cur].as<float4>();
    PLS_CUDA(cudaMemsetAsync(kd.normals.p, 0, (size_t)M * sizeof(float4), st));
    kd.grid_hdr.reserve(sizeof(KdGridHeader), st);
    static const float cell_target = getenv("PLS_KD_CELL") ? (float)atof(getenv("PLS_KD_CELL")) : KD_CELL_TARGET;
    kd_grid_header_kernel<<<1, 32, 0, st>>>(kd.bbox.as<int>(), kd.grid_hdr.as<KdGridHeader>(), cell_target);
    PLS_CHECK_LAUNCH();
    kd.bbox_clean = true;
    kd_morton_kernel<<<grid_for(M, 256, 8 * kNumSMs), 256, 0, st>>>(pts, M, kd.grid_hdr.as<KdGridHeader>(),
                                                                     kd.morton.as<uint64_t>(), kd.order.as<uint32_t>());
    PLS_CHECK_LAUNCH();
    uint64_t* sk;
    uint32_t* sv;
    radix_sort_pairs(ctx, kd.morton.as<uint64_t>(), kd.order.as<uint32_t>(), M, (3 * KD_COORD_BITS + 7) / 8, &sk, &sv,
                     kd.cap_points);
    kd_gather_kernel<<<grid_for(M, 256, 8 * kNumSMs), 256, 0, st>>>(pts, sv, M, kd.sorted.as<float4>(),
                                                                     kd.inv_order.as<uint32_t>());
    PLS_CHECK_LAUNCH();
    {   // cell tables: level l gets a power-of-two table of >= 1.5 M / 2^l slots (overflow falls back to the BVH)
        CellTables T;
        const size_t off = cell_table_bytes(M, kd.table_mask, kd.table_offset);
        kd.cells.reserve(off, st);
        PLS_CUDA(cudaMemsetAsync(kd.cells.p, 0, off, st));
        for (int l = 0; l < KD_LEVELS; ++l) {
            T.table[l] = reinterpret_cast<uint4*>(reinterpret_cast<char*>(kd.cells.p) + kd.table_offset[l]);
            T.mask[l] = kd.table_mask[l];
        }
        kd_cells_kernel<<<grid_for(M, 256, 8 * kNumSMs), 256, 0, st>>>(sk, M, T, kd.grid_hdr.as<KdGridHeader>());
        PLS_CHECK_LAUNCH();
    }
    if (M > 1) {
        int* visit = kd.visit.as<int>();
        int4* ranges = reinterpret_cast<int4*>(reinterpret_cast<char*>(kd.visit.p) + (((size_t)M * sizeof(int) + 15) / 16) * 16);
        PLS_CUDA(cudaMemsetAsync(visit, 0, (size_t)M * sizeof(int), st));
        kd_hierarchy_kernel<<<grid_for(M - 1, 128, 1 << 20), 128, 0, st>>>(sk, (int)M, ranges, kd.parent.as<int>());
        PLS_CHECK_LAUNCH();
        kd_boxes_kernel<<<grid_for(2 * M, 128, 1 << 20), 128, 0, st>>>(kd.sorted.as<float4>(), (int)M, ranges,
                                                                        kd.parent.as<int>(), visit, kd.nodes.as<float4>());
        PLS_CHECK_LAUNCH();
    }
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
    ctx->kd.max_frame = 0;   // the buffers (cap_points) are kept: a re-initialised sequence reuses them
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

// One fused ICP iteration over the device-resident queries (float4 in ctx->queries, count in
// SC_QUERY_COUNT); writes block partials to ctx->partials and returns the block count.
int kdmap_icp_iteration(pls_context* ctx, int64_t query_bound, int rank, int num_ranks, bool first, float fuse_threshold,
                        bool* solved) {
    PLS_REQUIRE(ctx->kd.valid, "kd map: search before any update");
    *solved = false;
    const int64_t mine = (query_bound + num_ranks - 1) / num_ranks;
    const uint32_t* nq_dev = reinterpret_cast<const uint32_t*>(&frame_result_dev(ctx)->counts[1]);
    // credited per executed iteration by the caller (the launch is a no-op once ICP converged)
    ProfileScope ps(ctx, 0, 0.0, false);
    // lanes per query of the split search kernels (0 = the fused thread-per-query kernel)
    static const int group = getenv("PLS_KD_GROUP") ? atoi(getenv("PLS_KD_GROUP")) : 4;
    if (group == 2 || group == 4 || group == 8) {
        *solved = fuse_threshold >= 0.f;
        if (group == 2) return launch_group_iteration<2>(ctx, make_index(ctx), mine, nq_dev, rank, num_ranks, first, fuse_threshold);
        if (group == 4) return launch_group_iteration<4>(ctx, make_index(ctx), mine, nq_dev, rank, num_ranks, first, fuse_threshold);
        return launch_group_iteration<8>(ctx, make_index(ctx), mine, nq_dev, rank, num_ranks, first, fuse_threshold);
    }
    if (first) PLS_CUDA(cudaMemsetAsync(ctx->nn_prev.p, 0xff, (size_t)query_bound * sizeof(int), ctx->stream));
    const int blocks = grid_for(mine, KD_ITER_THREADS, 8 * kNumSMs);
    ctx->partials.reserve((size_t)blocks * NACC * sizeof(double), ctx->stream);
    kd_icp_iter_kernel<<<blocks, KD_ITER_THREADS, 0, ctx->stream>>>(
        make_index(ctx), ctx->cfg.num_neighbors_normals, ctx->query_ptr, nq_dev, (int64_t)rank, (int64_t)num_ranks,
        frame_result_dev(ctx), ctx->cfg.scheme, ctx->cfg.sigma, ctx->nn_prev.as<int>(), ctx->partials.as<double>());
    PLS_CHECK_LAUNCH();
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
    kd_search_kernel<<<grid_for(n, 128, 8 * kNumSMs), 128, 0, ctx->stream>>>(
        make_index(ctx), ctx->cfg.num_neighbors_normals, d, n, (float*)onb.dev, (float*)onr.dev, (long long*)oix.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, onb);
    finish_out(ctx, onr);
    finish_out(ctx, oix);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

}  // extern "C"
