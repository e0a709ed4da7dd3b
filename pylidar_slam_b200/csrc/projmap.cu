// K3/K4 -- the projective local map (replaces ProjectiveLocalMap, slam/odometry/local_map.py:91-240,
// compute_normal_map and compute_neighbors, slam/common/geometry.py:240-295,397-439).
//
//   normal_map_kernel       5x5 (k x k) zero-padded box sums of v and v v^T from a shared-memory
//                           tile, n ~ (sum v v^T)^-1 sum v through the cofactor matrix, float32
//   model_zbuf_kernel       build_model (local_map.py:177-202): every stored frame k is re-expressed
//   model_resolve_kernel    in the newest frame (p' = R_k v + t_k, n' = R_k n, masked) and re-projected
//                           with its own closest-wins z-buffer into _model_vmap / _model_nmap [K,3,H,W]
//   query_zbuf_kernel       nearest_neighbor_search (local_map.py:205-235): the transformed queries
//                           are z-buffered into the target vertex map (one survivor per pixel)
//   proj_icp_iter_kernel    per pixel: argmin_k |p - v_k| over the K model maps (first minimum wins,
//                           null candidates skipped), gather the winner's point and normal, point-to-
//                           plane residual / Jacobian / weight and the block-reduced normal equations.
//                           Streams HW*12*(K+1) bytes per launch: the HBM-bound correspondence kernel.
//   proj_pairs_kernel       the same association materialised per pixel for the fine-grained API
#include <stdlib.h>

#include "gn_device.cuh"
#include "internal.cuh"
#include "pose_device.cuh"
#include "projection_device.cuh"

namespace pls {

namespace {

inline int grid_for(int64_t n, int threads, int cap_blocks = 16 * kNumSMs) {
    int64_t b = (n + threads - 1) / threads;
    return (int)(b < 1 ? 1 : (b > cap_blocks ? cap_blocks : b));
}

// ------------------------------------------------------------------------------------------ K3
constexpr int NM_TX = 32, NM_TY = 8, NM_MAXR = 4;  // kernel sizes up to 9

__global__ void __launch_bounds__(NM_TX* NM_TY)
normal_map_kernel(const float* __restrict__ vmap, int batch, int H, int W, int ksize, float* __restrict__ out) {
    __shared__ float tile[3][NM_TY + 2 * NM_MAXR][NM_TX + 2 * NM_MAXR + 1];
    const int r = ksize / 2;
    const int b = blockIdx.z;
    const int x0 = blockIdx.x * NM_TX, y0 = blockIdx.y * NM_TY;
    const int64_t hw = (int64_t)H * W;
    const float* v = vmap + (size_t)b * 3 * hw;
    const int tw = NM_TX + 2 * r, th = NM_TY + 2 * r;
    for (int i = threadIdx.y * NM_TX + threadIdx.x; i < tw * th; i += NM_TX * NM_TY) {
        int ty = i / tw, tx = i - ty * tw;
        int gx = x0 + tx - r, gy = y0 + ty - r;
        bool in = gx >= 0 && gx < W && gy >= 0 && gy < H;
        int64_t g = (int64_t)gy * W + gx;
        tile[0][ty][tx] = in ? v[g] : 0.f;
        tile[1][ty][tx] = in ? v[hw + g] : 0.f;
        tile[2][ty][tx] = in ? v[2 * hw + g] : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= W || y >= H) return;
    // Operation order and rounding follow the reference exactly (its normals are dominated by float32
    // rounding: it inverts the UNCENTRED second-moment matrix, cond ~ r^2/sigma^2 >> 1/eps, so only a
    // bit-faithful evaluation reproduces them):
    //  * box sums = sequential float32 adds over the window, rows outer / columns inner, of separately
    //    rounded products (conv2d with a ones kernel over `v` and `v v^T`, geometry.py:258-268)
    float sx = 0.f, sy = 0.f, sz = 0.f, sxx = 0.f, sxy = 0.f, sxz = 0.f, syy = 0.f, syz = 0.f, szz = 0.f;
    for (int dy = 0; dy < ksize; ++dy)
        for (int dx = 0; dx < ksize; ++dx) {
            const float px = tile[0][threadIdx.y + dy][threadIdx.x + dx];
            const float py = tile[1][threadIdx.y + dy][threadIdx.x + dx];
            const float pz = tile[2][threadIdx.y + dy][threadIdx.x + dx];
            sx = __fadd_rn(sx, px); sy = __fadd_rn(sy, py); sz = __fadd_rn(sz, pz);
            sxx = __fadd_rn(sxx, __fmul_rn(px, px)); sxy = __fadd_rn(sxy, __fmul_rn(px, py));
            sxz = __fadd_rn(sxz, __fmul_rn(px, pz)); syy = __fadd_rn(syy, __fmul_rn(py, py));
            syz = __fadd_rn(syz, __fmul_rn(py, pz)); szz = __fadd_rn(szz, __fmul_rn(pz, pz));
        }
    //  * cofactor rows c_i = A[i-2] x A[i-1], each component fma(a1, b2, -(a2 * b1))  (geometry.py:65-76)
    const float A0[3] = {sxx, sxy, sxz}, A1[3] = {sxy, syy, syz}, A2[3] = {sxz, syz, szz};
    auto cross = [](const float* a, const float* b, float* c) {
        c[0] = __fmaf_rn(a[1], b[2], -__fmul_rn(a[2], b[1]));
        c[1] = __fmaf_rn(a[2], b[0], -__fmul_rn(a[0], b[2]));
        c[2] = __fmaf_rn(a[0], b[1], -__fmul_rn(a[1], b[0]));
    };
    auto dot3 = [](float a0, float b0, float a1, float b1, float a2, float b2) {
        return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
    };
    float c0[3], c1[3], c2[3];
    cross(A1, A2, c0);
    cross(A2, A0, c1);
    cross(A0, A1, c2);
    //  * det = mean of the three row expansions, ((d0 + d1) + d2) / 3  (geometry.py:87)
    const float d0 = dot3(c0[0], A0[0], c0[1], A0[1], c0[2], A0[2]);
    const float d1 = dot3(c1[0], A1[0], c1[1], A1[1], c1[2], A1[2]);
    const float d2 = dot3(c2[0], A2[0], c2[1], A2[1], c2[2], A2[2]);
    const float det = __fdiv_rn(__fadd_rn(__fadd_rn(d0, d1), d2), 3.0f);
    float n[3] = {0.f, 0.f, 0.f};
    if (fabsf(det) > 1e-6f) {
        //  * n = (cof / det)^T b : n_i = sum_j (cof[j][i] / det) b_j  (geometry.py:89-114,273)
        n[0] = dot3(__fdiv_rn(c0[0], det), sx, __fdiv_rn(c1[0], det), sy, __fdiv_rn(c2[0], det), sz);
        n[1] = dot3(__fdiv_rn(c0[1], det), sx, __fdiv_rn(c1[1], det), sy, __fdiv_rn(c2[1], det), sz);
        n[2] = dot3(__fdiv_rn(c0[2], det), sx, __fdiv_rn(c1[2], det), sy, __fdiv_rn(c2[2], det), sz);
        //  * norm = sqrt(fma(n2, n2, fma(n1, n1, n0 * n0)))  (torch.norm's vectorised kernel)
        float nn = __fsqrt_rn(__fmaf_rn(n[2], n[2], __fmaf_rn(n[1], n[1], __fmul_rn(n[0], n[0]))));
        if (nn == 0.f) nn = 1.f;
        n[0] = __fdiv_rn(n[0], nn); n[1] = __fdiv_rn(n[1], nn); n[2] = __fdiv_rn(n[2], nn);
    }
    const float cx = tile[0][threadIdx.y + r][threadIdx.x + r], cy = tile[1][threadIdx.y + r][threadIdx.x + r],
                cz = tile[2][threadIdx.y + r][threadIdx.x + r];
    if (cx == 0.f && cy == 0.f && cz == 0.f) n[0] = n[1] = n[2] = 0.f;  // torch.norm(vertex) == 0
    float* o = out + (size_t)b * 3 * hw + (int64_t)y * W + x;
    o[0] = n[0];
    o[hw] = n[1];
    o[2 * hw] = n[2];
}

// ------------------------------------------------------------------------------------------ model layout
// The re-projected model maps are stored TILE-INTERLEAVED in HBM: [tile = pix / 128][row = k * 3 + c][pix % 128]
// with a fixed tile stride of Kcap * 3 * 128 floats (Kcap = local_map_size).  All candidates of a 128-pixel
// tile are then ONE contiguous block (K * 3 * 512 bytes): a single TMA bulk copy per tile and a purely
// sequential HBM stream for the correspondence kernel.  The reference's planar [K,3,H,W] view exists only
// at the API boundary (pls_projmap_model converts).
constexpr int PT_TILE = 128;
__device__ __host__ __forceinline__ size_t model_off(int64_t pix, int row, int kcap) {
    return (size_t)(pix / PT_TILE) * ((size_t)kcap * 3 * PT_TILE) + (size_t)row * PT_TILE + (size_t)(pix % PT_TILE);
}
// The model NORMALS are only ever gathered for the winning candidate of a pixel, so they are stored as one
// float4 per (tile, k, pixel): the gather is a single 16-byte access (one 32-byte sector) instead of three.
__device__ __host__ __forceinline__ size_t normal_off(int64_t pix, int k, int kcap) {
    return ((size_t)(pix / PT_TILE) * kcap + (size_t)k) * PT_TILE + (size_t)(pix % PT_TILE);  // float4 units
}

// ------------------------------------------------------------------------------------------ model rebuild
struct PoseSet {
    const float* poses;  // [K][16] device
};

__device__ __forceinline__ bool model_point(const float* __restrict__ vmaps, const float* __restrict__ P, int64_t hw,
                                            int64_t src, float* p) {
    const float x = vmaps[src], y = vmaps[hw + src], z = vmaps[2 * hw + src];
    // mask_not_null (geometry.py:157-177): any channel non-zero
    if (fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 0.f) {
        p[0] = x * P[0] + y * P[1] + z * P[2] + P[3];
        p[1] = x * P[4] + y * P[5] + z * P[6] + P[7];
        p[2] = x * P[8] + y * P[9] + z * P[10] + P[11];
        return true;
    }
    return false;
}

// (a rank of a sharded job keeps only the pixels [pix_lo, pix_hi) of its own tiles)
__global__ void model_zbuf_kernel(const float* __restrict__ vmaps, const float* __restrict__ poses, int K, int head,
                                  int slots, ProjConst pc, int pix_lo, int pix_hi, unsigned long long* __restrict__ zbuf) {
    const int64_t hw = (int64_t)pc.H * pc.W;
    const int64_t total = (int64_t)K * hw;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = g / hw, src = g - k * hw;
        float p[3];
        if (!model_point(vmaps + (size_t)((head + k) % slots) * 3 * hw, poses + 16 * k, hw, src, p)) continue;
        int pix;
        float r;
        if (project_to_pixel(p[0], p[1], p[2], pc, pix, r) && pix >= pix_lo && pix < pix_hi) {
            unsigned long long key = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(uint32_t)src;
            atomicMin(&zbuf[k * hw + pix], key);
        }
    }
}

__global__ void model_resolve_kernel(const float* __restrict__ vmaps, const float* __restrict__ nmaps,
                                     const float* __restrict__ poses, int K, int head, int slots, int kcap, int64_t hw,
                                     int64_t pix_lo, int64_t pix_hi, const unsigned long long* __restrict__ zbuf,
                                     float* __restrict__ model_v, float4* __restrict__ model_n) {
    const int64_t span = pix_hi - pix_lo, total = (int64_t)K * span;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = g / span, pix = pix_lo + (g - k * span);
        const unsigned long long key = zbuf[k * hw + pix];
        float p[3] = {0.f, 0.f, 0.f}, n[3] = {0.f, 0.f, 0.f};
        if (key != ~0ull) {
            const int64_t src = (int64_t)(uint32_t)(key & 0xffffffffull);
            const float* P = poses + 16 * k;
            const size_t slot = (size_t)((head + k) % slots);
            model_point(vmaps + slot * 3 * hw, P, hw, src, p);
            const float* nm = nmaps + slot * 3 * hw;
            const float nx = nm[src], ny = nm[hw + src], nz = nm[2 * hw + src];
            n[0] = P[0] * nx + P[1] * ny + P[2] * nz;
            n[1] = P[4] * nx + P[5] * ny + P[6] * nz;
            n[2] = P[8] * nx + P[9] * ny + P[10] * nz;
        }
        const size_t o = model_off(pix, (int)k * 3, kcap);
        model_v[o] = p[0]; model_v[o + PT_TILE] = p[1]; model_v[o + 2 * PT_TILE] = p[2];
        model_n[normal_off(pix, (int)k, kcap)] = make_float4(n[0], n[1], n[2], 0.f);
    }
}

// ------------------------------------------------------------------------------------------ search
__device__ __forceinline__ void load_T(const float* __restrict__ T, float* sT) {
    if (threadIdx.x < 12) sT[threadIdx.x] = T[threadIdx.x];
    __syncthreads();
}

__global__ void query_zbuf_kernel(const float4* __restrict__ queries, const uint32_t* __restrict__ nq_dev, int64_t nq_host,
                                  const float* __restrict__ T, const int* __restrict__ done, ProjConst pc, int pix_lo,
                                  int pix_hi, unsigned long long* __restrict__ zbuf) {
    if (done && *done) return;
    __shared__ float sT[12];
    if (T) load_T(T, sT);
    const int64_t nq = nq_dev ? (int64_t)*nq_dev : nq_host;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p0 = queries[i];
        float p[3] = {p0.x, p0.y, p0.z};
        if (T) {
            p[0] = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
            p[1] = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
            p[2] = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
        }
        int pix;
        float r;
        if (project_to_pixel(p[0], p[1], p[2], pc, pix, r) && pix >= pix_lo && pix < pix_hi) {
            unsigned long long key = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned long long)(uint32_t)i;
            atomicMin(&zbuf[pix], key);
        }
    }
}

// z-buffer winners -> the target vertex map of this iteration, float4 (p transformed, valid flag) per pixel
__global__ void query_resolve_kernel(unsigned long long* __restrict__ zbuf, const float4* __restrict__ queries,
                                     const float* __restrict__ T, const int* __restrict__ done, int64_t pix_lo,
                                     int64_t pix_hi, float4* __restrict__ tgt) {
    if (done && *done) return;
    __shared__ float sT[12];
    load_T(T, sT);
    for (int64_t pix = pix_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < pix_hi; pix += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = zbuf[pix];
        zbuf[pix] = ~0ull;  // leave the z-buffer cleared for the next iteration (no separate memset)
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (key != ~0ull) {
            const float4 p0 = queries[(uint32_t)(key & 0xffffffffull)];
            o.x = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
            o.y = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
            o.z = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
            o.w = 1.f;
        }
        tgt[pix] = o;
    }
}

// Arg-min over distances WITHOUT the square roots in the common case.  The reference takes the arg-min over
// torch.norm (then torch.min keeps the first minimum: geometry.py:424-428), and two different squared distances can share
// one correctly rounded root -- but only if they lie within 2^-22 of each other (the pre-image of a float under sqrt is at
// most that wide, relatively).  The tile loop therefore compares squares against best2 * (1 - 2^-21), records whether any
// comparison fell inside that band, and redoes such a pixel (about one in 50 000) with the roots.

// acc += [ (wJ)(wJ)^T upper, (wJ)(wr), (wr)^2, r^2, 1 ] in float32: a thread of the tile-streaming kernel meets a
// handful of pixels only (tiles / CTAs), their sum goes to float64 before the block reduction -- each partial sum
// carries ~1e-7 relative rounding, the half-million-term totals stay far more accurate than the reference's float32
// sgemm while the accumulators cost 30 registers instead of 60
__device__ __forceinline__ void accumulate_normal_equations_f32(float* acc, const float* J, float w, float wr, float r) {
    float wj[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) wj[a] = J[a] * w;
    int k = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b = a; b < 6; ++b) acc[k++] += wj[a] * wj[b];
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += wj[a] * wr;
    acc[27] += wr * wr;
    acc[28] += r * r;
    acc[29] += 1.0f;
}

// argmin over the K candidates at one pixel; returns false if none is valid
__device__ __forceinline__ bool pixel_argmin(const float* __restrict__ model_v, const float4* __restrict__ model_n, int K,
                                             int kcap, int64_t pix, const float* p, float* q, float* n) {
    float best = __int_as_float(0x7f800000);
    int kbest = -1;
    float bq[3] = {0.f, 0.f, 0.f};
    const float* mvb = model_v + model_off(pix, 0, kcap);
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float* mv = mvb + (size_t)k * 3 * PT_TILE;
        const float x = mv[0], y = mv[PT_TILE], z = mv[2 * PT_TILE];
        if (fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 0.f) {
            const float dx = p[0] - x, dy = p[1] - y, dz = p[2] - z;
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            if (d < best) {  // torch.min keeps the first minimum
                best = d;
                kbest = k;
                bq[0] = x; bq[1] = y; bq[2] = z;
            }
        }
    }
    if (kbest < 0) return false;
    q[0] = bq[0]; q[1] = bq[1]; q[2] = bq[2];
    const float4 mn = model_n[normal_off(pix, kbest, kcap)];
    n[0] = mn.x; n[1] = mn.y; n[2] = mn.z;
    return true;
}

constexpr int PJ_THREADS = 256;

__global__ void __launch_bounds__(PJ_THREADS)
proj_icp_iter_kernel(const float* __restrict__ model_v, const float4* __restrict__ model_n, int K, int kcap,
                     const unsigned long long* __restrict__ zbuf, const float4* __restrict__ queries,
                     const FrameResult* __restrict__ fr, int64_t pix_begin, int64_t pix_end, int scheme, float sigma,
                     double* __restrict__ partials) {
    if (fr->done) return;
    __shared__ float sT[12];
    load_T(fr->T, sT);
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (int64_t pix = pix_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < pix_end;
         pix += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = zbuf[pix];
        if (key == ~0ull) continue;
        const float4 p0 = queries[(uint32_t)(key & 0xffffffffull)];
        float p[3], q[3], n[3];
        p[0] = p0.x * sT[0] + p0.y * sT[1] + p0.z * sT[2] + sT[3];
        p[1] = p0.x * sT[4] + p0.y * sT[5] + p0.z * sT[6] + sT[7];
        p[2] = p0.x * sT[8] + p0.y * sT[9] + p0.z * sT[10] + sT[11];
        if (!pixel_argmin(model_v, model_n, K, kcap, pix, p, q, n)) continue;
        float J[6];
        const float r = p2plane_residual_jacobian_identity(p, q, n, J);
        const float w = ls_weight<float>(scheme, sigma, r, p, q);
        accumulate_normal_equations<float>(acc, J, w, r * w, r);
    }
    block_reduce_store<PJ_THREADS>(acc, partials + (size_t)blockIdx.x * NACC);
}

// ---- TMA-staged variant ---------------------------------------------------------------------------
// Persistent CTAs (2 per SM); each loops over 128-pixel tiles.  For a tile, ONE thread issues K*3 bulk
// async copies (cp.async.bulk, the TMA engine; 512 contiguous bytes per candidate plane) that land in a
// [K*3][128] shared-memory stage and complete on that stage's mbarrier; 3 stages keep two tiles (up to
// 60 KB per CTA) in flight while the third is consumed, with no registers tied up by outstanding loads.
// The consumers (one thread per pixel) read conflict-free from shared memory.
constexpr int PT_STAGES = 3;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// L2 eviction-priority policies (the fixed encodings of createpolicy.fractional.L2::evict_last / evict_first, fraction 1)
constexpr unsigned long long L2_EVICT_LAST = 0x14F0000000000000ull, L2_EVICT_FIRST = 0x12F0000000000000ull;
__device__ __forceinline__ void bulk_copy_g2s_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ float ldg_f32_hint(const float* p, unsigned long long policy) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4* p, unsigned long long policy) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(PT_TILE)
proj_icp_tma_kernel(const float* __restrict__ model_v, const float4* __restrict__ model_n, int K, int kcap,
                    const float4* __restrict__ tgt, const FrameResult* __restrict__ fr, int64_t tile_begin, int64_t tile_end, int scheme, float sigma,
                    int stages, int ktma, int64_t resident_end, unsigned long long policy_resident,
                    unsigned long long policy_stream, double* __restrict__ partials) {
    if (fr->done) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[PT_STAGES];
    float* stage_base = reinterpret_cast<float*>(smem_raw);
    // candidates [0, ktma) are staged through shared memory by the TMA engine; candidates [ktma, K) are read by
    // the consumers themselves with coalesced loads (the tile-interleaved rows are 512 contiguous bytes), issued
    // BEFORE the mbarrier wait: the LSU path and the TMA path pull from HBM concurrently
    const int rows = ktma * 3;
    const uint32_t model_bytes = (uint32_t)rows * PT_TILE * sizeof(float);
    const uint32_t stage_floats = (uint32_t)(rows + 4) * PT_TILE;  // + the tile of the target vertex map (float4 per pixel)
    const uint32_t stage_bytes = stage_floats * sizeof(float);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t first_tile = tile_begin + blockIdx.x;
    const int64_t stride = gridDim.x;
    auto issue = [&](int64_t tile, int s) {  // thread 0 only
        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
        // the tile's K*3 candidate rows are contiguous in the tile-interleaved layout: one bulk copy; a second
        // one brings the tile of the (z-buffered, already transformed) target vertex map
        // the same model is streamed once per ICP iteration: the tiles below `resident_end` ask L2 to keep them
        // (evict-last) and are served from L2 from the second iteration on; the rest passes through (evict-first)
        // without pushing them out
        float* dst = stage_base + (size_t)s * stage_floats;
        bulk_copy_g2s_hint(dst, model_v + (size_t)tile * ((size_t)kcap * 3 * PT_TILE), model_bytes, &full_bar[s],
                           tile < resident_end ? policy_resident : policy_stream);
        bulk_copy_g2s(dst + (size_t)rows * PT_TILE, tgt + tile * PT_TILE, PT_TILE * sizeof(float4), &full_bar[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            const int64_t t = first_tile + (int64_t)s * stride;
            if (t < tile_end) issue(t, s);
        }
    }
    float acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.f;
    bool pend = false;
    float pp[3] = {0.f, 0.f, 0.f}, pq[3] = {0.f, 0.f, 0.f}, pn[3] = {0.f, 0.f, 0.f};
    int it = 0;
    for (int64_t tile = first_tile; tile < tile_end; tile += stride, ++it) {
        const int s = it % stages;
        const uint32_t parity = (uint32_t)((it / stages) & 1);
        const int64_t pix = tile * PT_TILE + threadIdx.x;
        constexpr int KDIRECT_MAX = 10;
        float dv[3 * KDIRECT_MAX];
        {
            const float* g = model_v + (size_t)tile * ((size_t)kcap * 3 * PT_TILE) + (size_t)ktma * 3 * PT_TILE + threadIdx.x;
            const unsigned long long policy = tile < resident_end ? policy_resident : policy_stream;
#pragma unroll
            for (int j = 0; j < 3 * KDIRECT_MAX; ++j)
                dv[j] = (j < (K - ktma) * 3) ? ldg_f32_hint(g + (size_t)j * PT_TILE, policy) : 0.f;
        }
        mbar_wait(&full_bar[s], parity);
        bool matched = false;
        float q[3] = {0.f, 0.f, 0.f};
        int kbest = -1;
        const float4 tp = reinterpret_cast<const float4*>(stage_base + (size_t)s * stage_floats + (size_t)rows * PT_TILE)[threadIdx.x];
        const float p[3] = {tp.x, tp.y, tp.z};
        const bool has = tp.w != 0.f;
        if (has) {
            const float* st = stage_base + (size_t)s * stage_floats + threadIdx.x;
            // Branch-free arg-min on squared distances.  `band` records whether any comparison fell inside the
            // 2^-22 band where the squares cannot decide (see above): the pixel is then redone with the roots.
            float best2 = __int_as_float(0x7f800000);  // squared distance of the best candidate so far
            bool band = false;
#pragma unroll 4
            for (int k = 0; k < ktma; ++k) {
                const float x = st[(3 * k) * PT_TILE], y = st[(3 * k + 1) * PT_TILE], z = st[(3 * k + 2) * PT_TILE];
                const float dx = p[0] - x, dy = p[1] - y, dz = p[2] - z;
                const float d2 = dx * dx + dy * dy + dz * dz;
                // torch.min keeps the first minimum; null candidates (all channels 0) do not compete
                const bool live = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 0.f;
                const bool lt = live && d2 < best2 * 0.99999952f;
                band |= live && !lt && d2 < best2;
                best2 = lt ? d2 : best2;
                kbest = lt ? k : kbest;
                q[0] = lt ? x : q[0]; q[1] = lt ? y : q[1]; q[2] = lt ? z : q[2];
            }
#pragma unroll
            for (int j = 0; j < KDIRECT_MAX; ++j) {
                if (j < K - ktma) {
                    const float x = dv[3 * j], y = dv[3 * j + 1], z = dv[3 * j + 2];
                    const float dx = p[0] - x, dy = p[1] - y, dz = p[2] - z;
                    const float d2 = dx * dx + dy * dy + dz * dz;
                    const bool live = fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 0.f;
                    const bool lt = live && d2 < best2 * 0.99999952f;
                    band |= live && !lt && d2 < best2;
                    best2 = lt ? d2 : best2;
                    kbest = lt ? ktma + j : kbest;
                    q[0] = lt ? x : q[0]; q[1] = lt ? y : q[1]; q[2] = lt ? z : q[2];
                }
            }
            if (band) {  // about one pixel in 50 000: the reference's arg-min over the roots, first minimum wins
                float best = __int_as_float(0x7f800000);
                kbest = -1;
                for (int k = 0; k < K; ++k) {
                    float x, y, z;
                    if (k < ktma) {
                        x = st[(3 * k) * PT_TILE]; y = st[(3 * k + 1) * PT_TILE]; z = st[(3 * k + 2) * PT_TILE];
                    } else {
                        const float* gk = model_v + (size_t)tile * ((size_t)kcap * 3 * PT_TILE) + (size_t)k * 3 * PT_TILE + threadIdx.x;
                        x = __ldg(gk); y = __ldg(gk + PT_TILE); z = __ldg(gk + 2 * PT_TILE);
                    }
                    if (fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 0.f) {
                        const float dx = p[0] - x, dy = p[1] - y, dz = p[2] - z;
                        const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                        if (d < best) {
                            best = d;
                            kbest = k;
                            q[0] = x; q[1] = y; q[2] = z;
                        }
                    }
                }
            }
            matched = kbest >= 0;
        }
        __syncthreads();  // every consumer is done with stage s before the async proxy refills it
        if (threadIdx.x == 0) {
            const int64_t nt = tile + (int64_t)stages * stride;
            if (nt < tile_end) issue(nt, s);
        }
        // software pipeline: first consume the PREVIOUS tile's correspondence (its normal gather was issued
        // one iteration ago and has had this tile's wait + arg-min to land) ...
        if (pend) {
            float J[6];
            const float r = p2plane_residual_jacobian_identity(pp, pq, pn, J);
            const float w = ls_weight<float>(scheme, sigma, r, pp, pq);
            accumulate_normal_equations_f32(acc, J, w, r * w, r);
        }
        // ... then issue THIS tile's winner-normal gather straight into the pending registers (no copy that
        // would force the load to complete here)
        pend = matched;
        if (matched) {
            const float4 mn = ldg_f4_hint(model_n + normal_off(pix, kbest, kcap), policy_stream);  // one 16-byte gather, no reuse
            pn[0] = mn.x; pn[1] = mn.y; pn[2] = mn.z;
#pragma unroll
            for (int c = 0; c < 3; ++c) { pp[c] = p[c]; pq[c] = q[c]; }
        }
    }
    if (pend) {
        float J[6];
        const float r = p2plane_residual_jacobian_identity(pp, pq, pn, J);
        const float w = ls_weight<float>(scheme, sigma, r, pp, pq);
        accumulate_normal_equations_f32(acc, J, w, r * w, r);
    }
    double acc64[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc64[a] = (double)acc[a];
    block_reduce_store<PT_TILE>(acc64, partials + (size_t)blockIdx.x * NACC);
    // (finishing the iteration in the CTA that arrives last -- as kd_residual_kernel does -- was measured and lost here:
    // 1.51-1.54 vs 1.46 ms per cfg5 frame; the solve stays in icp_step_kernel)
}

// per-pixel association for the fine-grained API: flag + (q, n, p)
__global__ void proj_pairs_kernel(const float* __restrict__ model_v, const float4* __restrict__ model_n, int K, int kcap,
                                  int64_t hw, const unsigned long long* __restrict__ zbuf, const float4* __restrict__ queries,
                                  uint8_t* __restrict__ flags, float* __restrict__ pairs /* [hw][9] */) {
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long key = zbuf[pix];
        uint8_t ok = 0;
        if (key != ~0ull) {
            const float4 p0 = queries[(uint32_t)(key & 0xffffffffull)];
            float p[3] = {p0.x, p0.y, p0.z}, q[3], n[3];
            if (pixel_argmin(model_v, model_n, K, kcap, pix, p, q, n)) {
                ok = 1;
                float* o = pairs + 9 * pix;
                o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
                o[3] = n[0]; o[4] = n[1]; o[5] = n[2];
                o[6] = p[0]; o[7] = p[1]; o[8] = p[2];
            }
        }
        flags[pix] = ok;
    }
}

__global__ void proj_compact_kernel(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ pos,
                                    const float* __restrict__ pairs, int64_t hw, float* __restrict__ out_q,
                                    float* __restrict__ out_n, float* __restrict__ out_p) {
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (int64_t)gridDim.x * blockDim.x) {
        if (!flags[pix]) continue;
        const uint32_t d = pos[pix];
        const float* s = pairs + 9 * pix;
        for (int c = 0; c < 3; ++c) {
            out_q[3 * (size_t)d + c] = s[c];
            if (out_n) out_n[3 * (size_t)d + c] = s[3 + c];
            if (out_p) out_p[3 * (size_t)d + c] = s[6 + c];
        }
    }
}

// stateless compute_neighbors (geometry.py:397-439)
__global__ void compute_neighbors_kernel(const float* __restrict__ tgt, const float* __restrict__ ref,
                                         const float* __restrict__ fields, int K, int C, int64_t hw,
                                         float* __restrict__ out_nb, float* __restrict__ out_f) {
    for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < hw; pix += (int64_t)gridDim.x * blockDim.x) {
        const float p[3] = {tgt[pix], tgt[hw + pix], tgt[2 * hw + pix]};
        const bool tgt_ok = fmaxf(fmaxf(fabsf(p[0]), fabsf(p[1])), fabsf(p[2])) > 0.f;
        float best = __int_as_float(0x7f800000);
        int kbest = 0;  // torch.min over all-inf returns index 0
        if (tgt_ok) {
            for (int k = 0; k < K; ++k) {
                const float* mv = ref + (size_t)k * 3 * hw + pix;
                const float x = mv[0], y = mv[hw], z = mv[2 * hw];
                if (fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) > 0.f) {
                    const float dx = p[0] - x, dy = p[1] - y, dz = p[2] - z;
                    const float d = sqrtf(dx * dx + dy * dy + dz * dz);
                    if (d < best) { best = d; kbest = k; }
                }
            }
        }
        const float* mv = ref + (size_t)kbest * 3 * hw + pix;
        out_nb[pix] = tgt_ok ? mv[0] : 0.f;
        out_nb[hw + pix] = tgt_ok ? mv[hw] : 0.f;
        out_nb[2 * hw + pix] = tgt_ok ? mv[2 * hw] : 0.f;
        if (fields)
            for (int c = 0; c < C; ++c) out_f[(size_t)c * hw + pix] = fields[((size_t)kbest * C + c) * hw + pix];
    }
}

// tile-interleaved -> the reference's planar [K,3,H,W]
__global__ void model_export_kernel(const float* __restrict__ tiled, int K, int kcap, int64_t hw, float* __restrict__ planar) {
    const int64_t total = (int64_t)K * 3 * hw;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = g / hw, pix = g - row * hw;
        planar[g] = tiled[model_off(pix, (int)row, kcap)];
    }
}

__global__ void normal_export_kernel(const float4* __restrict__ tiled, int K, int kcap, int64_t hw, float* __restrict__ planar) {
    const int64_t total = (int64_t)K * hw;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = g / hw, pix = g - k * hw;
        const float4 n = tiled[normal_off(pix, (int)k, kcap)];
        planar[(k * 3 + 0) * hw + pix] = n.x;
        planar[(k * 3 + 1) * hw + pix] = n.y;
        planar[(k * 3 + 2) * hw + pix] = n.z;
    }
}

// The pixel range of the model this rank needs: all of it, or -- when the ICP iterations of this job are split across
// ranks (icp_shards) and the TMA path applies -- the tiles the rank reduces (projmap_icp_iteration takes the same split).
void model_range(pls_context* ctx, bool full, int64_t* lo, int64_t* hi) {
    const int64_t hw = (int64_t)ctx->cfg.height * ctx->cfg.width;
    *lo = 0;
    *hi = hw;
    if (full || hw % PT_TILE != 0 || !icp_shards(ctx, hw)) return;
    const int64_t tiles = hw / PT_TILE;
    const int rank = comm_rank(ctx), size = comm_size(ctx);
    *lo = (tiles * rank / size) * PT_TILE;
    *hi = (tiles * (rank + 1) / size) * PT_TILE;
}

void rebuild_model(pls_context* ctx, bool full = false) {
    ProjMap& pm = ctx->pm;
    cudaStream_t st = ctx->stream;
    const int H = ctx->cfg.height, W = ctx->cfg.width;
    const int64_t hw = (int64_t)H * W;
    const int K = pm.K;
    if (K == 0) return;
    int64_t lo, hi;
    model_range(ctx, full, &lo, &hi);
    ProfileScope ps(ctx, 2, (double)K * (hw * 24.0 + (hi - lo) * (24.0 + 16.0)));
    pm.poses.reserve((size_t)ctx->cfg.local_map_size * 16 * sizeof(float) + 64, st);
    PLS_CUDA(cudaMemcpyAsync(pm.poses.p, pm.host_poses.data(), (size_t)K * 16 * sizeof(float), cudaMemcpyHostToDevice, st));
    pm.zbuf.reserve((size_t)(K > 1 ? K : 1) * hw * sizeof(unsigned long long), st);
    const int kcap = ctx->cfg.local_map_size;
    const size_t model_floats = (size_t)((hw + PT_TILE - 1) / PT_TILE) * kcap * 3 * PT_TILE;
    pm.model_v.reserve(model_floats * sizeof(float), st);
    pm.model_n.reserve((size_t)((hw + PT_TILE - 1) / PT_TILE) * kcap * PT_TILE * sizeof(float4), st);
    if (lo == 0 && hi == hw)
        PLS_CUDA(cudaMemsetAsync(pm.zbuf.p, 0xff, (size_t)K * hw * sizeof(unsigned long long), st));
    else  // only the rank's own pixel columns of each of the K z-buffers
        PLS_CUDA(cudaMemset2DAsync(pm.zbuf.as<unsigned long long>() + lo, (size_t)hw * sizeof(unsigned long long), 0xff,
                                   (size_t)(hi - lo) * sizeof(unsigned long long), (size_t)K, st));
    ProjConst pc = make_proj_const(H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg);
    const int slots = ctx->cfg.local_map_size + 1;
    model_zbuf_kernel<<<grid_for(K * hw, 256), 256, 0, st>>>(pm.vmaps.as<float>(), pm.poses.as<float>(), K, pm.head, slots,
                                                             pc, (int)lo, (int)hi, pm.zbuf.as<unsigned long long>());
    PLS_CHECK_LAUNCH();
    model_resolve_kernel<<<grid_for(K * (hi - lo), 256), 256, 0, st>>>(pm.vmaps.as<float>(), pm.nmaps.as<float>(),
                                                                       pm.poses.as<float>(), K, pm.head, slots, kcap, hw, lo, hi,
                                                                       pm.zbuf.as<unsigned long long>(),
                                                                       pm.model_v.as<float>(), pm.model_n.as<float4>());
    PLS_CHECK_LAUNCH();
    pm.valid = true;
    pm.built_lo = lo;
    pm.built_hi = hi;
}

// Callers that read the whole model (export, the stand-alone search) on a rank that built only its share.
void ensure_full_model(pls_context* ctx) {
    const int64_t hw = (int64_t)ctx->cfg.height * ctx->cfg.width;
    if (ctx->pm.valid && (ctx->pm.built_lo != 0 || ctx->pm.built_hi != hw)) rebuild_model(ctx, true);
}

}  // namespace

void launch_normal_map(pls_context* ctx, const float* vmap, int batch, int H, int W, int ksize, float* out) {
    PLS_REQUIRE(ksize >= 1 && ksize <= 2 * NM_MAXR + 1 && (ksize & 1), "normal map: odd kernel size 1..9");
    dim3 grid((W + NM_TX - 1) / NM_TX, (H + NM_TY - 1) / NM_TY, batch), block(NM_TX, NM_TY);
    normal_map_kernel<<<grid, block, 0, ctx->stream>>>(vmap, batch, H, W, ksize, out);
    PLS_CHECK_LAUNCH();
}

void projmap_reset(pls_context* ctx) {
    ctx->pm.K = 0;
    ctx->pm.head = 0;
    ctx->pm.valid = false;
    ctx->pm.host_poses.clear();
}

// ProjectiveLocalMap.update (local_map.py:126-174): poses_k <- rel^-1 poses_k, append, evict, rebuild.
void projmap_update(pls_context* ctx, const float* rel_pose_host, const float* vmap_dev) {
    ProjMap& pm = ctx->pm;
    cudaStream_t st = ctx->stream;
    const int H = ctx->cfg.height, W = ctx->cfg.width;
    const int64_t hw = (int64_t)H * W;
    const int cap = ctx->cfg.local_map_size;
    const size_t frame_bytes = (size_t)3 * hw * sizeof(float);
    pm.vmaps.reserve((size_t)(cap + 1) * frame_bytes, st, true);
    pm.nmaps.reserve((size_t)(cap + 1) * frame_bytes, st, true);
    if (pm.K == 0) {
        PLS_REQUIRE(vmap_dev != nullptr, "projective map: the first update needs a vertex map");
        pm.host_poses.assign(rel_pose_host, rel_pose_host + 16);
        PLS_CUDA(cudaMemcpyAsync(pm.vmaps.p, vmap_dev, frame_bytes, cudaMemcpyDeviceToDevice, st));
        launch_normal_map(ctx, vmap_dev, 1, H, W, ctx->cfg.normals_kernel_size, pm.nmaps.as<float>());
        pm.K = 1;
    } else {
        float inv[16], tmp[16];
        rigid_inverse(rel_pose_host, inv);
        for (int k = 0; k < pm.K; ++k) {
            mat4_mul(inv, &pm.host_poses[16 * k], tmp);
            memcpy(&pm.host_poses[16 * k], tmp, sizeof(tmp));
        }
        const int slots = cap + 1;
        if (vmap_dev) {
            float eye[16];
            for (int i = 0; i < 16; ++i) eye[i] = (i % 5 == 0) ? 1.f : 0.f;
            pm.host_poses.insert(pm.host_poses.end(), eye, eye + 16);
            const size_t slot = (size_t)((pm.head + pm.K) % slots);  // ring of cap+1 frame slots
            char* vdst = reinterpret_cast<char*>(pm.vmaps.p) + slot * frame_bytes;
            char* ndst = reinterpret_cast<char*>(pm.nmaps.p) + slot * frame_bytes;
            PLS_CUDA(cudaMemcpyAsync(vdst, vmap_dev, frame_bytes, cudaMemcpyDeviceToDevice, st));
            launch_normal_map(ctx, vmap_dev, 1, H, W, ctx->cfg.normals_kernel_size, reinterpret_cast<float*>(ndst));
            pm.K += 1;
        }
        if (pm.K > cap) {  // drop the oldest frame: advance the ring head
            pm.head = (pm.head + 1) % slots;
            pm.host_poses.erase(pm.host_poses.begin(), pm.host_poses.begin() + 16);
            pm.K -= 1;
        }
    }
    rebuild_model(ctx);
}

// one ICP iteration on the projective map; pixels [rank*hw/R, (rank+1)*hw/R) are reduced by this rank
int projmap_icp_iteration(pls_context* ctx, int64_t query_bound, int rank, int num_ranks) {
    ProjMap& pm = ctx->pm;
    PLS_REQUIRE(pm.valid, "projective map: search before any update");
    cudaStream_t st = ctx->stream;
    const int H = ctx->cfg.height, W = ctx->cfg.width;
    const int64_t hw = (int64_t)H * W;
    FrameResult* fr = frame_result_dev(ctx);
    ctx->tmp[3].reserve((size_t)hw * sizeof(unsigned long long), st);
    unsigned long long* zbuf = ctx->tmp[3].as<unsigned long long>();
    // the TMA path's resolve kernel leaves the z-buffer cleared, so only the first iteration of a frame memsets
    if (!pm.zbuf_clean) PLS_CUDA(cudaMemsetAsync(zbuf, 0xff, (size_t)hw * sizeof(unsigned long long), st));
    pm.zbuf_clean = false;
    ProjConst pc = make_proj_const(H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg);
    const int K = pm.K;
    // split of the K candidates between the TMA path and the direct-load path (at most 10 direct)
    static const int kdirect_env = getenv("PLS_PROJ_KDIRECT") ? atoi(getenv("PLS_PROJ_KDIRECT")) : 4;
    int kdirect = kdirect_env < 0 ? 0 : (kdirect_env > 10 ? 10 : kdirect_env);
    if (kdirect > K - 1) kdirect = K > 1 ? K - 1 : 0;
    const int ktma = K - kdirect;
    const size_t stage_bytes = (size_t)(ktma * 3 + 4) * PT_TILE * sizeof(float);
    static const bool no_tma = getenv("PLS_PROJ_NO_TMA") != nullptr;
    static const int stages = getenv("PLS_PROJ_STAGES") ? atoi(getenv("PLS_PROJ_STAGES")) : 2;
    const bool use_tma = !no_tma && hw % PT_TILE == 0 && stages >= 1 && stages <= PT_STAGES && stage_bytes * stages <= 200 * 1024;
    // the pixels this rank reduces: whole 128-pixel tiles on the TMA path.  Only they take part in the z-buffer of the
    // transformed queries and in the target map, and only they need the model (a sharded rank builds just its share)
    const int64_t tiles = hw / PT_TILE;
    const int64_t tile_begin = tiles * rank / num_ranks, tile_end = tiles * (rank + 1) / num_ranks;
    const int64_t need_lo = use_tma ? tile_begin * PT_TILE : hw * rank / num_ranks;
    const int64_t need_hi = use_tma ? tile_end * PT_TILE : hw * (rank + 1) / num_ranks;
    if (need_lo < pm.built_lo || need_hi > pm.built_hi) rebuild_model(ctx, true);  // (the split rule changed since the update)
    query_zbuf_kernel<<<grid_for(query_bound, 256), 256, 0, st>>>(
        ctx->query_ptr, reinterpret_cast<const uint32_t*>(&fr->counts[1]), 0, fr->T, &fr->done, pc,
        use_tma ? (int)need_lo : 0, use_tma ? (int)need_hi : (int)hw, zbuf);
    PLS_CHECK_LAUNCH();
    int blocks;
    if (use_tma) {
        // TMA-staged persistent kernel over 128-pixel tiles; ranks take contiguous tile ranges
        const size_t smem = stage_bytes * stages;
        static bool attr_set = false;
        if (!attr_set) {
            PLS_CUDA(cudaFuncSetAttribute(proj_icp_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            attr_set = true;
        }
        int per_sm = (int)((220 * 1024) / (smem + 2048));
        per_sm = per_sm < 1 ? 1 : (per_sm > 8 ? 8 : per_sm);
        {   // equal tile counts per CTA: ceil(tiles / ceil(tiles / max_ctas)) persistent CTAs
            const int64_t my_tiles = tile_end - tile_begin;
            const int64_t max_ctas = (int64_t)per_sm * kNumSMs;
            const int64_t per_cta = (my_tiles + max_ctas - 1) / max_ctas;
            blocks = (int)((my_tiles + (per_cta > 0 ? per_cta : 1) - 1) / (per_cta > 0 ? per_cta : 1));
            if (blocks < 1) blocks = 1;
        }
        ctx->partials.reserve((size_t)blocks * NACC * sizeof(double), st);
        ctx->tmp[7].reserve((size_t)hw * sizeof(float4), st);
        query_resolve_kernel<<<grid_for(need_hi - need_lo, 256), 256, 0, st>>>(zbuf, ctx->query_ptr, fr->T, &fr->done, need_lo, need_hi,
                                                                               ctx->tmp[7].as<float4>());
        PLS_CHECK_LAUNCH();
        // how much of this rank's share of the model asks to stay in the 126 MB L2 between iterations
        static const int resident_mb = getenv("PLS_PROJ_RESIDENT_MB") ? atoi(getenv("PLS_PROJ_RESIDENT_MB")) : 56;
        const int64_t tile_bytes = (int64_t)ctx->cfg.local_map_size * 3 * PT_TILE * sizeof(float);
        const int64_t resident_end = tile_begin + ((int64_t)resident_mb << 20) / tile_bytes;
        // (a persisting access-policy window on the stream -- a set-aside part of L2 -- was tried instead of the
        // per-instruction priorities and lost: 38.5 us vs 33.2 us per launch at cfg5, profiles/r2_proj_l2_residency.md)
        ProfileScope ps(ctx, 1, 0.0, false);
        proj_icp_tma_kernel<<<blocks, PT_TILE, smem, st>>>(pm.model_v.as<float>(), pm.model_n.as<float4>(), K,
                                                           ctx->cfg.local_map_size, ctx->tmp[7].as<float4>(), fr, tile_begin,
                                                           tile_end, ctx->cfg.scheme, ctx->cfg.sigma, stages, ktma, resident_end,
                                                           L2_EVICT_LAST, L2_EVICT_FIRST, ctx->partials.as<double>());
        PLS_CHECK_LAUNCH();
        pm.zbuf_clean = true;
        return blocks;
    }
    const int64_t pix_begin = hw * rank / num_ranks, pix_end = hw * (rank + 1) / num_ranks;
    blocks = grid_for(pix_end - pix_begin, PJ_THREADS, 4 * kNumSMs);
    ctx->partials.reserve((size_t)blocks * NACC * sizeof(double), st);
    {
        ProfileScope ps(ctx, 1, 0.0, false);
        proj_icp_iter_kernel<<<blocks, PJ_THREADS, 0, st>>>(pm.model_v.as<float>(), pm.model_n.as<float4>(), pm.K,
                                                            ctx->cfg.local_map_size, zbuf,
                                                            ctx->query_ptr, fr, pix_begin, pix_end, ctx->cfg.scheme,
                                                            ctx->cfg.sigma, ctx->partials.as<double>());
        PLS_CHECK_LAUNCH();
    }
    return blocks;
}

}  // namespace pls

using namespace pls;

extern "C" {

int pls_normal_map(pls_context* ctx, const float* vertex_map, int batch, int height, int width, int kernel_size, float* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(vertex_map && out && batch > 0 && height > 0 && width > 0, "pls_normal_map: bad arguments");
    const size_t bytes = (size_t)batch * 3 * height * width * sizeof(float);
    const float* d = (const float*)to_device(ctx, vertex_map, bytes, ctx->stage_in[0]);
    OutArg o = out_arg(ctx, out, bytes, ctx->stage_out[0]);
    launch_normal_map(ctx, d, batch, height, width, kernel_size, (float*)o.dev);
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_compute_neighbors(pls_context* ctx, const float* tgt, const float* ref, const float* fields, int num_ref,
                          int num_field_channels, int height, int width, float* out_nb, float* out_fields) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(tgt && ref && out_nb && num_ref > 0 && height > 0 && width > 0, "pls_compute_neighbors: bad arguments");
    PLS_REQUIRE(!fields || (out_fields && num_field_channels > 0), "pls_compute_neighbors: fields need an output");
    const int64_t hw = (int64_t)height * width;
    const float* d_t = (const float*)to_device(ctx, tgt, (size_t)3 * hw * sizeof(float), ctx->stage_in[0]);
    const float* d_r = (const float*)to_device(ctx, ref, (size_t)num_ref * 3 * hw * sizeof(float), ctx->stage_in[1]);
    const float* d_f = (const float*)to_device(ctx, fields, (size_t)num_ref * num_field_channels * hw * sizeof(float), ctx->stage_in[2]);
    OutArg onb = out_arg(ctx, out_nb, (size_t)3 * hw * sizeof(float), ctx->stage_out[0]);
    OutArg of = out_arg(ctx, fields ? out_fields : nullptr, (size_t)num_field_channels * hw * sizeof(float), ctx->stage_out[1]);
    compute_neighbors_kernel<<<grid_for(hw, 256), 256, 0, ctx->stream>>>(d_t, d_r, d_f, num_ref, num_field_channels, hw,
                                                                         (float*)onb.dev, (float*)of.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, onb);
    finish_out(ctx, of);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_projmap_update(pls_context* ctx, const float* rel_pose, const float* vertex_map) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(rel_pose, "pls_projmap_update: rel_pose required");
    PLS_REQUIRE(ctx->cfg.local_map_type == PLS_MAP_PROJECTIVE, "context holds a kd map");
    float rel[16];
    if (is_device_ptr(rel_pose)) PLS_CUDA(cudaMemcpy(rel, rel_pose, sizeof(rel), cudaMemcpyDeviceToHost));
    else memcpy(rel, rel_pose, sizeof(rel));
    const size_t bytes = (size_t)3 * ctx->cfg.height * ctx->cfg.width * sizeof(float);
    const float* d = (const float*)to_device(ctx, vertex_map, bytes, ctx->stage_in[0]);
    projmap_update(ctx, rel, d);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_projmap_num_frames(pls_context* ctx, int* num_frames) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(num_frames, "pls_projmap_num_frames: null output");
    *num_frames = ctx->pm.K;
    PLS_API_END(ctx)
}

int pls_projmap_model(pls_context* ctx, float* out_vmap, float* out_nmap) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(ctx->pm.valid, "pls_projmap_model: empty map");
    ensure_full_model(ctx);
    const int64_t hw = (int64_t)ctx->cfg.height * ctx->cfg.width;
    const size_t bytes = (size_t)ctx->pm.K * 3 * hw * sizeof(float);
    auto put = [&](float* dst, const float* tiled, DBuf& stage) {
        if (!dst) return;
        OutArg o = out_arg(ctx, dst, bytes, stage);
        model_export_kernel<<<grid_for((int64_t)ctx->pm.K * 3 * hw, 256), 256, 0, ctx->stream>>>(tiled, ctx->pm.K, ctx->cfg.local_map_size,
                                                                                              hw, (float*)o.dev);
        PLS_CHECK_LAUNCH();
        finish_out(ctx, o);
    };
    put(out_vmap, ctx->pm.model_v.as<float>(), ctx->stage_out[0]);
    if (out_nmap) {
        OutArg o = out_arg(ctx, out_nmap, bytes, ctx->stage_out[1]);
        normal_export_kernel<<<grid_for((int64_t)ctx->pm.K * hw, 256), 256, 0, ctx->stream>>>(ctx->pm.model_n.as<float4>(), ctx->pm.K,
                                                                                           ctx->cfg.local_map_size, hw, (float*)o.dev);
        PLS_CHECK_LAUNCH();
        finish_out(ctx, o);
    }
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_projmap_nn_search(pls_context* ctx, const float* queries, int64_t n, float* out_neighbors, float* out_normals,
                          float* out_targets, int64_t* out_count) {
    PLS_API_BEGIN(ctx)
    map_stream_wait(ctx);
    PLS_REQUIRE(queries && out_neighbors && out_count && n > 0, "pls_projmap_nn_search: bad arguments");
    if (!ctx->pm.valid) throw pls::Error{PLS_E_STATE, "pls_projmap_nn_search: the map is empty"};
    cudaStream_t st = ctx->stream;
    const int H = ctx->cfg.height, W = ctx->cfg.width;
    const int64_t hw = (int64_t)H * W;
    const float* d = (const float*)to_device(ctx, queries, (size_t)n * 3 * sizeof(float), ctx->stage_in[0]);
    // queries as float4 (no NaN filtering here: the reference passes them through unchanged)
    ctx->queries.reserve((size_t)n * sizeof(float4), st);
    uint32_t* cnt = scalar_u32(ctx, SC_NAN_COUNT);
    pack_valid_rows(ctx, d, n, ctx->queries.as<float4>(), cnt);
    ctx->tmp[3].reserve((size_t)hw * sizeof(unsigned long long), st);
    unsigned long long* zbuf = ctx->tmp[3].as<unsigned long long>();
    PLS_CUDA(cudaMemsetAsync(zbuf, 0xff, (size_t)hw * sizeof(unsigned long long), st));
    ProjConst pc = make_proj_const(H, W, ctx->cfg.up_fov_deg, ctx->cfg.down_fov_deg);
    ensure_full_model(ctx);
    query_zbuf_kernel<<<grid_for(n, 256), 256, 0, st>>>(ctx->queries.as<float4>(), cnt, 0, nullptr, nullptr, pc, 0, (int)hw, zbuf);
    PLS_CHECK_LAUNCH();
    ctx->tmp[1].reserve((size_t)hw, st);
    ctx->tmp[2].reserve((size_t)hw * sizeof(uint32_t), st);
    ctx->tmp[7].reserve((size_t)hw * 9 * sizeof(float), st);
    proj_pairs_kernel<<<grid_for(hw, 256), 256, 0, st>>>(ctx->pm.model_v.as<float>(), ctx->pm.model_n.as<float4>(), ctx->pm.K,
                                                          ctx->cfg.local_map_size, hw, zbuf, ctx->queries.as<float4>(), ctx->tmp[1].as<uint8_t>(),
                                                          ctx->tmp[7].as<float>());
    PLS_CHECK_LAUNCH();
    uint32_t* total = scalar_u32(ctx, SC_PROJ_NC);
    exclusive_scan_flags(ctx, ctx->tmp[1].as<uint8_t>(), hw, ctx->tmp[2].as<uint32_t>(), total);
    OutArg oq = out_arg(ctx, out_neighbors, (size_t)hw * 3 * sizeof(float), ctx->stage_out[0]);
    OutArg on = out_arg(ctx, out_normals, (size_t)hw * 3 * sizeof(float), ctx->stage_out[1]);
    OutArg op = out_arg(ctx, out_targets, (size_t)hw * 3 * sizeof(float), ctx->stage_out[2]);
    proj_compact_kernel<<<grid_for(hw, 256), 256, 0, st>>>(ctx->tmp[1].as<uint8_t>(), ctx->tmp[2].as<uint32_t>(),
                                                            ctx->tmp[7].as<float>(), hw, (float*)oq.dev, (float*)on.dev,
                                                            (float*)op.dev);
    PLS_CHECK_LAUNCH();
    uint32_t nc = 0;
    PLS_CUDA(cudaMemcpyAsync(&nc, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    PLS_CUDA(cudaStreamSynchronize(st));
    *out_count = nc;
    finish_out(ctx, oq, (size_t)nc * 3 * sizeof(float));
    finish_out(ctx, on, (size_t)nc * 3 * sizeof(float));
    finish_out(ctx, op, (size_t)nc * 3 * sizeof(float));
    PLS_CUDA(cudaStreamSynchronize(st));
    PLS_API_END(ctx)
}

}  // extern "C"
