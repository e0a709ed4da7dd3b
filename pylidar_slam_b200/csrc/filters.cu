// Motion de-skew of a frame -- the filter that precedes GridSample in the shipped preprocessing chain
// (config/slam/preprocessing/grid_sample.yaml).
//
// Replaces Distortion.filter (slam/preprocessing.py:148-191):
//   alpha_i = (t_i - min t) / (max t - min t)            in the timestamps' dtype (all zeros when max == min)
//   R_i     = Slerp(identity -> R)(alpha_i) = exp(alpha_i log R)     float64
//   out_i   = R_i p_i + alpha_i t                                     float64 [n,3]
//
//   ts_minmax_kernel : grid-stride min / max of the timestamps (NaN-propagating like np.min / np.max), one
//                      partial pair per block (deterministic two-stage reduction, no atomics).
//   distort_kernel   : every block folds the <= 1184 partial pairs in its prologue (cheaper than a third launch),
//                      then per point: alpha, Rodrigues rotation about the fixed axis of log R by alpha * angle
//                      (one sincos), rotate, translate, three coalesced float64 stores.
//
// HBM-bound: 12 (xyz) + 4|8 (t, read twice) + 24 (out) bytes per point; nothing is re-read except the timestamps.
#include "filters_device.cuh"
#include "internal.cuh"

namespace pls {

namespace {

constexpr int DS_THREADS = 256;

template <typename TS>
__global__ void __launch_bounds__(DS_THREADS)
ts_minmax_kernel(const TS* __restrict__ ts, int64_t n, double* __restrict__ partials /*[blocks][3]: min, max, nan*/) {
    double mn = INFINITY, mx = -INFINITY, bad = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double t = (double)ts[i];  // exact for float32 timestamps
        if (t != t) bad = 1.0;
        mn = fmin(mn, t);
        mx = fmax(mx, t);
    }
    __shared__ double s_mn[DS_THREADS / 32], s_mx[DS_THREADS / 32], s_bad[DS_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad = fmax(bad, __shfl_xor_sync(0xffffffffu, bad, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < DS_THREADS / 32; ++w) {
            mn = fmin(mn, s_mn[w]);
            mx = fmax(mx, s_mx[w]);
            bad = fmax(bad, s_bad[w]);
        }
        partials[3 * blockIdx.x] = mn;
        partials[3 * blockIdx.x + 1] = mx;
        partials[3 * blockIdx.x + 2] = bad;
    }
}

template <typename TP, typename TS>
__global__ void __launch_bounds__(DS_THREADS)
distort_kernel(const TP* __restrict__ pc, const TS* __restrict__ ts, int64_t n, const double* __restrict__ partials,
               int num_partials, DistortParams prm, double* __restrict__ out) {
    __shared__ double s_red[3][DS_THREADS / 32];
    __shared__ TS s_min, s_den;
    __shared__ int s_mode;  // 0 = regular, 1 = max == min (alpha = t * 0), 2 = NaN among the timestamps
    {
        double mn = INFINITY, mx = -INFINITY, bad = 0.0;
        for (int b = threadIdx.x; b < num_partials; b += DS_THREADS) {
            mn = fmin(mn, partials[3 * b]);
            mx = fmax(mx, partials[3 * b + 1]);
            bad = fmax(bad, partials[3 * b + 2]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            bad = fmax(bad, __shfl_xor_sync(0xffffffffu, bad, o));
        }
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { s_red[0][warp] = mn; s_red[1][warp] = mx; s_red[2][warp] = bad; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < DS_THREADS / 32; ++w) {
                mn = fmin(mn, s_red[0][w]);
                mx = fmax(mx, s_red[1][w]);
                bad = fmax(bad, s_red[2][w]);
            }
            // preprocessing.py:180-182 -- the subtraction happens in the timestamps' dtype
            const TS tmn = (TS)mn, tmx = (TS)mx;
            const TS diff = tmx - tmn;
            s_min = tmn;
            s_den = diff;
            s_mode = bad != 0.0 ? 2 : (diff == (TS)0 ? 1 : 0);
        }
        __syncthreads();
    }
    const TS tmin = s_min, den = s_den;
    const int mode = s_mode;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const TS t = ts[i];
        TS a;
        if (mode == 0) a = (t - tmin) / den;
        else if (mode == 1) a = t * (TS)0;
        else a = (TS)NAN;  // np.max / np.min propagate NaN: diff, and so every alpha, is NaN
        double ox, oy, oz;
        distort_point<TS>(a, prm, (double)pc[3 * i], (double)pc[3 * i + 1], (double)pc[3 * i + 2], ox, oy, oz);
        out[3 * i] = ox;
        out[3 * i + 1] = oy;
        out[3 * i + 2] = oz;
    }
}

inline int blocks_for(int64_t n) {
    int64_t b = (n + DS_THREADS - 1) / DS_THREADS;
    const int64_t cap = 8 * kNumSMs;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

template <typename TP, typename TS>
void distort_impl(pls_context* ctx, const void* d_pc, const void* d_ts, int64_t n, const DistortParams& prm, double* out) {
    cudaStream_t st = ctx->stream;
    const int blocks = blocks_for(n);
    ctx->next_buf[0].reserve((size_t)blocks * 3 * sizeof(double), st);
    double* partials = ctx->next_buf[0].as<double>();
    ts_minmax_kernel<TS><<<blocks, DS_THREADS, 0, st>>>((const TS*)d_ts, n, partials);
    PLS_CHECK_LAUNCH();
    distort_kernel<TP, TS><<<blocks, DS_THREADS, 0, st>>>((const TP*)d_pc, (const TS*)d_ts, n, partials, blocks, prm, out);
    PLS_CHECK_LAUNCH();
}

}  // namespace
}  // namespace pls

using namespace pls;

extern "C" {

int pls_distort(pls_context* ctx, const void* xyz, int xyz_is_f64, const void* timestamps, int ts_is_f64, int64_t n,
                const void* rel_pose, int pose_is_f64, double* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(xyz && timestamps && rel_pose && out && n > 0, "pls_distort: need [n,3] points, [n] timestamps, a 4x4 pose");
    double P[16];
    {
        unsigned char raw[16 * sizeof(double)];
        const size_t pb = 16 * (pose_is_f64 ? sizeof(double) : sizeof(float));
        if (is_device_ptr(rel_pose)) PLS_CUDA(cudaMemcpy(raw, rel_pose, pb, cudaMemcpyDeviceToHost));
        else memcpy(raw, rel_pose, pb);
        for (int i = 0; i < 16; ++i)
            P[i] = pose_is_f64 ? reinterpret_cast<const double*>(raw)[i] : (double)reinterpret_cast<const float*>(raw)[i];
    }
    DistortParams prm;
    const double R[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
    rotation_vector(R, prm.axis, &prm.angle);
    prm.t[0] = P[3]; prm.t[1] = P[7]; prm.t[2] = P[11];
    prm.tr_f32 = (!ts_is_f64 && !pose_is_f64) ? 1 : 0;
    const void* d_pc = to_device(ctx, xyz, (size_t)n * 3 * (xyz_is_f64 ? sizeof(double) : sizeof(float)), ctx->stage_in[0]);
    const void* d_ts = to_device(ctx, timestamps, (size_t)n * (ts_is_f64 ? sizeof(double) : sizeof(float)), ctx->stage_in[1]);
    OutArg o = out_arg(ctx, out, (size_t)n * 3 * sizeof(double), ctx->stage_out[0]);
    if (xyz_is_f64) {
        if (ts_is_f64) distort_impl<double, double>(ctx, d_pc, d_ts, n, prm, (double*)o.dev);
        else distort_impl<double, float>(ctx, d_pc, d_ts, n, prm, (double*)o.dev);
    } else {
        if (ts_is_f64) distort_impl<float, double>(ctx, d_pc, d_ts, n, prm, (double*)o.dev);
        else distort_impl<float, float>(ctx, d_pc, d_ts, n, prm, (double*)o.dev);
    }
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

}  // extern "C"
