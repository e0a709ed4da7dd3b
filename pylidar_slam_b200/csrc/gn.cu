// K6/K7 -- point-to-plane (and, behind the same registry, point-to-point) Gauss-Newton alignment on given
// correspondences.
//
//   gn_accumulate_kernel : residual r = n.(R(x) p + t(x) - q), Jacobian row
//                          J = [n, (dR/de_k p).n], robust weight w, and the reduction of the
//                          30 normal-equation accumulators (21 upper JtWJ, 6 JtWr, sum (w r)^2,
//                          sum r^2, count): per-thread fp64 accumulation of fp32 (or fp64) terms,
//                          warp-shuffle then shared-memory block reduction, one partial row per
//                          block (deterministic two-stage sum, no atomics).
//   gn_solve_kernel      : sums the block partials in fixed order, applies the reference's two
//                          guards (|r| < 1e-7 -> warn/stop, |det H| < 1e-7 -> error), solves the
//                          6x6 system and updates x.
//
// Replaces PointToPlaneCost.get_residual_fun / get_residual_jac_fun
// (slam/common/optimization.py:356-435), _WLSScheme.weights + the seven cost functions
// (:45-50,61-226), GaussNewton.compute (:296-344) and GaussNewtonPointToPlaneAlignment.align
// (slam/odometry/alignment.py:91-127).  COST_POINT swaps in PointToPointCost's closures (optimization.py:458-541)
// for GaussNewtonPointToPointAlignment.align (alignment.py:144-189); everything after the residual/Jacobian row is
// shared.
#include "gn_device.cuh"
#include "internal.cuh"
#include "pose_device.cuh"

namespace pls {

namespace {

template <typename T>
struct GnState {
    T x[6];
    T dT[16];
    double sums[NACC];
    double dx_norm;
    int done;
    int status;
    int iters;
    int pad;
};

constexpr int GN_THREADS = 256;
enum { COST_PLANE = 0, COST_POINT = 1 };

template <typename T, int COST>
__global__ void __launch_bounds__(GN_THREADS)
gn_accumulate_kernel(const T* __restrict__ ref, const T* __restrict__ tgt, const T* __restrict__ nrm, int64_t n,
                     const GnState<T>* __restrict__ state, int scheme, T sigma, T* __restrict__ loss_out,
                     double* __restrict__ partials) {
    if (state->done) return;
    __shared__ T sR[9], st[3], sdR[27];
    if (threadIdx.x == 0) {
        T M[16];
        build_pose(state->x, M);
        sR[0] = M[0]; sR[1] = M[1]; sR[2] = M[2];
        sR[3] = M[4]; sR[4] = M[5]; sR[5] = M[6];
        sR[6] = M[8]; sR[7] = M[9]; sR[8] = M[10];
        st[0] = M[3]; st[1] = M[7]; st[2] = M[11];
        euler_jacobian(state->x + 3, sdR);
    }
    __syncthreads();
    double acc[NACC];
#pragma unroll
    for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T p[3] = {tgt[3 * i], tgt[3 * i + 1], tgt[3 * i + 2]};
        T q[3] = {ref[3 * i], ref[3 * i + 1], ref[3 * i + 2]};
        T J[6];
        T r;
        if constexpr (COST == COST_PLANE) {
            T nn[3] = {nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]};
            r = p2plane_residual_jacobian<T>(p, q, nn, sR, st, sdR, J);
        } else {
            r = p2point_residual_jacobian<T>(p, q, sR, st, sdR, J);
        }
        T w = ls_weight<T>(scheme, sigma, r, p, q);
        T wr = r * w;
        if (loss_out) loss_out[i] = wr * wr;
        accumulate_normal_equations<T>(acc, J, w, wr, r);
    }
    block_reduce_store<GN_THREADS>(acc, partials + (size_t)blockIdx.x * NACC);
}

template <typename T>
__global__ void gn_solve_kernel(GnState<T>* state, const double* __restrict__ partials, int num_blocks,
                                T norm_stop) {
    if (state->done) return;
    __shared__ double sums[NACC];
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int a = warp; a < NACC; a += 8) {
            double s = 0.0;
            for (int b = lane; b < num_blocks; b += 32) s += partials[(size_t)b * NACC + a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) { sums[a] = s; state->sums[a] = s; }
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    state->iters += 1;
    // optimization.py:323-327 -- tiny residual norm: warn and return x unchanged
    if (sqrt(sums[28]) < 1e-7) {
        state->status = PLS_W_TINY_RESIDUAL;
        state->done = 1;
        build_pose(state->x, state->dT);
        return;
    }
    double dx[6];
    double det = solve6(sums, dx);
    // optimization.py:334-336 -- singular normal equations: raise
    if (!(fabs(det) >= 1e-7)) {
        state->status = PLS_E_SINGULAR;
        state->done = 1;
        return;
    }
    double nrm2 = 0.0;
    for (int i = 0; i < 6; ++i) {
        T d = (T)dx[i];
        state->x[i] = state->x[i] + d;
        nrm2 += (double)d * (double)d;
    }
    state->dx_norm = sqrt(nrm2);
    build_pose(state->x, state->dT);
    if (state->dx_norm < (double)norm_stop) state->done = 1;
}

template <typename T, int COST>
void align_impl(pls_context* ctx, const void* ref, const void* tgt, const void* nrm, int64_t n, int scheme,
                double sigma, int max_iters, double norm_stop, const void* x0, void* out_dT, void* out_x,
                void* out_loss, int* status_out) {
    cudaStream_t st = ctx->stream;
    const size_t pts_bytes = (size_t)n * 3 * sizeof(T);
    const T* d_ref = (const T*)to_device(ctx, ref, pts_bytes, ctx->stage_in[0]);
    const T* d_tgt = (const T*)to_device(ctx, tgt, pts_bytes, ctx->stage_in[1]);
    const T* d_nrm = COST == COST_PLANE ? (const T*)to_device(ctx, nrm, pts_bytes, ctx->stage_in[2]) : nullptr;
    OutArg o_loss = out_arg(ctx, out_loss, (size_t)n * sizeof(T), ctx->stage_out[0]);

    ctx->tmp[0].reserve(sizeof(GnState<T>), st);
    GnState<T>* d_state = ctx->tmp[0].as<GnState<T>>();
    GnState<T> h_state;
    memset(&h_state, 0, sizeof(h_state));
    if (x0) {
        if (is_device_ptr(x0)) PLS_CUDA(cudaMemcpy(h_state.x, x0, 6 * sizeof(T), cudaMemcpyDeviceToHost));
        else memcpy(h_state.x, x0, 6 * sizeof(T));
    }
    build_pose(h_state.x, h_state.dT);
    PLS_CUDA(cudaMemcpyAsync(d_state, &h_state, sizeof(h_state), cudaMemcpyHostToDevice, st));

    int blocks = (int)((n + GN_THREADS - 1) / GN_THREADS);
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    if (blocks < 1) blocks = 1;
    ctx->partials.reserve((size_t)blocks * NACC * sizeof(double), st);
    int iters = max_iters < 1 ? 1 : max_iters;
    for (int it = 0; it < iters; ++it) {
        {
            ProfileScope ps(ctx, 5, (double)n * (COST == COST_PLANE ? 9 : 6) * sizeof(T) + NACC * 8.0);
            gn_accumulate_kernel<T, COST><<<blocks, GN_THREADS, 0, st>>>(d_ref, d_tgt, d_nrm, n, d_state, scheme, (T)sigma,
                                                                  (T*)o_loss.dev, ctx->partials.as<double>());
            PLS_CHECK_LAUNCH();
        }
        gn_solve_kernel<T><<<1, 256, 0, st>>>(d_state, ctx->partials.as<double>(), blocks, (T)norm_stop);
        PLS_CHECK_LAUNCH();
    }
    PLS_CUDA(cudaMemcpyAsync(&h_state, d_state, sizeof(h_state), cudaMemcpyDeviceToHost, st));
    finish_out(ctx, o_loss);
    PLS_CUDA(cudaStreamSynchronize(st));
    auto put = [&](void* dst, const void* src, size_t bytes) {
        if (!dst) return;
        if (is_device_ptr(dst)) PLS_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
        else memcpy(dst, src, bytes);
    };
    put(out_dT, h_state.dT, 16 * sizeof(T));
    put(out_x, h_state.x, 6 * sizeof(T));
    *status_out = h_state.status;
}

__global__ void pose_build_kernel(const float* params, int batch, float* out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) build_pose(params + 6 * b, out + 16 * b);
}
__global__ void pose_from_kernel(const float* mats, int batch, float* out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) from_pose(mats + 16 * b, out + 6 * b);
}

}  // namespace
}  // namespace pls

using namespace pls;

extern "C" {

int pls_align_p2plane(pls_context* ctx, const void* ref, const void* tgt, const void* nrm, int64_t n, int is_f64,
                      int scheme, double sigma, int max_iters, double norm_stop, const void* x0, void* out_dT,
                      void* out_x, void* out_loss) {
    int status = PLS_OK;
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(ref && tgt && nrm && n > 0, "pls_align_p2plane: ref/tgt/nrm must be [n,3] with n > 0");
    PLS_REQUIRE(scheme >= 0 && scheme <= PLS_SCHEME_CAUCHY, "pls_align_p2plane: unknown weighting scheme");
    if (is_f64)
        align_impl<double, COST_PLANE>(ctx, ref, tgt, nrm, n, scheme, sigma, max_iters, norm_stop, x0, out_dT, out_x, out_loss, &status);
    else
        align_impl<float, COST_PLANE>(ctx, ref, tgt, nrm, n, scheme, sigma, max_iters, norm_stop, x0, out_dT, out_x, out_loss, &status);
    if (status == PLS_E_SINGULAR) throw pls::Error{PLS_E_SINGULAR, "Invalid Jacobian in Gauss Newton minimization"};
    if (status == PLS_W_TINY_RESIDUAL) {
        ctx->err = "The residual norm is lower than threshold 1e-7";
        return PLS_W_TINY_RESIDUAL;
    }
    PLS_API_END(ctx)
}

int pls_align_p2point(pls_context* ctx, const void* ref, const void* tgt, int64_t n, int is_f64, int scheme, double sigma,
                      int max_iters, double norm_stop, const void* x0, void* out_dT, void* out_x, void* out_loss) {
    int status = PLS_OK;
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(ref && tgt && n > 0, "pls_align_p2point: ref/tgt must be [n,3] with n > 0");
    PLS_REQUIRE(scheme >= 0 && scheme <= PLS_SCHEME_CAUCHY, "pls_align_p2point: unknown weighting scheme");
    if (is_f64)
        align_impl<double, COST_POINT>(ctx, ref, tgt, nullptr, n, scheme, sigma, max_iters, norm_stop, x0, out_dT, out_x, out_loss, &status);
    else
        align_impl<float, COST_POINT>(ctx, ref, tgt, nullptr, n, scheme, sigma, max_iters, norm_stop, x0, out_dT, out_x, out_loss, &status);
    if (status == PLS_E_SINGULAR) throw pls::Error{PLS_E_SINGULAR, "Invalid Jacobian in Gauss Newton minimization"};
    if (status == PLS_W_TINY_RESIDUAL) {
        ctx->err = "The residual norm is lower than threshold 1e-7";
        return PLS_W_TINY_RESIDUAL;
    }
    PLS_API_END(ctx)
}

int pls_build_pose_matrix(pls_context* ctx, const float* params, int batch, float* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(params && out && batch > 0, "pls_build_pose_matrix: bad arguments");
    const float* d_in = (const float*)to_device(ctx, params, (size_t)batch * 6 * sizeof(float), ctx->stage_in[0]);
    OutArg o = out_arg(ctx, out, (size_t)batch * 16 * sizeof(float), ctx->stage_out[0]);
    pose_build_kernel<<<(batch + 63) / 64, 64, 0, ctx->stream>>>(d_in, batch, (float*)o.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_from_pose_matrix(pls_context* ctx, const float* mats, int batch, float* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(mats && out && batch > 0, "pls_from_pose_matrix: bad arguments");
    const float* d_in = (const float*)to_device(ctx, mats, (size_t)batch * 16 * sizeof(float), ctx->stage_in[0]);
    OutArg o = out_arg(ctx, out, (size_t)batch * 6 * sizeof(float), ctx->stage_out[0]);
    pose_from_kernel<<<(batch + 63) / 64, 64, 0, ctx->stream>>>(d_in, batch, (float*)o.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

}  // extern "C"
