// The rows either side of the path, rank 4 of SURVEY.md section 8f: dataset -> vertex-map ingestion and pose chains.
//
//   kitti_correct_kernel     KITTIOdometrySequence.correct_scan (slam/dataset/kitti_dataset.py:200-231): one thread per
//                            point, [n, stride] float32 scan rows (x, y, z[, reflectance]) -> float64 [n,3]
//   pls_ingest_scan          ... followed by the float64 spherical projection + closest-wins z-buffer of
//                            KITTIOdometrySequence.__getitem__ (:241-249), i.e. the `numpy_pc` and the vertex map a
//                            DataLoader worker produces on the CPU in the reference
//   relative_poses_kernel    compute_relative_poses (slam/eval/eval_odometry.py:80-83): inv(pose[i-1]) @ pose[i]
//   absolute_poses_kernel    compute_absolute_poses (:86-96): the running product, sequential like the reference (one warp:
//                            lane (i, j) carries entry (i, j) of the accumulated pose)
#include "ingest_device.cuh"
#include "internal.cuh"

namespace pls {

namespace {

__global__ void kitti_correct_kernel(const float* __restrict__ scan, int stride, int64_t n, double c, double s,
                                     double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double o[3];
        kitti_correct_point(scan[stride * i], scan[stride * i + 1], scan[stride * i + 2], c, s, o);
        out[3 * i] = o[0];
        out[3 * i + 1] = o[1];
        out[3 * i + 2] = o[2];
    }
}

// an uncorrected scan is only widened to float64 (raw-lidar branch of __getitem__, kitti_dataset.py:251-259)
__global__ void widen_scan_kernel(const float* __restrict__ scan, int stride, int64_t n, double* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        out[3 * i] = (double)scan[stride * i];
        out[3 * i + 1] = (double)scan[stride * i + 1];
        out[3 * i + 2] = (double)scan[stride * i + 2];
    }
}

template <typename T>
__global__ void relative_poses_kernel(const T* __restrict__ poses, int64_t n, T* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        T prev[16], inv[16], cur[16], r[16];
        for (int k = 0; k < 16; ++k) {
            prev[k] = i > 0 ? poses[16 * (i - 1) + k] : ((k % 5 == 0) ? (T)1 : (T)0);  // shift_poses: identity first
            cur[k] = poses[16 * i + k];
        }
        inverse4<T>(prev, inv);
        matmul4<T>(inv, cur, r);
        for (int k = 0; k < 16; ++k) out[16 * i + k] = r[k];
    }
}

// absolute[0] = relative[0]; absolute[i + 1] = absolute[i] @ relative[i + 1]
template <typename T>
__global__ void absolute_poses_kernel(const T* __restrict__ rel, int64_t n, T* __restrict__ out) {
    const int lane = threadIdx.x;  // 16 working lanes: entry (lane / 4, lane % 4)
    const int i = (lane >> 2) & 3, j = lane & 3;
    T acc = lane < 16 ? rel[lane] : (T)0;
    if (lane < 16) out[lane] = acc;
    for (int64_t k = 1; k < n; ++k) {
        T s = 0;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const T a = __shfl_sync(0xffffffffu, acc, 4 * i + t);
            const T b = lane < 16 ? rel[16 * k + 4 * t + j] : (T)0;
            s += a * b;
        }
        acc = s;
        if (lane < 16) out[16 * k + lane] = acc;
    }
}

inline int grid_for(int64_t n, int threads = 256) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = 8 * kNumSMs;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

// projection.cu: float64 cloud -> float64 vertex map [3,H,W]
void launch_projection_f64_out64(pls_context* ctx, const double* xyz, int64_t n, int H, int W, float up, float down, double* out,
                                 unsigned long long* zbuf);

}  // namespace pls

using namespace pls;

extern "C" {

int pls_kitti_correct_scan(pls_context* ctx, const float* scan, int64_t n, int stride, double* out_xyz) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(scan && out_xyz && n > 0 && (stride == 3 || stride == 4), "pls_kitti_correct_scan: scan is [n,3] or [n,4] float32");
    const float* d = (const float*)to_device(ctx, scan, (size_t)n * stride * sizeof(float), ctx->stage_in[0]);
    OutArg o = out_arg(ctx, out_xyz, (size_t)n * 3 * sizeof(double), ctx->stage_out[0]);
    const double theta = 0.205 * 3.141592653589793 / 180.0;  // kitti_dataset.py:212
    kitti_correct_kernel<<<grid_for(n), 256, 0, ctx->stream>>>(d, stride, n, cos(theta), sin(theta), (double*)o.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_ingest_scan(pls_context* ctx, const float* scan, int64_t n, int stride, int correct, int height, int width,
                    float up_fov_deg, float down_fov_deg, double* out_xyz, double* out_vertex_map) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(scan && out_vertex_map && n > 0 && (stride == 3 || stride == 4) && height > 0 && width > 0,
                "pls_ingest_scan: bad arguments");
    cudaStream_t st = ctx->stream;
    const int64_t hw = (int64_t)height * width;
    const float* d = (const float*)to_device(ctx, scan, (size_t)n * stride * sizeof(float), ctx->stage_in[0]);
    OutArg ox = out_arg(ctx, out_xyz, (size_t)n * 3 * sizeof(double), ctx->stage_out[0]);
    DBuf& xyz_buf = ctx->next_buf[0];
    double* xyz = (double*)ox.dev;
    if (!xyz) {  // the caller does not want the cloud itself
        xyz_buf.reserve((size_t)n * 3 * sizeof(double), st);
        xyz = xyz_buf.as<double>();
    }
    const double theta = 0.205 * 3.141592653589793 / 180.0;
    if (correct) kitti_correct_kernel<<<grid_for(n), 256, 0, st>>>(d, stride, n, cos(theta), sin(theta), xyz);
    else widen_scan_kernel<<<grid_for(n), 256, 0, st>>>(d, stride, n, xyz);
    PLS_CHECK_LAUNCH();
    OutArg ov = out_arg(ctx, out_vertex_map, (size_t)3 * hw * sizeof(double), ctx->stage_out[1]);
    ctx->tmp[3].reserve((size_t)hw * sizeof(unsigned long long), st);
    launch_projection_f64_out64(ctx, xyz, n, height, width, up_fov_deg, down_fov_deg, (double*)ov.dev,
                                ctx->tmp[3].as<unsigned long long>());
    finish_out(ctx, ox);
    finish_out(ctx, ov);
    PLS_CUDA(cudaStreamSynchronize(st));
    PLS_API_END(ctx)
}

int pls_relative_poses(pls_context* ctx, const void* poses, int64_t n, int is_f64, void* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(poses && out && n > 0, "pls_relative_poses: poses is [n,4,4]");
    const size_t bytes = (size_t)n * 16 * (is_f64 ? sizeof(double) : sizeof(float));
    const void* d = to_device(ctx, poses, bytes, ctx->stage_in[0]);
    OutArg o = out_arg(ctx, out, bytes, ctx->stage_out[0]);
    if (is_f64) relative_poses_kernel<double><<<grid_for(n, 64), 64, 0, ctx->stream>>>((const double*)d, n, (double*)o.dev);
    else relative_poses_kernel<float><<<grid_for(n, 64), 64, 0, ctx->stream>>>((const float*)d, n, (float*)o.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

int pls_absolute_poses(pls_context* ctx, const void* relative_poses, int64_t n, int is_f64, void* out) {
    PLS_API_BEGIN(ctx)
    PLS_REQUIRE(relative_poses && out && n > 0, "pls_absolute_poses: relative_poses is [n,4,4]");
    const size_t bytes = (size_t)n * 16 * (is_f64 ? sizeof(double) : sizeof(float));
    const void* d = to_device(ctx, relative_poses, bytes, ctx->stage_in[0]);
    OutArg o = out_arg(ctx, out, bytes, ctx->stage_out[0]);
    if (is_f64) absolute_poses_kernel<double><<<1, 32, 0, ctx->stream>>>((const double*)d, n, (double*)o.dev);
    else absolute_poses_kernel<float><<<1, 32, 0, ctx->stream>>>((const float*)d, n, (float*)o.dev);
    PLS_CHECK_LAUNCH();
    finish_out(ctx, o);
    PLS_CUDA(cudaStreamSynchronize(ctx->stream));
    PLS_API_END(ctx)
}

}  // extern "C"
