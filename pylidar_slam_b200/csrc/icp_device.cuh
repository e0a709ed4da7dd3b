// Device-side tail of an ICP iteration, shared by the stand-alone step kernels (odometry.cu) and the
// correspondence kernels that finish the iteration in their last block (kdmap.cu).
#pragma once
#include "internal.cuh"
#include "pose_device.cuh"

namespace pls {

// Deterministic parallel sum of the block partial rows (THREADS = 256 or 128 threads, all must call): slice w of
// eight sums the rows w, w+8, ... of every accumulator (lane = accumulator, so a row is one coalesced 240-byte read and
// the loads of a lane are independent), then the eight slices are added in fixed order -- the same additions in the
// same order whatever the block size (a 128-thread block gives two slices to each warp).
template <int THREADS = 256>
__device__ __forceinline__ void sum_partials_256(const double* __restrict__ partials, int num_blocks, double* sums) {
    static_assert(THREADS == 256 || THREADS == 128, "sum_partials: 4 or 8 warps");
    __shared__ double slice_sum[8][NACC];
    const int a = threadIdx.x & 31;
    if (a < NACC) {
        for (int slice = threadIdx.x >> 5; slice < 8; slice += THREADS / 32) {
            // sixteen rows in flight per lane (the additions keep their order; a missing row adds +0.0): the sum of a
            // few hundred rows is a chain of L2 round trips otherwise
            double s = 0.0;
            for (int b0 = slice; b0 < num_blocks; b0 += 8 * 16) {
                double v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int b = b0 + 8 * j;
                    v[j] = b < num_blocks ? __ldcg(partials + (size_t)b * NACC + a) : 0.0;
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) s += v[j];
            }
            slice_sum[slice][a] = s;
        }
    }
    __syncthreads();
    if (threadIdx.x < NACC) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += slice_sum[k][threadIdx.x];
        sums[threadIdx.x] = s;
    }
}

// The serial tail of an ICP iteration (thread 0): Gauss-Newton guards, 6x6 solve, stop test, pose update.
__device__ __forceinline__ void icp_solve_and_update(FrameResult* fr, const double* sums, float threshold_delta) {
    const int it = fr->iters;
    fr->iters = it + 1;
    // optimization.py:323-327: |r| < 1e-7 -> warning, x stays 0, the residuals are r^2.  The ICP loop then sees delta = 0:
    // it breaks if 0 < threshold_delta_pose (icp_odometry.py:292) and otherwise goes on, the pose unchanged
    // (delta_pose_matrix = I, so `new_pose_matrix` only passes through the Euler round trip again)
    if (sqrt(sums[28]) < 1e-7) {
        fr->losses[it] = (float)sums[28];
        fr->status = PLS_W_TINY_RESIDUAL;
        if (0.f < threshold_delta) {
            fr->done = 1;
        } else {
            float prm[6];
            from_pose(fr->T, prm);
            build_pose(prm, fr->T);
            for (int i = 0; i < 6; ++i) fr->params[i] = prm[i];
        }
        return;
    }
    double dx[6];
    const double det = solve6(sums, dx);
    if (!(fabs(det) >= 1e-7)) {  // optimization.py:334-336
        fr->status = PLS_E_SINGULAR;
        fr->done = 1;
        return;
    }
    fr->losses[it] = (float)sums[27];
    float delta[6];
    float n2 = 0.f;
    for (int i = 0; i < 6; ++i) {
        delta[i] = (float)dx[i];
        n2 += delta[i] * delta[i];
    }
    if (sqrtf(n2) < threshold_delta) {  // icp_odometry.py:292-293: the last delta is not applied
        fr->done = 1;
        return;
    }
    float dT[16], Tn[16], prm[6];
    build_pose(delta, dT);
    mat4_mul(dT, fr->T, Tn);
    from_pose(Tn, prm);          // icp_odometry.py:296
    build_pose(prm, fr->T);      // icp_odometry.py:297
    for (int i = 0; i < 6; ++i) fr->params[i] = prm[i];
}

// Called by every block of a correspondence kernel (256 or 128 threads) after it stored its partial row: the block
// that arrives last (ticket in fr->pad) sums all rows in the fixed order and runs the solve -- one launch
// and one dependent-launch gap less per ICP iteration than a separate step kernel.
template <int THREADS = 256>
__device__ __forceinline__ void icp_finish_in_last_block(FrameResult* fr, const double* __restrict__ partials,
                                                         float threshold_delta) {
    __shared__ int s_last;
    __shared__ double s_sums[NACC];
    __threadfence();  // this block's partial row is visible before its ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const int ticket = atomicAdd(&fr->pad, 1);
        s_last = ticket == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    sum_partials_256<THREADS>(partials, (int)gridDim.x, s_sums);
    __syncthreads();
    if (threadIdx.x < NACC) fr->last_sums[threadIdx.x] = s_sums[threadIdx.x];
    if (threadIdx.x != 0) return;
    fr->pad = 0;
    icp_solve_and_update(fr, s_sums, threshold_delta);
}

}  // namespace pls
