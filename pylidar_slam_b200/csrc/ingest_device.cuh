// Device/host math of the rows either side of the path: KITTI scan rectification
// (slam/dataset/kitti_dataset.py:200-231) and the pose chains of slam/eval/eval_odometry.py:80-96.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pls {

// KITTIOdometrySequence.correct_scan for one point: the HDL-64 intrinsic correction rotates every point by 0.205 degrees
// about the horizontal axis normal to its ray, u = normalise(p x e_z).  The reference's dtypes are reproduced:
// the axis and its outer product in float32 (np.cross / np.linalg.norm / `/=` on float32 arrays), the Rodrigues matrix
// c I + s [u]x + (1 - c) u u^T and the product with the point in float64 (c, s are numpy float64 scalars).  A point on
// the vertical axis has a zero cross product and comes out NaN, as in the reference.
__host__ __device__ inline void kitti_correct_point(float x, float y, float z, double c, double s, double* out) {
    // p x (0, 0, 1) = (y, -x, 0), exact in float32
    const float a0 = y, a1 = -x;
    // np.linalg.norm over (a0, a1, 0) in float32: two rounded products, one rounded sum (no fused multiply-add: the
    // device compiler would contract a0 * a0 + a1 * a1 and move the axis by an ulp)
#ifdef __CUDA_ARCH__
    const float nrm = sqrtf(__fadd_rn(__fmul_rn(a0, a0), __fmul_rn(a1, a1)));
#else
    const float s0 = a0 * a0, s1 = a1 * a1;
    const float nrm = sqrtf(s0 + s1);
#endif
    const float u0 = a0 / nrm, u1 = a1 / nrm;
    const double o00 = (double)(u0 * u0), o01 = (double)(u0 * u1), o11 = (double)(u1 * u1);  // float32 outer product
    const double k = 1.0 - c;
    // rows of c I + s [u]x + (1 - c) u u^T with u = (u0, u1, 0):  [u]x = [[0, 0, u1], [0, 0, -u0], [-u1, u0, 0]]
    const double r00 = c + k * o00, r01 = k * o01, r02 = s * (double)u1;
    const double r10 = k * o01, r11 = c + k * o11, r12 = s * (double)(-u0);
    const double r20 = s * (double)(-u1), r21 = s * (double)u0, r22 = c;
    const double px = (double)x, py = (double)y, pz = (double)z;
    out[0] = r00 * px + r01 * py + r02 * pz;
    out[1] = r10 * px + r11 * py + r12 * pz;
    out[2] = r20 * px + r21 * py + r22 * pz;
}

// General 4x4 inverse (Gauss-Jordan, partial pivoting) -- np.linalg.inv on a pose matrix.
template <typename T>
__host__ __device__ inline void inverse4(const T* A, T* out) {
    T m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m[i][j] = A[4 * i + j];
            m[i][4 + j] = (i == j) ? (T)1 : (T)0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r)
            if (fabs((double)m[r][c]) > fabs((double)m[piv][c])) piv = r;
        if (piv != c)
            for (int j = 0; j < 8; ++j) {
                const T t = m[c][j];
                m[c][j] = m[piv][j];
                m[piv][j] = t;
            }
        const T inv = (T)1 / m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] *= inv;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                const T f = m[r][c];
                for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
            }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[4 * i + j] = m[i][4 + j];
}

template <typename T>
__host__ __device__ inline void matmul4(const T* A, const T* B, T* C) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            T s = 0;
            for (int k = 0; k < 4; ++k) s += A[4 * i + k] * B[4 * k + j];
            C[4 * i + j] = s;
        }
}

}  // namespace pls
