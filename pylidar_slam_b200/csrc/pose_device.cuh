// Device/host pose algebra of the hot path (SE(3) <-> (tx,ty,tz,ex,ey,ez), Euler xyz).
// Follows slam/common/rotation.py:144-150 (R = Rz Ry Rx), :253-270 (inverse with the gimbal
// branch), :166-216 (analytic Jacobian) and slam/common/pose.py:120-207.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pls {

template <typename T>
__host__ __device__ inline void euler_to_mat(const T* e, T* R /*3x3 row-major*/) {
    T cx = cos(e[0]), sx = sin(e[0]);
    T cy = cos(e[1]), sy = sin(e[1]);
    T cz = cos(e[2]), sz = sin(e[2]);
    // Rz * Ry * Rx
    R[0] = cz * cy;  R[1] = cz * sy * sx - sz * cx;  R[2] = cz * sy * cx + sz * sx;
    R[3] = sz * cy;  R[4] = sz * sy * sx + cz * cx;  R[5] = sz * sy * cx - cz * sx;
    R[6] = -sy;      R[7] = cy * sx;                 R[8] = cy * cx;
}

template <typename T>
__host__ __device__ inline void mat_to_euler(const T* R, T* e) {
    T sy = sqrt(R[0] * R[0] + R[3] * R[3]);
    bool singular = sy < (T)1e-6;
    if (!singular) {
        e[0] = atan2(R[7], R[8]);
        e[1] = atan2(-R[6], sy);
        e[2] = atan2(R[3], R[0]);
    } else {
        e[0] = atan2(-R[5], R[4]);
        e[1] = atan2(-R[6], sy);
        e[2] = (T)0;
    }
}

template <typename T>
__host__ __device__ inline void build_pose(const T* params, T* M /*4x4 row-major*/) {
    T R[9];
    euler_to_mat(params + 3, R);
    M[0] = R[0]; M[1] = R[1]; M[2] = R[2];  M[3] = params[0];
    M[4] = R[3]; M[5] = R[4]; M[6] = R[5];  M[7] = params[1];
    M[8] = R[6]; M[9] = R[7]; M[10] = R[8]; M[11] = params[2];
    M[12] = 0;   M[13] = 0;   M[14] = 0;    M[15] = 1;
}

template <typename T>
__host__ __device__ inline void from_pose(const T* M, T* params) {
    T R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
    params[0] = M[3]; params[1] = M[7]; params[2] = M[11];
    mat_to_euler(R, params + 3);
}

template <typename T>
__host__ __device__ inline void mat4_mul(const T* A, const T* B, T* C) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            T s = 0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            C[i * 4 + j] = s;
        }
}

// dR/de_k for k = x,y,z : three 3x3 row-major matrices (rotation.py:166-184)
template <typename T>
__host__ __device__ inline void euler_jacobian(const T* e, T* dR /*[3][9]*/) {
    T cx = cos(e[0]), sx = sin(e[0]);
    T cy = cos(e[1]), sy = sin(e[1]);
    T cz = cos(e[2]), sz = sin(e[2]);
    T Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx};
    T Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    T Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    T Jx[9] = {0, 0, 0, 0, -sx, -cx, 0, cx, -sx};
    T Jy[9] = {-sy, 0, cy, 0, 0, 0, -cy, 0, -sy};
    T Jz[9] = {-sz, -cz, 0, cz, -sz, 0, 0, 0, 0};
    auto mul3 = [](const T* A, const T* B, T* C) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                T s = 0;
                for (int k = 0; k < 3; ++k) s += A[i * 3 + k] * B[k * 3 + j];
                C[i * 3 + j] = s;
            }
    };
    T tmp[9];
    mul3(Rz, Ry, tmp); mul3(tmp, Jx, dR);
    mul3(Rz, Jy, tmp); mul3(tmp, Rx, dR + 9);
    mul3(Jz, Ry, tmp); mul3(tmp, Rx, dR + 18);
}

// Rigid inverse of a 4x4 pose (R^T, -R^T t), computed in double.
__host__ __device__ inline void rigid_inverse(const float* M, float* out) {
    double R[9] = {M[0], M[1], M[2], M[4], M[5], M[6], M[8], M[9], M[10]};
    double t[3] = {M[3], M[7], M[11]};
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out[i * 4 + j] = (float)R[j * 3 + i];
        out[i * 4 + 3] = (float)(-(R[0 * 3 + i] * t[0] + R[1 * 3 + i] * t[1] + R[2 * 3 + i] * t[2]));
    }
    out[12] = 0; out[13] = 0; out[14] = 0; out[15] = 1;
}

// Solves H dx = -g for the symmetric positive semi-definite 6x6 system given the 21 upper-triangle entries
// and g; returns det(H).  Gaussian elimination WITHOUT pivoting (stable for SPD matrices), fully unrolled so
// that the matrix lives in registers: this runs on one thread between two correspondence launches, where a
// local-memory, dynamically indexed pivoting version cost microseconds of pure latency.  A zero pivot
// (rank-deficient H) turns det into 0 or NaN, which the callers' |det| >= 1e-7 guard rejects.
__host__ __device__ inline double solve6(const double* sums /*21 upper + 6*/, double* dx) {
    double A[6][7];
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) {
                A[i][j] = sums[k];
                A[j][i] = sums[k];
                ++k;
            }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i][6] = -sums[21 + i];
    double det = 1.0;
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const double piv = A[c][c];
        det *= piv;
        const double inv = 1.0 / piv;
#pragma unroll
        for (int r = c + 1; r < 6; ++r) {
            const double f = A[r][c] * inv;
#pragma unroll
            for (int j = c + 1; j < 7; ++j) A[r][j] -= f * A[c][j];
        }
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = A[i][6];
#pragma unroll
        for (int j = i + 1; j < 6; ++j) s -= A[i][j] * dx[j];
        dx[i] = s / A[i][i];
    }
    return det;
}

}  // namespace pls
