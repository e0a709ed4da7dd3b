// Closed-form tail of the Procrustes registration, host/device so that the CPU test-suite can exercise the exact
// code the solve kernel runs (tests/test_abi.py builds a host harness with nvcc).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pls {

// Cyclic Jacobi on a symmetric 3x3 (float64): A = V diag(d) V^T, columns of V orthonormal.
__host__ __device__ inline void jacobi3(double A[3][3], double V[3][3], double d[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 32; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) d[i] = A[i][i];
}

// R = U diag(1, 1, sign(det U det V)) V^T of the cross-covariance C (row-major, reference rows x target columns),
// t = mu_r - R mu_t; mu = (mu_t, mu_r).  Registration.py:48-73.
__host__ __device__ inline void kabsch_from_cross(const double* C9, const double* mu, double* out_T /*[16]*/) {
    double Cm[3][3], A[3][3], V[3][3], d[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Cm[i][j] = C9[3 * i + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = Cm[0][i] * Cm[0][j] + Cm[1][i] * Cm[1][j] + Cm[2][i] * Cm[2][j];  // C^T C
    jacobi3(A, V, d);
    // order the eigenpairs by descending eigenvalue (LAPACK's singular-value order)
    int o[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[o[j]] > d[o[i]]) { const int t = o[i]; o[i] = o[j]; o[j] = t; }
    double v[3][3], u[3][3];  // v[k] = k-th right singular vector, u[k] = k-th left singular vector
    for (int k = 0; k < 3; ++k)
        for (int i = 0; i < 3; ++i) v[k][i] = V[i][o[k]];
    for (int k = 0; k < 2; ++k) {
        for (int i = 0; i < 3; ++i) u[k][i] = Cm[i][0] * v[k][0] + Cm[i][1] * v[k][1] + Cm[i][2] * v[k][2];
        if (k == 1) {  // Gram-Schmidt against u_0: C v_1 is orthogonal to it only up to rounding
            const double dp = u[1][0] * u[0][0] + u[1][1] * u[0][1] + u[1][2] * u[0][2];
            for (int i = 0; i < 3; ++i) u[1][i] -= dp * u[0][i];
        }
        const double nrm = sqrt(u[k][0] * u[k][0] + u[k][1] * u[k][1] + u[k][2] * u[k][2]);
        for (int i = 0; i < 3; ++i) u[k][i] /= nrm;
    }
    const double detV = v[0][0] * (v[1][1] * v[2][2] - v[1][2] * v[2][1]) - v[0][1] * (v[1][0] * v[2][2] - v[1][2] * v[2][0]) +
                        v[0][2] * (v[1][0] * v[2][1] - v[1][1] * v[2][0]);
    const double sg = detV < 0.0 ? -1.0 : 1.0;
    u[2][0] = sg * (u[0][1] * u[1][2] - u[0][2] * u[1][1]);
    u[2][1] = sg * (u[0][2] * u[1][0] - u[0][0] * u[1][2]);
    u[2][2] = sg * (u[0][0] * u[1][1] - u[0][1] * u[1][0]);
    double R[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = u[0][i] * v[0][j] + u[1][i] * v[1][j] + u[2][i] * v[2][j];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) out_T[4 * i + j] = R[i][j];
        out_T[4 * i + 3] = mu[3 + i] - (R[i][0] * mu[0] + R[i][1] * mu[1] + R[i][2] * mu[2]);
    }
    out_T[12] = 0.0; out_T[13] = 0.0; out_T[14] = 0.0; out_T[15] = 1.0;
}

}  // namespace pls
