// Smallest-eigenvalue direction of a symmetric 3x3 matrix: the normal of a map point from the second moments of its
// neighbourhood (replaces np.linalg.svd(covs)[2][:, 2, :], slam/odometry/local_map.py:414-416; the sign of the vector
// is arbitrary there and irrelevant downstream: J^T J and J^T r are even in the normal).
//
// Two solvers in float64 (the reference's LAPACK sgesdd works in float32, so both are more accurate than what they
// stand in for):
//   * closed form: eigenvalues by the trigonometric solution of the characteristic cubic, the vector as the largest
//     cross product of two rows of (A - lambda I).  ~300 instructions, no loop.  Accurate to ~eps * lambda_max / gap,
//     gap = lambda_mid - lambda_min: used when the gap is at least 1e-3 of the spectrum;
//   * cyclic Jacobi: unconditionally stable, ~6x the instructions: the fall-back for nearly degenerate
//     neighbourhoods (lines, isotropic blobs), where the direction is ill-defined anyway.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pls {

__host__ __device__ inline void smallest_eigenvector_jacobi(const float* c /*xx,xy,xz,yy,yz,zz*/, float* n) {
    double A[3][3] = {{c[0], c[1], c[2]}, {c[1], c[3], c[4]}, {c[2], c[4], c[5]}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-18 * diag || off == 0.0) break;
#pragma unroll
        for (int pq = 0; pq < 3; ++pq) {
            const int p = pq == 2 ? 1 : 0;
            const int q = pq == 0 ? 1 : 2;
            double apq = A[p][q];
            if (apq == 0.0) continue;
            double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double akp = A[k][p], akq = A[k][q];
                A[k][p] = cs * akp - sn * akq;
                A[k][q] = sn * akp + cs * akq;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double apk = A[p][k], aqk = A[q][k];
                A[p][k] = cs * apk - sn * aqk;
                A[q][k] = sn * apk + cs * aqk;
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double vkp = V[k][p], vkq = V[k][q];
                V[k][p] = cs * vkp - sn * vkq;
                V[k][q] = sn * vkp + cs * vkq;
            }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    double nx = V[0][m], ny = V[1][m], nz = V[2][m];
    double inv = 1.0 / sqrt(nx * nx + ny * ny + nz * nz);
    n[0] = (float)(nx * inv);
    n[1] = (float)(ny * inv);
    n[2] = (float)(nz * inv);
}

// Returns false when the closed form should not be trusted (caller falls back to Jacobi).
__host__ __device__ inline bool smallest_eigenvector_closed_form(const float* c, float* n) {
    const double a = c[0], b = c[1], cc = c[2], d = c[3], e = c[4], f = c[5];
    const double q = (a + d + f) / 3.0;
    const double p1 = b * b + cc * cc + e * e;
    const double aq = a - q, dq = d - q, fq = f - q;
    const double p2 = aq * aq + dq * dq + fq * fq + 2.0 * p1;
    if (!(p2 > 0.0)) return false;  // isotropic (or NaN)
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    // r = det((A - q I) / p) / 2 in [-1, 1]
    const double b11 = aq * ip, b22 = dq * ip, b33 = fq * ip, b12 = b * ip, b13 = cc * ip, b23 = e * ip;
    double r = 0.5 * (b11 * (b22 * b33 - b23 * b23) - b12 * (b12 * b33 - b23 * b13) + b13 * (b12 * b23 - b22 * b13));
    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    const double phi = acos(r) / 3.0;
    const double lmax = q + 2.0 * p * cos(phi);
    const double lmin = q + 2.0 * p * cos(phi + 2.0943951023931953);  // + 2 pi / 3
    const double lmid = 3.0 * q - lmax - lmin;
    if (!((lmid - lmin) > 1e-3 * (lmax - lmin))) return false;  // plane direction nearly degenerate
    // rows of A - lmin I; the eigenvector is orthogonal to all of them: take the best-conditioned cross product
    const double r0[3] = {a - lmin, b, cc}, r1[3] = {b, d - lmin, e}, r2[3] = {cc, e, f - lmin};
    const double c01[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    const double c02[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    const double c12[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double n01 = c01[0] * c01[0] + c01[1] * c01[1] + c01[2] * c01[2];
    const double n02 = c02[0] * c02[0] + c02[1] * c02[1] + c02[2] * c02[2];
    const double n12 = c12[0] * c12[0] + c12[1] * c12[1] + c12[2] * c12[2];
    const double* v = c01;
    double nn = n01;
    if (n02 > nn) { v = c02; nn = n02; }
    if (n12 > nn) { v = c12; nn = n12; }
    if (!(nn > 0.0)) return false;
    const double inv = 1.0 / sqrt(nn);
    n[0] = (float)(v[0] * inv);
    n[1] = (float)(v[1] * inv);
    n[2] = (float)(v[2] * inv);
    return true;
}

__host__ __device__ inline void smallest_eigenvector(const float* c, float* n) {
    if (!smallest_eigenvector_closed_form(c, n)) smallest_eigenvector_jacobi(c, n);
}

}  // namespace pls
