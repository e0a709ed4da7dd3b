// Per-point math of the motion de-skew (filters.cu), host/device so that the CPU test-suite can run the exact code
// of the kernel against the reference goldens (tests/host_harness.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pls {

struct DistortParams {
    double axis[3];   // unit axis of log R
    double angle;     // |log R|
    double t[3];      // translation of the relative pose
    int tr_f32;       // alpha and the pose are both float32: numpy forms alpha * t in float32
};

// axis * angle of a (near-)rotation matrix, angle in [0, pi]: what scipy's Slerp derives from
// rot[0].inv() * rot[1] with rot[0] = identity (preprocessing.py:175-178).
__host__ __device__ inline void rotation_vector(const double* R /*3x3 row-major*/, double* axis, double* angle) {
    const double w[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
    const double s = 0.5 * sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const double c = 0.5 * (R[0] + R[4] + R[8] - 1.0);
    *angle = atan2(s, c);
    if (s > 1e-8) {
        for (int i = 0; i < 3; ++i) axis[i] = w[i] / (2.0 * s);
        return;
    }
    if (c > 0.0) {  // angle ~ 0: log R ~ w / 2
        const double nrm = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        if (nrm > 0.0) {
            for (int i = 0; i < 3; ++i) axis[i] = w[i] / nrm;
            *angle = 0.5 * nrm;
        } else {
            axis[0] = 1.0; axis[1] = 0.0; axis[2] = 0.0;
            *angle = 0.0;
        }
        return;
    }
    // angle ~ pi: axis from the largest diagonal entry of (R + I) / 2 = k k^T
    const double B[9] = {0.5 * (R[0] + 1.0), 0.5 * R[1], 0.5 * R[2], 0.5 * R[3], 0.5 * (R[4] + 1.0), 0.5 * R[5],
                         0.5 * R[6], 0.5 * R[7], 0.5 * (R[8] + 1.0)};
    int k = 0;
    if (B[4] > B[0]) k = 1;
    if (B[8] > B[4 * k]) k = 2;
    const double d = sqrt(B[4 * k]);
    for (int i = 0; i < 3; ++i) axis[i] = B[3 * i + k] / d;
}

// out = exp(alpha log R) p + alpha t  (preprocessing.py:184-190); alpha arrives in the timestamps' dtype.
template <typename TS>
__host__ __device__ inline void distort_point(TS a, const DistortParams& prm, double px, double py, double pz, double& ox,
                                              double& oy, double& oz) {
    const double kx = prm.axis[0], ky = prm.axis[1], kz = prm.axis[2];
    const double ad = (double)a;
    double s, c;
    sincos(ad * prm.angle, &s, &c);
    const double v = 1.0 - c;
    // Rodrigues: R = I + s K + (1 - c) K^2 = c I + s K + (1 - c) k k^T
    const double kp = kx * px + ky * py + kz * pz;
    const double cx = ky * pz - kz * py, cy = kz * px - kx * pz, cz = kx * py - ky * px;
    ox = c * px + s * cx + v * kp * kx;
    oy = c * py + s * cy + v * kp * ky;
    oz = c * pz + s * cz + v * kp * kz;
    if (prm.tr_f32) {
        ox += (double)((float)a * (float)prm.t[0]);
        oy += (double)((float)a * (float)prm.t[1]);
        oz += (double)((float)a * (float)prm.t[2]);
    } else {
        ox += ad * prm.t[0];
        oy += ad * prm.t[1];
        oz += ad * prm.t[2];
    }
}

}  // namespace pls
