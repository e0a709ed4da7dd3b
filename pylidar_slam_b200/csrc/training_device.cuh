// Per-pixel arithmetic of the unsupervised point-to-plane training loss (training.cu), host/device so that the CPU
// test-suite can run the exact code of the kernel against the reference's autograd (tests/host_harness.cu).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include "../../include/plslam_b200.h"

namespace pls {

// C(a) of _LS_SCHEME[scheme].cost (slam/common/optimization.py:61-208) for a = |r| >= 0, its slope dC/da and -- for
// the neighbourhood scheme, whose weights exp(-|p' - q|^2 / sigma^2) are NOT detached in the loss
// (loss_modules.py:101) -- dC/d(|p' - q|^2).
__host__ __device__ inline void loss_cost(int scheme, double sigma, double a, double d2, double& C, double& dCa, double& dCd2) {
    dCd2 = 0.0;
    const double a2 = a * a;
    switch (scheme) {
        case PLS_SCHEME_DEFAULT:
        case PLS_SCHEME_LEAST_SQUARE:
            C = a2; dCa = 2.0 * a;
            break;
        case PLS_SCHEME_HUBER:
            if (a < sigma) { C = a2; dCa = 2.0 * a; }
            else { C = 2.0 * sigma * a - sigma * sigma; dCa = 2.0 * sigma; }
            break;
        case PLS_SCHEME_EXP: {
            const double e = exp(-a2 / (sigma * sigma));
            C = a2 * e; dCa = 2.0 * a * e * (1.0 - a2 / (sigma * sigma));
            break;
        }
        case PLS_SCHEME_NEIGHBORHOOD: {
            const double w = exp(-d2 / (sigma * sigma));
            C = a2 * w; dCa = 2.0 * a * w; dCd2 = -a2 * w / (sigma * sigma);
            break;
        }
        case PLS_SCHEME_GEMAN_MCCLURE: {
            const double s = sigma + a2;
            C = sigma * a2 / s; dCa = 2.0 * sigma * sigma * a / (s * s);
            break;
        }
        case PLS_SCHEME_SQUARE_GEMAN_MCCLURE: {
            const double s = sigma + a2;
            C = a2 * (sigma / s) * (sigma / s); dCa = 2.0 * a * sigma * sigma * (sigma - a2) / (s * s * s);
            break;
        }
        default: {  // PLS_SCHEME_CAUCHY
            C = log(1.0 + a2 / (sigma * sigma)); dCa = 2.0 * a / (sigma * sigma + a2);
            break;
        }
    }
}

// One pixel of the loss (loss_modules.py:90-102): pw = the transformed target point that won the pixel (zeros if the
// pixel is empty), q / n = reference vertex / normal of the pixel.  Returns mask (0/1), C(|r|)^2 and
// g = d(C^2)/d(pw) (not yet divided by the mask count and the batch size).
__host__ __device__ inline void loss_pixel_terms(int scheme, double sigma, const float* pw, const float* q, const float* n,
                                                 double& mask, double& c2, double* g) {
    const bool ok = !(n[0] == 0.f && n[1] == 0.f && n[2] == 0.f) && !(q[0] == 0.f && q[1] == 0.f && q[2] == 0.f) &&
                    !(pw[0] == 0.f && pw[1] == 0.f && pw[2] == 0.f);
    mask = ok ? 1.0 : 0.0;
    c2 = 0.0;
    g[0] = g[1] = g[2] = 0.0;
    if (!ok) return;  // residual = mask * |.| = 0 and every cost function vanishes (with zero slope) at 0
    const double dx = (double)q[0] - (double)pw[0], dy = (double)q[1] - (double)pw[1], dz = (double)q[2] - (double)pw[2];
    const double r = dx * (double)n[0] + dy * (double)n[1] + dz * (double)n[2];
    const double a = fabs(r);
    double C, dCa, dCd2;
    loss_cost(scheme, sigma, a, dx * dx + dy * dy + dz * dz, C, dCa, dCd2);
    c2 = C * C;
    const double sgn = r > 0.0 ? 1.0 : (r < 0.0 ? -1.0 : 0.0);
    const double ka = 2.0 * C * dCa * sgn;   // d(C^2)/dr
    const double kd = 2.0 * C * dCd2 * 2.0;  // d(C^2)/d(d2) * d(d2)/d(pw - q)
    // dr/dpw = -n ;  d(d2)/dpw = 2 (pw - q) = -2 (q - pw)
    g[0] = -ka * (double)n[0] - kd * dx;
    g[1] = -ka * (double)n[1] - kd * dy;
    g[2] = -ka * (double)n[2] - kd * dz;
}

}  // namespace pls
