// TEST INFRASTRUCTURE ONLY -- host harness around the __host__ __device__ math headers of the CUDA library.
//
// The kernels of filters.cu / registration.cu / gn.cu keep their per-element arithmetic in host/device inline
// functions (filters_device.cuh, registration_device.cuh, gn_device.cuh, pose_device.cuh).  This file wraps that
// arithmetic in plain host loops behind a tiny C ABI so that tests/test_host_math.py can check -- on a machine
// WITHOUT a GPU -- the very code the kernels execute against the goldens of the unmodified reference.  It is not
// part of the product and is never linked into libplslam_b200.so; only the reduction ORDER differs from the
// kernels (sequential here; warp-shuffle / block partials there), both accumulate in float64.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../pylidar_slam_b200/csrc/eigen_device.cuh"
#include "../pylidar_slam_b200/csrc/filters_device.cuh"
#include "../pylidar_slam_b200/csrc/gn_device.cuh"
#include "../pylidar_slam_b200/csrc/ingest_device.cuh"
#include "../pylidar_slam_b200/csrc/pose_device.cuh"
#include "../pylidar_slam_b200/csrc/projection_device.cuh"
#include "../pylidar_slam_b200/csrc/registration_device.cuh"
#include "../pylidar_slam_b200/csrc/training_device.cuh"

using namespace pls;

namespace {

template <typename TP, typename TS>
void distort_host(const TP* pc, const TS* ts, int64_t n, const DistortParams& prm, double* out) {
    // ts_minmax_kernel + the prologue of distort_kernel
    double mn = INFINITY, mx = -INFINITY;
    bool bad = false;
    for (int64_t i = 0; i < n; ++i) {
        const double t = (double)ts[i];
        if (t != t) bad = true;
        mn = fmin(mn, t);
        mx = fmax(mx, t);
    }
    const TS tmin = (TS)mn, den = (TS)mx - (TS)mn;
    const int mode = bad ? 2 : (den == (TS)0 ? 1 : 0);
    for (int64_t i = 0; i < n; ++i) {
        const TS t = ts[i];
        TS a;
        if (mode == 0) a = (t - tmin) / den;
        else if (mode == 1) a = t * (TS)0;
        else a = (TS)NAN;
        distort_point<TS>(a, prm, (double)pc[3 * i], (double)pc[3 * i + 1], (double)pc[3 * i + 2], out[3 * i],
                          out[3 * i + 1], out[3 * i + 2]);
    }
}

// host replica of accumulate_normal_equations (device-only in gn_device.cuh because of its unrolled registers)
template <typename T>
void accumulate_host(double* acc, const T* J, T w, T wr, T r) {
    double wj[6];
    for (int a = 0; a < 6; ++a) wj[a] = (double)(J[a] * w);
    int k = 0;
    for (int a = 0; a < 6; ++a)
        for (int b = a; b < 6; ++b) acc[k++] += wj[a] * wj[b];
    for (int a = 0; a < 6; ++a) acc[21 + a] += wj[a] * (double)wr;
    acc[27] += (double)wr * (double)wr;
    acc[28] += (double)r * (double)r;
    acc[29] += 1.0;
}

// gn_accumulate_kernel + gn_solve_kernel, sequentially
template <typename T>
int align_host(int cost, const T* ref, const T* tgt, const T* nrm, int64_t n, int scheme, T sigma, int max_iters,
               T norm_stop, const T* x0, T* out_x, T* out_dT, T* out_loss) {
    T x[6] = {0, 0, 0, 0, 0, 0};
    if (x0) memcpy(x, x0, sizeof(x));
    int status = 0;
    build_pose(x, out_dT);
    for (int it = 0; it < (max_iters < 1 ? 1 : max_iters); ++it) {
        T M[16], R[9], t[3], dR[27];
        build_pose(x, M);
        R[0] = M[0]; R[1] = M[1]; R[2] = M[2]; R[3] = M[4]; R[4] = M[5]; R[5] = M[6]; R[6] = M[8]; R[7] = M[9]; R[8] = M[10];
        t[0] = M[3]; t[1] = M[7]; t[2] = M[11];
        euler_jacobian(x + 3, dR);
        double acc[30];
        for (int a = 0; a < 30; ++a) acc[a] = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            const T* p = tgt + 3 * i;
            const T* q = ref + 3 * i;
            T J[6];
            T r = cost == 0 ? p2plane_residual_jacobian<T>(p, q, nrm + 3 * i, R, t, dR, J)
                            : p2point_residual_jacobian<T>(p, q, R, t, dR, J);
            T w = ls_weight<T>(scheme, sigma, r, p, q);
            T wr = r * w;
            if (out_loss) out_loss[i] = wr * wr;
            accumulate_host<T>(acc, J, w, wr, r);
        }
        if (sqrt(acc[28]) < 1e-7) { status = PLS_W_TINY_RESIDUAL; build_pose(x, out_dT); break; }
        double dx[6];
        const double det = solve6(acc, dx);
        if (!(fabs(det) >= 1e-7)) { status = PLS_E_SINGULAR; break; }
        double nrm2 = 0.0;
        for (int i = 0; i < 6; ++i) {
            const T d = (T)dx[i];
            x[i] = x[i] + d;
            nrm2 += (double)d * (double)d;
        }
        build_pose(x, out_dT);
        if (sqrt(nrm2) < (double)norm_stop) break;
    }
    memcpy(out_x, x, sizeof(x));
    return status;
}

}  // namespace

extern "C" {

void hh_distort(const void* pc, int pc_is_f64, const void* ts, int ts_is_f64, int64_t n, const double* pose16,
                int pose_is_f64, double* out) {
    DistortParams prm;
    const double R[9] = {pose16[0], pose16[1], pose16[2], pose16[4], pose16[5], pose16[6], pose16[8], pose16[9], pose16[10]};
    rotation_vector(R, prm.axis, &prm.angle);
    prm.t[0] = pose16[3]; prm.t[1] = pose16[7]; prm.t[2] = pose16[11];
    prm.tr_f32 = (!ts_is_f64 && !pose_is_f64) ? 1 : 0;
    if (pc_is_f64) {
        if (ts_is_f64) distort_host<double, double>((const double*)pc, (const double*)ts, n, prm, out);
        else distort_host<double, float>((const double*)pc, (const float*)ts, n, prm, out);
    } else {
        if (ts_is_f64) distort_host<float, double>((const float*)pc, (const double*)ts, n, prm, out);
        else distort_host<float, float>((const float*)pc, (const float*)ts, n, prm, out);
    }
}

void hh_rotation_vector(const double* R9, double* axis3, double* angle) { rotation_vector(R9, axis3, angle); }

void hh_kabsch(const double* C9, const double* mu6, double* T16) { kabsch_from_cross(C9, mu6, T16); }

int hh_align(int cost, int is_f64, const void* ref, const void* tgt, const void* nrm, int64_t n, int scheme, double sigma,
             int max_iters, double norm_stop, const void* x0, void* out_x, void* out_dT, void* out_loss) {
    if (is_f64)
        return align_host<double>(cost, (const double*)ref, (const double*)tgt, (const double*)nrm, n, scheme, sigma, max_iters,
                                  norm_stop, (const double*)x0, (double*)out_x, (double*)out_dT, (double*)out_loss);
    return align_host<float>(cost, (const float*)ref, (const float*)tgt, (const float*)nrm, n, scheme, (float)sigma, max_iters,
                             (float)norm_stop, (const float*)x0, (float*)out_x, (float*)out_dT, (float*)out_loss);
}

void hh_loss_pixel(int scheme, double sigma, const float* pw, const float* q, const float* n, double* out5 /*mask, c2, g[3]*/) {
    loss_pixel_terms(scheme, sigma, pw, q, n, out5[0], out5[1], out5 + 2);
}

// loss_zbuf_kernel + loss_accumulate_kernel + loss_finalize_kernel of training.cu, sequentially (same per-point code)
void hh_p2plane_loss(const float* vt, const float* vr, const float* nr, const float* mats, const float* params, int B, int H,
                     int W, float up, float down, int scheme, float sigma, float* out_loss, float* out_pb, float* out_gm,
                     float* out_gp) {
    const int64_t hw = (int64_t)H * W;
    const ProjConst pc = make_proj_const(H, W, up, down);
    unsigned long long* zbuf = new unsigned long long[(size_t)B * hw];
    double total = 0.0;
    for (int b = 0; b < B; ++b) {
        float Mb[16];
        if (mats) memcpy(Mb, mats + 16 * b, sizeof(Mb));
        else build_pose(params + 6 * b, Mb);
        const float* vm = vt + 3 * hw * b;
        auto moved = [&](int64_t i, float* p, float* pm) {
            p[0] = vm[i]; p[1] = vm[hw + i]; p[2] = vm[2 * hw + i];
            if (p[0] == 0.f && p[1] == 0.f && p[2] == 0.f) return false;
            pm[0] = p[0] * Mb[0] + p[1] * Mb[1] + p[2] * Mb[2] + Mb[3];
            pm[1] = p[0] * Mb[4] + p[1] * Mb[5] + p[2] * Mb[6] + Mb[7];
            pm[2] = p[0] * Mb[8] + p[1] * Mb[9] + p[2] * Mb[10] + Mb[11];
            return true;
        };
        unsigned long long* zb = zbuf + hw * b;
        for (int64_t i = 0; i < hw; ++i) zb[i] = ~0ull;
        for (int64_t i = 0; i < hw; ++i) {
            float p[3], pm[3], r;
            int pix;
            if (!moved(i, p, pm) || !project_to_pixel(pm[0], pm[1], pm[2], pc, pix, r)) continue;
            unsigned int bits;
            memcpy(&bits, &r, 4);
            const unsigned long long key = ((unsigned long long)bits << 32) | (unsigned long long)(uint32_t)i;
            if (key < zb[pix]) zb[pix] = key;
        }
        double row[16];
        for (int a = 0; a < 16; ++a) row[a] = 0.0;
        for (int64_t i = 0; i < hw; ++i) {
            float p[3], pm[3], r;
            int pix;
            if (!moved(i, p, pm) || !project_to_pixel(pm[0], pm[1], pm[2], pc, pix, r)) continue;
            const uint32_t win = (uint32_t)(zb[pix] & 0xffffffffull);
            float pw[3] = {pm[0], pm[1], pm[2]};
            if (win != (uint32_t)i) {
                float pj[3];
                moved((int64_t)win, pj, pw);
            }
            const float q[3] = {vr[3 * hw * b + pix], vr[3 * hw * b + hw + pix], vr[3 * hw * b + 2 * hw + pix]};
            const float n[3] = {nr[3 * hw * b + pix], nr[3 * hw * b + hw + pix], nr[3 * hw * b + 2 * hw + pix]};
            double mask, c2, g[3];
            loss_pixel_terms(scheme, (double)sigma, pw, q, n, mask, c2, g);
            if (win == (uint32_t)i) { row[0] += c2; row[1] += mask; }
            for (int a = 0; a < 3; ++a) {
                row[2 + a] += g[a];
                for (int c = 0; c < 3; ++c) row[5 + 3 * a + c] += g[a] * (double)p[c];
            }
        }
        const double Mc = row[1], lb = row[0] / Mc, sc = 1.0 / (Mc * (double)B);
        total += lb;
        if (out_pb) out_pb[b] = (float)lb;
        double G[12];
        for (int a = 0; a < 3; ++a) {
            for (int c = 0; c < 3; ++c) G[4 * a + c] = row[5 + 3 * a + c] * sc;
            G[4 * a + 3] = row[2 + a] * sc;
        }
        if (out_gm) {
            for (int k = 0; k < 12; ++k) out_gm[16 * b + k] = (float)G[k];
            for (int k = 12; k < 16; ++k) out_gm[16 * b + k] = 0.f;
        }
        if (out_gp && params) {
            float dR[27];
            euler_jacobian(params + 6 * b + 3, dR);
            for (int a = 0; a < 3; ++a) out_gp[6 * b + a] = (float)G[4 * a + 3];
            for (int k = 0; k < 3; ++k) {
                double s2 = 0.0;
                for (int a = 0; a < 3; ++a)
                    for (int c = 0; c < 3; ++c) s2 += G[4 * a + c] * (double)dR[9 * k + 3 * a + c];
                out_gp[6 * b + 3 + k] = (float)s2;
            }
        }
    }
    *out_loss = (float)(total / (double)B);
    delete[] zbuf;
}

// rank 4: kitti_correct_kernel's per-point math and the pose chains (relative_poses_kernel / absolute_poses_kernel)
void hh_kitti_correct(const float* scan, int64_t n, int stride, double* out) {
    const double theta = 0.205 * 3.141592653589793 / 180.0;
    for (int64_t i = 0; i < n; ++i)
        kitti_correct_point(scan[stride * i], scan[stride * i + 1], scan[stride * i + 2], cos(theta), sin(theta), out + 3 * i);
}

void hh_pose_chains(const double* poses, int64_t n, double* rel, double* absolute) {
    for (int64_t i = 0; i < n; ++i) {
        double prev[16], inv[16];
        for (int k = 0; k < 16; ++k) prev[k] = i > 0 ? poses[16 * (i - 1) + k] : ((k % 5 == 0) ? 1.0 : 0.0);
        inverse4<double>(prev, inv);
        matmul4<double>(inv, poses + 16 * i, rel + 16 * i);
    }
    for (int k = 0; k < 16; ++k) absolute[k] = rel[k];
    for (int64_t i = 1; i < n; ++i) matmul4<double>(absolute + 16 * (i - 1), rel + 16 * i, absolute + 16 * i);
}

// smallest-eigenvalue directions of n symmetric 3x3 matrices (xx,xy,xz,yy,yz,zz): which = 0 the product's solver
// (closed form with Jacobi fall-back), 1 Jacobi only, 2 closed form only (used[i] = 0 where it declined)
void hh_smallest_eigenvectors(const float* cov6, int64_t n, int which, float* out3, int* used) {
    for (int64_t i = 0; i < n; ++i) {
        bool ok = true;
        if (which == 0) smallest_eigenvector(cov6 + 6 * i, out3 + 3 * i);
        else if (which == 1) smallest_eigenvector_jacobi(cov6 + 6 * i, out3 + 3 * i);
        else ok = smallest_eigenvector_closed_form(cov6 + 6 * i, out3 + 3 * i);
        if (used) used[i] = ok ? 1 : 0;
    }
}

}  // extern "C"
