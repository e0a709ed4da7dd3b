// Spherical projection device math (slam/common/projection.py:11-73,393-401), float32 in the
// reference's operation order.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pls {

struct ProjConst {
    int H, W;
    float Hf, Wf;
    float abs_down;  // |fov_down| in radians, rounded to float like the torch scalar
    float fov;       // |fov_down| + |fov_up|
};

inline ProjConst make_proj_const(int H, int W, float up_deg, float down_deg) {
    ProjConst pc;
    pc.H = H;
    pc.W = W;
    double up = (double)up_deg / 180.0 * 3.141592653589793;
    double down = (double)down_deg / 180.0 * 3.141592653589793;
    pc.abs_down = (float)fabs(down);
    pc.fov = (float)(fabs(down) + fabs(up));
    pc.Wf = (float)W;
    pc.Hf = (float)H;
    return pc;
}

#ifdef __CUDACC__
// Float pixel coordinates as torch__spherical_projection returns them: (-1,-1) for the null point.
__host__ __device__ __forceinline__ void project_point(float x, float y, float z, const ProjConst& pc, float& row, float& col,
                                              float& r_out) {
    const float kPi = 3.14159274101257324f;  // float(np.pi)
    float r = sqrtf(x * x + y * y + z * z);
    r_out = r;
    bool null = (r == 0.0f);
    float rr = null ? 0.001f : r;
    float theta = -atan2f(y, x);
    float phi = asinf(z / rr);
    float c = 0.5f * (theta / kPi + 1.0f);
    float rw = 1.0f - (phi + pc.abs_down) / pc.fov;
    c = c * pc.Wf;
    rw = rw * pc.Hf;
    row = null ? -1.0f : rw;
    col = null ? -1.0f : c;
}

// Rounded pixel index + validity (projection.py:393-401,408).  False for NaN / null / out of image.
__host__ __device__ __forceinline__ bool project_to_pixel(float x, float y, float z, const ProjConst& pc, int& pix, float& r) {
    float row, col;
    project_point(x, y, z, pc, row, col, r);
    float pr = rintf(row), pcn = rintf(col);
    bool ok = (pr >= 0.0f) && (pr <= (float)(pc.H - 1)) && (pcn >= 0.0f) && (pcn <= (float)(pc.W - 1)) && (r > 0.0f);
    if (!ok) return false;
    pix = (int)pr * pc.W + (int)pcn;
    return true;
}
#endif

}  // namespace pls
