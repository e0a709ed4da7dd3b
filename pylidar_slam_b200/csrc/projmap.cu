// placeholder until the projective local map lands (K3/K4)
#include "internal.cuh"
namespace pls {
void projmap_reset(pls_context* ctx) { ctx->pm.K = 0; ctx->pm.valid = false; ctx->pm.host_poses.clear(); }
void projmap_update(pls_context*, const float*, const float*) { throw Error{PLS_E_STATE, "projective map: not built yet"}; }
int projmap_icp_iteration(pls_context*, int64_t, int, int) { throw Error{PLS_E_STATE, "projective map: not built yet"}; }
void launch_normal_map(pls_context*, const float*, int, int, int, int, float*) { throw Error{PLS_E_STATE, "normal map: not built yet"}; }
}
extern "C" {
int pls_normal_map(pls_context* ctx, const float*, int, int, int, int, float*) { return PLS_E_STATE; }
int pls_compute_neighbors(pls_context* ctx, const float*, const float*, const float*, int, int, int, int, float*, float*) { return PLS_E_STATE; }
int pls_projmap_update(pls_context* ctx, const float*, const float*) { return PLS_E_STATE; }
int pls_projmap_num_frames(pls_context* ctx, int*) { return PLS_E_STATE; }
int pls_projmap_model(pls_context* ctx, float*, float*) { return PLS_E_STATE; }
int pls_projmap_nn_search(pls_context* ctx, const float*, int64_t, float*, float*, float*, int64_t*) { return PLS_E_STATE; }
}
