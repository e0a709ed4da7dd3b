"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the rows either side of the hot path (SURVEY.md section 8f,
ranks 1-2): the Distortion and Voxelization filters that precede GridSample in the shipped preprocessing
chains, and the point-to-point Gauss-Newton alignment + weighted Procrustes that sit behind the same
RIGID_ALIGNMENT registry as the point-to-plane alignment.

Same rules as oracle/icp_oracle.py: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
it; the product path never does.

Parity status: PINNED against the unmodified reference -- tests/golden/next_rows.npz is written by
tests/golden/make_golden_next.py (reference imported under oracle/ref_shims.py) and
tests/test_next_rows_oracle.py checks every function below against it.  scipy's Slerp (the reference's
dependency for the de-skew) is restated here with the closed-form rotation-vector interpolation it
implements, so the oracle is independent of the library call it pins.

Paths cited are relative to the reference root.
"""
import numpy as np
import torch

from . import icp_oracle as orc


# --------------------------------------------------------------------------------------
# Distortion.filter                         slam/preprocessing.py:148-191
# --------------------------------------------------------------------------------------
def rotation_vector(R: np.ndarray) -> np.ndarray:
    """axis * angle of a rotation matrix (float64), angle in [0, pi] -- what
    (rot[0].inv() * rot[1]).as_rotvec() yields inside scipy's Slerp for rot[0] = identity."""
    R = np.asarray(R, np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = 0.5 * np.linalg.norm(w)            # sin(angle)
    c = 0.5 * (np.trace(R) - 1.0)          # cos(angle)
    angle = np.arctan2(s, c)
    if s > 1e-8:
        return w / (2.0 * s) * angle
    if c > 0:                               # angle ~ 0: first-order
        return 0.5 * w
    # angle ~ pi: axis from the largest diagonal entry of (R + I) / 2
    B = 0.5 * (R + np.eye(3))
    k = int(np.argmax(np.diag(B)))
    axis = B[:, k] / np.sqrt(B[k, k])
    return axis * angle


def rodrigues(rotvec: np.ndarray) -> np.ndarray:
    """[n,3] rotation vectors -> [n,3,3] matrices (float64)."""
    rv = np.asarray(rotvec, np.float64).reshape(-1, 3)
    th = np.linalg.norm(rv, axis=1)
    small = th < 1e-12
    k = rv / np.where(small, 1.0, th)[:, None]
    K = np.zeros((rv.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    R = np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)
    R[small] = np.eye(3)
    return R


def distort(pc: np.ndarray, timestamps: np.ndarray, rpose: np.ndarray) -> np.ndarray:
    """De-skew a frame with the estimated relative motion (preprocessing.py:171-191):
    alpha = (t - min t) / (max t - min t) in the timestamps' dtype (all zeros when max == min); per point
    the rotation Slerp(identity -> R)(alpha) = exp(alpha * log R) in float64, the translation alpha * t in
    numpy's promoted dtype of (alpha, rpose); output float64 [n,3]."""
    ts = np.asarray(timestamps).reshape(-1)
    diff = np.max(ts) - np.min(ts)
    alpha = ts * 0 if diff == 0.0 else (ts - np.min(ts)) / (np.max(ts) - np.min(ts))
    rv = rotation_vector(rpose[:3, :3].astype(np.float64))
    rots = rodrigues(alpha.astype(np.float64)[:, None] * rv[None, :])
    tr = alpha.reshape(-1, 1) * rpose[:3, 3].reshape(1, 3)
    return np.einsum("nij,nj->ni", rots, pc) + tr


# --------------------------------------------------------------------------------------
# Voxelization.filter                       slam/preprocessing.py:71-97
#   voxel_normal_distribution               slam/common/pointcloud.py:83-167
# --------------------------------------------------------------------------------------
def voxel_normal_distribution(points: np.ndarray, hashes: np.ndarray):
    """Per distinct hash (ascending): number of points, mean, scatter matrix sum (x - mean)(x - mean)^T
    (NOT divided by the count), and for every point the rank of its voxel.  The reference accumulates in
    the cloud's dtype in numba's (unstable) argsort order; this restatement accumulates in float64 in
    index order and casts, so float32 outputs agree to float32 rounding of the sums."""
    order = np.argsort(hashes, kind="stable")
    hs = hashes[order]
    head = np.ones(hs.shape[0], dtype=bool)
    head[1:] = hs[1:] != hs[:-1]
    vid_sorted = np.cumsum(head) - 1
    V = int(vid_sorted[-1]) + 1
    p = points[order].astype(np.float64)
    sizes = np.bincount(vid_sorted, minlength=V).astype(np.int64)
    sums = np.zeros((V, 3))
    np.add.at(sums, vid_sorted, p)
    means = sums / sizes[:, None]
    c = p - means[vid_sorted]
    covs = np.zeros((V, 3, 3))
    np.add.at(covs, vid_sorted, c[:, :, None] * c[:, None, :])
    ids = np.empty(points.shape[0], dtype=np.int64)
    ids[order] = vid_sorted
    return sizes, means.astype(points.dtype), covs.astype(points.dtype), ids


def voxelization(points: np.ndarray, voxel: float):
    """-> dict with the keys Voxelization.filter writes (preprocessing.py:80-97)."""
    coords = orc.voxel_coords(points, voxel)
    hashes = orc.voxel_hashes(coords)
    sizes, means, covs, ids = voxel_normal_distribution(points, hashes)
    return dict(voxel_hashes=hashes, voxel_coordinates=coords, voxel_sizes=sizes, voxel_means=means,
                voxel_covariances=covs, voxel_indices=ids)


# --------------------------------------------------------------------------------------
# GaussNewtonPointToPointAlignment.align    slam/odometry/alignment.py:144-189
#   PointToPointCost                        slam/common/optimization.py:458-541
# --------------------------------------------------------------------------------------
def p2point_residual(x, tgt, ref):
    """r_i = |R(x) p_i + t(x) - q_i|  (optimization.py:526-539)."""
    d = orc.apply_transformation(tgt, orc.build_pose_matrix(x)) - ref
    return torch.sqrt((d * d).sum(dim=-1))


def p2point_jacobian(x, tgt, ref):
    """J_i[k] = (dT/dx_k p~_i) . (R p_i + t - q_i)  (optimization.py:485-501).  NOTE: this is the reference's
    Jacobian as written -- the gradient of r^2 / 2, i.e. r times the Jacobian of r -- not d r / d x."""
    B, N, _ = tgt.shape
    d = orc.apply_transformation(tgt, orc.build_pose_matrix(x)) - ref
    dR = orc.euler_jacobian(x[:, 3:])  # [B,3,3,3]
    J = torch.zeros(B, N, 6, dtype=tgt.dtype)
    J[:, :, :3] = d
    for k in range(3):
        J[:, :, 3 + k] = (torch.einsum("bij,bnj->bni", dR[:, k], tgt) * d).sum(-1)
    return J


def align_p2point(ref, tgt, scheme="default", sigma=0.5, max_iters=1, norm_stop=1e-3, x0=None):
    """GaussNewton.compute (optimization.py:296-344) on the point-to-point closures ->
    (dT [B,4,4], x [B,6], (w r)^2 [B,N], status)."""
    B = ref.shape[0]
    x = torch.zeros(B, 6, dtype=ref.dtype) if x0 is None else x0
    res, status = None, "ok"
    for _ in range(max(max_iters, 1)):
        J = p2point_jacobian(x, tgt, ref)
        res = p2point_residual(x, tgt, ref)
        if res.norm() < 1e-7:
            status = "tiny_residual"
            break
        w = orc.ls_weights(scheme, sigma, res, tgt, ref)
        res = res * w
        J = J * w.unsqueeze(-1)
        Jt = J.permute(0, 2, 1)
        Hm = Jt @ J
        if torch.any(Hm.det().abs() < 1e-7):
            raise orc.SingularHessian("Invalid Jacobian in Gauss Newton minimization")
        dx = -Hm.inverse() @ Jt @ res.unsqueeze(-1)
        x = x + dx[:, :, 0]
        if dx.norm() < norm_stop:
            break
    return orc.build_pose_matrix(x), x, res * res, status


# --------------------------------------------------------------------------------------
# weighted_procrustes (numpy path)          slam/common/registration.py:15-76
# --------------------------------------------------------------------------------------
def weighted_procrustes(pc_target: np.ndarray, pc_reference: np.ndarray, weights=None) -> np.ndarray:
    """Rigid T (float64 4x4) with T * target ~ reference.  The weights only enter the two centroids; the
    cross-covariance sum (ref - mu_ref)(tgt - mu_tgt)^T is unweighted (registration.py:44-46)."""
    if weights is None:
        weights = np.ones((pc_target.shape[0], 1), dtype=pc_target.dtype)
    aw = weights / weights.sum(axis=0)
    mu_t = (pc_target * aw).sum(axis=0).reshape(1, 3)
    mu_r = (pc_reference * aw).sum(axis=0).reshape(1, 3)
    Cm = (pc_reference - mu_r).T.astype(np.float64) @ (pc_target - mu_t).astype(np.float64)
    U, _, Vt = np.linalg.svd(Cm)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    T = np.eye(4)
    T[:3, :3] = U @ S @ Vt
    T[:3, 3] = mu_r.astype(np.float64).reshape(3) - T[:3, :3] @ mu_t.astype(np.float64).reshape(3)
    return T


# --------------------------------------------------------------------------------------
# _PointToPlaneLossModule.point_to_plane_loss      slam/training/loss_modules.py:51-104
#   (the other caller of the projection / normal-map helpers)
# --------------------------------------------------------------------------------------
def _cost_and_slope(scheme: str, sigma: float, a: np.ndarray, d2: np.ndarray):
    """C(a) of _LS_SCHEME[scheme].cost (optimization.py:61-208) and dC/da, float64; `d2` = |p' - q|^2 feeds the
    neighbourhood weights.  Returns (C, dC/da, dC/d(d2))."""
    zero = np.zeros_like(a)
    if scheme in ("default", "least_square"):
        return a * a, 2 * a, zero
    if scheme == "huber":
        quad = a < sigma
        return np.where(quad, a * a, 2 * sigma * a - sigma ** 2), np.where(quad, 2 * a, 2 * sigma), zero
    if scheme == "exp":
        e = np.exp(-a * a / sigma ** 2)
        return a * a * e, 2 * a * e * (1 - a * a / sigma ** 2), zero
    if scheme == "neighborhood":
        w = np.exp(-d2 / sigma ** 2)
        return a * a * w, 2 * a * w, -a * a * w / sigma ** 2
    if scheme == "geman_mcclure":
        return sigma * a * a / (sigma + a * a), 2 * sigma ** 2 * a / (sigma + a * a) ** 2, zero
    if scheme == "square_geman_mcclure":
        return a * a * (sigma / (sigma + a * a)) ** 2, 2 * a * sigma ** 2 * (sigma - a * a) / (sigma + a * a) ** 3, zero
    if scheme == "cauchy":
        return np.log(1 + (a / sigma) ** 2), 2 * a / (sigma ** 2 + a * a), zero
    raise AssertionError(f"unknown scheme {scheme}")


def p2plane_training_loss(vm_target: torch.Tensor, vm_reference: torch.Tensor, nm_reference: torch.Tensor,
                          pose_mats: torch.Tensor, projector: "orc.Projector", scheme="geman_mcclure", sigma=0.5):
    """loss = mean_b( sum_pix C(|r|)^2 / sum_pix mask ) with r = n . (q - p'), p' the transformed target point that wins
    pixel `pix` of the re-projected target map (loss_modules.py:76-104), and its ANALYTIC gradient with respect to the
    top three rows of the pose matrices (the reference gets it from autograd: values flow through the scatter -- to
    every point written to a pixel, not only the surviving one -- the rounded pixel coordinates carry none).  float32 where the reference's rounding decides something (transform,
    projection, masks), float64 for the sums.  Returns (loss, per-batch losses [B], grad [B,4,4])."""
    B, _, H, W = vm_target.shape
    pts = vm_target.permute(0, 2, 3, 1).reshape(B, H * W, 3)
    alive = (pts.norm(dim=2, keepdim=True) != 0.0)
    moved = orc.apply_transformation(pts, pose_mats) * alive
    index = torch.arange(H * W, dtype=torch.float32).reshape(1, -1, 1).expand(B, -1, -1)
    vm_t = projector.build_projection_map(moved, channels=torch.cat([moved, index + 1.0], dim=2), height=H, width=W)
    pt = vm_t[:, :3].permute(0, 2, 3, 1).reshape(B, H * W, 3).numpy().astype(np.float64)
    win = vm_t[:, 3].reshape(B, H * W).numpy().astype(np.int64) - 1  # -1 = empty pixel
    q = vm_reference.permute(0, 2, 3, 1).reshape(B, H * W, 3).numpy()
    n = nm_reference.permute(0, 2, 3, 1).reshape(B, H * W, 3).numpy()
    mask = (np.linalg.norm(n, axis=-1) != 0) & (np.linalg.norm(q, axis=-1) != 0) & (np.linalg.norm(pt.astype(np.float32), axis=-1) != 0)
    q, n = q.astype(np.float64), n.astype(np.float64)
    r = ((q - pt) * n).sum(-1)
    a = np.abs(r) * mask
    d2 = ((pt - q) ** 2).sum(-1)
    Cv, dCa, dCd2 = _cost_and_slope(scheme, sigma, a, d2)
    M = mask.sum(1).astype(np.float64)
    per_batch = (Cv * Cv).sum(1) / M
    # d(C^2)/dp' = 2 C (dC/da * sign(r) * (-n) + dC/d(d2) * 2 (p' - q)), masked
    g = (2 * Cv * (dCa * np.sign(r) * mask))[..., None] * (-n) + (2 * Cv * dCd2 * mask)[..., None] * 2 * (pt - q)
    g = g / (M[:, None, None] * B)
    # Backward of the z-buffer scatter `image[b, :, row, col] = values[b, order, :]` (projection.py:415): autograd's
    # index_put backward hands EVERY scattered point the gradient of the pixel it was written to, including the
    # points a closer one overwrote -- so all points landing in a pixel contribute g[pixel], each with its own
    # coordinates in dL/dR.  Restated as is: the gradient must be the reference's.
    row, col = projector.pixels(moved, H, W)
    prow, pcol = row.round(), col.round()
    landed = ((prow >= 0.0) & (prow <= H - 1) & (pcol >= 0.0) & (pcol <= W - 1) & (moved.norm(dim=2) > 0.0)).numpy()
    pix = (prow.long() * W + pcol.long()).numpy()
    src = pts.numpy().astype(np.float64)
    grad = np.zeros((B, 4, 4))
    for b in range(B):
        gp = g[b][pix[b][landed[b]]]            # gradient of the pixel each landed point was written to
        grad[b, :3, 3] = gp.sum(0)
        grad[b, :3, :3] = gp.T @ src[b][landed[b]]
    return float(per_batch.mean()), per_batch, grad


def pose_matrix_grad_to_params(params: np.ndarray, grad_mats: np.ndarray) -> np.ndarray:
    """Chain rule through Pose.build_pose_matrix (pose.py:120-144, rotation.py:166-184): [B,4,4] -> [B,6]."""
    dR = orc.euler_jacobian(torch.from_numpy(np.asarray(params, np.float64)[:, 3:])).numpy()  # [B,3,3,3]
    out = np.zeros((params.shape[0], 6))
    out[:, :3] = grad_mats[:, :3, 3]
    for k in range(3):
        out[:, 3 + k] = (grad_mats[:, :3, :3] * dR[:, k]).sum((1, 2))
    return out
