"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the ICP-odometry hot path.

A from-scratch restatement (torch CPU / numpy / scipy) of the algorithm of
Kitware/pyLiDAR-SLAM's frame-to-model ICP odometry.  It is the *checker* for
the CUDA path and the ``cpu_baseline`` / ``--impl reference`` arm of
``bench.py``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU legs may import it; the product path (``pylidar_slam_b200``) never does.

Parity status: PINNED against the unmodified reference, imported in the build
container under ``oracle/ref_shims.py``; golden vectors and the generating
script live in ``tests/golden``.  One boundary is unpinned by the reference
itself: ``pykdtree`` (requirements.txt:5, unpinned version, not vendored) is
replaced by ``scipy.spatial.cKDTree`` on both sides -- both are exact k-NN
searches, so they agree except on sub-ulp distance ties.

Each function cites the reference file:line it restates (paths relative to
the reference root).  The operation order follows the reference where float32
rounding matters (projection pixels, z-buffer order, weights clamp).
"""
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch
from scipy.spatial import cKDTree

_KD_WORKERS = -1  # pykdtree is OpenMP-parallel; give the stand-in all cores too


# --------------------------------------------------------------------------------------
# a1 -- voxel-grid subsample          slam/common/pointcloud.py:13-23,40-79,170-195
# --------------------------------------------------------------------------------------
HASH_PX, HASH_PY, HASH_PZ = 73856093, 19349669, 83492791


def voxel_coords(points: np.ndarray, voxel: float) -> np.ndarray:
    """int64 round-half-even of p / voxel in float64 (pointcloud.py:54-79)."""
    return np.rint(points.astype(np.float64) / float(voxel)).astype(np.int64)


def voxel_hashes(coords: np.ndarray) -> np.ndarray:
    """Signed 64-bit linear hash, no xor / modulo (pointcloud.py:13-23,40-51)."""
    c = coords.astype(np.int64)
    return HASH_PX * c[:, 0] + HASH_PY * c[:, 1] + HASH_PZ * c[:, 2]


def grid_sample(points: np.ndarray, voxel: float):
    """One point per distinct hash: first occurrence, ascending hash order
    (pointcloud.py:170-195, preprocessing.py:213-226)."""
    h = voxel_hashes(voxel_coords(points, voxel))
    order = np.argsort(h, kind="stable")
    hs = h[order]
    head = np.ones(hs.shape[0], dtype=bool)
    head[1:] = hs[1:] != hs[:-1]
    idx = order[head]
    return points[idx], idx.astype(np.int64)


# --------------------------------------------------------------------------------------
# a16 -- pose algebra                 slam/common/rotation.py:144-150,253-270; pose.py:120-207
# --------------------------------------------------------------------------------------
def euler_to_mat(e: torch.Tensor) -> torch.Tensor:
    """R = Rz(ez) Ry(ey) Rx(ex), e = [B,3] (rotation.py:144-150)."""
    c, s = torch.cos(e), torch.sin(e)
    B = e.shape[0]

    def mat(rows):
        return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)

    one, zero = torch.ones(B, dtype=e.dtype), torch.zeros(B, dtype=e.dtype)
    rx = mat([[one, zero, zero], [zero, c[:, 0], -s[:, 0]], [zero, s[:, 0], c[:, 0]]])
    ry = mat([[c[:, 1], zero, s[:, 1]], [zero, one, zero], [-s[:, 1], zero, c[:, 1]]])
    rz = mat([[c[:, 2], -s[:, 2], zero], [s[:, 2], c[:, 2], zero], [zero, zero, one]])
    return rz @ ry @ rx


def mat_to_euler(R: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """Inverse of euler_to_mat with the gimbal branch sy < eps (rotation.py:253-270)."""
    sy = torch.sqrt(R[:, 0, 0] * R[:, 0, 0] + R[:, 1, 0] * R[:, 1, 0])
    sing = sy < eps
    x = torch.where(sing, torch.atan2(-R[:, 1, 2], R[:, 1, 1]), torch.atan2(R[:, 2, 1], R[:, 2, 2]))
    y = torch.atan2(-R[:, 2, 0], sy)
    z = torch.where(sing, torch.zeros_like(sy), torch.atan2(R[:, 1, 0], R[:, 0, 0]))
    return torch.stack([x, y, z], dim=1)


def build_pose_matrix(params: torch.Tensor) -> torch.Tensor:
    """[B,6] (tx,ty,tz,ex,ey,ez) -> [B,4,4] (pose.py:120-144)."""
    B = params.shape[0]
    T = torch.zeros(B, 4, 4, dtype=params.dtype)
    T[:, :3, :3] = euler_to_mat(params[:, 3:])
    T[:, :3, 3] = params[:, :3]
    T[:, 3, 3] = 1.0
    return T


def from_pose_matrix(T: torch.Tensor) -> torch.Tensor:
    """[B,4,4] -> [B,6] (pose.py:188-207)."""
    return torch.cat([T[:, :3, 3], mat_to_euler(T[:, :3, :3])], dim=1)


def apply_transformation(points: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """P R^T + t for [B,N,3] points, [B,4,4] poses (pose.py:169-186)."""
    return torch.matmul(points, T[:, :3, :3].transpose(1, 2)) + T[:, :3, 3].unsqueeze(1)


def euler_jacobian(e: torch.Tensor) -> torch.Tensor:
    """d R / d e_k, [B,3,3,3] (rotation.py:166-184)."""
    c, s = torch.cos(e), torch.sin(e)
    B = e.shape[0]
    one, zero = torch.ones(B, dtype=e.dtype), torch.zeros(B, dtype=e.dtype)

    def mat(rows):
        return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)

    rx = mat([[one, zero, zero], [zero, c[:, 0], -s[:, 0]], [zero, s[:, 0], c[:, 0]]])
    ry = mat([[c[:, 1], zero, s[:, 1]], [zero, one, zero], [-s[:, 1], zero, c[:, 1]]])
    rz = mat([[c[:, 2], -s[:, 2], zero], [s[:, 2], c[:, 2], zero], [zero, zero, one]])
    jx = mat([[zero, zero, zero], [zero, -s[:, 0], -c[:, 0]], [zero, c[:, 0], -s[:, 0]]])
    jy = mat([[-s[:, 1], zero, c[:, 1]], [zero, zero, zero], [-c[:, 1], zero, -s[:, 1]]])
    jz = mat([[-s[:, 2], -c[:, 2], zero], [c[:, 2], -s[:, 2], zero], [zero, zero, zero]])
    return torch.stack([rz @ ry @ jx, rz @ jy @ rx, jz @ ry @ rx], dim=1)


# --------------------------------------------------------------------------------------
# a2/a3 -- spherical projection + closest-wins z-buffer      slam/common/projection.py:11-73,331-418
# --------------------------------------------------------------------------------------
@dataclass
class Projector:
    height: int
    width: int
    up_fov: float = 3.0
    down_fov: float = -24.0

    def pixels(self, xyz: torch.Tensor, height=None, width=None):
        """Float pixel coordinates (row, col) of [B,N,3] points (projection.py:11-73)."""
        H = self.height if height is None else height
        W = self.width if width is None else width
        up = self.up_fov / 180.0 * np.pi
        down = self.down_fov / 180.0 * np.pi
        fov = abs(down) + abs(up)
        r = torch.norm(xyz, p=2, dim=2)
        null = (r == 0.0).to(xyz.dtype)
        ok = 1.0 - null
        r = null * 0.001 + ok * r
        theta = -torch.atan2(xyz[:, :, 1], xyz[:, :, 0])
        phi = torch.asin(xyz[:, :, 2] / r)
        col = 0.5 * (theta / np.pi + 1.0)
        row = 1.0 - (phi + abs(down)) / fov
        col = col * W
        row = row * H
        return row * ok - null, col * ok - null

    def build_projection_map(self, xyz: torch.Tensor, channels: Optional[torch.Tensor] = None,
                             height=None, width=None) -> torch.Tensor:
        """[B,N,3] (+ optional [B,N,C] channels) -> [B,C,H,W]; the closest point per
        pixel survives; empty pixels are 0 (projection.py:331-418)."""
        H = self.height if height is None else height
        W = self.width if width is None else width
        B, N, _ = xyz.shape
        values = xyz if channels is None else channels
        C = values.shape[2]
        out = torch.zeros(B, C, H, W, dtype=xyz.dtype)
        row, col = self.pixels(xyz[:, :, :3], H, W)
        r = xyz.norm(dim=2)
        prow, pcol = row.round(), col.round()
        bad = ~((prow >= 0.0) & (prow <= H - 1) & (pcol >= 0.0) & (pcol <= W - 1))
        r = r.clone()
        r[bad] = -1.0
        # The reference sorts by descending range and scatters, so the closest point is
        # written last.  Restated deterministically: ascending stable sort, first hit per
        # pixel wins (ties on range -> lowest point index).
        order = torch.argsort(r, dim=1, descending=False, stable=True)
        flat = out.view(B, C, H * W)
        for b in range(B):
            o = order[b]
            o = o[r[b, o] > 0.0]
            pix = (prow[b, o].long() * W + pcol[b, o].long()).numpy()
            upix, first = np.unique(pix, return_index=True)
            sel = o[torch.from_numpy(first)]
            flat[b][:, torch.from_numpy(upix)] = values[b, sel, :].t()
        return out


# --------------------------------------------------------------------------------------
# a19 -- vertex-map helpers          slam/common/geometry.py:157-204; utils.py:169-196
# --------------------------------------------------------------------------------------
def map_to_points(pmap: torch.Tensor) -> torch.Tensor:
    """[K,C,H,W] -> [K,H*W,C] row-major pixels (geometry.py:181-204)."""
    K, C, H, W = pmap.shape
    return pmap.permute(0, 2, 3, 1).reshape(K, H * W, C)


def not_null(t: torch.Tensor, dim: int) -> torch.Tensor:
    """True where any entry along dim is non-zero (geometry.py:157-177)."""
    return t.abs().max(dim=dim, keepdim=True)[0] > 0


# --------------------------------------------------------------------------------------
# a4 -- box-filter normal map        slam/common/geometry.py:65-114,240-295
# --------------------------------------------------------------------------------------
def normal_map(vmap: torch.Tensor, kernel_size: int = 5) -> torch.Tensor:
    """[B,3,H,W] vertex map -> unit normals n ~ (sum v v^T)^-1 sum v over a zero-padded
    k x k window, via the cofactor matrix; 0 where |det| <= 1e-6 or the vertex is null."""
    B, _, H, W = vmap.shape
    box = torch.ones(1, 1, kernel_size, kernel_size, dtype=torch.float32)
    pad = kernel_size // 2
    outer = (vmap.unsqueeze(1) * vmap.unsqueeze(2)).reshape(B * 9, 1, H, W)
    sv = torch.nn.functional.conv2d(vmap.reshape(B * 3, 1, H, W), box, padding=(pad, pad))
    sv = sv.reshape(B, 3, H, W).permute(0, 2, 3, 1)  # [B,H,W,3]
    A = torch.nn.functional.conv2d(outer, box, padding=(pad, pad)).reshape(B, 3, 3, H, W).permute(0, 3, 4, 1, 2)
    cof = torch.empty_like(A)
    for i in range(3):
        cof[..., i, :] = torch.cross(A[..., i - 2, :], A[..., i - 1, :], dim=-1)
    det = (cof * A).sum(-1).mean(-1)  # mean of the three row expansions
    keep = det.abs()[..., None, None] > 1e-6
    safe = torch.where(keep, det[..., None, None], torch.ones_like(det[..., None, None]))
    inv_t = torch.where(keep.expand_as(cof), cof / safe, torch.zeros_like(cof))
    n = torch.einsum("...ij,...j->...i", inv_t.transpose(-1, -2), sv)
    good = det.abs() > 1e-6
    nm = n[good]
    norms = nm.norm(dim=1, keepdim=True)
    norms = norms + (norms == 0.0).to(torch.float32)
    n[good] = nm / norms
    n[~good] = 0.0
    n[vmap.norm(dim=1) == 0.0] = 0.0
    return n.permute(0, 3, 1, 2)


# --------------------------------------------------------------------------------------
# a6 -- projective data association  slam/common/geometry.py:397-439
# --------------------------------------------------------------------------------------
def compute_neighbors(tgt: torch.Tensor, ref: torch.Tensor, fields: Optional[torch.Tensor] = None):
    """Per pixel: the reference map (of K) whose vertex is closest to the target vertex."""
    tgt_ok = not_null(tgt, 1)
    ref_ok = not_null(ref, 1)
    inf = float("inf")
    d = (tgt - ref).norm(dim=1, keepdim=True)
    d = d + torch.where(ref_ok, 0.0, inf) + torch.where(tgt_ok, 0.0, inf)
    _, k = torch.min(d, dim=0, keepdim=True)
    nb = torch.gather(ref, 0, k.expand(1, *ref.shape[1:])).clone()
    nb[~tgt_ok.expand_as(nb)] = 0.0
    nf = None
    if fields is not None:
        nf = torch.gather(fields, 0, k.expand(1, *fields.shape[1:]))
    return nb, nf


# --------------------------------------------------------------------------------------
# a13 -- robust weights              slam/common/optimization.py:45-50,61-226
# --------------------------------------------------------------------------------------
SCHEMES = ("default", "least_square", "huber", "exp", "neighborhood", "geman_mcclure",
           "square_geman_mcclure", "cauchy")


def ls_weights(scheme: str, sigma: float, r: torch.Tensor, tgt=None, ref=None, eps: float = 1e-4):
    """w = sqrt(cost(r)) / max(|r|, eps); `default`/`least_square` short-circuit to 1."""
    if scheme in ("default", "least_square"):
        return torch.ones([1] * r.dim(), dtype=r.dtype)
    a = r.abs()
    if scheme == "huber":
        quad = a < sigma
        cost = quad * (r * r) + ~quad * (2 * sigma * a - sigma ** 2)
    elif scheme == "exp":
        cost = (r * r) * torch.exp(-r ** 2 / sigma ** 2)
    elif scheme == "neighborhood":
        cost = r * r * torch.exp(-(tgt - ref).norm(dim=-1) ** 2 / sigma ** 2)
    elif scheme == "geman_mcclure":
        r2 = r ** 2
        cost = sigma * r2 / (sigma + r2)
    elif scheme == "square_geman_mcclure":
        r2 = r ** 2
        cost = r2 * (sigma / (sigma + r2)) ** 2
    elif scheme == "cauchy":
        cost = torch.log(1 + (r / sigma) ** 2)
    else:
        raise AssertionError(f"unknown scheme {scheme}")
    return cost.sqrt() / a.clamp(eps, float("inf"))


# --------------------------------------------------------------------------------------
# a11/a12/a14/a15 -- point-to-plane Gauss-Newton   optimization.py:296-344,356-435; alignment.py:91-127
# --------------------------------------------------------------------------------------
class SingularHessian(RuntimeError):
    """optimization.py:334-336 -> RuntimeError("Invalid Jacobian in Gauss Newton minimization")."""


def p2plane_residual(x, tgt, ref, nrm):
    T = build_pose_matrix(x)
    return ((apply_transformation(tgt, T) - ref) * nrm).sum(dim=-1)


def p2plane_jacobian(x, tgt, nrm):
    B, N, _ = tgt.shape
    dR = euler_jacobian(x[:, 3:])  # [B,3,3,3]
    J = torch.zeros(B, N, 6, dtype=tgt.dtype)
    J[:, :, :3] = nrm
    for k in range(3):
        J[:, :, 3 + k] = (torch.einsum("bij,bnj->bni", dR[:, k], tgt) * nrm).sum(-1)
    return J


def gauss_newton_p2plane(ref, tgt, nrm, scheme="default", sigma=0.5, max_iters=1,
                         norm_stop=1e-3, x0=None):
    """Returns (x [B,6], (w r)^2 [B,N], status) with status in {"ok","tiny_residual"};
    raises SingularHessian on |det H| < 1e-7."""
    B = ref.shape[0]
    x = torch.zeros(B, 6, dtype=ref.dtype) if x0 is None else x0
    res = None
    for _ in range(max(max_iters, 1)):
        J = p2plane_jacobian(x, tgt, nrm)
        res = p2plane_residual(x, tgt, ref, nrm)
        if res.norm() < 1e-7:
            return x, res * res, "tiny_residual"
        w = ls_weights(scheme, sigma, res, tgt, ref)
        res = res * w
        J = J * w.unsqueeze(-1)
        Jt = J.permute(0, 2, 1)
        Hm = Jt @ J
        if torch.any(Hm.det().abs() < 1e-7):
            raise SingularHessian("Invalid Jacobian in Gauss Newton minimization")
        dx = -Hm.inverse() @ Jt @ res.unsqueeze(-1)
        x = x + dx[:, :, 0]
        if dx.norm() < norm_stop:
            break
    return x, res * res, "ok"


def align_p2plane(ref, tgt, nrm, scheme="default", sigma=0.5, max_iters=1, norm_stop=1e-3):
    """alignment.py:91-127 -> (dT [B,4,4], delta [B,6], loss [B,N])."""
    x, loss, _ = gauss_newton_p2plane(ref, tgt, nrm, scheme, sigma, max_iters, norm_stop)
    return build_pose_matrix(x), x, loss


# --------------------------------------------------------------------------------------
# a7/a8/a9 -- kd-tree local map      slam/odometry/local_map.py:254-427
# --------------------------------------------------------------------------------------
class KdTreeLocalMap:
    def __init__(self, local_map_size=20, num_neighbors_normals=10):
        self.size = local_map_size
        self.k = num_neighbors_normals
        self.init()

    def init(self):
        self.points = None
        self.counts = []
        self.normals = None
        self.tree = None

    def update(self, rel_pose: np.ndarray, new_points: Optional[np.ndarray] = None,
               new_vertex_map: Optional[torch.Tensor] = None):
        """rel_pose [4,4] f32; new points either raw [n,3] or a vertex map (pixels with
        |p| <= 0.01 dropped) (local_map.py:302-362)."""
        pts = None
        if new_points is not None:
            pts = np.asarray(new_points, dtype=np.float32).reshape(-1, 3)
        elif new_vertex_map is not None:
            flat = new_vertex_map[0].permute(1, 2, 0).reshape(-1, 3)
            pts = flat[flat.norm(dim=-1) > 0.01].numpy()
        if pts is not None:
            pts = pts[~np.isnan(pts).any(axis=1)]
        if self.points is None:
            self.points = pts
            self.counts.append(0 if pts is None else pts.shape[0])
        else:
            inv = np.linalg.inv(rel_pose)
            moved = np.einsum("ij,nj->ni", inv[:3, :3], self.points) + inv[:3, 3].reshape(1, 3)
            if pts is not None:
                self.points = np.concatenate([moved, pts], axis=0)
                self.counts.append(pts.shape[0])
            else:
                self.points = moved
            if len(self.counts) > self.size:
                self.points = self.points[self.counts.pop(0):]
        self._rebuild()

    def _rebuild(self):
        """Tree rebuilt and normal cache cleared on every update (local_map.py:365-369)."""
        self.normals = np.zeros((self.points.shape[0], 4), dtype=np.float32)
        self.tree = cKDTree(self.points)

    def nearest_neighbor_search(self, queries: np.ndarray):
        """-> (neighbor points [N,3], normals [N,3], indices [N]) (local_map.py:372-422)."""
        _, idx = self.tree.query(queries, k=1, workers=_KD_WORKERS)
        return self.points[idx], self._normals_for(idx), idx

    def _normals_for(self, idx):
        todo = idx[self.normals[idx, 3] == 0.0]
        if todo.shape[0] > 0:
            centre = self.points[todo]
            _, nb = self.tree.query(centre, k=self.k + 1, workers=_KD_WORKERS)
            nb = nb[:, 1:]
            d = self.points[nb.reshape(-1)].reshape(-1, self.k, 3) - centre[:, None, :]
            cov = (d[:, :, :, None] * d[:, :, None, :]).mean(axis=1)
            _, _, vh = np.linalg.svd(cov)
            self.normals[todo, :3] = vh[:, 2, :]
            self.normals[todo, 3] = 1.0
        return self.normals[idx, :3]


# --------------------------------------------------------------------------------------
# a5/a6 -- projective local map      slam/odometry/local_map.py:91-240
# --------------------------------------------------------------------------------------
class ProjectiveLocalMap:
    def __init__(self, projector: Projector, local_map_size=20, normals_kernel_size=5):
        self.projector = projector
        self.size = local_map_size
        self.ksize = normals_kernel_size
        self.init()

    def init(self):
        self.vmaps = None
        self.nmaps = None
        self.masks = None
        self.poses = None
        self.model_vmap = None
        self.model_nmap = None

    def update(self, rel_pose: torch.Tensor, new_vertex_map: Optional[torch.Tensor] = None,
               mask: Optional[torch.Tensor] = None):
        """rel_pose [1,4,4] (local_map.py:126-174)."""
        if new_vertex_map is not None:
            nmap = normal_map(new_vertex_map, self.ksize)
            if mask is None:
                mask = not_null(new_vertex_map, 1)
        if self.vmaps is None:
            self.vmaps, self.nmaps, self.masks, self.poses = new_vertex_map, nmap, mask, rel_pose
        else:
            old = rel_pose.inverse() @ self.poses
            if new_vertex_map is not None:
                self.poses = torch.cat([old, torch.eye(4, dtype=old.dtype).unsqueeze(0)], dim=0)
                self.vmaps = torch.cat([self.vmaps, new_vertex_map], dim=0)
                self.nmaps = torch.cat([self.nmaps, nmap], dim=0)
                self.masks = torch.cat([self.masks, mask], dim=0)
            else:
                self.poses = old
            if self.poses.shape[0] > self.size:
                self.vmaps, self.nmaps = self.vmaps[1:], self.nmaps[1:]
                self.poses, self.masks = self.poses[1:], self.masks[1:]
        self._rebuild()

    def _rebuild(self):
        """Re-express all K frames in the newest frame and re-project (local_map.py:177-202)."""
        _, _, H, W = self.vmaps.shape
        pts = apply_transformation(map_to_points(self.vmaps), self.poses)
        nrm = torch.einsum("bij,bnj->bni", self.poses[:, :3, :3], map_to_points(self.nmaps))
        both = torch.cat([pts, nrm], dim=2) * map_to_points(self.masks.to(pts.dtype))
        model = self.projector.build_projection_map(both[:, :, :3], channels=both, height=H, width=W)
        self.model_vmap, self.model_nmap = model[:, :3], model[:, 3:6]

    def nearest_neighbor_search(self, queries: torch.Tensor):
        """[N,3] -> (neighbor points, normals, surviving queries), each [1,Nc,3] (local_map.py:205-235)."""
        tgt = self.projector.build_projection_map(queries.unsqueeze(0))
        nb_v, nb_n = compute_neighbors(tgt, self.model_vmap, self.model_nmap)
        q = map_to_points(nb_v)
        p = map_to_points(tgt)
        keep = (not_null(p, -1) & not_null(q, -1))[:, :, 0]
        return q[keep].unsqueeze(0), map_to_points(nb_n)[keep].unsqueeze(0), p[keep].unsqueeze(0)


# --------------------------------------------------------------------------------------
# a17/a18 -- ICP frame-to-model odometry     slam/odometry/icp_odometry.py:72-380
# --------------------------------------------------------------------------------------
@dataclass
class ICPConfig:
    max_num_alignments: int = 100
    threshold_delta_pose: float = 1e-4
    threshold_trans: float = 0.1
    threshold_rot: float = 0.3
    data_key: str = "vertex_map"
    local_map: str = "kdtree"  # or "projective"
    local_map_size: int = 20
    num_neighbors_normals: int = 10
    normals_kernel_size: int = 5
    scheme: str = "default"
    sigma: float = 0.5
    gn_max_iters: int = 1
    gn_norm_stop: float = 1e-3


class ICPFrameToModelOracle:
    """process_next_frame(data_dict) with the reference's three input layouts, frame-0
    insertion, early-break semantics and key-frame policy."""

    def __init__(self, config: ICPConfig, projector: Projector):
        self.cfg = config
        self.projector = projector
        if config.local_map == "kdtree":
            self.local_map = KdTreeLocalMap(config.local_map_size, config.num_neighbors_normals)
        else:
            self.local_map = ProjectiveLocalMap(projector, config.local_map_size, config.normals_kernel_size)
        self.init()

    def init(self):
        self.local_map.init()
        self.relative_poses = []
        self.absolute_poses = []
        self.iter = 0
        self.sample_pointcloud = False
        self.delta_since_update = torch.eye(4, dtype=torch.float32).reshape(1, 4, 4)
        self.losses = []
        self.num_iters = []

    # icp_odometry.py:319-358
    def _read_input(self, data_dict):
        data = data_dict[self.cfg.data_key]
        if isinstance(data, np.ndarray):
            self.sample_pointcloud = True
            pc = torch.from_numpy(data).unsqueeze(0)
            vmap = self.projector.build_projection_map(pc)
        elif data.dim() in (3, 4):
            vmap = data if data.dim() == 4 else data.unsqueeze(0)
            pc = vmap.permute(0, 2, 3, 1).reshape(1, -1, 3)
            pc = pc[not_null(pc, -1)[:, :, 0]]
            # Reference quirk (icp_odometry.py:342-344,356-358): the masked cloud is 2-D, so
            # `_tgt_pc[0]` keeps only the FIRST non-null pixel; mirrored on purpose.
            pc = pc[:1]
        else:
            pc = data.unsqueeze(0)
            vmap = self.projector.build_projection_map(pc)
        vmap = vmap.to(torch.float32).clone()
        pc = pc.to(torch.float32)
        nan_px = torch.isnan(vmap).any(dim=1, keepdim=True).expand_as(vmap)
        vmap[nan_px] = 0.0
        pc = pc.reshape(-1, 3)
        pc = pc[~torch.isnan(pc).any(dim=1)]
        self.vmap, self.pc = vmap, pc.unsqueeze(0)

    # icp_odometry.py:301-308
    def _sample_points(self):
        if self.sample_pointcloud:
            return self.pc[0]
        pts = map_to_points(self.vmap)[0]
        return pts[pts.norm(dim=-1) > 0.0]

    def _nn(self, pts: torch.Tensor):
        if isinstance(self.local_map, KdTreeLocalMap):
            q, n, _ = self.local_map.nearest_neighbor_search(pts.numpy())
            return torch.from_numpy(q).unsqueeze(0), torch.from_numpy(n).unsqueeze(0), pts.unsqueeze(0)
        return self.local_map.nearest_neighbor_search(pts)

    # icp_odometry.py:248-299
    def register_new_frame(self, points: torch.Tensor, initial: torch.Tensor):
        T = initial
        params = torch.zeros(6, dtype=points.dtype)
        losses = []
        for _ in range(self.cfg.max_num_alignments):
            moved = apply_transformation(points.unsqueeze(0), T)[0]
            q, n, p = self._nn(moved)
            dT, delta, loss = align_p2plane(q, p, n, self.cfg.scheme, self.cfg.sigma,
                                            self.cfg.gn_max_iters, self.cfg.gn_norm_stop)
            losses.append(float(loss.sum()))
            if delta.norm() < self.cfg.threshold_delta_pose:
                break
            params = from_pose_matrix(dT @ T)
            T = build_pose_matrix(params)
        return params, T, losses

    # icp_odometry.py:360-380
    def _update_map(self, T: torch.Tensor):
        delta = self.delta_since_update @ T
        dp = from_pose_matrix(delta.reshape(1, 4, 4))
        if dp[0, :3].norm() > self.cfg.threshold_trans or \
                dp[0, 3:].norm() * 180 / np.pi > self.cfg.threshold_rot:
            if isinstance(self.local_map, KdTreeLocalMap):
                self.local_map.update(T[0].numpy(), new_points=self.pc.reshape(-1, 3).numpy())
            else:
                self.local_map.update(T, new_vertex_map=self.vmap, mask=not_null(self.vmap, 1))
            self.delta_since_update = torch.eye(4, dtype=torch.float32)
        else:
            if isinstance(self.local_map, KdTreeLocalMap):
                self.local_map.update(T[0].numpy())
            else:
                self.local_map.update(T)
            self.delta_since_update = delta

    # icp_odometry.py:157-246
    def process_next_frame(self, data_dict: dict):
        self._read_input(data_dict)
        if self.iter == 0:
            eye = torch.eye(4, dtype=torch.float32).unsqueeze(0)
            if isinstance(self.local_map, KdTreeLocalMap):
                self.local_map.update(eye[0].numpy(), new_vertex_map=self.vmap)
            else:
                self.local_map.update(eye, new_vertex_map=self.vmap)
            self.relative_poses.append(eye.numpy())
            self.absolute_poses.append(eye.to(torch.float64).numpy()[0])
            self.iter += 1
            return
        init = data_dict.get("init_rpose", None)
        init = np.eye(4) if init is None else init
        init = torch.from_numpy(np.asarray(init)).to(torch.float32).reshape(1, 4, 4)
        params, T, losses = self.register_new_frame(self._sample_points(), init)
        self.losses.append(losses)
        self._update_map(T)
        self.relative_poses.append(T.numpy())
        step = build_pose_matrix(params.to(torch.float64).reshape(1, 6))[0].numpy()
        self.absolute_poses.append(self.absolute_poses[-1].dot(step))
        data_dict["odometry_pc"] = data_dict.get("distorted", self.pc.numpy().reshape(-1, 3))
        data_dict["odometry_pose"] = T.numpy().reshape(4, 4)
        self.iter += 1

    def get_relative_poses(self):
        return np.concatenate(self.relative_poses, axis=0) if self.relative_poses else None
