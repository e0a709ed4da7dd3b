"""Host-side mirrors of the reference's `slam/common` helpers on the hot path, each a thin call
into the C ABI (names, argument meaning and error behaviour follow the reference; arithmetic
runs in the CUDA library).

Arrays may be numpy arrays, CPU torch tensors or CUDA torch tensors; outputs are returned in the
same container kind / device as the first input.
"""
import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib

_default_ctx = None


def default_context() -> _lib.Context:
    """A lazily created context on cuda:0 for the stateless helpers."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _lib.Context()
    return _default_ctx


def assert_debug(condition: bool, message: str = ""):
    """slam/common/utils.py:30-38"""
    if not condition:
        raise AssertionError(message)


def check_tensor(tensor, sizes: list):
    """slam/common/utils.py:54-74 (shape check; -1 matches any size)."""
    shape = tensor.shape
    ok = len(shape) == len(sizes)
    if ok:
        for s, t in zip(sizes, shape):   # a plain loop: this runs several times per frame
            if s != -1 and s != t:
                ok = False
                break
    if not ok:
        raise AssertionError(f"[BAD TENSOR SHAPE] Wrong tensor shape got {tuple(tensor.shape)} expected {sizes}")


def _as_f32(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.float32)
    return x.to(torch.float32).contiguous()


def _empty_like_kind(x, shape, dtype=np.float32):
    if isinstance(x, np.ndarray):
        return np.empty(shape, dtype=dtype)
    tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int64: torch.int64}[dtype]
    return torch.empty(shape, dtype=tdt, device=x.device)


# ------------------------------------------------------------------------------------------
# slam/common/pointcloud.py
# ------------------------------------------------------------------------------------------
def voxelise(pointcloud, voxel_x: float = 0.2, voxel_y: float = -1.0, voxel_z: float = -1.0, ctx=None):
    """int64 voxel coordinates `(n, 3)` (pointcloud.py:54-79); voxel_y / voxel_z default to voxel_x."""
    ctx = ctx or default_context()
    voxel_y = voxel_x if voxel_y == -1.0 else voxel_y
    voxel_z = voxel_x if voxel_z == -1.0 else voxel_z
    check_tensor(pointcloud, [-1, 3])
    is64 = (pointcloud.dtype == np.float64) if isinstance(pointcloud, np.ndarray) else (pointcloud.dtype == torch.float64)
    pc = np.ascontiguousarray(pointcloud) if isinstance(pointcloud, np.ndarray) else pointcloud.contiguous()
    if not is64:
        pc = _as_f32(pc)
    n = pc.shape[0]
    coords = _empty_like_kind(pc, (n, 3), np.int64)
    hashes = _empty_like_kind(pc, (n,), np.int64)
    ctx.call("pls_voxel_hash_xyz", _lib.ptr(pc), int(is64), n, float(voxel_x), float(voxel_y), float(voxel_z),
             _lib.ptr(coords), _lib.ptr(hashes))
    return coords


def voxel_hashing(pointcloud, voxel_size: float, ctx=None):
    """Signed 64-bit voxel hashes `(n,)` of a point cloud (pointcloud.py:13-23,40-51 fused with voxelise)."""
    ctx = ctx or default_context()
    check_tensor(pointcloud, [-1, 3])
    is64 = (pointcloud.dtype == np.float64) if isinstance(pointcloud, np.ndarray) else (pointcloud.dtype == torch.float64)
    pc = np.ascontiguousarray(pointcloud) if isinstance(pointcloud, np.ndarray) else pointcloud.contiguous()
    if not is64:
        pc = _as_f32(pc)
    hashes = _empty_like_kind(pc, (pc.shape[0],), np.int64)
    ctx.call("pls_voxel_hash", _lib.ptr(pc), int(is64), pc.shape[0], float(voxel_size), None, _lib.ptr(hashes))
    return hashes


def grid_sample(pointcloud, voxel_size: float, ctx=None):
    """One point per voxel hash: `(sample_points, sample_indices)` (pointcloud.py:182-195)."""
    ctx = ctx or default_context()
    check_tensor(pointcloud, [-1, 3])
    is64 = (pointcloud.dtype == np.float64) if isinstance(pointcloud, np.ndarray) else (pointcloud.dtype == torch.float64)
    pc = np.ascontiguousarray(pointcloud) if isinstance(pointcloud, np.ndarray) else pointcloud.contiguous()
    if not is64:
        pc = _as_f32(pc)
    n = pc.shape[0]
    count = C.c_int64(0)
    if isinstance(pc, np.ndarray):
        # host caller (GridSample.filter): the gather kernel writes the result into pinned host memory -- one
        # synchronisation, no pageable device->host copy -- and the library keeps a device-resident twin, which is
        # published so that ICPFrameToModel can consume it without sending the samples back to the device
        fdt = np.dtype(np.float64 if is64 else np.float32)
        xyz_buf, idx_buf = _lib.PinnedPool.take(n * 3 * fdt.itemsize), _lib.PinnedPool.take(n * 8)
        dx = C.c_void_p()
        if xyz_buf is not None and idx_buf is not None:
            # the kernel writes into pinned buffers of the pool; the arrays handed out are views of them (zero copies)
            hx, hi = C.c_void_p(_lib.ptr(xyz_buf)), C.c_void_p(_lib.ptr(idx_buf))
            ctx.call("pls_grid_sample_staged", _lib.ptr(pc), int(is64), n, float(voxel_size), C.byref(hx), C.byref(hi),
                     C.byref(dx), C.byref(count))
            S = count.value
            out = xyz_buf[:S * 3 * fdt.itemsize].view(fdt).reshape(S, 3)
            idx = idx_buf[:S * 8].view(np.int64)
        else:
            # pool exhausted (the caller keeps many frames alive): the library's own staging, copied out
            del xyz_buf, idx_buf
            hx, hi = C.c_void_p(), C.c_void_p()
            ctx.call("pls_grid_sample_staged", _lib.ptr(pc), int(is64), n, float(voxel_size), C.byref(hx), C.byref(hi),
                     C.byref(dx), C.byref(count))
            S = count.value
            out = _lib.host_view(hx.value, (S, 3), fdt).copy()
            idx = _lib.host_view(hi.value, (S,), np.int64).copy()
        _lib.Handoff.publish(out, dx.value or 0, int(ctx.cfg.device), is64)
        return out, idx
    out = _empty_like_kind(pc, (n, 3), np.float64 if is64 else np.float32)
    idx = _empty_like_kind(pc, (n,), np.int64)
    ctx.call("pls_grid_sample", _lib.ptr(pc), int(is64), n, float(voxel_size), _lib.ptr(out), _lib.ptr(idx), C.byref(count))
    return out[:count.value], idx[:count.value]


def _is64(x) -> bool:
    return x.dtype == (np.float64 if isinstance(x, np.ndarray) else torch.float64)


def voxel_statistics(pointcloud, voxel_size: float, ctx=None):
    """voxelise + voxel_hashing + voxel_normal_distribution in one pass over the sorted hashes
    (pointcloud.py:54-79,40-51,83-167): returns `(coords [n,3], hashes [n], sizes [V], means [V,3], covs [V,3,3],
    voxel_indices [n])`, voxels in ascending-hash order, covs = the un-normalised scatter matrices."""
    ctx = ctx or default_context()
    check_tensor(pointcloud, [-1, 3])
    is64 = _is64(pointcloud)
    pc = np.ascontiguousarray(pointcloud) if isinstance(pointcloud, np.ndarray) else pointcloud.contiguous()
    if not is64:
        pc = _as_f32(pc)
    n = pc.shape[0]
    assert_debug(n > 0, "cannot voxelise an empty point cloud")
    fdt = np.float64 if is64 else np.float32
    coords, hashes = _empty_like_kind(pc, (n, 3), np.int64), _empty_like_kind(pc, (n,), np.int64)
    sizes, ids = _empty_like_kind(pc, (n,), np.int64), _empty_like_kind(pc, (n,), np.int64)
    means, covs = _empty_like_kind(pc, (n, 3), fdt), _empty_like_kind(pc, (n, 3, 3), fdt)
    count = C.c_int64(0)
    ctx.call("pls_voxel_statistics", _lib.ptr(pc), int(is64), n, float(voxel_size), _lib.ptr(coords), _lib.ptr(hashes),
             _lib.ptr(sizes), _lib.ptr(means), _lib.ptr(covs), _lib.ptr(ids), C.byref(count))
    V = count.value
    return coords, hashes, sizes[:V], means[:V], covs[:V], ids


def voxel_normal_distribution(pointcloud, voxel_size: float, ctx=None):
    """`(voxel_sizes, means, covs, voxel_ids)` of voxel_normal_distribution (pointcloud.py:154-167).  The reference
    takes precomputed hashes; here they come from the same pass, so the argument is the voxel size."""
    _, _, sizes, means, covs, ids = voxel_statistics(pointcloud, voxel_size, ctx=ctx)
    return sizes, means, covs, ids


def distort_frame(pointcloud, timestamps, relative_pose, ctx=None):
    """The arithmetic of Distortion.filter (preprocessing.py:171-191): `Slerp(I -> R)(alpha) p + alpha t` with
    `alpha = (t - min t) / (max t - min t)`; float64 `[n,3]` out."""
    ctx = ctx or default_context()
    check_tensor(pointcloud, [-1, 3])
    n = pointcloud.shape[0]
    check_tensor(timestamps, [n])
    check_tensor(relative_pose, [4, 4])
    is_np = isinstance(pointcloud, np.ndarray)
    pc64 = _is64(pointcloud)
    pc = (np.ascontiguousarray(pointcloud) if is_np else pointcloud.contiguous()) if pc64 else _as_f32(pointcloud)
    ts64 = _is64(timestamps)
    ts = (np.ascontiguousarray(timestamps) if isinstance(timestamps, np.ndarray) else timestamps.contiguous()) if ts64 \
        else _as_f32(timestamps)
    pose = np.asarray(relative_pose.detach().cpu().numpy() if isinstance(relative_pose, torch.Tensor) else relative_pose)
    pose64 = pose.dtype == np.float64
    pose = np.ascontiguousarray(pose, dtype=np.float64 if pose64 else np.float32)
    out = _empty_like_kind(pc, (n, 3), np.float64)
    ctx.call("pls_distort", _lib.ptr(pc), int(pc64), _lib.ptr(ts), int(ts64), n, _lib.ptr(pose), int(pose64), _lib.ptr(out))
    return out


# ------------------------------------------------------------------------------------------
# slam/common/registration.py
# ------------------------------------------------------------------------------------------
def weighted_procrustes(pc_target, pc_reference, weights=None, ctx=None):
    """Rigid transform (float64 `[4,4]` numpy) with `T * target ~ reference` (registration.py:15-76, numpy path:
    `[n,3]` clouds, optional `[n,1]` weights that only enter the centroids)."""
    ctx = ctx or default_context()
    check_tensor(pc_target, [-1, 3])
    check_tensor(pc_reference, [*pc_target.shape])
    is64 = _is64(pc_target)
    conv = (lambda a: (np.ascontiguousarray(a, dtype=np.float64) if isinstance(a, np.ndarray) else a.to(torch.float64).contiguous())) \
        if is64 else _as_f32
    tgt, ref = conv(pc_target), conv(pc_reference)
    w = None
    if weights is not None:
        w = conv(weights.reshape(-1))
        check_tensor(w, [tgt.shape[0]])
    out = np.empty((4, 4), dtype=np.float64)
    ctx.call("pls_weighted_procrustes", _lib.ptr(tgt), _lib.ptr(ref), _lib.ptr(w), tgt.shape[0], int(is64), _lib.ptr(out))
    return out


# ------------------------------------------------------------------------------------------
# slam/common/projection.py
# ------------------------------------------------------------------------------------------
class SphericalProjector:
    """SphericalProjector (projection.py:426-508): spherical range-image projection."""

    def __init__(self, height: Optional[int] = None, width: Optional[int] = None, num_channels: Optional[int] = None,
                 up_fov: Optional[float] = None, down_fov: Optional[float] = None, conversion=None, ctx=None, **kwargs):
        self.height, self.width, self.num_channels = height, width, num_channels
        self.up_fov, self.down_fov = up_fov, down_fov
        # the image channels of build_projection_map when no `transform` is passed: the reference's member transform,
        # `xyz_conversion` by default (projection.py:76-95,439-446) -- the first three channels of the cloud
        self.conversion = conversion
        self._ctx = ctx

    @property
    def ctx(self):
        return self._ctx or default_context()

    def project_pointcloud(self, pointcloud, height=None, width=None, up_fov=None, down_fov=None, **kwargs):
        """[B,N,K>=3] -> float pixel coordinates [B,N,2] (row, col) (projection.py:452-484)."""
        check_tensor(pointcloud, [-1, -1, -1])
        H, W = height or self.height, width or self.width
        up, down = (self.up_fov if up_fov is None else up_fov), (self.down_fov if down_fov is None else down_fov)
        xyz = _as_f32(pointcloud[:, :, :3])
        B, N, _ = xyz.shape
        out = _empty_like_kind(xyz, (B, N, 2))
        self.ctx.call("pls_project_pixels", _lib.ptr(xyz), B * N, H, W, float(up), float(down), _lib.ptr(out))
        return out

    def build_projection_map(self, pointcloud, default_value: float = 0.0, height=None, width=None,
                             transform=None, **kwargs):
        """[B,N,C>=3] -> [B,C_dest,H,W]; the closest point per pixel survives (projection.py:331-418)."""
        check_tensor(pointcloud, [-1, -1, -1])
        H, W = height or self.height, width or self.width
        pc = _as_f32(pointcloud)
        B, N, _ = pc.shape
        transform = self.conversion if transform is None else transform      # Projector.swap (projection.py:322-329,368)
        channels = _as_f32(transform(pc)) if transform is not None else None  # None: xyz_conversion, the xyz are scattered
        xyz = _as_f32(pc[:, :, :3])
        Cd = 3 if channels is None else channels.shape[2]
        out = _empty_like_kind(xyz, (B, Cd, H, W))
        self.ctx.call("pls_build_projection_map_filled", _lib.ptr(xyz), _lib.ptr(channels), B, N, Cd, H, W,
                      float(self.up_fov), float(self.down_fov), float(default_value), _lib.ptr(out))
        return out


# ------------------------------------------------------------------------------------------
# slam/common/geometry.py
# ------------------------------------------------------------------------------------------
def compute_normal_map(vertex_map, kernel_size: int = 5, ctx=None):
    """[B,3,H,W] -> unit normals [B,3,H,W] (geometry.py:240-295)."""
    ctx = ctx or default_context()
    check_tensor(vertex_map, [-1, 3, -1, -1])
    vm = _as_f32(vertex_map)
    B, _, H, W = vm.shape
    out = _empty_like_kind(vm, (B, 3, H, W))
    ctx.call("pls_normal_map", _lib.ptr(vm), B, H, W, int(kernel_size), _lib.ptr(out))
    return out


def compute_neighbors(vm_target, vm_reference, reference_fields=None, ctx=None, **kwargs):
    """Per pixel nearest of D reference vertex maps (geometry.py:397-439)."""
    ctx = ctx or default_context()
    check_tensor(vm_target, [1, 3, -1, -1])
    _, _, H, W = vm_target.shape
    check_tensor(vm_reference, [-1, 3, H, W])
    tgt, ref = _as_f32(vm_target), _as_f32(vm_reference)
    K = ref.shape[0]
    fields = None if reference_fields is None else _as_f32(reference_fields)
    Cf = 0 if fields is None else fields.shape[1]
    nb = _empty_like_kind(tgt, (1, 3, H, W))
    nf = None if fields is None else _empty_like_kind(tgt, (1, Cf, H, W))
    ctx.call("pls_compute_neighbors", _lib.ptr(tgt), _lib.ptr(ref), _lib.ptr(fields), K, Cf, H, W, _lib.ptr(nb), _lib.ptr(nf))
    return nb, nf


# ------------------------------------------------------------------------------------------
# slam/common/pose.py (Euler xyz only: the hot path's representation)
# ------------------------------------------------------------------------------------------
class Pose:
    def __init__(self, pose_type: str = "euler", ctx=None):
        assert_debug(pose_type == "euler", "only the euler pose representation is on the hot path")
        self.pose_type = pose_type
        self._ctx = ctx

    @property
    def ctx(self):
        return self._ctx or default_context()

    @staticmethod
    def num_params() -> int:
        return 6

    def build_pose_matrix(self, params):
        """[B,6] -> [B,4,4] (pose.py:120-144)."""
        check_tensor(params, [-1, 6])
        p = _as_f32(params)
        out = _empty_like_kind(p, (p.shape[0], 4, 4))
        self.ctx.call("pls_build_pose_matrix", _lib.ptr(p), p.shape[0], _lib.ptr(out))
        return out

    def from_pose_matrix(self, mats):
        """[B,4,4] -> [B,6] (pose.py:188-207)."""
        check_tensor(mats, [-1, 4, 4])
        m = _as_f32(mats)
        out = _empty_like_kind(m, (m.shape[0], 6))
        self.ctx.call("pls_from_pose_matrix", _lib.ptr(m), m.shape[0], _lib.ptr(out))
        return out


def euler_pose_matrix_f64(params: np.ndarray) -> np.ndarray:
    """float64 host build_pose_matrix for the absolute-pose bookkeeping (icp_odometry.py:200-202):
    R = Rz(ez) Ry(ey) Rx(ex) written out (this runs once per frame on the host; scalar math beats numpy's
    small-matrix overheads by 10x)."""
    import math
    tx, ty, tz, ex, ey, ez = [float(v) for v in params]
    cx, sx, cy, sy, cz, sz = math.cos(ex), math.sin(ex), math.cos(ey), math.sin(ey), math.cos(ez), math.sin(ez)
    return np.array([[cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx, tx],
                     [sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx, ty],
                     [-sy, cy * sx, cy * cx, tz],
                     [0.0, 0.0, 0.0, 1.0]], dtype=np.float64)
