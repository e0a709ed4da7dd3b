"""Mirrors of the reference's odometry plug-in interfaces for the ICP hot path:

  OdometryAlgorithm / OdometryConfig              slam/odometry/odometry.py:13-81
  LocalMap, KdTreeLocalMap, ProjectiveLocalMap    slam/odometry/local_map.py:21-445
  RigidAlignment, GaussNewtonPointToPlaneAlignment slam/odometry/alignment.py:23-127
  ICPFrameToModel, ICPFrameToModelConfig           slam/odometry/icp_odometry.py:27-380

Same class / method / config-field names, same data_dict keys, same error behaviour; the
arithmetic runs in libplslam_b200.so (hand-written sm_100a CUDA) through the C ABI.  The
registries (`ODOMETRY`, `LOCAL_MAP`, `RIGID_ALIGNMENT`) keep the reference's discriminator
fields (`algorithm`, `type`, `mode`) and add `*_b200` members; INTEGRATION.md shows the two-line
patch that registers them inside the reference tree.
"""
import ctypes as C
import dataclasses
import time
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import _lib
from .common import Pose, SphericalProjector, assert_debug, check_tensor, euler_pose_matrix_f64

MISSING = "???"


def _cfg_to_dict(cfg) -> dict:
    if cfg is None:
        return {}
    if dataclasses.is_dataclass(cfg):
        return {f.name: getattr(cfg, f.name) for f in dataclasses.fields(cfg)}
    if isinstance(cfg, dict) or hasattr(cfg, "keys"):
        return {k: cfg[k] for k in cfg.keys()}
    raise AssertionError(f"cannot interpret {cfg!r} as a config")


class ObjectLoaderEnum:
    """slam/common/utils.py:266-302: discriminator-driven loading from a config."""

    @classmethod
    def load(cls, config, **kwargs):
        d = _cfg_to_dict(config)
        assert_debug(cls.type_name() in d, f"The config does not contains the key : '{cls.type_name()}'")
        _type = d[cls.type_name()]
        assert_debug(_type in cls.__members__,
                     f"Unknown type `{_type}`. Existing members are : {cls.__members__.keys()}")
        _class, _config = cls.__members__[_type].value
        if not isinstance(config, _config):
            names = {f.name for f in dataclasses.fields(_config)}
            config = _config(**{k: v for k, v in d.items() if k in names and v != MISSING})
        return _class(config, **kwargs)


# ----------------------------------------------------------------------------------------------------------------------
# Local maps
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class LocalMapConfig:
    pose: str = "euler"
    type: str = MISSING


@dataclass
class KdTreeLocalMapConfig(LocalMapConfig):
    local_map_size: int = 20
    num_neighbors_normals: int = 10
    type: str = "kdtree_local_map"


@dataclass
class ProjectiveLocalMapConfig(LocalMapConfig):
    local_map_size: int = 20
    type: str = "projective_local_map"
    normals_kernel_size: int = 5


class LocalMap(ABC):
    """slam/odometry/local_map.py:31-79"""

    @dataclass
    class NeighborhoodResult:
        neighbor_points: Optional[Any] = None
        neighbor_normals: Optional[Any] = None
        new_target_points: Optional[Any] = None

    def __init__(self, config: LocalMapConfig, **kwargs):
        self.config = config
        self.pose = Pose(config.pose)

    @abstractmethod
    def init(self):
        raise NotImplementedError("")

    @abstractmethod
    def update(self, new_relative_pose, new_pc_data=None, new_vertex_map=None, **kwargs) -> None:
        raise NotImplementedError("")

    @abstractmethod
    def nearest_neighbor_search(self, points, with_normals: bool = True, with_new_target_points: bool = True, **kwargs):
        raise NotImplementedError("")


def _pose16(relative_pose) -> np.ndarray:
    if isinstance(relative_pose, torch.Tensor):
        relative_pose = relative_pose.detach().cpu().numpy()
    rel = np.ascontiguousarray(np.asarray(relative_pose, dtype=np.float32).reshape(4, 4))
    return rel


def _f32c(x):
    if isinstance(x, np.ndarray):
        return np.ascontiguousarray(x, dtype=np.float32)
    return x.to(torch.float32).contiguous()


def _new(x, shape, dtype=torch.float32):
    if isinstance(x, np.ndarray):
        return np.empty(shape, dtype={torch.float32: np.float32, torch.int64: np.int64}[dtype])
    return torch.empty(shape, dtype=dtype, device=x.device)


class KdTreeLocalMap(LocalMap):
    """KdTreeLocalMap (local_map.py:254-427) on the GPU: exact 1-NN over the hashed cell pyramid + lazily cached 10-NN normals."""

    def __init__(self, config: KdTreeLocalMapConfig, projector=None, ctx: Optional[_lib.Context] = None, **kwargs):
        super().__init__(config)
        self.ctx = ctx or _lib.Context(local_map_type=_lib.MAP_KDTREE, local_map_size=config.local_map_size,
                                       num_neighbors_normals=config.num_neighbors_normals)

    def init(self):
        self.ctx.call("pls_map_init")

    def update(self, relative_pose, new_pc_data=None, new_vertex_map=None, **kwargs):
        rel = _pose16(relative_pose)
        if new_pc_data is not None:
            pts = _f32c(new_pc_data.reshape(-1, 3))
            self.ctx.call("pls_kdmap_update_points", _lib.ptr(rel), _lib.ptr(pts), pts.shape[0])
        elif new_vertex_map is not None:
            check_tensor(new_vertex_map, [1, 3, -1, -1])
            vm = _f32c(new_vertex_map)
            self.ctx.call("pls_kdmap_update_vertex_map", _lib.ptr(rel), _lib.ptr(vm), vm.shape[2], vm.shape[3])
        else:
            self.ctx.call("pls_kdmap_update_points", _lib.ptr(rel), None, 0)

    def num_points(self) -> int:
        n = C.c_int64(0)
        self.ctx.call("pls_kdmap_size", C.byref(n))
        return n.value

    def points(self) -> np.ndarray:
        out = np.empty((self.num_points(), 3), dtype=np.float32)
        self.ctx.call("pls_kdmap_points", _lib.ptr(out))
        return out

    def nearest_neighbor_search(self, target_points, with_normals: bool = True, with_new_target_points: bool = True,
                                **kwargs):
        check_tensor(target_points, [-1, 3])
        q = _f32c(target_points)
        n = q.shape[0]
        nb = _new(q, (n, 3))
        nrm = _new(q, (n, 3)) if with_normals else None
        self.ctx.call("pls_kdmap_nn_search", _lib.ptr(q), n, _lib.ptr(nb), _lib.ptr(nrm), None)
        result = self.NeighborhoodResult()
        is_torch = isinstance(q, torch.Tensor)
        result.neighbor_points = nb.unsqueeze(0) if is_torch else nb
        if with_normals:
            result.neighbor_normals = nrm.unsqueeze(0) if is_torch else nrm
        if with_new_target_points:
            result.new_target_points = q.reshape(1, n, 3) if is_torch else q
        return result


class ProjectiveLocalMap(LocalMap):
    """ProjectiveLocalMap (local_map.py:91-240) on the GPU."""

    def __init__(self, config: ProjectiveLocalMapConfig, projector: SphericalProjector = None,
                 ctx: Optional[_lib.Context] = None, **kwargs):
        super().__init__(config)
        assert_debug(projector is not None)
        self.projector = projector
        self.ctx = ctx or _lib.Context(local_map_type=_lib.MAP_PROJECTIVE, local_map_size=config.local_map_size,
                                       normals_kernel_size=config.normals_kernel_size, height=projector.height,
                                       width=projector.width, up_fov_deg=projector.up_fov,
                                       down_fov_deg=projector.down_fov)

    def init(self):
        self.ctx.call("pls_map_init")

    def update(self, relative_pose, new_vertex_map=None, new_normal_map=None, mask=None, **kwargs):
        rel = _pose16(relative_pose)
        vm = None
        if new_vertex_map is not None:
            check_tensor(new_vertex_map, [1, 3, self.projector.height, self.projector.width])
            vm = _f32c(new_vertex_map)
        self.ctx.call("pls_projmap_update", _lib.ptr(rel), _lib.ptr(vm))

    def model(self):
        """(_model_vmap, _model_nmap), each [K,3,H,W] numpy."""
        k = C.c_int(0)
        self.ctx.call("pls_projmap_num_frames", C.byref(k))
        H, W = self.projector.height, self.projector.width
        v = np.empty((k.value, 3, H, W), dtype=np.float32)
        n = np.empty((k.value, 3, H, W), dtype=np.float32)
        self.ctx.call("pls_projmap_model", _lib.ptr(v), _lib.ptr(n))
        return v, n

    def nearest_neighbor_search(self, target_points, with_normals: bool = True, with_new_target_points: bool = True,
                                **kwargs):
        check_tensor(target_points, [-1, 3])
        q = _f32c(target_points)
        hw = self.projector.height * self.projector.width
        nb, nrm, tgt = _new(q, (hw, 3)), _new(q, (hw, 3)), _new(q, (hw, 3))
        count = C.c_int64(0)
        self.ctx.call("pls_projmap_nn_search", _lib.ptr(q), q.shape[0], _lib.ptr(nb), _lib.ptr(nrm), _lib.ptr(tgt),
                      C.byref(count))
        nc = count.value
        res = self.NeighborhoodResult()
        wrap = (lambda a: a[:nc].unsqueeze(0)) if isinstance(q, torch.Tensor) else (lambda a: a[:nc][None])
        res.neighbor_points = wrap(nb)
        if with_normals:
            res.neighbor_normals = wrap(nrm)
        if with_new_target_points:
            res.new_target_points = wrap(tgt)
        return res


class LOCAL_MAP(ObjectLoaderEnum, Enum):
    """slam/odometry/local_map.py:437-445 (+ explicit *_b200 aliases)"""
    projective_local_map = (ProjectiveLocalMap, ProjectiveLocalMapConfig)
    kdtree_local_map = (KdTreeLocalMap, KdTreeLocalMapConfig)

    @classmethod
    def type_name(cls):
        return "type"


# ----------------------------------------------------------------------------------------------------------------------
# Rigid alignment
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class RigidAlignmentConfig:
    mode: str = MISSING
    pose: str = "euler"
    scheme: str = "huber"


@dataclass
class GaussNewtonPointToPlaneConfig(RigidAlignmentConfig):
    mode: str = "point_to_plane_gauss_newton"
    num_gn_iters: int = 1
    gauss_newton_config: Dict[str, Any] = field(default_factory=lambda: dict(max_iters=1))


def _reject_mask(mask, n):
    """`mask` of RigidAlignment.align ([1,n,1], alignment.py:96-121,158-182).  The reference cannot apply one: its cost
    functions multiply the [b,n,6] Jacobian IN PLACE by `mask.unsqueeze(1)` ([b,1,n,1]; optimization.py:391-392,500-501),
    which fails to broadcast -- every call with a mask ends in a RuntimeError.  Same error type here, after the
    reference's own shape check."""
    check_tensor(mask, [1, n, 1])
    raise RuntimeError(f"output with shape [1, {n}, 6] doesn't match the broadcast shape [1, 1, {n}, 6] "
                       f"(the reference's alignments cannot apply a mask: optimization.py:391-392, 500-501)")


def _gn_settings(gn_cfg) -> dict:
    """GaussNewton(**gauss_newton_config) defaults (optimization.py:287-294, :61-208)."""
    d = dict(_cfg_to_dict(gn_cfg))
    scheme = d.get("scheme", "default")
    assert_debug(scheme in _lib.SCHEMES, f"unknown weighting scheme {scheme}")
    return dict(scheme=scheme, sigma=float(d.get("sigma", 0.5)), max_iters=max(int(d.get("max_iters", 10)), 1),
                norm_stop=float(d.get("norm_stop_criterion", 1e-3)))


class RigidAlignment(ABC):
    def __init__(self, alignment_config: RigidAlignmentConfig, **kwargs):
        self.config = alignment_config
        self.pose = Pose(self.config.pose)

    def align(self, ref_points, tgt_points, *args, **kwargs):
        raise NotImplementedError("")


class GaussNewtonPointToPlaneAlignment(RigidAlignment):
    """GaussNewtonPointToPlaneAlignment.align (alignment.py:91-127) -> pls_align_p2plane."""

    def __init__(self, config: GaussNewtonPointToPlaneConfig, ctx: Optional[_lib.Context] = None, **kwargs):
        super().__init__(config, **kwargs)
        self.gn = _gn_settings(config.gauss_newton_config)
        self._ctx = ctx

    @property
    def ctx(self):
        from .common import default_context
        return self._ctx or default_context()

    def align(self, ref_points, tgt_points, ref_normals=None, initial_estimate=None, mask=None, **kwargs):
        assert_debug(ref_normals is not None, "The argument 'ref_normals' is required for a point to plane alignemnt")
        check_tensor(tgt_points, [1, -1, 3])
        n = tgt_points.shape[1]
        if mask is not None:
            _reject_mask(mask, n)
        check_tensor(ref_points, [1, n, 3])
        check_tensor(ref_normals, [1, n, 3])
        is_np = isinstance(tgt_points, np.ndarray)
        is64 = (tgt_points.dtype == (np.float64 if is_np else torch.float64))
        dt_np, dt_t = (np.float64, torch.float64) if is64 else (np.float32, torch.float32)
        conv = (lambda a: np.ascontiguousarray(a, dtype=dt_np)) if is_np else (lambda a: a.to(dt_t).contiguous())
        ref, tgt, nrm = conv(ref_points), conv(tgt_points), conv(ref_normals)
        x0 = None
        if initial_estimate is not None:
            x0 = conv(initial_estimate)
            if x0.ndim == 3:
                assert_debug(not is64, "pose-matrix initial estimates are float32")
                x0 = self.pose.from_pose_matrix(x0)
            x0 = conv(x0.reshape(6))
        mk = (lambda s: np.empty(s, dtype=dt_np)) if is_np else (lambda s: torch.empty(s, dtype=dt_t, device=tgt.device))
        dT, x, loss = mk((1, 4, 4)), mk((1, 6)), mk((1, n))
        self.ctx.call("pls_align_p2plane", _lib.ptr(ref), _lib.ptr(tgt), _lib.ptr(nrm), n, int(is64),
                      _lib.SCHEMES[self.gn["scheme"]], self.gn["sigma"], self.gn["max_iters"], self.gn["norm_stop"],
                      _lib.ptr(x0), _lib.ptr(dT), _lib.ptr(x), _lib.ptr(loss))
        return dT, x, loss


@dataclass
class GNPointToPointConfig(RigidAlignmentConfig):
    """slam/odometry/alignment.py:131-141"""
    mode: str = "point_to_point_gn"
    num_gn_iters: int = 1
    initialize_with_svd: bool = False
    gauss_newton_config: Dict[str, Any] = field(default_factory=lambda: dict(max_iters=1))


class GaussNewtonPointToPointAlignment(RigidAlignment):
    """GaussNewtonPointToPointAlignment.align (alignment.py:144-189) -> pls_align_p2point.

    Faithful to the reference, including its Jacobian (optimization.py:485-501 is r * dr/dx, see gn_device.cuh).
    `initialize_with_svd=True` is rejected: the reference's torch weighted_procrustes builds its output with
    `.repeat(b, 4, 4)` (registration.py:58-59), a [b,16,16] tensor that its own from_pose_matrix shape check then
    refuses -- there is no reference behaviour to reproduce; `weighted_procrustes` (numpy path) is offered
    stand-alone in pylidar_slam_b200.common."""

    def __init__(self, config: GNPointToPointConfig, ctx: Optional[_lib.Context] = None, **kwargs):
        super().__init__(config, **kwargs)
        self.gn = _gn_settings(config.gauss_newton_config)
        self._ctx = ctx

    @property
    def ctx(self):
        from .common import default_context
        return self._ctx or default_context()

    def align(self, ref_points, tgt_points, initial_estimate=None, mask=None, **kwargs):
        assert_debug(not self.config.initialize_with_svd,
                     "initialize_with_svd fails inside the reference itself (registration.py:58-59 builds a [b,16,16] "
                     "tensor); use pylidar_slam_b200.common.weighted_procrustes and pass it as initial_estimate")
        check_tensor(tgt_points, [1, -1, 3])
        n = tgt_points.shape[1]
        check_tensor(ref_points, [1, n, 3])
        if mask is not None:
            _reject_mask(mask, n)
        is_np = isinstance(tgt_points, np.ndarray)
        is64 = (tgt_points.dtype == (np.float64 if is_np else torch.float64))
        dt_np, dt_t = (np.float64, torch.float64) if is64 else (np.float32, torch.float32)
        conv = (lambda a: np.ascontiguousarray(a, dtype=dt_np)) if is_np else (lambda a: a.to(dt_t).contiguous())
        ref, tgt = conv(ref_points), conv(tgt_points)
        x0 = None
        if initial_estimate is not None:
            x0 = conv(initial_estimate)
            if x0.ndim == 3:
                check_tensor(x0, [1, 4, 4])
                assert_debug(not is64, "pose-matrix initial estimates are float32")
                x0 = self.pose.from_pose_matrix(x0)
            x0 = conv(x0.reshape(6))
        mk = (lambda s: np.empty(s, dtype=dt_np)) if is_np else (lambda s: torch.empty(s, dtype=dt_t, device=tgt.device))
        dT, x, loss = mk((1, 4, 4)), mk((1, 6)), mk((1, n))
        self.ctx.call("pls_align_p2point", _lib.ptr(ref), _lib.ptr(tgt), n, int(is64), _lib.SCHEMES[self.gn["scheme"]],
                      self.gn["sigma"], self.gn["max_iters"], self.gn["norm_stop"], _lib.ptr(x0), _lib.ptr(dT), _lib.ptr(x),
                      _lib.ptr(loss))
        return dT, x, loss


class RIGID_ALIGNMENT(ObjectLoaderEnum, Enum):
    """slam/odometry/alignment.py:200-208"""
    point_to_plane_gauss_newton = (GaussNewtonPointToPlaneAlignment, GaussNewtonPointToPlaneConfig)
    point_to_point_gauss_newton = (GaussNewtonPointToPointAlignment, GNPointToPointConfig)

    @classmethod
    def type_name(cls):
        return "mode"


# ----------------------------------------------------------------------------------------------------------------------
# Odometry
# ----------------------------------------------------------------------------------------------------------------------
@dataclass
class OdometryConfig:
    algorithm: str = MISSING


class OdometryAlgorithm(ABC):
    """slam/odometry/odometry.py:21-81"""

    def __init__(self, config: OdometryConfig, **kwargs):
        self.config = config
        self.elapsed: list = []

    @abstractmethod
    def init(self):
        self.elapsed = []

    def process_next_frame(self, data_dict: dict):
        beginning = time.time()
        self.do_process_next_frame(data_dict)
        self.elapsed.append(time.time() - beginning)

    @abstractmethod
    def do_process_next_frame(self, data_dict: dict):
        raise NotImplementedError("")

    def get_relative_poses(self) -> np.ndarray:
        raise NotImplementedError("")

    def get_elapsed(self) -> float:
        return sum(self.elapsed)

    @staticmethod
    def pointcloud_key() -> str:
        return "odometry_pc"

    @staticmethod
    def relative_pose_key() -> str:
        return "odometry_pose"


@dataclass
class ICPFrameToModelConfig(OdometryConfig):
    """icp_odometry.py:27-64 (visualisation fields are accepted and ignored)."""
    algorithm: str = "icp_F2M"
    device: str = "cuda:0"
    pose: str = "euler"
    max_num_alignments: int = 100
    initialization: Any = MISSING
    local_map: Any = MISSING
    alignment: Any = MISSING
    threshold_delta_pose: float = 1.e-4
    threshold_trans: float = 0.1
    threshold_rot: float = 0.3
    sigma: float = 0.1
    data_key: str = "vertex_map"
    viz_debug: bool = False
    viz_with_edl: bool = True
    viz_color_by_elevation: bool = True
    viz_grayscale: str = "viridis"
    viz_num_pcs: int = 500
    viz_z_min: float = 0.0
    viz_z_max: float = 30

    def completed(self):
        """RuntimeDefaultDict.completed (utils.py:199-262): runtime defaults kdtree + point_to_plane_GN."""
        if self.local_map is None or self.local_map == MISSING:
            self.local_map = KdTreeLocalMapConfig()
        if self.alignment is None or self.alignment == MISSING:
            self.alignment = GaussNewtonPointToPlaneConfig()
        return self


class ICPFrameToModel(OdometryAlgorithm):
    """ICPFrameToModel (icp_odometry.py:72-380): the whole frame -- input normalisation, ICP loop,
    key-frame policy, local-map update -- runs inside one `pls_process_frame` call."""

    def __init__(self, config: ICPFrameToModelConfig, projector: SphericalProjector = None, pose: Pose = None,
                 device=None, stream: Optional[int] = None, **kwargs):
        if not isinstance(config, ICPFrameToModelConfig):
            names = {f.name for f in dataclasses.fields(ICPFrameToModelConfig)}
            config = ICPFrameToModelConfig(**{k: v for k, v in _cfg_to_dict(config).items() if k in names})
        config = config.completed()
        super().__init__(config)
        assert_debug(projector is not None)
        self.projector = projector
        self.pose = pose or Pose("euler")
        dev = torch.device(device if device is not None else config.device)
        assert_debug(dev.type == "cuda", "ICPFrameToModel (B200) runs on a CUDA device; there is no CPU path")
        self.device = dev

        lm = _cfg_to_dict(self.config.local_map)
        al = _cfg_to_dict(self.config.alignment)
        assert_debug(lm.get("type") in LOCAL_MAP.__members__, f"Unknown type `{lm.get('type')}`")
        assert_debug(al.get("mode", "point_to_plane_gauss_newton") in RIGID_ALIGNMENT.__members__,
                     f"Unknown mode `{al.get('mode')}`")
        # the reference's ICP loop calls align(neigh, tgt, normals) (icp_odometry.py:285-288): with the point-to-point
        # alignment the normals land in `initial_estimate` and its shape check raises -- same error type here
        assert_debug(al.get("mode", "point_to_plane_gauss_newton") == "point_to_plane_gauss_newton",
                     "ICPFrameToModel needs the point-to-plane alignment (the reference's loop cannot drive point_to_point)")
        gn = _gn_settings(al.get("gauss_newton_config", dict(max_iters=1)))
        is_kd = lm["type"] == "kdtree_local_map"
        self.ctx = _lib.Context(
            height=projector.height, width=projector.width, up_fov_deg=float(projector.up_fov),
            down_fov_deg=float(projector.down_fov),
            local_map_type=_lib.MAP_KDTREE if is_kd else _lib.MAP_PROJECTIVE,
            local_map_size=int(lm.get("local_map_size", 20)),
            num_neighbors_normals=int(lm.get("num_neighbors_normals", 10)),
            normals_kernel_size=int(lm.get("normals_kernel_size", 5)),
            scheme=_lib.SCHEMES[gn["scheme"]], sigma=gn["sigma"], gn_max_iters=1,
            gn_norm_stop=gn["norm_stop"], max_num_alignments=int(self.config.max_num_alignments),
            threshold_delta_pose=float(self.config.threshold_delta_pose),
            threshold_trans=float(self.config.threshold_trans), threshold_rot=float(self.config.threshold_rot),
            device=dev.index or 0, stream=stream)
        self.gn_max_iters = self.config.max_num_alignments
        # The fused C-ABI path runs ONE Gauss-Newton step per ICP iteration (the shipped configuration,
        # alignment.py:77).  gauss_newton_config.max_iters > 1 (alignment.py:69-77,110-127: several re-linearised steps on
        # the same correspondences) takes the reference-shaped loop below, built from the fine-grained GPU plug-ins.
        self._fine_grained = gn["max_iters"] > 1
        if self._fine_grained:
            self.ctx.cfg.gn_max_iters = 1  # the context's own fused entry points stay usable (register_new_frame)
            self.local_map = LOCAL_MAP.load(self.config.local_map, projector=projector, ctx=self.ctx)
            self.rigid_alignment = RIGID_ALIGNMENT.load(self.config.alignment, ctx=self.ctx)
            self._pose_ops = Pose("euler", ctx=self.ctx)
            self._delta_since_map_update = torch.eye(4, dtype=torch.float32)
        self.relative_poses: list = []
        self.absolute_poses: list = []
        self._iter = 0
        self.last_info = np.zeros(12, dtype=np.float64)
        self._pose_out = np.zeros((4, 4), dtype=np.float32)
        self._params_out = np.zeros(6, dtype=np.float32)
        # per-frame call arguments that never change: addresses of the persistent output arrays, the has-pose cell
        self._has_pose = C.c_int(0)
        self._out_args = (_lib.ptr(self._pose_out), _lib.ptr(self._params_out), C.byref(self._has_pose),
                          _lib.ptr(self.last_info))

    def init(self):
        super().init()
        self.relative_poses = []
        self.absolute_poses = []
        self._iter = 0
        self.ctx.call("pls_odometry_init")
        if self._fine_grained:
            self.local_map.init()
            self._delta_since_map_update = torch.eye(4, dtype=torch.float32)
            self._sample_pointcloud = False

    def _interpret(self, data):
        """_read_input's three layouts (icp_odometry.py:319-358)."""
        H, W = self.projector.height, self.projector.width
        # float64 clouds keep their precision up to the projection, like the reference (icp_odometry.py:331-352)
        if isinstance(data, np.ndarray):
            check_tensor(data, [-1, 3])
            if data.dtype == np.float64:
                return _lib.INPUT_NDARRAY_F64 | _lib.PTR_HOST, np.ascontiguousarray(data), data.shape[0]
            return _lib.INPUT_NDARRAY | _lib.PTR_HOST, np.ascontiguousarray(data, dtype=np.float32), data.shape[0]
        if isinstance(data, torch.Tensor):
            if data.dim() in (3, 4):
                vm = data if data.dim() == 4 else data.unsqueeze(0)
                assert_debug(vm.shape[0] == 1, "Unexpected batched data format.")
                check_tensor(vm, [1, 3, H, W])
                return _lib.INPUT_VERTEX_MAP, vm.to(torch.float32).contiguous(), 0
            assert_debug(data.dim() == 2)
            check_tensor(data, [-1, 3])
            hint = _lib.PTR_DEVICE if data.is_cuda else _lib.PTR_HOST
            if data.dtype == torch.float64:
                return _lib.INPUT_TENSOR_F64 | hint, data.contiguous(), data.shape[0]
            if data.dtype != torch.float32 or not data.is_contiguous():
                data = data.to(torch.float32).contiguous()
            return _lib.INPUT_TENSOR | hint, data, data.shape[0]
        raise RuntimeError(f"Could not interpret the data: {data} as a pointcloud tensor")

    # -- the reference-shaped loop over the fine-grained plug-ins (icp_odometry.py:157-380), for configurations the fused
    #    path does not cover: the NN search, the normals, every Gauss-Newton step and the map update are the same CUDA
    #    kernels, the loop itself runs here (one host round trip per ICP iteration, like the reference)
    def _process_fine_grained(self, data_dict: dict):
        dev = self.device
        data = data_dict[self.config.data_key]
        if isinstance(data, np.ndarray):                                   # _read_input, icp_odometry.py:319-358
            check_tensor(data, [-1, 3])
            self._sample_pointcloud = True
            pc = torch.from_numpy(data).to(dev).unsqueeze(0)
            vmap = self.projector.build_projection_map(pc.to(torch.float32))
        elif isinstance(data, torch.Tensor):
            if data.dim() in (3, 4):
                vmap = data.to(dev) if data.dim() == 4 else data.to(dev).unsqueeze(0)
                assert_debug(vmap.shape[0] == 1, "Unexpected batched data format.")
                check_tensor(vmap, [1, 3, -1, -1])
                pc = vmap.permute(0, 2, 3, 1).reshape(1, -1, 3)
                pc = pc[:, (pc[0] != 0).any(dim=-1)][:, :1]   # the reference keeps the first non-null pixel only (:342-358)
            else:
                assert_debug(data.dim() == 2)
                pc = data.to(dev).unsqueeze(0)
                vmap = self.projector.build_projection_map(pc.to(torch.float32))
        else:
            raise RuntimeError(f"Could not interpret the data: {data} as a pointcloud tensor")
        vmap = vmap.to(torch.float32)
        vmap = torch.where(torch.isnan(vmap).any(dim=1, keepdim=True), torch.zeros_like(vmap), vmap)   # modify_nan_pmap
        pc = pc.to(torch.float32)
        pc = pc[:, ~torch.isnan(pc[0]).any(dim=-1)]                                                    # remove_nan
        if self._iter == 0:
            eye = torch.eye(4, dtype=torch.float32).unsqueeze(0)
            self.local_map.update(eye, new_vertex_map=vmap)
            self.relative_poses.append(eye.numpy())
            self.absolute_poses.append(np.eye(4, dtype=np.float64))
            self._iter += 1
            return
        init = data_dict.get("init_rpose", None)
        T = torch.eye(4, dtype=torch.float32, device=dev).unsqueeze(0) if init is None else \
            torch.from_numpy(np.asarray(init, dtype=np.float32).reshape(1, 4, 4)).to(dev)
        if self._sample_pointcloud:                                       # sample_points, icp_odometry.py:301-308
            points = pc[0]
        else:
            flat = vmap[0].permute(1, 2, 0).reshape(-1, 3)
            points = flat[flat.norm(dim=-1) > 0.0]
        params = torch.zeros(1, 6, dtype=torch.float32, device=dev)
        self.last_losses = []
        for _ in range(self.config.max_num_alignments):                   # register_new_frame, icp_odometry.py:274-297
            moved = points @ T[0, :3, :3].T + T[0, :3, 3]
            res = self.local_map.nearest_neighbor_search(moved)
            dT, delta, residuals = self.rigid_alignment.align(res.neighbor_points, res.new_target_points, res.neighbor_normals)
            self.last_losses.append(float(residuals.sum()))
            if float(torch.as_tensor(delta).norm()) < self.config.threshold_delta_pose:
                break
            params = self._pose_ops.from_pose_matrix(torch.as_tensor(dT).to(dev) @ T)
            T = self._pose_ops.build_pose_matrix(params)
        self.last_info[0] = len(self.last_losses)
        T_host = T.detach().cpu()
        new_delta = self._delta_since_map_update @ T_host[0]              # __update_map, icp_odometry.py:360-380
        dp = self._pose_ops.from_pose_matrix(new_delta.unsqueeze(0))
        dp = torch.as_tensor(dp).cpu()
        if float(dp[0, :3].norm()) > self.config.threshold_trans or float(dp[0, 3:].norm()) * 180 / np.pi > self.config.threshold_rot:
            self.local_map.update(T_host, new_vertex_map=vmap, new_pc_data=pc[0])
            self._delta_since_map_update = torch.eye(4, dtype=torch.float32)
        else:
            self.local_map.update(T_host)
            self._delta_since_map_update = new_delta
        T_np = T_host.numpy().reshape(4, 4).astype(np.float32)
        self.relative_poses.append(T_np.reshape(1, 4, 4))
        self.absolute_poses.append(self.absolute_poses[-1].dot(euler_pose_matrix_f64(torch.as_tensor(params).cpu().numpy().reshape(6))))
        if "distorted" in data_dict:
            tgt_np_pc = data_dict["distorted"]
        else:
            tgt_np_pc = pc[0].detach().cpu().numpy()
        data_dict[self.pointcloud_key()] = tgt_np_pc
        data_dict[self.relative_pose_key()] = T_np
        self._iter += 1

    def do_process_next_frame(self, data_dict: dict):
        assert_debug(self.config.data_key in data_dict,
                     f"Could not find the key `{self.config.data_key}` in the input dictionary.\n"
                     f"With keys : {data_dict.keys()}). Set the parameter `slam.odometry.data_key` to the desired key")
        if self._fine_grained:
            return self._process_fine_grained(data_dict)
        layout, data, n = self._interpret(data_dict[self.config.data_key])
        init = data_dict.get("init_rpose", None)
        init = None if init is None else np.ascontiguousarray(np.asarray(init, dtype=np.float32).reshape(4, 4))
        has_pose = self._has_pose
        address = _lib.ptr(data)
        if layout & _lib.PTR_HOST and n == _lib.Handoff.rows:
            # the array GridSample.filter handed out (possibly wrapped by ToTensor): its device-resident twin is used
            # instead of copying the samples host -> device again
            twin = _lib.Handoff.match(address, n, bool((layout & 0xff) >= _lib.INPUT_NDARRAY_F64), int(self.ctx.cfg.device))
            if twin:
                address, layout = twin, (layout & 0xff) | _lib.PTR_DEVICE
        self.ctx.call("pls_process_frame", address, layout, n, _lib.ptr(init), *self._out_args)
        layout &= 0xff
        if int(self.last_info[6]) == _lib.PLS_W_TINY_RESIDUAL:   # GaussNewton.compute's warning (optimization.py:323-327)
            import logging
            logging.warning("The residual norm is lower than threshold 1e-7. "
                            "This would lead to invalid jacobian. We prefer Stopping ICP")
        if not has_pose.value:
            eye = np.eye(4, dtype=np.float32).reshape(1, 4, 4)
            self.relative_poses.append(eye)
            self.absolute_poses.append(np.eye(4, dtype=np.float64))
            self._iter += 1
            return
        T = self._pose_out.copy()
        self.relative_poses.append(T.reshape(1, 4, 4))
        self.absolute_poses.append(self.absolute_poses[-1].dot(euler_pose_matrix_f64(self._params_out)))
        if "distorted" in data_dict:
            tgt_np_pc = data_dict["distorted"]
        elif layout == _lib.INPUT_VERTEX_MAP:
            tgt_np_pc = self.last_info[8:11].astype(np.float32).reshape(1, 3)  # icp_odometry.py:342-358 quirk
        else:
            tgt_np_pc = data if isinstance(data, np.ndarray) else \
                (data.detach().cpu().numpy() if data.is_cuda or data.requires_grad else data.numpy())
            tgt_np_pc = tgt_np_pc.astype(np.float32, copy=False)  # _tgt_pc is float32 (icp_odometry.py:352)
            if self.last_info[5] > 0:
                tgt_np_pc = tgt_np_pc[~np.isnan(tgt_np_pc).any(axis=1)]
        data_dict[self.pointcloud_key()] = tgt_np_pc
        data_dict[self.relative_pose_key()] = T
        self._iter += 1

    def get_relative_poses(self) -> np.ndarray:
        if len(self.relative_poses) == 0:
            return None
        return np.concatenate(self.relative_poses, axis=0)

    # -- fine-grained entry kept for parity tests: register_new_frame (icp_odometry.py:248-299)
    def register_new_frame(self, target_points, initial_estimate=None, **kwargs):
        pts = _f32c(target_points)
        T0 = None if initial_estimate is None else _pose16(initial_estimate)
        T, params = np.zeros((1, 4, 4), np.float32), np.zeros((1, 6), np.float32)
        losses = np.zeros(self.config.max_num_alignments, np.float32)
        iters = C.c_int(0)
        self.ctx.call("pls_register_frame", _lib.ptr(pts), pts.shape[0], _lib.ptr(T0), _lib.ptr(T), _lib.ptr(params),
                      _lib.ptr(losses), C.byref(iters))
        return params, T, list(losses[:iters.value])


class ODOMETRY(ObjectLoaderEnum, Enum):
    """slam/odometry/__init__.py:23-32"""
    icp_F2M = (ICPFrameToModel, ICPFrameToModelConfig)

    @classmethod
    def type_name(cls):
        return "algorithm"
