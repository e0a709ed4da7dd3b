"""Dataset -> vertex-map ingestion on the GPU.

In the reference a DataLoader worker turns every KITTI scan into the odometry's inputs on the CPU
(`KITTIOdometrySequence.__getitem__`, slam/dataset/kitti_dataset.py:233-249): read the `.bin`, rectify the HDL-64's
intrinsic error (`correct_scan`, :200-231), keep the cloud as `numpy_pc` and project it into the `vertex_map` with
`SphericalProjector.build_projection_map`.  The same three steps here, the arithmetic in the CUDA library
(pls_kitti_correct_scan / pls_ingest_scan), names, keys and dtypes as in the reference.
"""
import numpy as np
import torch

from . import _lib
from .common import assert_debug, check_tensor, default_context


def kitti_read_scan(file_path: str) -> np.ndarray:
    """A KITTI velodyne `.bin` as an `(N, 4)` float32 array: x, y, z, reflectance (kitti_dataset.py:20-38)."""
    try:
        return np.fromfile(file_path, dtype=np.float32).reshape((-1, 4))
    except (Exception, ValueError) as e:
        print(f"Error reading scan : {file_path}")
        raise e


def correct_scan(scan: np.ndarray, ctx=None) -> np.ndarray:
    """KITTIOdometrySequence.correct_scan (kitti_dataset.py:200-231): `(N, 3+)` float32 -> `(N, 3)` float64."""
    ctx = ctx or default_context()
    assert_debug(isinstance(scan, np.ndarray) and scan.ndim == 2 and scan.shape[1] in (3, 4), "scan must be [N,3] or [N,4]")
    s = np.ascontiguousarray(scan, dtype=np.float32)
    out = np.empty((s.shape[0], 3), dtype=np.float64)
    ctx.call("pls_kitti_correct_scan", _lib.ptr(s), s.shape[0], s.shape[1], _lib.ptr(out))
    return out


def kitti_frame(scan, projector, corrected_lidar_channel: str = "vertex_map", with_numpy_pc: bool = True,
                correct: bool = True, ctx=None) -> dict:
    """The `data_dict` entries `KITTIOdometrySequence.__getitem__` produces for one scan (kitti_dataset.py:239-249):
    `numpy_pc` (the rectified cloud, float64 `[N,3]`) and the vertex map (`torch` float64 `[3,H,W]`) under
    `corrected_lidar_channel`.  `scan` is the `[N,4]` array or the path of a `.bin`."""
    ctx = ctx or default_context()
    if isinstance(scan, str):
        scan = kitti_read_scan(scan)
    check_tensor(scan, [-1, scan.shape[1]])
    assert_debug(scan.shape[1] in (3, 4), "scan must be [N,3] or [N,4]")
    s = np.ascontiguousarray(scan, dtype=np.float32)
    H, W = projector.height, projector.width
    xyz = np.empty((s.shape[0], 3), dtype=np.float64) if with_numpy_pc else None
    vmap = np.empty((3, H, W), dtype=np.float64)
    ctx.call("pls_ingest_scan", _lib.ptr(s), s.shape[0], s.shape[1], int(bool(correct)), H, W, float(projector.up_fov),
             float(projector.down_fov), _lib.ptr(xyz), _lib.ptr(vmap))
    data_dict = {}
    if with_numpy_pc:
        data_dict["numpy_pc"] = xyz
    data_dict[corrected_lidar_channel] = torch.from_numpy(vmap)
    return data_dict
