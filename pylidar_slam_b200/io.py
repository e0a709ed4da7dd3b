"""Pose I/O and pose chains: the 12-column pose CSV of slam/common/io.py:17-76 -- the
`<sequence>.poses.txt` files the runner writes and the evaluation reads -- and compute_relative_poses /
compute_absolute_poses of slam/eval/eval_odometry.py:80-96.  The file format is byte-compatible with the reference's
pandas writer (header `0,...,11`, `,` separated, every value in its dtype's shortest round-trip representation); the pose
chains run in the CUDA library (pls_relative_poses / pls_absolute_poses)."""
from pathlib import Path

import numpy as np

from . import _lib
from .common import assert_debug, check_tensor, default_context


def delimiter():
    return ","


def poses_to_rows(poses_array: np.ndarray) -> np.ndarray:
    """poses_to_df's payload (io.py:63-76): `[N,4,4]` -> `[N,12]`, the first three rows of every pose."""
    shape = poses_array.shape
    assert_debug(len(shape) == 3)
    assert_debug(shape[1] == 4 and shape[2] == 4)
    return poses_array[:, :3, :].reshape([shape[0], 12])


def rows_to_poses(array: np.ndarray) -> np.ndarray:
    """df_to_poses (io.py:44-60): `[N,12]` (cast to float32, as the reference does) -> float64 `[N,4,4]`."""
    array = np.asarray(array, dtype=np.float32)
    assert_debug(array.ndim == 2 and array.shape[1] == 12)
    n = array.shape[0]
    last_row = np.concatenate((np.zeros((n, 3)), np.ones((n, 1))), axis=1)[:, None, :]
    return np.concatenate((array.reshape([n, 3, 4]), last_row), axis=1)


def write_poses_to_disk(file_path: str, poses: np.ndarray):
    """write_poses_to_disk (io.py:17-29)."""
    check_tensor(poses, [-1, 4, 4])
    path = Path(file_path)
    assert_debug(path.parent.exists())
    rows = poses_to_rows(np.asarray(poses))
    lines = [delimiter().join(str(i) for i in range(12))]
    lines += [delimiter().join(str(v) for v in row) for row in rows]
    path.write_text("\n".join(lines) + "\n")


def read_poses_from_disk(file_path: str, _delimiter: str = delimiter()) -> np.ndarray:
    """read_poses_from_disk (io.py:32-41): float64 `[N,4,4]` (values rounded through float32, like the reference)."""
    path = Path(file_path)
    assert_debug(path.exists() and path.is_file())
    rows = np.loadtxt(str(path), delimiter=_delimiter, skiprows=1, dtype=np.float64, ndmin=2)
    return rows_to_poses(rows)


def _pose_chain(name, poses, ctx):
    ctx = ctx or default_context()
    check_tensor(poses, [-1, 4, 4])
    # the reference's identity (shift_poses) and copies are float64, so float32 inputs are promoted before the arithmetic
    p = np.ascontiguousarray(poses, dtype=np.float64)
    out = np.empty_like(p)
    ctx.call(name, _lib.ptr(p), p.shape[0], 1, _lib.ptr(out))
    return out


def compute_relative_poses(poses: np.ndarray, ctx=None) -> np.ndarray:
    """compute_relative_poses (eval_odometry.py:80-83): `inv(poses[i-1]) @ poses[i]`, the first pose kept."""
    return _pose_chain("pls_relative_poses", poses, ctx)


def compute_absolute_poses(relative_poses: np.ndarray, ctx=None) -> np.ndarray:
    """compute_absolute_poses (eval_odometry.py:86-96): the running product of the relative poses.  A float32 input keeps
    its dtype in the reference (`relative_poses.copy()`); here the product is formed in float64 and cast back."""
    out = _pose_chain("pls_absolute_poses", relative_poses, ctx)
    return out.astype(relative_poses.dtype, copy=False) if np.asarray(relative_poses).dtype == np.float32 else out
