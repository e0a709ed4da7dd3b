"""TEST INFRASTRUCTURE ONLY -- CPU restatement of SURVEY.md section 8f rank 4: dataset -> vertex-map ingestion
(slam/dataset/kitti_dataset.py:200-249), pose I/O (slam/common/io.py:17-76) and the pose chains of
slam/eval/eval_odometry.py:80-96.  Pinned against the unmodified reference by tests/golden/io_rows.npz
(tests/golden/make_golden_io.py) in tests/test_io_oracle.py.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs
may import anything under oracle/."""
import numpy as np


def kitti_correct_scan(scan: np.ndarray) -> np.ndarray:
    """correct_scan (kitti_dataset.py:200-231): Rodrigues rotation of every point by 0.205 deg about normalise(p x e_z).
    float32 axes, float64 rotation and result -- the dtypes numpy's promotion gives the reference."""
    xyz = np.ascontiguousarray(scan[:, :3], dtype=np.float32)
    a0, a1 = xyz[:, 1].copy(), -xyz[:, 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        nrm = np.sqrt(a0 * a0 + a1 * a1)
        u0, u1 = a0 / nrm, a1 / nrm
    theta = 0.205 * np.pi / 180.0
    c, s = np.cos(theta), np.sin(theta)
    o00, o01, o11 = (u0 * u0).astype(np.float64), (u0 * u1).astype(np.float64), (u1 * u1).astype(np.float64)
    u0, u1 = u0.astype(np.float64), u1.astype(np.float64)
    k = 1.0 - c
    x, y, z = (xyz[:, i].astype(np.float64) for i in range(3))
    out = np.empty((xyz.shape[0], 3), np.float64)
    out[:, 0] = (c + k * o00) * x + (k * o01) * y + (s * u1) * z
    out[:, 1] = (k * o01) * x + (c + k * o11) * y + (s * -u0) * z
    out[:, 2] = (s * -u1) * x + (s * u0) * y + c * z
    return out


def project_f64(xyz: np.ndarray, H: int, W: int, up_fov=3.0, down_fov=-24.0) -> np.ndarray:
    """Projector.build_projection_map of a float64 cloud (projection.py:11-73,331-418): float64 pixel math, closest point
    per pixel (lowest index on exact range ties), empty pixels 0; [3,H,W] float64."""
    up, down = up_fov / 180.0 * np.pi, down_fov / 180.0 * np.pi
    fov = abs(down) + abs(up)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.sqrt((xyz * xyz).sum(1))
        null = r == 0.0
        rr = np.where(null, 0.001, r)
        theta = -np.arctan2(xyz[:, 1], xyz[:, 0])
        phi = np.arcsin(xyz[:, 2] / rr)
        col = 0.5 * (theta / np.pi + 1.0) * W
        row = (1.0 - (phi + abs(down)) / fov) * H
        pr, pc = np.rint(np.where(null, -1.0, row)), np.rint(np.where(null, -1.0, col))
        ok = (pr >= 0) & (pr <= H - 1) & (pc >= 0) & (pc <= W - 1) & (r > 0)
    out = np.zeros((3, H * W), np.float64)
    idx = np.nonzero(ok)[0]
    pix = (pr[idx].astype(np.int64) * W + pc[idx].astype(np.int64))
    order = np.lexsort((idx, r[idx]))            # closest first, then lowest index
    first = np.unique(pix[order], return_index=True)[1]
    win = idx[order][first]
    out[:, pix[order][first]] = xyz[win].T
    return out.reshape(3, H, W)


def poses_to_rows(poses: np.ndarray) -> np.ndarray:
    """poses_to_df (io.py:63-76): the first three rows of every pose, [N,12]."""
    assert poses.ndim == 3 and poses.shape[1:] == (4, 4)
    return poses[:, :3, :].reshape(poses.shape[0], 12)


def rows_to_poses(rows: np.ndarray) -> np.ndarray:
    """df_to_poses (io.py:44-60): float32 rows, a float64 last row (0, 0, 0, 1) -> float64 [N,4,4]."""
    rows = np.asarray(rows, dtype=np.float32)
    assert rows.shape[1] == 12
    poses = rows.reshape(-1, 3, 4)
    last = np.concatenate((np.zeros((rows.shape[0], 3)), np.ones((rows.shape[0], 1))), axis=1)[:, None, :]
    return np.concatenate((poses, last), axis=1)


def poses_csv_text(poses: np.ndarray) -> str:
    """The bytes write_poses_to_disk produces (pandas to_csv, sep ',', header 0..11, no index): every value in the shortest
    representation that round-trips in its dtype."""
    rows = poses_to_rows(poses)
    lines = [",".join(str(i) for i in range(12))]
    lines += [",".join(str(v) for v in r) for r in rows]
    return "\n".join(lines) + "\n"


def relative_poses(poses: np.ndarray) -> np.ndarray:
    """compute_relative_poses (eval_odometry.py:74-83): inv(shifted) @ poses with the identity shifted in first."""
    shifted = np.concatenate([np.eye(4)[None], poses[:-1]], axis=0)
    return np.linalg.inv(shifted) @ poses


def absolute_poses(relative: np.ndarray) -> np.ndarray:
    """compute_absolute_poses (eval_odometry.py:86-96): the running product, sequentially."""
    out = relative.copy()
    for i in range(out.shape[0] - 1):
        out[i + 1] = out[i] @ relative[i + 1]
    return out
