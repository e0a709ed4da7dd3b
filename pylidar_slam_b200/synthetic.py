"""Seeded synthetic rotating-LiDAR stream (BASELINE.md section 2 "Synthetic input").

A level spinning LiDAR (H beams x W columns, vertical FOV up_fov..down_fov)
drives along a circular arc (0.8 m forward + 0.01 rad yaw per frame) inside a
large hall: ground + ceiling planes, four walls and a fixed pseudo-random
forest of vertical pillars.  Ranges carry 1 cm Gaussian noise seeded by the
frame index (noise is mandatory: with exact correspondences the reference's
robust weights collapse, slam/common/optimization.py:49-50,334-336).

Both bench arms and every parity test draw their scans from here, so the
GPU path and the CPU oracle always see bit-identical inputs.  Pure numpy.

Pixel convention follows slam/common/projection.py:11-73: column j <-> theta =
-atan2(y, x) with col = W/2 (theta/pi + 1); row i <-> phi = asin(z/r) with
row = H (1 - (phi + |down|)/fov).
"""
import numpy as np

SENSOR_HEIGHT = 1.8
CEILING = 8.0
HALL = (-130.0, 130.0, -50.0, 210.0)  # xmin, xmax, ymin, ymax
STEP_FORWARD = 0.8
STEP_YAW = 0.01


def _pillars():
    rng = np.random.RandomState(1234)
    xs = np.arange(HALL[0] + 10, HALL[1] - 5, 18.0)
    ys = np.arange(HALL[2] + 10, HALL[3] - 5, 18.0)
    gx, gy = np.meshgrid(xs, ys, indexing="ij")
    c = np.stack([gx.ravel(), gy.ravel()], 1) + rng.uniform(-5, 5, (gx.size, 2))
    rad = rng.uniform(0.4, 1.6, gx.size)
    # keep the circular path (radius 80 centred on (0, 80)) clear
    d = np.abs(np.linalg.norm(c - np.array([0.0, 80.0]), axis=1) - 80.0)
    keep = d > (rad + 3.5)
    return c[keep], rad[keep]


_PILLAR_C, _PILLAR_R = _pillars()


def gt_pose(k: int) -> np.ndarray:
    """World<-sensor pose of frame k (float64 4x4), circular arc of radius 80 m."""
    radius = STEP_FORWARD / (2.0 * np.sin(STEP_YAW / 2.0))
    yaw = STEP_YAW * k
    # chord-integrated positions on a circle centred at (0, radius)
    x = radius * np.sin(yaw)
    y = radius * (1.0 - np.cos(yaw))
    T = np.eye(4)
    T[:3, :3] = np.array([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0, 0, 1.0]])
    T[:3, 3] = [x, y, 0.0]
    return T


def gt_relative_pose(k: int) -> np.ndarray:
    """Pose of frame k expressed in frame k-1 (what the odometry estimates)."""
    return np.linalg.inv(gt_pose(k - 1)) @ gt_pose(k)


def _horizontal_range(origin_xy, az_world):
    """Distance along each horizontal direction to the nearest vertical surface."""
    dx, dy = np.cos(az_world), np.sin(az_world)
    ox, oy = origin_xy
    with np.errstate(divide="ignore", invalid="ignore"):
        tx = np.where(dx > 0, (HALL[1] - ox) / dx, np.where(dx < 0, (HALL[0] - ox) / dx, np.inf))
        ty = np.where(dy > 0, (HALL[3] - oy) / dy, np.where(dy < 0, (HALL[2] - oy) / dy, np.inf))
    s = np.minimum(tx, ty)
    # pillars: |o + s d - c|^2 = r^2
    rel = np.array([ox, oy])[None, :] - _PILLAR_C  # [P,2]
    b = dx[:, None] * rel[None, :, 0] + dy[:, None] * rel[None, :, 1]  # [W,P]
    cc = (rel ** 2).sum(1)[None, :] - (_PILLAR_R ** 2)[None, :]
    disc = b * b - cc
    hit = (disc > 0) & (b < 0)
    root = np.where(hit, -b - np.sqrt(np.where(hit, disc, 0.0)), np.inf)
    root = np.where(root > 0, root, np.inf)
    return np.minimum(s, root.min(axis=1))


def scan(k: int, height: int = 64, width: int = 2048, up_fov: float = 3.0, down_fov: float = -24.0,
         sigma_range: float = 0.01, dtype=np.float32) -> np.ndarray:
    """Frame k as an [H*W, 3] point cloud in the sensor frame (row-major beams x columns)."""
    fov_up = up_fov / 180.0 * np.pi
    fov_down = down_fov / 180.0 * np.pi
    fov = abs(fov_up) + abs(fov_down)
    rows = np.arange(height, dtype=np.float64)
    cols = np.arange(width, dtype=np.float64)
    phi = (1.0 - rows / height) * fov - abs(fov_down)  # [H]
    theta = (2.0 * cols / width - 1.0) * np.pi  # [W]
    az = -theta  # sensor-frame azimuth

    T = gt_pose(k)
    yaw = STEP_YAW * k
    s = _horizontal_range(T[:2, 3], az + yaw)  # [W]

    cphi, sphi = np.cos(phi)[:, None], np.sin(phi)[:, None]
    with np.errstate(divide="ignore"):
        t_vert = s[None, :] / cphi
        t_ground = np.where(sphi < 0, SENSOR_HEIGHT / np.maximum(-sphi, 1e-12), np.inf)
        t_ceil = np.where(sphi > 0, (CEILING - 0.0) / np.maximum(sphi, 1e-12), np.inf)
    t = np.minimum(t_vert, np.minimum(t_ground, t_ceil))  # [H,W]
    rng = np.random.RandomState(1000003 + k)
    t = t + rng.normal(0.0, sigma_range, t.shape)
    d = np.stack([cphi * np.cos(az)[None, :], cphi * np.sin(az)[None, :], np.broadcast_to(sphi, t.shape)], axis=-1)
    pts = (t[..., None] * d).reshape(-1, 3)
    return np.ascontiguousarray(pts.astype(dtype))


def vertex_map_from_scan(points: np.ndarray, height: int, width: int) -> np.ndarray:
    """The organised scan as a [1, 3, H, W] vertex map (no re-projection)."""
    return np.ascontiguousarray(points.reshape(height, width, 3).transpose(2, 0, 1)[None])
