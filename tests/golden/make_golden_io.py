"""Golden vectors for SURVEY.md section 8f rank 4 (dataset -> vertex-map ingestion, pose I/O), produced by the
UNMODIFIED reference under oracle/ref_shims.py.  Build container only.

    python tests/golden/make_golden_io.py        -> tests/golden/io_rows.npz

  kitti_*     KITTIOdometrySequence.correct_scan (slam/dataset/kitti_dataset.py:200-231) on a synthetic HDL-64-like scan
              [N,4] (x, y, z, reflectance), then __getitem__'s projection (:241-249): numpy_pc, vertex map
  poses_*     poses_to_df / df_to_poses and the CSV text write_poses_to_disk produces (slam/common/io.py:17-76)
  rel_*, abs_ compute_relative_poses / compute_absolute_poses (slam/eval/eval_odometry.py:80-96), float64 and float32
"""
import io
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)
ref_shims.install()
import slam.common.io as ref_io  # noqa: E402
import slam.common.projection as ref_proj  # noqa: E402
import slam.dataset.kitti_dataset as ref_kitti  # noqa: E402
import slam.eval.eval_odometry as ref_eval  # noqa: E402

out = {}
rng = np.random.RandomState(11)

# ---- KITTI scan ingestion (64 x 720 projection as in config/dataset/kitti.yaml:4-7)
H, W = 64, 720
pts = syn.scan(3, 64, 1024)
keep = rng.rand(pts.shape[0]) > 0.15                      # ragged like a real scan
scan = np.concatenate([pts[keep], rng.rand(int(keep.sum()), 1).astype(np.float32)], axis=1)
scan[7] = [0.0, 0.0, 5.0, 0.3]                            # on the rotation axis: cross product 0 -> NaN axis (reference quirk)
corrected = ref_kitti.KITTIOdometrySequence.correct_scan(scan)
projector = ref_proj.SphericalProjector(height=H, width=W, num_channels=3, up_fov=3.0, down_fov=-24.0)
vmap = projector.build_projection_map(torch.from_numpy(corrected[:, :3]).unsqueeze(0))[0].numpy()
out.update(kitti_scan=scan, kitti_corrected=corrected, kitti_vmap=vmap, kitti_hw=np.array([H, W]))

# ---- pose I/O
traj = np.stack([syn.gt_pose(k) for k in range(40)]).astype(np.float32)
traj[:, :3, 3] += rng.randn(40, 3).astype(np.float32) * 0.01
df = ref_io.poses_to_df(traj)
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "seq.poses.txt")
    ref_io.write_poses_to_disk(path, traj)
    text = open(path).read()
    back = ref_io.read_poses_from_disk(path)
out.update(poses_in=traj, poses_df=df.to_numpy(), poses_csv=np.frombuffer(text.encode(), dtype=np.uint8), poses_back=back)

# ---- relative / absolute poses
for tag, dt in (("f64", np.float64), ("f32", np.float32)):
    P = np.stack([syn.gt_pose(k) for k in range(60)]).astype(dt)
    rel = ref_eval.compute_relative_poses(P)
    absolute = ref_eval.compute_absolute_poses(rel)
    out[f"rel_{tag}_in"], out[f"rel_{tag}_out"], out[f"abs_{tag}_out"] = P, rel, absolute
np.savez_compressed(os.path.join(HERE, "io_rows.npz"), **out)
for k, v in out.items():
    print(k, v.shape, v.dtype)
