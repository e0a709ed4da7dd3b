"""Runs the four kernels either side of the hot path once each at the BASELINE scan sizes on device-resident inputs
(meant to run under `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`), and prints
the algorithmic bytes per launch so that achieved GB/s can be set against the HBM roofline."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pylidar_slam_b200 as b200  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

out = {}
for H, W in ((64, 2048), (128, 4096)):
    pc = syn.scan(3, H, W)
    n = pc.shape[0]
    az = np.arctan2(pc[:, 1].astype(np.float64), pc[:, 0].astype(np.float64))
    ts = 1.6e9 + 0.1 * ((az + np.pi) / (2 * np.pi))
    pose = syn.gt_relative_pose(3)
    d_pc, d_ts = torch.from_numpy(pc).cuda(), torch.from_numpy(ts).cuda()
    for rep in range(2):
        d = b200.distort_frame(d_pc, d_ts, pose)
        stats = b200.voxel_statistics(d_pc, 0.2)
    V = int(stats[2].shape[0])
    # a general 6-DoF motion plus noise: a planar motion leaves the z column of the reference's point-to-point Jacobian
    # identically zero and its normal equations singular (the reference raises there too)
    from scipy.spatial.transform import Rotation
    Rg = Rotation.from_euler("xyz", [0.004, -0.003, 0.006]).as_matrix()
    moved = pc.astype(np.float64) @ Rg.T + np.array([0.05, -0.03, 0.02]) + np.random.RandomState(0).normal(0, 0.01, pc.shape)
    moved = torch.from_numpy(moved.astype(np.float32)).cuda()
    al = b200.RIGID_ALIGNMENT.load(dict(mode="point_to_point_gauss_newton", gauss_newton_config=dict(scheme="huber", sigma=0.3, max_iters=1)))
    for rep in range(2):
        al.align(moved.unsqueeze(0), d_pc.unsqueeze(0))
        T = b200.weighted_procrustes(d_pc, moved)
    out[f"{H}x{W}"] = {
        "points": n, "voxels": V,
        "distort_kernel_bytes": n * (12 + 8 + 24), "ts_minmax_kernel_bytes": n * 8,
        "voxel_stats_kernel_bytes": n * (4 + 12) * 2 + V * (4 + 8 + 12 + 36),
        "gn_accumulate_kernel_p2point_bytes": n * (24 + 4) + 240,
        "procrustes_moments_kernel_bytes": n * 24, "procrustes_cross_kernel_bytes": n * 24,
        "procrustes_residual_max": float(np.abs(T[:3, :3] - Rg).max()),
    }
# training loss, batch of 4 pairs of 64x2048 vertex maps
H, W, B = 64, 2048, 4
pairs = np.stack([np.stack([syn.vertex_map_from_scan(syn.scan(k, H, W), H, W)[0], syn.vertex_map_from_scan(syn.scan(k + 1, H, W), H, W)[0]])
                  for k in range(B)]).astype(np.float32)
vm = torch.from_numpy(pairs).cuda()
mod = b200._PointToPlaneLossModule(b200.PointToPlaneLossConfig(), b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0))
x = torch.tensor([[0.8, 0.0, 0.0, 0.0, 0.0, 0.01]] * B, device="cuda", requires_grad=True)
for rep in range(2):
    loss, dd = mod({"vertex_map": vm, "pose_params": x})
    loss.backward()
out["training_loss"] = {"batch": B, "pixels": H * W, "loss": float(loss.detach()),
                        "loss_zbuf_kernel_bytes": B * H * W * (12 + 8), "loss_accumulate_kernel_bytes": B * H * W * (12 + 8 + 12 + 24)}
torch.cuda.synchronize()
print(json.dumps(out))
