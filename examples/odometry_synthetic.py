#!/usr/bin/env python
"""The shipped grid_sample pipeline of pyLiDAR-SLAM on a synthetic 64x2048 stream, through the reference-shaped API of
this package (what `python run.py slam/odometry=icp_odometry_b200 slam/preprocessing=grid_sample ...` does per frame):

    GridSample(voxel 0.3) -> ToTensor -> ICPFrameToModel.process_next_frame      (needs a B200; there is no CPU path)

    python examples/odometry_synthetic.py [frames]

Prints the per-frame relative pose error against the stream's ground truth and the frames per second.
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import pylidar_slam_b200 as b200                     # noqa: E402
from pylidar_slam_b200 import synthetic as syn       # noqa: E402


def main(frames: int = 50, height: int = 64, width: int = 2048, voxel: float = 0.3, device: str = "cuda:0"):
    preprocessing = b200.Preprocessing(b200.PreprocessingConfig(filters={
        "1": dict(filter_name="grid_sample", voxel_size=voxel, pointcloud_key="numpy_pc"),
        "2": dict(filter_name="to_tensor", keys=dict(sample_points="input_data"))}))
    odometry = b200.ICPFrameToModel(
        b200.ICPFrameToModelConfig(
            local_map=b200.KdTreeLocalMapConfig(local_map_size=20),
            alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
            max_num_alignments=10, data_key="input_data"),
        projector=b200.SphericalProjector(height=height, width=width, up_fov=3.0, down_fov=-24.0), device=device)
    odometry.init()
    previous, errors, t0 = None, [], time.perf_counter()
    for k in range(frames):
        data_dict = {"numpy_pc": syn.scan(k, height, width), "init_rpose": previous}      # constant-velocity initialisation
        preprocessing.forward(data_dict)
        odometry.process_next_frame(data_dict)
        if "odometry_pose" in data_dict:                                                  # (nothing on frame 0, like the reference)
            previous = data_dict["odometry_pose"].astype(np.float64)
            truth = np.linalg.inv(syn.gt_pose(k - 1)) @ syn.gt_pose(k)
            errors.append(float(np.linalg.norm(previous[:3, 3] - truth[:3, 3])))
    elapsed = time.perf_counter() - t0
    poses = odometry.get_relative_poses()
    print(f"{frames} frames ({height}x{width}, voxel {voxel}): {frames / elapsed:.1f} frames/s including the synthetic ray casting; "
          f"translation error vs ground truth: mean {1e3 * np.mean(errors):.1f} mm, max {1e3 * np.max(errors):.1f} mm; "
          f"{len(poses)} relative poses")
    return poses, errors


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 50)
