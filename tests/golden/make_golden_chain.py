"""Golden poses of the WHOLE shipped pipeline run by the unmodified reference (imported under oracle/ref_shims.py):

    Distortion -> GridSample(0.4) -> ToTensor -> ICPFrameToModel (kd-tree map, point-to-plane GN)

on 32x512 synthetic scans with per-point timestamps, constant-velocity initialisation.  The de-skewed frame is float64,
so `input_data` is a float64 tensor: the reference projects it in float64 (icp_odometry.py:331-352).  Two variants:
the yaml stop rule (threshold_delta_pose 1e-4) with per-iteration losses, and 6 fixed iterations.  A float64
`numpy_pc` run (data_key=numpy_pc fed with the float64 samples) covers the ndarray layout.

    python tests/golden/make_golden_chain.py        (build container only)  ->  chain_icp.npz
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_golden as mg  # noqa: E402  (re-uses make_algo / single_thread; importing it does not regenerate anything)
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

pre = importlib.import_module("slam.preprocessing")
H, W, FRAMES, VOXEL = 32, 512, 7, 0.4


def timestamps_for(points: np.ndarray, seed: int) -> np.ndarray:
    az = np.arctan2(points[:, 1].astype(np.float64), points[:, 0].astype(np.float64))
    return 1.6e9 + 0.1 * ((az + np.pi) / (2 * np.pi)) + np.random.RandomState(seed).uniform(0, 1e-4, points.shape[0])


def run(data_key, max_iters, thr):
    chain = [pre.Distortion(pre.DistortionConfig(output_key="distorted")),
             pre.GridSample(pre.GridSampleConfig(voxel_size=VOXEL, pointcloud_key="distorted")),
             pre.ToTensor(pre.ToTensorConfig(keys=dict(sample_points="input_data")))]
    algo = mg.make_algo("kdtree", H, W, data_key, max_iters=max_iters, lm_size=4, threshold_delta_pose=thr)
    poses, losses, dtypes, prev = [], [], set(), None
    orig = algo.register_new_frame

    def wrapped(*a, **k):
        p, T, ls = orig(*a, **k)
        losses.append(np.array([float(x) for x in ls] + [np.nan] * (algo.gn_max_iters - len(ls)), dtype=np.float64))
        return p, T, ls
    algo.register_new_frame = wrapped
    for k in range(FRAMES):
        pc = syn.scan(k, H, W)
        dd = {"numpy_pc": pc, "numpy_pc_timestamps": timestamps_for(pc, k), "init_rpose": prev}
        for f in chain:
            mg.single_thread()
            f.filter(dd)
        if data_key == "numpy_pc":  # ndarray layout fed with the (float64 once de-skewed) samples
            dd["numpy_pc"] = dd["sample_points"]
        dtypes.add(str(dd["sample_points"].dtype))
        mg.single_thread()
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            poses.append(dd["odometry_pose"].copy())
            prev = dd["odometry_pose"].astype(np.float64)
    return np.stack(poses), np.stack(losses), dtypes


def main():
    out = {}
    for name, key, iters, thr in (("tensor", "input_data", 8, 1e-4), ("tensor_fixed6", "input_data", 6, 0.0),
                                  ("ndarray_fixed6", "numpy_pc", 6, 0.0)):
        poses, losses, dtypes = run(key, iters, thr)
        out[f"{name}_poses"], out[f"{name}_losses"] = poses, losses
        print(name, poses.shape, dtypes, "max t err vs gt",
              max(np.abs(poses[i][:3, 3] - syn.gt_relative_pose(i + 1)[:3, 3]).max() for i in range(len(poses))))
    np.savez_compressed(os.path.join(HERE, "chain_icp.npz"), **out)


if __name__ == "__main__":
    main()
