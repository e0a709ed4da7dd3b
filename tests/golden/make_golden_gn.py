"""Goldens for gauss_newton_config.max_iters > 1 inside the ICP loop (alignment.py:69-77,110-127), from the UNMODIFIED
reference under oracle/ref_shims.py.  Build container only:  python tests/golden/make_golden_gn.py -> icp_gn.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)
ns = ref_shims.load_reference(kdtree_workers=-1)
H, W, VOXEL, FRAMES = 32, 512, 0.4, 6
out = {}
for name, lm, key in (("kd_gn3", "kdtree", "numpy_pc"), ("proj_gn2", "projective", "vertex_map")):
    gn_iters = 3 if lm == "kdtree" else 2
    lmc = ns.local_map.KdTreeLocalMapConfig(local_map_size=4) if lm == "kdtree" else ns.local_map.ProjectiveLocalMapConfig(local_map_size=4)
    cfg = ns.icp.ICPFrameToModelConfig(
        data_key=key, max_num_alignments=5, device="cpu", threshold_delta_pose=0.0, local_map=lmc,
        alignment=ns.alignment.GaussNewtonPointToPlaneConfig(
            gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=gn_iters, norm_stop_criterion=1e-9)))
    algo = ns.icp.ICPFrameToModel(cfg, projector=ns.projection.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                  pose=ns.pose.Pose("euler"), device=torch.device("cpu"))
    algo.init()
    prev, poses = None, []
    for k in range(FRAMES):
        if torch.get_num_threads() != 1:
            torch.set_num_threads(1)
        pc = syn.scan(k, H, W)
        if key == "numpy_pc":
            pc, _ = ns.pointcloud.grid_sample(pc, VOXEL)
            dd = {"numpy_pc": pc}
        else:
            dd = {"vertex_map": torch.from_numpy(syn.vertex_map_from_scan(pc, H, W))}
        dd["init_rpose"] = prev
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            prev = dd["odometry_pose"].astype(np.float64)
            poses.append(dd["odometry_pose"].copy())
    out[f"{name}_poses"] = np.stack(poses)
    out[f"{name}_gn_iters"] = np.array(gn_iters)
    print(name, np.stack(poses)[:, :3, 3])
np.savez_compressed(os.path.join(HERE, "icp_gn.npz"), **out)
