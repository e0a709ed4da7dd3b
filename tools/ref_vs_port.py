"""Times the UNMODIFIED reference (imported under oracle/ref_shims.py) and the oracle port on the same cfg2 frames, in the
build container (the GPU box has no /root/reference), and writes profiles/ref_vs_port.json -- so that the
`cpu_baseline.kind: "port"` number bench.py reports has a stated relation to the real thing.

Both run the shipped pipeline: GridSample(voxel 0.3) -> ToTensor -> ICPFrameToModel (kd-tree local map of 20 frames,
point-to-plane GN geman_mcclure 0.3, <= 10 alignments, CV init) on the seeded 64x2048 stream; warm-up frames fill the
map (and pay numba's JIT), the timed frames follow.  The reference is timed the way it times itself: the `elapsed` list
of OdometryAlgorithm.process_next_frame (slam/odometry/odometry.py:44-46), GridSample.filter separately.  All host
threads are allowed on both sides (torch intra-op, numba prange, cKDTree workers=-1 standing in for pykdtree's OpenMP).

    python tools/ref_vs_port.py [warmup=24] [timed=30]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import icp_oracle as orc  # noqa: E402
from oracle import ref_shims  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

H, W, VOXEL = 64, 2048, 0.3
WARM = int(sys.argv[1]) if len(sys.argv) > 1 else 24
TIMED = int(sys.argv[2]) if len(sys.argv) > 2 else 30
scans = [syn.scan(k, H, W) for k in range(WARM + TIMED)]


def run_reference():
    ns = ref_shims.load_reference(kdtree_workers=-1)
    import slam.preprocessing as pre
    cfg = ns.icp.ICPFrameToModelConfig(
        data_key="input_data", max_num_alignments=10, device="cpu",
        local_map=ns.local_map.KdTreeLocalMapConfig(local_map_size=20),
        alignment=ns.alignment.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)))
    algo = ns.icp.ICPFrameToModel(cfg, projector=ns.projection.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                  pose=ns.pose.Pose("euler"), device=torch.device("cpu"))
    algo.init()
    gs = pre.GridSample(pre.GridSampleConfig(voxel_size=VOXEL, pointcloud_key="numpy_pc"))
    tt = pre.ToTensor(pre.ToTensorConfig(keys=dict(sample_points="input_data")))
    prev, t_gs = None, []
    for k, pc in enumerate(scans):
        dd = {"numpy_pc": pc, "init_rpose": prev}
        t0 = time.perf_counter()
        gs.filter(dd)
        tt.filter(dd)
        t_gs.append(time.perf_counter() - t0)
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            prev = dd["odometry_pose"].astype(np.float64)
    odo = np.array(algo.elapsed[WARM:])
    pre_t = np.array(t_gs[WARM:])
    return odo, pre_t, np.stack([p.reshape(4, 4) for p in algo.relative_poses[WARM:]])


def run_port():
    cfg = orc.ICPConfig(max_num_alignments=10, data_key="input_data", local_map="kdtree", local_map_size=20,
                        scheme="geman_mcclure", sigma=0.3)
    algo = orc.ICPFrameToModelOracle(cfg, orc.Projector(H, W))
    prev, t_odo, t_gs, poses = None, [], [], []
    for k, pc in enumerate(scans):
        t0 = time.perf_counter()
        s, _ = orc.grid_sample(pc, VOXEL)
        dd = {"input_data": torch.from_numpy(s), "init_rpose": prev}
        t1 = time.perf_counter()
        algo.process_next_frame(dd)
        t2 = time.perf_counter()
        if "odometry_pose" in dd:
            prev = dd["odometry_pose"].astype(np.float64)
            poses.append(prev)
        t_gs.append(t1 - t0)
        t_odo.append(t2 - t1)
    return np.array(t_odo[WARM:]), np.array(t_gs[WARM:]), np.stack(poses[-TIMED:])


ref_odo, ref_pre, ref_poses = run_reference()
port_odo, port_pre, port_poses = run_port()
dt = np.linalg.norm(ref_poses[:, :3, 3] - port_poses[:, :3, 3], axis=1) / np.linalg.norm(ref_poses[:, :3, 3], axis=1)
out = {
    "what": "unmodified reference (under oracle/ref_shims.py) vs the oracle port, same seeded cfg2 frames, same box",
    "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "warmup_frames": WARM, "timed_frames": TIMED,
    "reference_ms_per_frame": float(1e3 * (ref_odo.mean() + ref_pre.mean())),
    "reference_process_next_frame_ms": float(1e3 * ref_odo.mean()), "reference_grid_sample_to_tensor_ms": float(1e3 * ref_pre.mean()),
    "port_ms_per_frame": float(1e3 * (port_odo.mean() + port_pre.mean())),
    "port_process_next_frame_ms": float(1e3 * port_odo.mean()), "port_grid_sample_ms": float(1e3 * port_pre.mean()),
    "port_over_reference": float((port_odo.mean() + port_pre.mean()) / (ref_odo.mean() + ref_pre.mean())),
    "max_rel_translation_difference_of_the_timed_poses": float(dt.max()),
    "note": "kd-tree = scipy cKDTree (workers=-1) on both sides: pykdtree is not installable here (SURVEY.md 8c)",
}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "ref_vs_port.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
