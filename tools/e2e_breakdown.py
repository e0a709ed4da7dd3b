"""Development helper: where the end-to-end frame goes (reference-shaped Python API, pinned host scans):
C-ABI call times vs the Python around them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pylidar_slam_b200 as b200
from pylidar_slam_b200 import _lib, synthetic as syn

H, W, F, WARM = 64, 2048, 70, 30
scans = [torch.from_numpy(syn.scan(k, H, W)).pin_memory().numpy() for k in range(F)]
dev = torch.device("cuda", 0)
cfg = b200.ICPFrameToModelConfig(local_map=b200.KdTreeLocalMapConfig(local_map_size=20),
    alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
    max_num_alignments=10, data_key="input_data")
algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0), device=dev)
algo.init()
pre = b200.Preprocessing(b200.PreprocessingConfig(filters={
    "2": dict(filter_name="grid_sample", voxel_size=0.3, pointcloud_key="numpy_pc"),
    "3": dict(filter_name="to_tensor", keys=dict(sample_points="input_data"))}))
for f in pre.filters:
    if hasattr(f, "ctx"):
        f.ctx = algo.ctx
calls = {}
orig = _lib.Context.call
def timed_call(self, name, *a):
    t0 = time.perf_counter()
    r = orig(self, name, *a)
    calls.setdefault(name, []).append(time.perf_counter() - t0)
    return r
_lib.Context.call = timed_call
t_pre, t_odo, prev = [], [], None
for k in range(F):
    if k == WARM:
        calls.clear(); t_pre.clear(); t_odo.clear()
        for slot in (0, 3, 4):
            orig(algo.ctx, "pls_profile_enable", slot, 1)
    dd = {"numpy_pc": scans[k], "init_rpose": prev}
    t0 = time.perf_counter()
    pre.forward(dd)
    t1 = time.perf_counter()
    algo.process_next_frame(dd)
    t2 = time.perf_counter()
    t_pre.append(t1 - t0); t_odo.append(t2 - t1)
    if "odometry_pose" in dd:
        prev = dd["odometry_pose"].astype(np.float64)
n = F - WARM
print(f"frame {1e6*(sum(t_pre)+sum(t_odo))/n:.0f} us = preprocessing {1e6*sum(t_pre)/n:.0f} + process_next_frame {1e6*sum(t_odo)/n:.0f}")
for name, v in calls.items():
    print(f"   C ABI {name:32s} {1e6*sum(v)/n:8.1f} us/frame ({len(v)/n:.1f} calls)")
for slot, name in ((4, "grid sample kernels"), (0, "ICP kernels"), (3, "index build (map stream)")):
    ms, launches, _ = algo.ctx.profile(slot)
    print(f"   device events, {name:26s} {1e3 * ms / n:8.1f} us/frame")
print(f"   Python around the C ABI: {1e6*(sum(t_pre)+sum(t_odo)-sum(sum(v) for v in calls.values()))/n:.0f} us/frame")
