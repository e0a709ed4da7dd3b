"""Development helper: per-frame timing of the cfg2 pipeline through the Python API."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pylidar_slam_b200 as b200
from pylidar_slam_b200 import synthetic as syn

H, W = 64, 2048
F = int(sys.argv[1]) if len(sys.argv) > 1 else 40
layout = sys.argv[2] if len(sys.argv) > 2 else "tensor"
scans = [syn.scan(k, H, W) for k in range(F)]
cfg = b200.ICPFrameToModelConfig(local_map=b200.KdTreeLocalMapConfig(local_map_size=20),
    alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
    max_num_alignments=10, data_key="input_data" if layout == "tensor" else "numpy_pc")
algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0), device="cuda:0")
algo.init()
prev = None
tg, ti = [], []
for k in range(F):
    t0 = time.perf_counter()
    s, idx = b200.grid_sample(scans[k], 0.3)
    t1 = time.perf_counter()
    dd = {"input_data": torch.from_numpy(s)} if layout == "tensor" else {"numpy_pc": s}
    dd["init_rpose"] = prev
    algo.process_next_frame(dd)
    t2 = time.perf_counter()
    tg.append(t1 - t0); ti.append(t2 - t1)
    if "odometry_pose" in dd:
        prev = dd["odometry_pose"].astype(np.float64)
        gt = syn.gt_relative_pose(k)
        if k % 8 == 0 or k == F - 1:
            print(k, "S", s.shape[0], "iters", int(algo.last_info[0]), "queries", int(algo.last_info[2]), "map", int(algo.last_info[3]),
                  f"gs {1e3*tg[-1]:.3f} ms icp {1e3*ti[-1]:.3f} ms terr {np.abs(prev[:3,3]-gt[:3,3]).max():.4f}")
print(f"steady-state (last 10): grid_sample {1e3*np.mean(tg[-10:]):.3f} ms, process_next_frame {1e3*np.mean(ti[-10:]):.3f} ms")

import ctypes
st = (ctypes.c_ulonglong * 16)()
algo.ctx.call("pls_kdmap_stats", st)
st = list(st)
if st[0]:
    print(f"nn: {st[0]} queries, exact at level 0 {st[1]/st[0]:.4f}, coarser levels {st[2]/st[0]:.4f}, cand/query {st[3]/st[0]:.1f}")
    print(f"knn: {st[4]} queries, exact at level 0 {st[5]/max(st[4],1):.4f}, coarser levels {st[6]/max(st[4],1):.4f}, cand/query {st[7]/max(st[4],1):.1f}")
