"""Development helper: config-3/5 style projective run (vertex-map input, K -> 20) with the in-library
CUDA-event profile of proj_icp_iter_kernel (slot 1) and the model rebuild (slot 2)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pylidar_slam_b200 as b200
from pylidar_slam_b200 import synthetic as syn

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 2048)
F = int(sys.argv[3]) if len(sys.argv) > 3 else 26
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
cfg = b200.ICPFrameToModelConfig(local_map=b200.ProjectiveLocalMapConfig(local_map_size=20),
    alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
    max_num_alignments=iters, data_key="vertex_map", threshold_delta_pose=0.0 if iters == 20 else 1e-4)
algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0), device="cuda:0")
algo.init()
vms = [torch.from_numpy(syn.vertex_map_from_scan(syn.scan(k, H, W), H, W)).cuda() for k in range(F)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
prev, times = None, []
for k in range(F):
    if k == F - 5:
        algo.ctx.call("pls_profile_enable", 1, 1)
        algo.ctx.call("pls_profile_enable", 2, 1)
    flush.zero_()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dd = {"vertex_map": vms[k], "init_rpose": prev}
    algo.process_next_frame(dd)
    times.append(time.perf_counter() - t0)
    if "odometry_pose" in dd:
        prev = dd["odometry_pose"].astype(np.float64)
ms, n, by = algo.ctx.profile(1)
ms2, n2, by2 = algo.ctx.profile(2)
gt = syn.gt_relative_pose(F - 1)
print(f"{H}x{W} K=20 iters<={iters}: frame {1e3*np.mean(times[-5:]):.3f} ms; proj_icp_iter: {n} executed launches, "
      f"{1e3*ms/max(n,1):.1f} us/launch, {by/max(n,1)/1e6:.1f} MB/launch algorithmic -> {by/ms/1e6 if ms else 0:.0f} GB/s; "
      f"model rebuild {1e3*ms2/max(n2,1):.1f} us, {by2/ms2/1e6 if ms2 else 0:.0f} GB/s; terr {np.abs(prev[:3,3]-gt[:3,3]).max():.4f}")
