"""Development helper (runs without a GPU): the Python cost of one end-to-end frame of the reference-shaped API
(GridSample -> ToTensor -> ICPFrameToModel.process_next_frame) with the C ABI replaced by a stand-in that returns at
once (fixed sample count, identity pose).  What remains is what the interpreter adds to every frame.  (Without a GPU the
pinned pool cannot allocate, so GridSample hands out pageable COPIES of the staged samples: about 30 us of memcpy per frame
that a GPU box does not pay -- compare runs of this tool with each other, not with profiles/r2_e2e_breakdown.log.)
    python tools/host_overhead.py [frames] [--profile]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import pylidar_slam_b200 as b200
from pylidar_slam_b200 import _lib, synthetic as syn

H, W, S = 64, 2048, 31600
staging_xyz = np.random.rand(S, 3).astype(np.float32)
staging_idx = np.arange(S, dtype=np.int64)


class NullContext:
    """Every entry point succeeds immediately; outputs get plausible values."""

    def __init__(self, **kwargs):
        self.cfg = _lib.PlsConfig()
        for k, v in kwargs.items():
            setattr(self.cfg, k, v)
        self.handle = 1
        self.frames = 0

    def close(self):
        pass

    def call(self, name, *a):
        if name == "pls_grid_sample_staged":
            a[4]._obj.value = staging_xyz.ctypes.data
            a[5]._obj.value = staging_idx.ctypes.data
            a[6]._obj.value = 0xdead0000
            a[7]._obj.value = S
        elif name == "pls_process_frame":
            pose = _lib_arr(a[4], (4, 4), np.float32)
            pose[:] = np.eye(4)
            a[6]._obj.value = 1 if self.frames else 0
            self.frames += 1
        return 0


def _lib_arr(addr, shape, dtype):
    return _lib.host_view(int(addr), shape, dtype)


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 300
    _lib.Context = NullContext
    _lib.host_fingerprint = lambda address, nbytes: 7
    import pylidar_slam_b200.odometry as odo
    odo._lib.Context = NullContext
    scans = [syn.scan(k, H, W) for k in range(4)]
    cfg = b200.ICPFrameToModelConfig(local_map=b200.KdTreeLocalMapConfig(local_map_size=20),
        alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
        max_num_alignments=10, data_key="input_data")
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                device=torch.device("cuda", 0))
    algo.init()
    pre = b200.Preprocessing(b200.PreprocessingConfig(filters={
        "2": dict(filter_name="grid_sample", voxel_size=0.3, pointcloud_key="numpy_pc"),
        "3": dict(filter_name="to_tensor", keys=dict(sample_points="input_data"))}))
    for f in pre.filters:
        if hasattr(f, "ctx"):
            f.ctx = algo.ctx

    def loop(n):
        prev = None
        for k in range(n):
            dd = {"numpy_pc": scans[k & 3], "init_rpose": prev}
            pre.forward(dd)
            algo.process_next_frame(dd)
            if "odometry_pose" in dd:
                prev = dd["odometry_pose"].astype(np.float64)

    loop(20)
    if "--profile" in sys.argv:
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        loop(frames)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(28)
    t0 = time.perf_counter()
    loop(frames)
    print(f"Python per frame with a no-op C ABI: {1e6 * (time.perf_counter() - t0) / frames:.1f} us")


if __name__ == "__main__":
    main()
