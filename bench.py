#!/usr/bin/env python
"""bench.py -- ICP odometry frames/s on synthetic 64x2048 scans (BASELINE.json metric).

One "step" = one frame of the BASELINE config-2 pipeline: GridSample(voxel 0.3) -> ToTensor ->
ICPFrameToModel (kd-tree local map of 20 frames, point-to-plane Gauss-Newton, geman_mcclure
sigma 0.3, <= 10 alignments, constant-velocity initialisation), on a seeded synthetic stream.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

* `value`  : frames/s with every raw scan already resident in HBM; one C-ABI call per frame
             (pls_process_frame_grid_sample) on device pointers; CUDA events on the library's
             stream; L2 flushed (untimed) between frames.
* `e2e`    : frames/s through the reference-shaped Python API (Preprocessing[GridSample, ToTensor]
             -> ICPFrameToModel.process_next_frame) from PINNED HOST buffers, host<->device copies
             inside the timed region.
* roofline : the kd correspondence+reduction kernels of one ICP iteration (kd_nn_group_kernel<4> +
             kd_normals_group_kernel<4> + kd_residual_kernel, which also runs the fused solve), CUDA-event
             timed inside the library during the timed frames (a second pass over the same frames, so the
             event records do not perturb `value`).
* cpu_baseline / --impl reference: the CPU oracle port of the reference path (oracle/) on the host
             cores, on a bounded sample of the same stream.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

H, W, VOXEL = 64, 2048, 0.3
MAX_ALIGN, LM_SIZE, SCHEME, SIGMA = 10, 20, "geman_mcclure", 0.3
WORKLOAD = ("cfg2: icp_odometry + grid_sample(0.3) preprocessing, synthetic 64x2048 rotating-LiDAR stream, "
            "kd-tree local map (20 frames), point-to-plane GN geman_mcclure 0.3, <=10 alignments, CV init")


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index=0):
        self.index, self.rows, self._p, self._t = index, [], None, None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self._p = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                                        "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._p = None
            return
        self._t = threading.Thread(target=self._read, daemon=True)
        self._t.start()

    def _read(self):
        for line in self._p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self._p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self._p.terminate()
        try:
            self._p.wait(timeout=2)
        except Exception:
            self._p.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for nme, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_scans(n_frames):
    from pylidar_slam_b200 import synthetic as syn
    return [syn.scan(k, H, W) for k in range(n_frames)]


# ------------------------------------------------------------------------------------------ CPU arm
def run_cpu_port(scans, warmup, steps, threads=None):
    """The oracle port of the reference path on the host cores; returns (fps, ms/frame list)."""
    import torch
    from oracle import icp_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    cfg = orc.ICPConfig(max_num_alignments=MAX_ALIGN, data_key="input_data", local_map="kdtree", local_map_size=LM_SIZE,
                        scheme=SCHEME, sigma=SIGMA)
    algo = orc.ICPFrameToModelOracle(cfg, orc.Projector(H, W))
    prev, times = None, []
    for k in range(warmup + steps):
        t0 = time.perf_counter()
        s, _ = orc.grid_sample(scans[k], VOXEL)                       # GridSample.filter
        dd = {"input_data": torch.from_numpy(s), "init_rpose": prev}  # ToTensor
        algo.process_next_frame(dd)
        dt = time.perf_counter() - t0
        if "odometry_pose" in dd:
            prev = dd["odometry_pose"].astype(np.float64)
        if k >= warmup:
            times.append(dt)
    return len(times) / sum(times), times


def reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded sample: the map needs ~20 frames to reach steady state; cap the CPU work at ~60 frames
    warmup = min(args.warmup, 24)
    steps = min(args.steps, 30)
    scans = make_scans(warmup + steps)
    t0 = time.perf_counter()
    fps, times = run_cpu_port(scans, warmup, steps)
    line = {
        "impl": "reference", "metric": "icp_odometry_frames_per_sec", "value": fps, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * float(np.mean(times)),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "voxel": VOXEL},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"frames {warmup}..{warmup + steps - 1} of the same seeded stream after {warmup} warm-up "
                                   f"frames (oracle/icp_oracle.py: torch CPU + scipy cKDTree workers=-1), "
                                   f"{time.perf_counter() - t0:.1f} s wall"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------ GPU arm
def b200_arm(args):
    import torch
    import torch.distributed as dist
    import pylidar_slam_b200 as b200
    from pylidar_slam_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    W_, K_ = args.warmup, args.steps
    n_frames = 1 + W_ + K_
    scans = make_scans(n_frames)
    n_raw = scans[0].shape[0]
    stream = torch.cuda.Stream(device=dev)
    projector = b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0)

    comm_used = [args.comm]  # init_comm reports a fall-back from p2p to nccl

    def make_algo():
        cfg = b200.ICPFrameToModelConfig(
            local_map=b200.KdTreeLocalMapConfig(local_map_size=LM_SIZE),
            alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme=SCHEME, sigma=SIGMA, max_iters=1)),
            max_num_alignments=MAX_ALIGN, data_key="input_data")
        algo = b200.ICPFrameToModel(cfg, projector=projector, device=dev, stream=stream.cuda_stream)
        algo.init()
        if world > 1:
            from pylidar_slam_b200.distributed import init_comm
            comm_used[0] = init_comm(algo.ctx, dist, rank, world, dev, mode=args.comm)
        return algo

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def flush_l2():
        with torch.cuda.stream(stream):
            flush_buf.zero_()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: inputs resident in HBM, one C-ABI call per frame
    dev_scans = torch.from_numpy(np.stack(scans)).to(dev)
    torch.cuda.synchronize(dev)

    def device_pass(profile_slot=None, flush_each_step=False):
        """W warm-up frames, then exactly K timed frames inside ONE CUDA-event bracket on the library's
        stream (the local-map update of frame k runs on the library's second stream and overlaps frame
        k+1, so per-step brackets with untimed gaps would hide work; the bracket closes only after
        pls_synchronize has drained both streams)."""
        algo = make_algo()
        ctx = algo.ctx
        pose = np.zeros((4, 4), np.float32)
        params = np.zeros(6, np.float32)
        info = np.zeros(12, np.float64)
        import ctypes as C
        has = C.c_int(0)
        prev = None
        launches0, iters = 0, []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in range(n_frames):
            timed = k > W_
            if k == W_ + 1:
                ctx.call("pls_synchronize")
                flush_l2()
                barrier()
                launches0 = ctx.launch_count()
                if profile_slot is not None:
                    ctx.call("pls_profile_enable", profile_slot, 1)
                e0.record(stream)
            if timed and flush_each_step:
                flush_l2()        # inside the bracket: counted
            ctx.call("pls_process_frame_grid_sample", dev_scans[k].data_ptr(), n_raw, VOXEL, _lib.INPUT_TENSOR,
                     _lib.ptr(prev), _lib.ptr(pose), _lib.ptr(params), C.byref(has), _lib.ptr(info))
            if timed:
                iters.append(int(info[0]))
            if has.value:
                prev = pose.copy()
        ctx.call("pls_synchronize")
        e1.record(stream)
        barrier()
        launches = ctx.launch_count() - launches0
        total_ms = e0.elapsed_time(e1)
        prof = ctx.profile(profile_slot) if profile_slot is not None else None
        stats = {"samples": int(info[4]), "queries": int(info[2]), "map_points": int(info[3]), "iters_mean": float(np.mean(iters))}
        return total_ms, launches, prof, stats, ctx

    clocks = ClockSampler(local_rank)
    clocks.start()
    ms_dev, launches, _, stats, _ = device_pass()
    clock_info = clocks.stop()
    if args.quick:
        if world > 1:
            t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_dev = float(t[0])
            dist.destroy_process_group()
        if rank == 0:
            print(json.dumps({"quick": True, "n_gpus": world, "comm": comm_used[0] if world > 1 else None,
                              "ms_per_step": ms_dev / K_, "gpu_launches": launches, **stats}))
        return
    ms_dev_flushed, _, _, _, _ = device_pass(flush_each_step=True)
    # roofline passes: same frames with CUDA events around one kernel family inside the library
    _, _, prof_nn, _, _ = device_pass(profile_slot=0)
    _, _, prof_idx, _, _ = device_pass(profile_slot=3)
    _, _, prof_gs, _, _ = device_pass(profile_slot=4)

    # ---------------- e2e: reference-shaped Python API from pinned host buffers
    pinned = [torch.from_numpy(s).pin_memory() for s in scans]
    host_scans = [p.numpy() for p in pinned]
    pre = b200.Preprocessing(b200.PreprocessingConfig(filters={
        "2": dict(filter_name="grid_sample", voxel_size=VOXEL, pointcloud_key="numpy_pc"),
        "3": dict(filter_name="to_tensor", keys=dict(sample_points="input_data"))}))
    algo = make_algo()
    gs_ctx = algo.ctx
    for f in pre.filters:
        if hasattr(f, "ctx"):
            f.ctx = gs_ctx
    prev, h2d, d2h = None, 0, 0
    t_start = 0.0
    for k in range(n_frames):
        timed = k > W_
        if k == W_ + 1:
            gs_ctx.call("pls_synchronize")
            flush_l2()
            barrier()
            t_start = time.perf_counter()
        dd = {"numpy_pc": host_scans[k], "init_rpose": prev}
        pre.forward(dd)
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            prev = dd["odometry_pose"].astype(np.float64)
        if timed:
            S = dd["sample_points"].shape[0]
            h2d += host_scans[k].nbytes + S * 12 + 64
            d2h += S * 12 + S * 8 + 8 + 1400  # samples + indices + count + FrameResult
    gs_ctx.call("pls_synchronize")
    t_e = time.perf_counter() - t_start
    barrier()

    # ---------------- aggregate over ranks (max of the per-rank time)
    t_dev = ms_dev / 1e3
    if world > 1:
        t = torch.tensor([t_dev, t_e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peaks()
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r1_kd_traffic.json")))["traffic_bytes_per_launch"]
    except Exception:
        pass
    nn_ms, nn_launches, nn_bytes = prof_nn
    achieved = (nn_bytes / max(nn_launches, 1)) / (nn_ms / max(nn_launches, 1) * 1e-3) / 1e9 if nn_ms > 0 else 0.0
    frame_ms = ms_dev / K_
    line = {
        "metric": "icp_odometry_frames_per_sec", "value": K_ / t_dev, "unit": "frames/s", "n_gpus": world,
        "steps": K_, "warmup": W_, "ms_per_step": 1e3 * t_dev / K_, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "height": H, "width": W, "voxel": VOXEL,
                   "l2": "L2 flushed once before the timed bracket; every step streams a NEW 1.5 MB scan from HBM while the "
                         "62 MB local map legitimately stays L2-resident across frames (production behaviour); "
                         "value_l2_flushed_every_step re-measures with a 256 MiB flush INSIDE the bracket before every frame",
                   "value_l2_flushed_every_step": K_ / (ms_dev_flushed / 1e3),
                   "parallelism": "1 GPU" if world == 1 else f"queries sharded over {world} GPUs, map replicated, "
                                                             f"one 30-double all-reduce per ICP iteration ({comm_used[0]})",
                   **stats},
        "e2e": {"value": K_ / t_e, "unit": "frames/s", "h2d_bytes_per_step": int(h2d / K_), "d2h_bytes_per_step": int(d2h / K_),
                "ms_per_step": 1e3 * t_e / K_},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": {"bound": "hbm", "kernel": "kd_nn_group_kernel<4> + kd_normals_group_kernel<4> + kd_residual_kernel: one ICP iteration "
                               "(exact 1-NN, lazy 10-NN normals, point-to-plane reduction + fused solve)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "launches": int(nn_launches), "avg_us": 1e3 * nn_ms / max(nn_launches, 1),
                     "algorithmic_bytes_per_launch": nn_bytes / max(nn_launches, 1),
                     "share_of_step": (nn_ms / K_) / frame_ms},
        "kernels": {"index_build_ms_per_frame": prof_idx[0] / K_, "grid_sample_ms_per_frame": prof_gs[0] / K_,
                    "correspondence_ms_per_frame": nn_ms / K_},
    }
    if world == 1 and not args.no_cpu:
        t0 = time.perf_counter()
        nb = min(len(scans), 30)
        fps_cpu, times = run_cpu_port(scans[:nb], min(22, nb - 6), nb - min(22, nb - 6))
        line["cpu_baseline"] = {"value": fps_cpu, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"frames {min(22, nb - 6)}..{nb - 1} of the same stream (oracle port: torch CPU + scipy "
                                          f"cKDTree workers=-1), {time.perf_counter() - t0:.1f} s wall"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--quick", action="store_true", help="device-resident pass only (for ncu captures)")
    ap.add_argument("--comm", default="p2p", choices=["p2p", "nccl"],
                    help="N>1 exchange: one-shot NVLink peer-to-peer all-reduce fused into the solve kernel, or NCCL")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
