"""bench.py's CPU arm (the oracle port timed on the host cores): the thread calibration replays frames on COPIES of the
algorithm's state and must leave the timed stream untouched; the --impl reference line keeps the contract's keys."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_thread_calibration_does_not_disturb_the_stream():
    import torch
    import bench
    before = torch.get_num_threads()
    try:
        scans = bench.make_scans(5)
        _, t_plain, _, tried_plain, pose_plain = bench.run_cpu_port(scans, 2, 3, threads=1)
        _, t_cal, best, tried, pose_cal = bench.run_cpu_port(scans, 2, 3, calibrate=True)
        assert tried_plain == {} and len(t_plain) == len(t_cal) == 3
        assert 1 in tried and best in tried and all(ms > 0 for ms in tried.values())
        # same frames, same arithmetic: the calibration ran on deep copies of the state after the warm-up
        np.testing.assert_allclose(pose_cal, pose_plain, rtol=0, atol=1e-6)
    finally:
        torch.set_num_threads(before)


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "icp_odometry_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 3 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) / d["value"] < 1e-6
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == os.cpu_count() and cb["torch_threads"] >= 1
    assert d["config"]["workload"].startswith("cfg2") and d["config"]["height"] == 64 and d["config"]["width"] == 2048


def test_reference_arm_is_silent_on_other_ranks():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert out.returncode == 0 and out.stdout.strip() == ""
