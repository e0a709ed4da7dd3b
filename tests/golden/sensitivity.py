"""How far do 1-ulp input differences move the REFERENCE's own poses?  (TEST INFRASTRUCTURE; build container only.)

Some parity tests of tests/test_gpu_parity.py compare configurations in which the reference itself is numerically
ill-conditioned (float32 normal maps from an uncentred second-moment inverse, few iterations).  A GPU path that is
arithmetically equivalent but not bit-identical cannot be closer to the reference than the reference is to itself
under perturbations of the last bit of its inputs.  This script measures that: the unmodified reference (under
oracle/ref_shims.py) runs each configuration once on the seeded scans and three more times with every coordinate
multiplied by (1 + s 2^-23), s uniform in {-1, 0, +1} (seeds 0-2); the largest pose deviation from the unperturbed run
is recorded per configuration in tests/golden/sensitivity.json.  The tests then use
max(north-star tolerance, 1.5 x measured) and cite the file.

    python tests/golden/sensitivity.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import pose_errors  # noqa: E402
from oracle import ref_shims  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)  # deterministic closest-wins scatter
ns = ref_shims.load_reference(kdtree_workers=-1)


def perturb(a, seed):
    if seed is None:
        return a
    rng = np.random.RandomState(seed)
    s = rng.randint(-1, 2, size=a.shape).astype(np.float32)
    return (a * (np.float32(1.0) + s * np.float32(2.0 ** -23))).astype(np.float32)


def run(lm, H, W, key, iters, frames, seed, lm_size=4, thr=1e-4, nan_case=False):
    if torch.get_num_threads() != 1:
        torch.set_num_threads(1)
    lmc = ns.local_map.KdTreeLocalMapConfig(local_map_size=lm_size) if lm == "kdtree" else \
        ns.local_map.ProjectiveLocalMapConfig(local_map_size=lm_size)
    cfg = ns.icp.ICPFrameToModelConfig(
        data_key=key, max_num_alignments=iters, device="cpu", threshold_delta_pose=thr, local_map=lmc,
        alignment=ns.alignment.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)))
    algo = ns.icp.ICPFrameToModel(cfg, projector=ns.projection.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                  pose=ns.pose.Pose("euler"), device=torch.device("cpu"))
    algo.init()
    prev, poses = None, []
    for k in range(frames):
        pc = perturb(syn.scan(k, H, W), None if seed is None else 1000 * seed + k).copy()
        if nan_case:
            pc[5::97] = np.nan
        if key == "vertex_map":
            dd = {"vertex_map": torch.from_numpy(syn.vertex_map_from_scan(pc, H, W))}
        else:
            if nan_case:
                pc, _ = ns.pointcloud.grid_sample(pc[~np.isnan(pc).any(1)], 0.4)
                pc = pc.copy()
                pc[3::41, 1] = np.nan
            dd = {"numpy_pc": pc}
        dd["init_rpose"] = prev
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            prev = dd["odometry_pose"].astype(np.float64)
            poses.append(prev)
    return poses


CASES = {
    # name: (local map, H, W, data key, alignments, frames, kwargs)      <- the test each one backs
    "proj_vmap_32x512": ("projective", 32, 512, "vertex_map", 8, 7, {}),                       # test_icp_projective_small[proj_vmap]
    "proj_ndarray_32x512": ("projective", 32, 512, "numpy_pc", 8, 7, {}),                     # test_icp_projective_small[proj_ndarray]
    "nan_kd_ndarray_32x512": ("kdtree", 32, 512, "numpy_pc", 8, 3, dict(thr=0.0, nan_case=True)),     # test_nan_rows_and_nan_pixels
    "nan_proj_vmap_32x512": ("projective", 32, 512, "vertex_map", 8, 3, dict(thr=0.0, nan_case=True)),
    "cfg5_proj_128x4096_20it": ("projective", 128, 4096, "vertex_map", 20, 3, dict(thr=0.0, lm_size=20)),  # test_cfg5_projective
}

out = {"perturbation": "each coordinate times (1 + s 2^-23), s in {-1, 0, 1} uniformly, seeds 0..2", "cases": {}}
only = sys.argv[1:]
for name, (lm, H, W, key, iters, frames, kw) in CASES.items():
    if only and name not in only:
        continue
    base = run(lm, H, W, key, iters, frames, None, **kw)
    per_frame = np.zeros((len(base), 2))
    for seed in range(3):
        got = run(lm, H, W, key, iters, frames, seed, **kw)
        for k, (a, b) in enumerate(zip(got, base)):
            dt, ang = pose_errors(a, b)
            per_frame[k] = np.maximum(per_frame[k], [dt, ang])
    out["cases"][name] = {"max_rel_translation": float(per_frame[:, 0].max()), "max_rotation_rad": float(per_frame[:, 1].max()),
                          "per_frame_rel_translation": [float(v) for v in per_frame[:, 0]],
                          "per_frame_rotation_rad": [float(v) for v in per_frame[:, 1]]}
    print(name, out["cases"][name], flush=True)
path = os.path.join(HERE, "sensitivity.json")
if only and os.path.exists(path):
    old = json.load(open(path))
    old["cases"].update(out["cases"])
    out = old
json.dump(out, open(path, "w"), indent=1)
