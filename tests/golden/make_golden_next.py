"""Golden vectors for the rows either side of the hot path, produced by
the UNMODIFIED reference (/root/reference imported under oracle/ref_shims.py).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_next.py

Output  next_rows.npz
  dist_*   Distortion.filter (slam/preprocessing.py:148-191): float64 / float32 timestamps and poses,
           constant timestamps, a large rotation, and the inactive paths
  chain_*  the shipped grid_sample.yaml chain distortion -> grid_sample -> to_tensor
  vox_*    Voxelization.filter (slam/preprocessing.py:71-97) = voxelise + voxel_hashing +
           voxel_normal_distribution (slam/common/pointcloud.py:83-167), float32 and float64 clouds
  p2p_*    GaussNewtonPointToPointAlignment.align (slam/odometry/alignment.py:144-189), one step per
           weighting scheme, a non-zero initial estimate, several iterations, float64
  proc_*   weighted_procrustes, numpy path (slam/common/registration.py:15-76)
"""
import importlib
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
pre = importlib.import_module("slam.preprocessing")
reg = importlib.import_module("slam.common.registration")
SCHEMES = ["default", "huber", "exp", "neighborhood", "geman_mcclure", "square_geman_mcclure", "cauchy"]


def timestamps_for(points: np.ndarray, seed: int) -> np.ndarray:
    """Acquisition time of each point of a spinning LiDAR: the azimuth fraction of the revolution plus jitter."""
    az = np.arctan2(points[:, 1].astype(np.float64), points[:, 0].astype(np.float64))
    rng = np.random.RandomState(seed)
    return 1.6e9 + 0.1 * ((az + np.pi) / (2 * np.pi)) + rng.uniform(0, 1e-4, points.shape[0])


def rot(axis, angle):
    from scipy.spatial.transform import Rotation
    a = np.asarray(axis, np.float64)
    return Rotation.from_rotvec(a / np.linalg.norm(a) * angle).as_matrix()


def main():
    out = {}
    H, W = 16, 256
    # ------------------------------------------------------------------ Distortion
    pc = syn.scan(2, H, W)
    ts = timestamps_for(pc, 5)
    rpose = syn.gt_relative_pose(2)
    big = np.eye(4)
    big[:3, :3] = rot([0.3, -0.5, 0.8], 2.6)
    big[:3, 3] = [1.5, -0.7, 0.2]
    cases = {
        "f64": (pc, ts, rpose),
        "f32": (pc, (ts - 1.6e9).astype(np.float32), rpose.astype(np.float32)),
        "mixed": (pc, (ts - 1.6e9).astype(np.float32), rpose),
        "const": (pc, np.full(pc.shape[0], 3.25), rpose),
        "big": (pc, ts, big),
        "pc64": (pc.astype(np.float64) * 1.0000001, ts, rpose.astype(np.float32)),
    }
    flt = pre.Distortion(pre.DistortionConfig(output_key="distorted"))
    for name, (p, t, T) in cases.items():
        dd = {"numpy_pc": p, "numpy_pc_timestamps": t, "init_rpose": T}
        flt.filter(dd)
        out[f"dist_{name}_pc"] = p
        out[f"dist_{name}_ts"] = t
        out[f"dist_{name}_pose"] = T
        out[f"dist_{name}_out"] = dd["distorted"]
        print("distortion", name, dd["distorted"].dtype, dd["distorted"].shape)
    # inactive paths return the input array itself
    dd = {"numpy_pc": pc, "init_rpose": rpose}
    flt.filter(dd)
    assert dd["distorted"] is pc
    dd = {"numpy_pc": pc, "numpy_pc_timestamps": ts, "init_rpose": None}
    flt.filter(dd)
    assert dd["distorted"] is pc

    # ------------------------------------------------------------------ the shipped chain
    Hc, Wc = 32, 512
    pcc = syn.scan(4, Hc, Wc)
    tsc = timestamps_for(pcc, 9)
    chain = [pre.Distortion(pre.DistortionConfig(output_key="distorted")),
             pre.GridSample(pre.GridSampleConfig(voxel_size=0.3, pointcloud_key="distorted")),
             pre.ToTensor(pre.ToTensorConfig(keys=dict(sample_points="input_data")))]
    dd = {"numpy_pc": pcc, "numpy_pc_timestamps": tsc, "init_rpose": syn.gt_relative_pose(4)}
    for f in chain:
        f.filter(dd)
    out.update(chain_pc=pcc, chain_ts=tsc, chain_pose=syn.gt_relative_pose(4), chain_distorted=dd["distorted"],
               chain_sample=dd["sample_points"], chain_indices=dd["sample_indices"])
    print("chain", dd["sample_points"].dtype, dd["sample_points"].shape, dd["input_data"].dtype)

    # ------------------------------------------------------------------ Voxelization
    rng = np.random.RandomState(11)
    v32 = (rng.randn(6000, 3) * np.array([8.0, 8.0, 1.0])).astype(np.float32)
    v64 = rng.randn(3000, 3) * 0.5
    scan_pts = syn.scan(1, H, W)
    for name, p, vs in (("f32", v32, 0.2), ("f64", v64, 0.1), ("scan", scan_pts, 0.5), ("one", v32[:1], 0.2),
                        ("coarse", v32, 50.0)):
        vox = pre.Voxelization(pre.VoxelizationConfig(voxel_size=vs))
        dd = {"numpy_pc": p}
        vox.filter(dd)
        out[f"vox_{name}_pc"] = p
        out[f"vox_{name}_voxel"] = np.float64(vs)
        for key in ("voxel_hashes", "voxel_coordinates", "voxel_means", "voxel_covariances", "voxel_sizes", "voxel_indices"):
            out[f"vox_{name}_{key}"] = np.asarray(dd[key])
        print("voxelization", name, {k: (np.asarray(dd[k]).dtype, np.asarray(dd[k]).shape) for k in
                                     ("voxel_means", "voxel_covariances", "voxel_sizes", "voxel_indices")})

    # ------------------------------------------------------------------ point-to-point Gauss-Newton
    pose = ns.pose.Pose("euler")
    N = 2000
    torch.manual_seed(4)
    tgt = torch.randn(1, N, 3) * 10.0
    xs = torch.tensor([[0.05, -0.03, 0.02, 0.004, -0.003, 0.006]])
    ref = pose.apply_transformation(tgt, xs) + 0.01 * torch.randn(1, N, 3)
    out.update(p2p_tgt=tgt[0].numpy(), p2p_ref=ref[0].numpy())
    for sch in SCHEMES:
        al = ns.alignment.GaussNewtonPointToPointAlignment(
            ns.alignment.GNPointToPointConfig(gauss_newton_config=dict(scheme=sch, sigma=0.3, max_iters=1)), pose=pose)
        dT, x, loss = al.align(ref, tgt)
        out[f"p2p_{sch}_dT"] = dT[0].numpy()
        out[f"p2p_{sch}_x"] = x[0].numpy()
        out[f"p2p_{sch}_loss"] = loss[0].numpy()
    al = ns.alignment.GaussNewtonPointToPointAlignment(
        ns.alignment.GNPointToPointConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=4,
                                                                   norm_stop_criterion=1e-9)), pose=pose)
    x0 = torch.tensor([[0.02, 0.01, -0.01, 0.001, 0.002, -0.001]])
    dT, x, loss = al.align(ref, tgt, initial_estimate=x0)
    out.update(p2p_multi_x0=x0[0].numpy(), p2p_multi_dT=dT[0].numpy(), p2p_multi_x=x[0].numpy(), p2p_multi_loss=loss[0].numpy())
    # pose-matrix initial estimate
    al1 = ns.alignment.GaussNewtonPointToPointAlignment(
        ns.alignment.GNPointToPointConfig(gauss_newton_config=dict(scheme="huber", sigma=0.3, max_iters=1)), pose=pose)
    dT, x, loss = al1.align(ref, tgt, initial_estimate=pose.build_pose_matrix(x0))
    out.update(p2p_mat_x=x[0].numpy(), p2p_mat_loss=loss[0].numpy())
    # float64
    t64, r64 = tgt.to(torch.float64), ref.to(torch.float64)
    al64 = ns.alignment.GaussNewtonPointToPointAlignment(
        ns.alignment.GNPointToPointConfig(gauss_newton_config=dict(scheme="default", max_iters=6, norm_stop_criterion=1e-12)),
        pose=pose)
    dT, x, loss = al64.align(r64, t64)
    out.update(p2p_f64_x=x[0].numpy(), p2p_f64_loss=loss[0].numpy(), p2p_f64_dT=dT[0].numpy())
    print("p2p", out["p2p_default_x"], out["p2p_multi_x"], out["p2p_f64_x"])

    # ------------------------------------------------------------------ weighted Procrustes (numpy path)
    rng = np.random.RandomState(21)
    pt = rng.randn(500, 3) * np.array([5.0, 3.0, 1.0])
    Tgt = np.eye(4)
    Tgt[:3, :3] = rot([0.2, 0.9, -0.4], 1.1)
    Tgt[:3, 3] = [0.5, -1.0, 2.0]
    pr = pt @ Tgt[:3, :3].T + Tgt[:3, 3] + 0.01 * rng.randn(500, 3)
    w = rng.uniform(0.1, 1.0, (500, 1))
    out.update(proc_tgt=pt, proc_ref=pr, proc_w=w, proc_T=reg.weighted_procrustes(pt, pr),
               proc_T_w=reg.weighted_procrustes(pt, pr, w),
               proc_T_f32=reg.weighted_procrustes(pt.astype(np.float32), pr.astype(np.float32)))
    # reflection case: the optimal orthogonal map is improper, S = diag(1, 1, -1) applies
    prm = pr.copy()
    prm[:, 2] *= -1.0
    out.update(proc_ref_mirror=prm, proc_T_mirror=reg.weighted_procrustes(pt, prm))
    # planar cloud (smallest singular value ~ 0)
    pl = pt.copy()
    pl[:, 2] = 0.0
    plr = pl @ Tgt[:3, :3].T + Tgt[:3, 3]
    out.update(proc_planar_tgt=pl, proc_planar_ref=plr, proc_T_planar=reg.weighted_procrustes(pl, plr))
    print("procrustes", np.abs(out["proc_T"] - Tgt).max(), np.linalg.det(out["proc_T_mirror"][:3, :3]),
          np.abs(out["proc_T_planar"] - Tgt).max())

    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **out)
    print("next_rows.npz", os.path.getsize(os.path.join(HERE, "next_rows.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
