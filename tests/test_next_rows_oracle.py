"""oracle/next_rows_oracle.py against the golden vectors of the unmodified reference
(tests/golden/next_rows.npz, written by tests/golden/make_golden_next.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import icp_oracle as orc
from oracle import next_rows_oracle as nxt
from conftest import CHAIN_CASES, chain_timestamps, check_pose_sequence, check_voxel_stats, dist_tol

SCHEMES = ["default", "huber", "exp", "neighborhood", "geman_mcclure", "square_geman_mcclure", "cauchy"]
DIST_CASES = ["f64", "f32", "mixed", "const", "big", "pc64"]
VOX_CASES = ["f32", "f64", "scan", "one", "coarse"]
P2P_MULTI_F32_TOL = 2e-3


def test_distortion_matches_reference(golden_next):
    g = golden_next
    for name in DIST_CASES:
        out = nxt.distort(g[f"dist_{name}_pc"], g[f"dist_{name}_ts"], g[f"dist_{name}_pose"])
        ref = g[f"dist_{name}_out"]
        assert out.dtype == ref.dtype == np.float64 and out.shape == ref.shape
        assert np.abs(out - ref).max() <= dist_tol(g[f"dist_{name}_pose"]), (name, np.abs(out - ref).max())


def test_distortion_chain_matches_reference(golden_next):
    """distortion -> grid_sample -> to_tensor, the shipped grid_sample.yaml chain."""
    g = golden_next
    d = nxt.distort(g["chain_pc"], g["chain_ts"], g["chain_pose"])
    assert np.abs(d - g["chain_distorted"]).max() <= 1e-11
    s, i = orc.grid_sample(d, 0.3)
    assert np.array_equal(i, g["chain_indices"])
    assert np.abs(s - g["chain_sample"]).max() <= 1e-11


@pytest.mark.parametrize("name", VOX_CASES)
def test_voxelization_matches_reference(golden_next, name):
    g = golden_next
    out = nxt.voxelization(g[f"vox_{name}_pc"], float(g[f"vox_{name}_voxel"]))
    check_voxel_stats(out, g, name)


def test_voxelization_reference_property(golden_next):
    """The reference's own test (tests/test_pointcloud.py:7-25): every point is within the voxel diagonal of
    its voxel's mean."""
    rng = np.random.RandomState(0)
    pc = rng.randn(100000, 3)
    out = nxt.voxelization(pc, 0.1)
    assert out["voxel_coordinates"].dtype == np.int64
    diff = np.linalg.norm(pc - out["voxel_means"][out["voxel_indices"]], axis=-1).max()
    assert diff < 0.18


@pytest.mark.parametrize("scheme", SCHEMES)
def test_p2point_step_matches_reference(golden_next, scheme):
    g = golden_next
    tgt, ref = torch.from_numpy(g["p2p_tgt"]).unsqueeze(0), torch.from_numpy(g["p2p_ref"]).unsqueeze(0)
    dT, x, loss, _ = nxt.align_p2point(ref, tgt, scheme=scheme, sigma=0.3, max_iters=1)
    assert np.abs(x[0].numpy() - g[f"p2p_{scheme}_x"]).max() <= 2e-5 * max(1.0, np.abs(g[f"p2p_{scheme}_x"]).max())
    assert np.abs(dT[0].numpy() - g[f"p2p_{scheme}_dT"]).max() <= 2e-5
    rl = g[f"p2p_{scheme}_loss"]
    assert np.abs(loss[0].numpy() - rl).max() <= 1e-5 * max(1.0, np.abs(rl).max())


def test_p2point_initial_estimates_multi_iter_f64(golden_next):
    g = golden_next
    tgt, ref = torch.from_numpy(g["p2p_tgt"]).unsqueeze(0), torch.from_numpy(g["p2p_ref"]).unsqueeze(0)
    x0 = torch.from_numpy(g["p2p_multi_x0"]).unsqueeze(0)
    dT, x, loss, _ = nxt.align_p2point(ref, tgt, scheme="geman_mcclure", sigma=0.3, max_iters=4, norm_stop=1e-9, x0=x0)
    # four float32 iterations of the reference's r-scaled Jacobian (H = sum r^2 grad grad^T, float32 matmul):
    # every step overshoots ~1/r-fold and re-amplifies the rounding of the previous one -> 2e-3 after 4 steps
    # (the float64 run below pins the same code path to 1e-9)
    assert np.abs(x[0].numpy() - g["p2p_multi_x"]).max() <= P2P_MULTI_F32_TOL
    assert np.abs(loss[0].numpy() - g["p2p_multi_loss"]).max() <= P2P_MULTI_F32_TOL * max(1.0, np.abs(g["p2p_multi_loss"]).max())
    # a pose-matrix initial estimate goes through from_pose_matrix (alignment.py:177-178)
    x0m = orc.from_pose_matrix(orc.build_pose_matrix(x0))
    _, x, loss, _ = nxt.align_p2point(ref, tgt, scheme="huber", sigma=0.3, max_iters=1, x0=x0m)
    assert np.abs(x[0].numpy() - g["p2p_mat_x"]).max() <= 2e-5 * max(1.0, np.abs(g["p2p_mat_x"]).max())
    t64, r64 = tgt.to(torch.float64), ref.to(torch.float64)
    dT, x, loss, _ = nxt.align_p2point(r64, t64, scheme="default", max_iters=6, norm_stop=1e-12)
    assert np.abs(x[0].numpy() - g["p2p_f64_x"]).max() <= 1e-9
    assert np.abs(dT[0].numpy() - g["p2p_f64_dT"]).max() <= 1e-9
    assert np.abs(loss[0].numpy() - g["p2p_f64_loss"]).max() <= 1e-9


def test_procrustes_matches_reference(golden_next):
    g = golden_next
    pt, pr = g["proc_tgt"], g["proc_ref"]
    assert np.abs(nxt.weighted_procrustes(pt, pr) - g["proc_T"]).max() <= 1e-12
    assert np.abs(nxt.weighted_procrustes(pt, pr, g["proc_w"]) - g["proc_T_w"]).max() <= 1e-12
    assert np.abs(nxt.weighted_procrustes(pt.astype(np.float32), pr.astype(np.float32)) - g["proc_T_f32"]).max() <= 1e-6
    Tm = nxt.weighted_procrustes(pt, g["proc_ref_mirror"])
    assert np.abs(Tm - g["proc_T_mirror"]).max() <= 1e-12 and np.linalg.det(Tm[:3, :3]) > 0.999
    assert np.abs(nxt.weighted_procrustes(g["proc_planar_tgt"], g["proc_planar_ref"]) - g["proc_T_planar"]).max() <= 1e-9


@pytest.mark.parametrize("case", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_whole_shipped_chain_oracle_vs_reference(golden_chain, case):
    """Distortion -> GridSample(0.4) -> ToTensor -> ICP (kd map) against the reference's poses.  The de-skewed
    samples are float64, so this also pins the oracle's float64 projection (icp_odometry.py:331-352)."""
    from pylidar_slam_b200 import synthetic as syn
    name, key, iters, thr = case
    H, W = 32, 512
    algo = orc.ICPFrameToModelOracle(orc.ICPConfig(max_num_alignments=iters, threshold_delta_pose=thr, data_key=key,
                                                   local_map="kdtree", local_map_size=4, scheme="geman_mcclure", sigma=0.3),
                                     orc.Projector(H, W))
    prev, poses = None, []
    for k in range(7):
        pc = syn.scan(k, H, W)
        d = pc if prev is None else nxt.distort(pc, chain_timestamps(pc, k), prev)
        s, _ = orc.grid_sample(d, 0.4)
        assert s.dtype == (np.float32 if prev is None else np.float64)
        dd = {"init_rpose": prev, key: (torch.from_numpy(s) if key == "input_data" else s)}
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            poses.append(dd["odometry_pose"].copy())
            prev = dd["odometry_pose"].astype(np.float64)
    its = [len(l) for l in algo.losses]
    check_pose_sequence(np.stack(poses), its, golden_chain[f"{name}_poses"], golden_chain[f"{name}_losses"],
                        threshold_delta_pose=max(thr, 1e-12), name=name)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_training_loss_and_analytic_gradient_match_reference_autograd(golden_loss, scheme):
    """loss_modules.py:51-132: the oracle's analytic gradient against the reference's autograd (float32)."""
    g = golden_loss
    vm, nm, x = torch.from_numpy(g["vertex_map"]), torch.from_numpy(g["normal_map"]), g["pose_params"]
    mats = orc.build_pose_matrix(torch.from_numpy(x))
    loss, per_batch, grad = nxt.p2plane_training_loss(vm[:, 1], vm[:, 0], nm[:, 0], mats, orc.Projector(16, 256), scheme, 0.5)
    ref = float(g[f"{scheme}_loss"])
    assert abs(loss - ref) <= 2e-5 * abs(ref), (scheme, loss, ref)
    assert abs(float(g[f"{scheme}_loss_matrix"]) - ref) <= 1e-6 * abs(ref)
    gm = g[f"{scheme}_grad_matrix"].astype(np.float64)
    scale = np.abs(gm).max()
    assert np.abs(grad[:, :3] - gm[:, :3]).max() <= 2e-4 * scale, (scheme, np.abs(grad[:, :3] - gm[:, :3]).max() / scale)
    assert np.abs(gm[:, 3]).max() == 0.0
    gp = nxt.pose_matrix_grad_to_params(x, grad)
    rp = g[f"{scheme}_grad_params"].astype(np.float64)
    assert np.abs(gp - rp).max() <= 2e-4 * np.abs(rp).max(), (scheme, np.abs(gp - rp).max() / np.abs(rp).max())
