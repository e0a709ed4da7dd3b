"""Parity of the CUDA path (through the C ABI) for the rows either side of the hot path (SURVEY.md section 8f,
ranks 1-2): Distortion, Voxelization, point-to-point Gauss-Newton, weighted Procrustes -- against the goldens of
the unmodified reference (tests/golden/next_rows.npz) and, at the BASELINE scan sizes, against the CPU oracle.
Everything here needs a GPU (`-m gpu`).  Tolerances:
  * integer / index outputs (voxel coordinates, hashes, sizes, voxel ids, sample indices): bit-exact
  * de-skewed points: 1e-11 m for float64 poses, 1e-6 m for float32 poses (conftest.dist_tol explains)
  * voxel means / scatter matrices: the rounding of the reference's own float32 accumulation (conftest.check_voxel_stats)
  * Gauss-Newton steps: 2e-5 relative on x (float32), 1e-9 (float64); Procrustes 1e-9 (float64 data)
"""
import logging

import numpy as np
import pytest
import torch

from conftest import CHAIN_CASES, chain_timestamps, check_pose_sequence, check_voxel_stats, dist_tol

pytestmark = pytest.mark.gpu

SCHEMES = ["default", "huber", "exp", "neighborhood", "geman_mcclure", "square_geman_mcclure", "cauchy"]
DIST_CASES = ["f64", "f32", "mixed", "const", "big", "pc64"]
VOX_CASES = ["f32", "f64", "scan", "one", "coarse"]


@pytest.fixture(scope="module")
def b200():
    import pylidar_slam_b200 as p
    return p


@pytest.fixture(scope="module")
def nxt():
    from oracle import next_rows_oracle
    return next_rows_oracle


@pytest.fixture(scope="module")
def orc():
    from oracle import icp_oracle
    return icp_oracle


@pytest.fixture(scope="module")
def syn():
    from pylidar_slam_b200 import synthetic
    return synthetic


def _timestamps(points, seed):
    az = np.arctan2(points[:, 1].astype(np.float64), points[:, 0].astype(np.float64))
    return 1.6e9 + 0.1 * ((az + np.pi) / (2 * np.pi)) + np.random.RandomState(seed).uniform(0, 1e-4, points.shape[0])


# ------------------------------------------------------------------------------------------ Distortion
@pytest.mark.parametrize("name", DIST_CASES)
def test_distortion_golden(b200, golden_next, name):
    g = golden_next
    flt = b200.Distortion(b200.DistortionConfig(output_key="distorted"))
    dd = {"numpy_pc": g[f"dist_{name}_pc"], "numpy_pc_timestamps": g[f"dist_{name}_ts"], "init_rpose": g[f"dist_{name}_pose"]}
    flt.filter(dd)
    out, ref = dd["distorted"], g[f"dist_{name}_out"]
    assert isinstance(out, np.ndarray) and out.dtype == np.float64 and out.shape == ref.shape
    assert np.abs(out - ref).max() <= dist_tol(g[f"dist_{name}_pose"]), (name, np.abs(out - ref).max())


def test_distortion_inactive_paths_return_the_input(b200, golden_next):
    """preprocessing.py:160-166: no timestamps / no pose / activate=False -> the input array itself."""
    g = golden_next
    pc, ts, pose = g["dist_f64_pc"], g["dist_f64_ts"], g["dist_f64_pose"]
    flt = b200.Distortion(b200.DistortionConfig(output_key="distorted"))
    for dd in ({"numpy_pc": pc, "init_rpose": pose}, {"numpy_pc": pc, "numpy_pc_timestamps": ts, "init_rpose": None}):
        flt.filter(dd)
        assert dd["distorted"] is pc
    off = b200.Distortion(b200.DistortionConfig(output_key="distorted", activate=False))
    dd = {"numpy_pc": pc, "numpy_pc_timestamps": ts, "init_rpose": pose}
    off.filter(dd)
    assert dd["distorted"] is pc
    with pytest.raises(AssertionError):
        flt.filter({"numpy_pc": pc, "numpy_pc_timestamps": ts[:-1], "init_rpose": pose})
    with pytest.raises(AssertionError):
        flt.filter({"numpy_pc": torch.from_numpy(pc), "numpy_pc_timestamps": ts, "init_rpose": pose})


@pytest.mark.parametrize("H,W", [(64, 2048), (128, 4096)])
def test_distortion_full_size_vs_oracle_and_properties(b200, nxt, syn, H, W):
    pc = syn.scan(7, H, W)
    ts = _timestamps(pc, 3)
    pose = syn.gt_relative_pose(7)
    out = b200.distort_frame(pc, ts, pose)
    ref = nxt.distort(pc, ts, pose)
    assert np.abs(out - ref).max() <= 1e-11
    # size-independent properties: the earliest point is untouched, the latest is moved by the full pose,
    # and every point keeps its distance to the interpolated sensor origin alpha * t (a rotation about it)
    i0, i1 = int(np.argmin(ts)), int(np.argmax(ts))
    assert np.abs(out[i0] - pc[i0].astype(np.float64)).max() <= 1e-12
    assert np.abs(out[i1] - (pose[:3, :3] @ pc[i1].astype(np.float64) + pose[:3, 3])).max() <= 1e-11
    alpha = (ts - ts.min()) / (ts.max() - ts.min())
    r_in = np.linalg.norm(pc.astype(np.float64), axis=1)
    r_out = np.linalg.norm(out - alpha[:, None] * pose[:3, 3][None], axis=1)
    assert np.abs(r_in - r_out).max() <= 1e-10
    # device tensors in, device tensor out
    d_pc = torch.from_numpy(pc).cuda()
    out_d = b200.distort_frame(d_pc, torch.from_numpy(ts).cuda(), pose)
    assert out_d.device == d_pc.device and out_d.dtype == torch.float64
    assert np.array_equal(out_d.cpu().numpy(), out)


def test_distortion_nan_and_single_point(b200, golden_next):
    g = golden_next
    pc, ts, pose = g["dist_f64_pc"], g["dist_f64_ts"].copy(), g["dist_f64_pose"]
    ts[5] = np.nan
    assert np.isnan(b200.distort_frame(pc, ts, pose)).all()  # np.max / np.min propagate NaN (preprocessing.py:180)
    one = b200.distort_frame(pc[:1], g["dist_f64_ts"][:1], pose)  # max == min -> alpha = 0
    assert np.array_equal(one, pc[:1].astype(np.float64))


def test_shipped_chain_distortion_grid_sample_to_tensor(b200, golden_next):
    """config/slam/preprocessing/grid_sample.yaml, filters 1-3, through Preprocessing."""
    g = golden_next
    pre = b200.Preprocessing(b200.PreprocessingConfig(filters={
        "1": dict(filter_name="distortion", force=False, activate=True, pointcloud_key="numpy_pc",
                  timestamps_key="numpy_pc_timestamps", output_key="distorted"),
        "2": dict(filter_name="grid_sample", voxel_size=0.3, pointcloud_key="distorted"),
        "3": dict(filter_name="to_tensor", keys=dict(sample_points="input_data"))}))
    dd = {"numpy_pc": g["chain_pc"], "numpy_pc_timestamps": g["chain_ts"], "init_rpose": g["chain_pose"]}
    pre.forward(dd)
    assert np.abs(dd["distorted"] - g["chain_distorted"]).max() <= 1e-11
    np.testing.assert_array_equal(dd["sample_indices"], g["chain_indices"])
    assert np.abs(dd["sample_points"] - g["chain_sample"]).max() <= 1e-11
    assert dd["input_data"].dtype == torch.float64 and tuple(dd["input_data"].shape) == g["chain_sample"].shape


@pytest.mark.parametrize("case", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_whole_shipped_chain_vs_reference_golden(b200, syn, golden_chain, case):
    """The whole shipped pipeline -- Distortion -> GridSample(0.4) -> ToTensor -> ICPFrameToModel (kd map) -- against
    the poses of the unmodified reference (tests/golden/chain_icp.npz).  The de-skewed samples are float64: the frame's
    vertex map is projected in float64 like the reference's (PLS_INPUT_TENSOR_F64 / _NDARRAY_F64).  Poses: 1e-4
    relative translation / 1e-5 rad, stop-rule aware for the yaml threshold, strict for the fixed-iteration runs."""
    name, key, iters, thr = case
    H, W = 32, 512
    cfg = b200.ICPFrameToModelConfig(
        local_map=b200.KdTreeLocalMapConfig(local_map_size=4),
        alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
        max_num_alignments=iters, threshold_delta_pose=thr, data_key=key)
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                device="cuda:0")
    algo.init()
    pre = b200.Preprocessing(b200.PreprocessingConfig(filters={
        "1": dict(filter_name="distortion", output_key="distorted"),
        "2": dict(filter_name="grid_sample", voxel_size=0.4, pointcloud_key="distorted"),
        "3": dict(filter_name="to_tensor", keys=dict(sample_points="input_data"))}))
    prev, poses, its = None, [], []
    for k in range(7):
        pc = syn.scan(k, H, W)
        dd = {"numpy_pc": pc, "numpy_pc_timestamps": chain_timestamps(pc, k), "init_rpose": prev}
        pre.forward(dd)
        if k >= 2:
            assert dd["sample_points"].dtype == np.float64 and dd["input_data"].dtype == torch.float64
        if key == "numpy_pc":
            dd["numpy_pc"] = dd["sample_points"]
        algo.process_next_frame(dd)
        if k == 0:
            assert "odometry_pose" not in dd
            continue
        assert dd["odometry_pose"].dtype == np.float32 and dd["odometry_pc"] is dd["distorted"]  # icp_odometry.py:204-205
        poses.append(dd["odometry_pose"].copy())
        its.append(int(algo.last_info[0]))
        prev = dd["odometry_pose"].astype(np.float64)
    check_pose_sequence(np.stack(poses), its, golden_chain[f"{name}_poses"], golden_chain[f"{name}_losses"],
                        threshold_delta_pose=max(thr, 1e-12), name=name)


def test_float64_cloud_projection_matches_oracle(b200, orc, syn):
    """A float64 `input_data` tensor / `numpy_pc` array keeps float64 up to the vertex map (icp_odometry.py:331-352):
    first-frame map initialisation from a float64 cloud equals the oracle's float64 projection rounded to float32."""
    H, W = 32, 512
    pc = syn.scan(2, H, W).astype(np.float64) * (1.0 + 1e-9)  # not representable in float32
    vm_ref = orc.Projector(H, W).build_projection_map(torch.from_numpy(pc).unsqueeze(0)).to(torch.float32)[0].numpy()
    for key, data in (("input_data", torch.from_numpy(pc)), ("numpy_pc", pc), ("input_data", torch.from_numpy(pc).cuda())):
        cfg = b200.ICPFrameToModelConfig(local_map=b200.KdTreeLocalMapConfig(local_map_size=4), max_num_alignments=4, data_key=key,
                                         alignment=b200.GaussNewtonPointToPlaneConfig())
        algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                    device="cuda:0")
        algo.init()
        algo.process_next_frame({key: data})
        # frame 0 inserts the pixels of the vertex map with |p| > 0.01 (local_map.py:320-328), row-major
        pts = vm_ref.reshape(3, -1).T
        pts = pts[np.linalg.norm(pts, axis=1) > 0.01]
        from pylidar_slam_b200 import _lib
        import ctypes as C
        m = C.c_int64(0)
        algo.ctx.call("pls_kdmap_size", C.byref(m))
        got = np.zeros((m.value, 3), np.float32)
        algo.ctx.call("pls_kdmap_points", _lib.ptr(got))
        assert got.shape == pts.shape, (key, got.shape, pts.shape)
        np.testing.assert_array_equal(got, pts)


# ------------------------------------------------------------------------------------------ Voxelization
@pytest.mark.parametrize("name", VOX_CASES)
def test_voxelization_golden(b200, golden_next, name):
    g = golden_next
    flt = b200.Voxelization(b200.VoxelizationConfig(voxel_size=float(g[f"vox_{name}_voxel"])))
    dd = {"numpy_pc": g[f"vox_{name}_pc"]}
    flt.filter(dd)
    check_voxel_stats(dd, g, name)


def test_voxelization_without_statistics_and_errors(b200, golden_next):
    g = golden_next
    flt = b200.Voxelization(b200.VoxelizationConfig(voxel_size=0.2, with_normal_distribution=False))
    dd = {"numpy_pc": g["vox_f32_pc"]}
    flt.filter(dd)
    np.testing.assert_array_equal(dd["voxel_hashes"], g["vox_f32_voxel_hashes"])
    np.testing.assert_array_equal(dd["voxel_coordinates"], g["vox_f32_voxel_coordinates"])
    assert "voxel_means" not in dd
    with pytest.raises(AssertionError):
        flt.filter({"other": g["vox_f32_pc"]})
    with pytest.raises(AssertionError):
        b200.Voxelization(b200.VoxelizationConfig()).filter({"numpy_pc": np.zeros((0, 3), np.float32)})


@pytest.mark.parametrize("n,voxel", [(131072, 0.2), (524288, 0.4), (2049, 0.05), (200000, 25.0)])
def test_voxelization_full_size_vs_oracle_and_properties(b200, nxt, syn, n, voxel):
    """BASELINE scan sizes, a ragged size around the sort tile and a few-huge-voxels case (warp-strided segments)."""
    if n in (131072, 524288):
        pts = syn.scan(5, 64 if n == 131072 else 128, 2048 if n == 131072 else 4096)
    else:
        pts = (np.random.RandomState(n).randn(n, 3) * np.array([30.0, 30.0, 3.0])).astype(np.float32)
    coords, hashes, sizes, means, covs, ids = b200.voxel_statistics(pts, voxel)
    ref = nxt.voxelization(pts, voxel)
    np.testing.assert_array_equal(coords, ref["voxel_coordinates"])
    np.testing.assert_array_equal(hashes, ref["voxel_hashes"])
    np.testing.assert_array_equal(sizes, ref["voxel_sizes"])
    np.testing.assert_array_equal(ids, ref["voxel_indices"])
    # both sides accumulate in float64 and round once to float32: equal up to one float32 ulp
    scale = np.abs(pts).max()
    assert np.abs(means.astype(np.float64) - ref["voxel_means"]).max() <= 2 * np.finfo(np.float32).eps * scale
    cmag = np.abs(ref["voxel_covariances"]).reshape(len(sizes), -1).max(axis=1)
    err = np.abs(covs.astype(np.float64) - ref["voxel_covariances"]).reshape(len(sizes), -1).max(axis=1)
    assert (err <= 4 * np.finfo(np.float32).eps * np.maximum(cmag, 1e-6)).all()
    # size-independent properties: counts add up; the mean of the means weighted by size is the cloud's mean;
    # every point is within the voxel diagonal of its voxel's mean (the reference's own test, test_pointcloud.py:7-25)
    assert sizes.sum() == n and ids.min() == 0 and ids.max() == len(sizes) - 1
    assert np.array_equal(np.bincount(ids, minlength=len(sizes)), sizes)
    wmean = (means.astype(np.float64) * sizes[:, None]).sum(0) / n
    assert np.abs(wmean - pts.astype(np.float64).mean(0)).max() <= 1e-4
    if voxel < 5.0:  # hash collisions between far-apart voxels only show up with few, huge voxels
        assert np.linalg.norm(pts - means[ids], axis=-1).max() < 0.9 * voxel * np.sqrt(3.0) + 1e-3


def test_voxelization_device_tensor(b200, nxt):
    pts = (np.random.RandomState(8).randn(50000, 3) * 5).astype(np.float32)
    d_pts = torch.from_numpy(pts).cuda()
    out = b200.voxel_statistics(d_pts, 0.3)
    assert all(o.device == d_pts.device for o in out)
    ref = nxt.voxelization(pts, 0.3)
    np.testing.assert_array_equal(out[2].cpu().numpy(), ref["voxel_sizes"])
    np.testing.assert_array_equal(out[5].cpu().numpy(), ref["voxel_indices"])


# ------------------------------------------------------------------------------------------ point-to-point GN
def _p2p(b200, scheme, max_iters=1, norm_stop=1e-3, sigma=0.3):
    return b200.RIGID_ALIGNMENT.load(dict(mode="point_to_point_gauss_newton",
                                          gauss_newton_config=dict(scheme=scheme, sigma=sigma, max_iters=max_iters,
                                                                   norm_stop_criterion=norm_stop)))


@pytest.mark.parametrize("scheme", SCHEMES)
def test_p2point_step_golden(b200, golden_next, scheme):
    g = golden_next
    tgt, ref = torch.from_numpy(g["p2p_tgt"]).unsqueeze(0), torch.from_numpy(g["p2p_ref"]).unsqueeze(0)
    dT, x, loss = _p2p(b200, scheme).align(ref, tgt)
    rx = g[f"p2p_{scheme}_x"]
    assert np.abs(x[0].numpy() - rx).max() <= 2e-5 * max(1.0, np.abs(rx).max())
    assert np.abs(dT[0].numpy() - g[f"p2p_{scheme}_dT"]).max() <= 2e-5
    rl = g[f"p2p_{scheme}_loss"]
    assert np.abs(loss[0].numpy() - rl).max() <= 1e-5 * max(1.0, np.abs(rl).max())


def test_p2point_initial_estimates_multi_iter_f64_and_device(b200, golden_next):
    g = golden_next
    tgt, ref = torch.from_numpy(g["p2p_tgt"]).unsqueeze(0), torch.from_numpy(g["p2p_ref"]).unsqueeze(0)
    x0 = torch.from_numpy(g["p2p_multi_x0"]).unsqueeze(0)
    dT, x, loss = _p2p(b200, "geman_mcclure", 4, 1e-9).align(ref, tgt, initial_estimate=x0)
    assert np.abs(x[0].numpy() - g["p2p_multi_x"]).max() <= 2e-3  # amplified float32 rounding, see test_next_rows_oracle.py
    dT, x, loss = _p2p(b200, "huber").align(ref, tgt, initial_estimate=b200.Pose("euler").build_pose_matrix(x0))
    assert np.abs(x[0].numpy() - g["p2p_mat_x"]).max() <= 2e-5 * max(1.0, np.abs(g["p2p_mat_x"]).max())
    dT, x, loss = _p2p(b200, "default", 6, 1e-12, sigma=0.5).align(ref.double(), tgt.double())
    assert x.dtype == torch.float64
    assert np.abs(x[0].numpy() - g["p2p_f64_x"]).max() <= 1e-9 and np.abs(dT[0].numpy() - g["p2p_f64_dT"]).max() <= 1e-9
    assert np.abs(loss[0].numpy() - g["p2p_f64_loss"]).max() <= 1e-9
    d_ref = ref.cuda()
    dTd, xd, lossd = _p2p(b200, "huber").align(d_ref, tgt.cuda())
    assert xd.device == d_ref.device and np.abs(xd[0].cpu().numpy() - g["p2p_huber_x"]).max() <= 2e-5 * max(1.0, np.abs(g["p2p_huber_x"]).max())


def test_p2point_large_vs_oracle_and_error_behaviour(b200, nxt, caplog):
    """cfg1-sized (10 000 correspondences) step against the oracle; the reference's error behaviour."""
    torch.manual_seed(2)
    N = 10000
    tgt = torch.randn(1, N, 3) * 10
    from oracle import icp_oracle as orc
    xs = torch.tensor([[0.04, -0.02, 0.03, 0.003, -0.002, 0.004]])
    ref = orc.apply_transformation(tgt, orc.build_pose_matrix(xs)) + 0.01 * torch.randn(1, N, 3)
    dT, x, loss = _p2p(b200, "cauchy").align(ref, tgt)
    dTo, xo, losso, _ = nxt.align_p2point(ref, tgt, scheme="cauchy", sigma=0.3, max_iters=1)
    assert np.abs(x.numpy() - xo.numpy()).max() <= 5e-5 * max(1.0, np.abs(xo.numpy()).max())
    assert np.abs(loss.numpy() - losso.numpy()).max() <= 1e-5 * max(1.0, np.abs(losso.numpy()).max())
    # identical clouds: |r| < 1e-7 -> warning, x unchanged (optimization.py:323-327)
    with caplog.at_level(logging.WARNING):
        dT, x, loss = _p2p(b200, "default").align(tgt, tgt)
    assert np.abs(x.numpy()).max() == 0.0 and any("residual norm" in r.message for r in caplog.records)
    # a single correspondence: rank-1 normal equations -> RuntimeError (optimization.py:334-336)
    with pytest.raises(RuntimeError, match="Invalid Jacobian"):
        _p2p(b200, "default").align(ref[:, :1], tgt[:, :1])
    with pytest.raises(AssertionError):
        _p2p(b200, "default").align(ref[:, :10], tgt[:, :11])
    with pytest.raises(AssertionError):  # the reference's ICP loop cannot drive this alignment either
        b200.ICPFrameToModel(b200.ICPFrameToModelConfig(alignment=b200.GNPointToPointConfig(mode="point_to_point_gauss_newton")),
                             projector=b200.SphericalProjector(height=16, width=64, up_fov=3.0, down_fov=-24.0), device="cuda:0")


# ------------------------------------------------------------------------------------------ Procrustes
def test_procrustes_golden(b200, golden_next):
    g = golden_next
    pt, pr = g["proc_tgt"], g["proc_ref"]
    assert np.abs(b200.weighted_procrustes(pt, pr) - g["proc_T"]).max() <= 1e-9
    assert np.abs(b200.weighted_procrustes(pt, pr, g["proc_w"]) - g["proc_T_w"]).max() <= 1e-9
    assert np.abs(b200.weighted_procrustes(pt.astype(np.float32), pr.astype(np.float32)) - g["proc_T_f32"]).max() <= 2e-6
    Tm = b200.weighted_procrustes(pt, g["proc_ref_mirror"])
    assert np.abs(Tm - g["proc_T_mirror"]).max() <= 1e-9 and np.linalg.det(Tm[:3, :3]) > 0.999
    assert np.abs(b200.weighted_procrustes(g["proc_planar_tgt"], g["proc_planar_ref"]) - g["proc_T_planar"]).max() <= 1e-9


def test_procrustes_scan_size_round_trip(b200, nxt, syn):
    """A 64x2048 scan against its rigidly moved copy: recovers the motion; matches the oracle; device tensors."""
    pc = syn.scan(3, 64, 2048).astype(np.float64)
    T = syn.gt_relative_pose(3)
    moved = pc @ T[:3, :3].T + T[:3, 3]
    est = b200.weighted_procrustes(pc, moved)
    assert np.abs(est - T).max() <= 1e-9
    assert np.abs(est - nxt.weighted_procrustes(pc, moved)).max() <= 1e-9
    est_d = b200.weighted_procrustes(torch.from_numpy(pc).cuda(), torch.from_numpy(moved).cuda())
    assert np.abs(est_d - est).max() <= 1e-12
    with pytest.raises(AssertionError):
        b200.weighted_procrustes(pc, moved[:-1])


# ------------------------------------------------------------------------------------------ training loss (rank 3)
def _loss_module(b200, scheme, H, W, sigma=0.5):
    return b200._PointToPlaneLossModule(b200.PointToPlaneLossConfig(least_square_scheme=dict(scheme=scheme, sigma=sigma)),
                                        b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0), b200.Pose("euler"))


@pytest.mark.parametrize("scheme", SCHEMES)
def test_training_loss_and_gradients_golden(b200, golden_loss, scheme):
    """_PointToPlaneLossModule.forward + backward (loss_modules.py:51-132) against the loss and the autograd gradients of
    the unmodified reference: pose parameters [B,6] and pose matrices [B,4,4].  float32 sums of 4096 pixels on the
    reference side: 2e-5 relative on the loss, 2e-4 of the largest component on the gradients."""
    g = golden_loss
    vm, nm = torch.from_numpy(g["vertex_map"]).cuda(), torch.from_numpy(g["normal_map"]).cuda()
    _, _, _, H, W = vm.shape
    mod = _loss_module(b200, scheme, H, W)
    x = torch.from_numpy(g["pose_params"]).cuda().requires_grad_(True)
    loss, dd = mod({"vertex_map": vm, "normal_map": nm, "pose_params": x})
    assert loss.dim() == 0 and loss.device == vm.device
    loss.backward()
    ref = float(g[f"{scheme}_loss"])
    assert abs(float(loss.detach()) - ref) <= 2e-5 * abs(ref), (scheme, float(loss.detach()), ref)
    rp = g[f"{scheme}_grad_params"]
    assert np.abs(x.grad.cpu().numpy() - rp).max() <= 2e-4 * np.abs(rp).max()
    M = b200.Pose("euler").build_pose_matrix(torch.from_numpy(g["pose_params"]).cuda()).requires_grad_(True)
    loss_m, _ = mod({"vertex_map": vm, "normal_map": nm, "pose_params": M})
    (3.0 * loss_m).backward()  # the incoming gradient scales the stored one
    assert abs(float(loss_m.detach()) - ref) <= 2e-5 * abs(ref)
    rm = g[f"{scheme}_grad_matrix"]
    assert np.abs(M.grad.cpu().numpy() / 3.0 - rm).max() <= 2e-4 * np.abs(rm).max()
    assert float(M.grad[:, 3].abs().max()) == 0.0


def test_training_loss_computes_missing_normal_maps_and_checks_shapes(b200, golden_loss):
    g = golden_loss
    vm = torch.from_numpy(g["vertex_map"]).cuda()
    mod = _loss_module(b200, "geman_mcclure", 16, 256)
    dd = {"vertex_map": vm, "pose_params": torch.from_numpy(g["pose_params"]).cuda()}
    loss, dd = mod(dd)
    assert tuple(dd["normal_map"].shape) == tuple(vm.shape)
    ref = float(g["geman_mcclure_loss"])
    assert abs(float(loss.detach()) - ref) <= 1e-4 * abs(ref)  # normal maps recomputed by the K3 kernel
    with pytest.raises(AssertionError):
        mod({"vertex_map": vm[:, :1], "normal_map": dd["normal_map"][:, :1], "pose_params": dd["pose_params"]})
    with pytest.raises(AssertionError):
        mod.point_to_plane_loss(vm[:, 1], vm[:, 0], dd["normal_map"][:, 0], dd["pose_params"][:2])
    with pytest.raises(AssertionError):
        mod.point_to_plane_loss(vm[:, 1].cpu(), vm[:, 0].cpu(), dd["normal_map"][:, 0].cpu(), dd["pose_params"].cpu())


def test_training_loss_scan_size_vs_oracle(b200, nxt, orc, syn):
    """BASELINE scan size (64x2048), batch of 2, against the oracle's analytic loss and gradient."""
    H, W, B = 64, 2048, 2
    pairs, params = [], []
    for b in range(B):
        k = 3 * b + 1
        pairs.append(np.stack([syn.vertex_map_from_scan(syn.scan(k, H, W), H, W)[0], syn.vertex_map_from_scan(syn.scan(k + 1, H, W), H, W)[0]]))
        T = torch.from_numpy(syn.gt_relative_pose(k + 1).astype(np.float32)).unsqueeze(0)
        params.append(orc.from_pose_matrix(T)[0].numpy() + np.array([0.03, -0.02, 0.01, 0.001, -0.001, 0.002], np.float32))
    vm = torch.from_numpy(np.stack(pairs).astype(np.float32))
    x = torch.from_numpy(np.stack(params).astype(np.float32))
    nm = b200.compute_normal_map(vm.reshape(B * 2, 3, H, W).cuda()).reshape(B, 2, 3, H, W)
    for scheme in ("geman_mcclure", "neighborhood"):
        mod = _loss_module(b200, scheme, H, W)
        xg = x.clone().cuda().requires_grad_(True)
        loss, _ = mod({"vertex_map": vm.cuda(), "normal_map": nm, "pose_params": xg})
        loss.backward()
        lo, per_batch, gm = nxt.p2plane_training_loss(vm[:, 1], vm[:, 0], nm[:, 0].cpu(), orc.build_pose_matrix(x), orc.Projector(H, W),
                                                      scheme, 0.5)
        assert abs(float(loss.detach()) - lo) <= 1e-5 * abs(lo), (scheme, float(loss.detach()), lo)
        assert np.abs(mod.last_loss_per_batch.cpu().numpy() - per_batch).max() <= 1e-5 * np.abs(per_batch).max()
        gp = nxt.pose_matrix_grad_to_params(x.numpy(), gm)
        assert np.abs(xg.grad.cpu().numpy() - gp).max() <= 1e-4 * np.abs(gp).max(), scheme
