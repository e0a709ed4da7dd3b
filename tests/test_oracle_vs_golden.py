"""Pins the CPU oracle (oracle/icp_oracle.py) against golden vectors produced by the
UNMODIFIED reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import icp_oracle as orc
from pylidar_slam_b200 import synthetic as syn
from conftest import check_pose_sequence, pose_errors

SCHEMES = ["default", "huber", "exp", "neighborhood", "geman_mcclure", "square_geman_mcclure", "cauchy"]


def test_a1_voxel_hash_grid_sample(golden_helpers):
    g = golden_helpers
    coords = orc.voxel_coords(g["a1_points"], 0.3)
    assert coords.dtype == np.int64
    np.testing.assert_array_equal(coords, g["a1_coords"])
    np.testing.assert_array_equal(orc.voxel_hashes(coords), g["a1_hashes"])
    s, i = orc.grid_sample(g["a1_points"], 0.3)
    np.testing.assert_array_equal(i, g["a1_indices"])
    np.testing.assert_array_equal(s, g["a1_sample"])
    s, i = orc.grid_sample(g["a1_points64"], 0.1)
    np.testing.assert_array_equal(i, g["a1_indices64"])
    np.testing.assert_array_equal(s, g["a1_sample64"])


def test_a3_projection(golden_helpers):
    g = golden_helpers
    proj = orc.Projector(16, 256)
    pts = torch.from_numpy(g["a3_points"]).unsqueeze(0)
    row, col = proj.pixels(pts)
    ok = ~np.isnan(g["a3_pixels"][:, 0])
    np.testing.assert_array_equal(row[0].numpy()[ok], g["a3_pixels"][ok, 0])
    np.testing.assert_array_equal(col[0].numpy()[ok], g["a3_pixels"][ok, 1])
    vmap = proj.build_projection_map(pts)[0].numpy()
    np.testing.assert_array_equal(vmap, g["a3_vmap"])


def test_a4_normal_map(golden_helpers):
    g = golden_helpers
    vm = torch.from_numpy(g["a4_vmap"]).unsqueeze(0)
    for k, key in ((5, "a4_nmap"), (3, "a4_nmap_k3")):
        n = orc.normal_map(vm, k)[0].numpy()
        np.testing.assert_allclose(n, g[key], atol=1e-6)


def test_a6_compute_neighbors(golden_helpers):
    g = golden_helpers
    nb, nf = orc.compute_neighbors(torch.from_numpy(g["a6_tgt"]).unsqueeze(0), torch.from_numpy(g["a6_ref"]),
                                   torch.from_numpy(g["a6_fields"]))
    np.testing.assert_array_equal(nb[0].numpy(), g["a6_nb"])
    np.testing.assert_array_equal(nf[0].numpy(), g["a6_nf"])


def test_reference_test_geometry_property():
    """tests/test_geometry.py:6-24 of the reference, seeded."""
    torch.manual_seed(0)
    tgt, ref = torch.randn(1, 3, 10, 10), torch.randn(10, 3, 10, 10)
    tgt[0, :, 0, 0] = 0.0
    nb, _ = orc.compute_neighbors(tgt, ref)
    assert nb[0, :, 0, 0].norm() == 0.0
    d_nb = (nb - tgt).norm(dim=1)[0]
    d_all = (ref - tgt).norm(dim=1)
    mask = torch.ones(10, 10, dtype=torch.bool)
    mask[0, 0] = False
    assert bool(((d_nb.unsqueeze(0) <= d_all)[:, mask]).all())


def test_a16_pose(golden_helpers):
    g = golden_helpers
    p = torch.from_numpy(g["a16_params"])
    m = orc.build_pose_matrix(p)
    np.testing.assert_allclose(m.numpy(), g["a16_mats"], atol=1e-7)
    np.testing.assert_allclose(orc.from_pose_matrix(torch.from_numpy(g["a16_mats"])).numpy(), g["a16_back"], atol=1e-6)


@pytest.mark.parametrize("scheme", SCHEMES)
def test_a11_a15_gauss_newton_step(golden_helpers, scheme):
    g = golden_helpers
    tgt, ref, nrm = (torch.from_numpy(g[k]).unsqueeze(0) for k in ("gn_tgt", "gn_ref", "gn_nrm"))
    dT, delta, loss = orc.align_p2plane(ref, tgt, nrm, scheme, 0.3, 1)
    np.testing.assert_allclose(delta[0].numpy(), g[f"gn_{scheme}_delta"], rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(dT[0].numpy(), g[f"gn_{scheme}_dT"], atol=1e-6)
    np.testing.assert_allclose(loss[0].numpy(), g[f"gn_{scheme}_loss"], rtol=1e-4, atol=1e-9)


def test_gauss_newton_multi_iter_and_known_answer(golden_helpers):
    g = golden_helpers
    tgt, ref, nrm = (torch.from_numpy(g[k]).unsqueeze(0) for k in ("gn_tgt", "gn_ref", "gn_nrm"))
    x, loss, st = orc.gauss_newton_p2plane(ref, tgt, nrm, "geman_mcclure", 0.3, 5, 1e-9)
    np.testing.assert_allclose(x[0].numpy(), g["gn_multi_x"], rtol=1e-3, atol=1e-6)
    # tests/test_optimization.py:9-32 with scheme="default" (the shipped huber/1e-4 variant raises)
    t64, r64, n64 = (torch.from_numpy(g[k]) for k in ("ka_tgt", "ka_ref", "ka_nrm"))
    est, loss, st = orc.gauss_newton_p2plane(r64, t64, n64, "default", 0.5, 100, 1e-10)
    assert np.abs(est.numpy() - g["ka_params"]).max() <= 1e-7
    assert np.abs(est.numpy() - g["ka_est"]).max() <= 1e-9
    assert float(loss.abs().sum()) <= 1e-7


def test_singular_hessian_raises():
    """optimization.py:49-50,334-336: noise-free data + robust scheme -> weights collapse -> raise."""
    torch.manual_seed(0)
    tgt = torch.randn(1, 100, 3, dtype=torch.float64)
    nrm = torch.randn(1, 100, 3, dtype=torch.float64)
    nrm /= nrm.norm(dim=-1, keepdim=True)
    x = torch.tensor([[0.01, 0.01, 0.01, 0.001, 0.001, 0.001]], dtype=torch.float64)
    ref = orc.apply_transformation(tgt, orc.build_pose_matrix(x))
    with pytest.raises(orc.SingularHessian):
        orc.gauss_newton_p2plane(ref, tgt, nrm, "huber", 1e-4, 100, 1e-10)


def test_a7_a9_kd_local_map(golden_helpers):
    g = golden_helpers
    lm = orc.KdTreeLocalMap(local_map_size=2)
    lm.update(np.eye(4, dtype=np.float32), new_vertex_map=torch.from_numpy(g["kd_v0"]).unsqueeze(0))
    lm.update(g["kd_rel"], new_points=g["kd_pc1"])
    np.testing.assert_allclose(lm.points, g["kd_map"], atol=1e-6)
    nb, nrm, _ = lm.nearest_neighbor_search(g["kd_queries"])
    np.testing.assert_allclose(nb, g["kd_nb"], atol=1e-6)
    dots = np.abs((nrm * g["kd_normals"]).sum(-1))
    assert np.mean(dots > 1 - 1e-4) > 0.995  # sign-free; a few ill-conditioned patches may differ


def _drive(algo, frame_fn, n):
    prev, poses = None, []
    for k in range(n):
        dd = frame_fn(k)
        dd["init_rpose"] = prev
        algo.process_next_frame(dd)
        if "odometry_pose" in dd:
            poses.append(dd["odometry_pose"].copy())
            prev = dd["odometry_pose"].astype(np.float64)
    return np.stack(poses)


def _frames(layout, H, W, voxel):
    def fn(k):
        pc = syn.scan(k, H, W)
        if layout == "vertex_map":
            return {"vertex_map": torch.from_numpy(syn.vertex_map_from_scan(pc, H, W))}
        if voxel:
            pc, _ = orc.grid_sample(pc, voxel)
        return {"numpy_pc": pc} if layout == "ndarray" else {"input_data": torch.from_numpy(pc)}
    return fn


CASES = [("kd_ndarray", "kdtree", "ndarray", "numpy_pc", 0.4, 7, "geman_mcclure", 0.3, 8),
         ("kd_tensor", "kdtree", "tensor", "input_data", 0.4, 7, "geman_mcclure", 0.3, 8),
         ("kd_vmap", "kdtree", "vertex_map", "vertex_map", None, 4, "geman_mcclure", 0.3, 8),
         ("proj_vmap", "projective", "vertex_map", "vertex_map", None, 7, "geman_mcclure", 0.3, 8),
         ("proj_ndarray", "projective", "ndarray", "numpy_pc", None, 7, "geman_mcclure", 0.3, 8),
         ("kd_default", "kdtree", "ndarray", "numpy_pc", 0.4, 5, "default", 0.5, 6)]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_icp_small_end_to_end(golden_icp_small, case):
    name, lm, layout, key, voxel, nf, scheme, sigma, iters = case
    H, W = 32, 512
    cfg = orc.ICPConfig(max_num_alignments=iters, data_key=key, local_map=lm, local_map_size=4,
                        scheme=scheme, sigma=sigma)
    algo = orc.ICPFrameToModelOracle(cfg, orc.Projector(H, W))
    poses = _drive(algo, _frames(layout, H, W, voxel), nf)
    ref = golden_icp_small[f"{name}_poses"]
    assert poses.shape == ref.shape
    for T, Tr in zip(poses, ref):
        dt, ang = pose_errors(T, Tr)
        assert dt <= 1e-4 and ang <= 1e-5, (name, dt, ang)
    gl = golden_icp_small[f"{name}_losses"]
    for mine, theirs in zip(algo.losses, gl):
        theirs = theirs[~np.isnan(theirs)]
        assert len(mine) == len(theirs), name
        np.testing.assert_allclose(mine, theirs, rtol=2e-3)


def test_icp_cfg2_full_size_oracle_vs_reference(golden_icp_full):
    """The oracle at BASELINE config-2 size (first 12 frames of the golden run): strict tolerance with
    the fixed-iteration golden; stop-rule aware with the default threshold."""
    H, W = 64, 2048
    for name, thr, iters, nf in (("cfg2_tensor_fixed6", 0.0, 6, 9), ("cfg2_tensor", 1e-4, 10, 12)):
        cfg = orc.ICPConfig(max_num_alignments=iters, data_key="input_data", local_map="kdtree", local_map_size=20,
                            scheme="geman_mcclure", sigma=0.3, threshold_delta_pose=thr)
        algo = orc.ICPFrameToModelOracle(cfg, orc.Projector(H, W))
        poses = _drive(algo, _frames("tensor", H, W, 0.3), nf + 1)
        its = [len(l) for l in algo.losses]
        check_pose_sequence(poses, its, golden_icp_full[f"{name}_poses"][:nf], golden_icp_full[f"{name}_losses"][:nf],
                            threshold_delta_pose=max(thr, 1e-12), name=name)
