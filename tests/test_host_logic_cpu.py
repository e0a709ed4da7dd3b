"""Host-logic tests on a machine without a GPU.

The Python mirrors (filters, alignments, loss module, registries) marshal arguments into the C ABI and map status codes
onto the reference's error behaviour.  Here `_lib.Context` is replaced -- for the duration of one test, by a
monkeypatch fixture -- with tests/dryrun_next_rows.FakeContext, a TEST-ONLY stand-in that follows
include/plslam_b200.h argument by argument and computes with the CPU oracle.  The bodies of the GPU parity tests
(tests/test_next_rows_gpu.py) then run unchanged against the reference goldens: what is checked is the plumbing
(argument order, dtypes, shapes, output slicing, registries, exceptions, the autograd hand-off), not the kernels --
those are pinned by tests/test_host_math.py on CPU and by the `-m gpu` run on the B200."""
import numpy as np
import pytest
import torch

import dryrun_next_rows as dry
import test_next_rows_gpu as T
from oracle import icp_oracle as orc
from oracle import next_rows_oracle as nxt
from pylidar_slam_b200 import synthetic as syn


class _Caplog:
    def __init__(self, caplog):
        self._c = caplog

    def at_level(self, lvl):
        return self._c.at_level(lvl)

    @property
    def records(self):
        return self._c.records


@pytest.fixture
def b200(monkeypatch):
    import pylidar_slam_b200 as pkg
    from pylidar_slam_b200 import _lib, common, training
    monkeypatch.setattr(_lib, "Context", dry.FakeContext)
    monkeypatch.setattr(common, "_default_ctx", dry.FakeContext())
    monkeypatch.setattr(training, "_require_cuda", lambda t: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)  # device legs degrade to CPU tensors
    return pkg


@pytest.mark.parametrize("name", T.DIST_CASES)
def test_distortion_filter_plumbing(b200, golden_next, name):
    T.test_distortion_golden(b200, golden_next, name)


def test_distortion_inactive_paths_and_shape_errors(b200, golden_next):
    T.test_distortion_inactive_paths_return_the_input(b200, golden_next)
    T.test_distortion_nan_and_single_point(b200, golden_next)


def test_shipped_chain_through_the_preprocessing_registry(b200, golden_next):
    T.test_shipped_chain_distortion_grid_sample_to_tensor(b200, golden_next)


@pytest.mark.parametrize("name", T.VOX_CASES)
def test_voxelization_filter_plumbing(b200, golden_next, name):
    T.test_voxelization_golden(b200, golden_next, name)


def test_voxelization_without_statistics_and_errors(b200, golden_next):
    T.test_voxelization_without_statistics_and_errors(b200, golden_next)
    T.test_voxelization_full_size_vs_oracle_and_properties(b200, nxt, syn, 2049, 0.05)


@pytest.mark.parametrize("scheme", T.SCHEMES)
def test_point_to_point_alignment_plumbing(b200, golden_next, scheme):
    T.test_p2point_step_golden(b200, golden_next, scheme)


def test_point_to_point_initial_estimates_errors_and_registry(b200, golden_next, caplog):
    T.test_p2point_initial_estimates_multi_iter_f64_and_device(b200, golden_next)
    T.test_p2point_large_vs_oracle_and_error_behaviour(b200, nxt, _Caplog(caplog))


def test_procrustes_plumbing(b200, golden_next):
    T.test_procrustes_golden(b200, golden_next)


@pytest.mark.parametrize("scheme", T.SCHEMES)
def test_training_loss_module_and_autograd_hand_off(b200, golden_loss, scheme):
    T.test_training_loss_and_gradients_golden(b200, golden_loss, scheme)


KD_CASES = [("kd_ndarray", "ndarray", "numpy_pc", 0.4, 7), ("kd_tensor", "tensor", "input_data", 0.4, 7),
            ("kd_vmap", "vertex_map", "vertex_map", None, 4)]


@pytest.mark.parametrize("case", KD_CASES, ids=[c[0] for c in KD_CASES])
def test_icp_frame_to_model_mirror_host_logic(b200, golden_icp_small, case):
    """ICPFrameToModel mirror (icp_odometry.py:157-246): the three input layouts, frame 0 writes nothing, data_dict keys
    and dtypes, relative / absolute pose bookkeeping -- against the reference's poses (the arithmetic behind the fake C
    ABI is the oracle's, so the poses themselves restate tests/test_oracle_vs_golden.py)."""
    from conftest import pose_errors
    name, layout, key, voxel, nf = case
    H, W = 32, 512
    cfg = b200.ICPFrameToModelConfig(
        local_map=b200.KdTreeLocalMapConfig(local_map_size=4),
        alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=1)),
        max_num_alignments=8, data_key=key)
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0), device="cuda:0")
    algo.init()
    prev, poses = None, []
    for k in range(nf):
        pc = syn.scan(k, H, W)
        if layout == "vertex_map":
            dd = {"vertex_map": torch.from_numpy(syn.vertex_map_from_scan(pc, H, W))}
        else:
            s, _ = orc.grid_sample(pc, voxel)
            dd = {"numpy_pc": s} if layout == "ndarray" else {"input_data": torch.from_numpy(s)}
        dd["init_rpose"] = prev
        algo.process_next_frame(dd)
        if k == 0:
            assert "odometry_pose" not in dd and "odometry_pc" not in dd
            continue
        assert dd["odometry_pose"].dtype == np.float32 and dd["odometry_pose"].shape == (4, 4)
        assert dd["odometry_pc"].dtype == np.float32 and dd["odometry_pc"].shape[1] == 3
        poses.append(dd["odometry_pose"].copy())
        prev = dd["odometry_pose"].astype(np.float64)
    ref = golden_icp_small[f"{name}_poses"]
    assert len(poses) == len(ref)
    for T_, Tr in zip(poses, ref):
        dt, ang = pose_errors(T_, Tr)
        assert dt <= 1e-4 and ang <= 1e-5, (name, dt, ang)
    rel = algo.get_relative_poses()
    assert rel.shape == (nf, 4, 4) and np.array_equal(rel[0], np.eye(4, dtype=np.float32))
    assert len(algo.absolute_poses) == nf
    absolute = np.eye(4)
    for T_ in poses:
        absolute = absolute @ T_.astype(np.float64)
    assert np.abs(algo.absolute_poses[-1] - absolute).max() <= 1e-4  # Euler re-normalised float64 product
    with pytest.raises(AssertionError):
        algo.process_next_frame({"other_key": pc})
    with pytest.raises(AssertionError):
        algo.process_next_frame({key: np.zeros((5, 4), np.float32) if layout != "vertex_map" else torch.zeros(1, 3, 8, 8)})


def test_icp_mirror_float64_layouts_are_announced(b200):
    """float64 clouds must reach the C ABI as the _F64 layouts (float64 projection), float32 ones as before."""
    from pylidar_slam_b200 import _lib
    cfg = b200.ICPFrameToModelConfig(local_map=b200.KdTreeLocalMapConfig(), data_key="input_data",
                                     alignment=b200.GaussNewtonPointToPlaneConfig())
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=16, width=64, up_fov=3.0, down_fov=-24.0), device="cuda:0")
    pts = np.random.RandomState(0).randn(10, 3)
    # point layouts carry a residency hint in the high bits (host memory here)
    assert algo._interpret(pts)[0] == _lib.INPUT_NDARRAY_F64 | _lib.PTR_HOST and algo._interpret(pts)[1].dtype == np.float64
    assert algo._interpret(pts.astype(np.float32))[0] == _lib.INPUT_NDARRAY | _lib.PTR_HOST
    assert algo._interpret(torch.from_numpy(pts))[0] == _lib.INPUT_TENSOR_F64 | _lib.PTR_HOST
    assert algo._interpret(torch.from_numpy(pts).float())[0] == _lib.INPUT_TENSOR | _lib.PTR_HOST
    assert algo._interpret(torch.zeros(1, 3, 16, 64, dtype=torch.float64))[0] == _lib.INPUT_VERTEX_MAP
    assert algo._interpret(torch.zeros(1, 3, 16, 64, dtype=torch.float64))[1].dtype == torch.float32


def test_registries_expose_the_reference_names(b200):
    assert set(b200.FILTER.__members__) == {"distortion", "voxelization", "grid_sample", "to_tensor"}
    assert set(b200.RIGID_ALIGNMENT.__members__) == {"point_to_plane_gauss_newton", "point_to_point_gauss_newton"}
    assert set(b200.LOCAL_MAP.__members__) == {"kdtree_local_map", "projective_local_map"} or len(b200.LOCAL_MAP.__members__) == 2
    assert "icp_F2M" in b200.ODOMETRY.__members__
    with pytest.raises(AssertionError):
        b200.FILTER.load(dict(filter_name="ground_detection"))
    with pytest.raises(AssertionError):
        b200.RIGID_ALIGNMENT.load(dict(mode="nope"))
    f = b200.FILTER.load(dict(filter_name="voxelization", voxel_size=0.4))
    assert isinstance(f, b200.Voxelization) and f.config.voxel_size == 0.4 and f.config.input_channel == "numpy_pc"


def test_the_fake_backend_does_not_leak():
    """The stand-in lives only inside the fixture: outside it the product still refuses to run without a GPU."""
    from pylidar_slam_b200 import _lib
    assert _lib.Context is not dry.FakeContext
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            _lib.Context()


def test_handoff_twin_is_matched_by_identity_and_dropped_after_a_write(b200):
    """GridSample hands out a host array and publishes its device-resident twin; the odometry may substitute the twin
    only for that very array with unchanged content (odometry.do_process_next_frame)."""
    from pylidar_slam_b200 import _lib
    pts = syn.scan(0, 16, 256).astype(np.float32)
    ctx = dry.FakeContext()
    samples, _ = b200.grid_sample(pts, 0.3, ctx=ctx)
    H = _lib.Handoff
    assert H.array is samples and H.rows == samples.shape[0] and H.dev_ptr
    addr = _lib.ptr(torch.from_numpy(samples))            # what ToTensor produces shares the address
    dev = int(ctx.cfg.device)
    assert H.match(addr, samples.shape[0], False, dev) == H.dev_ptr
    assert H.match(addr, samples.shape[0], True, dev) == 0                    # other dtype
    assert H.match(addr, samples.shape[0] - 1, False, dev) == 0               # a slice is another cloud
    assert H.match(_lib.ptr(samples.copy()), samples.shape[0], False, dev) == 0   # a copy lives elsewhere
    assert H.match(addr, samples.shape[0], False, dev + 1) == 0               # twin on another device
    samples[0, 0] += 1.0                                                       # caller edits the array: host content wins
    assert H.match(addr, samples.shape[0], False, dev) == 0
    H.clear()
    assert H.match(addr, samples.shape[0], False, dev) == 0


def test_pinned_pool_reuses_a_buffer_only_when_nothing_refers_to_it(monkeypatch):
    """_lib.PinnedPool: arrays handed out are views of pooled buffers; a buffer takes the next frame only after every
    view (numpy or torch) is gone; an exhausted pool answers None (the caller then copies into pageable arrays)."""
    import ctypes as C
    from pylidar_slam_b200 import _lib

    class Allocator:            # stands in for pls_pinned_alloc / pls_pinned_free (no CUDA on this box)
        def __init__(self):
            self.live, self.freed = {}, []

        def pls_pinned_alloc(self, n, ref):
            block = (C.c_char * n)()
            self.live[C.addressof(block)] = block
            ref._obj.value = C.addressof(block)
            return 0

        def pls_pinned_free(self, p):
            self.freed.append(p)
            self.live.pop(p, None)
            return 0

    alloc = Allocator()
    monkeypatch.setattr(_lib, "load", lambda: alloc)
    monkeypatch.setattr(_lib.PinnedPool, "buffers", [])
    monkeypatch.setattr(_lib.PinnedPool, "limit", 3)
    take = _lib.PinnedPool.take
    a = take(120)
    view = a[:48].view(np.float32).reshape(-1, 3)
    b = take(120)
    assert a is not b and len(_lib.PinnedPool.buffers) == 2
    del a, b
    assert take(120) is _lib.PinnedPool.buffers[1]          # buffer 0 is still referenced through `view`
    tensor = torch.from_numpy(view)
    del view
    assert take(64) is _lib.PinnedPool.buffers[1]           # ... and now through the tensor
    del tensor
    assert take(64) is _lib.PinnedPool.buffers[0]
    big = take(4096)                                         # too small buffers do not qualify: a third one
    assert big.nbytes == 4096 and len(_lib.PinnedPool.buffers) == 3
    held = [take(8), take(8)]
    assert held[0] is not held[1] and take(8) is None        # limit reached, everything referenced
    del held
    bigger = take(8192)                                      # limit reached: a free smaller buffer makes room
    assert bigger is not None and len(_lib.PinnedPool.buffers) == 3 and len(alloc.freed) == 1


def test_voxelise_with_three_voxel_lengths_and_alignment_mask_error(b200, golden_misc):
    """voxelise(voxel_x, voxel_y, voxel_z) reaches pls_voxel_hash_xyz with the three lengths; a `mask` handed to either
    alignment ends in the reference's own error type (tests/test_dropin_reference.py shows the reference raising it)."""
    g = golden_misc
    vx, vy, vz = (float(v) for v in g["vox_sizes"])
    np.testing.assert_array_equal(b200.voxelise(g["vox_points"], vx, vy, vz, ctx=dry.FakeContext()), g["vox_coords"])
    n = 64
    pts = np.random.RandomState(0).randn(1, n, 3).astype(np.float32)
    mask = np.ones((1, n, 1), np.float32)
    plane = b200.GaussNewtonPointToPlaneAlignment(b200.GaussNewtonPointToPlaneConfig(), ctx=dry.FakeContext())
    point = b200.GaussNewtonPointToPointAlignment(b200.GNPointToPointConfig(), ctx=dry.FakeContext())
    with pytest.raises(RuntimeError, match="broadcast"):
        plane.align(pts, pts, pts, mask=mask)
    with pytest.raises(RuntimeError, match="broadcast"):
        point.align(pts, pts, mask=torch.from_numpy(mask))
    with pytest.raises(AssertionError):                       # the reference's shape check comes first
        plane.align(pts, pts, pts, mask=mask[:, :, 0])


# ---- the fine-grained plug-ins (LOCAL_MAP / RIGID_ALIGNMENT mirrors) and the reference-shaped loop built from them
def test_local_map_and_alignment_plugins_plumbing(b200):
    """KdTreeLocalMap / ProjectiveLocalMap / GaussNewtonPointToPlaneAlignment mirrors: argument marshalling, output
    shapes for ndarray and tensor inputs, the NeighborhoodResult fields -- against the oracle's classes on the same data."""
    H, W = 16, 128
    proj = b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0)
    vm0 = torch.from_numpy(syn.vertex_map_from_scan(syn.scan(0, H, W), H, W))
    eye = torch.eye(4).unsqueeze(0)
    rel = b200.Pose("euler").build_pose_matrix(torch.tensor([[0.3, 0.0, 0.0, 0.0, 0.0, 0.01]]))

    kd = b200.KdTreeLocalMap(b200.KdTreeLocalMapConfig(local_map_size=3), projector=proj, ctx=dry.FakeContext(
        local_map_size=3, num_neighbors_normals=10))
    kd.init()
    kd.update(eye, new_vertex_map=vm0)
    ref = orc.KdTreeLocalMap(3, 10)
    ref.update(np.eye(4, dtype=np.float32), new_vertex_map=vm0)
    assert kd.num_points() == ref.points.shape[0] > 0
    pts1 = syn.scan(1, H, W)
    pts1 = pts1[np.linalg.norm(pts1, axis=1) > 0][::3].copy()
    kd.update(rel, new_pc_data=torch.from_numpy(pts1))
    ref.update(rel[0].numpy(), new_points=pts1)
    np.testing.assert_array_equal(kd.points(), ref.points)
    q = pts1[:50] + 0.01
    want_p, want_n, _ = ref.nearest_neighbor_search(q)
    for queries in (q, torch.from_numpy(q)):                    # ndarray in -> ndarrays out; tensor in -> [1,N,3] tensors
        res = kd.nearest_neighbor_search(queries)
        got_p, got_n, got_t = res.neighbor_points, res.neighbor_normals, res.new_target_points
        if isinstance(queries, torch.Tensor):
            assert got_p.shape == got_n.shape == got_t.shape == (1, 50, 3)
            got_p, got_n, got_t = got_p[0].numpy(), got_n[0].numpy(), got_t[0].numpy()
        np.testing.assert_array_equal(got_p, want_p)
        np.testing.assert_array_equal(got_n, want_n)
        np.testing.assert_array_equal(got_t, q)
    kd.update(rel)                                              # no new frame: the map only moves
    ref.update(rel[0].numpy())
    np.testing.assert_array_equal(kd.points(), ref.points)

    pm = b200.ProjectiveLocalMap(b200.ProjectiveLocalMapConfig(local_map_size=3), projector=proj, ctx=dry.FakeContext(
        local_map_size=3, normals_kernel_size=5, height=H, width=W, up_fov_deg=3.0, down_fov_deg=-24.0))
    pm.init()
    rp = orc.ProjectiveLocalMap(orc.Projector(H, W), 3, 5)
    vm1 = torch.from_numpy(syn.vertex_map_from_scan(syn.scan(1, H, W), H, W))
    pm.update(eye, new_vertex_map=vm0)
    rp.update(eye, new_vertex_map=vm0)
    pm.update(rel, new_vertex_map=vm1)
    rp.update(rel, new_vertex_map=vm1)
    v, n = pm.model()
    np.testing.assert_array_equal(v, rp.model_vmap.numpy())
    np.testing.assert_array_equal(n, rp.model_nmap.numpy())
    qs = torch.from_numpy(pts1)
    a, b, c = rp.nearest_neighbor_search(qs)
    res = pm.nearest_neighbor_search(qs)
    assert res.neighbor_points.shape == a.shape and a.shape[1] > 0
    np.testing.assert_array_equal(res.neighbor_points.numpy(), a.numpy())
    np.testing.assert_array_equal(res.neighbor_normals.numpy(), b.numpy())
    np.testing.assert_array_equal(res.new_target_points.numpy(), c.numpy())

    al = b200.GaussNewtonPointToPlaneAlignment(b200.GaussNewtonPointToPlaneConfig(
        gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=3, norm_stop_criterion=1e-9)), ctx=dry.FakeContext())
    dT, x, loss = al.align(a, c, b)
    x_ref, loss_ref, _ = orc.gauss_newton_p2plane(a, c, b, "geman_mcclure", 0.3, 3, 1e-9)
    assert tuple(dT.shape) == (1, 4, 4) and tuple(x.shape) == (1, 6) and tuple(loss.shape) == (1, a.shape[1])
    np.testing.assert_allclose(np.asarray(x), x_ref.numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.asarray(dT), orc.build_pose_matrix(x_ref).numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.asarray(loss), loss_ref.numpy(), rtol=0, atol=1e-7)


@pytest.mark.parametrize("name,lm,key", [("kd_gn3", "kdtree", "numpy_pc"), ("proj_gn2", "projective", "vertex_map")])
def test_fine_grained_icp_loop_host_logic_vs_reference_golden(b200, monkeypatch, name, lm, key):
    """gauss_newton_config.max_iters > 1 takes ICPFrameToModel._process_fine_grained: input normalisation, query sampling,
    the ICP loop, the key-frame policy and the pose bookkeeping run in Python over the plug-ins.  With the oracle behind the
    C ABI the poses must reproduce the unmodified reference's (tests/golden/icp_gn.npz, made by make_golden_gn.py)."""
    import os
    import test_gpu_parity as G
    dry.cuda_stand_ins(monkeypatch)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "icp_gn.npz"))
    H, W = 32, 512
    lmc = b200.KdTreeLocalMapConfig(local_map_size=4) if lm == "kdtree" else b200.ProjectiveLocalMapConfig(local_map_size=4)
    cfg = b200.ICPFrameToModelConfig(
        local_map=lmc, alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(
            scheme="geman_mcclure", sigma=0.3, max_iters=int(g[f"{name}_gn_iters"]), norm_stop_criterion=1e-9)),
        max_num_alignments=5, data_key=key, threshold_delta_pose=0.0)
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                pose=b200.Pose("euler"), device="cuda:0")
    assert algo._fine_grained
    algo.init()
    layout = "ndarray" if key == "numpy_pc" else "vertex_map"
    poses = G._drive(algo, G._frames(syn, b200.grid_sample, layout, H, W, 0.4 if layout == "ndarray" else None), 6)
    ref = g[f"{name}_poses"]
    assert len(poses) == len(ref) == len(algo.get_relative_poses()) - 1
    for k, (T, Tr) in enumerate(zip(poses, ref)):
        dt, ang = G.pose_errors(T, Tr)
        assert dt <= 1e-4 and ang <= 1e-5, (name, k, dt, ang)
    assert int(algo.last_info[0]) == 5 and len(algo.last_losses) == 5


def test_fine_grained_icp_loop_tensor_layout_stop_rule_no_keyframe_and_distorted_key(b200, monkeypatch):
    """The remaining branches of ICPFrameToModel._process_fine_grained against the oracle configured alike: a torch [N,3]
    input (pixel queries), the default stop rule breaking the loop, frames that do NOT become key frames (the map only
    moves, the motion accumulates until the thresholds are passed) and the `distorted` cloud passed through as odometry_pc."""
    dry.cuda_stand_ins(monkeypatch)
    H, W = 32, 512
    kw = dict(max_num_alignments=6, threshold_delta_pose=1e-4, threshold_trans=2.0, threshold_rot=30.0)   # a key frame every 3rd
    cfg = b200.ICPFrameToModelConfig(
        local_map=b200.KdTreeLocalMapConfig(local_map_size=3), alignment=b200.GaussNewtonPointToPlaneConfig(
            gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=2, norm_stop_criterion=1e-9)),
        data_key="input_data", **kw)
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                pose=b200.Pose("euler"), device="cuda:0")
    algo.init()
    ref = orc.ICPFrameToModelOracle(orc.ICPConfig(data_key="input_data", local_map="kdtree", local_map_size=3, scheme="geman_mcclure",
                                                  sigma=0.3, gn_max_iters=2, gn_norm_stop=1e-9, **kw), orc.Projector(H, W))
    prev_a = prev_b = None
    keyframes, iters = [], []
    for k in range(6):
        pc, _ = orc.grid_sample(syn.scan(k, H, W), 0.4)
        marker = np.full((5, 3), float(k), np.float32)                      # stands for the Distortion filter's output
        da = {"input_data": torch.from_numpy(pc.copy()), "init_rpose": prev_a, "distorted": marker}
        db = {"input_data": torch.from_numpy(pc.copy()), "init_rpose": prev_b, "distorted": marker}
        size_before = algo.local_map.num_points() if k else 0
        algo.process_next_frame(da)
        ref.process_next_frame(db)
        if k == 0:
            assert "odometry_pose" not in da and "odometry_pose" not in db
            continue
        np.testing.assert_allclose(da["odometry_pose"], db["odometry_pose"], rtol=0, atol=1e-6)
        assert da["odometry_pc"] is marker and da["odometry_pose"].dtype == np.float32
        assert len(algo.last_losses) == len(ref.losses[-1])
        iters.append(len(algo.last_losses))
        keyframes.append(algo.local_map.num_points() != size_before)
        assert algo.local_map.num_points() == ref.local_map.points.shape[0]
        prev_a, prev_b = da["odometry_pose"].astype(np.float64), db["odometry_pose"].astype(np.float64)
    assert not all(keyframes) and any(keyframes)                             # both branches of the key-frame policy ran
    assert min(iters) < 6                                                    # the stop rule broke at least one loop
    np.testing.assert_allclose(algo.get_relative_poses(), ref.get_relative_poses(), rtol=0, atol=1e-6)
    np.testing.assert_allclose(np.stack(algo.absolute_poses), np.stack(ref.absolute_poses), rtol=0, atol=1e-5)


def test_point_to_plane_initial_estimates_and_register_new_frame_plumbing(b200):
    """GaussNewtonPointToPlaneAlignment.align with an initial estimate given as parameters [1,6] or as a pose matrix
    [1,4,4] (alignment.py:110-118), float64 inputs staying float64; ICPFrameToModel.register_new_frame (the fused ICP loop
    as a stand-alone call, icp_odometry.py:248-299) returning (params [1,6], T [1,4,4], the executed iterations' losses)."""
    g = torch.Generator().manual_seed(3)
    n = 400
    tgt = torch.randn(1, n, 3, generator=g)
    nrm = torch.nn.functional.normalize(torch.randn(1, n, 3, generator=g), dim=-1)
    x_true = torch.tensor([[0.02, -0.01, 0.015, 0.002, -0.001, 0.003]])
    ref = orc.apply_transformation(tgt, orc.build_pose_matrix(x_true)) + 0.005 * torch.randn(1, n, 3, generator=g)
    al = b200.GaussNewtonPointToPlaneAlignment(b200.GaussNewtonPointToPlaneConfig(
        gauss_newton_config=dict(scheme="geman_mcclure", sigma=0.3, max_iters=4, norm_stop_criterion=1e-12)), ctx=dry.FakeContext())
    x0 = torch.tensor([[0.01, 0.0, 0.01, 0.0, 0.0, 0.002]])
    want, _, _ = orc.gauss_newton_p2plane(ref, tgt, nrm, "geman_mcclure", 0.3, 4, 1e-12, x0.clone())
    _, x_a, _ = al.align(ref, tgt, nrm, initial_estimate=x0)
    _, x_b, _ = al.align(ref, tgt, nrm, initial_estimate=orc.build_pose_matrix(x0))
    np.testing.assert_allclose(np.asarray(x_a), want.numpy(), rtol=0, atol=1e-7)
    np.testing.assert_allclose(np.asarray(x_b), want.numpy(), rtol=0, atol=2e-6)      # through the Euler round trip of x0
    dT64, x64, loss64 = al.align(ref.double().numpy(), tgt.double().numpy(), nrm.double().numpy())
    assert dT64.dtype == x64.dtype == loss64.dtype == np.float64 and dT64.shape == (1, 4, 4) and loss64.shape == (1, n)
    with pytest.raises(AssertionError):
        al.align(ref.double(), tgt.double(), nrm.double(), initial_estimate=orc.build_pose_matrix(x0).double())

    H, W = 32, 512
    cfg = b200.ICPFrameToModelConfig(local_map=b200.KdTreeLocalMapConfig(local_map_size=3),
                                     alignment=b200.GaussNewtonPointToPlaneConfig(gauss_newton_config=dict(
                                         scheme="geman_mcclure", sigma=0.3, max_iters=1)),
                                     max_num_alignments=5, data_key="numpy_pc", threshold_delta_pose=0.0)
    algo = b200.ICPFrameToModel(cfg, projector=b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0),
                                device="cuda:0")
    algo.init()
    pc0, _ = orc.grid_sample(syn.scan(0, H, W), 0.4)
    pc1, _ = orc.grid_sample(syn.scan(1, H, W), 0.4)
    algo.process_next_frame({"numpy_pc": pc0})
    params, T, losses = algo.register_new_frame(pc1, initial_estimate=np.eye(4, dtype=np.float32)[None])
    assert params.shape == (1, 6) and T.shape == (1, 4, 4) and len(losses) == 5 and all(np.isfinite(losses))
    np.testing.assert_allclose(T[0], orc.build_pose_matrix(torch.from_numpy(params))[0].numpy(), rtol=0, atol=1e-6)
    gt = np.linalg.inv(syn.gt_pose(0)) @ syn.gt_pose(1)
    assert np.linalg.norm(T[0, :3, 3] - gt[:3, 3]) < 0.05                      # one frame of the stream: ~0.8 m forward


def test_slam_common_helper_mirrors_plumbing(b200, golden_helpers, golden_misc):
    """The slam/common mirrors (voxelise / voxel_hashing / grid_sample, SphericalProjector, compute_normal_map,
    compute_neighbors, Pose): the bodies of their GPU parity tests against the reference goldens, with the oracle behind
    the C ABI -- argument order, batch handling, output shapes and dtypes, numpy and tensor inputs."""
    import test_gpu_parity as G
    G.test_a1_voxel_hash_golden(b200, golden_helpers)
    G.test_a3_projection_golden(b200, golden_helpers)
    G.test_a3_projection_default_value_and_default_channels(b200, golden_misc)
    G.test_a16_pose_golden(b200, golden_helpers)
    G.test_a6_compute_neighbors_golden(b200, golden_helpers)
    G.test_a6_reference_geometry_property(b200)
    G.test_a4_normal_map_golden_and_oracle(b200, orc, syn, golden_helpers)


def test_more_gpu_parity_bodies_through_the_stand_in(b200, monkeypatch, caplog, golden_helpers, golden_misc):
    """Bodies of further GPU parity tests (tests/test_gpu_parity.py) with the oracle behind the C ABI: the Gauss-Newton
    step goldens of all seven schemes, its error behaviour, the local-map mirrors' life cycles, NaN handling, re-init,
    device-tensor inputs -- the tests' own logic and the mirrors' plumbing, checked before GPU minutes are spent."""
    import test_gpu_parity as G
    dry.cuda_stand_ins(monkeypatch)
    for scheme in G.SCHEMES:
        G.test_gn_step_golden(b200, golden_helpers, scheme)
    G.test_gn_multi_iter_and_known_answer(b200, golden_helpers)
    G.test_gn_singular_raises_and_tiny_residual_warns(b200, orc, _Caplog(caplog))
    G.test_align_bad_shape_is_assertion(b200)
    G.test_kd_local_map_golden(b200, golden_helpers)
    G.test_kd_map_lifecycle_vs_oracle(b200, orc, syn)
    G.test_a5_projective_map_lifecycle_vs_oracle(b200, orc, syn)
    G.test_preprocessing_chain_matches_reference_layout(b200, orc, syn)
    G.test_icp_missing_key_and_bad_shape(b200)
    G.test_nan_rows_and_nan_pixels_match_oracle(b200, orc, syn)
    G.test_odometry_reinit_is_clean(b200, syn)
    G.test_icp_device_tensor_input_and_outputs(b200, syn)
    G.test_a1_voxelise_with_one_voxel_length_per_axis(b200, golden_misc)
    G.test_a3_projection_with_channels_and_batch(b200, orc, syn)


def test_example_script_runs_through_the_stand_in(b200, monkeypatch, capsys):
    """examples/odometry_synthetic.py (the shipped pipeline through the reference-shaped API) stepped on a small sensor with
    the oracle behind the C ABI: the script's own logic, and the poses it reports against the stream's ground truth."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "odometry_synthetic.py")
    spec = importlib.util.spec_from_file_location("odometry_synthetic_example", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    poses, errors = mod.main(frames=5, height=32, width=512, voxel=0.4)
    assert poses.shape == (5, 4, 4) and len(errors) == 4 and max(errors) < 0.05
    assert "frames/s" in capsys.readouterr().out
