"""Development helper (NOT part of the product, never imported by it): runs the bodies of tests/test_next_rows_gpu.py on
a machine without a GPU by standing a CPU fake in for the four new C-ABI entry points (and the few old ones those
tests touch).  The fake follows include/plslam_b200.h argument by argument and computes with the CPU oracle / the host
harness, so it checks the Python mirrors' plumbing and the tests' own logic (shapes, dtypes, tolerances, error
paths) before GPU minutes are spent.  It proves nothing about the kernels -- tests/test_host_math.py and the `-m gpu`
run do that.  `FakeContext` is also what tests/test_host_logic_cpu.py patches in (per test, via monkeypatch) to check
the Python mirrors' host logic in the CPU suite.

    python tests/dryrun_next_rows.py
"""
import ctypes as C
import os
import sys
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))  # conftest helpers

from oracle import icp_oracle as orc  # noqa: E402
from oracle import next_rows_oracle as nxt  # noqa: E402
from pylidar_slam_b200 import _lib  # noqa: E402

SCHEME_NAMES = {v: k for k, v in _lib.SCHEMES.items()}


def arr(addr, shape, dtype):
    if addr is None:
        return None
    if hasattr(addr, "value"):
        addr = addr.value
    n = int(np.prod(shape))
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(addr))
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


def cuda_stand_ins(monkeypatch):
    """TEST-ONLY: tensor factories and Tensor.to asked for a cuda device hand back CPU tensors, so that host logic written
    for device tensors (the reference-shaped ICP loop, bench.py's control flow) can be stepped through without a GPU."""
    real_to = torch.Tensor.to

    def on_cpu(fn):
        def wrapped(*a, **k):
            if k.get("device") is not None and torch.device(k["device"]).type == "cuda":
                k["device"] = "cpu"
            return fn(*a, **k)
        return wrapped

    def to(self, *a, **k):
        if a and isinstance(a[0], (torch.device, str)) and torch.device(a[0]).type == "cuda":
            a = a[1:]
            if not a and not k:
                return self
        if k.get("device") is not None and torch.device(k["device"]).type == "cuda":
            k = {kk: v for kk, v in k.items() if kk != "device"}
            if not a and not k:
                return self
        return real_to(self, *a, **k)

    for name in ("empty", "zeros", "ones", "eye", "tensor", "full"):
        monkeypatch.setattr(torch, name, on_cpu(getattr(torch, name)))
    monkeypatch.setattr(torch.Tensor, "to", to)


class FakeContext:
    def __init__(self, **kwargs):
        self.kwargs = kwargs
        self.cfg = _lib.PlsConfig()
        self.cfg.device = int(kwargs.get("device", 0) or 0)
        self._staging = {}  # "device" twins of pls_grid_sample_staged live in host memory here

    def close(self):
        pass

    def call(self, name, *a):
        return getattr(self, name)(*a)

    def pls_distort(self, xyz, x64, ts, t64, n, pose, p64, out):
        pc = arr(xyz, (n, 3), np.float64 if x64 else np.float32)
        t = arr(ts, (n,), np.float64 if t64 else np.float32)
        P = arr(pose, (4, 4), np.float64 if p64 else np.float32)
        arr(out, (n, 3), np.float64)[:] = nxt.distort(pc, t, P)

    def pls_voxel_statistics(self, xyz, is64, n, voxel, coords, hashes, sizes, means, covs, ids, count):
        dt = np.float64 if is64 else np.float32
        r = nxt.voxelization(arr(xyz, (n, 3), dt), voxel)
        V = len(r["voxel_sizes"])
        if coords:
            arr(coords, (n, 3), np.int64)[:] = r["voxel_coordinates"]
        if hashes:
            arr(hashes, (n,), np.int64)[:] = r["voxel_hashes"]
        arr(sizes, (n,), np.int64)[:V] = r["voxel_sizes"]
        arr(means, (n, 3), dt)[:V] = r["voxel_means"]
        arr(covs, (n, 3, 3), dt)[:V] = r["voxel_covariances"]
        arr(ids, (n,), np.int64)[:] = r["voxel_indices"]
        count._obj.value = V

    def pls_voxel_hash(self, xyz, is64, n, voxel, coords, hashes):
        c = orc.voxel_coords(arr(xyz, (n, 3), np.float64 if is64 else np.float32), voxel)
        if coords:
            arr(coords, (n, 3), np.int64)[:] = c
        if hashes:
            arr(hashes, (n,), np.int64)[:] = orc.voxel_hashes(c)

    def pls_voxel_hash_xyz(self, xyz, is64, n, vx, vy, vz, coords, hashes):
        pts = arr(xyz, (n, 3), np.float64 if is64 else np.float32).astype(np.float64)
        c = np.rint(pts / np.array([vx, vy, vz])).astype(np.int64)       # round half to even, like np.round_ (pointcloud.py:73-75)
        if coords:
            arr(coords, (n, 3), np.int64)[:] = c
        if hashes:
            arr(hashes, (n,), np.int64)[:] = orc.voxel_hashes(c)

    def pls_grid_sample(self, xyz, is64, n, voxel, out, idx, count):
        dt = np.float64 if is64 else np.float32
        s, i = orc.grid_sample(arr(xyz, (n, 3), dt), voxel)
        arr(out, (n, 3), dt)[:len(i)] = s
        if idx:
            arr(idx, (n,), np.int64)[:len(i)] = i
        count._obj.value = len(i)

    def pls_grid_sample_staged(self, xyz, is64, n, voxel, out_host, idx_host, out_dev, count):
        dt = np.float64 if is64 else np.float32
        s, i = orc.grid_sample(arr(xyz, (n, 3), dt), voxel)
        self._staging = {"xyz": np.ascontiguousarray(s, dtype=dt), "idx": np.ascontiguousarray(i, dtype=np.int64),
                         "twin": np.ascontiguousarray(s, dtype=dt).copy()}
        if out_host._obj.value and idx_host._obj.value:   # caller-owned staging (pls_pinned_alloc buffers)
            arr(out_host._obj.value, (len(i), 3), dt)[:] = self._staging["xyz"]
            arr(idx_host._obj.value, (len(i),), np.int64)[:] = self._staging["idx"]
        else:
            out_host._obj.value = self._staging["xyz"].ctypes.data
            idx_host._obj.value = self._staging["idx"].ctypes.data
        out_dev._obj.value = self._staging["twin"].ctypes.data
        count._obj.value = len(i)

    def pls_wait_stream(self, stream):
        pass

    # ---- rank 4: ingestion and pose chains
    def pls_kitti_correct_scan(self, scan, n, stride, out):
        from oracle import io_oracle as ioo
        arr(out, (n, 3), np.float64)[:] = ioo.kitti_correct_scan(arr(scan, (n, stride), np.float32))

    def pls_ingest_scan(self, scan, n, stride, correct, H, W, up, down, out_xyz, out_vmap):
        from oracle import io_oracle as ioo
        s = arr(scan, (n, stride), np.float32)
        xyz = ioo.kitti_correct_scan(s) if correct else s[:, :3].astype(np.float64)
        if out_xyz:
            arr(out_xyz, (n, 3), np.float64)[:] = xyz
        arr(out_vmap, (3, H, W), np.float64)[:] = ioo.project_f64(xyz, H, W, up, down)

    def pls_relative_poses(self, poses, n, is64, out):
        from oracle import io_oracle as ioo
        dt = np.float64 if is64 else np.float32
        arr(out, (n, 4, 4), dt)[:] = ioo.relative_poses(arr(poses, (n, 4, 4), dt))

    def pls_absolute_poses(self, rel, n, is64, out):
        from oracle import io_oracle as ioo
        dt = np.float64 if is64 else np.float32
        arr(out, (n, 4, 4), dt)[:] = ioo.absolute_poses(arr(rel, (n, 4, 4), dt))

    def pls_align_p2point(self, ref, tgt, n, is64, scheme, sigma, max_iters, norm_stop, x0, dT, x, loss):
        dt = np.float64 if is64 else np.float32
        r, t = (torch.from_numpy(arr(p, (1, n, 3), dt).copy()) for p in (ref, tgt))
        x0t = None if not x0 else torch.from_numpy(arr(x0, (1, 6), dt).copy())
        try:
            dTo, xo, lo, status = nxt.align_p2point(r, t, SCHEME_NAMES[scheme], sigma, max_iters, norm_stop, x0t)
        except orc.SingularHessian:
            return _lib.check(None, _lib.PLS_E_SINGULAR)
        arr(dT, (4, 4), dt)[:] = dTo[0].numpy()
        arr(x, (6,), dt)[:] = xo[0].numpy()
        arr(loss, (n,), dt)[:] = lo[0].numpy()
        if status == "tiny_residual":
            import logging
            logging.warning("The residual norm is lower than threshold 1e-7. ")

    def pls_weighted_procrustes(self, tgt, ref, w, n, is64, out):
        dt = np.float64 if is64 else np.float32
        ww = None if not w else arr(w, (n, 1), dt)
        arr(out, (4, 4), np.float64)[:] = nxt.weighted_procrustes(arr(tgt, (n, 3), dt), arr(ref, (n, 3), dt), ww)

    def pls_p2plane_loss(self, vt, vr, nr, mats, params, B, H, W, up, down, scheme, sigma, loss, pb, gm, gp):
        maps = [torch.from_numpy(arr(p, (B, 3, H, W), np.float32).copy()) for p in (vt, vr, nr)]
        if mats:
            M = torch.from_numpy(arr(mats, (B, 4, 4), np.float32).copy())
        else:
            M = orc.build_pose_matrix(torch.from_numpy(arr(params, (B, 6), np.float32).copy()))
        lo, per_batch, g = nxt.p2plane_training_loss(maps[0], maps[1], maps[2], M, orc.Projector(H, W, up, down), SCHEME_NAMES[scheme], sigma)
        arr(loss, (1,), np.float32)[:] = lo
        if pb:
            arr(pb, (B,), np.float32)[:] = per_batch
        if gm:
            arr(gm, (B, 4, 4), np.float32)[:] = g
        if gp:
            arr(gp, (B, 6), np.float32)[:] = nxt.pose_matrix_grad_to_params(arr(params, (B, 6), np.float32), g)

    def pls_normal_map(self, vmap, B, H, W, ksize, out):
        arr(out, (B, 3, H, W), np.float32)[:] = orc.normal_map(torch.from_numpy(arr(vmap, (B, 3, H, W), np.float32).copy()), ksize).numpy()

    # ---- the odometry entry points (host logic of the ICPFrameToModel mirror) -- the oracle's ICP behind the C ABI
    def pls_odometry_init(self):
        k = self.kwargs
        names = {v: n for n, v in _lib.SCHEMES.items()}
        cfg = orc.ICPConfig(max_num_alignments=k["max_num_alignments"], threshold_delta_pose=k["threshold_delta_pose"],
                            threshold_trans=k["threshold_trans"], threshold_rot=k["threshold_rot"], data_key="data",
                            local_map="kdtree" if k["local_map_type"] == _lib.MAP_KDTREE else "projective",
                            local_map_size=k["local_map_size"], num_neighbors_normals=k["num_neighbors_normals"],
                            normals_kernel_size=k["normals_kernel_size"], scheme=names[k["scheme"]], sigma=k["sigma"])
        self.algo = orc.ICPFrameToModelOracle(cfg, orc.Projector(k["height"], k["width"], k["up_fov_deg"], k["down_fov_deg"]))

    def pls_process_frame(self, data, layout, n, init, out_pose, out_params, has_pose, info):
        layout &= 0xff  # residency hints: everything is host memory here
        H, W = self.kwargs["height"], self.kwargs["width"]
        f64 = layout in (_lib.INPUT_NDARRAY_F64, _lib.INPUT_TENSOR_F64)
        if layout == _lib.INPUT_VERTEX_MAP:
            x = torch.from_numpy(arr(data, (1, 3, H, W), np.float32).copy())
        else:
            a = arr(data, (n, 3), np.float64 if f64 else np.float32).copy()
            x = a if layout in (_lib.INPUT_NDARRAY, _lib.INPUT_NDARRAY_F64) else torch.from_numpy(a)
        dd = {"data": x, "init_rpose": None if not init else arr(init, (4, 4), np.float32).astype(np.float64)}
        nan_rows = 0 if layout == _lib.INPUT_VERTEX_MAP else int(np.isnan(np.asarray(x, dtype=np.float64)).any(axis=1).sum())
        self.algo.process_next_frame(dd)
        out = arr(info, (12,), np.float64)
        out[:] = 0.0
        out[5] = nan_rows
        if "odometry_pose" not in dd:
            has_pose._obj.value = 0
            return
        has_pose._obj.value = 1
        T = dd["odometry_pose"].astype(np.float32)
        arr(out_pose, (4, 4), np.float32)[:] = T
        arr(out_params, (6,), np.float32)[:] = orc.from_pose_matrix(torch.from_numpy(T).unsqueeze(0))[0].numpy()
        out[0] = len(self.algo.losses[-1])
        if layout == _lib.INPUT_VERTEX_MAP:
            out[8:11] = self.algo.pc.reshape(-1, 3)[0].numpy()

    # ---- the fine-grained plug-ins (LocalMap / RigidAlignment mirrors, the projector): the oracle's classes behind the
    #      header's signatures
    def pls_map_init(self):
        k = self.kwargs
        self.kd = orc.KdTreeLocalMap(k.get("local_map_size", 20), k.get("num_neighbors_normals", 10))
        self.pm = None
        if "height" in k:
            self.pm = orc.ProjectiveLocalMap(orc.Projector(k["height"], k["width"], k.get("up_fov_deg", 3.0), k.get("down_fov_deg", -24.0)),
                                             k.get("local_map_size", 20), k.get("normals_kernel_size", 5))

    def _maps(self):
        if not hasattr(self, "kd"):
            self.pls_map_init()

    def pls_kdmap_update_points(self, rel, pts, n):
        self._maps()
        self.kd.update(arr(rel, (4, 4), np.float32).copy(), new_points=None if not pts else arr(pts, (n, 3), np.float32).copy())

    def pls_kdmap_update_vertex_map(self, rel, vm, H, W):
        self._maps()
        self.kd.update(arr(rel, (4, 4), np.float32).copy(), new_vertex_map=torch.from_numpy(arr(vm, (1, 3, H, W), np.float32).copy()))

    def pls_kdmap_size(self, out):
        self._maps()
        out._obj.value = 0 if self.kd.points is None else self.kd.points.shape[0]

    def pls_kdmap_points(self, out):
        arr(out, self.kd.points.shape, np.float32)[:] = self.kd.points

    def pls_kdmap_nn_search(self, q, n, nb, nrm, idx):
        pts, normals, ids = self.kd.nearest_neighbor_search(arr(q, (n, 3), np.float32).copy())
        arr(nb, (n, 3), np.float32)[:] = pts
        if nrm:
            arr(nrm, (n, 3), np.float32)[:] = normals
        if idx:
            arr(idx, (n,), np.int64)[:] = ids

    def pls_projmap_update(self, rel, vm):
        self._maps()
        H, W = self.kwargs["height"], self.kwargs["width"]
        self.pm.update(torch.from_numpy(arr(rel, (1, 4, 4), np.float32).copy()),
                       None if not vm else torch.from_numpy(arr(vm, (1, 3, H, W), np.float32).copy()))

    def pls_projmap_num_frames(self, out):
        self._maps()
        out._obj.value = 0 if self.pm.vmaps is None else self.pm.vmaps.shape[0]

    def pls_projmap_model(self, out_v, out_n):
        arr(out_v, tuple(self.pm.model_vmap.shape), np.float32)[:] = self.pm.model_vmap.numpy()
        arr(out_n, tuple(self.pm.model_nmap.shape), np.float32)[:] = self.pm.model_nmap.numpy()

    def pls_projmap_nn_search(self, q, n, nb, nrm, tgt, count):
        a, b, c = self.pm.nearest_neighbor_search(torch.from_numpy(arr(q, (n, 3), np.float32).copy()))
        nc = a.shape[1]
        hw = self.kwargs["height"] * self.kwargs["width"]
        arr(nb, (hw, 3), np.float32)[:nc] = a[0].numpy()
        arr(nrm, (hw, 3), np.float32)[:nc] = b[0].numpy()
        arr(tgt, (hw, 3), np.float32)[:nc] = c[0].numpy()
        count._obj.value = nc

    def pls_align_p2plane(self, ref, tgt, nrm, n, is64, scheme, sigma, max_iters, norm_stop, x0, dT, x, loss):
        dt = np.float64 if is64 else np.float32
        r, t, m = (torch.from_numpy(arr(p, (1, n, 3), dt).copy()) for p in (ref, tgt, nrm))
        x0t = None if not x0 else torch.from_numpy(arr(x0, (1, 6), dt).copy())
        try:
            xo, lo, status = orc.gauss_newton_p2plane(r, t, m, SCHEME_NAMES[scheme], sigma, max_iters, norm_stop, x0t)
        except orc.SingularHessian:
            return _lib.check(None, _lib.PLS_E_SINGULAR)
        arr(dT, (4, 4), dt)[:] = orc.build_pose_matrix(xo)[0].numpy()
        arr(x, (6,), dt)[:] = xo[0].numpy()
        arr(loss, (n,), dt)[:] = lo[0].numpy()
        if status == "tiny_residual":
            import logging
            logging.warning("The residual norm is lower than threshold 1e-7. ")

    def pls_project_pixels(self, xyz, n, H, W, up, down, out):
        row, col = orc.Projector(H, W, up, down).pixels(torch.from_numpy(arr(xyz, (1, n, 3), np.float32).copy()))
        arr(out, (n, 2), np.float32)[:] = torch.stack([row[0], col[0]], dim=1).numpy()

    def pls_compute_neighbors(self, tgt, ref, fields, K, Cf, H, W, out_nb, out_nf):
        t = torch.from_numpy(arr(tgt, (1, 3, H, W), np.float32).copy())
        r = torch.from_numpy(arr(ref, (K, 3, H, W), np.float32).copy())
        f = None if not fields else torch.from_numpy(arr(fields, (K, Cf, H, W), np.float32).copy())
        nb, nf = orc.compute_neighbors(t, r, f)
        arr(out_nb, (3, H, W), np.float32)[:] = nb[0].numpy()
        if f is not None and out_nf:
            arr(out_nf, (Cf, H, W), np.float32)[:] = nf[0].numpy()

    def pls_register_frame(self, pts, n, T0, out_T, out_params, out_losses, out_iters):
        init = torch.eye(4).unsqueeze(0) if not T0 else torch.from_numpy(arr(T0, (1, 4, 4), np.float32).copy())
        params, T, losses = self.algo.register_new_frame(torch.from_numpy(arr(pts, (n, 3), np.float32).copy()), init)
        arr(out_T, (4, 4), np.float32)[:] = T[0].numpy()
        arr(out_params, (6,), np.float32)[:] = params.reshape(6).numpy()
        arr(out_losses, (len(losses),), np.float32)[:] = losses
        out_iters._obj.value = len(losses)

    def pls_build_projection_map_filled(self, xyz, channels, B, n, C_, H, W, up, down, default_value, out):
        pts = torch.from_numpy(arr(xyz, (B, n, 3), np.float32).copy())
        ch = None if not channels else torch.from_numpy(arr(channels, (B, n, C_), np.float32).copy())
        m = orc.Projector(H, W, up, down).build_projection_map(pts, channels=ch, height=H, width=W)
        if default_value != 0.0:   # pixels no point lands on (projection.py:378-391)
            filled = orc.Projector(H, W, up, down).build_projection_map(pts, channels=torch.ones(B, n, 1), height=H, width=W)
            m = torch.where(filled.expand_as(m) > 0, m, torch.full_like(m, default_value))
        arr(out, (B, C_, H, W), np.float32)[:] = m.numpy()

    def pls_build_pose_matrix(self, params, batch, out):
        arr(out, (batch, 4, 4), np.float32)[:] = orc.build_pose_matrix(torch.from_numpy(arr(params, (batch, 6), np.float32).copy())).numpy()

    def pls_from_pose_matrix(self, mats, batch, out):
        arr(out, (batch, 6), np.float32)[:] = orc.from_pose_matrix(torch.from_numpy(arr(mats, (batch, 4, 4), np.float32).copy())).numpy()


def main():
    import pylidar_slam_b200 as b200
    from pylidar_slam_b200 import common
    _lib.Context = FakeContext
    common._default_ctx = FakeContext()
    torch.Tensor.cuda = lambda self, *a, **k: self  # device-tensor legs degrade to CPU tensors
    import conftest
    import test_next_rows_gpu as T
    g = np.load(os.path.join(ROOT, "tests", "golden", "next_rows.npz"))
    from pylidar_slam_b200 import synthetic as syn

    class Caplog:
        records = []

        def at_level(self, lvl):
            import contextlib
            import logging
            cap = self

            class H(logging.Handler):
                def emit(self, record):
                    record.message = record.getMessage()
                    cap.records.append(record)

            @contextlib.contextmanager
            def cm():
                h = H()
                logging.getLogger().addHandler(h)
                try:
                    yield
                finally:
                    logging.getLogger().removeHandler(h)
            return cm()

    runs = []
    for name in T.DIST_CASES:
        runs.append((f"distortion_golden[{name}]", lambda name=name: T.test_distortion_golden(b200, g, name)))
    runs += [("distortion_inactive", lambda: T.test_distortion_inactive_paths_return_the_input(b200, g)),
             ("distortion_full_64", lambda: T.test_distortion_full_size_vs_oracle_and_properties(b200, nxt, syn, 64, 2048)),
             ("distortion_nan_single", lambda: T.test_distortion_nan_and_single_point(b200, g)),
             ("chain", lambda: T.test_shipped_chain_distortion_grid_sample_to_tensor(b200, g))]
    for name in T.VOX_CASES:
        runs.append((f"vox_golden[{name}]", lambda name=name: T.test_voxelization_golden(b200, g, name)))
    runs += [("vox_nostats_errors", lambda: T.test_voxelization_without_statistics_and_errors(b200, g))]
    for n, v in [(131072, 0.2), (2049, 0.05), (200000, 25.0)]:
        runs.append((f"vox_full[{n}]", lambda n=n, v=v: T.test_voxelization_full_size_vs_oracle_and_properties(b200, nxt, syn, n, v)))
    runs += [("vox_device", lambda: T.test_voxelization_device_tensor(b200, nxt))]
    for s in T.SCHEMES:
        runs.append((f"p2p_step[{s}]", lambda s=s: T.test_p2point_step_golden(b200, g, s)))
    runs += [("p2p_x0_multi", lambda: T.test_p2point_initial_estimates_multi_iter_f64_and_device(b200, g)),
             ("p2p_large_errors", lambda: T.test_p2point_large_vs_oracle_and_error_behaviour(b200, nxt, Caplog())),
             ("procrustes_golden", lambda: T.test_procrustes_golden(b200, g)),
             ("procrustes_scan", lambda: T.test_procrustes_scan_size_round_trip(b200, nxt, syn))]
    gl = np.load(os.path.join(ROOT, "tests", "golden", "p2plane_loss.npz"))
    from pylidar_slam_b200 import training
    training._require_cuda = lambda t: None
    for sch in T.SCHEMES:
        runs.append((f"loss_golden[{sch}]", lambda sch=sch: T.test_training_loss_and_gradients_golden(b200, gl, sch)))
    runs += [("loss_normals_shapes", lambda: T.test_training_loss_computes_missing_normal_maps_and_checks_shapes(b200, gl)),
             ("loss_scan_size", lambda: T.test_training_loss_scan_size_vs_oracle(b200, nxt, orc, syn))]
    bad = 0
    for name, fn in runs:
        try:
            fn()
            print("ok  ", name)
        except BaseException as e:
            bad += 1
            print("FAIL", name, type(e).__name__, str(e)[:300])
            traceback.print_exc(limit=4)
    print("dry run:", len(runs) - bad, "ok,", bad, "failed (is_cuda asserts are expected to fail here)")


if __name__ == "__main__":
    main()
