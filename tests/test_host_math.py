"""The per-element arithmetic of the new kernels, run on the CPU.

filters.cu / registration.cu / gn.cu keep their math in __host__ __device__ headers; tests/host_harness.cu wraps
those headers in host loops (nvcc, host code only, no GPU needed).  Checked here against the goldens of the
unmodified reference, so that the device code is pinned even on a box without a GPU; the `-m gpu` tests then only
have to confirm the launch plumbing and the parallel reductions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import dist_tol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCHEMES = {"default": 0, "least_square": 1, "huber": 2, "exp": 3, "neighborhood": 4, "geman_mcclure": 5,
           "square_geman_mcclure": 6, "cauchy": 7}
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.fixture(scope="module")
def hh():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "host_harness.so")
    src = os.path.join(ROOT, "tests", "host_harness.cu")
    deps = [src] + [os.path.join(ROOT, "pylidar_slam_b200", "csrc", f) for f in
                    ("filters_device.cuh", "registration_device.cuh", "gn_device.cuh", "pose_device.cuh",
                     "projection_device.cuh", "training_device.cuh", "eigen_device.cuh", "ingest_device.cuh")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
                               "-o", so, src])
    lib = C.CDLL(so)
    lib.hh_align.restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def test_distort_point_math_matches_reference(hh, golden_next):
    g = golden_next
    for name in ["f64", "f32", "mixed", "const", "big", "pc64"]:
        pc, ts, pose = (np.ascontiguousarray(g[f"dist_{name}_{k}"]) for k in ("pc", "ts", "pose"))
        out = np.empty((pc.shape[0], 3), np.float64)
        pose64 = np.ascontiguousarray(pose.astype(np.float64))
        hh.hh_distort(_p(pc), int(pc.dtype == np.float64), _p(ts), int(ts.dtype == np.float64), C.c_int64(pc.shape[0]),
                      _p(pose64), int(pose.dtype == np.float64), _p(out))
        err = np.abs(out - g[f"dist_{name}_out"]).max()
        assert err <= dist_tol(pose), (name, err)


def test_distort_nan_timestamp_poisons_everything(hh, golden_next):
    g = golden_next
    pc, ts, pose = (np.ascontiguousarray(g[f"dist_f64_{k}"]) for k in ("pc", "ts", "pose"))
    ts = ts.copy()
    ts[17] = np.nan
    out = np.zeros((pc.shape[0], 3), np.float64)
    hh.hh_distort(_p(pc), 0, _p(ts), 1, C.c_int64(pc.shape[0]), _p(pose), 1, _p(out))
    assert np.isnan(out).all()  # np.max / np.min propagate the NaN into every alpha (preprocessing.py:180-182)


def test_rotation_vector_branches(hh):
    from scipy.spatial.transform import Rotation
    rng = np.random.RandomState(3)
    for angle in [0.0, 1e-12, 1e-9, 1e-5, 0.3, 2.0, np.pi - 1e-5, np.pi - 1e-10, np.pi]:
        axis = rng.randn(3)
        axis /= np.linalg.norm(axis)
        R = np.ascontiguousarray(Rotation.from_rotvec(axis * angle).as_matrix())
        ax, an = np.zeros(3), C.c_double(0)
        hh.hh_rotation_vector(_p(R), _p(ax), C.byref(an))
        back = Rotation.from_rotvec(ax * an.value).as_matrix()
        assert np.abs(back - R).max() <= 1e-7 if angle > 3.1 else np.abs(back - R).max() <= 1e-12, angle


def test_kabsch_tail_matches_reference(hh, golden_next):
    g = golden_next

    def run(pt, pr, w=None):
        w = np.ones((pt.shape[0], 1)) if w is None else w
        aw = w / w.sum()
        mu_t, mu_r = (pt * aw).sum(0), (pr * aw).sum(0)
        Cm = np.ascontiguousarray((pr - mu_r).T @ (pt - mu_t))
        mu = np.ascontiguousarray(np.concatenate([mu_t, mu_r]))
        T = np.zeros(16)
        hh.hh_kabsch(_p(Cm), _p(mu), _p(T))
        return T.reshape(4, 4)

    pt, pr = g["proc_tgt"], g["proc_ref"]
    assert np.abs(run(pt, pr) - g["proc_T"]).max() <= 1e-12
    assert np.abs(run(pt, pr, g["proc_w"]) - g["proc_T_w"]).max() <= 1e-12
    Tm = run(pt, g["proc_ref_mirror"])
    assert np.abs(Tm - g["proc_T_mirror"]).max() <= 1e-12 and np.linalg.det(Tm[:3, :3]) > 0.999
    assert np.abs(run(g["proc_planar_tgt"], g["proc_planar_ref"]) - g["proc_T_planar"]).max() <= 1e-9
    # random rotations incl. near-degenerate spectra
    rng = np.random.RandomState(0)
    from scipy.spatial.transform import Rotation
    for scale in ([1, 1, 1], [5, 1, 0.01], [3, 3, 1], [1, 1e-4, 1e-4]):
        p = rng.randn(200, 3) * np.array(scale)
        R = Rotation.random(random_state=rng.randint(1 << 30)).as_matrix()
        t = rng.randn(3)
        T = run(p, p @ R.T + t)
        tol = 1e-9 if min(scale) >= 0.01 else 1e-5
        assert np.abs(T[:3, :3] - R).max() <= tol and np.abs(T[:3, 3] - t).max() <= tol * 10, scale


def _align(hh, cost, ref, tgt, nrm, scheme, sigma, max_iters, norm_stop, x0=None):
    dt = ref.dtype
    n = ref.shape[0]
    x, dT, loss = np.zeros(6, dt), np.zeros(16, dt), np.zeros(n, dt)
    st = hh.hh_align(cost, int(dt == np.float64), _p(np.ascontiguousarray(ref)), _p(np.ascontiguousarray(tgt)),
                     _p(None if nrm is None else np.ascontiguousarray(nrm)), C.c_int64(n), SCHEMES[scheme], C.c_double(sigma),
                     max_iters, C.c_double(norm_stop), _p(None if x0 is None else np.ascontiguousarray(x0.astype(dt))),
                     _p(x), _p(dT), _p(loss))
    return st, x, dT.reshape(4, 4), loss


@pytest.mark.parametrize("scheme", [s for s in SCHEMES if s != "least_square"])
def test_p2point_device_math_matches_reference(hh, golden_next, scheme):
    g = golden_next
    st, x, dT, loss = _align(hh, 1, g["p2p_ref"], g["p2p_tgt"], None, scheme, 0.3, 1, 1e-3)
    assert st == 0
    rx = g[f"p2p_{scheme}_x"]
    assert np.abs(x - rx).max() <= 2e-5 * max(1.0, np.abs(rx).max())
    assert np.abs(dT - g[f"p2p_{scheme}_dT"]).max() <= 2e-5
    rl = g[f"p2p_{scheme}_loss"]
    assert np.abs(loss - rl).max() <= 1e-5 * max(1.0, np.abs(rl).max())


def test_p2point_device_math_x0_multi_f64(hh, golden_next):
    g = golden_next
    ref, tgt = g["p2p_ref"], g["p2p_tgt"]
    st, x, dT, loss = _align(hh, 1, ref, tgt, None, "geman_mcclure", 0.3, 4, 1e-9, x0=g["p2p_multi_x0"])
    assert st == 0 and np.abs(x - g["p2p_multi_x"]).max() <= 2e-3  # see tests/test_next_rows_oracle.py
    st, x, dT, loss = _align(hh, 1, ref, tgt, None, "huber", 0.3, 1, 1e-3, x0=g["p2p_multi_x0"])
    assert np.abs(x - g["p2p_mat_x"]).max() <= 2e-5 * max(1.0, np.abs(g["p2p_mat_x"]).max())
    st, x, dT, loss = _align(hh, 1, ref.astype(np.float64), tgt.astype(np.float64), None, "default", 0.5, 6, 1e-12)
    assert st == 0
    assert np.abs(x - g["p2p_f64_x"]).max() <= 1e-9 and np.abs(dT - g["p2p_f64_dT"]).max() <= 1e-9
    assert np.abs(loss - g["p2p_f64_loss"]).max() <= 1e-9


def test_p2plane_device_math_matches_reference(hh, golden_helpers):
    """Same harness on the hot path's own alignment: the goldens of a11-a15."""
    g = golden_helpers
    for scheme in ["default", "huber", "geman_mcclure", "cauchy"]:
        st, x, dT, loss = _align(hh, 0, g["gn_ref"], g["gn_tgt"], g["gn_nrm"], scheme, 0.3, 1, 1e-3)
        assert st == 0
        assert np.abs(x - g[f"gn_{scheme}_delta"]).max() <= 2e-6, scheme
        assert np.abs(loss - g[f"gn_{scheme}_loss"]).max() <= 1e-5 * max(1.0, np.abs(g[f"gn_{scheme}_loss"]).max())


@pytest.mark.parametrize("scheme", list(SCHEMES))
def test_training_loss_device_math_matches_reference_autograd(hh, golden_loss, scheme):
    """loss_modules.py:51-132: the kernel's per-point code (zbuf, pixel terms, index_put-style gradient routing, chain
    rule to the Euler parameters) against the loss and the autograd gradients of the unmodified reference."""
    if scheme == "least_square":
        pytest.skip("alias of default")
    g = golden_loss
    vm, nm, x = g["vertex_map"], g["normal_map"], np.ascontiguousarray(g["pose_params"])
    B, _, _, H, W = vm.shape
    vt, vr, nr = (np.ascontiguousarray(a) for a in (vm[:, 1], vm[:, 0], nm[:, 0]))
    loss, pb = np.zeros(1, np.float32), np.zeros(B, np.float32)
    gm, gp = np.zeros((B, 4, 4), np.float32), np.zeros((B, 6), np.float32)
    hh.hh_p2plane_loss(_p(vt), _p(vr), _p(nr), None, _p(x), B, H, W, C.c_float(3.0), C.c_float(-24.0), SCHEMES[scheme],
                       C.c_float(0.5), _p(loss), _p(pb), _p(gm), _p(gp))
    ref = float(g[f"{scheme}_loss"])
    assert abs(float(loss[0]) - ref) <= 2e-5 * abs(ref), (scheme, loss, ref)
    assert abs(float(pb.mean()) - ref) <= 2e-5 * abs(ref)
    rp, rm = g[f"{scheme}_grad_params"], g[f"{scheme}_grad_matrix"]
    assert np.abs(gp - rp).max() <= 2e-4 * np.abs(rp).max(), (scheme, np.abs(gp - rp).max() / np.abs(rp).max())
    assert np.abs(gm - rm).max() <= 2e-4 * np.abs(rm).max(), (scheme, np.abs(gm - rm).max() / np.abs(rm).max())


@pytest.mark.parametrize("scheme", [s for s in SCHEMES if s != "least_square"])
def test_loss_pixel_gradient_matches_finite_differences(hh, scheme):
    """d(C(|r|)^2)/d(p') of training_device.cuh against central differences of the same function's value, for every
    weighting scheme (incl. the neighbourhood scheme's extra term through |p' - q|^2) -- independent of any reference."""
    rng = np.random.RandomState(SCHEMES[scheme])
    worst = 0.0
    for _ in range(200):
        q = rng.randn(3).astype(np.float32) * 5
        n = rng.randn(3).astype(np.float32)
        n /= np.linalg.norm(n)
        pw = (q + rng.randn(3) * rng.choice([0.02, 0.3, 1.5])).astype(np.float32)
        out = np.zeros(5)
        hh.hh_loss_pixel(SCHEMES[scheme], C.c_double(0.5), _p(pw), _p(q), _p(n), _p(out))
        assert out[0] == 1.0
        g = out[2:5].copy()
        r = float(np.dot(n.astype(np.float64), q.astype(np.float64) - pw.astype(np.float64)))
        if scheme == "huber" and abs(abs(r) - 0.5) < 1e-2:
            continue  # the kink of the Huber cost
        fd = np.zeros(3)
        # float32 inputs: step on the float32 grid, large enough for a clean central difference
        for k in range(3):
            h = np.float32(1e-4)
            a, b = pw.copy(), pw.copy()
            a[k] += h
            b[k] -= h
            oa, ob = np.zeros(5), np.zeros(5)
            hh.hh_loss_pixel(SCHEMES[scheme], C.c_double(0.5), _p(a), _p(q), _p(n), _p(oa))
            hh.hh_loss_pixel(SCHEMES[scheme], C.c_double(0.5), _p(b), _p(q), _p(n), _p(ob))
            fd[k] = (oa[1] - ob[1]) / (float(a[k]) - float(b[k]))
        scale = max(np.abs(g).max(), np.abs(fd).max(), 1e-12)
        if abs(r) < 5e-3:
            continue  # |r| changes sign inside the stencil
        worst = max(worst, np.abs(g - fd).max() / scale)
    assert worst <= 5e-3, (scheme, worst)
    # masked pixels contribute nothing
    zero = np.zeros(3, np.float32)
    for args in ((zero, q, n), (pw, zero, n), (pw, q, zero)):
        out = np.ones(5)
        hh.hh_loss_pixel(SCHEMES[scheme], C.c_double(0.5), _p(args[0]), _p(args[1]), _p(args[2]), _p(out))
        assert not out.any()


def test_kabsch_properties_on_random_problems(hh):
    """kabsch_from_cross: always a proper rotation, agrees with numpy's SVD solution (U diag(1,1,det) V^T), on
    well-conditioned, near-planar and reflected problems."""
    rng = np.random.RandomState(7)
    for trial in range(300):
        scale = rng.choice([1.0, 1e-3, 1e3]) * np.array([1.0, rng.uniform(0.05, 1.0), rng.choice([1.0, 0.2, 1e-3])])
        A = rng.randn(50, 3) * scale
        Bm = rng.randn(50, 3) * scale if trial % 5 == 0 else A @ np.linalg.qr(rng.randn(3, 3))[0].T + 1e-3 * scale.max() * rng.randn(50, 3)
        mu_t, mu_r = A.mean(0), Bm.mean(0)
        Cm = np.ascontiguousarray((Bm - mu_r).T @ (A - mu_t))
        T = np.zeros(16)
        hh.hh_kabsch(_p(Cm), _p(np.ascontiguousarray(np.concatenate([mu_t, mu_r]))), _p(T))
        R = T.reshape(4, 4)[:3, :3]
        assert abs(np.linalg.det(R) - 1.0) <= 1e-9 and np.abs(R @ R.T - np.eye(3)).max() <= 1e-9
        U, S, Vt = np.linalg.svd(Cm)
        D = np.diag([1.0, 1.0, np.sign(np.linalg.det(U) * np.linalg.det(Vt))])
        if S[1] > 1e-6 * S[0] and (S[1] - S[2]) > 1e-6 * S[0] and (S[0] - S[1]) > 1e-6 * S[0]:  # unique solution
            assert np.abs(R - U @ D @ Vt).max() <= 1e-6, (trial, S)
        # the objective trace(R^T C) is never worse than the SVD solution's
        assert np.trace(R.T @ Cm) >= np.trace((U @ D @ Vt).T @ Cm) - 1e-9 * max(S[0], 1e-300)


def _neighbourhood_moments(rng, n, kind):
    """Second moments about a map point of 10 neighbours, float32 like the kernel forms them: noisy planar patches
    (the common case), near-lines and near-isotropic blobs (the ill-conditioned ones)."""
    covs = np.empty((n, 6), np.float32)
    for i in range(n):
        R, _ = np.linalg.qr(rng.standard_normal((3, 3)))
        scale = {"plane": (0.12, 0.1, 0.01), "line": (0.15, 0.004, 0.003), "blob": (0.1, 0.1, 0.1), "flat": (0.1, 0.1, 1e-6)}[kind]
        d = (rng.standard_normal((10, 3)) * np.array(scale) * rng.uniform(0.3, 3.0)) @ R.T + rng.standard_normal(3) * 0.02
        d = d.astype(np.float32)
        M = (d[:, :, None] * d[:, None, :]).mean(axis=0, dtype=np.float32)
        covs[i] = [M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]]
    return covs


def test_normal_eigen_solvers_agree_with_lapack(hh):
    """The normal's eigen-solver (closed form, Jacobi fall-back when the plane direction is nearly degenerate) against
    numpy's float64 eigh on the same float32 moments: |sin angle| <= 1e-9 * lambda_max / gap -- far below the float32
    resolution of the stored normal -- and the closed form may decline only ill-conditioned inputs."""
    rng = np.random.default_rng(7)
    for kind, must_use_closed_form in (("plane", True), ("flat", True), ("line", False), ("blob", False)):
        covs = _neighbourhood_moments(rng, 4000, kind)
        out = np.empty((len(covs), 3), np.float32)
        used = np.zeros(len(covs), np.int32)
        M = np.zeros((len(covs), 3, 3))
        M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2] = covs.astype(np.float64).T
        M[:, 1, 0], M[:, 2, 0], M[:, 2, 1] = M[:, 0, 1], M[:, 0, 2], M[:, 1, 2]
        w, v = np.linalg.eigh(M)
        ref = v[:, :, 0]
        gap = (w[:, 1] - w[:, 0]) / np.maximum(w[:, 2] - w[:, 0], 1e-300)
        for which in (0, 1, 2):
            hh.hh_smallest_eigenvectors(_p(covs), C.c_int64(len(covs)), which, _p(out), _p(used))
            o = out.astype(np.float64)
            assert np.abs(np.linalg.norm(o[used == 1], axis=1) - 1).max() <= 1e-6
            sin = np.linalg.norm(np.cross(o, ref), axis=1)
            ok = used == 1
            # float32 storage of the result: 6e-8; the solver's own error scales with 1 / gap
            assert (sin[ok] <= 2e-7 + 1e-12 / np.maximum(gap[ok], 1e-12)).all(), (kind, which, sin[ok].max())
            if which == 2:
                declined = gap[used == 0]
                assert declined.size == 0 or declined.max() <= 2e-3, (kind, declined.max())
                if must_use_closed_form:
                    assert used.mean() >= 0.99, (kind, used.mean())


def test_ingestion_and_pose_chain_device_math_matches_reference(hh):
    """kitti_correct_point / inverse4 / matmul4 (ingest_device.cuh) against the goldens of the unmodified reference."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "io_rows.npz"))
    scan = np.ascontiguousarray(g["kitti_scan"])
    out = np.empty((scan.shape[0], 3), np.float64)
    hh.hh_kitti_correct(_p(scan), C.c_int64(scan.shape[0]), 4, _p(out))
    ref = g["kitti_corrected"]
    assert np.array_equal(np.isnan(out), np.isnan(ref))
    ok = ~np.isnan(ref).any(axis=1)
    assert np.abs(out[ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()
    P = np.ascontiguousarray(g["rel_f64_in"])
    rel, absolute = np.empty_like(P), np.empty_like(P)
    hh.hh_pose_chains(_p(P), C.c_int64(P.shape[0]), _p(rel), _p(absolute))
    np.testing.assert_allclose(rel, g["rel_f64_out"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(absolute, g["abs_f64_out"], rtol=0, atol=1e-9)
