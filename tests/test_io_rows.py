"""SURVEY.md section 8f rank 4 -- dataset -> vertex-map ingestion, pose I/O, pose chains -- through the Python mirrors:
on the GPU (`-m gpu`) the CUDA kernels behind the C ABI against the goldens of the unmodified reference; on CPU the same
bodies against the test-only stand-in for the C ABI (host logic: marshalling, dtypes, file format)."""
import os

import numpy as np
import pytest
import torch

import dryrun_next_rows as dry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "io_rows.npz"))


def _ingestion(b200, g):
    H, W = (int(v) for v in g["kitti_hw"])
    scan, ref = g["kitti_scan"], g["kitti_corrected"]
    got = b200.correct_scan(scan)
    assert got.dtype == np.float64 and got.shape == ref.shape
    assert np.array_equal(np.isnan(got), np.isnan(ref))               # the point on the vertical axis is NaN in both
    ok = ~np.isnan(ref).any(axis=1)
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()
    proj = b200.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0)
    dd = b200.kitti_frame(scan, proj)
    assert set(dd) == {"numpy_pc", "vertex_map"} and isinstance(dd["vertex_map"], torch.Tensor)
    assert dd["vertex_map"].dtype == torch.float64 and tuple(dd["vertex_map"].shape) == (3, H, W)
    assert np.abs(dd["numpy_pc"][ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()
    vm, vref = dd["vertex_map"].numpy(), g["kitti_vmap"]
    # same winner in every pixel except where a pixel coordinate sits on a rounding boundary (libm differences)
    same = np.all(np.abs(vm - vref) <= 1e-9 * np.maximum(np.abs(vref), 1.0), axis=0)
    assert same.mean() >= 1 - 2e-3, same.mean()
    # an [N,3] scan without reflectance and without rectification: the cloud is only widened
    dd2 = b200.kitti_frame(scan[:, :3].copy(), proj, corrected_lidar_channel="vmap", correct=False)
    assert np.array_equal(dd2["numpy_pc"], scan[:, :3].astype(np.float64)) and "vmap" in dd2
    with pytest.raises(AssertionError):
        b200.correct_scan(np.zeros((5, 5), np.float32))


def _pose_io(b200, g, tmp_path):
    path = tmp_path / "seq.poses.txt"
    b200.write_poses_to_disk(str(path), g["poses_in"])
    assert path.read_bytes() == bytes(g["poses_csv"])                   # byte-identical to pandas' file
    back = b200.read_poses_from_disk(str(path))
    assert back.dtype == np.float64 and np.array_equal(back, g["poses_back"])
    with pytest.raises(AssertionError):
        b200.write_poses_to_disk(str(tmp_path / "missing_dir" / "x.txt"), g["poses_in"])
    with pytest.raises(AssertionError):
        b200.write_poses_to_disk(str(path), g["poses_in"][:, :3])


def _pose_chains(b200, g):
    for tag in ("f64", "f32"):
        P = g[f"rel_{tag}_in"]
        rel = b200.compute_relative_poses(P)
        assert rel.dtype == np.float64 and rel.shape == P.shape
        np.testing.assert_allclose(rel, g[f"rel_{tag}_out"], rtol=0, atol=1e-10)
        absolute = b200.compute_absolute_poses(g[f"rel_{tag}_out"])
        np.testing.assert_allclose(absolute, g[f"abs_{tag}_out"], rtol=0, atol=1e-9)
        # round trip on a long trajectory: absolute(relative(P)) == P
        from pylidar_slam_b200 import synthetic as syn
        long = np.stack([syn.gt_pose(k) for k in range(2000)])
        np.testing.assert_allclose(b200.compute_absolute_poses(b200.compute_relative_poses(long)), long, rtol=0, atol=1e-8)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture
def gpu_b200():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import pylidar_slam_b200 as b200
    return b200


@pytest.mark.gpu
def test_ingestion_gpu(gpu_b200, g):
    _ingestion(gpu_b200, g)


@pytest.mark.gpu
def test_pose_io_and_chains_gpu(gpu_b200, g, tmp_path):
    _pose_io(gpu_b200, g, tmp_path)
    _pose_chains(gpu_b200, g)


# ------------------------------------------------------------------------------------------------ CPU (host logic)
@pytest.fixture
def fake_b200(monkeypatch):
    import pylidar_slam_b200 as pkg
    from pylidar_slam_b200 import _lib, common
    monkeypatch.setattr(_lib, "Context", dry.FakeContext)
    monkeypatch.setattr(common, "_default_ctx", dry.FakeContext())
    return pkg


def test_ingestion_host_logic(fake_b200, g):
    _ingestion(fake_b200, g)


def test_pose_io_and_chains_host_logic(fake_b200, g, tmp_path):
    _pose_io(fake_b200, g, tmp_path)
    _pose_chains(fake_b200, g)
