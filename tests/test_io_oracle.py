"""The rank-4 oracle (oracle/io_oracle.py) against the goldens of the unmodified reference (tests/golden/io_rows.npz)."""
import os

import numpy as np
import pytest

from oracle import io_oracle as ioo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(ROOT, "tests", "golden", "io_rows.npz"))


def test_kitti_correct_scan_matches_reference(g):
    got = ioo.kitti_correct_scan(g["kitti_scan"])
    ref = g["kitti_corrected"]
    assert got.dtype == ref.dtype == np.float64
    assert np.array_equal(np.isnan(got), np.isnan(ref)) and np.isnan(ref).any()   # the point on the vertical axis
    ok = ~np.isnan(ref).any(axis=1)
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-12 * np.abs(ref[ok]).max()


def test_ingestion_vertex_map_matches_reference(g):
    H, W = g["kitti_hw"]
    vm = ioo.project_f64(g["kitti_corrected"], int(H), int(W))
    assert vm.shape == g["kitti_vmap"].shape and vm.dtype == np.float64
    assert np.array_equal(vm, g["kitti_vmap"])


def test_pose_io_bytes_and_round_trip(g):
    assert np.array_equal(ioo.poses_to_rows(g["poses_in"]), g["poses_df"])
    assert ioo.poses_csv_text(g["poses_in"]).encode() == bytes(g["poses_csv"])
    back = ioo.rows_to_poses(g["poses_df"])
    assert back.dtype == g["poses_back"].dtype and np.array_equal(back, g["poses_back"])


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_pose_chains_match_reference(g, tag):
    rel = ioo.relative_poses(g[f"rel_{tag}_in"])
    assert rel.dtype == g[f"rel_{tag}_out"].dtype
    np.testing.assert_allclose(rel, g[f"rel_{tag}_out"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(ioo.absolute_poses(rel), g[f"abs_{tag}_out"], rtol=0, atol=1e-10)
    # absolute(relative(P)) == P
    np.testing.assert_allclose(ioo.absolute_poses(rel), g[f"rel_{tag}_in"].astype(np.float64), rtol=0, atol=2e-5 if tag == "f32" else 1e-9)
