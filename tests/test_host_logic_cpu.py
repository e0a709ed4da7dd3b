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
