"""Filters that run before the odometry, GPU-backed, behind the reference's own plug-in surface.

The reference builds its preprocessing chain from `config/slam/preprocessing/{grid_sample,voxelization,none}.yaml`:
a dict of filter configs keyed by order, each naming a `filter_name` that is looked up in the `FILTER` enum
(slam/preprocessing.py:230-291).  The surface kept here is exactly that -- filter names, config field names and
defaults, `data_dict` keys and dtypes -- so either yaml runs unchanged; the work itself happens in the CUDA library:

    distortion    Distortion.filter    slam/preprocessing.py:148-191   -> pls_distort
    voxelization  Voxelization.filter  slam/preprocessing.py:71-97     -> pls_voxel_statistics / pls_voxel_hash
    grid_sample   GridSample.filter    slam/preprocessing.py:213-226   -> pls_grid_sample
    to_tensor     ToTensor.filter      slam/preprocessing.py:112-126   (host: numpy -> torch, renaming)
"""
import dataclasses
from enum import Enum
from typing import Any, Dict, Optional

import numpy as np
import torch

from .common import assert_debug, check_tensor, distort_frame, grid_sample, voxel_hashing, voxel_statistics, voxelise
from .odometry import MISSING, _cfg_to_dict


def _config_class(name: str, doc: str, base=None, **fields):
    """A config dataclass with the given `field=default` pairs (the hydra-visible surface of a filter)."""
    spec = [(k, Any if (v is None or v == MISSING) else type(v), dataclasses.field(default=v)) for k, v in fields.items()]
    cls = dataclasses.make_dataclass(name, spec, bases=(base,) if base else ())
    cls.__doc__ = doc
    return cls


FilterConfig = _config_class("FilterConfig", "slam/preprocessing.py:24-28", filter_name=MISSING)

VoxelizationConfig = _config_class(
    "VoxelizationConfig", "slam/preprocessing.py:43-60", FilterConfig,
    filter_name="voxelization", input_channel="numpy_pc", voxel_covariances_key="voxel_covariances",
    voxel_means_key="voxel_means", voxel_size_key="voxel_sizes", voxel_indices_key="voxel_indices",
    voxel_hashes_key="voxel_hashes", voxel_coordinates_key="voxel_coordinates", with_normal_distribution=True, voxel_size=0.2)

DistortionConfig = _config_class(
    "DistortionConfig", "slam/preprocessing.py:129-141 (pose_key = Initialization.initial_pose_key())", FilterConfig,
    filter_name="distortion", pointcloud_key="numpy_pc", timestamps_key="numpy_pc_timestamps", pose_key="init_rpose",
    output_key="input_data", force=False, activate=True)

GridSampleConfig = _config_class(
    "GridSampleConfig", "slam/preprocessing.py:195-204", FilterConfig,
    filter_name="grid_sample", voxel_size=0.3, pointcloud_key="numpy_pc", output_indices_key="sample_indices",
    output_sample_key="sample_points")

ToTensorConfig = _config_class(
    "ToTensorConfig", "slam/preprocessing.py:100-106", FilterConfig, filter_name="to_tensor", device="cpu", keys=MISSING)


class Filter:
    """A step of the chain: reads and extends the frame's `data_dict` in place (slam/preprocessing.py:31-40)."""

    def __init__(self, config, ctx=None, **kwargs):
        self.config = config
        self.ctx = ctx  # the CUDA context to run on (None: the package's default context on cuda:0)

    def filter(self, data_dict: dict):
        raise NotImplementedError("")

    @staticmethod
    def _host_cloud(data_dict: dict, key: str, what: str) -> np.ndarray:
        """The `[n,3]` numpy cloud stored under `key`; the reference's filters accept nothing else."""
        cloud = data_dict[key]
        assert_debug(isinstance(cloud, np.ndarray), what)
        check_tensor(cloud, [-1, 3])
        return cloud


class Voxelization(Filter):
    """Voxel coordinates and hashes of every point and -- unless `with_normal_distribution` is off -- each voxel's
    point count, mean and scatter matrix plus every point's voxel rank: one hash + stable radix sort + segmented
    warp reduction on the GPU."""

    def filter(self, data_dict: dict):
        c = self.config
        assert_debug(c.input_channel in data_dict, f"The input channel {c.input_channel} was not in the input channel")
        cloud = self._host_cloud(data_dict, c.input_channel, "voxelization needs a numpy point cloud")
        if c.with_normal_distribution:
            coords, hashes, sizes, means, covs, ranks = voxel_statistics(cloud, c.voxel_size, ctx=self.ctx)
            data_dict.update({c.voxel_means_key: means, c.voxel_covariances_key: covs, c.voxel_size_key: sizes,
                              c.voxel_indices_key: ranks})
        else:
            coords, hashes = voxelise(cloud, c.voxel_size, ctx=self.ctx), voxel_hashing(cloud, c.voxel_size, ctx=self.ctx)
        data_dict[c.voxel_hashes_key] = hashes
        data_dict[c.voxel_coordinates_key] = coords


class Distortion(Filter):
    """De-skews the frame with the initial motion estimate: every point is moved by the fraction of the relative pose
    that corresponds to its acquisition time.  Without timestamps, without a pose, or when deactivated, the input array
    itself is passed on (same object, as in the reference)."""

    def filter(self, data_dict: dict):
        c = self.config
        cloud = self._host_cloud(data_dict, c.pointcloud_key, "Cannot Distort a non numpy frame")
        pose_is_none = c.pose_key in data_dict and data_dict[c.pose_key] is None
        if not c.activate or c.timestamps_key not in data_dict or pose_is_none:
            data_dict[c.output_key] = cloud
            return
        rpose = data_dict[c.pose_key]  # a missing key raises KeyError, like the reference
        check_tensor(rpose, [4, 4])
        stamps = data_dict[c.timestamps_key]
        assert_debug(isinstance(stamps, np.ndarray))
        stamps = stamps.reshape(-1)
        check_tensor(stamps, [cloud.shape[0]])
        data_dict[c.output_key] = distort_frame(cloud, stamps, rpose, ctx=self.ctx)


class GridSample(Filter):
    """Keeps one point per voxel (the first one met, voxels in ascending hash order) and its index."""

    def filter(self, data_dict: dict):
        c = self.config
        cloud = self._host_cloud(data_dict, c.pointcloud_key, "Cannot Distort a non numpy frame")
        data_dict[c.output_sample_key], data_dict[c.output_indices_key] = grid_sample(cloud, c.voxel_size, ctx=self.ctx)


class ToTensor(Filter):
    """Renames numpy entries of the dict into torch tensors on `device` (host-side glue, no kernel)."""

    def __init__(self, config, device: str = "cpu", **kwargs):
        super().__init__(config, **kwargs)
        self.device = torch.device(device)
        self._keys = tuple(_cfg_to_dict(self.config.keys).items())
        self._on_host = self.device.type == "cpu"

    def filter(self, data_dict: dict):
        for source, target in self._keys:
            assert_debug(source in data_dict)
            value = data_dict[source]
            assert_debug(isinstance(value, np.ndarray))
            tensor = torch.from_numpy(value)
            data_dict[target] = tensor if self._on_host else tensor.to(self.device)


class FILTER(Enum):
    """The registry the yaml's `filter_name` indexes (slam/preprocessing.py:230-253)."""
    distortion = (Distortion, DistortionConfig)
    voxelization = (Voxelization, VoxelizationConfig)
    grid_sample = (GridSample, GridSampleConfig)
    to_tensor = (ToTensor, ToTensorConfig)

    @staticmethod
    def load(config, **kwargs) -> Filter:
        entries = _cfg_to_dict(config)
        assert_debug("filter_name" in entries)
        name = entries["filter_name"]
        assert_debug(name in FILTER.__members__, f"filter {name} is not on the B200 hot path")
        impl, config_cls = FILTER[name].value
        return impl(config_cls(**entries), **kwargs)


@dataclasses.dataclass
class PreprocessingConfig:
    """slam/preprocessing.py:257-260: `filters` maps an ordering key to one filter config."""
    filters: Optional[Dict[str, Any]] = None


class Preprocessing:
    """The chain itself (slam/preprocessing.py:268-291): filters instantiated once, applied in sorted-key order."""

    def __init__(self, preprocessing_config: PreprocessingConfig, **kwargs):
        self.config = preprocessing_config
        chain = _cfg_to_dict(self.config).get("filters") or {}
        self.filters = [FILTER.load(chain[key], **kwargs) for key in sorted(chain)]

    def forward(self, data_dict: dict):
        for step in self.filters:
            step.filter(data_dict)
