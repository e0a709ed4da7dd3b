"""Mirror of the Distortion / Voxelization / GridSample / ToTensor filters and the Preprocessing chain the shipped
`config/slam/preprocessing/{grid_sample,voxelization}.yaml` build (slam/preprocessing.py:43-291)."""
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Optional

import numpy as np
import torch

from .common import assert_debug, check_tensor, distort_frame, grid_sample, voxel_statistics
from .odometry import MISSING, _cfg_to_dict


@dataclass
class FilterConfig:
    filter_name: str = MISSING


class Filter:
    def __init__(self, config: FilterConfig, **kwargs):
        self.config = config

    def filter(self, data_dict: dict):
        raise NotImplementedError("")


@dataclass
class VoxelizationConfig(FilterConfig):
    """slam/preprocessing.py:43-60"""
    filter_name: str = "voxelization"
    input_channel: str = "numpy_pc"
    voxel_covariances_key: str = "voxel_covariances"
    voxel_means_key: str = "voxel_means"
    voxel_size_key: str = "voxel_sizes"
    voxel_indices_key: str = "voxel_indices"
    voxel_hashes_key: str = "voxel_hashes"
    voxel_coordinates_key: str = "voxel_coordinates"
    with_normal_distribution: bool = True
    voxel_size: float = 0.2


class Voxelization(Filter):
    """Voxelization.filter (preprocessing.py:71-97): voxel coordinates, hashes and -- optionally -- every voxel's
    point count, mean and scatter matrix, in one GPU pass (hash, stable radix sort, segmented warp reductions)."""

    def __init__(self, config: VoxelizationConfig, ctx=None, **kwargs):
        super().__init__(config)
        self.ctx = ctx

    def filter(self, data_dict: dict):
        cfg = self.config
        assert_debug(cfg.input_channel in data_dict,
                     f"The input channel {cfg.input_channel} was not in the input channel")
        pointcloud = data_dict[cfg.input_channel]
        assert_debug(isinstance(pointcloud, np.ndarray))
        check_tensor(pointcloud, [-1, 3])
        if not cfg.with_normal_distribution:
            from .common import voxelise, voxel_hashing
            data_dict[cfg.voxel_hashes_key] = voxel_hashing(pointcloud, cfg.voxel_size, ctx=self.ctx)
            data_dict[cfg.voxel_coordinates_key] = voxelise(pointcloud, cfg.voxel_size, ctx=self.ctx)
            return
        coords, hashes, sizes, means, covs, ids = voxel_statistics(pointcloud, cfg.voxel_size, ctx=self.ctx)
        data_dict[cfg.voxel_hashes_key] = hashes
        data_dict[cfg.voxel_coordinates_key] = coords
        data_dict[cfg.voxel_means_key] = means
        data_dict[cfg.voxel_covariances_key] = covs
        data_dict[cfg.voxel_size_key] = sizes
        data_dict[cfg.voxel_indices_key] = ids


@dataclass
class DistortionConfig(FilterConfig):
    """slam/preprocessing.py:129-141"""
    filter_name: str = "distortion"
    pointcloud_key: str = "numpy_pc"
    timestamps_key: str = "numpy_pc_timestamps"
    pose_key: str = "init_rpose"  # Initialization.initial_pose_key()
    output_key: str = "input_data"
    force: bool = False
    activate: bool = True


class Distortion(Filter):
    """Distortion.filter (preprocessing.py:148-191): de-skew the frame with the initial motion estimate."""

    def __init__(self, config: DistortionConfig, ctx=None, **kwargs):
        super().__init__(config)
        self.ctx = ctx

    def filter(self, data_dict: dict):
        cfg = self.config
        pc = data_dict[cfg.pointcloud_key]
        assert_debug(isinstance(pc, np.ndarray), "Cannot Distort a non numpy frame")
        check_tensor(pc, [-1, 3])
        no_distortion = not cfg.activate or (cfg.timestamps_key not in data_dict)
        no_distortion = no_distortion or (data_dict[cfg.pose_key] is None if cfg.pose_key in data_dict else False)
        if no_distortion:
            data_dict[cfg.output_key] = pc
            return
        rpose = data_dict[cfg.pose_key]
        check_tensor(rpose, [4, 4])
        timestamps = data_dict[cfg.timestamps_key]
        assert_debug(isinstance(timestamps, np.ndarray))
        timestamps = timestamps.reshape(-1)
        check_tensor(timestamps, [pc.shape[0]])
        data_dict[cfg.output_key] = distort_frame(pc, timestamps, rpose, ctx=self.ctx)


@dataclass
class GridSampleConfig(FilterConfig):
    filter_name: str = "grid_sample"
    voxel_size: float = 0.3
    pointcloud_key: str = "numpy_pc"
    output_indices_key: str = "sample_indices"
    output_sample_key: str = "sample_points"


class GridSample(Filter):
    """GridSample.filter (preprocessing.py:213-226): one point per voxel hash, on the GPU (K1)."""

    def __init__(self, config: GridSampleConfig, ctx=None, **kwargs):
        super().__init__(config)
        self.ctx = ctx

    def filter(self, data_dict: dict):
        pc = data_dict[self.config.pointcloud_key]
        assert_debug(isinstance(pc, np.ndarray), "Cannot Distort a non numpy frame")
        check_tensor(pc, [-1, 3])
        sample, indices = grid_sample(pc, self.config.voxel_size, ctx=self.ctx)
        data_dict[self.config.output_sample_key] = sample
        data_dict[self.config.output_indices_key] = indices


@dataclass
class ToTensorConfig(FilterConfig):
    filter_name: str = "to_tensor"
    device: str = "cpu"
    keys: Any = MISSING


class ToTensor(Filter):
    """ToTensor.filter (preprocessing.py:112-126)."""

    def __init__(self, config: ToTensorConfig, device: str = "cpu", **kwargs):
        super().__init__(config)
        self.device = torch.device(device)

    def filter(self, data_dict: dict):
        for old_key, new_key in _cfg_to_dict(self.config.keys).items():
            assert_debug(old_key in data_dict)
            np_array = data_dict[old_key]
            assert_debug(isinstance(np_array, np.ndarray))
            data_dict[new_key] = torch.from_numpy(np_array).to(self.device)


class FILTER(Enum):
    distortion = (Distortion, DistortionConfig)
    voxelization = (Voxelization, VoxelizationConfig)
    grid_sample = (GridSample, GridSampleConfig)
    to_tensor = (ToTensor, ToTensorConfig)

    @staticmethod
    def load(config, **kwargs) -> Filter:
        d = _cfg_to_dict(config)
        assert_debug("filter_name" in d)
        assert_debug(d["filter_name"] in FILTER.__members__, f"filter {d['filter_name']} is not on the B200 hot path")
        _class, _config = FILTER[d["filter_name"]].value
        return _class(_config(**d), **kwargs)


@dataclass
class PreprocessingConfig:
    filters: Optional[Dict[str, Any]] = None


class Preprocessing:
    """Preprocessing (preprocessing.py:268-291): filters applied in sorted-key order."""

    def __init__(self, preprocessing_config: PreprocessingConfig, **kwargs):
        self.config = preprocessing_config
        self.filters = []
        filters_config = _cfg_to_dict(self.config).get("filters")
        if filters_config is not None:
            for key in sorted(filters_config.keys()):
                self.filters.append(FILTER.load(filters_config[key], **kwargs))

    def forward(self, data_dict: dict):
        for _filter in self.filters:
            _filter.filter(data_dict)
