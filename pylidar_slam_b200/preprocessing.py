"""Mirror of the GridSample / ToTensor filters and the Preprocessing chain the shipped
`config/slam/preprocessing/grid_sample.yaml` builds (slam/preprocessing.py:100-126,195-226,230-291)."""
from dataclasses import dataclass
from enum import Enum
from typing import Any, Dict, Optional

import numpy as np
import torch

from .common import assert_debug, check_tensor, grid_sample
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
