"""Registration of the B200 odometry inside a pyLiDAR-SLAM checkout (INTEGRATION.md in code form).

The reference discovers its algorithms through closed `Enum` registries (`slam/odometry/__init__.py:23-32`) and its
default configs through hydra's ConfigStore (`icp_odometry.py:67-68`).  A maintainer adds one enum member and one
config node; `patched_odometry_enum` / `register_hydra_configs` do exactly that programmatically, for deployments that
prefer not to edit the checkout -- and for the test that drives the reference's own `SLAM` loop with this odometry.
"""
import os
from enum import Enum

from .odometry import ICPFrameToModel

ALGORITHM_NAME = "icp_F2M_b200"
CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")


def patched_odometry_enum(reference_enum, reference_config_class, odometry_class=ICPFrameToModel):
    """The reference's ODOMETRY enum plus the member INTEGRATION.md adds:
    `icp_F2M_b200 = (pylidar_slam_b200.ICPFrameToModel, ICPFrameToModelConfig)`.
    Rebind it where the reference looks the algorithm up: `slam.slam.ODOMETRY` and `slam.odometry.ODOMETRY`."""
    loader = [b for b in reference_enum.__mro__ if b.__name__ == "ObjectLoaderEnum"][0]

    class _Loader(loader):
        @classmethod
        def type_name(cls):
            return reference_enum.type_name()

    members = {name: member.value for name, member in reference_enum.__members__.items()}
    members[ALGORITHM_NAME] = (odometry_class, reference_config_class)
    return Enum(reference_enum.__name__, members, type=_Loader)


def register_hydra_configs(reference_config_class):
    """Stores the `slam/odometry/icp_odometry_b200` node (the structured-config twin of the shipped yaml,
    `config/slam/odometry/icp_odometry_b200.yaml`) in hydra's ConfigStore.  Needs hydra importable."""
    from hydra.core.config_store import ConfigStore
    cs = ConfigStore.instance()
    cs.store(group="slam/odometry", name="icp_odometry_b200", node=reference_config_class(algorithm=ALGORITHM_NAME))
    return cs
