"""TEST INFRASTRUCTURE ONLY -- import shims for the *unmodified* reference.

This module is used by ``tests/golden/make_golden.py`` (run in the build
container, where ``/root/reference`` exists) to import Kitware/pyLiDAR-SLAM
as-is and produce golden vectors.  Nothing in the product path, ``bench.py``
or the ``-m gpu`` tests may import it: ``/root/reference`` does not exist on
the GPU box.

The reference targets hydra 1.0 / typeguard 2 / numpy 1 / pykdtree, none of
which are installed here.  We install stub modules in
``sys.modules`` *without editing any reference file*:

* ``typeguard.check_type``: accept the 2.x ``(name, value, type)`` call form
  used at slam/common/utils.py:74
* ``np.round_``: removed in numpy 2 (slam/common/pointcloud.py:73-75)
* ``omegaconf`` / ``hydra`` / ``hydra.conf`` / ``hydra.core.config_store``
* ``pykdtree.kdtree.KDTree`` over ``scipy.spatial.cKDTree`` (exact k-NN,
  float64 distances) -- pykdtree itself is unpinned in requirements.txt:5 and
  not vendored; parity at that boundary is therefore "unpinned" by the
  reference and pinned only by exactness of both searches.
* ``matplotlib`` / ``seaborn`` / ``open3d`` auto-stubs
* ``slam.viz.color_map``: icp_odometry.py:20 obtains ``torch`` and
  ``assert_debug`` through ``from slam.viz.color_map import *``.
"""
import dataclasses
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PLS_REFERENCE_ROOT", "/root/reference")

_installed = False


class _AutoStub(types.ModuleType):
    """A module whose every attribute is another auto-stub / no-op callable."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _AutoStub(f"{self.__name__}.{name}")
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return _AutoStub("call")

    def __mro_entries__(self, bases):
        return (object,)

    N = 2


def _stub_package(name):
    mod = _AutoStub(name)
    mod.__path__ = []
    sys.modules[name] = mod
    return mod


def install(kdtree_workers=1):
    """Install the shims and make ``import slam`` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    import numpy as np
    if not hasattr(np, "round_"):
        np.round_ = np.round

    # numba's njit(parallel) inside the reference calls np.round_ from jitted
    # code; numba resolves it through its own table which still knows round_.
    import typeguard
    _orig_check = typeguard.check_type

    def _check_type(*args, **kwargs):
        if len(args) == 3 and isinstance(args[0], str):
            return _orig_check(args[1], args[2])
        return _orig_check(*args, **kwargs)

    typeguard.check_type = _check_type

    # ---- omegaconf -------------------------------------------------------
    oc = types.ModuleType("omegaconf")

    class DictConfig(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError:
                raise AttributeError(k)

        def __setattr__(self, k, v):
            self[k] = v

    class OmegaConf:
        @staticmethod
        def get_type(obj):
            return type(obj)

        @staticmethod
        def create(obj=None):
            return DictConfig(obj or {})

        @staticmethod
        def to_yaml(obj):
            return str(obj)

    oc.DictConfig = DictConfig
    oc.OmegaConf = OmegaConf
    oc.MISSING = "???"
    sys.modules["omegaconf"] = oc

    # ---- hydra -----------------------------------------------------------
    hydra = types.ModuleType("hydra")
    hydra.__path__ = []

    def _main(*a, **k):
        def deco(fn):
            return fn
        return deco

    hydra.main = _main
    conf = types.ModuleType("hydra.conf")
    conf.dataclass = dataclasses.dataclass
    conf.field = dataclasses.field
    conf.MISSING = "???"
    core = types.ModuleType("hydra.core")
    core.__path__ = []
    cstore = types.ModuleType("hydra.core.config_store")

    class _Node:
        def __init__(self, node):
            self.node = node

    class ConfigStore:
        _inst = None

        def __init__(self):
            self.repo = {}

        @classmethod
        def instance(cls):
            if cls._inst is None:
                cls._inst = ConfigStore()
            return cls._inst

        def store(self, name, node, group=None, **kw):
            key = f"{group}/{name}.yaml" if group else f"{name}.yaml"
            self.repo[key] = node

        def load(self, path):
            node = self.repo.get(path)
            if node is None:
                return None
            if isinstance(node, type):
                node = node()
            return _Node(node)

    cstore.ConfigStore = ConfigStore
    conf.ConfigStore = ConfigStore
    hydra.conf = conf
    hydra.core = core
    core.config_store = cstore
    sys.modules.update({"hydra": hydra, "hydra.conf": conf, "hydra.core": core,
                        "hydra.core.config_store": cstore})

    # ---- pykdtree over scipy cKDTree --------------------------------------
    from scipy.spatial import cKDTree
    pk = types.ModuleType("pykdtree")
    pk.__path__ = []
    pkk = types.ModuleType("pykdtree.kdtree")

    class KDTree:
        def __init__(self, data, leafsize=16):
            self._tree = cKDTree(data, leafsize=leafsize)

        def query(self, points, k=1, **kw):
            d, i = self._tree.query(points, k=k, workers=kdtree_workers)
            return d, i

    pkk.KDTree = KDTree
    pk.kdtree = pkk
    sys.modules.update({"pykdtree": pk, "pykdtree.kdtree": pkk})

    # ---- viz / plotting auto-stubs ----------------------------------------
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.colors",
                 "matplotlib.patches", "matplotlib.lines", "seaborn", "open3d", "cv2"):
        if name not in sys.modules or name == "cv2":
            _stub_package(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["matplotlib"].colors = sys.modules["matplotlib.colors"]

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # ---- slam.viz.color_map re-exporting torch / assert_debug --------------
    import torch
    from typing import Optional, Union
    importlib.import_module("slam")
    importlib.import_module("slam.common.utils")
    viz = types.ModuleType("slam.viz")
    viz.__path__ = []
    cmap = types.ModuleType("slam.viz.color_map")
    cmap.torch = torch
    cmap.np = np
    cmap.Optional = Optional
    cmap.Union = Union
    cmap.assert_debug = sys.modules["slam.common.utils"].assert_debug
    cmap.scalar_gray_cmap = lambda *a, **k: None
    cmap.tensor_to_image = lambda *a, **k: None
    cmap.__all__ = ["torch", "np", "Optional", "Union", "assert_debug",
                    "scalar_gray_cmap", "tensor_to_image"]
    viz.color_map = cmap
    sys.modules["slam.viz"] = viz
    sys.modules["slam.viz.color_map"] = cmap
    _installed = True


def load_reference(kdtree_workers=1):
    """Returns a namespace with the reference's hot-path symbols."""
    install(kdtree_workers)
    ns = types.SimpleNamespace()
    ns.icp = importlib.import_module("slam.odometry.icp_odometry")
    ns.local_map = importlib.import_module("slam.odometry.local_map")
    ns.alignment = importlib.import_module("slam.odometry.alignment")
    ns.optimization = importlib.import_module("slam.common.optimization")
    ns.projection = importlib.import_module("slam.common.projection")
    ns.geometry = importlib.import_module("slam.common.geometry")
    ns.pose = importlib.import_module("slam.common.pose")
    ns.rotation = importlib.import_module("slam.common.rotation")
    ns.pointcloud = importlib.import_module("slam.common.pointcloud")
    ns.utils = importlib.import_module("slam.common.utils")
    return ns
