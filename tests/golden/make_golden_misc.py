"""Goldens for two arguments of the reference's interfaces that no shipped config uses: anisotropic voxels
(voxelise, pointcloud.py:55-79) and `default_value` of Projector.build_projection_map (projection.py:333,378-391).
(The third one, the `mask` of the Gauss-Newton alignments, raises inside the reference itself -- optimization.py:391-392
and 500-501 multiply the [b,n,6] Jacobian in place by mask.unsqueeze(1), a [b,1,n,1] tensor -- so there is nothing to
pin; tests/test_dropin_reference.py checks that it still does.)  From the UNMODIFIED reference under
oracle/ref_shims.py.  Build container only:  python tests/golden/make_golden_misc.py -> misc.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)
ns = ref_shims.load_reference(kdtree_workers=-1)
rs = np.random.RandomState(7)
out = {}

# ---- voxelise with three voxel lengths
pts = (rs.randn(5000, 3) * np.array([25.0, 25.0, 2.5])).astype(np.float32)
out["vox_points"] = pts
out["vox_sizes"] = np.array([0.4, 0.25, 0.1])
out["vox_coords"] = np.asarray(ns.pointcloud.voxelise(pts, 0.4, 0.25, 0.1)).astype(np.int64)

# ---- build_projection_map with a non-zero default value (and an extra channel)
H, W = 16, 256
proj = ns.projection.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0)
scan = syn.scan(1, H, W)[::3]                                   # leaves pixels empty
cloud = np.concatenate([scan, rs.rand(len(scan), 1).astype(np.float32)], axis=1)
out["proj_cloud"] = cloud
out["proj_default"] = np.float32(-7.5)
out["proj_map"] = proj.build_projection_map(torch.from_numpy(cloud[None]), default_value=-7.5)[0].numpy()

np.savez_compressed(os.path.join(HERE, "misc.npz"), **out)
print({k: np.asarray(v).shape for k, v in out.items()})
