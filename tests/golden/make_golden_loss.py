"""Golden vectors of the unsupervised point-to-plane training loss, produced by the
UNMODIFIED reference module `_PointToPlaneLossModule` (slam/training/loss_modules.py:39-132) under oracle/ref_shims.py:
loss value and its autograd gradient with respect to the pose parameters / the pose matrices, per weighting scheme.

    python tests/golden/make_golden_loss.py        (build container only)  ->  p2plane_loss.npz
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from pylidar_slam_b200 import synthetic as syn  # noqa: E402

torch.set_num_threads(1)  # deterministic closest-wins scatter (see make_golden.py)
ns = ref_shims.load_reference(kdtree_workers=-1)
lm = importlib.import_module("slam.training.loss_modules")
SCHEMES = ["default", "huber", "exp", "neighborhood", "geman_mcclure", "square_geman_mcclure", "cauchy"]


def main():
    H, W, B = 16, 256, 3
    proj = ns.projection.SphericalProjector(height=H, width=W, up_fov=3.0, down_fov=-24.0)
    pose = ns.pose.Pose("euler")
    # pairs (reference frame k, target frame k+1); predicted motion = ground truth + a perturbation
    vms, params = [], []
    rng = np.random.RandomState(5)
    for b in range(B):
        k = 2 * b
        ref = syn.vertex_map_from_scan(syn.scan(k, H, W), H, W)[0]
        tgt = syn.vertex_map_from_scan(syn.scan(k + 1, H, W), H, W)[0].copy()
        tgt[:, 3, 20:40] = 0.0          # null pixels in the target
        vms.append(np.stack([ref, tgt]))
        T = syn.gt_relative_pose(k + 1)
        x = pose.from_pose_matrix(torch.from_numpy(T.astype(np.float32)).unsqueeze(0))[0].numpy()
        params.append(x + rng.normal(0, [0.05, 0.05, 0.02, 0.002, 0.002, 0.004]).astype(np.float32))
    vmap = torch.from_numpy(np.stack(vms).astype(np.float32))          # [B,2,3,H,W]
    vmap[1, 0, :, 5, 100:130] = 0.0                                    # null pixels in a reference map
    x0 = torch.from_numpy(np.stack(params).astype(np.float32))
    nmap = ns.geometry.compute_normal_map(vmap.view(B * 2, 3, H, W)).view(B, 2, 3, H, W)
    out = dict(vertex_map=vmap.numpy(), normal_map=nmap.numpy(), pose_params=x0.numpy())
    for sch in SCHEMES:
        mod = lm._PointToPlaneLossModule(lm.PointToPlaneLossConfig(least_square_scheme=dict(scheme=sch, sigma=0.5)), proj, pose)
        x = x0.clone().requires_grad_(True)
        torch.set_num_threads(1)
        loss, _ = mod({"vertex_map": vmap, "pose_params": x})
        loss.backward()
        out[f"{sch}_loss"] = loss.detach().numpy()
        out[f"{sch}_grad_params"] = x.grad.numpy().copy()
        # pose-matrix input (tgt_to_ref.size(-1) == 4, loss_modules.py:121-123)
        M = pose.build_pose_matrix(x0).clone().requires_grad_(True)
        loss_m, _ = mod({"vertex_map": vmap, "normal_map": nmap, "pose_params": M})
        loss_m.backward()
        out[f"{sch}_loss_matrix"] = loss_m.detach().numpy()
        out[f"{sch}_grad_matrix"] = M.grad.numpy().copy()
        print(sch, float(loss), float(loss_m), x.grad[0].numpy())
    np.savez_compressed(os.path.join(HERE, "p2plane_loss.npz"), **out)


if __name__ == "__main__":
    main()
