"""Mirror of the unsupervised point-to-plane training loss (slam/training/loss_modules.py:21-132): same config
dataclasses, same `nn.Module` interface (`forward(data_dict) -> (loss, data_dict)`, `point_to_plane_loss(...)`), the
arithmetic -- forward AND backward -- in one fused CUDA pass behind `pls_p2plane_loss`.  torch is only the autograd
plumbing: `_P2PlaneLossFn` hands the gradient the kernel already computed back to the graph.

Gradients follow the reference's autograd exactly, including index_put's backward through the z-buffer scatter (every
point written to a pixel receives that pixel's gradient, not only the survivor)."""
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

from . import _lib
from .common import Pose, SphericalProjector, assert_debug, check_tensor, compute_normal_map
from .odometry import MISSING, _cfg_to_dict


@dataclass
class LossConfig:
    """slam/training/loss_modules.py:21-24"""
    mode: str = MISSING


@dataclass
class PointToPlaneLossConfig(LossConfig):
    """slam/training/loss_modules.py:28-34"""
    mode: str = "unsupervised"
    least_square_scheme: Optional[Dict[str, Any]] = field(default_factory=lambda: dict(scheme="geman_mcclure", sigma=0.5))


def _require_cuda(t):
    assert_debug(t.is_cuda, "the B200 loss runs on CUDA tensors; there is no CPU path")


class _P2PlaneLossFn(torch.autograd.Function):
    @staticmethod
    def forward(fn_ctx, pose_tensor, vm_target, vm_reference, nm_reference, module):
        B, _, H, W = vm_target.shape
        dev = vm_target.device
        f32 = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731
        vt, vr, nr, pose = f32(vm_target), f32(vm_reference), f32(nm_reference), f32(pose_tensor)
        is_matrix = pose.shape[-1] == 4
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        per_batch = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.empty((B, 4, 4) if is_matrix else (B, 6), dtype=torch.float32, device=dev)
        proj = module.projector
        module.ctx.call("pls_p2plane_loss", _lib.ptr(vt), _lib.ptr(vr), _lib.ptr(nr), _lib.ptr(pose) if is_matrix else None,
                        None if is_matrix else _lib.ptr(pose), B, H, W, float(proj.up_fov), float(proj.down_fov),
                        module.scheme, module.sigma, _lib.ptr(loss), _lib.ptr(per_batch),
                        _lib.ptr(grad) if is_matrix else None, None if is_matrix else _lib.ptr(grad))
        fn_ctx.save_for_backward(grad)
        fn_ctx.pose_dtype = pose_tensor.dtype
        module.last_loss_per_batch = per_batch
        return loss.reshape(())

    @staticmethod
    def backward(fn_ctx, grad_output):
        (grad,) = fn_ctx.saved_tensors
        return (grad * grad_output).to(fn_ctx.pose_dtype), None, None, None, None


class _PointToPlaneLossModule(nn.Module):
    """_PointToPlaneLossModule (loss_modules.py:39-132)."""

    def __init__(self, config: PointToPlaneLossConfig, projector: SphericalProjector, pose: Pose = None, ctx=None):
        nn.Module.__init__(self)
        self.pose = pose or Pose("euler")
        self.projector = projector
        self.config = config
        ls = dict(_cfg_to_dict(config.least_square_scheme) if config.least_square_scheme is not None else {})
        name = ls.get("scheme", "default")
        assert_debug(name in _lib.SCHEMES, f"unknown weighting scheme {name}")
        self.scheme = _lib.SCHEMES[name]
        self.sigma = float(ls.get("sigma", 0.5))
        self._ctx = ctx
        self.last_loss_per_batch = None

    @property
    def ctx(self):
        from .common import default_context
        return self._ctx or default_context()

    def point_to_plane_loss(self, vm_target, vm_reference, nm_reference, pose_tensor, data_dict: dict = None):
        """Loss between a batch of target vertex maps and reference vertex / normal maps under the predicted
        target-to-reference poses `[B,6]` (Euler parameters) or `[B,4,4]` (loss_modules.py:51-104)."""
        check_tensor(vm_target, [-1, 3, -1, -1])
        b, _, h, w = vm_target.shape
        check_tensor(vm_reference, [b, 3, h, w])
        check_tensor(nm_reference, [b, 3, h, w])
        if pose_tensor.shape[-1] == 4:
            check_tensor(pose_tensor, [b, 4, 4])
        else:
            check_tensor(pose_tensor, [b, 6])
        _require_cuda(vm_target)
        return _P2PlaneLossFn.apply(pose_tensor, vm_target, vm_reference, nm_reference, self)

    def forward(self, data_dict: dict):
        vertex_map = data_dict["vertex_map"]
        if "normal_map" not in data_dict:
            check_tensor(vertex_map, [-1, 2, 3, -1, -1])
            b, seq, _, h, w = vertex_map.shape
            normal_map = compute_normal_map(vertex_map.reshape(b * seq, 3, h, w), ctx=self._ctx).reshape(b, seq, 3, h, w)
            data_dict["normal_map"] = normal_map
        normal_map = data_dict["normal_map"]
        b, s, _, h, w = vertex_map.shape
        assert_debug(s == 2)
        tgt_vmap, ref_vmap, ref_nmap = vertex_map[:, 1], vertex_map[:, 0], normal_map[:, 0]
        loss_icp = self.point_to_plane_loss(tgt_vmap, ref_vmap, ref_nmap, data_dict["pose_params"], data_dict).mean()
        return loss_icp, data_dict
