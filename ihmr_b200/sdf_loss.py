"""Drop-in for the ``sdf`` package the reference imports at
/root/reference/src/models/loss_utils.py:13 (``from sdf import SDFLoss, SDFLoss_Single``),
constructs at :34-38 and calls at :181-182 (boundary L0 of SURVEY.md §8(b)).

``SDFLoss(faces_right, faces_left, robustifier=None)(hand_verts (B,2,778,3),
return_per_vert_loss=True, return_origin_scale_loss=True)`` ->
``(losses (B,), per_vert (B,1556), origin_scale (B,1556))``; ``losses`` is differentiable
w.r.t. ``hand_verts``.  The voxel field, its trilinear sampling and the gradient are one CUDA
kernel in libihmr_b200.so (csrc/sdf.cu); no PyTorch implementation exists behind this module.
To use it for the unmodified reference: ``sys.modules['sdf'] = ihmr_b200.sdf_loss``.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .mano_layer import DeviceModel, _f32c, _ptr, _stream, faces_only_model


class _SdfFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, hand_verts, scale_factor):
        hv = _f32c(hand_verts)
        if hv.dim() != 4 or tuple(hv.shape[1:]) != (2, 778, 3):
            raise ValueError("hand_verts must be (B,2,778,3)")
        B, dev = hv.shape[0], hv.device
        model = module._device_model(dev)
        model.set_sdf_conventions(scale_factor, module.ray_axis)
        losses = torch.empty(B, device=dev, dtype=torch.float32)
        per_vert = torch.empty(B, 1556, device=dev, dtype=torch.float32)
        origin = torch.empty(B, 1556, device=dev, dtype=torch.float32)
        need_grad = hand_verts.requires_grad
        grad = torch.empty_like(hv) if need_grad else None
        rob = float(module.robustifier) if module.robustifier else 0.0
        ws = torch.empty(max(1, model._lib.ihmr_sdf_workspace_bytes(B)), device=dev, dtype=torch.uint8)
        if getattr(module, "exact", False):
            _lib.check(model._lib.ihmr_sdf_loss_exact(model.handle, B, _ptr(hv), _ptr(losses), _ptr(per_vert), _ptr(origin),
                                                      _ptr(grad), _ptr(ws), ws.numel(), _stream(dev)), "ihmr_sdf_loss_exact")
        else:
            _lib.check(model._lib.ihmr_sdf_loss(model.handle, B, _ptr(hv), _ptr(losses), _ptr(per_vert), _ptr(origin),
                                                _ptr(grad), rob, _ptr(ws), ws.numel(), _stream(dev)), "ihmr_sdf_loss")
        ctx.grad = grad
        ctx.mark_non_differentiable(per_vert, origin)
        return losses, per_vert, origin

    @staticmethod
    def backward(ctx, g_losses, _g_pv, _g_or):
        if ctx.grad is None:
            return None, None, None
        return None, ctx.grad * g_losses.view(-1, 1, 1, 1), None


class SDFLoss(nn.Module):
    def __init__(self, faces_right, faces_left, grid_size=32, robustifier=None, debugging=False, ray_axis=0):
        """``ray_axis`` (not a reference argument): world axis of the inside/outside parity ray, assumption A4 of
        SURVEY.md §8(c); 0 = +x is what the oracle restates."""
        super().__init__()
        if ray_axis not in (0, 1, 2):
            raise ValueError("ray_axis must be 0, 1 or 2")
        self.ray_axis = int(ray_axis)
        if grid_size != 32:
            raise ValueError("the penetration kernel is built for the reference's 32^3 grid (A1)")
        self.faces_right = np.asarray(faces_right).astype(np.int32)
        self.faces_left = np.asarray(faces_left).astype(np.int32)
        self.grid_size = grid_size
        self.robustifier = robustifier
        self._models: Dict[int, DeviceModel] = {}

    def _device_model(self, dev: torch.device) -> DeviceModel:
        if dev.type != "cuda":
            raise _lib.IhmrError("ihmr_b200 has no CPU path: hand_verts must live on a CUDA (sm_100) device")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if idx not in self._models:
            self._models[idx] = faces_only_model(self.faces_right, self.faces_left, idx)
        return self._models[idx]

    def forward(self, hand_verts, return_per_vert_loss=False, return_origin_scale_loss=False, scale_factor=0.2):
        losses, per_vert, origin = _SdfFn.apply(self, hand_verts, float(scale_factor))
        if return_per_vert_loss and return_origin_scale_loss:
            return losses, per_vert, origin
        if return_per_vert_loss:
            return losses, per_vert
        if return_origin_scale_loss:
            return losses, origin
        return losses


class SDFLossExact(SDFLoss):
    """NOT the reference's function — the exact, grid-free penetration mode of SURVEY.md §8(f) rank 3 behind the same
    call signature: per vertex the exact distance to the other hand's mesh if the vertex is inside it, else 0 (the limit
    of the 32^3 field for an infinitely fine grid; no 7 mm voxel quantisation).  Use it by name; ``SDFLoss`` stays the
    reference-parity module."""
    exact = True

    def __init__(self, faces_right, faces_left, grid_size=32, robustifier=None, debugging=False, ray_axis=0):
        if robustifier:
            raise ValueError("the exact mode has no robustifier")
        super().__init__(faces_right, faces_left, grid_size, None, debugging, ray_axis)


class SDFLoss_Single(SDFLoss):
    """Exported because loss_utils.py:13 imports the name; not used on the IHMR-OPT path."""
