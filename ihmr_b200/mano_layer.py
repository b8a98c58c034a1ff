"""Drop-in for the ``smplx`` MANO layer the reference builds at
/root/reference/src/models/optimize_model.py:105-106 and calls at :194-200 (boundary L0 of
SURVEY.md §8(b)).  ``create(model_path, 'mano', use_pca=False, is_rhand=..., batch_size=...)``
returns a module whose ``forward(global_orient=, hand_pose=, betas=)`` yields an object with
``.vertices`` (N,778,3) and ``.joints`` (N,16,3), differentiable w.r.t. all three inputs.

The arithmetic runs in libihmr_b200.so (hand-written sm_100a kernels) through the C ABI; there
is no PyTorch implementation behind it — on a machine without the library or a B200 the call
raises.  To use it in place of smplx for the unmodified reference:
``sys.modules['smplx'] = ihmr_b200.mano_layer``.
"""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .mano_io import load_mano_pkl

STANDARD_PARENTS = np.array([-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14], dtype=np.int32)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"ihmr_b200 kernels are fp32 only, got {t.dtype}")
    return t.contiguous()


class DeviceModel:
    """Owns one ihmr_model_t handle (constants resident on one CUDA device)."""

    def __init__(self, arrays: Dict[str, np.ndarray], faces_right: np.ndarray, faces_left: np.ndarray, device: int):
        lib = _lib.load()
        f32 = lambda a, shape: np.ascontiguousarray(np.asarray(a, dtype=np.float32).reshape(shape))
        i32 = lambda a, shape: np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(shape))
        self._keep = [
            f32(arrays["v_template"], (778, 3)), f32(arrays["shapedirs"], (778, 3, 10)),
            f32(arrays["posedirs"], (135, 2334)), f32(arrays["J_regressor"], (16, 778)),
            f32(arrays["lbs_weights"], (778, 16)), i32(arrays["parents"], (16,)),
            f32(arrays["hands_mean"], (45,)), i32(faces_right, (1538, 3)), i32(faces_left, (1538, 3)),
        ]
        handle = C.c_void_p()
        _lib.check(lib.ihmr_model_create(*[a.ctypes.data_as(C.c_void_p) for a in self._keep], int(device),
                                         C.byref(handle)), "ihmr_model_create")
        self.handle, self.device, self._lib = handle, int(device), lib
        self._sdf_conventions = (0.2, 0)

    def set_sdf_conventions(self, scale_factor: float = 0.2, ray_axis: int = 0):
        """Box scale factor and parity-ray axis of the penetration field for every later
        penetration call on this handle.  Not stream ordered: pending work on other streams is synchronised first."""
        conv = (float(scale_factor), int(ray_axis))
        if conv != self._sdf_conventions:
            torch.cuda.synchronize(self.device)
            _lib.check(self._lib.ihmr_model_set_sdf_conventions(self.handle, C.c_float(conv[0]), conv[1]),
                       "ihmr_model_set_sdf_conventions")
            self._sdf_conventions = conv

    def update_shapedirs(self, shapedirs: np.ndarray):
        sd = np.ascontiguousarray(np.asarray(shapedirs, dtype=np.float32).reshape(778, 3, 10))
        _lib.check(self._lib.ihmr_model_update_shapedirs(self.handle, sd.ctypes.data_as(C.c_void_p),
                                                         _stream(self.device)), "ihmr_model_update_shapedirs")

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self._lib.ihmr_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def faces_only_model(faces_right, faces_left, device: int) -> DeviceModel:
    """A handle that only carries the two face lists (what SDFLoss needs)."""
    z = dict(v_template=np.zeros((778, 3)), shapedirs=np.zeros((778, 3, 10)), posedirs=np.zeros((135, 2334)),
             J_regressor=np.zeros((16, 778)), lbs_weights=np.zeros((778, 16)), parents=STANDARD_PARENTS,
             hands_mean=np.zeros(45))
    return DeviceModel(z, faces_right, faces_left, device)


class _ManoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, layer, global_orient, hand_pose, betas):
        orient, pose, beta = _f32c(global_orient), _f32c(hand_pose), _f32c(betas)
        n = orient.shape[0]
        if orient.shape != (n, 3) or pose.shape != (n, 45) or beta.shape != (n, 10):
            raise ValueError("expected global_orient (N,3), hand_pose (N,45), betas (N,10)")
        dev = orient.device
        model = layer._device_model(dev)
        verts = torch.empty(n, 778, 3, device=dev, dtype=torch.float32)
        joints = torch.empty(n, 16, 3, device=dev, dtype=torch.float32)
        ws = layer._workspace(n, dev)
        _lib.check(model._lib.ihmr_mano_forward(model.handle, n, _ptr(orient), _ptr(pose), _ptr(beta), _ptr(verts),
                                                _ptr(joints), _ptr(ws), ws.numel(), _stream(dev)), "ihmr_mano_forward")
        ctx.save_for_backward(orient, pose, beta)
        ctx.layer = layer
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        orient, pose, beta = ctx.saved_tensors
        layer, dev, n = ctx.layer, orient.device, orient.shape[0]
        model = layer._device_model(dev)
        gv = None if g_verts is None else _f32c(g_verts)
        gj = None if g_joints is None else _f32c(g_joints)
        go, gp, gb = torch.empty_like(orient), torch.empty_like(pose), torch.empty_like(beta)
        ws = layer._workspace(n, dev)
        _lib.check(model._lib.ihmr_mano_backward(model.handle, n, _ptr(orient), _ptr(pose), _ptr(beta), _ptr(gv),
                                                 _ptr(gj), _ptr(go), _ptr(gp), _ptr(gb), _ptr(ws), ws.numel(),
                                                 _stream(dev)), "ihmr_mano_backward")
        return None, go, gp, gb


class ManoLayer(nn.Module):
    """Same attributes the reference touches on a smplx MANO object: ``shapedirs`` (a tensor it
    mutates in place, optimize_model.py:109-113), ``faces`` (ndarray, loss_utils.py:34-35,
    evaluator.py:27-29), ``J_regressor``; ``.cuda()`` works as for any module."""

    def __init__(self, model_path: str, is_rhand: bool = True, batch_size: int = 1, **_):
        super().__init__()
        m = load_mano_pkl(model_path)
        self._arrays = m
        self.is_rhand = is_rhand
        self.batch_size = batch_size
        self.faces = m["faces"]
        for name in ("v_template", "shapedirs", "posedirs", "J_regressor", "lbs_weights", "hands_mean"):
            self.register_buffer(name, torch.tensor(m[name], dtype=torch.float32))
        self.register_buffer("faces_tensor", torch.tensor(m["faces"], dtype=torch.long))
        self.register_buffer("parents", torch.tensor(m["parents"], dtype=torch.long))
        self._models: Dict[int, DeviceModel] = {}
        self._shapedirs_seen: Dict[int, tuple] = {}
        self._ws: Dict[int, torch.Tensor] = {}

    def _device_model(self, dev: torch.device) -> DeviceModel:
        if dev.type != "cuda":
            raise _lib.IhmrError("ihmr_b200 has no CPU path: inputs must live on a CUDA (sm_100) device")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        tag = (self.shapedirs.data_ptr(), self.shapedirs._version)
        if idx not in self._models:
            arrays = dict(self._arrays)
            arrays["shapedirs"] = self.shapedirs.detach().cpu().numpy()
            mirror = self.faces[:, ::-1].copy()
            fr, fl = (self.faces, mirror) if self.is_rhand else (mirror, self.faces)
            self._models[idx] = DeviceModel(arrays, fr, fl, idx)
            self._shapedirs_seen[idx] = tag
        elif self._shapedirs_seen[idx] != tag:        # the caller edited .shapedirs in place
            self._models[idx].update_shapedirs(self.shapedirs.detach().cpu().numpy())
            self._shapedirs_seen[idx] = tag
        return self._models[idx]

    def _workspace(self, n: int, dev: torch.device) -> torch.Tensor:
        need = _lib.load().ihmr_mano_workspace_bytes(n)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        ws = self._ws.get(idx)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            self._ws[idx] = ws
        return ws

    def forward(self, global_orient=None, hand_pose=None, betas=None, **_):
        verts, joints = _ManoFn.apply(self, global_orient, hand_pose, betas)
        return SimpleNamespace(vertices=verts, joints=joints, betas=betas, global_orient=global_orient,
                               hand_pose=hand_pose)


def create(model_path, model_type="mano", use_pca=False, is_rhand=True, batch_size=1, **kw):
    """Signature of ``smplx.create`` as used at src/models/optimize_model.py:105-106."""
    if str(model_type).lower() != "mano":
        raise ValueError("only the MANO layer is provided (the IHMR-OPT path uses nothing else)")
    if use_pca:
        raise ValueError("use_pca=True is not on the IHMR-OPT path (optimize_model.py:106 passes False)")
    return ManoLayer(model_path, is_rhand=is_rhand, batch_size=batch_size, **kw)
