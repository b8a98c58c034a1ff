"""``OptimizeModel`` — boundary L1 of SURVEY.md §8(b): the object /root/reference/src/optimize.py
drives (:47,64-70) and src/utils/evaluator.py reads (:27-29), re-implemented on the fused
C-ABI path.  Same constructor argument (the ``opt`` namespace), same methods
(``set_input``, ``init_optimize``, ``optimize``, ``forward``, ``get_pred_result``,
``get_current_errors``), same attributes (``inputSize``, ``batch_size``, ``mano_models``).

Where the reference runs several hundred eager launches, a CPU round trip and a fresh autograd
graph per iteration (src/models/optimize_model.py:390-414), one ``ihmr_opt_stage`` call
enqueues a whole stage on the current CUDA stream; nothing synchronises until
``get_pred_result``.
"""
from __future__ import annotations

import ctypes as C
import os.path as osp
import sys
from collections import OrderedDict
from typing import Dict

import numpy as np
import torch

from . import _lib
from .mano_layer import create as create_mano
from .mano_layer import _ptr, _stream
from .strategies import strategies

DEFAULT_LOSS_WEIGHTS = dict(joints_2d_loss=10.0, joints_3d_loss=1000.0, trans_loss_weight=100.0,
                            shape_reg_loss_weight=0.1, collision_loss_weight=1.0,
                            finger_reg_loss_weight=100000.0)   # optimize_model.py:84-92


class OptimizeModel:
    @property
    def name(self):
        return "OptimizeModel"

    def __init__(self, opt, device=None):
        self.opt = opt
        self.isTrain = getattr(opt, "isTrain", False)
        self.process_rank = getattr(opt, "process_rank", -1)
        self.inputSize = opt.inputSize
        self.total_params_dim = opt.total_params_dim
        self.cam_params_dim, self.pose_params_dim = opt.cam_params_dim, opt.pose_params_dim
        self.shape_params_dim, self.trans_params_dim = opt.shape_params_dim, opt.trans_params_dim
        assert self.total_params_dim == (self.cam_params_dim + self.trans_params_dim
                                         + self.pose_params_dim + self.shape_params_dim)
        assert (self.cam_params_dim, self.trans_params_dim, self.pose_params_dim, self.shape_params_dim) == (3, 3, 96, 20)
        self.batch_size = opt.batchSize
        # every batch-mean loss divides by this; a rank that holds a shard of a larger batch
        # passes the larger batch's size so its frames see identical gradients (Appendix D.9)
        self.bs_norm = int(getattr(opt, "bs_norm", None) or self.batch_size)
        if not torch.cuda.is_available():
            raise _lib.IhmrError("ihmr_b200.OptimizeModel needs a CUDA (sm_100) device; there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.lib = _lib.load()

        self.mano_models = dict()
        for hand_type in ("left", "right"):
            path = osp.join(opt.model_root, f"MANO_{hand_type.upper()}.pkl")
            self.mano_models[hand_type] = create_mano(path, "mano", use_pca=False, is_rhand=(hand_type == "right"),
                                                      batch_size=self.batch_size * 2)
        # same in-place fix-up as optimize_model.py:109-113 (no effect on this path: M6)
        sl, sr = self.mano_models["left"].shapedirs, self.mano_models["right"].shapedirs
        if torch.mean(torch.abs(sl[:, 0, :] - sr[:, 0, :])) < 1e-7:
            sl[:, 0, :] *= -1
        self._model = self._build_device_model()

        self.strategy = strategies[opt.strategy] if isinstance(opt.strategy, str) else opt.strategy
        self.default_loss_weights = dict(DEFAULT_LOSS_WEIGHTS)
        assert abs(self.default_loss_weights["collision_loss_weight"] - 1.0) < 1e-7
        self._ws = None
        self._buf: Dict[str, torch.Tensor] = {}
        self._pinned: Dict[str, torch.Tensor] = {}

    def _build_device_model(self):
        from .mano_layer import DeviceModel
        right, left = self.mano_models["right"], self.mano_models["left"]
        arrays = dict(right._arrays)
        arrays["shapedirs"] = right.shapedirs.detach().cpu().numpy()
        return DeviceModel(arrays, right.faces, left.faces, self.device.index or 0)

    # -------------------------------------------------------------------------- input
    def set_input(self, input):
        """H2D copy of one batch (keys of src/data/opt_dataset.py:176-196)."""
        dev = self.device

        def put(key):
            t = input[key]
            t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
            t = t.to(torch.float32)
            if t.device.type == "cpu" and not t.is_pinned():
                # pageable source: one staging copy; pinned sources go straight to the device
                return t.to(dev, non_blocking=False).contiguous()
            return t.to(dev, non_blocking=True).contiguous()

        self.hand_type_array = put("hand_type_array")
        self.hand_type_valid = put("hand_type_valid")
        self.joints_2d, self.joints_3d = put("joints_2d"), put("joints_3d")
        self.hand_trans = put("hand_trans")
        self.gt_pose_params, self.gt_shape_params = put("mano_pose"), put("mano_betas")
        self.mano_params_weight = put("mano_params_weight")
        self.init_cam, self.init_pose_params = put("init_cam"), put("init_pose_params")
        self.init_shape_params, self.init_hand_trans = put("init_shape_params"), put("init_hand_trans")
        self.init_joints_2d, self.init_joints_3d = put("init_joints_2d"), put("init_joints_3d")
        self.init_hand_trans_j = put("init_hand_trans_j")
        B = self.init_cam.shape[0]
        assert B == self.batch_size, "batch rows must equal opt.batchSize (optimize_model.py:185)"
        assert self.init_joints_2d.shape == (B, 42, 3) and self.init_joints_3d.shape == (B, 42, 4)
        self._targets = _lib.Targets(
            init_joints_2d=self.init_joints_2d.data_ptr(), init_joints_3d=self.init_joints_3d.data_ptr(),
            init_hand_trans_j=self.init_hand_trans_j.data_ptr(), gt_joints_3d=self.joints_3d.data_ptr(),
            hand_type_array=self.hand_type_array.data_ptr())

    def init_optimize(self):
        """optimize_model.py:235-251: start from the prior prediction. The seven parameter
        groups live in one (B,122) matrix [cam | trans | pose 96 | shape 20]."""
        self.params = torch.cat([self.init_cam, self.init_hand_trans[:, 0, :3], self.init_pose_params,
                                 self.init_shape_params], dim=1).contiguous()
        assert self.params.shape == (self.batch_size, 122)

    # views with the reference's attribute names
    pred_cam_params = property(lambda self: self.params[:, 0:3])
    pred_hand_trans = property(lambda self: self.params[:, 3:6].reshape(-1, 1, 3))
    pred_pose_params = property(lambda self: self.params[:, 6:102])
    pred_shape_params = property(lambda self: self.params[:, 102:122])
    pred_right_orient = property(lambda self: self.params[:, 6:9])
    pred_right_pose_params = property(lambda self: self.params[:, 9:54])
    pred_left_orient = property(lambda self: self.params[:, 54:57])
    pred_left_pose_params = property(lambda self: self.params[:, 57:102])
    pred_right_shape_params = property(lambda self: self.params[:, 102:112])
    pred_left_shape_params = property(lambda self: self.params[:, 112:122])

    # ------------------------------------------------------------------------ compute
    def _workspace(self):
        need = self.lib.ihmr_opt_workspace_bytes(self.batch_size)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def _out(self, name, *shape):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != shape:
            t = torch.empty(*shape, device=self.device, dtype=torch.float32)
            self._buf[name] = t
        return t

    def run_stage(self, stage: dict):
        ws = self._workspace()
        st = _lib.make_stage(stage)
        optimizer = _lib.OPTIMIZERS[getattr(self.opt, "optimizer", "adam")]
        _lib.check(self.lib.ihmr_opt_stage(self._model.handle, self.batch_size, self.bs_norm, _ptr(self.params),
                                           C.byref(self._targets), C.byref(st), int(self.opt.save_mid_freq),
                                           optimizer, _ptr(ws), ws.numel(), _stream(self.device)), "ihmr_opt_stage")

    def forward(self):
        """Final-style forward: fills pred_*_hand_verts, pred_joints_3d (root aligned, as the
        reference leaves it after __compute_loss), collision outputs."""
        B, ws = self.batch_size, self._workspace()
        self.pred_right_hand_verts = self._out("rv", B, 778, 3)
        self.pred_left_hand_verts = self._out("lv", B, 778, 3)
        self.pred_joints_3d = self._out("j3d", B, 42, 3)
        self.collision_loss_batch = self._out("col", B)
        self.collision_loss_origin_scale = self._out("ori", B, 1556)
        self.joints_3d_loss_p_batch = self._out("j3dp", B)
        _lib.check(self.lib.ihmr_opt_final(self._model.handle, B, _ptr(self.params), C.byref(self._targets),
                                           _ptr(self.pred_right_hand_verts), _ptr(self.pred_left_hand_verts),
                                           _ptr(self.pred_joints_3d), _ptr(self.collision_loss_batch),
                                           _ptr(self.collision_loss_origin_scale), _ptr(self.joints_3d_loss_p_batch),
                                           _ptr(ws), ws.numel(), _stream(self.device)), "ihmr_opt_final")

    def optimize(self, iter_id=0, num_iter=1):
        for stage_id, stage in enumerate(self.strategy):
            self.run_stage(stage)
            if self.process_rank <= 0 and not getattr(self.opt, "quiet", False):
                print(f"iter:{iter_id + 1:04d}/{num_iter:04d}, stage-{stage_id:02d} completes")
                sys.stdout.flush()
        self.forward()      # after optimization completes, forward again (optimize_model.py:413-414)

    def value_and_grad(self, stage: dict):
        """One iteration's six weighted batch losses and d loss / d params (B,122), no step."""
        ws = self._workspace()
        st = _lib.make_stage(stage)
        losses = torch.empty(6, device=self.device, dtype=torch.float32)
        grad = torch.empty_like(self.params)
        _lib.check(self.lib.ihmr_opt_value_and_grad(self._model.handle, self.batch_size, self.bs_norm,
                                                    _ptr(self.params), C.byref(self._targets), C.byref(st),
                                                    _ptr(losses), _ptr(grad), _ptr(ws), ws.numel(),
                                                    _stream(self.device)), "ihmr_opt_value_and_grad")
        return losses, grad

    # ------------------------------------------------------------------------- output
    def _to_host(self, name: str, t: torch.Tensor) -> np.ndarray:
        """D2H through a cached pinned staging buffer (async on the current stream)."""
        t = t.detach().contiguous()
        buf = self._pinned.get(name)
        if buf is None or buf.shape != t.shape or buf.dtype != t.dtype:
            buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            self._pinned[name] = buf
        buf.copy_(t, non_blocking=True)
        return buf.numpy()

    def get_pred_result(self):
        """optimize_model.py:418-435: the 13 numpy arrays the evaluator consumes. The arrays
        are views of pinned staging buffers and stay valid until the next call."""
        B = self.batch_size
        c = self._to_host
        res = OrderedDict(
            pred_cam_params=c("cam", self.pred_cam_params), pred_hand_trans=c("trans", self.pred_hand_trans),
            pred_shape_params=c("shape", self.pred_shape_params), pred_pose_params=c("pose", self.pred_pose_params),
            pred_right_hand_verts=c("rv", self.pred_right_hand_verts),
            pred_left_hand_verts=c("lv", self.pred_left_hand_verts),
            mano_params_weight=c("mpw", self.mano_params_weight), pred_joints_3d=c("j3d", self.pred_joints_3d),
            gt_joints_3d=c("gtj3d", self.joints_3d), collision_loss=c("col", self.collision_loss_batch),
            collision_loss_origin_scale=c("ori", self.collision_loss_origin_scale),
            do_flip=np.zeros(B).astype(np.int32), pred_hand_type=np.ones(B).astype(np.int32))
        torch.cuda.current_stream(self.device).synchronize()     # the one sync of the batch
        return res

    def profile_iteration(self, stage: dict):
        """Device milliseconds of each kernel class for one iteration of `stage` (measurement aid)."""
        ws = self._workspace()
        st = _lib.make_stage(stage)
        ms = (C.c_float * len(_lib.KERNEL_CLASSES))()
        _lib.check(self.lib.ihmr_opt_profile_iteration(self._model.handle, self.batch_size, self.bs_norm,
                                                       _ptr(self.params), C.byref(self._targets), C.byref(st), ms,
                                                       _ptr(ws), ws.numel(), _stream(self.device)),
                   "ihmr_opt_profile_iteration")
        return dict(zip(_lib.KERNEL_CLASSES, [float(x) for x in ms]))

    def get_current_errors(self):
        """optimize_model.py:438-455 (log-only values; plain tensor arithmetic on the exported
        outputs, not part of the refinement loop)."""
        losses, _ = self.value_and_grad(dict(update_params=[], loss_weights=self.default_loss_weights, lr=0.0,
                                             epoch=0, filter_loss=[("joints_3d_loss_p", "+0")],
                                             select_loss="joints_3d_loss_p"))
        l = losses.cpu().numpy()
        B = float(self.batch_size)
        with torch.no_grad():
            # the exported joints are root aligned; the 2-D projection needs the right wrist back
            root = self.mano_models["right"](global_orient=self.pred_right_orient.contiguous(),
                                             hand_pose=self.pred_right_pose_params.contiguous(),
                                             betas=self.pred_right_shape_params.contiguous()).joints[:, 0:1]
            has_r = (self.joints_3d[:, 0, 3] > 0.5).view(-1, 1, 1).float()
            world = self.pred_joints_3d + root * has_r
            cam = self.pred_cam_params.reshape(-1, 1, 3)
            j2d = cam[:, :, 0:1] * (world[:, :, :2] + cam[:, :, 1:])
            d2 = (self.joints_2d[:, :, :2] - j2d).abs() * self.joints_2d[:, :, 2:3]
            gt = self.joints_3d[:, :, :3] - self.joints_3d[:, 0:1, :3] * has_r
            d3 = (gt - self.pred_joints_3d) ** 2 * self.joints_3d[:, :, 3:4]
            dt = (self.hand_trans[:, :, :3] - self.pred_hand_trans) ** 2 * self.hand_trans[:, :, 3:4]
        return OrderedDict([
            ("joints_2d_loss", float(d2.mean().item())),
            ("joints_3d_loss", float(d3.mean().item()) * 1000),
            ("hand_trans_loss", float(dt.mean().item()) * 10),
            ("collision_loss", float(l[3]) * self.bs_norm / B),
            ("joints_3d_loss_p", float(l[1]) * self.bs_norm / B),
        ])
