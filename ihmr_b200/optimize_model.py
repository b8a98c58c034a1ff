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

INPUT_KEYS = ("hand_type_array", "hand_type_valid", "joints_2d", "joints_3d", "hand_trans", "mano_pose", "mano_betas",
              "mano_params_weight", "init_cam", "init_pose_params", "init_shape_params", "init_hand_trans",
              "init_joints_2d", "init_joints_3d", "init_hand_trans_j")     # opt_dataset.py:176-196 (tensor entries)


class _PinnedBuffer:
    """One pinned staging tensor exported to numpy through the array interface, so that numpy keeps a reference
    to THIS object (as the base of the array and of every view of it) for as long as the data is in use."""
    _TYPESTR = {torch.float32: "<f4", torch.int32: "<i4"}

    def __init__(self, shape, dtype):
        self.tensor = torch.empty(shape, dtype=dtype, pin_memory=True)
        self.__array_interface__ = {"data": (self.tensor.data_ptr(), False), "shape": tuple(shape),
                                    "typestr": self._TYPESTR[dtype], "version": 3}

    def array(self) -> np.ndarray:
        return np.asarray(self)


class _PinnedPool:
    """Pinned host staging buffers for get_pred_result.  A buffer is handed out again only when nothing outside
    the pool references it any more: the numpy arrays returned to the caller (and every view of them, e.g. the
    per-sample rows src/utils/evaluator.py:74-86 keeps) keep their buffer referenced, so results stay valid
    for as long as the caller holds them."""

    def __init__(self):
        self._bufs = []

    def get(self, shape, dtype) -> _PinnedBuffer:
        shape = tuple(shape)
        for b in self._bufs:
            # references when free: the list, the loop variable, getrefcount's argument
            if tuple(b.tensor.shape) == shape and b.tensor.dtype == dtype and sys.getrefcount(b) <= 3:
                return b
        b = _PinnedBuffer(shape, dtype)
        self._bufs.append(b)
        return b

    def trim(self):
        """Drops the buffers nobody holds (frees their pinned memory)."""
        self._bufs = [b for b in self._bufs if sys.getrefcount(b) > 3]


class PrefetchedInput:
    """Device copies of one batch issued on the H2D stream (OptimizeModel.prefetch_input)."""

    def __init__(self, tensors, event):
        self.tensors, self.event = tensors, event


class PendingResult:
    """Result of get_pred_result_async: the D2H copies are in flight on the D2H stream."""

    def __init__(self, arrays, event, keep):
        self._arrays, self._event, self._keep = arrays, event, keep

    def wait(self):
        """Blocks until THIS batch's copies are done (not the device) and returns the 13 arrays."""
        if self._event is not None:
            self._event.synchronize()
            self._event, self._keep = None, None
        return self._arrays


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
        # conventions of the un-vendored `sdf` package the oracle had to assume; the
        # reference call site (loss_utils.py:181-182) uses the package defaults
        self._model.set_sdf_conventions(getattr(opt, "sdf_scale_factor", 0.2), getattr(opt, "sdf_ray_axis", 0))

        self.strategy = strategies[opt.strategy] if isinstance(opt.strategy, str) else opt.strategy
        self.default_loss_weights = dict(DEFAULT_LOSS_WEIGHTS)
        assert abs(self.default_loss_weights["collision_loss_weight"] - 1.0) < 1e-7
        self._ws = None
        self._buf: Dict[str, torch.Tensor] = {}
        self._pinned = _PinnedPool()
        self._h2d_stream = torch.cuda.Stream(self.device)
        self._d2h_stream = torch.cuda.Stream(self.device)
        self._d2h_done = None        # event of the last result copy: the next final forward must not overtake it
        self._d2h_small = None       # event after its first part (parameters, pass-through inputs): next set_input / init
        self._dev_in: Dict[str, torch.Tensor] = {}
        self.params = None
        # CUDA graphs of the stage calls (launch-bound at small batches: ~2000 launches per step)
        self.use_cuda_graphs = bool(getattr(opt, "use_cuda_graphs", True))
        self._graphs: Dict[tuple, tuple] = {}
        self.replayed_launches = 0   # kernels launched through graph replays (ihmr_launch_count only sees direct launches)

    def _build_device_model(self):
        from .mano_layer import DeviceModel
        right, left = self.mano_models["right"], self.mano_models["left"]
        arrays = dict(right._arrays)
        arrays["shapedirs"] = right.shapedirs.detach().cpu().numpy()
        return DeviceModel(arrays, right.faces, left.faces, self.device.index or 0)

    # -------------------------------------------------------------------------- input
    def prefetch_input(self, input) -> PrefetchedInput:
        """Starts the H2D copy of a batch on the model's copy stream and returns at once; pass the handle to
        set_input later.  With pinned source tensors the copy overlaps whatever the compute stream is doing."""
        dev = self.device
        out = {}
        with torch.cuda.stream(self._h2d_stream):
            for key in INPUT_KEYS:
                t = input[key]
                t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
                t = t.to(torch.float32)
                # pageable source: torch stages it synchronously; pinned sources go straight to the device
                out[key] = t.to(dev, non_blocking=t.device.type != "cpu" or t.is_pinned()).contiguous()
            ev = torch.cuda.Event()
            ev.record(self._h2d_stream)
        return PrefetchedInput(out, ev)

    def set_input(self, input):
        """H2D copy of one batch (keys of src/data/opt_dataset.py:176-196), or adoption of a prefetched one."""
        pre = input if isinstance(input, PrefetchedInput) else self.prefetch_input(input)
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(pre.event)
        if self._d2h_small is not None:      # the previous batch's result copy still reads gt_joints_3d / mano_params_weight
            cur.wait_event(self._d2h_small)
        # The batch lives in persistent device buffers (same addresses for every batch, so captured CUDA graphs of the
        # stages stay valid); the prefetched copy is moved in with device-to-device copies on the compute stream.
        t = {}
        for key, v in pre.tensors.items():
            buf = self._dev_in.get(key)
            if buf is None or buf.shape != v.shape:
                buf = torch.empty_like(v)
                self._dev_in[key] = buf
                self._graphs.clear()
            buf.copy_(v, non_blocking=True)
            v.record_stream(cur)         # allocated on the copy stream, consumed on the compute stream
            t[key] = buf
        self.hand_type_array, self.hand_type_valid = t["hand_type_array"], t["hand_type_valid"]
        self.joints_2d, self.joints_3d = t["joints_2d"], t["joints_3d"]
        self.hand_trans = t["hand_trans"]
        self.gt_pose_params, self.gt_shape_params = t["mano_pose"], t["mano_betas"]
        self.mano_params_weight = t["mano_params_weight"]
        self.init_cam, self.init_pose_params = t["init_cam"], t["init_pose_params"]
        self.init_shape_params, self.init_hand_trans = t["init_shape_params"], t["init_hand_trans"]
        self.init_joints_2d, self.init_joints_3d = t["init_joints_2d"], t["init_joints_3d"]
        self.init_hand_trans_j = t["init_hand_trans_j"]
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
        if self._d2h_small is not None:      # the previous batch's result copy still reads the parameter matrix
            torch.cuda.current_stream(self.device).wait_event(self._d2h_small)
            self._d2h_small = None
        if getattr(self, "params", None) is None or self.params.shape != (self.batch_size, 122):
            self.params = torch.empty(self.batch_size, 122, device=self.device, dtype=torch.float32)
            self._graphs.clear()
        p = self.params
        p[:, 0:3].copy_(self.init_cam)
        p[:, 3:6].copy_(self.init_hand_trans[:, 0, :3])
        p[:, 6:102].copy_(self.init_pose_params)
        p[:, 102:122].copy_(self.init_shape_params)

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

    def _replay(self, key, fn):
        """Runs fn() (stream-ordered C-ABI launches only, no allocation) directly the first time and through a CUDA
        graph captured right after it from then on.  `key` must change whenever a pointer or argument of fn does."""
        if not self.use_cuda_graphs:
            fn()
            return
        hit = self._graphs.get(key)
        if hit is not None:
            hit[0].replay()
            self.replayed_launches += hit[1]
            return
        fn()                                  # this batch: direct launches (also does the one-time kernel attribute setup)
        g = torch.cuda.CUDAGraph()
        n0 = self.lib.ihmr_launch_count()
        with torch.cuda.graph(g, stream=self._capture_stream()):
            fn()                              # recorded, not executed
        self._graphs[key] = (g, int(self.lib.ihmr_launch_count() - n0))

    def _capture_stream(self):
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(self.device)
        return self._cap_stream

    def _graph_key(self, tag, stage=None):
        ptrs = (self.params.data_ptr(), self._workspace().data_ptr()) + tuple(t.data_ptr() for t in self._dev_in.values())
        extra = bytes(_lib.make_stage(stage)) if stage is not None else b""
        return (tag, self.batch_size, self.bs_norm, int(self.opt.save_mid_freq), getattr(self.opt, "optimizer", "adam"), ptrs, extra)

    def forward(self):
        """Final-style forward: fills pred_*_hand_verts, pred_joints_3d (root aligned, as the
        reference leaves it after __compute_loss), collision outputs."""
        B, ws = self.batch_size, self._workspace()
        if self._d2h_done is not None:       # the previous batch's result copy reads the buffers written below
            torch.cuda.current_stream(self.device).wait_event(self._d2h_done)
            self._d2h_done = None
        self.pred_right_hand_verts = self._out("rv", B, 778, 3)
        self.pred_left_hand_verts = self._out("lv", B, 778, 3)
        self.pred_joints_3d = self._out("j3d", B, 42, 3)
        self.collision_loss_batch = self._out("col", B)
        self.collision_loss_origin_scale = self._out("ori", B, 1556)
        self.joints_3d_loss_p_batch = self._out("j3dp", B)

        def launch():
            _lib.check(self.lib.ihmr_opt_final(self._model.handle, B, _ptr(self.params), C.byref(self._targets),
                                               _ptr(self.pred_right_hand_verts), _ptr(self.pred_left_hand_verts),
                                               _ptr(self.pred_joints_3d), _ptr(self.collision_loss_batch),
                                               _ptr(self.collision_loss_origin_scale), _ptr(self.joints_3d_loss_p_batch),
                                               _ptr(ws), ws.numel(), _stream(self.device)), "ihmr_opt_final")
        self._replay(self._graph_key("final") + (self.pred_right_hand_verts.data_ptr(), self.collision_loss_origin_scale.data_ptr()), launch)

    def optimize(self, iter_id=0, num_iter=1):
        for stage_id, stage in enumerate(self.strategy):
            self._replay(self._graph_key(("stage", stage_id), stage), lambda: self.run_stage(stage))
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
    def get_pred_result_async(self) -> PendingResult:
        """Starts the D2H copies of the 13 result arrays (optimize_model.py:418-435) on the model's copy stream,
        ordered after everything enqueued so far on the current stream, and returns without waiting."""
        B = self.batch_size
        cur = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        src = OrderedDict(
            pred_cam_params=self.pred_cam_params, pred_hand_trans=self.pred_hand_trans,
            pred_shape_params=self.pred_shape_params, pred_pose_params=self.pred_pose_params,
            pred_right_hand_verts=self.pred_right_hand_verts, pred_left_hand_verts=self.pred_left_hand_verts,
            mano_params_weight=self.mano_params_weight, pred_joints_3d=self.pred_joints_3d,
            gt_joints_3d=self.joints_3d, collision_loss=self.collision_loss_batch,
            collision_loss_origin_scale=self.collision_loss_origin_scale)
        res, keep = OrderedDict(), []
        # what the NEXT batch overwrites first (parameter matrix, pass-through inputs) is copied first and gets its own
        # event; the large arrays are only overwritten by the next batch's final forward
        early = ("pred_cam_params", "pred_hand_trans", "pred_shape_params", "pred_pose_params", "mano_params_weight", "gt_joints_3d")
        order = [k for k in src if k in early] + [k for k in src if k not in early]
        bufs = {}
        with torch.cuda.stream(self._d2h_stream):
            self._d2h_stream.wait_event(ready)
            small = None
            for i, name in enumerate(order):
                t = src[name].detach()
                buf = self._pinned.get(t.shape, t.dtype)
                buf.tensor.copy_(t, non_blocking=True)     # strided views (parameter columns) are gathered by the copy
                t.record_stream(self._d2h_stream)
                keep.append(t)
                bufs[name] = buf.array()
                del buf
                if i == len(early) - 1:
                    small = torch.cuda.Event()
                    small.record(self._d2h_stream)
            done = torch.cuda.Event()
            done.record(self._d2h_stream)
        for name in src:
            res[name] = bufs[name]
        self._d2h_small = small
        res["do_flip"] = np.zeros(B).astype(np.int32)
        res["pred_hand_type"] = np.ones(B).astype(np.int32)
        self._d2h_done = done
        return PendingResult(res, done, keep)

    def get_pred_result(self):
        """optimize_model.py:418-435: the 13 numpy arrays the evaluator consumes.  The arrays own their (pinned)
        memory: they stay valid after later calls, as the reference's freshly allocated arrays do.  Waits for this
        batch's copies only."""
        return self.get_pred_result_async().wait()

    def run_pipelined(self, batches, iter_id=0, num_iter=1):
        """Generator over an iterable of input dicts (the reference's per-batch loop, src/optimize.py:61-73, with
        the copies taken off the critical path): yields each batch's result dict in order.  While batch k
        refines, batch k+1 is copied to the device and batch k-1's results are copied back, on separate
        streams; the host runs at most one batch ahead of the results it hands out."""
        it = iter(batches)
        first = next(it, None)
        nxt = self.prefetch_input(first) if first is not None else None
        pending = None
        while nxt is not None:
            self.set_input(nxt)
            self.init_optimize()
            self.optimize(iter_id, num_iter)
            res = self.get_pred_result_async()
            upcoming = next(it, None)
            nxt = self.prefetch_input(upcoming) if upcoming is not None else None
            if pending is not None:
                yield pending.wait()
            pending = res
        if pending is not None:
            yield pending.wait()

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
        """optimize_model.py:438-455: the GT-based log values of __compute_loss (:276-306) plus the two criteria.  Log
        only — plain tensor arithmetic on the exported outputs, not part of the refinement loop.

        The exported ``pred_joints_3d`` were root aligned in place twice (loss_utils.py:91-104): by the GT wrist
        weights, then by the prior's (always 1, joint 0), so they are relative to the right wrist whatever the GT
        says.  The GT-based 3-D value needs the alignment of the first rule only: right wrist present (w > 0.5) ->
        joint 0, absent (w < 1e-7) -> joint 21, anything between -> not aligned at all."""
        losses, _ = self.value_and_grad(dict(update_params=[], loss_weights=self.default_loss_weights, lr=0.0,
                                             epoch=0, filter_loss=[("joints_3d_loss_p", "+0")],
                                             select_loss="joints_3d_loss_p"))
        l = losses.cpu().numpy()
        n = float(self.bs_norm)
        with torch.no_grad():
            root = self.mano_models["right"](global_orient=self.pred_right_orient.contiguous(),
                                             hand_pose=self.pred_right_pose_params.contiguous(),
                                             betas=self.pred_right_shape_params.contiguous()).joints[:, 0:1]
            world = self.pred_joints_3d + root                      # the joints before any alignment
            cam = self.pred_cam_params.reshape(-1, 1, 3)
            j2d = cam[:, :, 0:1] * (world[:, :, :2] + cam[:, :, 1:])
            d2 = (self.joints_2d[:, :, :2] - j2d).abs() * self.joints_2d[:, :, 2:3]
            w0 = self.joints_3d[:, 0, 3].view(-1, 1, 1)
            has_r, no_r = (w0 > 0.5).float(), (w0 < 1e-7).float()

            def align(j):
                return j - j[:, 0:1] * has_r - j[:, 21:22] * no_r
            d3 = (align(self.joints_3d[:, :, :3]) - align(world)) ** 2 * self.joints_3d[:, :, 3:4]
            dt = (self.hand_trans[:, :, :3] - self.pred_hand_trans) ** 2 * self.hand_trans[:, :, 3:4]
        return OrderedDict([
            ("joints_2d_loss", float(d2.sum().item()) / (n * 84)),
            ("joints_3d_loss", float(d3.sum().item()) / (n * 126) * 1000),
            ("hand_trans_loss", float(dt.sum().item()) / (n * 3) * 10),
            ("collision_loss", float(l[3])),
            ("joints_3d_loss_p", float(l[1])),
        ])
