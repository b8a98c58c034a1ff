"""Device-side counterpart of the reference ``Evaluator`` for the IHMR-OPT path
(/root/reference/src/utils/evaluator.py:20-181, metric code in src/utils/metric_utils.py).

The reference copies 26 KB per frame to the host (``get_pred_result``) and computes the four
numbers ``src/optimize.py:99-102`` prints — ``mpjpe_3d``, ``inter_mpjpe_3d``, ``collision_ave``,
``collision_max`` — with numpy loops.  Here one kernel (``ihmr_eval_metrics``) reduces every frame to
6 floats on the device; only that (frames, 6) table crosses to the host, where the final averages are
taken in float64 with the same pooling as the reference (errors of all frames are pooled, so frames
with more valid joints weigh more; collision values are averaged per frame).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from .mano_layer import _f32c, _ptr, _stream


def frame_metrics(pred_joints_3d: torch.Tensor, gt_joints_3d: torch.Tensor, collision_origin_scale: torch.Tensor,
                  scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(B,42,3), (B,42,4), (B,1556)[, (B,)] CUDA tensors -> (B,6) table, see include/ihmr_b200.h."""
    p, g, o = _f32c(pred_joints_3d), _f32c(gt_joints_3d), _f32c(collision_origin_scale)
    B = p.shape[0]
    if p.shape != (B, 42, 3) or g.shape != (B, 42, 4) or o.shape != (B, 1556):
        raise ValueError("expected pred (B,42,3), gt (B,42,4), collision_origin_scale (B,1556)")
    if p.device.type != "cuda":
        raise _lib.IhmrError("ihmr_b200 has no CPU path: metrics inputs must live on a CUDA (sm_100) device")
    s = None if scale is None else _f32c(scale.to(p.device))
    out = torch.empty(B, 6, device=p.device, dtype=torch.float32)
    _lib.check(_lib.load().ihmr_eval_metrics(B, _ptr(p), _ptr(g), _ptr(o), _ptr(s), _ptr(out), _stream(p.device)),
               "ihmr_eval_metrics")
    return out


class DeviceEvaluator:
    """``update`` after every batch, read the four properties at the end (same names as the reference)."""

    def __init__(self):
        self.tables: List[np.ndarray] = []
        self.interacting: List[np.ndarray] = []
        self.indices: List[np.ndarray] = []

    def clear(self):
        self.tables, self.interacting, self.indices = [], [], []

    def update(self, data_idxs, model, scale=None, interacting=None):
        """``model``: an ihmr_b200 OptimizeModel after ``optimize()``. ``interacting`` (B,) bool marks the
        frames whose hand_type is 'interacting' (the reference's default, evaluator.py:55-58)."""
        t = frame_metrics(model.pred_joints_3d, model.joints_3d, model.collision_loss_origin_scale, scale)
        self.tables.append(t.cpu().numpy().astype(np.float64))
        B = t.shape[0]
        self.interacting.append(np.ones(B, bool) if interacting is None else np.asarray(interacting, bool))
        self.indices.append(np.asarray(data_idxs).reshape(-1))

    def gather_pred(self, other: "DeviceEvaluator"):
        self.tables += other.tables
        self.interacting += other.interacting
        self.indices += other.indices

    def remove_redunc(self):
        """Drop repeated frame ids (the reference pads the dataset and de-duplicates, evaluator.py:137-146)."""
        idx = np.concatenate(self.indices)
        _, first = np.unique(idx, return_index=True)
        keep = np.sort(first)
        self.tables = [np.concatenate(self.tables)[keep]]
        self.interacting = [np.concatenate(self.interacting)[keep]]
        self.indices = [idx[keep]]

    def _table(self):
        return np.concatenate(self.tables), np.concatenate(self.interacting)

    @property
    def mpjpe_3d(self):
        t, _ = self._table()
        return float(t[:, 0].sum() / t[:, 1].sum())

    @property
    def inter_mpjpe_3d(self):
        t, _ = self._table()
        return float(t[:, 2].sum() / t[:, 3].sum())

    @property
    def collision_ave(self):
        t, m = self._table()
        return float(np.mean(t[m, 4] * 1000))

    @property
    def collision_max(self):
        t, m = self._table()
        return float(np.mean(t[m, 5] * 1000))
