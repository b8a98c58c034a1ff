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

    def __init__(self, data_list=None):
        """``data_list``: the dataset's per-sample records (``OPTDataset.data_list``).  When given, ``update`` reads
        ``scale`` and ``hand_type`` of every sample from it as the reference does (evaluator.py:48-58) and frames are
        de-duplicated by ``img_path`` like ``remove_redunc`` there; without it the caller passes them (defaults:
        scale 1, every frame interacting)."""
        self.data_list = data_list
        self.tables: List[np.ndarray] = []
        self.interacting: List[np.ndarray] = []
        self.indices: List[np.ndarray] = []

    def clear(self):
        self.tables, self.interacting, self.indices = [], [], []

    def update(self, data_idxs, model, scale=None, interacting=None):
        """``model``: an ihmr_b200 OptimizeModel after ``optimize()``. ``interacting`` (B,) bool marks the
        frames whose hand_type is 'interacting' (the reference's default, evaluator.py:55-58)."""
        idx = np.asarray(data_idxs).reshape(-1)
        if self.data_list is not None:
            recs = [self.data_list[int(i)] for i in idx]
            if scale is None:
                scale = torch.tensor([float(r.get("scale", 1.0)) for r in recs], dtype=torch.float32)
            if interacting is None:
                interacting = np.array([r.get("hand_type", "interacting") == "interacting" for r in recs], bool)
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
        key = idx
        if self.data_list is not None:           # the reference keys on img_path_relative (evaluator.py:137-146)
            names = {}
            key = np.array([names.setdefault(self.data_list[int(i)]["img_path"], len(names)) for i in idx])
        _, first = np.unique(key, return_index=True)
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


# ------------------------------------------------------------------------------------------------
# Host-side records in the reference's own format (result I/O)
def single_joints_error(pred, gt_xyz, valid, scale):
    """metric_utils.py:23-38: per-hand root-relative joint errors of the hands whose wrist is annotated (the second
    hand is measured after BOTH skeletons were already moved to the first wrist, as the reference does in place)."""
    p, g = np.array(pred, copy=True), np.array(gt_xyz, copy=True)
    errors = []
    for i in (0, 21):
        if valid[i, 0] > 0:
            p -= p[i:i + 1, :]
            g -= g[i:i + 1, :]
            for j in range(21):
                if valid[i + j, 0] > 0:
                    errors.append(np.linalg.norm(p[i + j] - g[i + j]) / scale)
    return errors


def single_pa_no_rot_error(pred, gt_xyz, valid, scale):
    """metric_utils.py:107-143 with use_rot=False: align mean and per-axis spread of the valid joints, no rotation."""
    v = valid[:, 0] if valid.ndim == 2 else valid
    if np.sum(v) < 2.0:
        return []
    s1, s2 = np.array(pred)[v > 0, :3], np.array(gt_xyz)[v > 0, :3]
    s1n = (s1 - np.mean(s1, axis=0).reshape(1, 3)) / np.std(s1, axis=0).reshape(1, 3)
    moved = s1n * np.std(s2, axis=0).reshape(1, 3) + np.mean(s2, axis=0).reshape(1, 3)
    return (np.linalg.norm(moved - s2, axis=1) / scale).tolist()


class Evaluator:
    """Same constructor, attributes, ``update`` records and summary properties as the reference's ``Evaluator``
    (/root/reference/src/utils/evaluator.py:20-181) for the refinement driver; ``save`` writes the pickle the reference's
    own tooling opens (``evaluate_results/optimize/<dataset>.pkl``, src/optimize.py:91-96)."""
    RECORD_DEFAULTS = dict(annot_type="machine", hand_type="interacting", hand_type_valid=1.0, scale=1.0)

    def __init__(self, opt, test_dataset, model):
        self.dataset_name = test_dataset.name
        self.data_list = test_dataset.data_list
        self.image_root = test_dataset.image_root
        self.inputSize = model.inputSize
        self.left_hand_faces = model.mano_models["left"].faces
        self.right_hand_faces = model.mano_models["right"].faces
        self.pred_results = list()

    def gather_pred(self, pred_results):
        self.pred_results += pred_results

    def clear(self):
        self.pred_results = list()

    def update(self, data_idxs, pred_results, save_verts=True):
        import os.path as osp
        self.save_verts = save_verts
        for i, data_idx in enumerate(data_idxs):
            data_idx = int(data_idx)
            src = self.data_list[data_idx]
            rec = dict(data_idx=data_idx)
            for k in ("pred_cam_params", "pred_shape_params", "pred_pose_params", "pred_hand_trans", "pred_joints_3d",
                      "collision_loss_origin_scale", "gt_joints_3d"):
                rec[k] = np.array(pred_results[k][i])        # owned copies: the flip below edits them in place
            rec["img_path"] = osp.join(self.image_root, src["img_path"])
            rec["img_path_relative"] = src["img_path"]
            for k, dflt in self.RECORD_DEFAULTS.items():
                rec[k] = src[k] if k in src else dflt
            if save_verts:
                for mode in ("pred", "gt"):
                    for hand in ("left", "right"):
                        k = f"{mode}_{hand}_hand_verts"
                        if k in pred_results:
                            rec[k] = pred_results[k][i].astype(np.float16)
            gt, valid = rec["gt_joints_3d"][:, :3], rec["gt_joints_3d"][:, 3:]
            rec["j3d_error"] = single_joints_error(rec["pred_joints_3d"], gt, valid, rec["scale"])
            rec["pa_no_rot_inter_j3d_error"] = single_pa_no_rot_error(rec["pred_joints_3d"], gt, valid, rec["scale"])
            if pred_results["do_flip"][i]:
                self._flip_back(rec)
            self.pred_results.append(rec)

    def _flip_back(self, rec):
        """evaluator.py:99-134: undo the left->right mirroring of a flipped sample."""
        rec["pred_cam_params"][1] *= -1
        rec["pred_hand_trans"][0] *= -1
        pose = rec["pred_pose_params"].copy()
        rec["pred_pose_params"][:48], rec["pred_pose_params"][48:] = pose[48:], pose[:48]
        rec["pred_pose_params"][1::3] *= -1
        rec["pred_pose_params"][2::3] *= -1
        for k in ("pred_joints_3d", "gt_joints_3d"):
            j = rec[k].copy()
            rec[k][:21], rec[k][21:] = j[21:], j[:21]
            rec[k][:, 0] *= -1
        c = rec["collision_loss_origin_scale"].copy()
        rec["collision_loss_origin_scale"][:778], rec["collision_loss_origin_scale"][778:] = c[778:], c[:778]
        if self.save_verts:
            saved = {k: rec[k].copy() for k in rec if k.endswith("_hand_verts")}
            for k, v in saved.items():
                other = k.replace("left", "@").replace("right", "left").replace("@", "right")
                if other in saved:
                    rec[k] = saved[other]
                    rec[k][:, 0] *= -1

    def remove_redunc(self):
        """evaluator.py:137-146: the dataset is padded to full batches; keep the first record per image."""
        seen, out = set(), []
        for rec in self.pred_results:
            if rec["img_path_relative"] not in seen:
                out.append(rec)
                seen.add(rec["img_path_relative"])
        self.pred_results = out

    @property
    def mpjpe_3d(self):
        return np.average([e for r in self.pred_results for e in r["j3d_error"]])

    @property
    def inter_mpjpe_3d(self):
        return np.average([e for r in self.pred_results for e in r["pa_no_rot_inter_j3d_error"]])

    @property
    def collision_ave(self):
        return np.average([np.mean(r["collision_loss_origin_scale"]) * 1000 for r in self.pred_results
                           if r["hand_type"] == "interacting"])

    @property
    def collision_max(self):
        return np.average([np.max(r["collision_loss_origin_scale"]) * 1000 for r in self.pred_results
                           if r["hand_type"] == "interacting"])

    # ---- pickle in the reference's format: an instance of `utils.evaluator.Evaluator`
    def save(self, path):
        """Writes the file src/optimize.py:91-96 writes: a pickled ``utils.evaluator.Evaluator`` whose attribute
        dictionary is this object's.  Where the reference's class is importable the file opens as that class; here a
        stand-in of the same qualified name is registered only while dumping."""
        import os
        import pickle
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with _reference_evaluator_class() as cls:
            obj = cls.__new__(cls)
            obj.__dict__.update(self.__dict__)
            with open(path, "wb") as fh:
                pickle.dump(obj, fh, protocol=2)

    @classmethod
    def load(cls, path):
        import pickle
        with _reference_evaluator_class():
            with open(path, "rb") as fh:
                obj = pickle.load(fh)
        out = cls.__new__(cls)
        out.__dict__.update(obj.__dict__)
        return out


class _reference_evaluator_class:
    """Context: makes `utils.evaluator.Evaluator` resolvable (the real class if the reference is importable,
    otherwise a bare stand-in that is removed again on exit)."""

    def __enter__(self):
        import importlib
        import sys
        import types
        self._added = []
        try:
            return importlib.import_module("utils.evaluator").Evaluator
        except Exception:
            pass
        for name in ("utils", "utils.evaluator"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
                self._added.append(name)
        mod = sys.modules["utils.evaluator"]
        if not hasattr(mod, "Evaluator"):
            mod.Evaluator = type("Evaluator", (object,), {"__module__": "utils.evaluator"})
            self._added.append("utils.evaluator:Evaluator")
        return mod.Evaluator

    def __exit__(self, *exc):
        import sys
        for name in reversed(self._added):
            if name.endswith(":Evaluator"):
                if "utils.evaluator" in sys.modules and hasattr(sys.modules["utils.evaluator"], "Evaluator"):
                    delattr(sys.modules["utils.evaluator"], "Evaluator")
            else:
                sys.modules.pop(name, None)
        return False
