"""ORACLE (test infrastructure, not product code) — CPU restatement of the evaluator metrics.

Follows /root/reference/src/utils/metric_utils.py:23-38 (get_single_joints_error), :107-118
(calc_transform_no_rot), :120-143 (get_single_pa_inter_joints_error, use_rot=False) and the four
properties of /root/reference/src/utils/evaluator.py:149-181.

PINNED: tests/golden/metrics.npz holds outputs of the reference's own, unmodified metric_utils
functions (imported by tests/golden/make_golden_metrics.py through oracle/ref_shims.py) on seeded
inputs that exercise missing wrists, missing joints and fewer than two valid joints.
"""
from __future__ import annotations

import numpy as np


def joints_error(pred, gt_xyz, valid, scale):
    p, g = np.array(pred, dtype=np.float32), np.array(gt_xyz, dtype=np.float32)
    out = []
    for root in (0, 21):
        if valid[root, 0] > 0:
            p = p - p[root:root + 1]
            g = g - g[root:root + 1]
            sel = np.where(valid[root:root + 21, 0] > 0)[0] + root
            out += list(np.linalg.norm(p[sel] - g[sel], axis=1) / scale)
    return out


def pa_no_rot_error(pred, gt_xyz, valid, scale):
    v = valid[:, 0] if valid.ndim == 2 else valid
    if np.sum(v) < 2.0:
        return []
    s1, s2 = np.array(pred, dtype=np.float32)[v > 0, :3], np.array(gt_xyz, dtype=np.float32)[v > 0, :3]
    m1, m2 = s1.mean(0, keepdims=True), s2.mean(0, keepdims=True)
    d1, d2 = s1.std(0, keepdims=True), s2.std(0, keepdims=True)
    moved = (s1 - m1) / d1 * d2 + m2
    return list(np.linalg.norm(moved - s2, axis=1) / scale)


def frame_table(pred, gt, origin, scale=None):
    """The (B,6) table ihmr_eval_metrics produces, from the restated per-frame functions."""
    B = pred.shape[0]
    out = np.zeros((B, 6))
    for b in range(B):
        sc = 1.0 if scale is None else float(scale[b])
        e1 = joints_error(pred[b], gt[b, :, :3], gt[b, :, 3:], sc)
        e2 = pa_no_rot_error(pred[b], gt[b, :, :3], gt[b, :, 3:], sc)
        out[b] = [np.sum(e1), len(e1), np.sum(e2), len(e2), origin[b].mean(), origin[b].max()]
    return out


def summary(table, interacting=None):
    m = np.ones(len(table), bool) if interacting is None else np.asarray(interacting, bool)
    return dict(mpjpe_3d=table[:, 0].sum() / table[:, 1].sum(), inter_mpjpe_3d=table[:, 2].sum() / table[:, 3].sum(),
                collision_ave=np.mean(table[m, 4] * 1000), collision_max=np.mean(table[m, 5] * 1000))
