"""Golden vectors for the evaluator metrics, from the reference's UNMODIFIED utils/metric_utils.py
(run once in the authoring container): python tests/golden/make_golden_metrics.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims as RS  # noqa: E402

RS._install_shims()
sys.path.insert(0, RS.REFERENCE_SRC)
import utils.metric_utils as mu  # noqa: E402  (the reference's file)

rng = np.random.default_rng(7)
B = 12
gt = np.concatenate([rng.normal(0, 0.08, (B, 42, 3)), np.ones((B, 42, 1))], 2).astype(np.float32)
pred = (gt[:, :, :3] + rng.normal(0, 0.01, (B, 42, 3))).astype(np.float32)
gt[1, 0, 3] = 0            # right wrist missing
gt[2, 21, 3] = 0           # left wrist missing
gt[3, 5:30, 3] = 0         # many joints missing
gt[4, :, 3] = 0
gt[4, 7, 3] = 1            # fewer than two valid joints, wrists missing
gt[5, :21, 3] = 0          # right hand missing
scale = np.ones(B, np.float32)
scale[6] = 1.3
origin = np.abs(rng.normal(0, 0.002, (B, 1556))).astype(np.float32) * (rng.random((B, 1556)) < 0.1)
origin = origin.astype(np.float32)
table = np.zeros((B, 4))
for b in range(B):
    e1 = mu.get_single_joints_error(pred[b], gt[b, :, :3], gt[b, :, 3:], scale[b])
    e2 = mu.get_single_pa_inter_joints_error(pred[b], gt[b, :, :3], gt[b, :, 3:], scale[b], use_rot=False)
    table[b] = [np.sum(e1), len(e1), np.sum(e2), len(e2)]
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "metrics.npz"),
                    pred=pred, gt=gt, scale=scale, origin=origin, table=table)
print(table)
