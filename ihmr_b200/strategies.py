"""Strategy registry with the dict layout of /root/reference/src/strategies/__init__.py:4-24:
a strategy is a list of stages; a stage holds ``update_params``, ``loss_weights``, ``lr``,
``epoch``, ``filter_loss`` [(criterion, '+p' | '-p'), ...] and ``select_loss``.

``opt_default`` carries the numbers of src/strategies/opt_default.py:1-78 (translation ->
global orientations -> finger poses -> shapes; the camera stage is commented out there).
"""
from __future__ import annotations

import copy
from typing import Dict, List

_FILTERS = [("joints_3d_loss_p", "+0"), ("collision_loss", "-10")]


def _stage(update_params, lr, joints_2d, trans, collision, finger, epoch=300) -> dict:
    return dict(
        update_params=list(update_params),
        loss_weights=dict(joints_2d_loss=joints_2d, joints_3d_loss=1000.0, trans_loss_weight=trans,
                          shape_reg_loss_weight=0.1, collision_loss_weight=collision,
                          finger_reg_loss_weight=finger),
        lr=lr, epoch=epoch, filter_loss=list(_FILTERS), select_loss="joints_3d_loss_p")


opt_default: List[dict] = [
    _stage(["pred_hand_trans"], 1e-4, joints_2d=100.0, trans=1000.0, collision=0.1, finger=0.0),
    _stage(["pred_left_orient", "pred_right_orient"], 1e-2, joints_2d=10.0, trans=100.0, collision=1.0, finger=0.0),
    _stage(["pred_left_pose_params", "pred_right_pose_params"], 1e-2, joints_2d=10.0, trans=100.0, collision=1.0,
           finger=100000.0),
    _stage(["pred_left_shape_params", "pred_right_shape_params"], 1e-2, joints_2d=10.0, trans=100.0, collision=1.0,
           finger=0.0),
]


def with_epochs(strategy: List[dict], epoch: int) -> List[dict]:
    """Copy of a strategy with every stage's epoch replaced (the fixed-iteration benchmark
    strategy of SURVEY.md §8(d) is ``with_epochs(opt_default, 24)``: 4 x 25 iterations)."""
    out = copy.deepcopy(strategy)
    for st in out:
        st["epoch"] = epoch
    return out


# IHMR-MLP (src/strategies/mlp_default.py): what the test-time path reads of each stage — which parameters the stage's
# residual MLP proposes, and the criteria select_better_params compares (mlp_model.py:592-637).  Training-only fields
# (loss weights, lr schedule, epochs) are not part of inference.
_MLP_F = [("joints_3d_loss_p", "+0"), ("collision_loss", "+0")]


def _mlp_stage(update_params, filter_loss=None, select_loss="collision_loss") -> dict:
    return dict(update_params=list(update_params), filter_loss=list(filter_loss or _MLP_F), select_loss=select_loss)


mlp_default: List[dict] = [
    _mlp_stage(["pred_hand_trans"]),
    _mlp_stage(["pred_left_orient"]),
    _mlp_stage(["pred_right_orient"]),
    _mlp_stage(["pred_left_pose_params", "pred_right_pose_params"]),
    _mlp_stage(["pred_left_shape_params", "pred_right_shape_params"]),
    _mlp_stage(["pred_cam_params"], [("joints_2d_loss_p", "+0")], "joints_2d_loss_p"),
]

strategies: Dict[str, List[dict]] = dict(opt_default=opt_default, opt_fixed100=with_epochs(opt_default, 24), mlp_default=mlp_default)
