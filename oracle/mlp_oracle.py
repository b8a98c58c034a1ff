"""ORACLE (test infrastructure, not product code) — CPU restatement of the IHMR-MLP test-time path.

Follows /root/reference/src/models/networks.py:83-105 (InterHandSubNetwork), src/models/mlp_model.py:458-472
(__update_params_single), :592-637 (select_better_params) and :683-699 (test), on top of the pinned host-loop
restatement (oracle/host_loop_oracle.py) for the MANO forward and the per-frame criteria.

PINNED where the reference can be imported (tests/test_mlp_oracle.py): the sub-network against the reference's own
``InterHandSubNetwork`` class with the same state dict (and through tests/golden/mlp.npz, generated from that class by
tests/golden/make_golden_mlp.py), and the selection rule against the reference's own unmodified
``MLPModel.select_better_params`` driven on a stand-in object.  The image backbone is outside this path.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn

PARAM_DIMS = {"pred_cam_params": 3, "pred_hand_trans": 3, "pred_right_orient": 3, "pred_right_pose_params": 45,
              "pred_left_orient": 3, "pred_left_pose_params": 45, "pred_right_shape_params": 10, "pred_left_shape_params": 10}
DEFAULT_WEIGHTS = dict(joints_2d_loss=10.0, joints_3d_loss=10.0, trans_loss_weight=0.0, collision_loss_weight=1.0,
                       shape_reg_loss_weight=0.0, finger_reg_loss_weight=0.0)          # mlp_model.py:219-229 (criteria only)


class SubNetworkOracle(nn.Module):
    """networks.py:83-105: 1146 -> 512 -> 256 -> 128 -> update dim with ReLU between; same parameter names."""

    def __init__(self, update_dim: int, input_dim: int = 1024 + 122):
        super().__init__()
        relu = nn.ReLU(inplace=True)
        self.regressor = nn.Sequential(nn.Linear(input_dim, 512), relu, nn.Linear(512, 256), relu, nn.Linear(256, 128), relu,
                                       nn.Linear(128, update_dim))

    def forward(self, x):
        return self.regressor(x)


def seeded_state_dict(update_dim: int, seed: int, scale: float = 0.04) -> Dict[str, torch.Tensor]:
    """Deterministic weights (numpy RandomState, independent of torch's initialisers) with the reference's parameter names;
    the golden generator loads them into the reference's class, the tests into the oracle and the CUDA path."""
    import numpy as np
    rng = np.random.RandomState(seed)
    dims = [(1024 + 122, 512), (512, 256), (256, 128), (128, update_dim)]
    sd = {}
    for i, (fi, fo) in enumerate(dims):
        sd[f"regressor.{2 * i}.weight"] = torch.from_numpy((rng.standard_normal((fo, fi)) * scale).astype(np.float32))
        sd[f"regressor.{2 * i}.bias"] = torch.from_numpy((rng.standard_normal(fo) * 0.05).astype(np.float32))
    return sd


def final_params(loop) -> torch.Tensor:
    """mlp_model.py:432-436: [cam | pose 96 | shape 20 | hand_trans 3]"""
    p = loop.p
    pose = torch.cat([p["pred_right_orient"], p["pred_right_pose_params"], p["pred_left_orient"], p["pred_left_pose_params"]], 1)
    shape = torch.cat([p["pred_right_shape_params"], p["pred_left_shape_params"]], 1)
    return torch.cat([loop.pred_cam_params, pose, shape, p["pred_hand_trans"].reshape(loop.B, 3)], 1)


def criteria(loop) -> Dict[str, torch.Tensor]:
    loop.forward()
    loop.compute_loss(DEFAULT_WEIGHTS)
    return dict(joints_3d_loss_p=loop.joints_3d_loss_p_batch.detach().clone(), collision_loss=loop.collision_loss_batch.detach().clone(),
                joints_2d_loss_p=loop.joints_2d_loss_p_batch.detach().clone())


def select_better(cur: Dict[str, torch.Tensor], prev: Dict[str, torch.Tensor], stage: dict) -> torch.Tensor:
    """mlp_model.py:596-611: keep (True) where every filter criterion is below its margin and the select criterion
    did not get worse."""
    ok = torch.ones_like(next(iter(cur.values())), dtype=torch.bool)
    for name, percent in stage["filter_loss"]:
        ok &= cur[name] < prev[name] * (1 + float(percent) / 100)
    ok &= cur[stage["select_loss"]] <= prev[stage["select_loss"]]
    return ok


def mlp_test(loop, nets: List[nn.Module], img_feat: torch.Tensor, batch, strategy: List[dict]):
    """mlp_model.py:683-699 on a HostLoopOracle: returns (result dict, list of kept masks)."""
    with torch.no_grad():
        loop.set_input(batch)
        loop.init_optimize()
        prev = criteria(loop)
        kept_all = []
        for stage, net in zip(strategy, nets):
            x = torch.cat([img_feat.to(loop.dtype), final_params(loop)], 1)
            res = net(x)
            old = {n: (loop.pred_cam_params if n == "pred_cam_params" else loop.p[n]).clone() for n in stage["update_params"]}
            off = 0
            for n in stage["update_params"]:                       # residual columns in list order (:462-470)
                d = PARAM_DIMS[n]
                new = old[n] + res[:, off:off + d].reshape(old[n].shape)
                off += d
                if n == "pred_cam_params":
                    loop.pred_cam_params = new
                else:
                    loop.p[n] = new
            cur = criteria(loop)
            keep = select_better(cur, prev, stage)
            for n in stage["update_params"]:
                tgt = loop.pred_cam_params if n == "pred_cam_params" else loop.p[n]
                m = keep.reshape((-1,) + (1,) * (tgt.dim() - 1))
                merged = torch.where(m, tgt, old[n])
                if n == "pred_cam_params":
                    loop.pred_cam_params = merged
                else:
                    loop.p[n] = merged
            prev = {k: torch.where(keep, cur[k], prev[k]) for k in prev}
            kept_all.append(keep)
        loop.forward()
        loop.compute_loss(dict(DEFAULT_WEIGHTS))
        return loop.get_pred_result(), kept_all, prev
