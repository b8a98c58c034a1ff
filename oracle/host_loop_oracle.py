"""ORACLE (test infrastructure, not product code) — restatement of the IHMR-OPT host loop.

The reference's loop is Python and runs here unmodified (oracle/ref_shims.py), but
/root/reference does not travel to the GPU box, so the CPU baseline and the GPU-side
drop-in tests need a restatement that does.  This file follows, function by function:

* OptimizeModel.set_input / init_optimize / get_mano_output / forward / __compute_loss /
  __set_optimize_target / __save_mid_results / __update_stage_results / optimize /
  get_pred_result           — /root/reference/src/models/optimize_model.py:120-435
* LossUtil._joints_2d_loss, __align_by_root, _joints_3d_loss, _hand_trans_loss,
  _shape_reg_loss, _finger_reg_loss, _collision_loss
                            — /root/reference/src/models/loss_utils.py:82-192
* batch_orthogonal_project  — /root/reference/src/models/transform_utils.py:47-53
* gather_params_losses, filter_by_losses, select_params
                            — /root/reference/src/utils/opt_utils.py:70-152
* opt_default               — /root/reference/src/strategies/opt_default.py:1-78

PINNED: tests/test_host_loop_oracle.py runs this against the unmodified reference loop
(same leaves, same inputs) wherever /root/reference exists, and against the committed
golden fixtures everywhere else.  The two leaves it drives (MANO layer, SDFLoss) are passed
in, so the same loop can drive the oracle leaves on CPU or the CUDA leaves on a GPU.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List

import numpy as np
import torch

TIP_IDS = (744, 320, 443, 554, 671)                       # optimize_model.py:99
FINGER_CHAINS = ((1, 2, 3, 17), (4, 5, 6, 18), (7, 8, 9, 20), (10, 11, 12, 19), (13, 14, 15, 16))

DEFAULT_LOSS_WEIGHTS = dict(joints_2d_loss=10.0, joints_3d_loss=1000.0, trans_loss_weight=100.0,
                            shape_reg_loss_weight=0.1, collision_loss_weight=1.0,
                            finger_reg_loss_weight=100000.0)  # optimize_model.py:84-92

PARAM_NAMES = ("pred_hand_trans", "pred_right_orient", "pred_left_orient", "pred_right_pose_params",
               "pred_left_pose_params", "pred_right_shape_params", "pred_left_shape_params")


def opt_default_strategy(epoch: int = 300) -> List[dict]:
    """The four stages of strategies/opt_default.py:1-78 (cam stage is commented out there)."""
    def stage(params, lr, j2d, trans, col, finger):
        return dict(update_params=params, lr=lr, epoch=epoch,
                    loss_weights=dict(joints_2d_loss=j2d, joints_3d_loss=1000.0, trans_loss_weight=trans,
                                      shape_reg_loss_weight=0.1, collision_loss_weight=col,
                                      finger_reg_loss_weight=finger),
                    filter_loss=[("joints_3d_loss_p", "+0"), ("collision_loss", "-10")],
                    select_loss="joints_3d_loss_p")
    return [
        stage(["pred_hand_trans"], 1e-4, 100.0, 1000.0, 0.1, 0.0),
        stage(["pred_left_orient", "pred_right_orient"], 1e-2, 10.0, 100.0, 1.0, 0.0),
        stage(["pred_left_pose_params", "pred_right_pose_params"], 1e-2, 10.0, 100.0, 1.0, 100000.0),
        stage(["pred_left_shape_params", "pred_right_shape_params"], 1e-2, 10.0, 100.0, 1.0, 0.0),
    ]


INVALID_CRITERIA = ("joints_3d_loss", "joints_2d_loss", "hand_trans_loss")   # opt_utils.py:57-67


def align_by_root(j3d: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """loss_utils.py:90-98, functional form: rows whose right-wrist weight > 0.5 become
    relative to joint 0, rows whose weight < 1e-7 relative to joint 21."""
    w0 = w[:, 0, 0]
    has_r = (w0 > 0.5).view(-1, 1, 1)
    j3d = torch.where(has_r, j3d - j3d[:, 0:1], j3d)
    no_r = (w0 < 1e-7).view(-1, 1, 1)
    return torch.where(no_r, j3d - j3d[:, 21:22], j3d)


def finger_reg(j3d: torch.Tensor):
    """loss_utils.py:138-171 (computed on whatever device j3d lives on)."""
    B = j3d.shape[0]
    idx = torch.tensor([c for base in (0, 21) for ch in FINGER_CHAINS for c in (np.array(ch) + base)],
                       device=j3d.device)
    ch = j3d[:, idx].view(B * 10, 4, 3)
    b0, b1, b2 = ch[:, 0] - ch[:, 1], ch[:, 1] - ch[:, 2], ch[:, 2] - ch[:, 3]
    c01 = torch.linalg.cross(b0, b1, dim=1)
    C1 = (b2 * c01).sum(1)
    C2 = (c01 * torch.linalg.cross(b1, b2, dim=1)).sum(1)
    per = (C1.abs() - torch.clamp(C2, max=0.0)).view(B, 10).sum(1)
    return per.mean(), per


class HostLoopOracle:
    def __init__(self, mano_right, faces_right, faces_left, sdf_loss, batch_size: int,
                 strategy: List[dict] = None, save_mid_freq: int = 1, optimizer: str = "adam",
                 device="cpu", dtype=torch.float32, bs_norm: int = None):
        self.mano = mano_right
        self.faces = dict(right=faces_right, left=faces_left)
        self.sdf_loss = sdf_loss
        self.B = batch_size
        # every batch mean divides by the batch size; bs_norm lets a shard of a larger batch
        # reproduce the larger batch's numbers
        self.bs_norm = batch_size if bs_norm is None else bs_norm
        self.strategy = strategy if strategy is not None else opt_default_strategy()
        self.save_mid_freq = save_mid_freq
        self.optimizer_name = optimizer
        self.device, self.dtype = device, dtype
        self.tip_ids = torch.tensor(TIP_IDS, device=device)

    # ------------------------------------------------------------------ input / state
    def set_input(self, data: Dict[str, torch.Tensor]):
        g = lambda k: data[k].to(self.device, self.dtype)
        self.hand_type_array = g("hand_type_array")
        self.joints_2d, self.joints_3d = g("joints_2d"), g("joints_3d")
        self.hand_trans = g("hand_trans")
        self.mano_params_weight = g("mano_params_weight")
        self.init_cam, self.init_pose_params = g("init_cam"), g("init_pose_params")
        self.init_shape_params, self.init_hand_trans = g("init_shape_params"), g("init_hand_trans")
        self.init_joints_2d, self.init_joints_3d = g("init_joints_2d"), g("init_joints_3d")
        self.init_hand_trans_j = g("init_hand_trans_j")

    def init_optimize(self):
        pose, shape = self.init_pose_params, self.init_shape_params
        self.p = dict(
            pred_hand_trans=self.init_hand_trans[..., :3].clone(),
            pred_right_orient=pose[:, 0:3].clone(), pred_right_pose_params=pose[:, 3:48].clone(),
            pred_left_orient=pose[:, 48:51].clone(), pred_left_pose_params=pose[:, 51:96].clone(),
            pred_right_shape_params=shape[:, :10].clone(), pred_left_shape_params=shape[:, 10:].clone())
        self.pred_cam_params = self.init_cam.clone()

    # ------------------------------------------------------------------------ forward
    def forward(self):
        p, B = self.p, self.B
        mir = torch.tensor([1.0, -1.0, -1.0], device=self.device, dtype=self.dtype)
        l_or = p["pred_left_orient"] * mir
        l_po = (p["pred_left_pose_params"].reshape(B, 15, 3) * mir).reshape(B, 45)
        out = self.mano(global_orient=torch.cat([p["pred_right_orient"], l_or]),
                        hand_pose=torch.cat([p["pred_right_pose_params"], l_po]),
                        betas=torch.cat([p["pred_right_shape_params"], p["pred_left_shape_params"]]))
        verts = out.vertices
        joints = torch.cat([out.joints, verts.index_select(1, self.tip_ids)], dim=1)
        flip = torch.tensor([-1.0, 1.0, 1.0], device=self.device, dtype=self.dtype)
        rv, rj = verts[:B], joints[:B]
        lv, lj = verts[B:] * flip, joints[B:] * flip
        shift = p["pred_hand_trans"].view(B, 1, 3) + (rj[:, 0:1] - lj[:, 0:1])
        self.pred_right_hand_verts, self.pred_left_hand_verts = rv, lv + shift
        self.pred_joints_3d = torch.cat([rj, lj + shift], dim=1)
        cam = self.pred_cam_params.view(B, 1, 3)
        self.pred_joints_2d = cam[:, :, 0:1] * (self.pred_joints_3d[:, :, :2] + cam[:, :, 1:])
        self.pred_shape_params = torch.cat([p["pred_right_shape_params"], p["pred_left_shape_params"]], 1)
        self.pred_pose_params = torch.cat([p["pred_right_orient"], p["pred_right_pose_params"],
                                           p["pred_left_orient"], p["pred_left_pose_params"]], 1)

    # -------------------------------------------------------------------------- losses
    def compute_loss(self, lw: Dict[str, float]):
        n = float(self.bs_norm)
        # 2-D reprojection (L1) against the prior prediction
        w2 = self.init_joints_2d[:, :, 2:3]
        l1 = (self.init_joints_2d[:, :, :2] - self.pred_joints_2d).abs() * w2
        self.joints_2d_loss_p_batch = l1.reshape(self.B, -1).mean(1) * lw["joints_2d_loss"]
        self.joints_2d_loss_p = l1.sum() / (n * 84) * lw["joints_2d_loss"]
        # GT-based log values: only their in-place root alignment of the prediction matters
        wgt = self.joints_3d[:, :, 3:4]
        pred = align_by_root(self.pred_joints_3d, wgt)
        gt = align_by_root(self.joints_3d[:, :, :3], wgt)
        self.joints_3d_loss = (((gt - pred) ** 2) * wgt).sum() / (n * 126) * 1000
        self.joints_2d_loss = ((self.joints_2d[:, :, :2] - self.pred_joints_2d).abs()
                               * self.joints_2d[:, :, 2:3]).sum() / (n * 84)
        # 3-D joints against the prior prediction
        w3 = self.init_joints_3d[:, :, 3:4]
        pred = align_by_root(pred, w3)
        tgt = align_by_root(self.init_joints_3d[:, :, :3], w3)
        sq = ((tgt - pred) ** 2) * w3
        self.joints_3d_loss_p_batch = sq.reshape(self.B, -1).mean(1) * lw["joints_3d_loss"]
        self.joints_3d_loss_p = sq.sum() / (n * 126) * lw["joints_3d_loss"]
        self.pred_joints_3d = pred                 # the reference aligns in place
        # relative translation
        d = self.init_hand_trans_j[:, :, :3] - self.p["pred_hand_trans"]
        self.hand_trans_loss_p = (d * d * self.init_hand_trans_j[:, :, 3:4]).sum() / (n * 3) * lw["trans_loss_weight"]
        dg = self.hand_trans[:, :, :3] - self.p["pred_hand_trans"]
        self.hand_trans_loss = (dg * dg * self.hand_trans[:, :, 3:4]).sum() / (n * 3) * 10
        # interpenetration
        hv = torch.stack([self.pred_right_hand_verts, self.pred_left_hand_verts], dim=1)
        col, _, origin = self.sdf_loss(hv, return_per_vert_loss=True, return_origin_scale_loss=True)
        both = (self.hand_type_array.sum(1) > 1.5).to(self.dtype)
        col = col.reshape(self.B) * both
        self.collision_loss_batch = col                                  # NOT weighted (:316-318)
        self.collision_loss_origin_scale = origin                       # NOT masked (loss_utils:189)
        self.collision_loss = col.sum() / n * lw["collision_loss_weight"]
        # shape regulariser and finger regulariser
        ds = self.p["pred_right_shape_params"] - self.p["pred_left_shape_params"]
        self.shape_reg_loss = (ds * ds).sum() / (n * 10) * lw["shape_reg_loss_weight"]
        fr, _ = finger_reg(self.pred_joints_3d)
        self.finger_reg_loss = fr * (self.B / n) * lw["finger_reg_loss_weight"]
        self.loss = (self.joints_2d_loss_p + self.joints_3d_loss_p + self.hand_trans_loss_p
                     + self.collision_loss + self.shape_reg_loss + self.finger_reg_loss)

    # ------------------------------------------------------------------- stage control
    def _begin_stage(self, stage):
        live = []
        for name in stage["update_params"]:
            assert name.startswith("pred_") and name in self.p, name
            self.p[name] = self.p[name].detach().clone().requires_grad_(True)
            live.append(self.p[name])
        if self.optimizer_name == "adam":
            self.optimizer = torch.optim.Adam(live, lr=stage["lr"], betas=(0.9, 0.999))
        else:
            assert self.optimizer_name == "sgd"
            self.optimizer = torch.optim.SGD(live, lr=stage["lr"], momentum=0.9)
        self.snapshots = []

    def _snapshot(self, stage):
        snap = {n: self.p[n].detach().clone() for n in stage["update_params"]}
        for crit in [c for c, _ in stage["filter_loss"]] + [stage["select_loss"]]:
            assert crit not in INVALID_CRITERIA
            snap[crit] = getattr(self, crit + "_batch").detach().clone()
        self.snapshots.append(snap)

    def _end_stage(self, stage):
        crits = [c for c, _ in stage["filter_loss"]] + [stage["select_loss"]]
        losses = {c: torch.stack([s[c] for s in self.snapshots]) for c in dict.fromkeys(crits)}   # (S,B)
        ok = torch.ones_like(next(iter(losses.values())), dtype=torch.bool)
        for name, crit in stage["filter_loss"]:
            assert crit[0] in "+-"
            bar = losses[name][0:1] * (1 + (float(crit) + 0.1) / 100)
            ok &= losses[name] <= bar
        sel = losses[stage["select_loss"]].clone()
        first = sel[0].clone()
        sel[~ok] = 100000000000.0
        sel[0] = first                                   # snapshot 0 always stays eligible
        best = torch.argmin(sel, dim=0)
        cols = torch.arange(self.B, device=best.device)
        for n in stage["update_params"]:
            self.p[n] = torch.stack([s[n] for s in self.snapshots])[best, cols]
        self.last_selected = best

    def optimize(self, iter_id=0, num_iter=1):
        for stage in self.strategy:
            self._begin_stage(stage)
            for j in range(stage["epoch"] + 1):
                self.forward()
                self.compute_loss(stage["loss_weights"])
                if j % self.save_mid_freq == 0:
                    self._snapshot(stage)
                self.optimizer.zero_grad()
                self.loss.backward()
                self.optimizer.step()
            self._end_stage(stage)
        with torch.no_grad():
            self.forward()
            self.compute_loss(DEFAULT_LOSS_WEIGHTS)

    # -------------------------------------------------------------------------- output
    def get_pred_result(self):
        c = lambda t: t.detach().cpu().to(torch.float32).numpy()
        return OrderedDict(
            pred_cam_params=c(self.pred_cam_params), pred_hand_trans=c(self.p["pred_hand_trans"]),
            pred_shape_params=c(self.pred_shape_params), pred_pose_params=c(self.pred_pose_params),
            pred_right_hand_verts=c(self.pred_right_hand_verts),
            pred_left_hand_verts=c(self.pred_left_hand_verts),
            mano_params_weight=c(self.mano_params_weight), pred_joints_3d=c(self.pred_joints_3d),
            gt_joints_3d=c(self.joints_3d), collision_loss=c(self.collision_loss_batch),
            collision_loss_origin_scale=c(self.collision_loss_origin_scale),
            do_flip=np.zeros(self.B, np.int32), pred_hand_type=np.ones(self.B, np.int32))
