"""IHMR-MLP inference: the test-time path of the reference's ``MLPModel``
(/root/reference/src/models/mlp_model.py:683-699) on the refinement path's kernels, forward only.

    test():   criteria of the prior prediction                         (forward_backbone + compute_loss, :685-687)
              for every stage of the strategy (src/strategies/mlp_default.py):
                  x        = [img_feat (1024) | final_params (122)]    (:461)
                  proposal = params + MLP_stage(x) on the stage's update_params   (networks.py:83-105, :462-470)
                  criteria of the proposal                             (MANO forward + losses, :480-583)
                  per frame keep the proposal iff the criteria improved  (select_better_params, :592-637)
              final forward with the selected parameters               (:698-699)

The four Linear layers of a stage run on the tcgen05 contraction of the library (``ihmr_linear``: 3xTF32, fp32
accuracy); MANO, the penetration loss and the joint criteria are ``ihmr_opt_criteria`` / ``ihmr_opt_final``; the
per-frame selection is ``ihmr_select_better``.  The image backbone is not part of this path: like the reference at test
time, the per-frame feature ``img_feat`` comes with the data (data_utils.py:63-64).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib
from .mano_layer import _ptr, _stream
from .optimize_model import OptimizeModel, PrefetchedInput
from .strategies import strategies

FEAT_DIM, IN_DIM, IN_PAD = 1024, 1024 + 122, 1152
HIDDEN = (512, 256, 128)
# column ranges of the (B,122) parameter matrix [cam | trans | pose 96 | shape 20]
PARAM_COLS = {"pred_cam_params": (0, 3), "pred_hand_trans": (3, 3), "pred_right_orient": (6, 3), "pred_right_pose_params": (9, 45),
              "pred_left_orient": (54, 3), "pred_left_pose_params": (57, 45), "pred_right_shape_params": (102, 10),
              "pred_left_shape_params": (112, 10)}
MLP_DEFAULT_WEIGHTS = dict(joints_2d_loss=10.0, joints_3d_loss=10.0, collision_loss=1.0)     # mlp_model.py:219-229


def _pad_rows(w: torch.Tensor, rows: int, cols: int) -> torch.Tensor:
    out = torch.zeros(rows, cols, dtype=torch.float32, device=w.device)
    out[:w.shape[0], :w.shape[1]] = w
    return out.contiguous()


class SubNetwork:
    """Weights of one ``InterHandSubNetwork`` (networks.py:83-105): Linear(1146,512) ReLU Linear(512,256) ReLU
    Linear(256,128) ReLU Linear(128, update_dim), held in the padded layouts ``ihmr_linear`` takes."""

    def __init__(self, update_dim: int, device, state_dict: Optional[Dict[str, torch.Tensor]] = None, seed: int = 0):
        self.update_dim = int(update_dim)
        dims = [(IN_DIM, HIDDEN[0]), (HIDDEN[0], HIDDEN[1]), (HIDDEN[1], HIDDEN[2]), (HIDDEN[2], self.update_dim)]
        if state_dict is None:                       # the reference's initialisation: xavier_uniform_(gain=0.01) weights
            g = torch.Generator().manual_seed(seed)
            state_dict = {}
            for i, (fan_in, fan_out) in enumerate(dims):
                bound = 0.01 * (6.0 / (fan_in + fan_out)) ** 0.5
                state_dict[f"regressor.{2 * i}.weight"] = (torch.rand(fan_out, fan_in, generator=g) * 2 - 1) * bound
                b = 1.0 / fan_in ** 0.5
                state_dict[f"regressor.{2 * i}.bias"] = (torch.rand(fan_out, generator=g) * 2 - 1) * b
        sd = {k.replace("module.", ""): v for k, v in state_dict.items()}       # DistributedDataParallel prefix (:383-385)
        self.state_dict = {k: v.detach().float().cpu() for k, v in sd.items()}
        self.layers = []
        for i, (fan_in, fan_out) in enumerate(dims):
            w, b = sd[f"regressor.{2 * i}.weight"].float().to(device), sd[f"regressor.{2 * i}.bias"].float().to(device)
            assert tuple(w.shape) == (fan_out, fan_in), (i, tuple(w.shape))
            in_pad = IN_PAD if i == 0 else fan_in
            self.layers.append((_pad_rows(w, (fan_out + 3) // 4 * 4, in_pad), b.contiguous(), in_pad, fan_out))


class MLPRefiner:
    """``MLPRefiner(opt)``; ``add_network(stage_id, state_dict)`` per stage (``net_mlp_stage_XX`` weights); ``test(input)``
    returns the same 13 result arrays as ``OptimizeModel.get_pred_result`` (mlp_model.py: get_pred_result)."""

    def __init__(self, opt, strategy=None, device=None):
        self.core = OptimizeModel(opt, device=device)            # inputs, workspace, final forward, result copies
        self.device, self.lib, self.batch_size = self.core.device, self.core.lib, self.core.batch_size
        self.strategy: List[dict] = strategy if strategy is not None else strategies["mlp_default"]
        self.loss_weights = dict(MLP_DEFAULT_WEIGHTS)
        self.sub_networks: List[SubNetwork] = []
        self.kept: List[torch.Tensor] = []

    def update_dim(self, stage_id: int) -> int:
        return sum(PARAM_COLS[p][1] for p in self.strategy[stage_id]["update_params"])

    def add_network(self, stage_id: int, state_dict=None, seed: int = 0) -> SubNetwork:
        assert stage_id == len(self.sub_networks), "stages are added in order (mlp_model.py:370-386)"
        net = SubNetwork(self.update_dim(stage_id), self.device, state_dict, seed=seed + stage_id)
        self.sub_networks.append(net)
        return net

    # ------------------------------------------------------------------------------------------
    def _criteria(self, params: torch.Tensor) -> torch.Tensor:
        ws = self.core._workspace()
        crit = torch.empty(self.batch_size, 3, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.ihmr_opt_criteria(self.core._model.handle, self.batch_size, _ptr(params), C.byref(self.core._targets),
                                              float(self.loss_weights["joints_2d_loss"]), float(self.loss_weights["joints_3d_loss"]),
                                              _ptr(crit), _ptr(ws), ws.numel(), _stream(self.device)), "ihmr_opt_criteria")
        return crit

    def mlp_forward(self, stage_id: int, img_feat: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
        """(B, ceil4(update_dim)) residual of the stage's network."""
        B, st = self.batch_size, _stream(self.device)
        x = torch.empty(B, IN_PAD, device=self.device, dtype=torch.float32)
        _lib.check(self.lib.ihmr_mlp_input(B, _ptr(img_feat), _ptr(params), _ptr(x), st), "ihmr_mlp_input")
        for i, (w, b, in_pad, out_dim) in enumerate(self.sub_networks[stage_id].layers):
            y = torch.empty(B, (out_dim + 3) // 4 * 4, device=self.device, dtype=torch.float32)
            _lib.check(self.lib.ihmr_linear(B, in_pad, out_dim, _ptr(x), x.shape[1], _ptr(w), _ptr(b), 1 if i < 3 else 0,
                                            _ptr(y), y.shape[1], st), "ihmr_linear")
            x = y
        return x

    def test(self, input):
        """mlp_model.py:683-699.  ``input``: the OPT batch keys plus ``img_feat`` (B,1024)."""
        assert len(self.sub_networks) == len(self.strategy), "add_network for every stage first"
        core, B, st = self.core, self.batch_size, _stream(self.device)
        feat = input["img_feat"]
        feat = (feat if isinstance(feat, torch.Tensor) else torch.as_tensor(feat)).to(self.device, torch.float32).contiguous()
        assert tuple(feat.shape) == (B, FEAT_DIM)
        core.set_input(input if isinstance(input, PrefetchedInput) else {k: v for k, v in input.items() if k != "img_feat"})
        core.init_optimize()                                           # params = the prior's prediction (:441-456)
        params = core.params
        prev = self._criteria(params)                                  # forward_backbone + compute_loss + save_pred_to_prev
        self.kept = []
        for stage_id, stage in enumerate(self.strategy):
            res = self.mlp_forward(stage_id, feat, params)
            cols = [PARAM_COLS[p] for p in stage["update_params"]]
            seg_col = (C.c_int32 * 8)(*[c for c, _ in cols]); seg_len = (C.c_int32 * 8)(*[n for _, n in cols])
            proposal = torch.empty_like(params)
            _lib.check(self.lib.ihmr_mlp_apply(B, _ptr(res), res.shape[1], len(cols), seg_col, seg_len, _ptr(params), _ptr(proposal), st),
                       "ihmr_mlp_apply")
            cur = self._criteria(proposal)
            kept = torch.empty(B, dtype=torch.int32, device=self.device)
            sstruct = _lib.make_stage(dict(stage, loss_weights=dict(joints_2d_loss=0, joints_3d_loss=0, trans_loss_weight=0,
                                                                     shape_reg_loss_weight=0, collision_loss_weight=0,
                                                                     finger_reg_loss_weight=0), lr=0.0, epoch=0))
            _lib.check(self.lib.ihmr_select_better(B, _ptr(cur), _ptr(prev), C.byref(sstruct), _ptr(proposal), _ptr(params), _ptr(kept), st),
                       "ihmr_select_better")
            self.kept.append(kept)
        self.criteria = prev
        core.forward()                                                 # forward mano after obtaining results (:698-699)
        return core.get_pred_result()
