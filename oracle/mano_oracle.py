"""ORACLE (test infrastructure, not product code) — CPU restatement of the MANO layer.

PARITY UNPINNED: the arithmetic of this leaf lives in the third-party package
``smplx==0.1.28`` (pin: /root/reference/docs/ihmr.yml:132), which is not vendored in the
reference and not installable offline.  This file restates its published algorithm
(``smplx/body_models.py::MANO.forward`` and ``smplx/lbs.py::{lbs, blend_shapes,
vertices2joints, batch_rodrigues, batch_rigid_transform}``) in plain torch, anchored on
the reference's own call site ``src/models/optimize_model.py:105-106,194-200`` and the
assumptions M1-M6 of SURVEY.md §8(c).  The reference ships no test or golden vector for
it.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may
import this module.

Works in float32 (CPU baseline) and float64 (numerical oracle).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ihmr_b200.mano_io import load_mano_pkl


def batch_rodrigues(rot_vecs: torch.Tensor) -> torch.Tensor:
    """[UPSTREAM] smplx.lbs.batch_rodrigues: angle = ||r + 1e-8|| (M3), R = I + sin K + (1-cos) K^2."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle)[:, None]
    sin = torch.sin(angle)[:, None]
    rx, ry, rz = rot_dir[:, 0:1], rot_dir[:, 1:2], rot_dir[:, 2:3]
    zeros = torch.zeros_like(rx)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """[UPSTREAM] smplx.lbs.batch_rigid_transform: kinematic chain, returns posed joints and
    the relative transforms A_j = G_j with translation G_j[:3,3] - G_j[:3,:3] J_j."""
    n, nj = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] = rel[:, 1:] - joints[:, parents[1:]]
    top = torch.cat([rot_mats.reshape(-1, 3, 3), rel.reshape(-1, 3, 1)], dim=2)
    bottom = torch.zeros(n * nj, 1, 4, dtype=joints.dtype, device=joints.device)
    bottom[:, 0, 3] = 1
    tm = torch.cat([top, bottom], dim=1).view(n, nj, 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, nj):
        chain.append(torch.matmul(chain[int(parents[i])], tm[:, i]))
    transforms = torch.stack(chain, dim=1)
    posed_joints = transforms[:, :, :3, 3]
    joints_h = F.pad(joints, [0, 0, 0, 1])
    rel_transforms = transforms - F.pad(torch.matmul(transforms, joints_h), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_transforms


def lbs(betas, full_pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights):
    """[UPSTREAM] smplx.lbs.lbs with pose2rot=True."""
    n = betas.shape[0]
    dtype = betas.dtype
    v_shaped = v_template[None] + torch.einsum("bl,mkl->bmk", betas, shapedirs)
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)
    ident = torch.eye(3, dtype=dtype, device=betas.device)
    rot_mats = batch_rodrigues(full_pose.reshape(-1, 3)).view(n, -1, 3, 3)
    pose_feature = (rot_mats[:, 1:] - ident).reshape(n, -1)
    v_posed = v_shaped + torch.matmul(pose_feature, posedirs).view(n, -1, 3)
    J_transformed, A = batch_rigid_transform(rot_mats, J, parents)
    nj = J_regressor.shape[0]
    W = lbs_weights[None].expand(n, -1, -1)
    T = torch.matmul(W, A.reshape(n, nj, 16)).view(n, -1, 4, 4)
    ones = torch.ones(n, v_posed.shape[1], 1, dtype=dtype, device=betas.device)
    v_h = torch.matmul(T, torch.cat([v_posed, ones], dim=2).unsqueeze(-1))
    return v_h[:, :, :3, 0], J_transformed


class ManoLayerOracle(nn.Module):
    """Duck-type of the object ``smplx.create(path, 'mano', use_pca=False, is_rhand=...,
    batch_size=...)`` returns (boundary L0 of SURVEY.md §8(b)): attributes ``shapedirs``
    (mutable tensor), ``faces`` (ndarray), ``J_regressor``; ``forward(global_orient,
    hand_pose, betas)`` -> object with ``.vertices`` (N,778,3) and ``.joints`` (N,16,3)."""

    def __init__(self, model_path, is_rhand=True, batch_size=1, dtype=torch.float32, **_):
        super().__init__()
        m = load_mano_pkl(model_path)
        self.is_rhand = is_rhand
        self.batch_size = batch_size
        self.faces = m["faces"]
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        self.register_buffer("v_template", t(m["v_template"]))
        self.register_buffer("shapedirs", t(m["shapedirs"]))
        self.register_buffer("posedirs", t(m["posedirs"]))
        self.register_buffer("J_regressor", t(m["J_regressor"]))
        self.register_buffer("lbs_weights", t(m["lbs_weights"]))
        self.register_buffer("hands_mean", t(m["hands_mean"]))
        self.register_buffer("parents", torch.tensor(m["parents"], dtype=torch.long))
        self.register_buffer("faces_tensor", torch.tensor(m["faces"], dtype=torch.long))

    def forward(self, global_orient, hand_pose, betas, **_):
        # M1: flat_hand_mean=False => hand_pose += hands_mean; pose_mean = [0,0,0, hands_mean]
        full_pose = torch.cat([global_orient, hand_pose + self.hands_mean[None]], dim=1)
        verts, joints = lbs(betas, full_pose, self.v_template, self.shapedirs, self.posedirs,
                            self.J_regressor, self.parents, self.lbs_weights)
        return SimpleNamespace(vertices=verts, joints=joints, betas=betas,
                               global_orient=global_orient, hand_pose=hand_pose, full_pose=full_pose)


def create(model_path, model_type="mano", use_pca=False, is_rhand=True, batch_size=1, **kw):
    """Signature of ``smplx.create`` as called at src/models/optimize_model.py:105-106."""
    assert model_type == "mano" and not use_pca
    return ManoLayerOracle(model_path, is_rhand=is_rhand, batch_size=batch_size, **kw)


TIP_IDS = (744, 320, 443, 554, 671)   # src/models/optimize_model.py:99


def two_hand_forward(layer: ManoLayerOracle, pose96, shape20, trans3):
    """Restates OptimizeModel.get_mano_output (src/models/optimize_model.py:171-232) for the
    frame generator: returns right verts, left verts, joints (B,42,3)."""
    B = pose96.shape[0]
    M = torch.tensor([1.0, -1.0, -1.0], dtype=pose96.dtype)
    r_or, r_po = pose96[:, 0:3], pose96[:, 3:48]
    l_or = pose96[:, 48:51] * M
    l_po = (pose96[:, 51:96].reshape(B, 15, 3) * M).reshape(B, 45)
    out = layer(global_orient=torch.cat([r_or, l_or]), hand_pose=torch.cat([r_po, l_po]),
                betas=torch.cat([shape20[:, :10], shape20[:, 10:]]))
    verts = out.vertices
    joints = torch.cat([out.joints, verts[:, list(TIP_IDS)]], dim=1)
    X = torch.tensor([-1.0, 1.0, 1.0], dtype=pose96.dtype)
    rv, rj = verts[:B], joints[:B]
    lv, lj = verts[B:] * X, joints[B:] * X
    shift = trans3.view(B, 1, 3) + rj[:, 0:1] - lj[:, 0:1]
    return rv, lv + shift, torch.cat([rj, lj + shift], dim=1)
