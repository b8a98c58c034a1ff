"""CPU tests of the oracle leaves: known-answer cases (the reference ships no tests, SURVEY.md
§4, so these pin the oracle to analytic facts and to the committed golden fixtures)."""
import copy
import os

import numpy as np
import pytest
import torch

from ihmr_b200 import synthetic
from oracle import mano_oracle, sdf_oracle
from tests import helpers as H


# ------------------------------------------------------------------------ synthetic model
def test_synthetic_mesh_topology():
    m = synthetic.make_mano_model(0)
    v, f = m["v_template"], m["f"]
    assert v.shape == (778, 3) and f.shape == (1538, 3)
    directed = {}
    for a, b, c in f:
        for e in ((a, b), (b, c), (c, a)):
            assert e not in directed          # consistently oriented manifold
            directed[e] = 1
    boundary = [e for e in directed if (e[1], e[0]) not in directed]
    assert len(boundary) == 16                # open wrist loop, like MANO
    assert np.allclose(m["weights"].sum(1), 1, atol=1e-5) and np.allclose(m["J_regressor"].sum(1), 1, atol=1e-5)
    assert list(m["kintree_table"][0][1:]) == [0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14]


def test_frames_do_not_depend_on_sharding():
    full = synthetic.make_raw_frames(0, 700, seed=0)
    part = synthetic.make_raw_frames(300, 150, seed=0)
    for k in full:
        assert np.array_equal(full[k][300:450], part[k]), k


# ---------------------------------------------------------------------------- MANO layer
def _zero_mean(layer):
    l = copy.deepcopy(layer).double()
    l.hands_mean.zero_()
    return l


def test_mano_rest_pose_is_blend_shape(oracle_layers):
    l = _zero_mean(oracle_layers[0])
    betas = torch.randn(3, 10, dtype=torch.float64)
    out = l(global_orient=torch.zeros(3, 3, dtype=torch.float64), hand_pose=torch.zeros(3, 45, dtype=torch.float64), betas=betas)
    v_shaped = l.v_template[None] + torch.einsum("bl,mkl->bmk", betas, l.shapedirs)
    assert torch.allclose(out.vertices, v_shaped, atol=1e-9)
    assert torch.allclose(out.joints, torch.einsum("bik,ji->bjk", v_shaped, l.J_regressor), atol=1e-9)


def test_mano_global_rotation_is_rigid_about_wrist(oracle_layers):
    from scipy.spatial.transform import Rotation
    l = _zero_mean(oracle_layers[0])
    pose = torch.randn(1, 45, dtype=torch.float64) * 0.3
    betas = torch.randn(1, 10, dtype=torch.float64)
    rv = np.array([0.3, -0.7, 0.5])
    a = l(global_orient=torch.zeros(1, 3, dtype=torch.float64), hand_pose=pose, betas=betas)
    b = l(global_orient=torch.tensor(rv)[None], hand_pose=pose, betas=betas)
    R = torch.tensor(Rotation.from_rotvec(rv).as_matrix())
    root = a.joints[:, 0:1]
    assert torch.allclose(b.vertices, (a.vertices - root) @ R.T + root, atol=1e-7)
    assert torch.allclose(b.joints[:, 0], a.joints[:, 0], atol=1e-12)      # wrist does not move


def test_rodrigues_matches_scipy_and_handles_zero():
    from scipy.spatial.transform import Rotation
    r = torch.randn(20, 3, dtype=torch.float64)
    R = mano_oracle.batch_rodrigues(r).numpy()
    assert np.allclose(R, Rotation.from_rotvec(r.numpy()).as_matrix(), atol=1e-6)
    assert torch.allclose(mano_oracle.batch_rodrigues(torch.zeros(1, 3, dtype=torch.float64))[0], torch.eye(3, dtype=torch.float64))


def test_two_hand_mirror_identity(oracle_layers):
    """Feeding mirrored parameters to the left hand yields the x-mirror of the right hand."""
    l = oracle_layers[0]
    g = torch.Generator().manual_seed(0)
    pose_r = torch.randn(2, 48, generator=g) * 0.3
    M = torch.tensor([1.0, -1.0, -1.0]).repeat(16)
    pose = torch.cat([pose_r, pose_r * M], 1)
    shape = torch.randn(2, 10, generator=g).repeat(1, 2)
    rv, lv, joints = mano_oracle.two_hand_forward(l, pose, shape, torch.zeros(2, 3))
    X = torch.tensor([-1.0, 1.0, 1.0])
    # left = mirror(right) translated so that the wrists coincide
    shift = joints[:, 21:22] - joints[:, 0:1] * X
    assert torch.allclose(lv, rv * X + shift, atol=1e-6)
    assert torch.allclose(joints[:, 21], joints[:, 0], atol=1e-6)          # hand_trans = 0


def test_hand_trans_moves_only_left_hand(oracle_layers):
    l = oracle_layers[0]
    raw = synthetic.make_raw_frames(0, 2)
    p, s = torch.tensor(raw["true_pose"]), torch.tensor(raw["true_shape"])
    t = torch.tensor([[0.01, -0.02, 0.03]]).repeat(2, 1)
    rv0, lv0, j0 = mano_oracle.two_hand_forward(l, p, s, torch.zeros(2, 3))
    rv1, lv1, j1 = mano_oracle.two_hand_forward(l, p, s, t)
    assert torch.equal(rv0, rv1) and torch.allclose(lv1 - lv0, t[:, None].expand_as(lv0), atol=1e-7)
    assert torch.allclose(j1[:, 21] - j1[:, 0], t, atol=1e-6)


def test_mano_gradcheck_fp64(oracle_layers):
    l = copy.deepcopy(oracle_layers[0]).double()
    g = torch.Generator().manual_seed(1)
    ins = [torch.randn(1, 3, dtype=torch.float64, generator=g).requires_grad_(True),
           (torch.randn(1, 45, dtype=torch.float64, generator=g) * 0.3).requires_grad_(True),
           torch.randn(1, 10, dtype=torch.float64, generator=g).requires_grad_(True)]
    sel = torch.tensor([0, 100, 320, 777])

    def f(o, p, b):
        out = l(global_orient=o, hand_pose=p, betas=b)
        return out.vertices[:, sel].sum(1), out.joints.sum(1)
    assert torch.autograd.gradcheck(f, ins, eps=1e-6, atol=1e-6)


def test_mano_leaf_golden(oracle_layers):
    z = np.load(os.path.join(H.GOLDEN, "leaves.npz"))
    out = oracle_layers[0](global_orient=torch.tensor(z["mano_orient"]), hand_pose=torch.tensor(z["mano_pose"]),
                           betas=torch.tensor(z["mano_betas"]))
    assert np.abs(out.vertices.numpy() - z["mano_vertices"]).max() < 1e-6
    assert np.abs(out.joints.numpy() - z["mano_joints"]).max() < 1e-6


# ------------------------------------------------------------------------------- SDF leaf
def _icosphere(radius, subdiv=2):
    t = (1 + 5 ** 0.5) / 2
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]])
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    for _ in range(subdiv):
        cache, nf = {}, []
        v = list(v)

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (np.asarray(v[a]) + np.asarray(v[b])) / 2
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f, v = np.array(nf), np.array(v)
    return v * radius, f


def test_sdf_grid_c_matches_numpy_twin():
    v, f = _icosphere(0.7, subdiv=1)
    v = v + np.array([0.013, -0.007, 0.021])
    G = 8
    ref = sdf_oracle.sdf_grid_numpy(v, f, G)
    for dtype in (torch.float32, torch.float64):
        phi = sdf_oracle.sdf_grid(torch.tensor(v, dtype=dtype)[None], torch.tensor(f), G)[0].double().numpy()
        assert np.array_equal(phi > 0, ref > 0)
        assert np.abs(phi - ref).max() < 1e-5


def test_sdf_sphere_depth_is_analytic():
    """Inside a (finely triangulated) sphere of radius r, phi(q) ~= r - |q|; outside it is 0."""
    r = 0.75
    v, f = _icosphere(r, subdiv=3)
    G = 32
    phi = sdf_oracle.sdf_grid(torch.tensor(v, dtype=torch.float32)[None], torch.tensor(f), G)[0].numpy()
    c = (2 * np.arange(G) + 1 - G) / G
    zz, yy, xx = np.meshgrid(c, c, c, indexing="ij")
    rad = np.sqrt(xx ** 2 + yy ** 2 + zz ** 2)
    inside = rad < r - 0.03
    outside = rad > r + 0.01
    assert np.all(phi[outside] == 0)
    assert np.all(phi[inside] > 0)
    assert np.abs(phi[inside] - (r - rad[inside])).max() < 0.012        # chordal error of the facets


def test_sdf_loss_zero_when_boxes_are_disjoint(oracle_layers):
    raw = synthetic.make_raw_frames(0, 1)
    with torch.no_grad():
        rv, lv, _ = mano_oracle.two_hand_forward(oracle_layers[0], torch.tensor(raw["true_pose"]),
                                                 torch.tensor(raw["true_shape"]), torch.tensor([[0.6, 0.0, 0.0]]))
    loss = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces)
    l, pv, o = loss(torch.stack([rv, lv], 1), True, True)
    assert float(l.abs().max()) == 0.0 and float(o.abs().max()) == 0.0 and o.shape == (1, 1556)


def test_sdf_leaf_golden_and_grad_flows_to_sampled_hand_only(oracle_layers):
    z = np.load(os.path.join(H.GOLDEN, "leaves.npz"))
    hv = torch.tensor(z["sdf_hand_verts"])[:1].clone().requires_grad_(True)
    loss = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces)
    l, pv, o = loss(hv, True, True)
    l.sum().backward()
    assert np.allclose(l.detach().numpy(), z["sdf_losses"][:1], rtol=1e-6)
    assert np.abs(o.numpy() - z["sdf_origin_scale"][:1]).max() < 1e-7
    assert np.abs(hv.grad.numpy() - z["sdf_grad"][:1]).max() < 1e-6 * np.abs(z["sdf_grad"]).max() + 1e-9
    # non-negative, metres, right-hand vertices first (evaluator.py:119-120)
    assert float(o.min()) >= 0 and float(l.min()) >= 0
    # loss = sum of per-vertex values / 4 (A6)
    assert np.isclose(float(l[0]), float(pv[0].sum()) / 4, rtol=1e-6)
