"""IHMR-MLP inference path on the GPU against the pinned oracle (oracle/mlp_oracle.py) and the
fixture generated from the reference's own InterHandSubNetwork (tests/golden/mlp.npz)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


def _to_our_layout(final_params):
    """reference final_params [cam 3 | pose 96 | shape 20 | trans 3] -> this library's [cam | trans | pose | shape]"""
    f = final_params
    return torch.cat([f[:, 0:3], f[:, 119:122], f[:, 3:99], f[:, 99:119]], 1).contiguous()


@pytest.mark.parametrize("dim,stage_id", [(3, 0), (90, 3)])
def test_subnetwork_forward_vs_reference_golden(model_root, dim, stage_id):
    from ihmr_b200.mlp_refiner import MLPRefiner
    from oracle import mlp_oracle as MO
    z = np.load(os.path.join(H.GOLDEN, "mlp.npz"))
    x = torch.from_numpy(z[f"d{dim}_x"])
    m = MLPRefiner(H.make_opt(model_root, x.shape[0]))
    assert m.update_dim(stage_id) == dim
    for s in range(stage_id + 1):
        m.add_network(s, MO.seeded_state_dict(m.update_dim(s), seed=m.update_dim(s)))
    y = m.mlp_forward(stage_id, x[:, :1024].contiguous().cuda(), _to_our_layout(x[:, 1024:]).cuda())
    ref = z[f"d{dim}_y"]
    assert np.abs(y[:, :dim].cpu().numpy() - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


def test_select_better_vs_oracle():
    from ihmr_b200 import _lib
    from ihmr_b200.strategies import mlp_default
    from oracle import mlp_oracle as MO
    lib, B = _lib.load(), 4096
    g = torch.Generator().manual_seed(0)
    vals = torch.tensor([0.0, 0.5, 0.999, 1.0, 1.0001, 1.001, 2.0])
    names = ("joints_3d_loss_p", "collision_loss", "joints_2d_loss_p")
    for stage in (mlp_default[0], mlp_default[3], mlp_default[5]):
        cur = {k: vals[torch.randint(0, 7, (B,), generator=g)] for k in names}
        prev = {k: vals[torch.randint(0, 7, (B,), generator=g)] for k in names}
        want = MO.select_better(cur, prev, stage)
        c = torch.stack([cur[k] for k in names], 1).cuda().contiguous()
        p = torch.stack([prev[k] for k in names], 1).cuda().contiguous()
        pnew, params = torch.ones(B, 122).cuda(), torch.zeros(B, 122).cuda()
        kept = torch.empty(B, dtype=torch.int32, device="cuda")
        st = _lib.make_stage(dict(stage, loss_weights=dict(joints_2d_loss=0, joints_3d_loss=0, trans_loss_weight=0, shape_reg_loss_weight=0,
                                                           collision_loss_weight=0, finger_reg_loss_weight=0), lr=0.0, epoch=0))
        ptr = lambda t: C.c_void_p(t.data_ptr())
        _lib.check(lib.ihmr_select_better(B, ptr(c), ptr(p), C.byref(st), ptr(pnew), ptr(params), ptr(kept),
                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)), "select_better")
        assert torch.equal(kept.cpu().bool(), want) and 0 < int(want.sum()) < B
        live = torch.zeros(122, dtype=torch.bool)
        from ihmr_b200.mlp_refiner import PARAM_COLS
        for n in stage["update_params"]:
            live[PARAM_COLS[n][0]:PARAM_COLS[n][0] + PARAM_COLS[n][1]] = True
        assert torch.equal(params.cpu() > 0.5, want[:, None] & live[None, :])          # only the stage's columns, only kept frames
        exp_prev = torch.stack([torch.where(want, cur[k], prev[k]) for k in names], 1)
        assert torch.equal(p.cpu(), exp_prev)


@pytest.mark.parametrize("mode", ["typical", "collision"])
def test_mlp_refiner_vs_oracle(model_root, oracle_layers, mode):
    """The whole test() path (mlp_model.py:683-699) against the oracle on synthetic frames with seeded networks."""
    from ihmr_b200.mlp_refiner import MLPRefiner
    from ihmr_b200.strategies import mlp_default
    from oracle import mlp_oracle as MO
    B = 6
    data = H.make_batch(oracle_layers[0], 0 if mode == "typical" else 512, B, mode=mode)
    batch = H.torch_batch(data)
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(B, 1024, generator=g) * 0.5
    sds = [MO.seeded_state_dict(sum(MO.PARAM_DIMS[p] for p in st["update_params"]), seed=10 + i, scale=0.01) for i, st in enumerate(mlp_default)]
    # oracle
    loop = H.oracle_loop(oracle_layers, B, 1, 1)
    nets = []
    for st, sd in zip(mlp_default, sds):
        net = MO.SubNetworkOracle(sum(MO.PARAM_DIMS[p] for p in st["update_params"]))
        net.load_state_dict(sd)
        nets.append(net)
    ref, kept_ref, crit_ref = MO.mlp_test(loop, nets, feat, batch, mlp_default)
    # CUDA path
    m = MLPRefiner(H.make_opt(model_root, B))
    for i, sd in enumerate(sds):
        m.add_network(i, sd)
    res = m.test(dict(batch, img_feat=feat))
    for k_gpu, k_ref in zip(m.kept, kept_ref):
        assert torch.equal(k_gpu.cpu().bool(), k_ref)
    if mode == "collision":
        assert sum(int(k.sum()) for k in kept_ref) > 0                  # some proposals are accepted ...
    assert sum(int((~k).sum()) for k in kept_ref) > 0                   # ... and some rejected
    crit = m.criteria.cpu().numpy()
    for i, name in enumerate(("joints_3d_loss_p", "collision_loss", "joints_2d_loss_p")):
        r = crit_ref[name].numpy()
        assert np.abs(crit[:, i] - r).max() <= 1e-4 * max(1e-6, np.abs(r).max()), name
    assert np.abs(res["pred_joints_3d"] - ref["pred_joints_3d"]).max() <= 1e-4
    assert np.abs(res["pred_left_hand_verts"] - ref["pred_left_hand_verts"]).max() <= 1e-4
    assert np.abs(res["pred_pose_params"] - ref["pred_pose_params"]).max() <= 1e-5
    assert np.abs(res["pred_cam_params"] - ref["pred_cam_params"]).max() <= 1e-5
    assert np.abs(res["collision_loss_origin_scale"] - ref["collision_loss_origin_scale"]).max() <= 1e-4
