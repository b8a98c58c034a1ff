"""BASELINE.json's full sizes through size-independent properties (the oracle cannot run them):
batch-composition independence (bitwise), agreement with small-batch runs that ARE checked
against the oracle / golden fixtures, exact zeros for disjoint hands."""
import os

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def layers(model_root):
    from ihmr_b200 import mano_layer
    right = mano_layer.create(os.path.join(model_root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True).cuda()
    left = mano_layer.create(os.path.join(model_root, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False).cuda()
    return right, left


def test_config2_mano_8192_hands(layers, oracle_layers):
    """config 2: batched MANO forward + backward, 2 x 4096 hands."""
    n = 8192
    g = torch.Generator().manual_seed(11)
    orient = ((torch.rand(n, 3, generator=g) - 0.5) * 3.0).cuda()
    pose = (torch.randn(n, 45, generator=g) * 0.4).cuda()
    betas = torch.randn(n, 10, generator=g).cuda()
    gv = torch.randn(n, 778, 3, generator=g).cuda()
    gj = torch.randn(n, 16, 3, generator=g).cuda()

    def run(sl):
        ins = [t[sl].clone().requires_grad_(True) for t in (orient, pose, betas)]
        out = layers[0](global_orient=ins[0], hand_pose=ins[1], betas=ins[2])
        ((out.vertices * gv[sl]).sum() + (out.joints * gj[sl]).sum()).backward()
        return out.vertices.detach(), out.joints.detach(), [t.grad for t in ins]

    v, j, grads = run(slice(0, n))
    for lo, hi in ((0, 7), (4000, 4133), (8191, 8192)):          # sub-batches reproduce the big batch bit for bit
        v2, j2, g2 = run(slice(lo, hi))
        assert torch.equal(v[lo:hi], v2) and torch.equal(j[lo:hi], j2)
        for a, b in zip(grads, g2):
            assert torch.equal(a[lo:hi], b)
    # spot check against the fp64 oracle
    import copy
    o64 = copy.deepcopy(oracle_layers[0]).double()
    idx = torch.tensor([0, 1, 777, 4095, 4096, 8191])
    ref = o64(global_orient=orient[idx].cpu().double(), hand_pose=pose[idx].cpu().double(), betas=betas[idx].cpu().double())
    assert (v[idx].cpu().double() - ref.vertices).abs().max().item() <= 1e-5
    assert (j[idx].cpu().double() - ref.joints).abs().max().item() <= 1e-5


@pytest.mark.parametrize("B", [1, 16, 1024, 16384])
def test_config3_penetration_sweep(layers, oracle_layers, B):
    """config 3: penetration loss fwd+bwd, batch 1 ... 16384 frames: replicas of 8 base frames
    (4 typical, 4 near-coincident) must reproduce the 8-frame result bit for bit."""
    from ihmr_b200 import sdf_loss, synthetic
    from oracle import mano_oracle
    raws = [synthetic.make_raw_frames(0, 4, seed=0, mode=m) for m in ("typical", "collision")]
    hv = []
    for raw in raws:
        with torch.no_grad():
            rv, lv, _ = mano_oracle.two_hand_forward(oracle_layers[0], torch.tensor(raw["true_pose"]),
                                                     torch.tensor(raw["true_shape"]), torch.tensor(raw["true_trans"]))
        hv.append(torch.stack([rv, lv], 1))
    base = torch.cat(hv).cuda()                                   # (8,2,778,3)
    sdf = sdf_loss.SDFLoss(layers[0].faces, layers[1].faces).cuda()

    def run(x):
        x = x.clone().requires_grad_(True)
        l, pv, o = sdf(x, return_per_vert_loss=True, return_origin_scale_loss=True)
        l.sum().backward()
        return l.detach(), o, x.grad

    l8, o8, g8 = run(base)
    assert float(l8[4:].min()) > 0                               # the near-coincident frames do collide
    idx = torch.arange(B) % 8
    l, o, g = run(base[idx].contiguous())
    assert torch.equal(l, l8[idx]) and torch.equal(o, o8[idx]) and torch.equal(g, g8[idx])


def test_config4_full_loop_65536_frames(model_root):
    """config 4: 65536 frames x 100 iterations on one GPU.  The batch tiles the golden config-1 frame and
    two other golden inputs, so every replica must equal the small-batch result (bitwise) and the
    config-1 replica must match the fixture the UNMODIFIED reference loop produced."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    d1, out1, epochs, freq = H.load_golden("loop_cfg1.npz")
    d2, _, _, _ = H.load_golden("loop_b2_short.npz")
    base = {k: np.concatenate([d1[k], d2[k]], 0) for k in d1}     # 3 distinct frames
    strat = with_epochs(opt_default, epochs)

    def run(B):
        idx = np.arange(B) % 3
        data = {k: v[idx] for k, v in base.items()}
        m = OptimizeModel(H.make_opt(model_root, B, save_mid_freq=freq, strategy=strat, bs_norm=1))
        m.set_input(H.torch_batch(data)); m.init_optimize(); m.optimize(0, 1)
        return {k: v.copy() for k, v in m.get_pred_result().items()}

    small = run(3)
    assert np.abs(small["pred_joints_3d"][0] - out1["pred_joints_3d"][0]).max() <= 1e-4        # 0.1 mm vs reference loop
    big = run(65536)
    idx = np.arange(65536) % 3
    for k in ("pred_pose_params", "pred_shape_params", "pred_hand_trans", "pred_joints_3d", "collision_loss",
              "pred_left_hand_verts"):
        assert np.array_equal(big[k], small[k][idx]), k


def test_config5_worst_case_collisions(model_root, oracle_layers):
    """config 5: near-coincident hands. One fused iteration on 2048 such frames: finite, collision loss
    positive everywhere, gradients of replicated frames identical."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default
    B = 2048
    data = H.make_batch(oracle_layers[0], 0, 8, mode="collision")
    idx = np.arange(B) % 8
    m = OptimizeModel(H.make_opt(model_root, B, bs_norm=512))
    m.set_input(H.torch_batch({k: v[idx] for k, v in data.items()}))
    m.init_optimize()
    losses, grad = m.value_and_grad(opt_default[2])
    assert torch.isfinite(losses).all() and torch.isfinite(grad).all()
    assert float(losses[3]) > 0
    assert torch.equal(grad, grad[:8][torch.arange(B) % 8])
