"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

Tolerances are the ones BASELINE.json's north_star states: max vertex error <= 1e-5 m per MANO
forward, loss and gradient relative error <= 1e-4, final refined joints within 0.1 mm.
"""
import copy
import os

import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

VERT_TOL = 1e-5        # metres
REL_TOL = 1e-4
JOINT_TOL = 1e-4       # metres (0.1 mm)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def cuda_layers(model_root):
    from ihmr_b200 import mano_layer
    right = mano_layer.create(os.path.join(model_root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True).cuda()
    left = mano_layer.create(os.path.join(model_root, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False).cuda()
    return right, left


@pytest.fixture(scope="module")
def oracle64(oracle_layers):
    r, l = copy.deepcopy(oracle_layers[0]).double(), copy.deepcopy(oracle_layers[1]).double()
    return r, l


def random_hands(n, seed):
    g = torch.Generator().manual_seed(seed)
    orient = (torch.rand(n, 3, generator=g) - 0.5) * 3.0
    pose = torch.randn(n, 45, generator=g) * 0.4
    betas = torch.randn(n, 10, generator=g)
    orient[0] = 0
    pose[0] = 0
    return orient, pose, betas


# ------------------------------------------------------------------------------ MANO layer
def test_mano_forward_golden(cuda_layers):
    z = np.load(os.path.join(H.GOLDEN, "leaves.npz"))
    out = cuda_layers[0](global_orient=torch.tensor(z["mano_orient"]).cuda(), hand_pose=torch.tensor(z["mano_pose"]).cuda(),
                         betas=torch.tensor(z["mano_betas"]).cuda())
    assert np.abs(out.vertices.cpu().numpy() - z["mano_vertices"]).max() <= VERT_TOL
    assert np.abs(out.joints.cpu().numpy() - z["mano_joints"]).max() <= VERT_TOL


@pytest.mark.parametrize("n", [1, 7, 64, 333])
def test_mano_forward_vs_oracle(cuda_layers, oracle64, n):
    orient, pose, betas = random_hands(n, seed=n)
    ref = oracle64[0](global_orient=orient.double(), hand_pose=pose.double(), betas=betas.double())
    out = cuda_layers[0](global_orient=orient.cuda(), hand_pose=pose.cuda(), betas=betas.cuda())
    assert out.vertices.shape == (n, 778, 3) and out.joints.shape == (n, 16, 3)
    assert (out.vertices.cpu().double() - ref.vertices).abs().max().item() <= VERT_TOL
    assert (out.joints.cpu().double() - ref.joints).abs().max().item() <= VERT_TOL


@pytest.mark.parametrize("n", [1, 9, 40])
def test_mano_backward_vs_oracle(cuda_layers, oracle64, n):
    orient, pose, betas = random_hands(n, seed=100 + n)
    g = torch.Generator().manual_seed(5)
    gv, gj = torch.randn(n, 778, 3, generator=g), torch.randn(n, 16, 3, generator=g) * 10
    ins64 = [t.double().requires_grad_(True) for t in (orient, pose, betas)]
    ref = oracle64[0](global_orient=ins64[0], hand_pose=ins64[1], betas=ins64[2])
    ((ref.vertices * gv.double()).sum() + (ref.joints * gj.double()).sum()).backward()
    ins = [t.cuda().requires_grad_(True) for t in (orient, pose, betas)]
    out = cuda_layers[0](global_orient=ins[0], hand_pose=ins[1], betas=ins[2])
    ((out.vertices * gv.cuda()).sum() + (out.joints * gj.cuda()).sum()).backward()
    for a, b, name in zip(ins, ins64, ("orient", "pose", "betas")):
        assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) <= REL_TOL, name


def test_mano_backward_only_vertices_or_joints(cuda_layers, oracle64):
    orient, pose, betas = random_hands(5, seed=9)
    for use in ("vertices", "joints"):
        ins64 = [t.double().requires_grad_(True) for t in (orient, pose, betas)]
        getattr(oracle64[0](global_orient=ins64[0], hand_pose=ins64[1], betas=ins64[2]), use).square().sum().backward()
        ins = [t.cuda().requires_grad_(True) for t in (orient, pose, betas)]
        getattr(cuda_layers[0](global_orient=ins[0], hand_pose=ins[1], betas=ins[2]), use).square().sum().backward()
        for a, b in zip(ins, ins64):
            assert rel_err(a.grad.cpu().numpy(), b.grad.numpy()) <= REL_TOL


# ---------------------------------------------------------------------- penetration loss
def _cuda_sdf(cuda_layers):
    from ihmr_b200 import sdf_loss
    return sdf_loss.SDFLoss(cuda_layers[0].faces, cuda_layers[1].faces).cuda()


def test_sdf_golden(cuda_layers):
    z = np.load(os.path.join(H.GOLDEN, "leaves.npz"))
    hv = torch.tensor(z["sdf_hand_verts"]).cuda().requires_grad_(True)
    losses, per_vert, origin = _cuda_sdf(cuda_layers)(hv, return_per_vert_loss=True, return_origin_scale_loss=True)
    losses.sum().backward()
    assert rel_err(losses.detach().cpu().numpy(), z["sdf_losses"]) <= REL_TOL
    assert np.abs(origin.cpu().numpy() - z["sdf_origin_scale"]).max() <= 1e-6
    assert rel_err(hv.grad.cpu().numpy(), z["sdf_grad"]) <= REL_TOL


@pytest.mark.parametrize("mode,start,count", [("typical", 0, 6), ("collision", 0, 3), ("typical", 40, 6)])
def test_sdf_vs_oracle(cuda_layers, oracle_layers, mode, start, count):
    from ihmr_b200 import synthetic
    from oracle import mano_oracle, sdf_oracle
    raw = synthetic.make_raw_frames(start, count, seed=0, mode=mode)
    with torch.no_grad():
        rv, lv, _ = mano_oracle.two_hand_forward(oracle_layers[0], torch.tensor(raw["true_pose"]),
                                                 torch.tensor(raw["true_shape"]), torch.tensor(raw["true_trans"]))
    hv_cpu = torch.stack([rv, lv], 1).clone().requires_grad_(True)
    ref = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces)
    l_ref, pv_ref, o_ref = ref(hv_cpu, True, True)
    l_ref.sum().backward()
    hv = hv_cpu.detach().cuda().requires_grad_(True)
    l, pv, o = _cuda_sdf(cuda_layers)(hv, return_per_vert_loss=True, return_origin_scale_loss=True)
    l.sum().backward()
    assert o.shape == (count, 1556)
    scale = max(float(l_ref.abs().max()), 1e-12)
    assert float((l.detach().cpu() - l_ref.detach()).abs().max()) <= REL_TOL * scale + 1e-9
    assert float((o.cpu() - o_ref).abs().max()) <= 2e-6
    assert float((pv.cpu() - pv_ref.detach()).abs().max()) <= 2e-5
    gscale = max(float(hv_cpu.grad.abs().max()), 1e-12)
    assert float((hv.grad.cpu() - hv_cpu.grad).abs().max()) <= REL_TOL * gscale + 1e-9


@pytest.mark.parametrize("scale_factor,ray_axis", [(0.2, 1), (0.2, 2), (0.35, 0), (0.1, 2)])
def test_sdf_conventions_flip_together_with_the_oracle(cuda_layers, oracle_layers, scale_factor, ray_axis):
    """The two assumptions about the un-vendored `sdf` package that are parameters on both sides (box
    scale factor, A4: axis of the inside/outside parity ray): changed in the oracle and in the kernels together the
    results still agree, and they differ from the default convention's (so the switch really reaches the kernels)."""
    from ihmr_b200 import sdf_loss, synthetic
    from oracle import mano_oracle, sdf_oracle
    frames = []
    for mode, start, count in (("collision", 8, 5), ("typical", 300, 7)):
        raw = synthetic.make_raw_frames(start, count, seed=0, mode=mode)
        with torch.no_grad():
            rv, lv, _ = mano_oracle.two_hand_forward(oracle_layers[0], torch.tensor(raw["true_pose"]),
                                                     torch.tensor(raw["true_shape"]), torch.tensor(raw["true_trans"]))
        frames.append(torch.stack([rv, lv], 1))
    hv_cpu = torch.cat(frames).clone().requires_grad_(True)
    ref = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces, ray_axis=ray_axis)
    l_ref, pv_ref, o_ref = ref(hv_cpu, True, True, scale_factor=scale_factor)
    l_ref.sum().backward()
    hv = hv_cpu.detach().cuda().requires_grad_(True)
    mod = sdf_loss.SDFLoss(cuda_layers[0].faces, cuda_layers[1].faces, ray_axis=ray_axis).cuda()
    l, pv, o = mod(hv, return_per_vert_loss=True, return_origin_scale_loss=True, scale_factor=scale_factor)
    l.sum().backward()
    scale = max(float(l_ref.abs().max()), 1e-12)
    assert float(l_ref.abs().max()) > 0
    assert float((l.detach().cpu() - l_ref.detach()).abs().max()) <= REL_TOL * scale + 1e-9
    assert float((o.cpu() - o_ref).abs().max()) <= 2e-6
    gscale = max(float(hv_cpu.grad.abs().max()), 1e-12)
    assert float((hv.grad.cpu() - hv_cpu.grad).abs().max()) <= REL_TOL * gscale + 1e-9
    l_def = _cuda_sdf(cuda_layers)(hv.detach())
    assert float((l_def.cpu() - l.detach().cpu()).abs().max()) > 1e-3 * scale


def test_sdf_disjoint_hands_is_exact_zero(cuda_layers):
    orient, pose, betas = random_hands(4, seed=3)
    v = cuda_layers[0](global_orient=orient.cuda(), hand_pose=pose.cuda(), betas=betas.cuda()).vertices
    hv = torch.stack([v[:2], v[2:] + 5.0], 1).contiguous().requires_grad_(True)
    l, pv, o = _cuda_sdf(cuda_layers)(hv, return_per_vert_loss=True, return_origin_scale_loss=True)
    l.sum().backward()
    assert float(l.abs().max()) == 0.0 and float(o.abs().max()) == 0.0 and float(hv.grad.abs().max()) == 0.0


# ------------------------------------------------------------------ one fused iteration
STAGE_IDS = [0, 1, 2, 3]


@pytest.mark.parametrize("stage_id", STAGE_IDS)
@pytest.mark.parametrize("mode", ["typical", "collision"])
def test_value_and_grad_vs_oracle(model_root, oracle_layers, stage_id, mode):
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default
    B = 3
    data = H.make_batch(oracle_layers[0], 0, B, mode=mode)
    batch = H.torch_batch(data)
    stage = opt_default[stage_id]
    # oracle: the host loop restatement with autograd over the oracle leaves, all params live
    hl = H.oracle_loop(oracle_layers, B, 1, 1)
    hl.set_input(batch)
    hl.init_optimize()
    for k in hl.p:
        hl.p[k] = hl.p[k].detach().clone().requires_grad_(True)
    hl.forward()
    hl.compute_loss(stage["loss_weights"])
    hl.loss.backward()
    ref_losses = [hl.joints_2d_loss_p, hl.joints_3d_loss_p, hl.hand_trans_loss_p, hl.collision_loss,
                  hl.shape_reg_loss, hl.finger_reg_loss]
    ref_grad = torch.cat([torch.zeros(B, 3), hl.p["pred_hand_trans"].grad.view(B, 3), hl.p["pred_right_orient"].grad,
                          hl.p["pred_right_pose_params"].grad, hl.p["pred_left_orient"].grad,
                          hl.p["pred_left_pose_params"].grad, hl.p["pred_right_shape_params"].grad,
                          hl.p["pred_left_shape_params"].grad], 1).numpy()

    model = OptimizeModel(H.make_opt(model_root, B))
    model.set_input(batch)
    model.init_optimize()
    losses, grad = model.value_and_grad(stage)
    losses, grad = losses.cpu().numpy(), grad.cpu().numpy()
    for i, r in enumerate(ref_losses):
        r = float(r)
        assert abs(losses[i] - r) <= REL_TOL * max(abs(r), 1e-6), (i, losses[i], r)
    groups = {"trans": slice(3, 6), "r_orient": slice(6, 9), "r_pose": slice(9, 54), "l_orient": slice(54, 57),
              "l_pose": slice(57, 102), "r_shape": slice(102, 112), "l_shape": slice(112, 122)}
    for name, sl in groups.items():
        assert rel_err(grad[:, sl], ref_grad[:, sl]) <= REL_TOL, name


@pytest.mark.parametrize("stage_id", [0, 2])
def test_value_and_grad_under_another_ray_axis(model_root, oracle_layers, stage_id):
    """One fused iteration (losses + parameter gradients) with the parity ray along +y in the oracle loop and in the
    CUDA loop (`opt.sdf_ray_axis`): the convention travels through the stage kernels, the stage-0 shortcut included."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default
    B = 3
    batch = H.torch_batch(H.make_batch(oracle_layers[0], 0, B, mode="collision"))
    stage = opt_default[stage_id]
    hl = H.oracle_loop(oracle_layers, B, 1, 1, ray_axis=1)
    hl.set_input(batch)
    hl.init_optimize()
    for k in hl.p:
        hl.p[k] = hl.p[k].detach().clone().requires_grad_(True)
    hl.forward()
    hl.compute_loss(stage["loss_weights"])
    hl.loss.backward()
    ref_grad = torch.cat([torch.zeros(B, 3), hl.p["pred_hand_trans"].grad.view(B, 3), hl.p["pred_right_orient"].grad,
                          hl.p["pred_right_pose_params"].grad, hl.p["pred_left_orient"].grad,
                          hl.p["pred_left_pose_params"].grad, hl.p["pred_right_shape_params"].grad,
                          hl.p["pred_left_shape_params"].grad], 1).numpy()
    opt = H.make_opt(model_root, B)
    opt.sdf_ray_axis = 1
    model = OptimizeModel(opt)
    model.set_input(batch)
    model.init_optimize()
    losses, grad = model.value_and_grad(stage)
    losses, grad = losses.cpu().numpy(), grad.cpu().numpy()
    r = float(hl.collision_loss)
    assert r > 0 and abs(losses[3] - r) <= REL_TOL * abs(r), (losses[3], r)
    live = {0: slice(3, 6), 2: slice(9, 54)}[stage_id]
    assert rel_err(grad[:, live], ref_grad[:, live]) <= REL_TOL
    ref_x = H.oracle_loop(oracle_layers, B, 1, 1)
    ref_x.set_input(batch); ref_x.init_optimize(); ref_x.forward(); ref_x.compute_loss(stage["loss_weights"])
    assert abs(float(ref_x.collision_loss) - r) > 1e-3 * abs(r)          # the +x convention gives another value


def test_current_errors_match_the_reference_log_values(model_root, oracle_layers):
    """get_current_errors (optimize_model.py:438-455): the GT-based log values, including frames whose GT has no right
    wrist (aligned to joint 21 first, loss_utils.py:91-99) or a wrist weight between the two thresholds (not aligned)."""
    from ihmr_b200.optimize_model import OptimizeModel
    B = 4
    data = H.make_batch(oracle_layers[0], 0, B, mode="collision")
    data["joints_3d"] = np.array(data["joints_3d"])
    data["joints_3d"][1, 0, 3] = 0.0          # no right wrist
    data["joints_3d"][2, 0, 3] = 0.3          # neither rule applies
    batch = H.torch_batch(data)
    hl = H.oracle_loop(oracle_layers, B, 1, 1)
    hl.set_input(batch); hl.init_optimize(); hl.forward(); hl.compute_loss(hl.strategy[0]["loss_weights"])
    model = OptimizeModel(H.make_opt(model_root, B))
    model.set_input(batch); model.init_optimize(); model.forward()
    got = model.get_current_errors()
    want = dict(joints_2d_loss=hl.joints_2d_loss, joints_3d_loss=hl.joints_3d_loss, hand_trans_loss=hl.hand_trans_loss)
    for k, v in want.items():
        assert abs(got[k] - float(v)) <= 1e-4 * max(abs(float(v)), 1e-6), (k, got[k], float(v))
    assert list(got) == ["joints_2d_loss", "joints_3d_loss", "hand_trans_loss", "collision_loss", "joints_3d_loss_p"]


# ------------------------------------------------------------------------- whole loop
@pytest.mark.parametrize("fixture", ["loop_b2_short.npz", "loop_collision_short.npz", "loop_cfg1.npz", "loop_mixed6.npz"])
def test_full_loop_vs_golden(model_root, fixture):
    """The fixtures were produced by the UNMODIFIED reference host loop + oracle leaves."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    data, out, epochs, freq = H.load_golden(fixture)
    B = data["init_cam"].shape[0]
    opt = H.make_opt(model_root, B, save_mid_freq=freq, strategy=with_epochs(opt_default, epochs))
    model = OptimizeModel(opt)
    model.set_input(H.torch_batch(data))
    model.init_optimize()
    model.optimize(0, 1)
    res = model.get_pred_result()
    assert list(res.keys()) == list(out.keys())
    for k in out:
        assert res[k].shape == out[k].shape and res[k].dtype == out[k].dtype, k
    check_result_arrays(res, out)


# every one of the 13 exported arrays by value: (absolute tolerance, relative-to-max tolerance)
RESULT_TOL = {
    "pred_cam_params": (0.0, 0.0),                    # never optimised by opt_default: a bit-exact copy of init_cam
    "pred_hand_trans": (JOINT_TOL, 0.0),
    "pred_shape_params": (2e-3, 0.0),                 # (dimensionless coefficients; 2e-3 moves a vertex by < 0.01 mm)
    "pred_pose_params": (2e-3, 0.0),                  # (radians; joints/verts below pin the geometry to 0.1 mm)
    "pred_right_hand_verts": (JOINT_TOL, 0.0),
    "pred_left_hand_verts": (JOINT_TOL, 0.0),
    "mano_params_weight": (0.0, 0.0),                 # input passed through
    "pred_joints_3d": (JOINT_TOL, 0.0),
    "gt_joints_3d": (0.0, 0.0),                       # input passed through
    "collision_loss": (1e-6, 1e-3),                   # sum of ~1e3 per-vertex values of refined (not identical) geometry
    "collision_loss_origin_scale": (JOINT_TOL, 0.0),  # metres
    "do_flip": (0.0, 0.0),
    "pred_hand_type": (0.0, 0.0),
}


def check_result_arrays(res, out):
    assert set(res.keys()) == set(RESULT_TOL.keys())
    for k, (atol, rtol) in RESULT_TOL.items():
        a, b = np.asarray(res[k], np.float64), np.asarray(out[k], np.float64)
        tol = atol + rtol * float(np.abs(b).max())
        assert np.abs(a - b).max() <= tol, (k, float(np.abs(a - b).max()), tol)


@pytest.mark.parametrize("fixture", ["loop_long_typical.npz", "loop_long_collision.npz"])
def test_full_loop_shipped_strategy_length(model_root, fixture):
    """The SHIPPED strategy length (opt_default.py:15,34,53,72: epoch=300 per stage = 1,204 Adam steps;
    bash/optimize.sh:33: save_mid_freq=10 = 31 snapshots per stage), one frame, fixture from the UNMODIFIED
    reference host loop: drift over the long run and the selection among 31 snapshots."""
    path = os.path.join(H.GOLDEN, fixture)
    if not os.path.exists(path):
        pytest.skip(f"{fixture} not generated")
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    data, out, epochs, freq = H.load_golden(fixture)
    assert (epochs, freq) == (300, 10)
    model = OptimizeModel(H.make_opt(model_root, 1, save_mid_freq=freq, strategy=with_epochs(opt_default, epochs)))
    model.set_input(H.torch_batch(data))
    model.init_optimize()
    model.optimize(0, 1)
    check_result_arrays(model.get_pred_result(), out)


def test_dropin_leaves_under_host_loop(model_root, oracle_layers, cuda_layers):
    """L0 boundary: the same host loop drives (i) oracle leaves on CPU and (ii) the CUDA leaves."""
    from ihmr_b200 import sdf_loss
    from oracle import host_loop_oracle as HL
    B, epochs, freq = 2, 2, 1
    batch = H.torch_batch(H.make_batch(oracle_layers[0], 0, B))
    cpu = H.oracle_loop(oracle_layers, B, epochs, freq)
    cpu.set_input(batch); cpu.init_optimize(); cpu.optimize()
    ref = cpu.get_pred_result()
    sdf = sdf_loss.SDFLoss(cuda_layers[0].faces, cuda_layers[1].faces).cuda()
    gpu = HL.HostLoopOracle(cuda_layers[0], cuda_layers[0].faces, cuda_layers[1].faces, sdf, B,
                            strategy=HL.opt_default_strategy(epochs), save_mid_freq=freq, device="cuda")
    gpu.set_input(batch); gpu.init_optimize(); gpu.optimize()
    res = gpu.get_pred_result()
    assert np.abs(res["pred_joints_3d"] - ref["pred_joints_3d"]).max() <= JOINT_TOL
    assert np.abs(res["pred_left_hand_verts"] - ref["pred_left_hand_verts"]).max() <= JOINT_TOL
    assert np.abs(res["collision_loss_origin_scale"] - ref["collision_loss_origin_scale"]).max() <= JOINT_TOL


# ------------------------------------------------------------------ selection (a13) on its own
def _select_gpu(crit, stage):
    """crit (S,B,3) = [joints_3d_loss_p, collision_loss, joints_2d_loss_p] -> chosen snapshot per frame (C ABI)."""
    import ctypes as C
    from ihmr_b200 import _lib
    lib = _lib.load()
    S, B = crit.shape[:2]
    c = torch.tensor(crit, dtype=torch.float32).cuda().contiguous()
    idx = torch.empty(B, dtype=torch.int32, device="cuda")
    st = _lib.make_stage(stage)
    _lib.check(lib.ihmr_select_snapshots(S, B, C.c_void_p(c.data_ptr()), C.byref(st), C.c_void_p(idx.data_ptr()),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "select")
    return idx.cpu().tolist()


def _select_oracle(crit, stage):
    from oracle import host_loop_oracle as HL

    class Fake(HL.HostLoopOracle):
        def __init__(self, B):
            self.B, self.p = B, {}
    S, B = crit.shape[:2]
    f = Fake(B)
    t = torch.tensor(crit, dtype=torch.float32)
    f.snapshots = [{"pred_hand_trans": torch.zeros(B, 1), "joints_3d_loss_p": t[s, :, 0], "collision_loss": t[s, :, 1],
                    "joints_2d_loss_p": t[s, :, 2]} for s in range(S)]
    f._end_stage(dict(stage, update_params=["pred_hand_trans"]))
    return f.last_selected.tolist()


def test_selection_semantics_handcrafted():
    """The device routine of the online selection (ihmr_select_snapshots runs the code ihmr_opt_stage applies after
    every snapshot) on hand-written criteria, against the pinned port of opt_utils.py:104-152: a filter rejects the
    best-scoring snapshot, ties keep the first, a zero origin collision, '+0' = +0.1 %, nothing valid -> snapshot 0."""
    from ihmr_b200.strategies import opt_default
    base = dict(opt_default[1])
    stage = dict(base, filter_loss=[("joints_3d_loss_p", "+0"), ("collision_loss", "-10")], select_loss="joints_3d_loss_p")
    j3d = np.array([[1.0, 1.0, 1.0, 1.0, 1.0], [0.5, 0.7, 0.2, 1.0005, 1.002], [0.4, 0.7, 0.1, 0.999, 0.5], [0.9, 0.9, 0.05, 0.9989, 0.5]])
    col = np.array([[1.0, 0.0, 1.0, 0.0, 2.0], [0.95, 0.0, 1.0, 0.0, 1.0], [0.90, 0.0, 0.95, 0.0, 1.9], [0.80, 0.0, 0.99, 0.0, 1.0]])
    crit = np.stack([j3d, col, np.zeros_like(j3d)], -1)
    want = _select_oracle(crit, stage)
    assert want == [2, 1, 0, 3, 3]     # frame 0: the best score (snapshot 2, col 0.90 of bar 0.901) passes, 0.5 at col 0.95 does not
    assert _select_gpu(crit, stage) == want
    # collision as the selected criterion, 2-D loss as a filter; random criteria incl. exact ties and zeros
    stage2 = dict(base, filter_loss=[("joints_2d_loss_p", "+0"), ("joints_3d_loss_p", "-10")], select_loss="collision_loss")
    rng = np.random.default_rng(0)
    crit = rng.choice([0.0, 0.25, 0.5, 0.9, 0.9009, 0.901, 1.0, 1.0009, 1.001, 1.0011, 2.0], size=(31, 4096, 3)).astype(np.float32)
    for st_ in (stage, stage2):
        assert _select_gpu(crit, st_) == _select_oracle(crit, st_)


def test_results_are_owned_and_pipelining_matches(model_root, oracle_layers):
    """get_pred_result hands out arrays that stay valid after later calls (the reference's evaluator keeps row
    views of them, evaluator.py:74-86), and the pipelined driver returns exactly what the per-batch calls return."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    B = 4
    batches = [H.torch_batch(H.make_batch(oracle_layers[0], s, B, mode=m)) for s, m in ((0, "typical"), (8, "collision"), (16, "typical"))]
    m = OptimizeModel(H.make_opt(model_root, B, save_mid_freq=1, strategy=with_epochs(opt_default, 2), bs_norm=B))
    seq, kept_rows = [], []
    for b in batches:
        m.set_input(b); m.init_optimize(); m.optimize(0, 1)
        r = m.get_pred_result()
        seq.append({k: v.copy() for k, v in r.items()})
        kept_rows.append((r["pred_joints_3d"][1], r["collision_loss_origin_scale"][2]))     # views, as the evaluator keeps
        del r
    for (j, o), ref in zip(kept_rows, seq):
        assert np.array_equal(j, ref["pred_joints_3d"][1]) and np.array_equal(o, ref["collision_loss_origin_scale"][2])
    assert not np.array_equal(seq[0]["pred_joints_3d"], seq[1]["pred_joints_3d"])
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in batches]
    got = list(m.run_pipelined(pinned))
    assert len(got) == len(seq)
    for g, ref in zip(got, seq):
        for k in ref:
            assert np.array_equal(g[k], ref[k]), k


def test_cuda_graph_replay_equals_direct_launches(model_root, oracle_layers):
    """Stages are replayed from CUDA graphs after the first batch (same kernels, same arguments): bitwise equal."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    B = 6
    batches = [H.torch_batch(H.make_batch(oracle_layers[0], s, B, mode=m)) for s, m in ((0, "typical"), (512, "collision"), (30, "typical"))]

    def run(use_graphs):
        opt = H.make_opt(model_root, B, save_mid_freq=2, strategy=with_epochs(opt_default, 5), bs_norm=B)
        opt.use_cuda_graphs = use_graphs
        m = OptimizeModel(opt)
        out = []
        for b in batches:
            m.set_input(b); m.init_optimize(); m.optimize(0, 1)
            out.append({k: v.copy() for k, v in m.get_pred_result().items()})
        return out, m

    direct, _ = run(False)
    graphed, m = run(True)
    assert m.replayed_launches > 0 and len(m._graphs) == 5           # 4 stages + the final forward
    for a, b in zip(direct, graphed):
        for k in a:
            assert np.array_equal(a[k], b[k]), k


def test_determinism_and_shard_invariance(model_root, oracle_layers):
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    B = 8
    data = H.make_batch(oracle_layers[0], 0, B)
    strat = with_epochs(opt_default, 3)

    def run(rows, bs_norm):
        sub = {k: v[rows] for k, v in data.items()}
        m = OptimizeModel(H.make_opt(model_root, len(rows), save_mid_freq=1, strategy=strat, bs_norm=bs_norm))
        m.set_input(H.torch_batch(sub)); m.init_optimize(); m.optimize(0, 1)
        return m.get_pred_result()

    full1, full2 = run(list(range(B)), B), run(list(range(B)), B)
    halves = [run(list(range(0, 4)), B), run(list(range(4, 8)), B)]
    for k in ("pred_pose_params", "pred_shape_params", "pred_hand_trans", "pred_joints_3d", "collision_loss"):
        assert np.array_equal(full1[k], full2[k]), k                       # run-to-run bitwise
        assert np.array_equal(np.concatenate([h[k] for h in halves]), full1[k]), k   # sharding bitwise


def test_stage_shortcuts_match_generic_path(model_root, oracle_layers):
    """The orientation-only (rigid) and shape-only (affine) stage kernels are algebraic rewrites of the
    generic chain: the same loop with IHMR_STAGE_GENERIC_KERNELS on every stage must agree to rounding."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    B = 48
    data = {k: np.concatenate([a, b]) for (k, a), (_, b) in zip(
        H.make_batch(oracle_layers[0], 0, B // 2).items(), H.make_batch(oracle_layers[0], 512, B // 2, mode="collision").items())}
    strat = with_epochs(opt_default, 6)

    def run(strategy):
        m = OptimizeModel(H.make_opt(model_root, B, save_mid_freq=2, strategy=strategy, bs_norm=B))
        m.set_input(H.torch_batch(data)); m.init_optimize(); m.optimize(0, 1)
        return m.get_pred_result()

    fast = run(strat)
    generic = run([dict(st, generic_kernels=True) for st in strat])
    for k in ("pred_joints_3d", "pred_right_hand_verts", "pred_left_hand_verts"):
        assert np.abs(fast[k] - generic[k]).max() <= 2e-5, k
    for k in ("pred_pose_params", "pred_shape_params", "pred_hand_trans"):
        assert np.abs(fast[k] - generic[k]).max() <= 2e-4, k
    assert np.abs(fast["collision_loss"] - generic["collision_loss"]).max() <= 1e-4 * max(1.0, np.abs(generic["collision_loss"]).max())


# ------------------------------------------------- less-travelled options of the reference
def test_sgd_optimizer_matches_oracle_loop(model_root, oracle_layers):
    """opt.optimizer == 'sgd' (optimize_model.py:345-347): SGD with momentum 0.9."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    from oracle import host_loop_oracle as HL
    from oracle import sdf_oracle
    B, epochs, freq = 2, 3, 1
    batch = H.torch_batch(H.make_batch(oracle_layers[0], 2, B))
    sdf = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces)
    cpu = HL.HostLoopOracle(oracle_layers[0], oracle_layers[0].faces, oracle_layers[1].faces, sdf, B,
                            strategy=HL.opt_default_strategy(epochs), save_mid_freq=freq, optimizer="sgd")
    cpu.set_input(batch); cpu.init_optimize(); cpu.optimize()
    ref = cpu.get_pred_result()
    opt = H.make_opt(model_root, B, save_mid_freq=freq, strategy=with_epochs(opt_default, epochs))
    opt.optimizer = "sgd"
    m = OptimizeModel(opt)
    m.set_input(batch); m.init_optimize(); m.optimize(0, 1)
    res = m.get_pred_result()
    assert np.abs(res["pred_joints_3d"] - ref["pred_joints_3d"]).max() <= JOINT_TOL
    assert np.abs(res["pred_pose_params"] - ref["pred_pose_params"]).max() <= 1e-4


def test_sdf_robustifier_matches_oracle(cuda_layers, oracle_layers):
    """robustifier > 0 is the training-time setting (loss_utils.py:36); rho(x) = (x/r)^2 / ((x/r)^2 + 1)."""
    from ihmr_b200 import sdf_loss, synthetic
    from oracle import mano_oracle, sdf_oracle
    raw = synthetic.make_raw_frames(0, 3, seed=0, mode="collision")
    with torch.no_grad():
        rv, lv, _ = mano_oracle.two_hand_forward(oracle_layers[0], torch.tensor(raw["true_pose"]),
                                                 torch.tensor(raw["true_shape"]), torch.tensor(raw["true_trans"]))
    hv_cpu = torch.stack([rv, lv], 1).clone().requires_grad_(True)
    l_ref, pv_ref, o_ref = sdf_oracle.SDFLoss(oracle_layers[0].faces, oracle_layers[1].faces, robustifier=0.05)(hv_cpu, True, True)
    l_ref.sum().backward()
    hv = hv_cpu.detach().cuda().requires_grad_(True)
    l, pv, o = sdf_loss.SDFLoss(cuda_layers[0].faces, cuda_layers[1].faces, robustifier=0.05).cuda()(
        hv, return_per_vert_loss=True, return_origin_scale_loss=True)
    l.sum().backward()
    assert rel_err(l.detach().cpu().numpy(), l_ref.detach().numpy()) <= REL_TOL
    assert float((pv.cpu() - pv_ref.detach()).abs().max()) <= 1e-4 * float(pv_ref.max())
    assert float((o.cpu() - o_ref).abs().max()) <= 2e-6
    assert rel_err(hv.grad.cpu().numpy(), hv_cpu.grad.numpy()) <= REL_TOL


def test_camera_gradient_and_stage(model_root, oracle_layers):
    """The commented-out camera stage of opt_default.py:81-97 is still expressible: 'pred_cam_params'
    as update target. Its gradient flows through the 2-D term only."""
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default
    B = 2
    batch = H.torch_batch(H.make_batch(oracle_layers[0], 0, B))
    stage = dict(opt_default[1], update_params=["pred_cam_params"], filter_loss=[("joints_2d_loss_p", "+0")],
                 select_loss="joints_2d_loss_p", epoch=5, lr=1e-2)
    hl = H.oracle_loop(oracle_layers, B, 1, 1)
    hl.set_input(batch); hl.init_optimize()
    hl.pred_cam_params = hl.pred_cam_params.clone().requires_grad_(True)
    hl.forward(); hl.compute_loss(stage["loss_weights"]); hl.loss.backward()
    m = OptimizeModel(H.make_opt(model_root, B, save_mid_freq=1))
    m.set_input(batch); m.init_optimize()
    _, grad = m.value_and_grad(stage)
    assert rel_err(grad[:, 0:3].cpu().numpy(), hl.pred_cam_params.grad.numpy()) <= REL_TOL
    before = m.value_and_grad(stage)[0][0].item()
    m.run_stage(stage)
    after = m.value_and_grad(stage)[0][0].item()
    assert after <= before * 1.001                        # selection never accepts a worse 2-D loss


def test_optimize_dataset_driver_writes_the_reference_result_file(model_root, tmp_path):
    """The caller side (src/optimize.py:40-102): dataset records -> pipelined refinement -> Evaluator records ->
    evaluate_results/optimize/<dataset>.pkl, and the records equal what the per-batch calls produce."""
    from ihmr_b200.evaluator import Evaluator
    from ihmr_b200.opt_dataset import OPTDataset
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.run_optimize import optimize_dataset
    from ihmr_b200.strategies import opt_default, with_epochs
    from tests.test_dataset_io import _write_dataset
    opt, info = _write_dataset(str(tmp_path), 6)
    opt.model_root, opt.strategy, opt.save_mid_freq, opt.opt_dataset = model_root, with_epochs(opt_default, 2), 1, "synthetic_ds"
    out_dir = os.path.join(str(tmp_path), "evaluate_results", "optimize")
    metrics = optimize_dataset(opt, info, out_dir=out_dir)
    assert set(metrics) == {"mpjpe_3d", "inter_mpjpe_3d", "collision_ave", "collision_max"} and all(np.isfinite(v) for v in metrics.values())
    ev = Evaluator.load(os.path.join(out_dir, "synthetic_ds.pkl"))
    assert len(ev.pred_results) == 6 and [r["data_idx"] for r in ev.pred_results] == list(range(6))
    ds = OPTDataset(opt, info)
    ds.load_data()
    m = OptimizeModel(opt)
    b0 = next(ds.batches(pin=False))
    m.set_input(b0); m.init_optimize(); m.optimize(0, 1)
    res = m.get_pred_result()
    for i in range(4):
        assert np.array_equal(ev.pred_results[i]["pred_joints_3d"], res["pred_joints_3d"][i])
        assert np.array_equal(ev.pred_results[i]["pred_pose_params"], res["pred_pose_params"][i])
