"""Pins oracle/host_loop_oracle.py: (i) against the committed golden fixtures, which were
produced by the UNMODIFIED reference host loop (tests/golden/make_golden.py), and (ii) against
that reference loop run live wherever /root/reference exists."""
import numpy as np
import pytest
import torch

from oracle import host_loop_oracle as HL
from oracle import ref_shims
from tests import helpers as H

KEYS = ("pred_cam_params", "pred_hand_trans", "pred_shape_params", "pred_pose_params", "pred_right_hand_verts",
        "pred_left_hand_verts", "mano_params_weight", "pred_joints_3d", "gt_joints_3d", "collision_loss",
        "collision_loss_origin_scale", "do_flip", "pred_hand_type")


@pytest.mark.parametrize("fixture", ["loop_b2_short.npz", "loop_collision_short.npz"])
def test_port_reproduces_reference_golden(oracle_layers, fixture):
    data, out, epochs, freq = H.load_golden(fixture)
    B = data["init_cam"].shape[0]
    loop = H.oracle_loop(oracle_layers, B, epochs, freq)
    loop.set_input(H.torch_batch(data))
    loop.init_optimize()
    loop.optimize()
    res = loop.get_pred_result()
    assert tuple(res.keys()) == KEYS == tuple(out.keys())
    for k in KEYS:
        assert res[k].shape == out[k].shape and res[k].dtype == out[k].dtype, k
        assert np.abs(res[k].astype(np.float64) - out[k]).max() <= 1e-6, k
    # shapes the evaluator relies on (a14)
    assert res["pred_pose_params"].shape == (B, 96) and res["collision_loss_origin_scale"].shape == (B, 1556)
    assert np.all(res["pred_joints_3d"][:, 0] == 0.0)          # root aligned in place (Appendix D.3)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_port_matches_unmodified_reference_live(model_root, oracle_layers):
    B, epochs, freq = 1, 2, 1
    data = H.make_batch(oracle_layers[0], 5, B)
    batch = H.torch_batch(data)
    ref = ref_shims.load_reference_model(ref_shims.make_opt(model_root, B, save_mid_freq=freq), epochs=epochs)
    ref.set_input(batch); ref.init_optimize(); ref.optimize(0, 1)
    want = ref.get_pred_result()
    loop = H.oracle_loop(oracle_layers, B, epochs, freq)
    loop.set_input(batch); loop.init_optimize(); loop.optimize()
    got = loop.get_pred_result()
    for k in KEYS:
        assert np.abs(got[k].astype(np.float64) - want[k]).max() <= 1e-6, k


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_log_values_match_the_reference_get_current_errors(model_root, oracle_layers):
    """The GT-based log values of __compute_loss (optimize_model.py:276-306) that get_current_errors (:438-455) returns,
    from the UNMODIFIED reference vs the port — with a frame whose GT has no right wrist (aligned to joint 21,
    loss_utils.py:96-99) and one whose wrist weight falls between the two thresholds (not aligned at all).  The CUDA
    model's get_current_errors is checked against the port on the GPU (test_current_errors_match_the_reference_log_values)."""
    B = 3
    data = H.make_batch(oracle_layers[0], 0, B, mode="collision")
    data["joints_3d"] = np.array(data["joints_3d"])
    data["joints_3d"][1, 0, 3] = 0.0
    data["joints_3d"][2, 0, 3] = 0.3
    batch = H.torch_batch(data)
    ref = ref_shims.load_reference_model(ref_shims.make_opt(model_root, B, save_mid_freq=1), epochs=1)
    ref.set_input(batch); ref.init_optimize(); ref.forward()
    weights = ref.strategy[0]["loss_weights"]
    getattr(ref, "_OptimizeModel__compute_loss")(weights)
    want = ref.get_current_errors()
    loop = H.oracle_loop(oracle_layers, B, 1, 1)
    loop.set_input(batch); loop.init_optimize(); loop.forward(); loop.compute_loss(loop.strategy[0]["loss_weights"])
    got = dict(joints_2d_loss=loop.joints_2d_loss, joints_3d_loss=loop.joints_3d_loss, hand_trans_loss=loop.hand_trans_loss,
               collision_loss=loop.collision_loss, joints_3d_loss_p=loop.joints_3d_loss_p)
    assert list(want) == list(got)
    for k, v in want.items():
        assert abs(float(got[k]) - float(v)) <= 1e-5 * max(abs(float(v)), 1e-6), (k, float(got[k]), float(v))


def test_bs_norm_makes_a_shard_equal_to_the_full_batch(oracle_layers):
    """1/bs inside every mean interacts with Adam's eps: a shard must normalise by the full size."""
    data = H.make_batch(oracle_layers[0], 0, 2)
    full = H.oracle_loop(oracle_layers, 2, 1, 1)
    full.set_input(H.torch_batch(data)); full.init_optimize(); full.optimize()
    a = full.get_pred_result()
    part = H.oracle_loop(oracle_layers, 1, 1, 1, bs_norm=2)
    part.set_input(H.torch_batch({k: v[1:2] for k, v in data.items()})); part.init_optimize(); part.optimize()
    b = part.get_pred_result()
    assert np.abs(a["pred_pose_params"][1] - b["pred_pose_params"][0]).max() <= 1e-6
    assert np.abs(a["pred_joints_3d"][1] - b["pred_joints_3d"][0]).max() <= 1e-6


# -------------------------------------------------- snapshot filtering / selection semantics
class _Fake(HL.HostLoopOracle):
    def __init__(self, B):
        self.B = B
        self.p = {}


def _select(j3d, col, params):
    """Runs _end_stage on hand-written snapshots: j3d/col (S,B), params (S,B,1)."""
    S, B = j3d.shape
    f = _Fake(B)
    f.snapshots = [{"pred_hand_trans": params[s], "joints_3d_loss_p": j3d[s], "collision_loss": col[s]} for s in range(S)]
    stage = dict(update_params=["pred_hand_trans"], filter_loss=[("joints_3d_loss_p", "+0"), ("collision_loss", "-10")],
                 select_loss="joints_3d_loss_p")
    f._end_stage(stage)
    return f.last_selected.tolist()


def test_filter_thresholds_and_first_minimum():
    t = torch.tensor
    params = torch.arange(4.0).view(4, 1, 1).repeat(1, 3, 1)
    #            frame0: col must drop 10 % (bar 0.901)   frame1: tie -> first   frame2: nothing valid -> 0
    j3d = t([[1.0, 1.0, 1.0], [0.5, 0.7, 0.2], [0.4, 0.7, 0.1], [0.9, 0.9, 0.05]])
    col = t([[1.0, 0.0, 1.0], [0.95, 0.0, 1.0], [0.90, 0.0, 0.95], [0.80, 0.0, 0.99]])
    assert _select(j3d, col, params) == [2, 1, 0]
    # '+0' means up to +0.1 %: 1.001 passes, 1.002 does not; col 0 stays 0 -> bar 0 -> valid
    j3d = t([[1.0], [1.0005]])
    col = t([[0.0], [0.0]])
    assert _select(j3d, col, params[:2, :1]) == [0]          # valid but not better than snapshot 0
    assert HL.INVALID_CRITERIA == ("joints_3d_loss", "joints_2d_loss", "hand_trans_loss")


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_unmodified_reference_class_binds_to_the_cuda_leaves(model_root, oracle_layers, monkeypatch):
    """Boundary L0 under the reference's OWN class: /root/reference/src/models/optimize_model.py,
    unmodified, with sys.modules['smplx'] / ['sdf'] set to the ihmr_b200 leaf modules.  This container has the
    reference but no GPU, and the GPU box has no reference, so the two can never compute together; what is checked
    here is everything up to the first kernel: the constructor path (smplx.create signature, the in-place
    `.shapedirs` edit of optimize_model.py:109-113, `.faces`, `SDFLoss(faces_right, faces_left, robustifier=None)`,
    `.cuda()`), set_input / init_optimize, and that forward() reaches OUR leaf, which refuses to run on a CPU tensor
    (no fallback) instead of silently computing something else."""
    import sys

    import ihmr_b200.mano_layer as mano_layer
    import ihmr_b200.sdf_loss as sdf_loss
    from ihmr_b200 import _lib
    ref_shims._install_shims()                       # ry_utils / opendr stand-ins, .cuda() -> identity on this CPU box
    monkeypatch.setitem(sys.modules, "smplx", mano_layer)
    monkeypatch.setitem(sys.modules, "sdf", sdf_loss)
    for name in [n for n in sys.modules if n == "models" or n.startswith("models.")]:
        monkeypatch.delitem(sys.modules, name)       # re-import the reference modules against the new leaves
    monkeypatch.syspath_prepend(ref_shims.REFERENCE_SRC)
    from models.optimize_model import OptimizeModel  # the reference's file, unmodified
    B = 2
    model = OptimizeModel(ref_shims.make_opt(model_root, B, save_mid_freq=1))
    assert isinstance(model.mano_models["right"], mano_layer.ManoLayer)
    assert isinstance(model.loss_util.sdf_loss, sdf_loss.SDFLoss) if hasattr(model, "loss_util") else True
    assert model.mano_models["left"].faces.shape == (1538, 3)
    model.set_input(H.torch_batch(H.make_batch(oracle_layers[0], 0, B)))
    model.init_optimize()
    with pytest.raises(_lib.IhmrError, match="no CPU path"):
        model.forward()
