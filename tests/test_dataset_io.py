"""Input records and result files either side of the path: ihmr_b200.opt_dataset against the
reference's own OPTDataset, ihmr_b200.evaluator.Evaluator against the reference's own Evaluator (both imported live
where /root/reference exists) and against the metric fixture generated from the reference (tests/golden/metrics.npz)."""
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from oracle import ref_shims
from tests import helpers as H


def _write_dataset(root, n, seed=0):
    """A tiny on-disk dataset in the reference's layout: annotation pkl, prior-prediction pkl, per-sample
    parameter pkls, images of different sizes (only their height / width matter)."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "data", "images", "seq"), exist_ok=True)
    os.makedirs(os.path.join(root, "params"), exist_ok=True)
    anno, pred = [], {}
    for i in range(n):
        img_path = f"seq/img_{i:03d}.png"
        h, w = int(rng.integers(120, 400)), int(rng.integers(120, 400))
        Image.fromarray(np.zeros((h, w, 3), np.uint8)).save(os.path.join(root, "data", "images", img_path))
        j3d = rng.normal(0, 0.05, (42, 4)).astype(np.float32)
        j3d[:, 3] = (rng.random(42) > 0.2).astype(np.float32)
        j3d[0, 3] = j3d[21, 3] = 1.0
        if i % 4 == 1:
            j3d[0, 3] = 0.0                                  # right wrist missing: no translation target
        rec = dict(img_path=img_path, param_path=f"p_{i:03d}.pkl", hand_type=("interacting", "right", "left")[i % 3] if i % 5 == 4 else "interacting",
                   hand_type_valid=1.0, joints_2d=(rng.random((42, 2)) * np.array([w, h])).astype(np.float32), joints_3d=j3d)
        if i % 2:
            rec["scale"] = float(rng.uniform(0.8, 1.2))
        anno.append(rec)
        hp = lambda: dict(pose=rng.normal(0, 0.3, 48).astype(np.float32), shape=rng.normal(0, 1, 10).astype(np.float32))
        with open(os.path.join(root, "params", rec["param_path"]), "wb") as fh:
            pickle.dump(dict(right_hand_param=hp(), left_hand_param=None if i % 3 == 2 else hp()), fh)
        pred[img_path] = dict(pred_cam_params=rng.normal(0, 1, 3).astype(np.float32), pred_shape_params=rng.normal(0, 1, 20).astype(np.float32),
                              pred_pose_params=rng.normal(0, 0.3, 96).astype(np.float32), pred_hand_trans=rng.normal(0, 0.05, 3).astype(np.float32),
                              joints_2d=(rng.random((42, 2)) * np.array([w, h])).astype(np.float32),
                              joints_3d=rng.normal(0, 0.05, (42, 3)).astype(np.float32), img_feat=rng.normal(0, 1, 8).astype(np.float32))
    with open(os.path.join(root, "data", "anno.pkl"), "wb") as fh:
        pickle.dump(anno, fh)
    with open(os.path.join(root, "data", "pred.pkl"), "wb") as fh:
        pickle.dump(pred, fh)
    opt = H.make_opt(root, 4)
    opt.data_root, opt.param_root = os.path.join(root, "data"), os.path.join(root, "params")
    return opt, ("synthetic", "anno.pkl", "pred.pkl", "images")


def test_dataset_padding_batches_and_keys(tmp_path):
    from ihmr_b200.opt_dataset import BATCH_KEYS, OPTDataset
    opt, info = _write_dataset(str(tmp_path), 10)
    ds = OPTDataset(opt, info)
    ds.load_data(world_size=2)                 # 10 samples -> padded to 16 = 2 ranks x 2 batches x 4
    assert len(ds) == 16 and ds.num_add == 6 and ds.data_list[10] is ds.data_list[0]
    got = [b for r in range(2) for b in ds.batches(r, 2, pin=False)]
    assert len(got) == 4 and all(tuple(b.keys()) == BATCH_KEYS for b in got)
    assert torch.cat([b["index"] for b in got]).tolist() == list(range(16))          # contiguous blocks per rank
    b = got[0]
    assert b["joints_2d"].shape == (4, 42, 3) and b["joints_3d"].shape == (4, 42, 4) and b["hand_trans"].shape == (4, 1, 4)
    assert b["init_hand_trans_j"].shape == (4, 1, 4) and b["mano_pose"].shape == (4, 96) and b["init_cam"].dtype == torch.float32
    for i in range(4):                          # a batch row is the single sample
        s = ds[i]
        for k in BATCH_KEYS:
            assert torch.equal(b[k][i].to(s[k].dtype), s[k]), k
    assert float(b["hand_trans"][1, 0, 3]) == 0.0 and float(b["hand_trans"][0, 0, 3]) == 1.0   # missing wrist -> weight 0


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_dataset_matches_the_reference_class(tmp_path, monkeypatch):
    """Sample by sample against /root/reference/src/data/opt_dataset.py, unmodified (its unavailable imports are
    stubbed: torchgeometry is only needed by augmentation code the refinement path never calls)."""
    from ihmr_b200.opt_dataset import BATCH_KEYS, OPTDataset
    opt, info = _write_dataset(str(tmp_path), 7)
    ref_shims._install_shims()
    ry = sys.modules["ry_utils"]
    ry.load_pkl = lambda p: pickle.load(open(p, "rb"))
    monkeypatch.setitem(sys.modules, "torchgeometry", types.ModuleType("torchgeometry"))
    monkeypatch.syspath_prepend(ref_shims.REFERENCE_SRC)
    for name in [n for n in sys.modules if n == "data" or n.startswith("data.")]:
        monkeypatch.delitem(sys.modules, name)
    from data.opt_dataset import OPTDataset as RefDataset      # the reference's file
    opt.dist, opt.model_type, opt.use_motion_blur = False, "opt", False       # fields DataProcessor's constructor reads
    ref = RefDataset(opt, info)
    ref.load_data()
    ours = OPTDataset(opt, info)
    ours.load_data()
    assert len(ref) == len(ours) == 8 and ref.num_add == ours.num_add == 1
    for i in range(len(ref)):
        a, b = ref[i], ours[i]
        assert set(a.keys()) == set(BATCH_KEYS) == set(b.keys())
        for k in BATCH_KEYS:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, (k, a[k].dtype, b[k].dtype)
            assert torch.allclose(a[k].double(), b[k].double(), atol=1e-6, rtol=0), k


def _fake_results(n, seed=3):
    z = np.load(os.path.join(H.GOLDEN, "metrics.npz"))
    B = z["pred"].shape[0]
    idx = np.arange(n) % B
    rng = np.random.default_rng(seed)
    res = dict(pred_cam_params=rng.normal(0, 1, (n, 3)).astype(np.float32), pred_shape_params=rng.normal(0, 1, (n, 20)).astype(np.float32),
               pred_pose_params=rng.normal(0, 0.3, (n, 96)).astype(np.float32), pred_hand_trans=rng.normal(0, 0.05, (n, 1, 3)).astype(np.float32),
               pred_joints_3d=z["pred"][idx].astype(np.float32), gt_joints_3d=z["gt"][idx].astype(np.float32),
               collision_loss_origin_scale=z["origin"][idx].astype(np.float32),
               pred_right_hand_verts=rng.normal(0, 0.05, (n, 778, 3)).astype(np.float32),
               pred_left_hand_verts=rng.normal(0, 0.05, (n, 778, 3)).astype(np.float32),
               do_flip=np.zeros(n, np.int32), pred_hand_type=np.ones(n, np.int32))
    return res, z, idx


class _Model:
    inputSize = 224
    mano_models = dict(left=types.SimpleNamespace(faces=np.zeros((1538, 3), np.int64)), right=types.SimpleNamespace(faces=np.ones((1538, 3), np.int64)))


def test_evaluator_records_metrics_and_file_roundtrip(tmp_path):
    from ihmr_b200.evaluator import Evaluator
    from ihmr_b200.opt_dataset import OPTDataset
    from oracle import metrics_oracle as MO
    opt, info = _write_dataset(str(tmp_path), 6)
    ds = OPTDataset(opt, info)
    ds.load_data()                                  # 6 -> 8 (two copies of sample 0)
    res, z, idx = _fake_results(8)
    ev = Evaluator(opt, ds, _Model())
    ev.update(np.arange(4), {k: v[:4] for k, v in res.items()})
    ev.update(np.arange(4, 8), {k: v[4:] for k, v in res.items()})
    assert len(ev.pred_results) == 8
    ev.remove_redunc()
    assert len(ev.pred_results) == 6                # the padded duplicates of sample 0 are gone
    rec = ev.pred_results[1]
    assert rec["img_path_relative"] == "seq/img_001.png" and rec["pred_right_hand_verts"].dtype == np.float16
    assert rec["scale"] == ds.data_list[1]["scale"] and ev.pred_results[0]["scale"] == 1.0
    # per-sample error lists against the pinned restatement of the reference's metric functions
    for r in ev.pred_results:
        i = idx[r["data_idx"]]
        assert np.allclose(r["j3d_error"], MO.joints_error(z["pred"][i], z["gt"][i, :, :3], z["gt"][i, :, 3:], r["scale"]), atol=1e-7)
        assert np.allclose(r["pa_no_rot_inter_j3d_error"], MO.pa_no_rot_error(z["pred"][i], z["gt"][i, :, :3], z["gt"][i, :, 3:], r["scale"]), atol=1e-6)
    before = {m: getattr(ev, m) for m in ("mpjpe_3d", "inter_mpjpe_3d", "collision_ave", "collision_max")}
    path = os.path.join(str(tmp_path), "evaluate_results", "optimize", "synthetic.pkl")
    ev.save(path)
    assert "utils.evaluator" not in sys.modules or not hasattr(sys.modules["utils.evaluator"], "Evaluator") or ref_shims.reference_available()
    back = Evaluator.load(path)
    assert {m: getattr(back, m) for m in before} == before and back.dataset_name == "synthetic"
    assert b"utils.evaluator" in open(path, "rb").read()[:200]          # pickled as the reference's class


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_evaluator_matches_the_reference_class_and_its_pickle(tmp_path, monkeypatch):
    """Same updates through /root/reference/src/utils/evaluator.py (unmodified): identical records and metrics, and the
    file ours writes opens as the reference's own Evaluator."""
    from ihmr_b200.evaluator import Evaluator
    from ihmr_b200.opt_dataset import OPTDataset
    opt, info = _write_dataset(str(tmp_path), 6)
    ds = OPTDataset(opt, info)
    ds.load_data()
    res, _, _ = _fake_results(8)
    res["do_flip"][2] = 1                            # exercises the flip-back branch (the reference's needs all four
    res["gt_right_hand_verts"] = res["pred_left_hand_verts"] * 0.5      # vertex arrays there; do_flip is 0 on the OPT path)
    res["gt_left_hand_verts"] = res["pred_right_hand_verts"] * 0.5
    ref_shims._install_shims()
    monkeypatch.syspath_prepend(ref_shims.REFERENCE_SRC)
    for name in [n for n in sys.modules if n == "utils" or n.startswith("utils.")]:
        monkeypatch.delitem(sys.modules, name)
    from utils.evaluator import Evaluator as RefEvaluator      # the reference's file
    ours, ref = Evaluator(opt, ds, _Model()), RefEvaluator(opt, ds, _Model())
    for ev in (ours, ref):
        ev.update(np.arange(8), {k: v.copy() for k, v in res.items()})
        ev.remove_redunc()
    assert len(ours.pred_results) == len(ref.pred_results) == 6
    for a, b in zip(ours.pred_results, ref.pred_results):
        assert set(a.keys()) == set(b.keys())
        for k in a:
            if isinstance(b[k], np.ndarray):
                assert np.array_equal(np.asarray(a[k]), b[k]), k
            elif isinstance(b[k], list):
                assert np.allclose(a[k], b[k], atol=1e-7), k
            else:
                assert a[k] == b[k], k
    for m in ("mpjpe_3d", "inter_mpjpe_3d", "collision_ave", "collision_max"):
        assert abs(getattr(ours, m) - getattr(ref, m)) <= 1e-9 * max(1.0, abs(getattr(ref, m))), m
    path = os.path.join(str(tmp_path), "out.pkl")
    ours.save(path)
    loaded = pickle.load(open(path, "rb"))
    assert type(loaded) is RefEvaluator and abs(loaded.mpjpe_3d - ref.mpjpe_3d) <= 1e-12
