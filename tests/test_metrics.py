"""Evaluator metrics: oracle vs the reference's own metric_utils (golden), device kernel vs both."""
import os

import numpy as np
import pytest
import torch

from oracle import metrics_oracle as MO
from tests import helpers as H


def golden():
    return np.load(os.path.join(H.GOLDEN, "metrics.npz"))


def test_oracle_reproduces_reference_metric_utils():
    z = golden()
    t = MO.frame_table(z["pred"], z["gt"], z["origin"], z["scale"])
    assert np.array_equal(t[:, [1, 3]], z["table"][:, [1, 3]])                 # counts
    assert np.abs(t[:, [0, 2]] - z["table"][:, [0, 2]]).max() <= 1e-6
    assert np.allclose(t[:, 4], z["origin"].mean(1)) and np.allclose(t[:, 5], z["origin"].max(1))


@pytest.mark.gpu
def test_device_metrics_match_reference_golden():
    from ihmr_b200 import evaluator
    z = golden()
    t = evaluator.frame_metrics(torch.tensor(z["pred"]).cuda(), torch.tensor(z["gt"]).cuda(),
                                torch.tensor(z["origin"]).cuda(), torch.tensor(z["scale"]).cuda()).cpu().numpy()
    assert np.array_equal(t[:, [1, 3]], z["table"][:, [1, 3]])
    assert np.abs(t[:, [0, 2]] - z["table"][:, [0, 2]]).max() <= 1e-5 * z["table"][:, [0, 2]].max()
    assert np.abs(t[:, 4] - z["origin"].mean(1)).max() <= 1e-9 and np.array_equal(t[:, 5], z["origin"].max(1))


@pytest.mark.gpu
def test_device_evaluator_end_to_end(model_root, oracle_layers):
    """The four numbers src/optimize.py:99-102 prints, from the device table vs the oracle on the
    arrays get_pred_result exports."""
    from ihmr_b200.evaluator import DeviceEvaluator
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    B = 6
    data = H.make_batch(oracle_layers[0], 0, B)
    m = OptimizeModel(H.make_opt(model_root, B, save_mid_freq=1, strategy=with_epochs(opt_default, 2)))
    m.set_input(H.torch_batch(data)); m.init_optimize(); m.optimize(0, 1)
    ev = DeviceEvaluator()
    ev.update(data["index"], m)
    ev.update(data["index"], m)          # padded duplicates, removed like the reference does
    ev.remove_redunc()
    res = m.get_pred_result()
    want = MO.summary(MO.frame_table(res["pred_joints_3d"], res["gt_joints_3d"], res["collision_loss_origin_scale"]))
    for k, v in want.items():
        assert abs(getattr(ev, k) - v) <= 1e-5 * max(abs(v), 1e-6), k


@pytest.mark.gpu
def test_device_evaluator_reads_scale_and_hand_type_from_the_data_list(model_root, oracle_layers):
    """With the dataset's records the device evaluator takes `scale` and `hand_type` per sample as the reference's
    Evaluator does (evaluator.py:48-58) and de-duplicates by image path."""
    from ihmr_b200.evaluator import DeviceEvaluator
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    B = 6
    data = H.make_batch(oracle_layers[0], 0, B)
    m = OptimizeModel(H.make_opt(model_root, B, save_mid_freq=1, strategy=with_epochs(opt_default, 2)))
    m.set_input(H.torch_batch(data)); m.init_optimize(); m.optimize(0, 1)
    scales = [1.0, 0.5, 2.0, 1.5, 1.0, 0.8]
    types = ["interacting", "right", "interacting", "left", "interacting", "interacting"]
    data_list = [dict(img_path=f"img_{i % 5}.jpg", scale=scales[i], hand_type=types[i]) for i in range(B)]   # 5 == 0: duplicate
    ev = DeviceEvaluator(data_list)
    ev.update(np.arange(B), m)
    ev.remove_redunc()
    res = m.get_pred_result()
    keep = np.arange(5)
    tab = MO.frame_table(res["pred_joints_3d"][keep], res["gt_joints_3d"][keep], res["collision_loss_origin_scale"][keep],
                         scale=np.array(scales)[keep])
    want = MO.summary(tab, interacting=np.array([t == "interacting" for t in types])[keep])
    for k, v in want.items():
        assert abs(getattr(ev, k) - v) <= 1e-5 * max(abs(v), 1e-6), k
