"""Pins oracle/mlp_oracle.py: the sub-network against tests/golden/mlp.npz (outputs of the reference's own
InterHandSubNetwork) and live against that class; the selection rule against the reference's own, unmodified
MLPModel.select_better_params driven on a stand-in object."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import mlp_oracle as MO
from oracle import ref_shims
from tests import helpers as H


@pytest.mark.parametrize("dim", [3, 90])
def test_subnetwork_matches_reference_golden(dim):
    z = np.load(os.path.join(H.GOLDEN, "mlp.npz"))
    net = MO.SubNetworkOracle(dim)
    net.load_state_dict(MO.seeded_state_dict(dim, seed=dim))
    with torch.no_grad():
        y = net(torch.from_numpy(z[f"d{dim}_x"]))
    assert np.abs(y.numpy() - z[f"d{dim}_y"]).max() <= 1e-6


def _reference_modules(monkeypatch):
    ref_shims._install_shims()
    monkeypatch.syspath_prepend(ref_shims.REFERENCE_SRC)
    for name in [n for n in sys.modules if n in ("models", "utils") or n.startswith("models.") or n.startswith("utils.")]:
        monkeypatch.delitem(sys.modules, name)


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
def test_subnetwork_matches_reference_class_live(monkeypatch):
    _reference_modules(monkeypatch)
    from models.networks import InterHandSubNetwork          # the reference's file
    ref = InterHandSubNetwork(None, 1024 + 122, 45)
    net = MO.SubNetworkOracle(45)
    net.load_state_dict(ref.state_dict())                    # same parameter names and shapes
    x = torch.randn(7, 1146)
    with torch.no_grad():
        assert torch.equal(net(x), ref(x))


def _criteria(seed, B):
    g = torch.Generator().manual_seed(seed)
    vals = torch.tensor([0.0, 0.5, 0.999, 1.0, 1.0001, 1.001, 2.0])
    pick = lambda: vals[torch.randint(0, len(vals), (B,), generator=g)]
    return {k: pick() for k in ("joints_3d_loss_p", "collision_loss", "joints_2d_loss_p")}


@pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("stage_id", [0, 3, 5])
def test_selection_rule_matches_reference_method(monkeypatch, stage_id):
    """MLPModel.select_better_params (mlp_model.py:592-637), unmodified, called on a stand-in that carries exactly the
    attributes it reads: which frames keep their new parameters must equal oracle.mlp_oracle.select_better."""
    _reference_modules(monkeypatch)
    from models.mlp_model import MLPModel                    # the reference's file
    from strategies.mlp_default import mlp_default
    B = 64
    stage = mlp_default[stage_id]
    cur, prev = _criteria(1, B), _criteria(2, B)
    fake = types.SimpleNamespace(strategy=mlp_default, batch_size=B, data_idxs=torch.arange(B),
                                 data_idxs_all=torch.ones(B, dtype=torch.bool),
                                 update_loss_name_list=set(cur.keys()),
                                 prev_losses={f"prev_{k}_batch": v.clone() for k, v in prev.items()}, prev_params={})
    for k, v in cur.items():
        setattr(fake, f"{k}_batch", v.clone())
    for name in stage["update_params"]:
        d = MO.PARAM_DIMS[name]
        setattr(fake, name, torch.ones(B, d))                                  # new parameters = 1, previous = 0
        fake.prev_params[name.replace("pred_", "prev_")] = torch.zeros(B, d)
    setattr(fake, "_MLPModel__gather_params", lambda: None)
    MLPModel.select_better_params(fake, stage_id)
    kept_ref = getattr(fake, stage["update_params"][0])[:, 0] > 0.5
    kept = MO.select_better(cur, prev, stage)
    assert torch.equal(kept, kept_ref)
    assert 0 < int(kept.sum()) < B                                             # both branches taken
    for k in cur:                                                              # rejected frames fall back to the previous losses
        assert torch.equal(getattr(fake, f"{k}_batch"), torch.where(kept, cur[k], prev[k]))
