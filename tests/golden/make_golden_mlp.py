"""Generates tests/golden/mlp.npz from the reference's own InterHandSubNetwork class
(/root/reference/src/models/networks.py:83-105, imported unmodified): for update dims 3 and 90 a seeded state dict,
inputs (5,1146) and the class's outputs (the state dicts come from oracle.mlp_oracle.seeded_state_dict, a numpy recipe).  Run once in the authoring container:  python tests/golden/make_golden_mlp.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims          # noqa: E402

ref_shims._install_shims()
sys.path.insert(0, ref_shims.REFERENCE_SRC)
from models.networks import InterHandSubNetwork      # noqa: E402  (the reference's file)

from oracle import mlp_oracle as MO                # noqa: E402

out = {}
rng = np.random.RandomState(7)
for dim in (3, 90):
    net = InterHandSubNetwork(None, 1024 + 122, dim)
    net.load_state_dict(MO.seeded_state_dict(dim, seed=dim))          # weights by recipe: only inputs / outputs are stored
    x = torch.from_numpy((rng.standard_normal((5, 1024 + 122)) * 0.5).astype(np.float32))
    with torch.no_grad():
        y = net(x)
    out[f"d{dim}_x"], out[f"d{dim}_y"] = x.numpy(), y.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mlp.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith("_y")}, float(np.abs(out["d90_y"]).max()))
