"""ORACLE (test infrastructure) — run the UNMODIFIED reference host loop on CPU.

/root/reference/src/models/optimize_model.py, loss_utils.py, transform_utils.py,
utils/opt_utils.py and strategies/ are imported from where they lie and executed as they
are; only the names they import that do not exist offline are provided:

* ``smplx``  -> oracle.mano_oracle   (un-vendored smplx==0.1.28, docs/ihmr.yml:132)
* ``sdf``    -> oracle.sdf_oracle    (un-vendored SDF_ihmr, docs/install.md:37)
* ``ry_utils`` / ``opendr.*``        -> empty stand-ins (helpers / rendering, unused here)
* ``.cuda()`` on tensors and modules, ``torch.cuda.FloatTensor`` -> CPU identity, because
  the reference hard-codes CUDA placement (base_model.py:19, optimize_model.py:99,117 ...).

This only works where /root/reference exists (the authoring container); it is used by
tests/golden/make_golden.py to produce the committed fixtures and by the not-gpu tests that
pin oracle/host_loop_oracle.py to the real reference.  Nothing under -m gpu uses it.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import torch

REFERENCE_SRC = "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(REFERENCE_SRC)


def _install_shims(dtype=torch.float32):
    from oracle import mano_oracle, sdf_oracle

    smplx = types.ModuleType("smplx")

    def create(*a, **k):
        return mano_oracle.create(*a, dtype=dtype, **k)

    smplx.create = create
    sys.modules["smplx"] = smplx

    sdf = types.ModuleType("sdf")
    sdf.SDFLoss = sdf_oracle.SDFLoss
    sdf.SDFLoss_Single = sdf_oracle.SDFLoss_Single
    sys.modules["sdf"] = sdf

    ry = types.ModuleType("ry_utils")
    ry.build_dir = lambda d: os.makedirs(d, exist_ok=True)
    sys.modules["ry_utils"] = ry

    for name in ("opendr", "opendr.camera", "opendr.renderer", "opendr.lighting"):
        m = types.ModuleType(name)
        for attr in ("ProjectPoints", "ColoredRenderer", "LambertianPointLight"):
            setattr(m, attr, object)
        sys.modules[name] = m

    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor if dtype == torch.float32 else torch.DoubleTensor


def make_opt(model_root: str, batch_size: int, save_mid_freq: int = 10, strategy: str = "opt_default"):
    """The option fields the path reads (SURVEY.md §5 'Config / flags'), with the defaults of
    src/options/base_options.py:11-45 and opt_options.py:3-19."""
    return argparse.Namespace(
        isTrain=False, dist=False, process_rank=-1, batchSize=batch_size, inputSize=224,
        total_params_dim=122, cam_params_dim=3, pose_params_dim=96, shape_params_dim=20,
        trans_params_dim=3, num_joints=42, model_root=model_root, strategy=strategy,
        optimizer="adam", save_mid_freq=save_mid_freq, sdf_robustifier=None,
        checkpoints_dir="./checkpoints", use_hand_rotation=False)


def load_reference_model(opt, epochs=None, dtype=torch.float32):
    """Instantiate the reference's own OptimizeModel on CPU. ``epochs`` (int) overrides the
    per-stage epoch count of the strategy (the reference ships 300; the fixed-iteration
    benchmark strategy of SURVEY.md §8(d) uses 24)."""
    assert reference_available(), "/root/reference is not present on this machine"
    _install_shims(dtype)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import copy
    from models.optimize_model import OptimizeModel   # the reference's file, unmodified
    model = OptimizeModel(opt)
    if epochs is not None:
        model.strategy = copy.deepcopy(model.strategy)
        for stage in model.strategy:
            stage["epoch"] = epochs
    return model
