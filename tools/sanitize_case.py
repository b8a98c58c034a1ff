"""A small pass over every kernel of the library for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).

    compute-sanitizer --tool racecheck python tools/sanitize_case.py

MANO layer forward + backward, the penetration op on typical and near-coincident frames, a short refinement loop
that goes through all four stage plans (generic, rigid, affine) with snapshots, the final forward, the evaluator
metrics and the selection routine.  CUDA graphs are off (the sanitizer sees every launch directly).
"""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ihmr_b200 import synthetic                      # noqa: E402
from ihmr_b200.optimize_model import OptimizeModel   # noqa: E402
from ihmr_b200.strategies import opt_default, with_epochs   # noqa: E402
from tests import helpers as H                      # noqa: E402
from tools.prof_iters import gpu_targets            # noqa: E402
from tools.sdf_bench import SdfOp, two_hand_verts   # noqa: E402


def main():
    dev = torch.device("cuda:0")
    root = tempfile.mkdtemp(prefix="ihmr_san_")
    synthetic.write_mano_pkls(root, seed=0)
    B = 6
    opt = H.make_opt(root, B, save_mid_freq=2, strategy=with_epochs(opt_default, 3), bs_norm=B)
    opt.use_cuda_graphs = False
    model = OptimizeModel(opt, device=dev)
    layer = model.mano_models["right"].to(dev)
    # MANO op
    g = torch.Generator().manual_seed(0)
    ins = [t.to(dev).requires_grad_(True) for t in ((torch.rand(5, 3, generator=g) - 0.5) * 3, torch.randn(5, 45, generator=g) * 0.4,
                                                    torch.randn(5, 10, generator=g))]
    out = layer(global_orient=ins[0], hand_pose=ins[1], betas=ins[2])
    (out.vertices.sum() + out.joints.sum()).backward()
    # penetration op
    op = SdfOp(model._model.handle, dev)
    for mode in ("typical", "collision"):
        hv = two_hand_verts(layer, synthetic.make_raw_frames(40, 4, seed=0, mode=mode), dev)
        losses, _, _ = op.run(hv)
        stats = torch.zeros(4, 32, dtype=torch.int32, device=dev)
        op.run(hv, stats=stats)
        print(mode, "losses", losses.cpu().numpy())
    # loop: three typical + three near-coincident frames
    raw = {k: np.concatenate([a, b]) for (k, a), (_, b) in zip(synthetic.make_raw_frames(0, 3, seed=0).items(),
                                                              synthetic.make_raw_frames(512, 3, seed=0, mode="collision").items())}
    model.set_input(H.torch_batch(gpu_targets(model, raw, dev)))
    model.init_optimize()
    model.optimize(0, 1)
    res = model.get_pred_result()
    print("loop collision_loss", res["collision_loss"])
    from ihmr_b200.evaluator import DeviceEvaluator
    ev = DeviceEvaluator()
    ev.update(np.arange(B), model)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
