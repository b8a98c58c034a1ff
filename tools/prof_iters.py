"""Runs a few fused iterations (and optionally whole steps) for profiling under ncu.

    python tools/prof_iters.py --frames 8192 --iters 4 --stage 2 [--mode collision] [--full-step]
"""
import argparse
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


def gpu_targets(model, raw, dev):
    """two-hand joints of the true parameters through the CUDA MANO layer"""
    layer = model.mano_models["right"].to(dev)
    M = torch.tensor([1.0, -1.0, -1.0], device=dev)
    X = torch.tensor([-1.0, 1.0, 1.0], device=dev)

    def fwd(pose, shape, trans):
        out = np.empty((pose.shape[0], 42, 3), np.float32)
        for s in range(0, pose.shape[0], 8192):
            p, sh, t = (torch.tensor(a[s:s + 8192], device=dev) for a in (pose, shape, trans))
            b = p.shape[0]
            with torch.no_grad():
                o = layer(global_orient=torch.cat([p[:, 0:3], p[:, 48:51] * M]).contiguous(),
                          hand_pose=torch.cat([p[:, 3:48], (p[:, 51:96].reshape(b, 15, 3) * M).reshape(b, 45)]).contiguous(),
                          betas=torch.cat([sh[:, :10], sh[:, 10:]]).contiguous())
                j = torch.cat([o.joints, o.vertices[:, [744, 320, 443, 554, 671]]], 1)
                rj, lj = j[:b], j[b:] * X
                lj = lj + (t.view(b, 1, 3) + rj[:, 0:1] - lj[:, 0:1])
                out[s:s + b] = torch.cat([rj, lj], 1).cpu().numpy()
        return out
    return synthetic.finish_frames(raw, fwd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--stage", type=int, default=2)
    ap.add_argument("--mode", default="typical")
    ap.add_argument("--full-step", action="store_true")
    ap.add_argument("--epochs", type=int, default=24)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    root = tempfile.mkdtemp(prefix="ihmr_prof_")
    synthetic.write_mano_pkls(root, seed=0)
    strategy = with_epochs(opt_default, args.epochs)
    model = OptimizeModel(H.make_opt(root, args.frames, save_mid_freq=10, strategy=strategy, bs_norm=512), device=dev)
    raw = synthetic.make_raw_frames(0, args.frames, seed=0, mode=args.mode)
    data = gpu_targets(model, raw, dev)
    model.set_input(H.torch_batch(data))
    model.init_optimize()
    if args.full_step:
        model.optimize(0, 1)
    else:
        for _ in range(args.iters):
            losses, grad = model.value_and_grad(strategy[args.stage])
        print("losses", losses.cpu().numpy())
    torch.cuda.synchronize()
    if os.environ.get("IHMR_STATS"):
        ms = model.profile_iteration(strategy[args.stage])
        print(ms)


if __name__ == "__main__":
    main()
