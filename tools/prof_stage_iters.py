"""One warm + one steady-state iteration of every opt_default stage (the launches `ihmr_opt_profile_iteration` makes),
for `ncu --set full`: the second half of each stage's launches is the steady state (hints warm, stage-specialised
kernels, fingertip-only backward for hands without collision gradient).

    ncu --set full --clock-control none --import-source on -k regex:^k_ -o gpurun_out/r02_full python tools/prof_stage_iters.py
"""
import argparse
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ihmr_b200 import synthetic                      # noqa: E402
from ihmr_b200.optimize_model import OptimizeModel   # noqa: E402
from ihmr_b200.strategies import opt_default, with_epochs   # noqa: E402
from tests import helpers as H                      # noqa: E402
from tools.prof_iters import gpu_targets            # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=65536)
ap.add_argument("--mode", default="typical")
ap.add_argument("--stages", default="0,1,2,3")
args = ap.parse_args()
dev = torch.device("cuda:0")
root = tempfile.mkdtemp(prefix="ihmr_prof_")
synthetic.write_mano_pkls(root, seed=0)
strategy = with_epochs(opt_default, 24)
model = OptimizeModel(H.make_opt(root, args.frames, save_mid_freq=10, strategy=strategy, bs_norm=512), device=dev)
raw = synthetic.make_raw_frames(0, args.frames, seed=0, mode=args.mode)
model.set_input(H.torch_batch(gpu_targets(model, raw, dev)))
model.init_optimize()
torch.cuda.synchronize()
print("PROFILE-BEGIN", flush=True)
for s in [int(x) for x in args.stages.split(",")]:
    ms = model.profile_iteration(strategy[s])
    print("stage", s, {k: round(v, 3) for k, v in ms.items()}, flush=True)
