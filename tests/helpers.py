"""Shared test helpers: synthetic frames through the oracle, golden fixture loading."""
import os

import numpy as np
import torch

from ihmr_b200 import synthetic
from oracle import host_loop_oracle as HL
from oracle import mano_oracle, sdf_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_batch(layer, start, count, mode="typical", seed=0):
    raw = synthetic.make_raw_frames(start, count, seed=seed, mode=mode)

    def fwd(p, s, t):
        with torch.no_grad():
            return mano_oracle.two_hand_forward(layer, torch.tensor(p), torch.tensor(s), torch.tensor(t))[2].numpy()
    return synthetic.finish_frames(raw, fwd)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    data = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    out = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    return data, out, int(z["epochs"]), int(z["save_mid_freq"])


def oracle_loop(layers, B, epochs, freq, bs_norm=None, dtype=torch.float32, ray_axis=0):
    right, left = layers
    if dtype != torch.float32:
        right, left = right.double(), left.double()
    sdf = sdf_oracle.SDFLoss(right.faces, left.faces, ray_axis=ray_axis)
    return HL.HostLoopOracle(right, right.faces, left.faces, sdf, B, strategy=HL.opt_default_strategy(epochs),
                             save_mid_freq=freq, bs_norm=bs_norm, dtype=dtype)


def torch_batch(data):
    return {k: torch.from_numpy(np.asarray(v)) for k, v in data.items()}


def make_opt(model_root, batch_size, save_mid_freq=10, strategy="opt_default", bs_norm=None):
    import argparse
    return argparse.Namespace(
        isTrain=False, dist=False, process_rank=-1, batchSize=batch_size, inputSize=224, total_params_dim=122,
        cam_params_dim=3, pose_params_dim=96, shape_params_dim=20, trans_params_dim=3, num_joints=42,
        model_root=model_root, strategy=strategy, optimizer="adam", save_mid_freq=save_mid_freq,
        sdf_robustifier=None, checkpoints_dir="./checkpoints", bs_norm=bs_norm, quiet=True)
