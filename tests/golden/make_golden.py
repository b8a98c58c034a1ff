"""Generates the committed golden fixtures (run once, in the authoring container).

    python tests/golden/make_golden.py

Everything below is produced by the UNMODIFIED reference host loop
(/root/reference/src/models/optimize_model.py and friends, imported through
oracle/ref_shims.py) driving the oracle leaves on CPU, on the seeded synthetic model and
frames of ihmr_b200/synthetic.py.  The reference itself has no tests or golden vectors
(SURVEY.md §4), so these fixtures are the pin for oracle/host_loop_oracle.py and, through it,
for the CUDA path.  /root/reference does not exist on the GPU box: the fixtures travel, the
reference does not.

Fixtures (float32 .npz, a few hundred KB in total):
  loop_cfg1.npz    config 1 of BASELINE.json: one frame, opt_default with epoch=24 per stage
                   (4 x 25 = 100 iterations), save_mid_freq=10        -> inputs + 13 result keys
  loop_b2_short.npz  two frames, epoch=3, save_mid_freq=2              -> inputs + results
  loop_collision_short.npz  one near-coincident frame (cfg5 style), epoch=3, save_mid_freq=1
  loop_mixed6.npz  six frames (three typical, three near-coincident), epoch=8, save_mid_freq=4
  loop_long_typical.npz / loop_long_collision.npz  one frame each at the SHIPPED strategy length
                   (opt_default.py:15,34,53,72: epoch=300 per stage = 1,204 Adam steps,
                   save_mid_freq=10 = 31 snapshots per stage; bash/optimize.sh:33)   [--only-long]
  leaves.npz       MANO leaf fwd (8 hands) and SDFLoss fwd/grad (2 frames) from the oracle leaves
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ihmr_b200 import synthetic as S          # noqa: E402
from oracle import mano_oracle as MO          # noqa: E402
from oracle import ref_shims as RS            # noqa: E402
from oracle import sdf_oracle as SO           # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
MODEL_ROOT = "/tmp/ihmr_golden_model"


def frames(start, count, mode, layer):
    raw = S.make_raw_frames(start, count, seed=0, mode=mode)

    def fwd(p, s, t):
        with torch.no_grad():
            return MO.two_hand_forward(layer, torch.tensor(p), torch.tensor(s), torch.tensor(t))[2].numpy()
    return S.finish_frames(raw, fwd)


def run_loop(data, epochs, freq):
    B = data["init_cam"].shape[0]
    opt = RS.make_opt(MODEL_ROOT, B, save_mid_freq=freq)
    model = RS.load_reference_model(opt, epochs=epochs)
    t = time.time()
    model.set_input({k: torch.from_numpy(v) for k, v in data.items()})
    model.init_optimize()
    model.optimize(0, 1)
    res = model.get_pred_result()
    print(f"  reference loop B={B} epochs={epochs}: {time.time() - t:.1f} s")
    out = {"in_" + k: v for k, v in data.items()}
    out.update({"out_" + k: v for k, v in res.items()})
    out["epochs"], out["save_mid_freq"] = np.int64(epochs), np.int64(freq)
    return out


def mixed6(layer):
    a, b = frames(7, 3, "typical", layer), frames(512, 3, "collision", layer)
    return {k: np.concatenate([a[k], b[k]]) for k in a}


def main():
    torch.manual_seed(0)
    S.write_mano_pkls(MODEL_ROOT, seed=0)
    layer = MO.create(os.path.join(MODEL_ROOT, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True)
    left = MO.create(os.path.join(MODEL_ROOT, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False)
    if "--only-long" in sys.argv:         # round 2: the shipped strategy length (minutes of CPU per frame)
        which = [a.split("=")[1] for a in sys.argv if a.startswith("--which=")] or ["typical", "collision"]
        for mode in which:
            start = 3 if mode == "typical" else 600
            np.savez_compressed(os.path.join(OUT, f"loop_long_{mode}.npz"), **run_loop(frames(start, 1, mode, layer), 300, 10))
        return
    if "--only-mixed6" in sys.argv:       # added later in the round; the other fixtures are left untouched
        np.savez_compressed(os.path.join(OUT, "loop_mixed6.npz"), **run_loop(mixed6(layer), 8, 4))
        return

    # leaves
    g = torch.Generator().manual_seed(1)
    N = 8
    orient = (torch.rand(N, 3, generator=g) - 0.5) * 3.0
    pose = torch.randn(N, 45, generator=g) * 0.4
    betas = torch.randn(N, 10, generator=g)
    orient[0] = 0
    pose[0] = 0      # exercises the ||r + 1e-8|| branch
    out = layer(global_orient=orient, hand_pose=pose, betas=betas)
    data2 = frames(0, 2, "typical", layer)
    raw2 = S.make_raw_frames(0, 2, seed=0)
    with torch.no_grad():
        rv, lv, _ = MO.two_hand_forward(layer, torch.tensor(raw2["true_pose"]),
                                        torch.tensor(raw2["true_shape"]), torch.tensor(raw2["true_trans"]))
    hv = torch.stack([rv, lv], 1).clone().requires_grad_(True)
    sdf = SO.SDFLoss(layer.faces, left.faces)
    losses, per_vert, origin = sdf(hv, True, True)
    losses.sum().backward()
    np.savez_compressed(os.path.join(OUT, "leaves.npz"),
                        mano_orient=orient.numpy(), mano_pose=pose.numpy(), mano_betas=betas.numpy(),
                        mano_vertices=out.vertices.numpy(), mano_joints=out.joints.numpy(),
                        sdf_hand_verts=hv.detach().numpy(), sdf_losses=losses.detach().numpy(),
                        sdf_origin_scale=origin.numpy(), sdf_grad=hv.grad.numpy())
    print("leaves.npz written; sdf losses", losses.tolist())

    np.savez_compressed(os.path.join(OUT, "loop_b2_short.npz"), **run_loop(data2, 3, 2))
    np.savez_compressed(os.path.join(OUT, "loop_collision_short.npz"),
                        **run_loop(frames(0, 1, "collision", layer), 3, 1))
    np.savez_compressed(os.path.join(OUT, "loop_cfg1.npz"), **run_loop(frames(0, 1, "typical", layer), 24, 10))
    np.savez_compressed(os.path.join(OUT, "loop_mixed6.npz"), **run_loop(mixed6(layer), 8, 4))


if __name__ == "__main__":
    main()
