"""The penetration op alone (BASELINE config 3): timing, work counters and an oracle check.

    python tools/sdf_bench.py --frames 8192 [--mode typical|collision] [--stats] [--check N] [--sweep]

Vertices come from the CUDA MANO layer on the seeded synthetic frames (true parameters); the op is
called through the C ABI (`ihmr_sdf_loss`, forward + gradient in one call).  `--check N` compares the
first N frames with the C oracle (loss, origin-scale values, gradient).  `--sweep` prints the
B = 1 ... 16384 table of config 3.  Prints one JSON object per line.
"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ihmr_b200 import _lib, synthetic                 # noqa: E402
from ihmr_b200.mano_layer import create as create_mano   # noqa: E402

ALG_BYTES_PER_FRAME = 43572          # verts in + grad out + loss + origin-scale


def two_hand_verts(layer, raw, dev, chunk=8192):
    """(B,2,778,3) world-frame vertices of the true parameters (optimize_model.py:171-232)."""
    M = torch.tensor([1.0, -1.0, -1.0], device=dev)
    X = torch.tensor([-1.0, 1.0, 1.0], device=dev)
    out = []
    for s in range(0, raw["true_pose"].shape[0], chunk):
        p, sh, t = (torch.tensor(raw[k][s:s + chunk], device=dev) for k in ("true_pose", "true_shape", "true_trans"))
        b = p.shape[0]
        with torch.no_grad():
            o = layer(global_orient=torch.cat([p[:, 0:3], p[:, 48:51] * M]).contiguous(),
                      hand_pose=torch.cat([p[:, 3:48], (p[:, 51:96].reshape(b, 15, 3) * M).reshape(b, 45)]).contiguous(),
                      betas=torch.cat([sh[:, :10], sh[:, 10:]]).contiguous())
            rv, lv = o.vertices[:b], o.vertices[b:] * X
            rj, lj = o.joints[:b, 0:1], o.joints[b:, 0:1] * X
            lv = lv + (t.view(b, 1, 3) + rj - lj)
            out.append(torch.stack([rv, lv], 1))
    return torch.cat(out).contiguous()


class SdfOp:
    def __init__(self, layer, dev):
        self.lib, self.dev = _lib.load(), dev
        self.handle = layer._device_model(dev).handle if hasattr(layer, "_device_model") else layer
        self.st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run(self, hv, grad=True, stats=None):
        B = hv.shape[0]
        losses = torch.empty(B, device=self.dev)
        origin = torch.empty(B, 1556, device=self.dev)
        g = torch.empty_like(hv) if grad else None
        ws = torch.empty(self.lib.ihmr_sdf_workspace_bytes(B), dtype=torch.uint8, device=self.dev)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        if stats is None:
            _lib.check(self.lib.ihmr_sdf_loss(self.handle, B, p(hv), p(losses), None, p(origin), p(g), 0.0, p(ws), ws.numel(), self.st), "sdf")
        else:
            _lib.check(self.lib.ihmr_sdf_stats(self.handle, B, p(hv), p(losses), p(stats), p(ws), ws.numel(), self.st), "stats")
        return losses, origin, g

    def time(self, hv, reps=20):
        for _ in range(3):
            self.run(hv)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self.run(hv)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=8192)
    ap.add_argument("--mode", default="typical")
    ap.add_argument("--stats", action="store_true")
    ap.add_argument("--stats-json", action="store_true", help="work counters as one JSON line (tests)")
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--start", type=int, default=0)
    ap.add_argument("--dump", default=None, help="with --stats: save the raw (B,32) counters + losses as .npz")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    root = tempfile.mkdtemp(prefix="ihmr_sdfb_")
    synthetic.write_mano_pkls(root, seed=0)
    right = create_mano(os.path.join(root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True).to(dev)
    raw = synthetic.make_raw_frames(args.start, args.frames, seed=0, mode=args.mode)
    hv = two_hand_verts(right, raw, dev)
    from ihmr_b200.optimize_model import OptimizeModel
    from tests import helpers as H
    model = OptimizeModel(H.make_opt(root, 1), device=dev)
    op = SdfOp(model._model.handle, dev)

    ms = op.time(hv)
    B = args.frames
    print(json.dumps({"what": "sdf fwd+bwd", "mode": args.mode, "frames": B, "ms": ms, "us_per_frame": ms * 1e3 / B,
                      "GBps_vs_alg": ALG_BYTES_PER_FRAME * B / (ms * 1e-3) / 1e9}), flush=True)

    if args.sweep:
        for b in (1, 4, 16, 64, 256, 1024, 4096, 16384):
            if b > B:
                break
            t = op.time(hv[:b].contiguous())
            print(json.dumps({"what": "sweep", "mode": args.mode, "frames": b, "ms": t, "us_per_frame": t * 1e3 / b}), flush=True)

    if args.stats_json:
        stats = torch.zeros(B, 32, dtype=torch.int32, device=dev)
        op.run(hv, stats=stats)
        torch.cuda.synchronize()
        st = stats.cpu().numpy().astype(np.int64)
        searching = (st[:, 0] > 0).astype(np.int64) + (st[:, 2] > 0)
        print(json.dumps({"what": "stats", "frames": B, "voxels": int(st[:, 0].sum() + st[:, 2].sum()), "passes": int(st[:, 10].sum()),
                          "max_passes_per_direction": int(np.max(st[:, 10] - np.maximum(searching - 1, 0))) if B else 0,
                          "candidates": int(st[:, 7].sum()),
                          "ray_flushes": int(st[:, 11].sum()), "candidate_flushes": int(st[:, 12].sum()),
                          "prep_done_directions": int(st[:, 16].sum() + st[:, 17].sum())}), flush=True)

    if args.stats:
        stats = torch.zeros(B, 32, dtype=torch.int32, device=dev)
        losses, _, _ = op.run(hv, stats=stats)
        torch.cuda.synchronize()
        st = stats.cpu().numpy().astype(np.float64)
        l = losses.cpu().numpy()
        if args.dump:
            np.savez_compressed(args.dump, stats=stats.cpu().numpy(), losses=l)
        rows = [(0, "voxels gridR"), (2, "voxels gridL"), (4, "activeQ gridR"), (5, "activeQ gridL"),
                (6, "pairs"), (7, "candidates"), (8, "marked"), (9, "ray items"), (10, "passes"), (11, "ray flushes"),
                (12, "cand flushes"), (16, "prep-done R"), (17, "prep-done L")]
        for i, n in rows:
            c = st[:, i]
            print(f"{n:>16}: mean {c.mean():9.1f}  p50 {np.percentile(c, 50):7.0f} p90 {np.percentile(c, 90):8.0f} "
                  f"p99 {np.percentile(c, 99):8.0f} max {c.max():8.0f}  nonzero {np.mean(c > 0) * 100:5.1f}%")
        ph = ["mark", "face boxes", "parity", "scan", "worklist", "seeds+cands", "exact tests", "finish", "sample+out"]
        tot = st[:, 19:28].sum()
        for i, n in enumerate(ph):
            c = st[:, 19 + i]
            print(f"{n:>14}: {c.sum() / tot * 100:5.1f}% of cycles  mean {c.mean():9.0f}  p99 {np.percentile(c, 99):9.0f}  max {c.max():9.0f}")
        print("mean cycles per frame (both directions, thread 0)", st[:, 19:28].sum(1).mean())
        print("loss>0 frames %.1f%%, mean loss %.4f" % (np.mean(l > 0) * 100, l.mean()))

    if args.check:
        from oracle import mano_oracle, sdf_oracle
        n = min(args.check, B)
        sub = hv[:n].contiguous()
        losses, origin, g = op.run(sub)
        torch.cuda.synchronize()
        ro = mano_oracle.create(os.path.join(root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True)
        lo = mano_oracle.create(os.path.join(root, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False)
        hc = sub.cpu().clone().requires_grad_(True)
        l_ref, _, o_ref = sdf_oracle.SDFLoss(ro.faces, lo.faces)(hc, True, True)
        l_ref.sum().backward()
        dl = (losses.cpu() - l_ref.detach()).abs()
        rel = dl / l_ref.detach().abs().clamp_min(1e-6)
        do = (origin.cpu() - o_ref).abs().max(1).values
        gs = hc.grad.abs().amax((1, 2, 3)).clamp_min(1e-9)
        dg = (g.cpu() - hc.grad).abs().amax((1, 2, 3)) / gs
        print(json.dumps({"what": "check", "mode": args.mode, "frames": n, "colliding": int((l_ref > 0).sum()),
                          "max_abs_loss_err": float(dl.max()), "max_rel_loss_err": float(rel.max()),
                          "max_origin_err_m": float(do.max()), "max_rel_grad_err": float(dg.max()),
                          "worst_frames": torch.topk(rel, min(5, n)).indices.tolist()}), flush=True)


if __name__ == "__main__":
    main()
