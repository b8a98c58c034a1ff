"""Distribution of the penetration kernel's per-frame work counters on synthetic frames."""
import argparse, os, sys, tempfile, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ihmr_b200 import synthetic, _lib
from ihmr_b200.optimize_model import OptimizeModel
from ihmr_b200.strategies import opt_default, with_epochs
from tests import helpers as H
from tools.prof_iters import gpu_targets

ap = argparse.ArgumentParser(); ap.add_argument("--frames", type=int, default=8192); ap.add_argument("--mode", default="typical")
args = ap.parse_args()
dev = torch.device("cuda:0"); root = tempfile.mkdtemp(); synthetic.write_mano_pkls(root, seed=0)
model = OptimizeModel(H.make_opt(root, args.frames, strategy=with_epochs(opt_default, 24), bs_norm=512), device=dev)
raw = synthetic.make_raw_frames(0, args.frames, seed=0, mode=args.mode)
model.set_input(H.torch_batch(gpu_targets(model, raw, dev))); model.init_optimize(); model.forward()
hv = torch.stack([model.pred_right_hand_verts, model.pred_left_hand_verts], 1).contiguous()
losses = torch.empty(args.frames, device=dev); stats = torch.zeros(args.frames, 32, dtype=torch.int32, device=dev)
lib = _lib.load()
_lib.check(lib.ihmr_sdf_stats(model._model.handle, args.frames, C.c_void_p(hv.data_ptr()), C.c_void_p(losses.data_ptr()),
                              C.c_void_p(stats.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "stats")
torch.cuda.synchronize()
st = stats.cpu().numpy().astype(np.float64); l = losses.cpu().numpy()
names = ["evalR", "farR", "evalL", "farL", "activeQ(gridR)", "activeQ(gridL)", "binsR", "binsL"]
for i, n in enumerate(names):
    c = st[:, i]
    print(f"{n:>16}: mean {c.mean():8.1f}  p50 {np.percentile(c,50):7.0f} p90 {np.percentile(c,90):7.0f} p99 {np.percentile(c,99):7.0f} max {c.max():7.0f}  nonzero {np.mean(c>0)*100:5.1f}%")
names[6:8] = ["pairs", "candidates"]
for i, n in [(6, "pairs"), (7, "candidates"), (20, "marked"), (21, "ray items"), (22, "pairs b0"), (23, "cands b0"), (24, "pairs b1"), (25, "cands b1"),
             (26, "pairs b2"), (27, "cands b2")]:
    c = st[:, i]
    print(f"{n:>16}: mean {c.mean():9.1f}  p50 {np.percentile(c,50):7.0f} p90 {np.percentile(c,90):8.0f} p99 {np.percentile(c,99):8.0f} max {c.max():8.0f}")
ph = ["-", "bbox", "mark", "normalise", "parity", "scan", "worklist", "candidates+tests", "classify+far", "sample", "outputs", "-"]
tot = st[:, 8:20].sum()
for i in range(1, 12):
    c = st[:, 8 + i]
    print(f"{ph[i]:>14}: {c.sum()/tot*100:5.1f}% of cycles  mean {c.mean():9.0f}  p99 {np.percentile(c,99):9.0f}  max {c.max():9.0f}")
print("mean cycles per frame", st[:, 8:20].sum(1).mean())
print("loss>0 frames %.1f%%, mean loss %.3f" % (np.mean(l > 0) * 100, l.mean()))
