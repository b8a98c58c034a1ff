"""Benchmark of the IHMR-OPT refinement hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--config 2|3|4|5] [--scaling weak|strong] [--frames F]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default (config 4): a step = one full refinement of one batch: the four opt_default stages with
epoch=24 (4 x 25 = 100 fwd+bwd+Adam iterations, snapshots every 10, online selection) plus the final
forward, on `--frames` synthetic two-hand frames per GPU (default 65536; frames are independent so
ranks hold disjoint blocks), followed by the one all-gather of refined parameters + loss statistics
when N > 1.  `--scaling weak` (default) keeps 65536 frames per GPU; `--scaling strong` splits 65536
frames over the N GPUs.  `--config 5` is the same loop on near-coincident hands (worst-case collision
density).  `--config 2` times the MANO layer op (forward + backward, 2 x 4096 hands) and `--config 3`
the penetration op (forward + backward) over B = 1 ... 16384 frames, both through the C ABI.

Prints ONE JSON line (rank 0).  `value` = frames refined per second with the batch resident in HBM;
`e2e` = the same through the public API (`OptimizeModel.run_pipelined`) from pinned host buffers: every
step copies its 15 input tensors to the device and its 11 result arrays back, on copy streams that
overlap the neighbouring steps' compute; `roofline` = the dominant kernel's algorithmic bytes /
measured device time against MEASURED_PEAKS.json; `cpu_baseline` = the oracle port of the reference
loop timed on this box's host cores.  `--impl reference` times that CPU loop alone with the same JSON
shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--frames", type=int, default=65536, help="frames per GPU (weak) / in total (strong)")
    ap.add_argument("--mode", default=None, choices=["typical", "collision"], help="frame generator (default: by config)")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per timed step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="length of the timed end-to-end run (default: max(12, --steps))")
    args = ap.parse_args()
    if args.mode is None:
        args.mode = "collision" if args.config == 5 else "typical"
    return args


ARGS = parse_args()
if ARGS.impl == "reference":
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (set before torch/OpenMP load)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np      # noqa: E402
import torch            # noqa: E402

METRIC = "two-hand frames refined/sec (fixed iters)"
EPOCHS, FREQ, BS_NORM = 24, 10, 512           # 4 x 25 iterations, bs_norm 512
ITERS = 4 * (EPOCHS + 1)

# Algorithmic bytes per launch unit: compulsory op-boundary traffic, fp32.
ALG_BYTES = {                       # per hand unless noted
    "pose_prep": 58 * 4 + 160 * 4 + 192 * 4 + 48 * 4,
    "blend_fwd": 160 * 4 + 2334 * 4,
    "skin_fwd": 2334 * 4 + 192 * 4 + 2334 * 4,
    "sdf": 43572 / 2,               # 43,572 B per FRAME
    "frame_loss": (48 * 4 + 15 * 4 + 63 * 4 + 122 * 4) / 1,
    "skin_bwd": 2334 * 4 * 3 + 192 * 4 * 2,
    "blend_bwd": 2334 * 4 + 160 * 4,
    "pose_bwd": 192 * 4 + 48 * 4 + 160 * 4 + 58 * 4,
    "step": 61 * 4 * 5,
}
STEP_BYTES_PER_FRAME_ITER = 76388   # SURVEY §8(d) op-boundary figure for the fused step
MANO_BYTES_PER_HAND = 19520         # MANO op fwd+bwd
MANO_TENSOR_FLOP_PER_HAND = 1.26e6  # posedirs contraction fwd + bwd-data
MANO_FP32_FLOP_PER_HAND = 1.2e6     # LBS, shape blend / regress, chain
SDF_BYTES_PER_FRAME = 43572
SDF_FLOP_PER_TEST = 150.0           # one exact point-triangle test (instruction count of pt_tri_dist2 + loads)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ CPU arms
def cpu_loop_sample(model_root, frames, epochs, mode="typical"):
    """Oracle port of the reference host loop + oracle leaves on the host cores: `frames` frames,
    `epochs` per stage (+ the final forward). Returns (seconds, iterations, torch threads)."""
    from oracle import mano_oracle
    from tests import helpers as H
    right = mano_oracle.create(os.path.join(model_root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True)
    left = mano_oracle.create(os.path.join(model_root, "MANO_LEFT.pkl"), "mano", use_pca=False, is_rhand=False)
    batch = H.torch_batch(H.make_batch(right, 0, frames, mode=mode))
    loop = H.oracle_loop((right, left), frames, epochs, FREQ, bs_norm=BS_NORM)
    t0 = time.perf_counter()
    loop.set_input(batch)
    loop.init_optimize()
    loop.optimize()
    loop.get_pred_result()
    return time.perf_counter() - t0, 4 * (epochs + 1), torch.get_num_threads()


def cpu_baseline_block(model_root, frames, mode):
    """One small warm-up (thread pools, library loads), then `frames` frames x the full 100 iterations."""
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_loop_sample(model_root, 1, 1, mode)
    sec, iters, thr = cpu_loop_sample(model_root, frames, EPOCHS, mode)
    return {"value": frames / sec, "unit": "frames/s", "cores": thr, "kind": "port",
            "sample": f"{frames} frame(s) x {iters} iterations (+final forward) = {sec:.1f} s on {thr} torch/OpenMP "
                      f"threads of {os.cpu_count()} host CPUs, after one 8-iteration warm-up; nothing extrapolated",
            "host_cpus": os.cpu_count(), "omp_num_threads": os.environ.get("OMP_NUM_THREADS")}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from ihmr_b200 import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    root = tempfile.mkdtemp(prefix="ihmr_ref_")
    synthetic.write_mano_pkls(root, seed=0)
    F = max(1, args.ref_frames)
    for _ in range(args.warmup):                 # warm-up steps: 1 frame x 8 iterations each
        cpu_loop_sample(root, 1, 1, args.mode)
    times = []
    for _ in range(args.steps):                  # timed steps: F frames x the full 100 iterations (+ final forward)
        sec, iters, thr = cpu_loop_sample(root, F, EPOCHS, args.mode)
        times.append(sec)
    sec = float(np.mean(times))
    value = F / sec
    sample = (f"{F} frame(s) x {iters} iterations (+final forward) per timed step, {thr} torch/OpenMP threads of "
              f"{os.cpu_count()} host CPUs; warm-up steps are 1 frame x 8 iterations; nothing extrapolated")
    cfg = workload_config(args, 1, F)
    cfg["workload"] = (f"IHMR-OPT full loop, CPU arm: {F} synthetic two-hand frame(s) x {ITERS} iterations per step "
                       f"(opt_default, epoch={EPOCHS}/stage, save_mid_freq={FREQ}, bs_norm={BS_NORM}); a bounded sample of "
                       f"BASELINE config {args.config}'s workload, rank 0 only")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": thr, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count(), "omp_num_threads": os.environ.get("OMP_NUM_THREADS"),
                         "step_seconds": times},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world, frames_per_gpu):
    name = {4: "IHMR-OPT full loop (BASELINE config 4)", 5: "IHMR-OPT full loop on near-coincident hands (BASELINE config 5)"}
    return {"workload": f"{name.get(args.config, 'IHMR-OPT')}: {frames_per_gpu} synthetic two-hand frames per GPU x "
                        f"{ITERS} iterations (opt_default, epoch={EPOCHS}/stage, save_mid_freq={FREQ}, bs_norm={BS_NORM})",
            "frames_per_gpu": frames_per_gpu, "global_frames": frames_per_gpu * world, "iterations": ITERS,
            "frame_mode": args.mode, "parallelism": f"frame-sharded x{world}, one all-gather at the end",
            "l2": "per-step working set (~82 KB/frame of intermediates) is far larger than the 126 MB L2; no flush needed"
                  if frames_per_gpu >= 4096 else "small batch: intermediates fit the 126 MB L2 (stated, not flushed)"}


# --------------------------------------------------------------------- op-level configs
def gpu_setup(args, frames, dev, rank=0):
    from ihmr_b200 import synthetic
    from ihmr_b200.optimize_model import OptimizeModel
    from ihmr_b200.strategies import opt_default, with_epochs
    from tests import helpers as H
    root = tempfile.mkdtemp(prefix=f"ihmr_bench_{rank}_")
    synthetic.write_mano_pkls(root, seed=0)
    strategy = with_epochs(opt_default, EPOCHS)
    opt = H.make_opt(root, frames, save_mid_freq=FREQ, strategy=strategy, bs_norm=BS_NORM)
    return root, strategy, OptimizeModel(opt, device=dev)


def event_ms(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def fp32_peak(model, dev):
    import ctypes as C
    from ihmr_b200 import _lib
    out = C.c_float()
    scratch = torch.zeros(4, device=dev)
    _lib.check(_lib.load().ihmr_measure_fp32_peak(model._model.handle, C.byref(out), C.c_void_p(scratch.data_ptr()),
                                                  C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "fp32 peak")
    return float(out.value)


def run_config2(args):
    """MANO op forward + backward on 2 x 4096 hands through ihmr_mano_forward / ihmr_mano_backward."""
    import ctypes as C
    from ihmr_b200 import _lib
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    n = 8192
    _, _, model = gpu_setup(args, 1, dev)
    lib, h = _lib.load(), model._model.handle
    g = torch.Generator().manual_seed(11)
    orient = ((torch.rand(n, 3, generator=g) - 0.5) * 3.0).to(dev)
    pose = (torch.randn(n, 45, generator=g) * 0.4).to(dev)
    betas = torch.randn(n, 10, generator=g).to(dev)
    gv, gj = torch.randn(n, 778, 3, generator=g).to(dev), torch.randn(n, 16, 3, generator=g).to(dev)
    verts, joints = torch.empty(n, 778, 3, device=dev), torch.empty(n, 16, 3, device=dev)
    go, gp, gb = torch.empty_like(orient), torch.empty_like(pose), torch.empty_like(betas)
    ws = torch.empty(lib.ihmr_mano_workspace_bytes(n), dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    # L2: forward + backward stream ~0.4 GB of intermediates (off, gposed, verts, grads) > 126 MB L2; no flush needed
    fwd = lambda: _lib.check(lib.ihmr_mano_forward(h, n, p(orient), p(pose), p(betas), p(verts), p(joints), p(ws), ws.numel(), st), "fwd")
    bwd = lambda: _lib.check(lib.ihmr_mano_backward(h, n, p(orient), p(pose), p(betas), p(gv), p(gj), p(go), p(gp), p(gb), p(ws), ws.numel(), st), "bwd")
    sampler = ClockSampler(0)
    l0 = lib.ihmr_launch_count()
    t_f, t_b = event_ms(fwd, 30), event_ms(bwd, 30)
    launches = (lib.ihmr_launch_count() - l0) // 33      # event_ms runs each of fwd / bwd 3 + 30 times
    clocks = sampler.stop()
    # the backward entry point recomputes the forward intermediates (pose_prep + blend) before its own kernels
    t = t_f + t_b
    peak_hbm, peak_bf16, kind = load_peaks()
    ffma = fp32_peak(model, dev)
    gbs = MANO_BYTES_PER_HAND * n / (t * 1e-3) / 1e9
    line = {
        "metric": "MANO layer fwd+bwd us/hand (BASELINE config 2)", "value": t * 1e3 / n, "unit": "us/hand", "n_gpus": 1,
        "steps": 30, "warmup": 3, "ms_per_step": t, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batched MANO forward+backward, 2 x 4096 hands fp32 through ihmr_mano_forward + ihmr_mano_backward "
                               "(random upstream grad_vertices / grad_joints)", "hands": n,
                   "l2": "~0.4 GB of intermediates per call, larger than the 126 MB L2; no flush"},
        "fwd_ms": t_f, "bwd_ms": t_b, "gpu_launches": int(launches),
        "roofline": {"bound": "fp32", "achieved": MANO_FP32_FLOP_PER_HAND * n / (t * 1e-3) / 1e12, "peak": ffma,
                     "unit": "TFLOP/s", "frac": MANO_FP32_FLOP_PER_HAND * n / (t * 1e-3) / 1e12 / ffma, "traffic": None,
                     "peak_source": "FFMA micro-kernel measured in this run (ihmr_measure_fp32_peak)",
                     "note": "arithmetic intensity 126 FLOP/B: the op is compute-bound (SURVEY §8(d)); HBM and tensor figures beside it"},
        "hbm": {"achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm, "alg_bytes_per_hand": MANO_BYTES_PER_HAND,
                "peak_source": f"{kind} copy bandwidth"},
        "tensor": {"achieved": MANO_TENSOR_FLOP_PER_HAND * n / (t * 1e-3) / 1e12, "unit": "TFLOP/s (algorithmic; 3xTF32 issues 3 MMAs per product)",
                   "bf16_peak_for_scale": peak_bf16},
        "clocks": clocks, "e2e": None, "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)


def run_config3(args):
    """Penetration op forward + backward, B = 1 ... 16384 frames, typical and near-coincident inputs."""
    from tools import sdf_bench as SB
    from ihmr_b200.mano_layer import create as create_mano
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    root, _, model = gpu_setup(args, 1, dev)
    from ihmr_b200 import synthetic
    right = create_mano(os.path.join(root, "MANO_RIGHT.pkl"), "mano", use_pca=False, is_rhand=True).to(dev)
    op = SB.SdfOp(model._model.handle, dev)
    peak_hbm, _, kind = load_peaks()
    ffma = fp32_peak(model, dev)
    sampler = ClockSampler(0)
    sweep = []
    for mode in ("typical", "collision"):
        hv = SB.two_hand_verts(right, synthetic.make_raw_frames(0, 16384, seed=0, mode=mode), dev)
        for b in (1, 4, 16, 64, 256, 1024, 4096, 16384):
            sub = hv[:b].contiguous()
            ms = op.time(sub, reps=20)
            stats = torch.zeros(b, 32, dtype=torch.int32, device=dev)
            op.run(sub, stats=stats)
            torch.cuda.synchronize()
            st = stats.sum(0).cpu().numpy().astype(np.float64)
            tests = float(st[7] + st[0] + st[2])       # queued candidates + one seed test per evaluated voxel
            sweep.append({"mode": mode, "frames": b, "ms": ms, "us_per_frame": ms * 1e3 / b,
                          "GBps_vs_alg": SDF_BYTES_PER_FRAME * b / (ms * 1e-3) / 1e9,
                          "exact_tests_per_frame": tests / b, "tests_per_s": tests / (ms * 1e-3),
                          "voxels_per_frame": float(st[0] + st[2]) / b,
                          "fp32_frac_exact_tests": tests * SDF_FLOP_PER_TEST / (ms * 1e-3) / 1e12 / ffma})
    clocks = sampler.stop()
    top = [s for s in sweep if s["mode"] == "typical" and s["frames"] == 16384][0]
    line = {
        "metric": "penetration loss fwd+bwd us/frame (BASELINE config 3)", "value": top["us_per_frame"], "unit": "us/frame",
        "n_gpus": 1, "steps": 20, "warmup": 3, "ms_per_step": top["ms"], "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "left/right penetration loss forward+backward through ihmr_sdf_loss (stateless: no seeds carried "
                               "between calls), B = 1 ... 16384 frames, typical and near-coincident (config-5 style) inputs; value = "
                               "typical frames at B = 16384",
                   "l2": "B >= 4096: inputs + outputs exceed the 126 MB L2; smaller B are L2-resident (stated, not flushed)"},
        "sweep": sweep,
        "roofline": {"bound": "issue", "achieved": top["GBps_vs_alg"], "peak": peak_hbm, "unit": "GB/s",
                     "frac": top["GBps_vs_alg"] / peak_hbm, "traffic": None, "peak_source": f"{kind} copy bandwidth",
                     "note": "overlapping frames are instruction-issue bound, not HBM bound: see fp32_frac_exact_tests and "
                             "profiles/ for the issue-slot utilisation", "fp32_peak_tflops": ffma},
        "clocks": clocks, "e2e": None, "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------- GPU arm
def main():
    args = ARGS
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config == 2:
        run_config2(args)
        return
    if args.config == 3:
        run_config3(args)
        return

    import torch.distributed as dist
    from ihmr_b200 import _lib, synthetic
    from ihmr_b200 import dist as idist

    rank, world, local_rank = idist.init_from_env("nccl")
    assert world == args.gpus or world == 1, f"launched with WORLD_SIZE={world} but --gpus {args.gpus}"
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if args.scaling == "strong":
        start, F = idist.shard_range(args.frames, rank, world)
        total = args.frames
        assert args.frames % world == 0, "strong scaling: --frames must be divisible by the number of GPUs"
    else:
        F, start, total = args.frames, rank * args.frames, args.frames * world
    root, strategy, model = gpu_setup(args, F, dev, rank)

    # synthetic frames of this rank (ids start ...), targets from the CUDA MANO layer (untimed)
    raw = synthetic.make_raw_frames(start, F, seed=0, mode=args.mode)
    from tools.prof_iters import gpu_targets
    data = gpu_targets(model, raw, dev)
    batch = {k: torch.from_numpy(np.ascontiguousarray(v)).pin_memory() for k, v in data.items()}
    from ihmr_b200.optimize_model import INPUT_KEYS
    h2d = sum(batch[k].numel() * batch[k].element_size() for k in INPUT_KEYS)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def gather():
        local = idist.pack_results(model.params, model.collision_loss_batch, model.joints_3d_loss_p_batch)
        return idist.all_gather_results(local, total)

    def step_resident():
        model.init_optimize()
        model.optimize(0, 1)
        return gather()

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    model.set_input(batch)
    for _ in range(args.warmup):
        step_resident()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = _lib.load().ihmr_launch_count() + model.replayed_launches
    ms_total, _ = timed(step_resident, args.steps)
    launches = _lib.load().ihmr_launch_count() + model.replayed_launches - launches0      # direct + replayed from CUDA graphs
    clocks = sampler.stop() if sampler else None
    ms_step = ms_total / args.steps
    value = total / (ms_step * 1e-3)

    # end to end through the public API: every step copies its inputs from pinned host memory and its results back
    # (run_pipelined overlaps those copies with the neighbouring steps' compute); the all-gather is part of each step
    # (the first H2D and the last D2H of a run are exposed; a production run streams hundreds of batches, so the timed
    # run is at least twelve steps long)
    e2e_steps = args.e2e_steps or max(12, args.steps)
    d2h_box = [0]

    def e2e_run(nsteps=None):
        last = None
        for res in model.run_pipelined([batch] * (nsteps or e2e_steps)):
            gather()
            last = res
        d2h_box[0] = sum(v.nbytes for k, v in last.items() if k not in ("do_flip", "pred_hand_type"))
        return last

    e2e_run(3)                                     # untimed: fills the pool of pinned staging buffers (three result
    ms_e2e, _ = timed(e2e_run, 1)                  # sets are alive at once) and warms the copy streams
    ms_e2e /= e2e_steps
    d2h = d2h_box[0]

    # per-kernel device time of one iteration of every stage (separate pass, CUDA events per launch)
    model.set_input(batch)
    model.init_optimize()
    per_stage = []
    for stage in strategy:
        model.profile_iteration(stage)
        acc = None
        reps = 3
        for _ in range(reps):
            ms = model.profile_iteration(stage)
            acc = ms if acc is None else {k: acc[k] + ms[k] for k in ms}
        per_stage.append({k: v / reps for k, v in acc.items()})
    mean_ms = {k: float(np.mean([s[k] for s in per_stage])) for k in per_stage[0]}
    dom = max(mean_ms, key=mean_ms.get)
    units = F if dom in ("frame_loss", "step") else 2 * F
    peak, _, peak_kind = load_peaks()
    achieved = ALG_BYTES[dom] * units / (mean_ms[dom] * 1e-3) / 1e9
    traffic = None
    tpath = next((q for q in (os.path.join(ROOT, "profiles", n) for n in ("r02_traffic.json", "traffic.json")) if os.path.exists(q)), "")
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if dom in t.get("kernels", {}):
            traffic = t["kernels"][dom]["dram_bytes_per_unit"] * units
    iter_ms = sum(mean_ms.values())
    step_achieved = STEP_BYTES_PER_FRAME_ITER * F / (iter_ms * 1e-3) / 1e9
    issue = None
    ipath = next((q for q in (os.path.join(ROOT, "profiles", n) for n in ("r02_issue.json", "issue.json")) if os.path.exists(q)), "")
    if os.path.exists(ipath):
        issue = json.load(open(ipath)).get(dom)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world, F),
            "us_per_frame_iteration": ms_step * 1e3 / (F * ITERS),
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "how": "OptimizeModel.run_pipelined over pinned host batches: H2D of step k+1 and D2H of step k-1 on copy "
                           "streams while step k refines; first H2D and last D2H are exposed and inside the timed region"},
            "gpu_launches": int(launches), "cuda_graphs": bool(model.use_cuda_graphs),
            "roofline": {"bound": "issue" if dom == "sdf" else "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": f"{peak_kind} copy bandwidth",
                         "avg_launch_ms": mean_ms[dom], "alg_bytes_per_launch": ALG_BYTES[dom] * units,
                         "issue": issue,
                         "note": "frac is the HBM fraction the contract asks for; the penetration kernels are instruction-issue "
                                 "bound (see `issue`: issue-slot utilisation from the committed ncu capture), DRAM traffic is "
                                 "below the algorithmic bytes" if dom == "sdf" else None},
            "step_roofline": {"bound": "issue", "alg_bytes_per_frame_iteration": STEP_BYTES_PER_FRAME_ITER,
                              "achieved": step_achieved, "peak": peak, "unit": "GB/s", "frac": step_achieved / peak,
                              "iteration_ms": iter_ms, "kernel_ms": mean_ms,
                              "kernel_ms_per_stage": per_stage},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_block(root, 1, args.mode)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
