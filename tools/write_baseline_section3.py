"""Rewrites section 3 of BASELINE.md ("numbers measured by the build") from the files under profiles/.

    python tools/write_baseline_section3.py [tag]      (default r02)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(ROOT, "profiles", tag + "_")


def load(name):
    path = P + name + ".json"
    return json.load(open(path)) if os.path.exists(path) else None


n1, ref, c2, c3, c5 = (load(k) for k in ("bench_n1", "bench_reference", "bench_cfg2", "bench_cfg3", "bench_cfg5"))
n2, n4, n8, s8 = (load(k) for k in ("bench_n2", "bench_n4", "bench_n8", "bench_strong_n8"))
c58 = load("bench_cfg5_n8")
c5_multi = (f"; 8 B200 (65536 frames per GPU): **{c58['value']:,.0f}** frames/s = {c58['value'] / (8 * c5['value']):.3f} of 8x" if c58 else
            "; per GPU the work is independent, so 8 GPUs scale it like config 4")
sw = {(s["mode"], s["frames"]): s for s in c3["sweep"]}
rf = n1["roofline"]
multi = ""
if n2 and n4 and n8:
    multi = (f"; 2 / 4 / 8 GPUs (65536 frames per GPU): {n2['value']:,.0f} / {n4['value']:,.0f} / **{n8['value']:,.0f}** "
             f"({n8['value'] / (8 * n1['value']):.3f} of 8x the 1-GPU value of that build; e2e {n8['e2e']['value']:,.0f})")
if s8:
    multi += f"; 65536 frames split over 8 GPUs: {s8['value']:,.0f} ({s8['ms_per_step']:.1f} ms per step)"
new = f"""## 3. Numbers measured by the build (B200, sm_100a; every figure is in a file under `profiles/`, tables in `profiles/README.md`)

| config (`BASELINE.json.configs`) | metric | value |
|---|---|---|
| 1. one synthetic frame, reference host loop on CPU | s / frame (100 iters); golden vectors written | {1 / ref['value']:.1f} s / frame on {ref['cpu_baseline']['cores']} host threads (`--impl reference`: 2 frames x 100 iterations per timed step, nothing extrapolated); goldens in `tests/golden/loop_*.npz` from the UNMODIFIED reference loop, including two frames at the shipped strategy length (1,204 steps) |
| 2. MANO fwd+bwd, 2x4096 hands, 1 B200 | us/hand; GB/s vs 19,520 B/hand; tensor-pipe %, FP32-pipe % | {c2['value']:.4f} us/hand (fwd {c2['fwd_ms']:.3f} ms + bwd {c2['bwd_ms']:.3f} ms per 8192 hands); {c2['hbm']['achieved']:.0f} GB/s vs the op-boundary bytes ({100 * c2['hbm']['frac']:.1f} % of HBM); {c2['roofline']['achieved']:.1f} TFLOP/s = {100 * c2['roofline']['frac']:.0f} % of the FFMA peak measured in the same run ({c2['roofline']['peak']:.1f} TFLOP/s); blend contraction {c2['tensor']['achieved']:.1f} algorithmic TFLOP/s on tcgen05 (3xTF32; tensor pipe 43-46 % at 65536 frames, per kernel in `profiles/{tag}_ncu_summary.csv`) |
| 3. penetration fwd+bwd, B = 1...16384, 1 B200 | us/frame; GB/s vs 43,572 B/frame; point-triangle tests/s | typical frames: {sw[('typical', 1)]['us_per_frame']:.0f} us (B=1), {sw[('typical', 64)]['us_per_frame']:.2f} us/frame (B=64), {sw[('typical', 1024)]['us_per_frame']:.3f} (B=1024), {sw[('typical', 16384)]['us_per_frame']:.3f} (B=16384) = {sw[('typical', 16384)]['GBps_vs_alg']:.0f} GB/s vs the algorithmic bytes, {sw[('typical', 16384)]['tests_per_s'] / 1e9:.1f} G exact tests/s; near-coincident hands: {sw[('collision', 16384)]['us_per_frame']:.3f} us/frame at B=16384, {sw[('collision', 16384)]['tests_per_s'] / 1e9:.1f} G exact tests/s (stateless calls; inside the loop the hints make it cheaper) |
| 4. full loop, 65536 frames x 100 iters, 1/2/4/8 B200 | frames/s; scaling efficiency; roofline.achieved | 1 GPU: **{n1['value']:,.0f}** frames/s resident ({n1['ms_per_step']:.1f} ms per step), **{n1['e2e']['value']:,.0f}** end to end with host buffers{multi}; penetration kernels {rf['avg_launch_ms']:.2f} ms per launch on average = {rf['achieved']:.0f} GB/s vs the algorithmic bytes = {rf['frac']:.3f} of the measured HBM peak (issue bound, not HBM bound); fused step {n1['step_roofline']['frac']:.3f} |
| 5. worst-case interpenetration | frames/s; slowdown vs config 4 | {c5['value']:,.0f} frames/s on one B200 ({n1['value'] / c5['value']:.2f}x slower than config 4){c5_multi} |
| CPU baseline (oracle port of the reference loop, {ref['cpu_baseline']['cores']} host threads) | frames/s | {ref['value']:.3f} (`--impl reference`), {n1['cpu_baseline']['value']:.3f} (`cpu_baseline` inside the GPU run: 1 frame x 100 iterations) |
"""
path = os.path.join(ROOT, "BASELINE.md")
s = open(path).read()
i = s.index("## 3. Numbers measured by the build")
open(path, "w").write(s[:i] + new)
print("BASELINE.md section 3 rewritten")
