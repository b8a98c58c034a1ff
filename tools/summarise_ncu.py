"""Reduce `ncu --page raw --csv` of tools/prof_stage_iters.py (65536 frames) to the tables under profiles/.

    python tools/summarise_ncu.py gpurun_out/r02_kernels_raw.csv profiles/r02 [frames]

Launch order of prof_stage_iters.py: per stage one warm iteration, one steady-state iteration, k_step.  The steady-state
instance of every kernel is the LAST one before the stage's k_step.  Writes <prefix>_ncu_summary.csv (one row per
stage x kernel), <prefix>_traffic.json (DRAM bytes per launch unit, read by bench.py) and <prefix>_issue.json
(issue-slot utilisation of the penetration kernel, read by bench.py)."""
import csv
import json
import sys

src, prefix = sys.argv[1], sys.argv[2]
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
rows = list(csv.reader(open(src)))
hdr, unit_row = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
SCALE = {"Tbyte/s": 1e12, "Gbyte/s": 1e9, "Mbyte/s": 1e6, "Kbyte/s": 1e3, "byte/s": 1.0, "s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6,
         "Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}      # times are normalised to ms


def val(r, name, default=None):
    i = idx.get(name)
    if i is None or r[i] in ("", "n/a"):
        return default
    try:
        return float(r[i].replace(",", "")) * SCALE.get(unit_row[i], 1.0)
    except ValueError:
        return default


def short(n):
    return n.split("(")[0].replace("void ", "").replace("ihmr::", "")


inst = [(short(r[idx["Kernel Name"]]), r) for r in rows[2:] if len(r) == len(hdr)]
# skip the target generation (MANO layer in 8192-hand chunks) that precedes the first penetration launch
first = next(i for i, (n, _) in enumerate(inst) if n == "k_sdf_prep")
start = max(j for j in range(first) if inst[j][0].startswith("k_pose_prep"))
inst = inst[start:]
stages, cur = [], []
for n, r in inst:
    cur.append((n, r))
    if n == "k_step":
        stages.append(cur)
        cur = []
UNIT = {"k_frame_loss": frames, "k_step": frames, "k_sdf_prep": frames, "k_sdf_dir": frames}
out_rows, traffic, issue = [], {}, {}
for s, launches in enumerate(stages):
    last = {}
    for n, r in launches:
        last[n] = r                      # the steady-state instance overwrites the warm one
    for n, r in last.items():
        ms = val(r, "gpu__time_duration.sum")
        dram = val(r, "dram__bytes.sum.per_second")
        dram_bytes = None
        if val(r, "dram__bytes_read.sum") is not None:
            dram_bytes = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
        elif dram is not None and ms is not None:
            dram_bytes = dram * ms * 1e-3
        units = UNIT.get(n, 2 * frames)
        row = {
            "stage": s, "kernel": n, "ms": ms,
            "dram_bytes": dram_bytes, "dram_bytes_per_unit": None if dram_bytes is None else dram_bytes / units,
            "unit": "frame" if units == frames else "hand",
            "dram_pct": val(r, "dram__cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "ipc_per_sm": val(r, "sm__inst_executed.avg.per_cycle_active"),
            "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or val(r, "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            "lanes_per_inst": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "fma_pipe_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "alu_pipe_pct": val(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
            "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            "lsu_wavefronts_pct": val(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers": val(r, "launch__registers_per_thread"),
            "stall_barrier": val(r, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
            "stall_long_scoreboard": val(r, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
            "stall_short_scoreboard": val(r, "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
            "stall_wait": val(r, "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
        }
        out_rows.append(row)
        if s == 2 or n not in traffic:
            traffic[n] = {"dram_bytes_per_unit": row["dram_bytes_per_unit"], "unit": row["unit"], "stage": s}
        if n.startswith("k_sdf_dir") and s == 2:
            issue["sdf"] = {"kernel": "k_sdf_dir (steady state, stage 2)", "issue_active_pct": row["issue_active_pct"],
                            "ipc_per_sm": row["ipc_per_sm"], "ipc_peak": 4.0, "lanes_per_inst": row["lanes_per_inst"],
                            "dram_pct": row["dram_pct"], "warps_active_pct": row["warps_active_pct"],
                            "source": src.split("/")[-1]}
cols = list(out_rows[0].keys())
with open(prefix + "_ncu_summary.csv", "w") as fh:
    w = csv.DictWriter(fh, fieldnames=cols)
    w.writeheader()
    for r in out_rows:
        w.writerow({k: (f"{v:.4g}" if isinstance(v, float) else v) for k, v in r.items()})
# bench.py looks the dominant kernel class up by its slot name
alias = {"sdf": ["k_sdf_dir", "k_sdf_prep"], "skin_bwd": ["k_skin_bwd", "k_skin_bwd_tips"], "skin_fwd": ["k_skin_fwd_tc"],
         "blend_fwd": ["k_gemm_tf32x3<256>"], "blend_bwd": ["k_gemm_tf32x3<160>"]}
kern = {}
for slot, names in alias.items():
    # (template arguments vary between builds: match the kernel name by prefix, the stage-2 instance wins)
    found = [next((k for k in traffic if k.startswith(nm) and traffic[k]["stage"] == 2), next((k for k in traffic if k.startswith(nm)), None))
             for nm in names]
    found = [k for k in found if k is not None and traffic[k]["dram_bytes_per_unit"] is not None]
    if found:
        per_hand = sum(traffic[k]["dram_bytes_per_unit"] / (2 if traffic[k]["unit"] == "frame" else 1) for k in found)
        kern[slot] = {"dram_bytes_per_unit": per_hand, "unit": "hand (half frame)" if slot == "sdf" else "hand", "kernels": found}
json.dump({"source": f"{src.split('/')[-1]}: ncu (SpeedOfLight / MemoryWorkloadAnalysis sections), {frames} frames, steady-state iteration of "
                     "stage 2 (dram__bytes.sum over the launch); per launch unit of bench.py (hand = half a frame)",
           "kernels": kern, "all": traffic}, open(prefix + "_traffic.json", "w"), indent=1)
json.dump(issue, open(prefix + "_issue.json", "w"), indent=1)
for r in out_rows:
    print(r["stage"], f"{r['kernel']:24s}", f"{r['ms']:.3f} ms" if r["ms"] else "", "dram/unit", None if r["dram_bytes_per_unit"] is None else round(r["dram_bytes_per_unit"]),
          "ipc", r["ipc_per_sm"], "issue%", r["issue_active_pct"], "fma%", r["fma_pipe_pct"], "tensor%", r["tensor_pipe_pct"], "dram%", r["dram_pct"])
