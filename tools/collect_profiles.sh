#!/bin/bash
# usage (on a GPU box): tools/collect_profiles.sh [tag]   -> gpurun_out/<tag>_*.{csv,ncu-rep,json,log}
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
# 1) launch list of one whole step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ihmr -c 2400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/prof_iters.py --frames 65536 --full-step > gpurun_out/${TAG}_launches.log 2>&1
# 2) --set full of every kernel of one warm + one steady iteration per stage
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:ihmr -o gpurun_out/${TAG}_full -f \
    python tools/prof_stage_iters.py > gpurun_out/${TAG}_full.log 2>&1
tail -3 gpurun_out/${TAG}_full.log
