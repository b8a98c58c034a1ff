#!/bin/bash
# usage (on a 1-GPU box): tools/final_measurements.sh [tag]  -> gpurun_out/<tag>_bench_*.json (+ the profiles of collect_profiles.sh)
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python bench.py "$@" 2> gpurun_out/${TAG}_bench_${name}.err | tail -1 > gpurun_out/${TAG}_bench_${name}.json; echo "$name: $(head -c 300 gpurun_out/${TAG}_bench_${name}.json)"; }
run n1 --steps 3 --warmup 3
run reference --impl reference --steps 2 --warmup 1
run cfg2 --config 2
run cfg3 --config 3
run cfg5 --config 5 --steps 3 --warmup 3 --no-cpu-baseline
run f8192 --frames 8192 --steps 3 --warmup 3 --no-cpu-baseline
run f512 --frames 512 --steps 3 --warmup 3 --no-cpu-baseline
tools/collect_profiles.sh $TAG all > gpurun_out/${TAG}_collect.log 2>&1
ls -la gpurun_out | tail -25
