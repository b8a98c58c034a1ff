#!/bin/bash
# usage (on an N-GPU box): tools/scaling_runs.sh N [tag]  -> gpurun_out/<tag>_bench_n<N>.json (weak) and <tag>_bench_strong_n<N>.json
cd "$(dirname "$0")/.."
N=$1; TAG=${2:-r02}
mkdir -p gpurun_out
run() { out=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
          bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/${out}.err | tail -1 > gpurun_out/${out}.json; echo "$out: $(head -c 260 gpurun_out/${out}.json)"; }
run ${TAG}_bench_n${N} --scaling weak
run ${TAG}_bench_strong_n${N} --scaling strong
