#!/bin/bash
# usage (on a GPU box): VARIANTS="name1 name2" tools/bench_variants.sh [bench.py args]
# Runs bench.py with the regular library and with every ihmr_b200/_lib/variants/libihmr_<name>.so built by
# tools/build_variant.sh (selected through IHMR_B200_LIB) and prints frames/s, ms per step and the
# penetration kernel's time per stage.
cd "$(dirname "$0")/.."
ARGS=${@:---steps 2 --warmup 3 --no-cpu-baseline}
P='import sys,json; d=json.loads(sys.stdin.read()); s=d["step_roofline"]["kernel_ms_per_stage"]; print(round(d["value"]), round(d["ms_per_step"],1), "sdf", [round(x["sdf"],3) for x in s])'
echo default; timeout 400 python bench.py $ARGS 2>&1 | tail -1 | python -c "$P"
for v in $VARIANTS; do echo $v; IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_$v.so timeout 400 python bench.py $ARGS 2>&1 | tail -1 | python -c "$P"; done
