#!/bin/bash
# usage (on a GPU box): VARIANTS="name1 name2" tools/sdf_variants.sh [frames]
# steady-state device time of the penetration kernels inside one stage-2 iteration (hints warm), regular
# library first, then every ihmr_b200/_lib/variants/libihmr_<name>.so built by tools/build_variant.sh
cd "$(dirname "$0")/.."
F=${1:-16384}
run() { IHMR_STATS=1 timeout 200 python tools/prof_iters.py --frames $F --iters 1 --stage 2 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(round(d['sdf'],4))"; }
echo -n "default "; run
for v in $VARIANTS; do echo -n "$v "; IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_$v.so run; done
