B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
P='import sys,json; d=json.loads(sys.stdin.read()); s=d["step_roofline"]["kernel_ms_per_stage"]; print(d["value"], d["ms_per_step"], [round(x["sdf"],3) for x in s], "stage3 fwd/bwd", round(s[3]["skin_fwd"],3), round(s[3]["skin_bwd"],3))'
for v in $VARIANTS; do echo $v; IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_$v.so timeout 300 $B 2>&1 | tail -1 | python -c "$P"; done
