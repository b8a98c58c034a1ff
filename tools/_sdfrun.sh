timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/sdf_stats.py 2>&1 | grep -E "cycles|pairs:|candidates:"
timeout 200 python bench.py --frames 8192 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], [round(x['sdf'],3) for x in d['step_roofline']['kernel_ms_per_stage']])"
