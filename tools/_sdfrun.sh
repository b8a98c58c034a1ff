timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"default\", d[\"value\"], d[\"ms_per_step\"], [round(x[\"sdf\"],3) for x in d[\"step_roofline\"][\"kernel_ms_per_stage\"]])"
