timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/sdf_stats.py 2>&1 | grep -E "cycles per|candidates:|pairs:|candidates\+"
