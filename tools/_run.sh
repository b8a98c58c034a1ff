python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sdf or loop" 2>&1 | tail -2
VARIANTS="base" tools/sdf_variants.sh 65536
