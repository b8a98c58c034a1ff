python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sdf" 2>&1 | tail -3
VARIANTS="old nodyn dyn4 dyn2 unroll" tools/sdf_variants.sh 65536
