python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15
VARIANTS="" tools/sdf_variants.sh 65536
