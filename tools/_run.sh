python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
VARIANTS="nodyn" tools/sdf_variants.sh 65536
python tools/sdf_bench.py --frames 16384 2>&1 | tail -1
IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_nodyn.so python tools/sdf_bench.py --frames 16384 2>&1 | tail -1
