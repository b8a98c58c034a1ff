python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -x -q -m gpu -k "sdf or loop or stage or capacity" 2>&1 | tail -2
VARIANTS="base" tools/sdf_variants.sh 65536
python tools/sdf_bench.py --frames 16384 2>&1 | tail -1
python tools/sdf_bench.py --frames 16384 --mode collision 2>&1 | tail -1
IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_base.so python tools/sdf_bench.py --frames 16384 --mode collision 2>&1 | tail -1
