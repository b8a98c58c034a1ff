VARIANTS="occ5 occ5b" tools/sdf_variants.sh 65536
for v in b200 ; do python tools/sdf_bench.py --frames 16384 --mode collision 2>&1 | tail -1; done
IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_occ5.so python tools/sdf_bench.py --frames 16384 --mode collision 2>&1 | tail -1
IHMR_B200_LIB=ihmr_b200/_lib/variants/libihmr_occ5b.so python tools/sdf_bench.py --frames 16384 --mode collision 2>&1 | tail -1
