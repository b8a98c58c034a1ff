#!/bin/bash
# usage (authoring container, after tools/final_measurements.sh / tools/scaling_runs.sh ran on the GPU box and gpurun_out/ was
# merged back): tools/ingest_profiles.sh [tag]  -> profiles/<tag>_*, profiles/README.md, BASELINE.md section 3
cd "$(dirname "$0")/.."
TAG=${1:-r02}
python tools/summarise_ncu.py gpurun_out/${TAG}_kernels_raw.csv profiles/${TAG} > /dev/null
for f in bench_n1 bench_reference bench_cfg2 bench_cfg3 bench_cfg5 bench_f8192 bench_f512 bench_n2 bench_n4 bench_n8 bench_strong_n2 bench_strong_n4 bench_strong_n8; do
  [ -s gpurun_out/${TAG}_$f.json ] && cp gpurun_out/${TAG}_$f.json profiles/${TAG}_$f.json
done
cp gpurun_out/${TAG}_launches.csv profiles/${TAG}_launches.csv
cp gpurun_out/${TAG}_sdf_details.txt profiles/${TAG}_sdf_details.txt
python tools/ncu_src_lines.py gpurun_out/${TAG}_sdf_src.csv > profiles/${TAG}_sdf_source_lines.txt
python tools/sass_histogram.py > profiles/${TAG}_sass_opcodes.csv
python tools/write_profiles_readme.py ${TAG}
python tools/write_baseline_section3.py ${TAG}
