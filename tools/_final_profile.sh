mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_n1.json 2> gpurun_out/bench_r01_n1.err; tail -c 300 gpurun_out/bench_r01_n1.err; cut -c1-200 gpurun_out/bench_r01_n1.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_ref.json 2>gpurun_out/bench_r01_ref.err; cut -c1-200 gpurun_out/bench_r01_ref.json
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/launches_r01.csv | cut -c1-120
timeout 500 ncu --set full --clock-control none -k regex:"k_shape_fwd|k_shape_bwd|k_rigid" -s 3 -c 5 -f -o gpurun_out/top_r01d python tools/prof_iters.py --frames 65536 --full-step --epochs 2 > gpurun_out/prof_top.log 2>&1; tail -1 gpurun_out/prof_top.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_sdf" -s 12 -c 1 -f -o gpurun_out/sdf_full_r1i python tools/prof_iters.py --frames 65536 --full-step --epochs 2 > gpurun_out/prof_sdf.log 2>&1; tail -1 gpurun_out/prof_sdf.log
ls -la gpurun_out | tail -8
