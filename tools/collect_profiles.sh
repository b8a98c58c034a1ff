#!/bin/bash
# usage (on a GPU box): tools/collect_profiles.sh [tag] [what]   -> gpurun_out/<tag>_*.csv (reports are reduced to CSV on
# the box: the .ncu-rep files are too large to bring back)      what: all | launches | kernels | sdf
cd "$(dirname "$0")/.."
TAG=${1:-r02}; WHAT=${2:-all}
mkdir -p gpurun_out /tmp/ncu
if [ $WHAT = all ] || [ $WHAT = launches ]; then
  # 1) launch list of one whole step (cold-cache, serialised: compare shares)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 2400 --csv --log-file gpurun_out/${TAG}_launches.csv \
      python tools/prof_iters.py --frames 65536 --full-step > gpurun_out/${TAG}_launches.log 2>&1
fi
if [ $WHAT = all ] || [ $WHAT = kernels ]; then
  # 2) every kernel of one warm + one steady iteration per stage: throughput, memory, scheduler, occupancy sections
  timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section SchedulerStats --section WarpStateStats \
      --section Occupancy --section LaunchStats --section InstructionStats --section ComputeWorkloadAnalysis \
      --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum \
      --clock-control none -k regex:^k_ -o /tmp/ncu/${TAG}_kernels -f python tools/prof_stage_iters.py > gpurun_out/${TAG}_kernels.log 2>&1
  ncu -i /tmp/ncu/${TAG}_kernels.ncu-rep --page raw --csv > gpurun_out/${TAG}_kernels_raw.csv 2>/dev/null
fi
if [ $WHAT = all ] || [ $WHAT = sdf ]; then
  # 3) --set full with source correlation of the steady-state stage-2 launch of the dominant kernel (6th k_sdf_dir launch)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sdf_dir -s 5 -c 1 -o /tmp/ncu/${TAG}_sdf -f \
      python tools/prof_stage_iters.py --stages 0,1,2 > gpurun_out/${TAG}_sdf.log 2>&1
  ncu -i /tmp/ncu/${TAG}_sdf.ncu-rep --page raw --csv > gpurun_out/${TAG}_sdf_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/${TAG}_sdf.ncu-rep --page source --csv > gpurun_out/${TAG}_sdf_sass.csv 2>/dev/null
  ncu -i /tmp/ncu/${TAG}_sdf.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/${TAG}_sdf_src.csv 2>/dev/null
  ncu -i /tmp/ncu/${TAG}_sdf.ncu-rep --page details > gpurun_out/${TAG}_sdf_details.txt 2>/dev/null
fi
ls -la gpurun_out | tail -12
