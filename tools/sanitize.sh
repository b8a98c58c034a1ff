#!/bin/bash
# usage (on a GPU box): tools/sanitize.sh  -> gpurun_out/sanitizer_<tool>.log + one summary line per tool
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  for lib in default smallcaps; do
    if [ $lib = smallcaps ]; then export IHMR_B200_LIB=ihmr_b200/_lib/libihmr_b200_smallcaps.so; else unset IHMR_B200_LIB; fi
    log=gpurun_out/sanitizer_${tool}_${lib}.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 400 python tools/sanitize_case.py > $log 2>&1
    echo "$tool $lib rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) ; $(grep -c 'done' $log) run(s) completed"
  done
done
