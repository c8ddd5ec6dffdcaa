#!/bin/bash
# Round-2 profiling pass (run under gpurun, one GPU): launch list with DRAM bytes of one B and one I picture of the bench workload, and an ncu --set full
# capture of the same launches.  Outputs in gpurun_out/ (copy the summaries to profiles/).
set -x
ncu --profile-from-start off --clock-control none --kernel-name-base demangled \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --csv --log-file gpurun_out/r02_launches_pictures.csv python tools/prof_pictures.py > gpurun_out/r02_prof_pictures.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled -f -o gpurun_out/r02_full_pictures python tools/prof_pictures.py >> gpurun_out/r02_prof_pictures.log 2>&1
ls -la gpurun_out/ | tail -5
