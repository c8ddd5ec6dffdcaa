#!/bin/bash
# usage: tools/ncu_capture.sh <out-name> <kernel-regex> <skip> <count>   (runs under gpurun; writes gpurun_out/<out-name>.ncu-rep)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s "$3" -c "$4" \
    -o "gpurun_out/$1" -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > "gpurun_out/$1.log" 2>&1
