#!/bin/bash
./tools/tma_probe
for shape in 128 64 32 2566; do echo "== K6 group shape $shape"; B200_MC_TMA=0 B200_INTRA_GROUP=$shape python tools/k6_probe.py 2>&1 | grep "picture ms"; done
echo "== TMA on"; python -m pytest tests/test_picture_gpu.py -x -q 2>&1 | tail -5
