#!/bin/bash
python -m pytest tests/test_k6_intra_gpu.py tests/test_k1_gpu.py -q 2>&1 | tail -4
for shape in 0 1286 1288 6412; do echo "== K6 group shape $shape"; B200_INTRA_GROUP=$shape python tools/k6_probe.py 2>&1 | grep "I picture ms"; done
for wh in "128 128" "256 128" "512 128" "3840 128" "128 256" "128 512" "128 2176" "1024 1024"; do set -- $wh; echo "== $1 x $2"; python tools/k6_probe.py --width $1 --height $2 2>&1 | grep "I picture ms"; done
