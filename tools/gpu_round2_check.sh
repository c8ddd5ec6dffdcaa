#!/bin/bash
# Round-2 GPU pass (one GPU, under gpurun): parity suite, K6 group-shape probe, bench with and without the TMA windows.  Outputs in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_tests_gpu.log
tail -6 gpurun_out/r02_tests_gpu.log
B200_MC_TMA=0 python -m pytest tests/test_picture_gpu.py tests/test_seam_gpu.py tests/test_baseline_configs_gpu.py tests/test_golden_gpu.py -q 2>&1 | tail -8 > gpurun_out/r02_tests_gpu_notma.log
tail -3 gpurun_out/r02_tests_gpu_notma.log
for shape in 128 64 32 3212; do echo "== K6 group shape $shape"; B200_INTRA_GROUP=$shape python tools/k6_probe.py 2>&1 | grep "picture ms"; done > gpurun_out/r02_k6_shapes.log 2>&1
cat gpurun_out/r02_k6_shapes.log
python bench.py > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; tail -c 400 gpurun_out/r02_bench_f.err
B200_MC_TMA=0 python bench.py --no-seam > gpurun_out/r02_bench_f_notma.json 2> gpurun_out/r02_bench_f_notma.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_f.json", "gpurun_out/r02_bench_f_notma.json"):
    try:
        d = json.load(open(f)); print(f, d["value"], d["e2e"]["value"], d.get("picture_ms"), {k: round(v["ms_per_step"], 4) for k, v in d["roofline"]["per_kernel"].items()})
    except Exception as e: print(f, "failed", e)
PY
