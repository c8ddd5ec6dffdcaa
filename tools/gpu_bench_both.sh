#!/bin/bash
# Both bench arms as the driver runs them (N = 1).  Outputs in gpurun_out/.
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_ref_final.json 2> gpurun_out/r02_ref_final.err; echo "ref rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "b200 rc=$?"; tail -c 300 gpurun_out/r02_bench_final.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_ref_final.json", "gpurun_out/r02_bench_final.json"):
    try:
        d = json.load(open(f)); print(f, d["value"], d["e2e"]["value"], d.get("picture_ms"), d.get("seam"), d.get("cpu_baseline"))
    except Exception as e: print(f, "failed", e)
PY
