#!/bin/bash
# Round-2 GPU pass (one GPU, under gpurun): parity suite, bench (both arms).  Outputs in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02_tests_gpu.log
tail -6 gpurun_out/r02_tests_gpu.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_ref_g.json 2> gpurun_out/r02_ref_g.err; tail -c 300 gpurun_out/r02_ref_g.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_g.json 2> gpurun_out/r02_bench_g.err; tail -c 400 gpurun_out/r02_bench_g.err
python - <<'PY'
import json
for f in ("gpurun_out/r02_ref_g.json", "gpurun_out/r02_bench_g.json"):
    try:
        d = json.load(open(f)); print(f, d["value"], d["e2e"]["value"], d.get("picture_ms"), d.get("seam"), d.get("cpu_baseline", {}).get("value"))
        if "roofline" in d: print({k: v for k, v in d["roofline"].items() if k != "per_kernel"}); print(d["roofline"]["per_kernel"])
    except Exception as e: print(f, "failed", e)
PY
python tools/k6_probe.py 2>&1 | grep "picture ms"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
