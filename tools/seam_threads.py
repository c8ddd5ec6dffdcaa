"""Seam throughput against the host thread count (run on the GPU box): stock DecLibRecon and DecLibReconB200, one recon instance and two taking pictures in turn.
Prints one JSON object; copy it to profiles/."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import bench
args = bench.parse()
wl = bench.Workload(args, 0)
H = wl.helpers
out = {"workload": bench.workload_config(args, 1)["workload"], "pictures": 8, "unit": "frames/s", "threads": {}}
cases = [wl.B[i % len(wl.B)] for i in range(8)]
for T in (8, 16, 32, 64, os.cpu_count()):
    r = {}
    for name, backend in (("stock", 0), ("b200", 1)):
        run = (lambda c: c.run_stock(threads=T)[2]) if backend == 0 else (lambda c: c.run_b200(threads=T)[2])
        run(cases[0])
        ts = [run(c) for c in cases]
        r[name + "_one_instance"] = round(len(ts) / sum(ts), 1)
        H.seam_pipelined(wl.ref, cases[:2], T, backend, 2, read=False)
        secs, _ = H.seam_pipelined(wl.ref, cases, T, backend, 2, read=False)
        r[name + "_two_instances"] = round(len(cases) / secs, 1) if secs > 0 else None
    out["threads"][T] = r
    print(T, r, file=sys.stderr, flush=True)
print(json.dumps(out))
