#!/usr/bin/env python
"""Experiment (round 2): GOP-parallel lanes on ONE GPU.  With an I picture per GOP the device spends ~30 % of its time in a latency-bound wave front (K6) that
keeps few SMs busy; a second context decoding another GOP fills them.  N contexts, schedules offset by GOP/N, b200_pic_run issued round-robin."""
import sys, os, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import vvdec_b200
from vvdec_b200 import abi
import bench
args = bench.parse()
lib = vvdec_b200.lib()
wl = bench.Workload(args, 0)
g = abi.make_geom(args.width, args.height, 10)
T = os.cpu_count()
flat = {}
for key, case in [("I", wl.I)] + [(k, c) for k, c in enumerate(wl.B)]:
    pic, _ = case.flatten(threads=T); pic["struct"].dstSlot = 5 if key == "I" else 4
    for a in bench.pic_arrays(pic): lib.b200_host_register(a.ctypes.data, a.nbytes)
    flat[key] = pic
steps = args.steps
for nctx in (1, 2, 3):
    ctxs = []
    for k in range(nctx):
        ctx = C.c_void_p(); vvdec_b200.check(lib.b200_ctx_create(C.byref(ctx), C.byref(g), 6, len(flat), 0))
        for s in range(6): vvdec_b200.check(lib.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(wl.base.refs[s % 4])))
        hs = {key: lib.b200_pic_upload(ctx, C.byref(p["struct"])) for key, p in flat.items()}
        vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
        ctxs.append((ctx, hs, k * args.gop // nctx))
    def h_of(hs, i):
        kind, _, k = wl.sched(i)
        return hs["I"] if kind == "I" else hs[k]
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for i in range(steps):
            for ctx, hs, off in ctxs: vvdec_b200.check(lib.b200_pic_run(ctx, h_of(hs, i + off)))
        for ctx, hs, off in ctxs: vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
        dt = time.perf_counter() - t0
    print(f"lanes {nctx}: {steps * nctx / dt:.1f} frames/s aggregate over {steps} steps per lane ({dt / steps * 1e3:.3f} ms per round)", flush=True)
    for ctx, hs, off in ctxs: lib.b200_ctx_destroy(ctx)
