#!/usr/bin/env python
"""Experiment: how much throughput is left on the table by running pictures strictly one after the other?  N contexts on ONE GPU, each
with its own streams and resident work lists, b200_pic_run issued round-robin; aggregate frames/s vs one context."""
import sys, os, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vvdec_b200
from vvdec_b200 import abi, synth
import torch

lib = vvdec_b200.lib()
W, H = 3840, 2160
g = abi.make_geom(W, H, 10)
rng = np.random.default_rng(1)
refs = [synth.noise_planes(rng, W, H, 10) for _ in range(4)]
pics = [synth.gen_picture(rng, W, H, 10, dst_slot=[4, 5, 0, 2, 1, 3][i % 6]) for i in range(4)]
for nctx in (1, 2, 3):
    ctxs = []
    for k in range(nctx):
        ctx = C.c_void_p(); vvdec_b200.check(lib.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 4, 0))
        for s in range(6): vvdec_b200.check(lib.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(refs[s % 4])))
        hs = [lib.b200_pic_upload(ctx, C.byref(p["struct"])) for p in pics]
        vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
        ctxs.append((ctx, hs))
    for it in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        n = 300
        for i in range(n):
            for ctx, hs in ctxs: vvdec_b200.check(lib.b200_pic_run(ctx, hs[i % 4]))
        for ctx, hs in ctxs: vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
        dt = time.perf_counter() - t0
    print(f"contexts {nctx}: {n * nctx / dt:.1f} frames/s aggregate ({dt / n * 1e3:.3f} ms per round)", flush=True)
    for ctx, hs in ctxs: lib.b200_ctx_destroy(ctx)
