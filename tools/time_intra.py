#!/usr/bin/env python
"""Device time of pictures with intra blocks (K6) at 4K: an I picture (all CUs intra) and a B picture with a share of intra CUs, in-loop filters on.
usage (GPU box): python tools/time_intra.py [W H]"""
import sys, os, json, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vvdec_b200
from vvdec_b200 import abi, synth

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
bd = 10
lib = vvdec_b200.lib()
g = abi.make_geom(W, H, bd)
rng = np.random.default_rng(7)
ctx = C.c_void_p()
vvdec_b200.check(lib.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 2, -1))
dpb = [synth.noise_planes(rng, W, H, bd) for _ in range(4)]
for s in range(4): vvdec_b200.check(lib.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(dpb[s])))
out = {}
for name, frac in (("inter_only", 0.0), ("intra_15pct", 0.15), ("intra_100pct", 1.0)):
    pic = synth.gen_picture(rng, W, H, bd, dst_slot=4, intra_frac=frac) if frac else synth.gen_picture(rng, W, H, bd, dst_slot=4)
    ms = []
    for rep in range(5):
        a = lib.b200_pic_upload(ctx, C.byref(pic["struct"])); assert a >= 0, lib.b200_last_error()
        vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
        vvdec_b200.check(lib.b200_ctx_mark(ctx, 0))
        vvdec_b200.check(lib.b200_pic_run(ctx, a))
        vvdec_b200.check(lib.b200_ctx_mark(ctx, 1))
        t = C.c_float(); vvdec_b200.check(lib.b200_ctx_elapsed_ms(ctx, C.byref(t)))
        vvdec_b200.check(lib.b200_wait_picture(ctx, a, None, 0))
        ms.append(round(t.value, 4))
    out[name] = {"ms": ms, "cus": len(pic["cus"]), "intra_blocks": int(len(pic.get("intraTus", [])))}
print(json.dumps({"geometry": [W, H], "device_ms_per_picture": out}))
lib.b200_ctx_destroy(ctx)
