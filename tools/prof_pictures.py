"""One B picture and one I picture of the bench workload through b200_pic_run between cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r02_full python tools/prof_pictures.py
   ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --csv --log-file ... python tools/prof_pictures.py"""
import sys, os, ctypes as C, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
import vvdec_b200
from vvdec_b200 import abi
import bench
args = bench.parse()
lib = vvdec_b200.lib()
wl = bench.Workload(args, 0)
g = abi.make_geom(args.width, args.height, 10)
ctx = C.c_void_p(); vvdec_b200.check(lib.b200_ctx_create(C.byref(ctx), C.byref(g), 6, 2, 0))
for s in range(6): vvdec_b200.check(lib.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(wl.base.refs[s % 4])))
hs = []
for name, case in (("B", wl.B[0]), ("I", wl.I)):
    pic, _ = case.flatten(threads=os.cpu_count()); pic["struct"].dstSlot = 4
    h = lib.b200_pic_upload(ctx, C.byref(pic["struct"])); assert h >= 0
    hs.append((name, h, pic))
for name, h, pic in hs:
    for _ in range(3): vvdec_b200.check(lib.b200_pic_run(ctx, h))
vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
torch.cuda.synchronize()
torch.cuda.profiler.start()
for name, h, pic in hs:
    vvdec_b200.check(lib.b200_pic_run(ctx, h)); vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
torch.cuda.profiler.stop()
print("profiled: 1 B picture (%d PUs, %d TUs, %d intra blocks) then 1 I picture (%d intra blocks)" % (len(hs[0][2]["pus"]), len(hs[0][2]["tus"]), len(hs[0][2].get("intraTus", [])), len(hs[1][2].get("intraTus", []))))
