#!/usr/bin/env python
"""bench.py — decoded frames/s of the VVC pixel-reconstruction back end on synthetic post-parse pictures.

Metric (BASELINE.json): decoded frames/sec, 4K 10-bit Main10 RA, bit-exact vs reference.
Workload (config.workload): `--width x --height` 10-bit 4:2:0 pictures (default 3840x2160 = BASELINE.json configs[2]) drawn by
vvdec_b200.synth.gen_picture (SURVEY §8d config 2/3 model: QT+BT partition, all CUs inter with uni / bi / BCW / BDOF / DMVR /
affine+PROF, residual on ~35 % of CUs with MTS / TS / joint-CbCr, deblocking grids, SAO on 40 % of CTUs, ALF + CC-ALF).
A *step* is one picture through the whole chain K2 -> K1 -> K3 -> K4 -> K5 into a device-resident DPB; pictures cycle through
a GOP of `--gop` distinct work lists and 6 DPB slots (each picture references slots written by earlier steps).

  value  = frames/s with the work lists already resident in HBM (b200_pic_run only), CUDA events on the launching stream.
  e2e    = frames/s through the reference-facing call b200_decompress_picture with pinned HOST work lists (H2D inside the timed
           region) + b200_get_frame of every output picture into pinned host planes (D2H inside the timed region).
  --impl reference: the reference's own CPU implementation (oracle/_ref = unmodified VVdeC kernels, SIMD on, all host threads)
           on the same pictures (see oracle/ref_shim.cpp: ref_decompress_picture_mt).
Multi-GPU (--gpus N under torchrun): closed GOPs are independent -> each rank decodes its own GOP, no data-path
collective; value = total frames / max-over-ranks time ("weak" scaling).
"""
import argparse, ctypes as C, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--gop", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=2, help="pictures timed for the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload
SLOT_ORDER = [4, 5, 0, 2, 1, 3]     # destination slot of step i (mod 6); refs are always slots 0..3 (synth.gen_pus)


def make_gop(args, rank=0):
    from vvdec_b200 import synth
    rng = np.random.default_rng(args.seed + 1000 * rank)
    pics = [synth.gen_picture(rng, args.width, args.height, 10, dst_slot=SLOT_ORDER[i % 6]) for i in range(args.gop)]
    refs = [synth.noise_planes(rng, args.width, args.height, 10) for _ in range(4)]
    return pics, refs


def algorithmic_bytes(args, pic):
    """SURVEY.md §8(d) algorithmic bytes per picture for each kernel family."""
    W, H = args.width, args.height
    S = W * H * 3 // 2
    pus, tus = pic["pus"], pic["tus"]
    nref = (pus["refSlot"] >= 0).sum(axis=1)
    w, h = pus["w"].astype(np.int64), pus["h"].astype(np.int64)
    aff = (pus["flags"] & 8) != 0
    # per PU: luma (w+7)(h+7) + 2 chroma (w/2+3)(h/2+3) read per list, w*h*1.5 written; affine: 6-tap per 4x4 -> (4+5)^2 per sub-block
    rd = np.where(aff, (w // 4) * (h // 4) * 81 + 2 * (w // 8) * (h // 8) * 49, (w + 7) * (h + 7) + 2 * (w // 2 + 3) * (h // 2 + 3))
    mc = 2 * (nref * rd).sum() + 2 * (w * h * 3 // 2).sum() + 64 * len(pus)
    ncoef = ((tus["maxX"].astype(np.int64) + 1) * (tus["maxY"].astype(np.int64) + 1)).sum()
    R = ((1 << tus["log2w"].astype(np.int64)) * (1 << tus["log2h"].astype(np.int64)) * np.where(tus["ict"] != 0, 2, 1)).sum()
    k1 = 2 * ncoef + 4 * R + 32 * len(tus)            # levels + pred read & reco write of the covered samples + records
    n4 = (W // 4) * (H // 4)
    lf = 2 * S + 2 * S + 2 * 6 * n4                    # planes read+written once (V+H counted once, §8d) + both grids
    nctu = ((W + 127) // 128) * ((H + 127) // 128)
    sao = 4 * S + 24 * nctu
    alf = 4 * S + 8 * nctu
    return {"mc": int(mc), "k1": int(k1), "lf": int(lf), "sao": int(sao), "alf": int(alf)}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx, self.samples, self.stop_flag = gpu_index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out: self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples: return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples if len(s) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.samples[0][1]) if self.samples[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.samples)}


def dist_env():
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def pin_pic(lib, pic):
    """Pin every host array of a synthetic picture (cudaHostRegister) so the H2D copies are truly asynchronous."""
    arrs = [pic["pus"], pic["tus"], pic["coefs"], pic.get("lfV"), pic.get("lfH"), pic.get("sao")]
    if "alf" in pic: arrs += [pic["alf"]["ctus"], pic["alf"]["lumaCoeff"], pic["alf"]["lumaClip"]]
    n = 0
    for a in arrs:
        if a is not None and a.nbytes:
            lib.b200_host_register(a.ctypes.data, a.nbytes); n += a.nbytes
    return n


def h2d_bytes(pic):
    n = pic["pus"].nbytes + pic["tus"].nbytes + pic["coefs"].nbytes
    n += 4 * sum(((int(w) + 15) // 16) * ((int(h) + 15) // 16) for w, h in zip(pic["pus"]["w"], pic["pus"]["h"]))   # tile list
    for k in ("lfV", "lfH", "sao"):
        if k in pic: n += pic[k].nbytes
    if "alf" in pic:
        a = pic["alf"]; n += a["ctus"].nbytes + a["lumaCoeff"].nbytes + a["lumaClip"].nbytes + a["chromaCoeff"].nbytes + a["chromaClip"].nbytes + sum(c.nbytes for c in a["cc"])
    return n


def bind_to_gpu_numa_node(local):
    """Run this process (and first-touch its pinned buffers) on the NUMA node the GPU hangs off: H2D/D2H then do not cross the socket link."""
    try:
        import subprocess
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bdf.startswith("00000000:"): bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0: return {"node": None}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-"); cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception as e:
        return {"node": None, "error": str(e)[:80]}


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch, torch.distributed as dist
    import vvdec_b200
    from vvdec_b200 import abi
    rank, world, local = dist_env()
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = vvdec_b200.lib()
    W, H = args.width, args.height
    g = abi.make_geom(W, H, 10)
    pics, refs = make_gop(args, rank)
    ctx = C.c_void_p()
    vvdec_b200.check(lib.b200_ctx_create(C.byref(ctx), C.byref(g), 6, args.gop, local))
    for s in range(4): vvdec_b200.check(lib.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(refs[s])))
    for s in (4, 5): vvdec_b200.check(lib.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(refs[0])))
    for p in pics: pin_pic(lib, p)
    outs = [[np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)] for _ in range(2)]
    for out in outs:
        for o in out: lib.b200_host_register(o.ctypes.data, o.nbytes)
    out = outs[0]
    structs = [p["struct"] for p in pics]

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1: return ms
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    # ---- value: work lists resident in HBM ----
    handles = []
    for st in structs:
        h = lib.b200_pic_upload(ctx, C.byref(st)); assert h >= 0, lib.b200_last_error(); handles.append(h)
    vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
    for i in range(args.warmup): vvdec_b200.check(lib.b200_pic_run(ctx, handles[i % args.gop]))
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = lib.b200_ctx_kernel_launches(ctx)
    vvdec_b200.check(lib.b200_ctx_mark(ctx, 0))
    for i in range(args.steps): vvdec_b200.check(lib.b200_pic_run(ctx, handles[i % args.gop]))
    vvdec_b200.check(lib.b200_ctx_mark(ctx, 1))
    ms = C.c_float(); vvdec_b200.check(lib.b200_ctx_elapsed_ms(ctx, C.byref(ms)))
    barrier()
    launches = lib.b200_ctx_kernel_launches(ctx) - l0
    ms_dev = max_over_ranks(ms.value)

    # ---- per-kernel device time (same workload, events around each family) ----
    lib.b200_ctx_set_profiling(ctx, 1)
    for i in range(args.steps): vvdec_b200.check(lib.b200_pic_run(ctx, handles[i % args.gop]))
    kms = (C.c_float * 8)(); kcnt = (C.c_int * 8)()
    vvdec_b200.check(lib.b200_ctx_get_kernel_ms(ctx, kms, kcnt))
    lib.b200_ctx_set_profiling(ctx, 0)

    # ---- e2e: host work lists in, host frames out, every step ----
    for i in range(max(3, args.warmup // 2)):
        h = lib.b200_decompress_picture(ctx, C.byref(structs[i % args.gop])); assert h >= 0
        vvdec_b200.check(lib.b200_get_frame(ctx, structs[i % args.gop].dstSlot, abi.plane_ptrs(out)))
    barrier()
    t0 = time.perf_counter()
    vvdec_b200.check(lib.b200_ctx_mark(ctx, 0))
    tickets = [None, None]
    nxt = lib.b200_pic_upload(ctx, C.byref(structs[0])); assert nxt >= 0, lib.b200_last_error()
    for i in range(args.steps):
        # every step: H2D of one picture's host work lists (the NEXT picture's: the caller keeps one upload in flight, as a decoder whose
        # parser runs ahead of reconstruction does) + this picture's kernels + D2H of its output frame into one of two pinned host frames;
        # the D2H of step i overlaps the H2D/kernels of step i+1 (copy stream), a host frame is reused only after its copy completed
        cur = nxt
        if i + 1 < args.steps:
            nxt = lib.b200_pic_upload(ctx, C.byref(structs[(i + 1) % args.gop])); assert nxt >= 0, lib.b200_last_error()
        vvdec_b200.check(lib.b200_pic_run(ctx, cur))
        k = i & 1
        if tickets[k] is not None: vvdec_b200.check(lib.b200_frame_wait(ctx, tickets[k]))
        tickets[k] = lib.b200_get_frame_async(ctx, structs[i % args.gop].dstSlot, abi.plane_ptrs(outs[k])); assert tickets[k] >= 0
    for t in tickets:
        if t is not None: vvdec_b200.check(lib.b200_frame_wait(ctx, t))
    vvdec_b200.check(lib.b200_ctx_mark(ctx, 1))
    ms2 = C.c_float(); vvdec_b200.check(lib.b200_ctx_elapsed_ms(ctx, C.byref(ms2)))
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(ms2.value, wall_e2e))
    # ---- the same end-to-end loop with the application's packed output format (pyuv, 4 samples in 5 bytes): the frame is
    # converted on the device, so 15.6 MB instead of 24.9 MB cross PCIe per frame.  Reported beside the headline, which stays 16-bit planes.
    pouts = [[np.zeros(lib.b200_frame_bytes(C.byref(g), 1, c), np.uint8) for c in range(3)] for _ in range(2)]
    for po in pouts:
        for o in po: lib.b200_host_register(o.ctypes.data, o.nbytes)
    pptrs = [(C.c_void_p * 3)(*[o.ctypes.data for o in po]) for po in pouts]
    n_p = min(args.steps, 200)
    barrier()
    t1 = time.perf_counter(); tickets = [None, None]
    nxt = lib.b200_pic_upload(ctx, C.byref(structs[0])); assert nxt >= 0, lib.b200_last_error()
    for i in range(n_p):
        cur = nxt
        if i + 1 < n_p:
            nxt = lib.b200_pic_upload(ctx, C.byref(structs[(i + 1) % args.gop])); assert nxt >= 0, lib.b200_last_error()
        vvdec_b200.check(lib.b200_pic_run(ctx, cur))
        k = i & 1
        if tickets[k] is not None: vvdec_b200.check(lib.b200_frame_wait(ctx, tickets[k]))
        tickets[k] = lib.b200_get_frame_fmt_async(ctx, structs[i % args.gop].dstSlot, 1, pptrs[k]); assert tickets[k] >= 0, lib.b200_last_error()
    for t in tickets:
        if t is not None: vvdec_b200.check(lib.b200_frame_wait(ctx, t))
    vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
    barrier()
    fps_pyuv = world * n_p / max_over_ranks(time.perf_counter() - t1)
    # ---- where the end-to-end time goes (untimed diagnostics): host time inside the upload call, H2D alone, D2H alone ----
    n_diag = min(args.steps, 64)
    t1 = time.perf_counter()
    for i in range(n_diag): assert lib.b200_pic_upload(ctx, C.byref(structs[i % args.gop])) >= 0
    host_up = (time.perf_counter() - t1) * 1e3 / n_diag
    vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
    up_total = (time.perf_counter() - t1) * 1e3 / n_diag
    t1 = time.perf_counter()
    for i in range(n_diag): vvdec_b200.check(lib.b200_get_frame(ctx, 0, abi.plane_ptrs(outs[i & 1])))
    d2h_ms = (time.perf_counter() - t1) * 1e3 / n_diag
    t1 = time.perf_counter(); tk = None
    for i in range(n_diag):
        assert lib.b200_pic_upload(ctx, C.byref(structs[i % args.gop])) >= 0
        if tk is not None: vvdec_b200.check(lib.b200_frame_wait(ctx, tk))
        tk = lib.b200_get_frame_async(ctx, 0, abi.plane_ptrs(outs[i & 1]))
    vvdec_b200.check(lib.b200_frame_wait(ctx, tk)); vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
    both_ms = (time.perf_counter() - t1) * 1e3 / n_diag
    e2e_diag = {"host_ms_in_upload_call": round(host_up, 4), "upload_ms_incl_h2d": round(up_total, 4), "d2h_frame_ms": round(d2h_ms, 4),
                "h2d_and_d2h_concurrent_ms": round(both_ms, 4),
                "e2e_pyuv_output_fps": round(fps_pyuv, 2), "pyuv_d2h_bytes_per_step": int(sum(o.nbytes for o in pouts[0]))}
    sampler.stop_flag = True; sampler.join(timeout=2)

    if rank != 0:
        if world > 1: dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel ----
    peaks = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))); peaks["src"] = "measured"
    except Exception:
        pass
    ab = [algorithmic_bytes(args, p) for p in pics]
    fam = {"mc": [0, 1], "k1": [2], "lf": [3, 4], "sao": [5], "alf": [6, 7]}
    per = {}
    for name, idx in fam.items():
        t = sum(kms[i] for i in idx); n = max(1, max(kcnt[i] for i in idx))
        per[name] = {"ms_per_picture": t / n, "bytes_per_picture": float(np.mean([a[name] for a in ab]))}
    dom = max(per, key=lambda k: per[k]["ms_per_picture"])
    ach = per[dom]["bytes_per_picture"] / (per[dom]["ms_per_picture"] * 1e-3) / 1e9
    total_k = sum(v["ms_per_picture"] for v in per.values())
    # DRAM bytes the family's launches of one picture moved, from the committed ncu pass of this same command (tools/ncu_traffic.py);
    # only valid for the configuration it was captured on
    traffic, traffic_src = None, None
    try:
        if (args.width, args.height) == (3840, 2160):
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic_v15.json")))
            traffic = int(tj["per_family"][dom]["dram_read_bytes"] + tj["per_family"][dom]["dram_write_bytes"]); traffic_src = "profiles/r01_traffic_v15.json"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 4),
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": int(per[dom]["bytes_per_picture"]),
            "note": "family of kernels launched together on forked streams (one launch list per tile class); issue-bound, not HBM-bound (DESIGN.md 5)", "peak_source": peaks["src"] + " (of measured)" if peaks["src"] == "measured" else "fallback",
            "share_of_step": round(per[dom]["ms_per_picture"] / total_k, 3),
            "per_kernel": {k: {"ms": round(v["ms_per_picture"], 4), "GBps": round(v["bytes_per_picture"] / (v["ms_per_picture"] * 1e-3) / 1e9, 1) if v["ms_per_picture"] > 0 else None} for k, v in per.items()}}
    fps = world * args.steps / (ms_dev * 1e-3)
    fps_e2e = world * args.steps / (ms_e2e * 1e-3)
    line = {"metric": "decoded frames/sec, 4K 10-bit Main10 RA, bit-exact YUV vs reference", "value": round(fps, 2), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16 samples / int32 accumulate", "data": "synthetic",
            "config": {"workload": f"{W}x{H} 10-bit 4:2:0 synthetic RA back-end pictures (SURVEY 8d config 3), GOP of {args.gop} work lists, 6-slot DPB, "
                                   "stages K2(MC uni/bi/BCW/BDOF/DMVR/affine+PROF)+K1(dequant/LFNST-off/DCT2/DST7/DCT8/TS/JCCR+reco)+K3 deblock+K4 SAO+K5 ALF/CC-ALF; "
                                   "all CUs inter (intra samples would be given pixels)",
                       "l2": "inputs larger than L2 (6x25 MB DPB + %d work-list arenas cycled)" % args.gop, "parallelism": f"gop-per-gpu x{world}"},
            "e2e": {"value": round(fps_e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": int(np.mean([h2d_bytes(p) for p in pics])),
                    "d2h_bytes_per_step": int(sum(o.nbytes for o in out)), "diag": e2e_diag, "api": "b200_pic_upload (one picture ahead) + b200_pic_run + b200_get_frame_async (D2H overlapped with the next picture), pinned host buffers; every step uploads one picture's work lists and downloads one frame"},
            "gpu_launches": int(launches), "numa": numa, "clocks": sampler.summary(), "roofline": roof}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, pics, refs)
    print(json.dumps(line), flush=True)
    lib.b200_ctx_destroy(ctx)
    if world > 1: dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ CPU legs (the only place bench.py touches oracle/)
def cpu_baseline(args, pics, refs, threads=None):
    """Reference kernels on the host cores for a bounded sample of the same pictures."""
    from tests import helpers
    from vvdec_b200 import abi
    ref = helpers.load_ref()
    g = abi.make_geom(args.width, args.height, 10)
    n = min(args.cpu_sample, len(pics))
    cores = threads or os.cpu_count()
    if ref is not None and hasattr(ref, "ref_decompress_picture_mt"):
        ref.ref_decompress_picture_mt.restype = C.c_double
        try: os.sched_setaffinity(0, range(os.cpu_count()))          # the GPU arm binds itself to the GPU's NUMA node; the CPU baseline gets every core
        except Exception: pass
        ref.ref_decompress_picture_mt(C.byref(g), helpers.ref_ptrs(refs), C.byref(pics[0]["struct"]), cores, 1)   # untimed warm-up (page faults, thread start-up)
        t = 0.0
        for i in range(n):
            t += ref.ref_decompress_picture_mt(C.byref(g), helpers.ref_ptrs(refs), C.byref(pics[i]["struct"]), cores, 1)
        return {"value": round(n / t, 3), "unit": "frames/s", "cores": cores, "kind": "reference",
                "sample": f"{n} of the {len(pics)} GOP pictures, VVdeC kernels ({ref.ref_simd_level().decode()}) via oracle/_ref, {cores} threads"}
    oracle = helpers.load_oracle()
    t0 = time.perf_counter()
    for i in range(n): helpers.oracle_decompress(oracle, g, refs, pics[i])
    t = time.perf_counter() - t0
    return {"value": round(n / t, 3), "unit": "frames/s", "cores": 1, "kind": "port", "sample": f"{n} pictures, scalar C oracle, 1 thread"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0: return
    pics, refs = make_gop(args, 0)
    W, H = args.width, args.height
    vals = []
    a2 = argparse.Namespace(**vars(args)); a2.cpu_sample = 1
    for i in range(args.warmup + args.steps):
        r = cpu_baseline(a2, [pics[i % len(pics)]], refs)
        if i >= args.warmup: vals.append(1.0 / r["value"])
    fps = len(vals) / sum(vals)
    r["value"] = round(fps, 3); r["sample"] = f"each step = 1 picture of the GOP; {r['sample']}"
    line = {"impl": "reference", "metric": "decoded frames/sec, 4K 10-bit Main10 RA, bit-exact YUV vs reference", "value": round(fps, 3), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / fps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16 samples / int32 accumulate", "data": "synthetic",
            "config": {"workload": f"{W}x{H} 10-bit 4:2:0 synthetic RA back-end pictures (same generator/seed as the b200 arm)"},
            "cpu_baseline": r, "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        if a.steps == 400 and a.warmup == 16: a.steps, a.warmup = 4, 1
        run_reference(a)
    else:
        run_b200(a)
