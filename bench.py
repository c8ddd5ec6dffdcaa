#!/usr/bin/env python
"""bench.py — decoded frames/s of the VVC pixel-reconstruction back end behind VVdeC's DecLibRecon seam.

Metric (BASELINE.json): decoded frames/sec, 4K 10-bit Main10 RA, bit-exact vs reference.
Workload (config.workload, identical in both arms): `--width x --height` 10-bit 4:2:0 synthetic PARSED pictures of a random-access GOP
(SURVEY §8d config 3): one I picture per `--gop` steps, B pictures otherwise with 85 % inter / 15 % intra CUs — merge / MMVD / GEO / CIIP /
affine (+PROF) / AMVP with AMVR, BCW, SMVD, BDOF and DMVR where the POC distances allow, residual with MTS / LFNST / SBT / TS / joint CbCr,
intra CUs with angular / MRL / MIP / CCLM / BDPCM — followed by deblocking, SAO, ALF + CC-ALF.  The pictures are built by oracle/ref_seam.h as
real VVdeC `Picture` objects (CodingStructure through the reference's Partitioner / addCU / addTU, levels in the reconstruction plane,
motion left as merge / AMVP syntax): exactly what DecLibRecon::decompressPicture receives from the parser.

  --impl reference : the reference's own DecLibRecon (create / decompressPicture / waitForPrevDecompressedPic, DecLibRecon.cpp:127,429,684)
                     with a ThreadPool of all host threads on those pictures — BASELINE.md level B1.  A step = one picture; the time of a
                     step is the time between decompressPicture() and the return of waitForPrevDecompressedPic().
  --impl b200 (default): the same pictures through the product:
     value  = frames/s of the device chain (b200_pic_run: K2 -> K1 -> K6 -> K3 -> K4 -> K5) with the pictures' work lists resident in HBM —
              the lists are what the drop-in class DecLibReconB200 flattens from each parsed Picture (untimed preparation, like the
              reference arm's picture construction); CUDA events on the launching stream.
     e2e    = frames/s through the C ABI with HOST buffers: every step uploads one picture's pinned host work lists (b200_pic_upload,
              bucketing kernels included), runs it and copies the 16-bit output frame to pinned host memory, inside the timed region.
     seam   = frames/s of DecLibReconB200::decompressPicture + waitForPrevDecompressedPic on live parsed Pictures (host MIDER + boundary
              strengths + flatten on the decoder's thread pool, H2D, device chain, DMVR read-back, TaskFinishMotionInfo, D2H into
              Picture::m_bufs) — the number a VVdeC built with this back end would see per recon instance, reported beside e2e.
Multi-GPU (--gpus N under torchrun): closed GOPs are independent -> each rank decodes its own GOPs (vvdec_b200.gop_shard), no
data-path collective while decoding; finished frames are gathered to rank 0 in display order over NCCL (--gather, on a side stream).
value = total frames / max-over-ranks time ("weak" scaling).

oracle/_ref (the compiled, unmodified reference) is used for three things only: building the parsed pictures (both arms), the reference arm /
cpu_baseline leg, and hosting the product's glue class for the seam leg (the glue lives inside a VVdeC build by design).  Every timed GPU
region runs libvvdec_b200.so alone.
"""
import argparse, ctypes as C, json, os, subprocess, sys, threading, time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np

METRIC = "decoded frames/sec, 4K 10-bit Main10 RA, bit-exact YUV vs reference"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--gop", type=int, default=32, help="one I picture every GOP steps")
    ap.add_argument("--distinct", type=int, default=6, help="distinct B pictures cycled through")
    ap.add_argument("--intra-pct", type=int, default=15)
    ap.add_argument("--isp-pct", type=int, default=20, help="share of the eligible intra luma CUs coded with intra sub-partitions")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=6, help="B pictures timed for the cpu_baseline leg (plus one I picture)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-seam", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="multi-GPU: leave the frames on their GPUs")
    ap.add_argument("--gather-timeout", type=float, default=120.0, help="multi-GPU: seconds the display-order gather pass may take before the line is printed without it")
    ap.add_argument("--host-threads", type=int, default=0, help="threads of the reference's ThreadPool (0: all)")
    ap.add_argument("--lanes", type=int, default=2, help="GOP-parallel lanes per GPU: device contexts that reconstruct different GOPs side by side (1: one picture at a time)")
    ap.add_argument("--recon-depth", type=int, default=2, help="recon instances taking pictures in turn, as DecLib runs them (DecLib.h:70); 1: one picture at a time")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ workload (both arms)
def workload_config(args, world):
    W, H = args.width, args.height
    return {"workload": f"{W}x{H} 10-bit 4:2:0 synthetic PARSED RA pictures at the DecLibRecon seam (oracle/ref_seam.h, seed {args.seed}): 1 I picture per {args.gop} steps, "
                        f"B pictures with {args.intra_pct} % intra CUs otherwise ({args.distinct} distinct, cycled); merge/MMVD/GEO/CIIP/affine+PROF/AMVP+AMVR/BCW/SMVD/BDOF/DMVR, "
                        f"residual MTS/LFNST/SBT/TS/JCCR, intra angular/MRL/MIP/CCLM/BDPCM/ISP ({args.isp_pct} % of the eligible CUs), deblocking + SAO + ALF/CC-ALF",
            "l2": "inputs larger than L2 (6 x 25 MB DPB buffers + work-list arenas cycled)", "parallelism": f"closed GOPs in parallel: {max(1, args.lanes)} lane(s) per GPU x {world} GPU(s)"}


class Workload:
    """The GOP's parsed pictures: one template (reference pictures, filter parameters) and per-picture generator seeds."""
    def __init__(self, args, rank):
        from tests import helpers
        self.helpers, self.args = helpers, args
        self.ref = helpers.load_ref()
        if self.ref is None or not hasattr(self.ref, "ref_seam_create"):
            raise RuntimeError("oracle/_ref/libvvdec_ref.so (the compiled reference + seam shim) is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
        rng = np.random.default_rng(args.seed + 1000 * rank)
        self.base = helpers.SeamCase(self.ref, rng, args.width, args.height, intra=args.intra_pct, isp=args.isp_pct)
        seeds = [int(s) for s in rng.integers(1, 1 << 30, size=args.distinct + 1)]
        self.B = [self.base.variant(s) for s in seeds[:-1]]
        self.I = self.base.variant(seeds[-1], slice_type=2)

    def sched(self, i):
        """Picture of step i: ('I', case) or ('B', case)."""
        if i % self.args.gop == 0: return "I", self.I, -1
        k = (i - 1 - i // self.args.gop) % len(self.B)
        return "B", self.B[k], k


def threads_all(args):
    return args.host_threads or os.cpu_count()


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
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def bind_to_gpu_numa_node(local):
    """Run this process (and first-touch its pinned buffers) on the NUMA node the GPU hangs off: H2D/D2H then do not cross the socket link."""
    try:
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


def pic_arrays(pic):
    arrs = [pic["pus"], pic["tus"], pic["coefs"], pic.get("lfV"), pic.get("lfH"), pic.get("sao"), pic.get("intraTus"), pic.get("wp")]
    if "alf" in pic: arrs += [pic["alf"]["ctus"]] + list(pic["alfArrays"].values())
    return [a for a in arrs if a is not None and a.nbytes]


def h2d_bytes(pic):
    n = sum(a.nbytes for a in pic_arrays(pic))
    n += 4 * sum(((int(w) + 15) // 16) * ((int(h) + 15) // 16) for w, h in zip(pic["pus"]["w"], pic["pus"]["h"])) if False else 0
    return n


def algorithmic_bytes(args, pic):
    """SURVEY.md §8(d) algorithmic bytes per picture for each kernel family."""
    W, H = args.width, args.height
    S = W * H * 3 // 2
    pus, tus = pic["pus"], pic["tus"]
    out = {}
    if len(pus):
        nref = (pus["refSlot"] >= 0).sum(axis=1)
        w, h = pus["w"].astype(np.int64), pus["h"].astype(np.int64)
        aff = (pus["flags"] & 8) != 0
        # per PU: luma (w+7)(h+7) + 2 chroma (w/2+3)(h/2+3) read per list, w*h*1.5 written; affine: 6-tap per 4x4 -> (4+5)^2 per sub-block
        rd = np.where(aff, (w // 4) * (h // 4) * 81 + 2 * (w // 8) * (h // 8) * 49, (w + 7) * (h + 7) + 2 * (w // 2 + 3) * (h // 2 + 3))
        out["mc"] = int(2 * (nref * rd).sum() + 2 * (w * h * 3 // 2).sum() + 64 * len(pus))
    else: out["mc"] = 0
    if len(tus):
        ncoef = ((tus["maxX"].astype(np.int64) + 1) * (tus["maxY"].astype(np.int64) + 1)).sum()
        R = ((1 << tus["log2w"].astype(np.int64)) * (1 << tus["log2h"].astype(np.int64)) * np.where(tus["ict"] != 0, 2, 1)).sum()
        out["k1"] = int(2 * ncoef + 4 * R + 32 * len(tus))     # levels + pred read & reco write of the covered samples + records
    else: out["k1"] = 0
    it = pic.get("intraTus")
    if it is not None and len(it):
        bw, bh = (1 << it["log2w"].astype(np.int64)), (1 << it["log2h"].astype(np.int64))
        resi = (it["flags"] & 4) != 0
        out["intra"] = int((2 * (2 * bw + 2 * bh + 1) + 2 * bw * bh + np.where(resi, 2 * bw * bh, 0)).sum() + 16 * len(it))   # reference samples + block write (+ residual read) + record
    else: out["intra"] = 0
    n4 = (W // 4) * (H // 4)
    out["lf"] = 2 * S + 2 * S + 2 * 6 * n4                    # planes read+written once (V+H counted once, §8d) + both grids
    nctu = ((W + 127) // 128) * ((H + 127) // 128)
    out["sao"] = 4 * S + 24 * nctu
    out["alf"] = 4 * S + 8 * nctu
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200(args):
    import torch, torch.distributed as dist
    import vvdec_b200
    from vvdec_b200 import abi
    rank, world, local = dist_env()
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = "NONE"                                 # stdout carries the JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = vvdec_b200.lib()
    W, H = args.width, args.height
    g = abi.make_geom(W, H, 10)
    try: os.sched_setaffinity(0, range(os.cpu_count()))               # the (untimed) preparation uses every core
    except Exception: pass
    wl = Workload(args, rank)
    T = threads_all(args)
    # ---- preparation (untimed): every distinct parsed picture flattened by the drop-in class's host stages into pinned work lists ----
    flat = {}
    host_stage_s = []
    for key, case in [("I", wl.I)] + [(k, c) for k, c in enumerate(wl.B)]:
        pic, secs = case.flatten(threads=T)
        assert pic is not None, f"DecLibReconB200 refused a workload picture ({secs})"
        pic["struct"].dstSlot = 5 if key == "I" else 4
        for a in pic_arrays(pic): lib.b200_host_register(a.ctypes.data, a.nbytes)
        flat[key] = pic; host_stage_s.append(secs)
    numa = bind_to_gpu_numa_node(local)

    def pic_of(i):
        kind, _, k = wl.sched(i)
        return flat["I"] if kind == "I" else flat[k]

    # GOP-parallel lanes: closed GOPs are independent, and an I picture is a latency-bound wave front that keeps few SMs busy (K6) — a second device context that
    # reconstructs another GOP fills them (tools/exp_two_lanes.py: 1146 -> 1423 frames/s with two, 1334 with three).  Lane j's schedule is offset by j * GOP / lanes.
    L = max(1, args.lanes)

    class Lane:
        pass
    lanes = []
    for j in range(L):
        ln = Lane(); ln.off = j * args.gop // L; ln.ctx = C.c_void_p(); ln.steps = args.steps // L + (1 if j < args.steps % L else 0)
        vvdec_b200.check(lib.b200_ctx_create(C.byref(ln.ctx), C.byref(g), 6, len(flat), local))
        for sl in range(4): vvdec_b200.check(lib.b200_ctx_load_slot(ln.ctx, sl, abi.plane_ptrs(wl.base.refs[sl])))
        for sl in (4, 5): vvdec_b200.check(lib.b200_ctx_load_slot(ln.ctx, sl, abi.plane_ptrs(wl.base.refs[0])))
        ln.outs = [[np.zeros((H, W), np.int16), np.zeros((H // 2, W // 2), np.int16), np.zeros((H // 2, W // 2), np.int16)] for _ in range(2)]
        for out in ln.outs:
            for o in out: lib.b200_host_register(o.ctypes.data, o.nbytes)
        lanes.append(ln)
    ctx, outs = lanes[0].ctx, lanes[0].outs                                 # lane 0 also serves the single-lane diagnostics

    def barrier():
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1: return ms
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    # ---- value: work lists resident in HBM ----
    def upload_all(ln):
        ln.handle = {}
        for key, pic in flat.items():
            h = lib.b200_pic_upload(ln.ctx, C.byref(pic["struct"])); assert h >= 0, lib.b200_last_error(); ln.handle[key] = h
        vvdec_b200.check(lib.b200_wait_picture(ln.ctx, -1, None, 0))

    def lane_h(ln, i):
        kind, _, k = wl.sched(i + ln.off)
        return ln.handle["I"] if kind == "I" else ln.handle[k]
    for ln in lanes: upload_all(ln)
    handle = lanes[0].handle

    def h_of(i):
        return lane_h(lanes[0], i)

    for ln in lanes:
        for i in range(args.warmup): vvdec_b200.check(lib.b200_pic_run(ln.ctx, lane_h(ln, i + 1)))
        vvdec_b200.check(lib.b200_pic_run(ln.ctx, ln.handle["I"]))           # the I picture's path is warm too
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = sum(lib.b200_ctx_kernel_launches(ln.ctx) for ln in lanes)
    for ln in lanes: vvdec_b200.check(lib.b200_ctx_mark(ln.ctx, 0))
    for i in range(max(ln.steps for ln in lanes)):                           # exactly args.steps pictures, dealt to the lanes round-robin
        for ln in lanes:
            if i < ln.steps: vvdec_b200.check(lib.b200_pic_run(ln.ctx, lane_h(ln, i)))
    for ln in lanes: vvdec_b200.check(lib.b200_ctx_mark(ln.ctx, 1))
    ms_l = []
    for ln in lanes:
        t = C.c_float(); vvdec_b200.check(lib.b200_ctx_elapsed_ms(ln.ctx, C.byref(t))); ms_l.append(t.value)
    barrier()
    launches = sum(lib.b200_ctx_kernel_launches(ln.ctx) for ln in lanes) - l0
    ms_dev = max_over_ranks(max(ms_l))                                       # the lanes start together: the slowest lane's span is the step time

    # ---- per-kernel-family device time (same schedule, events around each family) + the I / B split ----
    NF = 10
    lib.b200_ctx_set_profiling(ctx, 1)
    for i in range(args.steps): vvdec_b200.check(lib.b200_pic_run(ctx, h_of(i)))
    kms = (C.c_float * NF)(); kcnt = (C.c_int * NF)()
    vvdec_b200.check(lib.b200_ctx_get_kernel_ms_n(ctx, kms, kcnt, NF))
    lib.b200_ctx_set_profiling(ctx, 0)
    split = {}
    for name, hh in (("I_picture_ms", handle["I"]), ("B_picture_ms", handle[0])):
        for _ in range(2): vvdec_b200.check(lib.b200_pic_run(ctx, hh))
        vvdec_b200.check(lib.b200_ctx_mark(ctx, 0))
        for _ in range(8): vvdec_b200.check(lib.b200_pic_run(ctx, hh))
        vvdec_b200.check(lib.b200_ctx_mark(ctx, 1))
        t = C.c_float(); vvdec_b200.check(lib.b200_ctx_elapsed_ms(ctx, C.byref(t))); split[name] = round(t.value / 8, 4)

    # ---- e2e: host work lists in, host frames out, every step ----
    for ln in lanes:
        for i in range(max(3, args.warmup // 2)):
            p = pic_of(i + ln.off)
            h = lib.b200_decompress_picture(ln.ctx, C.byref(p["struct"])); assert h >= 0, lib.b200_last_error()
            vvdec_b200.check(lib.b200_get_frame(ln.ctx, p["struct"].dstSlot, abi.plane_ptrs(ln.outs[0])))
    barrier()
    t0 = time.perf_counter()
    for ln in lanes:
        vvdec_b200.check(lib.b200_ctx_mark(ln.ctx, 0))
        ln.tickets = [None, None]
        ln.nxt = lib.b200_pic_upload(ln.ctx, C.byref(pic_of(ln.off)["struct"])); assert ln.nxt >= 0, lib.b200_last_error()
    for i in range(max(ln.steps for ln in lanes)):
        # every step and lane: H2D of one picture's host work lists (the NEXT picture's: the caller keeps one upload in flight, as a decoder whose
        # parser runs ahead of reconstruction does) + this picture's kernels + D2H of its output frame into one of two pinned host frames;
        # the D2H of step i overlaps the H2D/kernels of step i+1 (copy stream), a host frame is reused only after its copy completed
        for ln in lanes:
            if i >= ln.steps: continue
            cur = ln.nxt
            if i + 1 < ln.steps:
                ln.nxt = lib.b200_pic_upload(ln.ctx, C.byref(pic_of(i + 1 + ln.off)["struct"])); assert ln.nxt >= 0, lib.b200_last_error()
            vvdec_b200.check(lib.b200_pic_run(ln.ctx, cur))
            k = i & 1
            if ln.tickets[k] is not None: vvdec_b200.check(lib.b200_frame_wait(ln.ctx, ln.tickets[k]))
            ln.tickets[k] = lib.b200_get_frame_async(ln.ctx, pic_of(i + ln.off)["struct"].dstSlot, abi.plane_ptrs(ln.outs[k])); assert ln.tickets[k] >= 0
    ms2 = 0.0
    for ln in lanes:
        for t in ln.tickets:
            if t is not None: vvdec_b200.check(lib.b200_frame_wait(ln.ctx, t))
        vvdec_b200.check(lib.b200_ctx_mark(ln.ctx, 1))
        t = C.c_float(); vvdec_b200.check(lib.b200_ctx_elapsed_ms(ln.ctx, C.byref(t))); ms2 = max(ms2, t.value)
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max_over_ranks(max(ms2, wall_e2e))

    # ---- where the end-to-end time goes (untimed diagnostics): the link alone, in each direction and both at once (the PCIe ceiling) ----
    n_diag = min(args.steps, 32)
    vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
    t1 = time.perf_counter()
    for i in range(n_diag): assert lib.b200_pic_upload(ctx, C.byref(pic_of(i + 1)["struct"])) >= 0
    host_up = (time.perf_counter() - t1) * 1e3 / n_diag
    vvdec_b200.check(lib.b200_wait_picture(ctx, -1, None, 0))
    up_total = (time.perf_counter() - t1) * 1e3 / n_diag
    t1 = time.perf_counter()
    for i in range(n_diag): vvdec_b200.check(lib.b200_get_frame(ctx, 0, abi.plane_ptrs(outs[i & 1])))
    d2h_ms = (time.perf_counter() - t1) * 1e3 / n_diag
    pcie = pcie_ceiling(torch)
    h2d_step = float(np.mean([h2d_bytes(pic_of(i)) for i in range(args.steps)])); d2h_step = int(sum(o.nbytes for o in outs[0]))
    e2e_diag = {"host_ms_in_upload_call": round(host_up, 4), "upload_ms_incl_h2d": round(up_total, 4), "d2h_frame_ms": round(d2h_ms, 4), "pcie": pcie,
                "link_bound_fps": round(1e3 / max(1e-9, max(h2d_step / (pcie["h2d_gbs_duplex"] * 1e6), d2h_step / (pcie["d2h_gbs_duplex"] * 1e6))), 1) if pcie.get("h2d_gbs_duplex") else None}
    sampler.stop_flag = True; sampler.join(timeout=2)

    # ---- seam: DecLibReconB200 live on parsed Pictures (its own device context) ----
    seam = None
    if not args.no_seam and rank == 0 and world == 1:
        try:
            try: os.sched_setaffinity(0, range(os.cpu_count()))
            except Exception: pass
            table = backend_sweep(args, wl, 1)
            fps_seam, Ts, Ds, runs = backend_schedule(args, wl, 1, table)
            seam = {"value": round(fps_seam, 2), "unit": "frames/s", "host_threads": Ts, "recon_instances": Ds, "pictures": args.steps, "sweep_estimate_fps": table, "schedule_fps": runs,
                    "host_stage_ms_per_picture": round(1e3 * float(np.mean(host_stage_s)), 3),
                    "api": "b200glue::DecLibReconB200::decompressPicture + waitForPrevDecompressedPic on the live parsed pictures of the schedule, measured like the reference arm: the faster of "
                           "the two best configurations of a sweep over host threads x recon instances taking pictures in turn (DecLib.h:70)"}
        except Exception as e:
            seam = {"error": f"{type(e).__name__}: {str(e)[:200]}"}

    line = None
    if rank == 0: line = assemble_line(args, world, wl, flat, kms, ms_dev, ms_e2e, h2d_step, d2h_step, e2e_diag, seam, split, launches, numa, sampler, pic_of)
    if world > 1 and not args.no_gather:
        try:
            gather_pass(args, rank, world, local, lib, lanes, upload_all, lane_h, pic_of, line, torch, dist, barrier, max_over_ranks)
        except Exception as e:                                               # the line (throughput without the gather) must not be lost to a failed exchange
            if rank == 0:
                line["gather"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
                print(json.dumps(line), flush=True)
            leave()
    if rank == 0: print(json.dumps(line), flush=True)
    for ln in lanes: lib.b200_ctx_destroy(ln.ctx)
    if world > 1: dist.destroy_process_group()
    leave()


def gather_pass(args, rank, world, local, lib, lanes, upload_all, lane_h, pic_of, line, torch, dist, barrier, max_over_ranks):
    """Multi-GPU: the same schedule once more with the finished frames leaving their GPUs — every lane of every rank reconstructs its own run of pictures, each
    finished frame is copied out of its DPB slot on the lane's side stream and goes to rank 0 in display order over NCCL
    (vvdec_b200/gather.py), overlapping the following pictures.  The timed region ends when the last frame has arrived on rank 0; its throughput becomes the
    line's `value`.  Everything else of the line is already assembled: a watchdog prints it with the gather marked as failed if the exchange does not finish."""
    import vvdec_b200
    from vvdec_b200 import gather
    W, H = args.width, args.height
    done = threading.Event()

    def dog():
        if done.wait(args.gather_timeout): return
        if rank == 0:
            line["gather"] = {"error": f"the display-order gather did not finish within {args.gather_timeout} s; value is the throughput without it"}
            print(json.dumps(line), flush=True)
        os._exit(0)
    barrier()                                                                # the other ranks wait here while rank 0 assembled its line
    threading.Thread(target=dog, daemon=True).start()
    for ln in lanes: upload_all(ln)                                          # the e2e leg cycled through the arenas: the work lists go back into HBM
    numel = W * H * 3                                                        # bytes of a 16-bit 4:2:0 frame
    # display order: every (rank, lane) reconstructs one run of consecutive pictures; the runs follow each other rank by rank, lane by lane.  A rank pushes its
    # frames in the order they finish (lanes interleaved): the n-th pushed frame of rank r is picture i of lane j
    display_of, base = {}, 0
    for r in range(world):
        n, bases = 0, []
        for ln in lanes: bases.append(base); base += ln.steps
        for i in range(max(ln.steps for ln in lanes)):
            for j, ln in enumerate(lanes):
                if i < ln.steps: display_of[(r, n)] = bases[j] + i; n += 1
    for ln in lanes: ln.side = torch.cuda.Stream()
    poff = [0, W * H, W * H + (W // 2) * (H // 2)]
    wall_ms = 0.0
    for timed in (False, True):                                              # the first pass is the warm-up (NCCL sets its peer-to-peer channels up with the first transfer)
        G = gather.FrameGather(rank, world, None, numel, torch.device("cuda", local), display_of=display_of)
        barrier()
        t_wall = time.perf_counter()
        n = 0
        for i in range(max(ln.steps for ln in lanes)):
            for ln in lanes:
                if i >= ln.steps: continue
                vvdec_b200.check(lib.b200_pic_run(ln.ctx, lane_h(ln, i)))
                base_ptr = G.slot(n).data_ptr()
                pl = (C.c_void_p * 3)(base_ptr + 2 * poff[0], base_ptr + 2 * poff[1], base_ptr + 2 * poff[2])
                vvdec_b200.check(lib.b200_get_frame_device_async(ln.ctx, pic_of(i + ln.off)["struct"].dstSlot, pl, C.c_void_p(ln.side.cuda_stream)))
                with torch.cuda.stream(ln.side): G.push(n)
                n += 1
        G.finish()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t_wall) * 1e3                       # the timed region ends when the last frame has arrived on rank 0
        barrier()
        if not timed: del G
    ms_g = max_over_ranks(wall_ms)
    done.set()
    if rank == 0:
        nb = G.bytes_received()
        line["value_without_gather"] = line["value"]
        line["value"] = round(world * args.steps / (ms_g * 1e-3), 2); line["ms_per_step"] = round(ms_g / args.steps, 4)
        line["gather"] = {"frames_received": int(nb // numel), "bytes": int(nb), "GBps_over_timed_region": round(nb / (ms_g * 1e-3) / 1e9, 2),
                          "note": "display-order gather to rank 0 (one batched NCCL send/recv group per step on a side stream), inside the timed region (host clock around it, "
                                  "max over ranks); NVLink 5: 900 GB/s per direction per GPU"}


def assemble_line(args, world, wl, flat, kms, ms_dev, ms_e2e, h2d_step, d2h_step, e2e_diag, seam, split, launches, numa, sampler, pic_of):
    W, H = args.width, args.height
    # ---- roofline of the dominant kernel family ----
    peaks = {"hbm_gbs": 6650.0, "src": "fallback"}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))); peaks["src"] = "measured"
    except Exception:
        pass
    ab = [algorithmic_bytes(args, pic_of(i)) for i in range(args.steps)]
    fam = {"mc": [0, 1], "k1": [2], "lf": [3, 4], "sao": [5], "alf": [6, 7], "intra": [8]}
    per = {}
    for name, idx in fam.items():
        per[name] = {"ms_per_step": sum(kms[i] for i in idx) / args.steps, "bytes_per_step": float(np.mean([a[name] for a in ab]))}
    dom = max(per, key=lambda k: per[k]["ms_per_step"])
    ach = per[dom]["bytes_per_step"] / (per[dom]["ms_per_step"] * 1e-3) / 1e9
    total_k = sum(v["ms_per_step"] for v in per.values())
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if (tj.get("width"), tj.get("height")) == (W, H):
            traffic = int(tj["per_family"][dom]["dram_read_bytes"] + tj["per_family"][dom]["dram_write_bytes"]); traffic_src = "profiles/r02_traffic.json"
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(ach / peaks["hbm_gbs"], 4),
            "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": int(per[dom]["bytes_per_step"]),
            "note": "family time = CUDA events around the family's launches inside b200_pic_run, averaged over the steps of the schedule (I picture included)",
            "peak_source": peaks["src"] + " (of measured)" if peaks["src"] == "measured" else "fallback",
            "share_of_step": round(per[dom]["ms_per_step"] / total_k, 3),
            "per_kernel": {k: {"ms": round(v["ms_per_step"], 4), "GBps": round(v["bytes_per_step"] / (v["ms_per_step"] * 1e-3) / 1e9, 1) if v["ms_per_step"] > 0 else None} for k, v in per.items()}}
    fps = world * args.steps / (ms_dev * 1e-3)
    fps_e2e = world * args.steps / (ms_e2e * 1e-3)
    line = {"metric": METRIC, "value": round(fps, 2), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16 samples / int32 accumulate", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": round(fps_e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": int(h2d_step), "d2h_bytes_per_step": d2h_step, "diag": e2e_diag,
                    "api": "b200_pic_upload (one picture ahead) + b200_pic_run + b200_get_frame_async (D2H overlapped with the next picture), pinned host buffers; every step uploads one picture's work lists and downloads one frame"},
            "seam": seam, "picture_ms": split, "gather": None, "lanes_per_gpu": max(1, args.lanes),
            "gpu_launches": int(launches), "numa": numa, "clocks": sampler.summary(), "roofline": roof}
    if not args.no_cpu_baseline and world == 1:                      # the CPU legs are N = 1 lines only
        try: line["cpu_baseline"] = cpu_baseline(args, wl)
        except Exception as e: line["cpu_baseline"] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    return line


def pcie_ceiling(torch, mb=256, reps=4):
    """Measured host<->device link: one large pinned buffer each way, alone and both directions at once (GB/s)."""
    try:
        n = mb << 20
        h1 = torch.empty(n, dtype=torch.uint8, pin_memory=True); h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
        d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def timed(fn):
            torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); return time.perf_counter() - t
        def up():
            with torch.cuda.stream(s1):
                for _ in range(reps): d1.copy_(h1, non_blocking=True)
        def down():
            with torch.cuda.stream(s2):
                for _ in range(reps): h2.copy_(d2, non_blocking=True)
        up(); down()
        tu, td = timed(up), timed(down)
        tb = timed(lambda: (up(), down()))
        gb = reps * n / 1e9
        return {"h2d_gbs": round(gb / tu, 1), "d2h_gbs": round(gb / td, 1), "h2d_gbs_duplex": round(gb / tb, 1), "d2h_gbs_duplex": round(gb / tb, 1), "buffer_mb": mb}
    except Exception as e:
        return {"error": str(e)[:100]}


# ------------------------------------------------------------------------------------------------ CPU legs (the reference's own DecLibRecon)
def leave():
    """The line is out: skip interpreter tear-down (the reference library's static thread pools and recon objects are destroyed in an order of the loader's choosing)."""
    sys.stdout.flush(); sys.stderr.flush()
    os._exit(0)


def thread_candidates(args):
    T = threads_all(args)
    if args.host_threads: return [T]
    return sorted({t for t in (16, 32, 64, T) if t <= T} | {T})


def run_one(case, backend, T):
    secs = case.run_stock(threads=T)[2] if backend == 0 else case.run_b200(threads=T)[2]
    assert secs >= 0, "the back end failed on a workload picture"
    return secs


def backend_fps(wl, cases, backend, T, D):
    """Pictures per second of `cases` through one back end behind the seam (0: the reference's DecLibRecon, 1: DecLibReconB200) on ThreadPool(T) with D recon instances
    taking pictures in turn."""
    if D == 1: return len(cases) / sum(run_one(c, backend, T) for c in cases)
    secs, _ = wl.helpers.seam_pipelined(wl.ref, cases, T, backend, D, read=False)
    assert secs > 0, "the back end failed on a workload picture"
    return len(cases) / secs


def backend_sweep(args, wl, backend, n=6):
    """Neither back end is fastest on every host thread: the pool's task scan contends (profiles/r02_seam_threads.json: stock DecLibRecon, 4K B pictures, 128-thread box:
    32 threads x 2 alternating instances 172 frames/s, 128 x 1 89, 128 x 2 34), and an I picture wants more threads than a B picture.  For every (threads, recon
    instances) pair: n B pictures and one I picture, combined in the schedule's proportion.  Returns {"TxD": estimated frames/s of the schedule}."""
    nI = sum(1 for i in range(args.steps) if wl.sched(i)[0] == "I"); nB = args.steps - nI
    cases = [wl.B[i % len(wl.B)] for i in range(n)]
    table = {}
    for T in thread_candidates(args):
        run_one(cases[0], backend, T)                            # pool start-up
        tI = run_one(wl.I, backend, T) if nI else 0.0
        for D in sorted({1, max(1, args.recon_depth)}):
            if D > 1: wl.helpers.seam_pipelined(wl.ref, cases[:D], T, backend, D, read=False)
            fB = backend_fps(wl, cases, backend, T, D)
            table[f"{T}x{D}"] = round(args.steps / (nI * tI + nB / fB), 1)
    return table


def backend_schedule(args, wl, backend, table):
    """The schedule's pictures in the two configurations the sweep rates fastest (two instances on many threads are bistable: the same pair can run 3x slower in the
    next call); returns (frames/s, threads, instances, {config: frames/s})."""
    runs = {}
    for key in sorted(table, key=lambda k: -table[k])[:2]:
        T, D = parse_cfg(key)
        if args.warmup: backend_fps(wl, [wl.sched(i + 1)[1] for i in range(min(args.warmup, 4))], backend, T, D)
        runs[key] = backend_fps(wl, [wl.sched(i)[1] for i in range(args.steps)], backend, T, D)
    best = max(runs, key=lambda k: runs[k]); T, D = parse_cfg(best)
    return runs[best], T, D, {k: round(v, 1) for k, v in runs.items()}


def parse_cfg(key):
    t, d = key.split("x"); return int(t), int(d)


def cpu_baseline(args, wl):
    """The reference's DecLibRecon + ThreadPool on the schedule's pictures, in its fastest (threads, recon instances) configuration."""
    try: os.sched_setaffinity(0, range(os.cpu_count()))          # the GPU arm binds itself to the GPU's NUMA node; the CPU baseline gets every core
    except Exception: pass
    table = backend_sweep(args, wl, 0)
    fps, T, D, runs = backend_schedule(args, wl, 0, table)
    return {"value": round(fps, 3), "unit": "frames/s", "cores": T, "kind": "reference", "recon_instances": D, "sweep_estimate_fps": table, "schedule_fps": runs,
            "sample": f"the {args.steps} pictures of the schedule through the reference's DecLibRecon (decompressPicture..waitForPrevDecompressedPic, {wl.ref.ref_simd_level().decode()}) in the faster of "
                      f"the two best configurations of a sweep over host threads x recon instances taking pictures in turn (DecLib.h:70; 6 B pictures + the I picture each): ThreadPool({T}) x {D}"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0: return
    wl = Workload(args, 0)
    T = threads_all(args)
    table = backend_sweep(args, wl, 0)
    fps, T, D, runs = backend_schedule(args, wl, 0, table)
    cb = {"value": round(fps, 3), "unit": "frames/s", "cores": T, "kind": "reference", "recon_instances": D, "sweep_estimate_fps": table, "schedule_fps": runs,
          "sample": f"the {args.steps} pictures of the schedule through the reference's DecLibRecon (decompressPicture..waitForPrevDecompressedPic, {wl.ref.ref_simd_level().decode()}) in the faster of "
                    f"the two best configurations of a sweep over host threads x recon instances taking pictures in turn (DecLib.h:70; 6 B pictures + the I picture each): ThreadPool({T}) x {D}"}
    line = {"impl": "reference", "metric": METRIC, "value": round(fps, 3), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / fps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16 samples / int32 accumulate", "data": "synthetic",
            "config": workload_config(args, args.gpus),
            "cpu_baseline": cb, "e2e": {"value": round(fps, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    leave()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
