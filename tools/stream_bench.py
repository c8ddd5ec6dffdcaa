"""Whole-decoder throughput on a generated bitstream (not part of bench.py's contract; a tool for the GPU box):
    python tools/stream_bench.py [--size 1920x1080] [--gops 4] [--threads 16] [--repeat 3] [--cpu-only]
writes a random-access stream (hierarchical GOPs of 8, every tool of the stream tests, ALF + CC-ALF + LMCS + SAO, hash SEIs off) with oracle/vvc_stream.py, then times
vvdec_decoder_open .. vvdec_flush over it for
  stock    the unmodified reference (parser + its own DecLibRecon on the host threads)
  swapped  the same decoder with b200glue::DecLibReconB200 behind the seam (parser on the host, reconstruction on the device through the C ABI)
and checks the frames of the two against each other.  Frames per second include parsing, picture management and the copy of every frame out of the decoder."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
from oracle import vvc_stream as vs
from tests import stream_util as su
from tests.test_stream_cpu import ALL, gop8, _diff


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080"); ap.add_argument("--gops", type=int, default=4); ap.add_argument("--threads", type=int, default=16)
    ap.add_argument("--repeat", type=int, default=3); ap.add_argument("--seed", type=int, default=1); ap.add_argument("--cpu-only", action="store_true")
    a = ap.parse_args()
    W, H = (int(v) for v in a.size.split("x")); W -= W % 8; H -= H % 8
    rng = np.random.default_rng(a.seed)
    cfg = vs.Config(**dict(ALL, width=W, height=H, ctu=128, dpb_size=8, alf=True, ccalf=True, lmcs=True, level_idc=102))
    pics = vs.with_lmcs(vs.with_alf(gop8(n_gops=a.gops), rng), rng, every=4)
    t0 = time.time(); aus, drawn, nbins = vs.build_stream(cfg, pics, seed=a.seed); t_build = time.time() - t0
    n = len(aus); fs = W * H * 2
    def run(fn):
        best = None
        for _ in range(a.repeat):
            t = time.time(); frames = fn(); dt = time.time() - t
            best = dt if best is None else min(best, dt)
        return frames, n / best
    stock, fps_stock = run(lambda: vs.decode(vs.REF_SO, aus, threads=a.threads, frame_samples=fs))
    out = dict(size=f"{W}x{H}", pictures=n, stream_bytes=sum(map(len, aus)), bins=int(sum(nbins)), threads=a.threads, build_seconds=round(t_build, 1),
               stock_fps=round(fps_stock, 2), writer_matches_stock=_diff(drawn, stock) == [0] * n)
    if not a.cpu_only:
        swapped, fps_swapped = run(lambda: su.decode_swapped_device(aus, threads=a.threads, frame_samples=fs))
        out.update(swapped_fps=round(fps_swapped, 2), swapped_matches_stock=_diff(swapped, stock) == [0] * n)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
