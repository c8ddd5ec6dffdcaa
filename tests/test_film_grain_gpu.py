"""Film grain on the device (b200_get_frame_grain_async) against the oracle (pinned to the reference's FilmGrain by
tests/test_film_grain_oracle_vs_ref.py) and against a golden fixture produced by the reference itself (tools/make_golden.py)."""
import ctypes as C
import os
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "film_grain_fgc.npz")


def make_fg(pattern, sLUT, pLUT, seeds, shift, present):
    fg = abi.FilmGrain()
    fg.pattern = pattern.ctypes.data; fg.sLUT = sLUT.ctypes.data; fg.pLUT = pLUT.ctypes.data; fg.lineSeeds = seeds.ctypes.data
    fg.scaleShift = shift
    for c in range(3): fg.compPresent[c] = int(present[c])
    return fg


def run_device(b200, g, planes, fg, fmt=0):
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 2, 1, -1))
    try:
        vvdec_b200.check(b200.b200_ctx_load_slot(ctx, 1, abi.plane_ptrs(planes)))
        outs = [np.zeros(b200.b200_frame_bytes(C.byref(g), fmt, c), np.uint8) for c in range(3)]
        ptrs = (C.c_void_p * 3)(*[o.ctypes.data for o in outs])
        t = b200.b200_get_frame_grain_async(ctx, 1, fmt, ptrs, C.byref(fg)); assert t >= 0, b200.b200_last_error()
        vvdec_b200.check(b200.b200_frame_wait(ctx, t))
        back = [np.zeros_like(p) for p in planes]                     # the DPB picture itself must be untouched
        vvdec_b200.check(b200.b200_get_frame(ctx, 1, abi.plane_ptrs(back)))
        for a, b in zip(back, planes): assert np.array_equal(a, b)
        return outs
    finally:
        b200.b200_ctx_destroy(ctx)


@pytest.mark.parametrize("W,H,bd,present", [(416, 240, 10, (1, 1, 1)), (200, 136, 10, (1, 0, 1)), (1920, 1080, 10, (1, 1, 1)), (416, 240, 8, (1, 1, 1)),
                                            (3840, 2160, 10, (1, 1, 1)), (264, 144, 8, (0, 1, 1))])
def test_film_grain_vs_oracle(b200, oracle, W, H, bd, present):
    rng = np.random.default_rng(W + bd)
    g = abi.make_geom(W, H, bd)
    planes = synth.noise_planes(rng, W, H, bd)
    planes[0][:8, :16] = (1 << bd) - 1; planes[0][8:16, :16] = 0
    pattern, sLUT, pLUT, seeds = synth.gen_film_grain_tables(rng, H)
    shift = int(rng.integers(8, 14)) - (bd - 8)
    pres = np.array(present, np.uint8)
    fg = make_fg(pattern, sLUT, pLUT, seeds, shift, pres)
    want = [p.copy() for p in planes]
    strides = (C.c_ssize_t * 3)(*[p.shape[1] for p in want])
    oracle.orc_film_grain(abi.plane_ptrs(want), strides, W, H, bd, pattern.ctypes.data, sLUT.ctypes.data, pLUT.ctypes.data, seeds.ctypes.data, shift, pres.ctypes.data)
    outs = run_device(b200, g, planes, fg)
    for c in range(3):
        cw, ch = (W, H) if c == 0 else (W // 2, H // 2)
        got = outs[c].view(np.int16).reshape(want[c].shape)
        assert np.array_equal(got[:ch, :cw], want[c][:ch, :cw]), c
    if bd == 10 and W % 8 == 0:                                       # grain, then the application's packed format
        outs = run_device(b200, g, planes, fg, fmt=1)
        for c in range(3):
            cw, ch = (W, H) if c == 0 else (W // 2, H // 2)
            packed = np.zeros(len(outs[c]), np.uint8)
            oracle.orc_pack_pyuv(want[c], want[c].shape[1], cw, ch, packed)
            assert np.array_equal(outs[c], packed), c


def test_film_grain_golden(b200):
    """Tables synthesised by the reference's own firmware (FilmGrain::updateFGC) from an SEI, output of its SIMD line kernels."""
    z = np.load(GOLD)
    W, H, bd = [int(v) for v in z["geom"]]
    g = abi.make_geom(W, H, bd)
    src = [np.ascontiguousarray(z[f"src{c}"]) for c in range(3)]
    tabs = [np.ascontiguousarray(z[k]) for k in ("pattern", "sLUT", "pLUT", "seeds")]          # must outlive the call: the struct only holds pointers
    fg = make_fg(tabs[0], tabs[1], tabs[2], tabs[3], int(z["shift"]), z["present"])
    outs = run_device(b200, g, src, fg)
    for c in range(3):
        cw, ch = (W, H) if c == 0 else (W // 2, H // 2)
        assert np.array_equal(outs[c].view(np.int16).reshape(src[c].shape)[:ch, :cw], z[f"out{c}"][:ch, :cw]), c


def test_film_grain_argument_checks(b200):
    rng = np.random.default_rng(3)
    W, H, bd = 416, 240, 10
    g = abi.make_geom(W, H, bd)
    planes = synth.noise_planes(rng, W, H, bd)
    pattern, sLUT, pLUT, seeds = synth.gen_film_grain_tables(rng, H)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 2, 1, -1))
    try:
        vvdec_b200.check(b200.b200_ctx_load_slot(ctx, 1, abi.plane_ptrs(planes)))
        outs = [np.zeros(b200.b200_frame_bytes(C.byref(g), 0, c), np.uint8) for c in range(3)]
        ptrs = (C.c_void_p * 3)(*[o.ctypes.data for o in outs])
        bad = pLUT.copy(); bad[1, 7] = 9 << 4
        assert b200.b200_get_frame_grain_async(ctx, 1, 0, ptrs, C.byref(make_fg(pattern, sLUT, bad, seeds, 9, (1, 1, 1)))) == -2 and b"pLUT" in b200.b200_last_error()
        assert b200.b200_get_frame_grain_async(ctx, 1, 0, ptrs, C.byref(make_fg(pattern, sLUT, pLUT, seeds, 3, (1, 1, 1)))) == -2 and b"scaleShift" in b200.b200_last_error()
    finally:
        b200.b200_ctx_destroy(ctx)
