"""Device-side output conversion (b200_get_frame_fmt_async) against the oracle's restatement of vvdecapp's writers."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("W,H", [(416, 240), (24, 16), (1920, 1080)])
def test_frame_formats(b200, oracle, W, H):
    rng = np.random.default_rng(W)
    bd = 10
    g = abi.make_geom(W, H, bd)
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 2, 1, -1))
    try:
        pl = synth.noise_planes(rng, W, H, bd)
        pl[0][0, :4] = [0, 1023, 1, 1022]
        vvdec_b200.check(b200.b200_ctx_load_slot(ctx, 1, abi.plane_ptrs(pl)))
        for fmt in (1, 2, 0):
            outs = [np.zeros(b200.b200_frame_bytes(C.byref(g), fmt, c), np.uint8) for c in range(3)]
            ptrs = (C.c_void_p * 3)(*[o.ctypes.data for o in outs])
            t = b200.b200_get_frame_fmt_async(ctx, 1, fmt, ptrs); assert t >= 0, b200.b200_last_error()
            vvdec_b200.check(b200.b200_frame_wait(ctx, t))
            for c in range(3):
                w, h = (W, H) if c == 0 else (W // 2, H // 2)
                if fmt == 0: want = pl[c].view(np.uint8).reshape(-1)
                else:
                    want = np.zeros(len(outs[c]), np.uint8)
                    if fmt == 1: oracle.orc_pack_pyuv(pl[c], pl[c].shape[1], w, h, want)
                    else: oracle.orc_narrow8(pl[c], pl[c].shape[1], w, h, bd, want)
                assert np.array_equal(outs[c], want), (fmt, c)
    finally:
        b200.b200_ctx_destroy(ctx)
