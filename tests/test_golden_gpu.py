"""The committed golden vectors (reference outputs) replayed through the CUDA kernels via the C ABI."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi
from tests import test_golden_cpu as G
from tests.helpers import ref_ptrs

pytestmark = pytest.mark.gpu


def test_k1_golden_gpu(b200):
    def fn(g, planes, recs, coefs):
        vvdec_b200.check(b200.b200_k1_residual(C.byref(g), abi.plane_ptrs(planes), recs, 1, coefs.ctypes.data, len(coefs), None, 0, 1))
    G.run_k1(fn)


def test_k2_golden_gpu(b200):
    G.run_k2(lambda g, out, refs, pus, dm: vvdec_b200.check(b200.b200_mc_predict(C.byref(g), abi.plane_ptrs(out), ref_ptrs(refs), 4, pus.ctypes.data, len(pus), dm.ctypes.data, len(dm))))


def test_k3_golden_gpu(b200):
    z, g, p, lfV, lfH, cs, sl, seq = G.k3_inputs()
    vvdec_b200.check(b200.b200_lf_deblock(C.byref(g), abi.plane_ptrs(p), lfV.ctypes.data, lfH.ctypes.data, cs.ctypes.data, sl.ctypes.data, len(sl), C.addressof(seq), 3))
    for c in range(3): assert np.array_equal(p[c], z[f"out{c}"])


def test_k4_golden_gpu(b200):
    z, g, src, sao, v = G.k4_inputs()
    out = [np.zeros_like(p) for p in src]
    vvdec_b200.check(b200.b200_sao_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(out), sao.ctypes.data, C.addressof(v)))
    for c in range(3): assert np.array_equal(out[c], z[f"out{c}"])


def test_k5_golden_gpu(b200):
    z, g, src, t, T = G.k5_inputs()
    out = [np.zeros_like(p) for p in src]
    vvdec_b200.check(b200.b200_alf_picture(C.byref(g), abi.plane_ptrs(src), abi.plane_ptrs(out), t["ctus"].ctypes.data, C.byref(T)))
    for c in range(3): assert np.array_equal(out[c], z[f"out{c}"])


def test_chain_golden_gpu(b200):
    """The stored output of the reference arm for a picture with GEO + explicit weighted prediction + LMCS, through b200_decompress_picture."""
    z, g, refs, pic = G.chain_inputs()
    ctx = C.c_void_p()
    vvdec_b200.check(b200.b200_ctx_create(C.byref(ctx), C.byref(g), 5, 1, -1))
    try:
        for s in range(4): vvdec_b200.check(b200.b200_ctx_load_slot(ctx, s, abi.plane_ptrs(refs[s])))
        pic["struct"].dstSlot = 4
        h = b200.b200_decompress_picture(ctx, C.byref(pic["struct"])); assert h >= 0, b200.b200_last_error()
        vvdec_b200.check(b200.b200_wait_picture(ctx, h, None, 0))
        got = [np.zeros_like(np.ascontiguousarray(z[f"out{c}"])) for c in range(3)]
        vvdec_b200.check(b200.b200_get_frame(ctx, 4, abi.plane_ptrs(got)))
        for c in range(3): assert np.array_equal(got[c], z[f"out{c}"]), f"plane {c}"
    finally:
        b200.b200_ctx_destroy(ctx)
