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
