"""Filter flatteners (vvdec_b200/vvdec_glue/flatten_filters.h): the reference's CtuData / APS / LoopFilterParam structures are filled from a
synthetic picture (the same way the filter shims feed the real LoopFilter / SAO / ALF), then flattened again by the glue — the result must be
the input, with the SAO availability coming from the real deriveLoopFilterBoundaryAvailibility."""
import ctypes as C
import numpy as np
import pytest
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("W,H,ctu", [(416, 240, 128), (384, 256, 64), (200, 136, 32)])
def test_filter_flatteners_round_trip(ref, W, H, ctu):
    rng = np.random.default_rng(W)
    bd = 10
    g = abi.make_geom(W, H, bd, ctu=ctu)
    pic = synth.gen_picture(rng, W, H, bd, ctu=ctu)
    lfV, lfH, sao, alf = pic["lfV"], pic["lfH"], pic["sao"], pic["alf"]
    T = pic["alfTabs"]
    lfVo, lfHo = np.zeros_like(lfV), np.zeros_like(lfH)
    saoO, alfO = np.zeros_like(sao), np.zeros_like(alf["ctus"])
    outs = [np.zeros_like(alf[k]) for k in ("lumaCoeff", "lumaClip", "chromaCoeff", "chromaClip")] + [np.zeros_like(alf["cc"][0]), np.zeros_like(alf["cc"][1])]
    counts = (C.c_int32 * 4)()
    V = C.c_void_p
    ref.ref_flatten_filters.argtypes = [C.POINTER(abi.Geom)] + [V] * 4 + [C.POINTER(abi.AlfTables)] + [V] * 10 + [C.POINTER(C.c_int32)]
    rc = ref.ref_flatten_filters(C.byref(g), lfV.ctypes.data, lfH.ctypes.data, sao.ctypes.data, alf["ctus"].ctypes.data, C.byref(T),
                                 lfVo.ctypes.data, lfHo.ctypes.data, saoO.ctypes.data, alfO.ctypes.data, *[o.ctypes.data for o in outs], counts)
    assert rc == 0
    assert np.array_equal(lfV.view(np.uint8), lfVo.view(np.uint8)) and np.array_equal(lfH.view(np.uint8), lfHo.view(np.uint8))
    assert np.array_equal(alf["ctus"].view(np.uint8), alfO.view(np.uint8))
    assert list(counts) == [alf["lumaCoeff"].shape[0], alf["chromaCoeff"].shape[0], alf["cc"][0].shape[0], alf["cc"][1].shape[0]]
    for k, o in zip(("lumaCoeff", "lumaClip", "chromaCoeff", "chromaClip"), outs): assert np.array_equal(alf[k], o), k
    assert np.array_equal(alf["cc"][0], outs[4]) and np.array_equal(alf["cc"][1], outs[5])
    # SAO: type / band / the offsets the filter reads / availability (one slice, one tile: picture boundaries only)
    assert np.array_equal(sao["type"], saoO["type"]) and np.array_equal(sao["avail"], saoO["avail"])
    for c in range(3):
        bo = sao["type"][:, c] == 4; eo = sao["type"][:, c] < 4
        assert np.array_equal(sao["band"][bo, c], saoO["band"][bo, c])
        assert np.array_equal(sao["offset"][bo, c, :4], saoO["offset"][bo, c, :4]) and np.array_equal(sao["offset"][eo, c], saoO["offset"][eo, c])
