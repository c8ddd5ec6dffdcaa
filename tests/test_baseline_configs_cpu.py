"""BASELINE.json configs that run without a GPU.
config[0]: 'single 64x64 CTU, 1080p I-frame, DCT-2-only inverse transform + deblock on CPU (reference path, no GPU)' — the oracle chain
against the reference's own kernels on exactly that shape: CTU 64, every CU intra (predictions come in as given samples),
residuals DCT-2 only, deblocking with intra boundary strengths, no SAO / ALF."""
import ctypes as C
import numpy as np
import pytest
from vvdec_b200 import abi, synth
from tests.helpers import ref_ptrs, oracle_decompress

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("simd", [0, 1])
def test_config0_1080p_intra_dct2_deblock(oracle, ref, simd):
    W, H, bd, ctu = 1920, 1080, 10, 64
    rng = np.random.default_rng(100)
    Hp = (H + 7) // 8 * 8
    g = abi.make_geom(W, Hp, bd, ctu=ctu)                     # the coded picture is 1088 rows (conformance window crops to 1080)
    pred = synth.noise_planes(rng, W, Hp, bd)
    pic = synth.gen_picture(rng, W, Hp, bd, ctu=ctu, dst_slot=0, inter=False, given=pred, sao=False, alf=False,
                            cu_intra=True, tu_kw=dict(p_cbf=0.8, p_mts=0.0, p_ts=0.0, p_jccr=0.0, p_intra=1.0))
    assert len(pic["pus"]) == 0 and (pic["tus"]["trType"] == 0).all() and (pic["tus"]["flags"] & 7 == 0).all()
    dpb = [pred] * 4
    want, _ = oracle_decompress(oracle, g, dpb, pic)
    got = [np.zeros_like(p) for p in want]
    ref.ref_decompress_picture_out(C.byref(g), ref_ptrs(dpb), C.byref(pic["struct"]), 4, simd, abi.plane_ptrs(got))
    for c in range(3):
        assert np.array_equal(want[c], got[c]), f"plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs"
    assert not np.array_equal(want[0], pred[0])
