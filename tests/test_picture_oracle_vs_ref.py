"""Whole back end, one picture: the oracle chain (K2->K1->K3->K4->K5) vs the reference arm of bench.py
(ref_decompress_picture_out: VVdeC's own kernels, multi-threaded). Bit-exact, which also shows that the number of host threads
does not change the reference result."""
import ctypes as C
import numpy as np
import pytest
from vvdec_b200 import abi, synth
from tests.helpers import oracle_decompress, ref_ptrs

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("seed,W,H,threads,simd", [(1, 416, 240, 1, 0), (2, 416, 240, 4, 1), (3, 832, 480, 8, 1)])
def test_reference_arm_equals_oracle_chain(oracle, ref, seed, W, H, threads, simd):
    rng = np.random.default_rng(seed)
    g = abi.make_geom(W, H, 10)
    refs = [synth.noise_planes(rng, W, H, 10) for _ in range(4)]
    pic = synth.gen_picture(rng, W, H, 10, tu_kw=dict(p_lfnst=0.1, p_intra=0.3, p_bdpcm=0.05))
    want, _ = oracle_decompress(oracle, g, refs, pic)
    got = [np.zeros_like(p) for p in want]
    secs = ref.ref_decompress_picture_out(C.byref(g), ref_ptrs(refs), C.byref(pic["struct"]), threads, simd, abi.plane_ptrs(got))
    assert secs > 0
    for c in range(3):
        assert np.array_equal(want[c], got[c]), f"plane {c}: {len(np.argwhere(want[c] != got[c]))} diffs, first {np.argwhere(want[c] != got[c])[:4]}"
