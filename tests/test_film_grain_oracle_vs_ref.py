"""Film grain synthesis: the oracle's per-sample restatement against the real FilmGrain — updateFGC firmware, SIMD line kernels,
seed state carried over frames — driven like VVDecImpl::xAddGrain; the tables reach the oracle through the glue flattener (flatten_output.h)."""
import ctypes as C
import numpy as np
import pytest
from vvdec_b200 import abi, synth

pytestmark = pytest.mark.ref


def run_ref(ref, sei, bd, w, h, frames, planes, scalar=0):
    pattern = np.zeros((2, 8, 64, 64), np.int8); sLUT = np.zeros((3, 256), np.uint8); pLUT = np.zeros((3, 256), np.uint8)
    seeds = np.zeros((h + 15) // 16, np.uint32); shift = C.c_int(0); present = np.zeros(3, np.uint8)
    strides = (C.c_ssize_t * 3)(*[p.shape[1] for p in planes])
    assert ref.ref_film_grain(sei.ctypes.data, scalar, bd, w, h, frames, abi.plane_ptrs(planes), strides, pattern.ctypes.data, sLUT.ctypes.data, pLUT.ctypes.data,
                              seeds.ctypes.data, C.byref(shift), present.ctypes.data) == 0
    return pattern, sLUT, pLUT, seeds, shift.value, present


@pytest.mark.parametrize("scalar", [0, 1])
@pytest.mark.parametrize("w,h,bd,model,present,frames", [
    (416, 240, 10, 0, (1, 1, 1), 1), (416, 240, 10, 1, (1, 1, 1), 2), (416, 240, 8, 0, (1, 1, 1), 1), (200, 136, 10, 0, (1, 0, 1), 3),
    (1920, 1080, 10, 0, (1, 1, 1), 1), (136, 72, 8, 1, (0, 1, 1), 2), (264, 144, 10, 1, (1, 1, 0), 1)])
def test_film_grain(oracle, ref, w, h, bd, model, present, frames, scalar):
    """scalar=1: the C model (FilmGrainImpl), any scale.  scalar=0: the SIMD class the decoder runs on x86; at 10 bit it sign-extends the uint8
    scale LUT entries (FilmGrainImpl_X86_SIMD.h:450,:478), so the two only agree below 128 — the range the oracle / device are pinned to there."""
    rng = np.random.default_rng(w + h + bd + model)
    sei = synth.gen_fgc_sei(rng, model, present, max_scale=256 if (scalar or bd == 8) else 128)
    src = synth.noise_planes(rng, w, h, bd)
    src[0][:8, :16] = (1 << bd) - 1; src[0][8:16, :16] = 0                       # clipping at both ends
    want = [p.copy() for p in src]
    pattern, sLUT, pLUT, seeds, shift, pres = run_ref(ref, sei, bd, w, h, frames, want, scalar)
    assert list(pres) == list(present)
    got = [p.copy() for p in src]
    strides = (C.c_ssize_t * 3)(*[p.shape[1] for p in got])
    oracle.orc_film_grain(abi.plane_ptrs(got), strides, w, h, bd, pattern.ctypes.data, sLUT.ctypes.data, pLUT.ctypes.data, seeds.ctypes.data, shift, pres.ctypes.data)
    changed = 0
    for c in range(3):
        cw, ch = (w, h) if c == 0 else (w // 2, h // 2)
        assert np.array_equal(got[c][:ch, :cw], want[c][:ch, :cw]), c
        changed += int(np.count_nonzero(want[c][:ch, :cw] != src[c][:ch, :cw]))
    assert changed > 500                                                         # the SEI really produced grain
