"""Output formats: the oracle's pyuv packing / 8-bit narrowing against vvdecapp's own plane writer (_writeComponentToFile)."""
import numpy as np
import pytest

pytestmark = pytest.mark.ref


@pytest.mark.parametrize("w,h", [(8, 2), (16, 4), (64, 36), (416, 240), (208, 120), (12, 6)])
def test_pyuv_and_8bit(oracle, ref, w, h):
    rng = np.random.default_rng(w * h)
    src = rng.integers(0, 1024, size=(h, w + 6)).astype(np.int16)
    src[0, :4] = [0, 1023, 1, 1022]
    for fmt, nbytes in ((1, w // 4 * 5 * h), (2, w * h)):
        want = np.zeros(nbytes + 8, np.uint8); got = np.zeros(nbytes, np.uint8)
        assert ref.ref_write_component(src, w + 6, w, h, fmt, want, len(want)) == nbytes
        if fmt == 1: oracle.orc_pack_pyuv(src, w + 6, w, h, got)
        else: oracle.orc_narrow8(src, w + 6, w, h, 10, got)
        assert np.array_equal(want[:nbytes], got), fmt
