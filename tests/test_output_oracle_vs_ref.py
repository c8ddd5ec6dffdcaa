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


@pytest.mark.parametrize("w,h,bd", [(16, 8, 10), (416, 240, 10), (416, 240, 8), (208, 120, 12), (1920, 1080, 10), (24, 2, 10)])
def test_picture_hash(oracle, ref, w, h, bd):
    """Decoded-picture hash: the oracle's CRC / checksum per plane against calcCRC / calcChecksum (CommonLib/PicYuvMD5.cpp)."""
    import ctypes as C
    from vvdec_b200 import abi
    rng = np.random.default_rng(w + bd)
    pl = [rng.integers(0, 1 << bd, size=(hh, ww + 10)).astype(np.int16) for ww, hh in ((w, h), (w // 2, h // 2), (w // 2, h // 2))]
    pl[0][0, :4] = [0, (1 << bd) - 1, 255, 256]
    strides = (C.c_ssize_t * 3)(*[p.shape[1] for p in pl])
    for method, n in ((1, 2), (2, 4)):
        want = np.zeros(16, np.uint8)
        assert ref.ref_picture_hash(method, bd, abi.plane_ptrs(pl), strides, w, h, want, 16) == 3 * n
        for c in range(3):
            got = np.zeros(4, np.uint8)
            assert oracle.orc_plane_hash(method, bd, pl[c], pl[c].shape[1], w >> (c > 0), h >> (c > 0), got) == n
            assert np.array_equal(got[:n], want[c * n:(c + 1) * n]), (method, c)
