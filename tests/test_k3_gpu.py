"""K3 parity on the GPU: CUDA deblocking (through the C ABI) vs the pinned oracle, bit-exact."""
import ctypes as C
import numpy as np
import pytest
import vvdec_b200
from vvdec_b200 import abi
from tests.test_k3_oracle_vs_ref import _picture_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,W,H,bd,ctu,nsl,ladf", [(1, 256, 128, 10, 128, 1, 0), (2, 416, 240, 10, 64, 2, 1), (3, 200, 136, 8, 32, 3, 0),
                                                     (4, 1920, 1080, 10, 128, 1, 0), (5, 384, 256, 12, 128, 1, 1),
                                                     (6, 3840, 2160, 10, 128, 1, 0)])
@pytest.mark.parametrize("dirs", [1, 2, 3])
def test_deblock_gpu_vs_oracle(b200, oracle, seed, W, H, bd, ctu, nsl, ladf, dirs):
    if W >= 1920 and dirs != 3:
        pytest.skip("large pictures: full V+H only")
    rng = np.random.default_rng(seed)
    cus, lfV, lfH, planes, sl, ctu_slice, seq = _picture_case(rng, W, H, bd, ctu, nsl, ladf)
    g = abi.make_geom(W, H, bd, ctu=ctu)
    a = [p.copy() for p in planes]; b = [p.copy() for p in planes]
    oracle.orc_lf_deblock(C.byref(g), abi.plane_ptrs(a), lfV.ctypes.data, lfH.ctypes.data, ctu_slice.ctypes.data,
                          sl.ctypes.data, C.addressof(seq), dirs)
    vvdec_b200.check(b200.b200_lf_deblock(C.byref(g), abi.plane_ptrs(b), lfV.ctypes.data, lfH.ctypes.data,
                                          ctu_slice.ctypes.data, sl.ctypes.data, nsl, C.addressof(seq), dirs))
    for c in range(3):
        assert np.array_equal(a[c], b[c]), f"plane {c}: {np.argwhere(a[c] != b[c])[:8]}"
    assert not np.array_equal(a[0], planes[0])
