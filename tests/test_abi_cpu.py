"""CPU-side checks of the boundary: the library loads without a GPU and exports every symbol of include/vvdec_b200.h."""
import os, re, ctypes as C
import numpy as np
import vvdec_b200
from vvdec_b200 import abi, bindings, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vvdec_b200.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", hdr))
    assert declared, "no B200_API declarations found"
    lib = C.CDLL(vvdec_b200.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(bindings.EXPORTS), declared ^ set(bindings.EXPORTS)


def test_struct_sizes_match_header():
    assert C.sizeof(abi.Tu) == 32 and abi.TU_DTYPE.itemsize == 32
    assert C.sizeof(abi.Geom) == 32


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    lib = vvdec_b200.lib()
    g = abi.make_geom(64, 64, 10)
    planes = [np.zeros((64, 64), np.int16), np.zeros((32, 32), np.int16), np.zeros((32, 32), np.int16)]
    rc = lib.b200_k1_residual(C.byref(g), abi.plane_ptrs(planes), None, 0, None, 0, None, 0, 0)
    assert rc == -3 and b"no CPU fallback" in lib.b200_last_error()


def test_partition_tiles_picture_exactly():
    rng = np.random.default_rng(0)
    for (W, H) in [(416, 240), (1920, 1080), (136, 72)]:
        cus = synth.partition(rng, W, H)
        cover = np.zeros((H, W), np.int32)
        for x, y, w, h in cus:
            assert x + w <= W and y + h <= H
            cover[y:y + h, x:x + w] += 1
        assert (cover == 1).all()
