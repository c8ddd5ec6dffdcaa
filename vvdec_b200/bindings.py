"""ctypes prototypes for libvvdec_b200.so (include/vvdec_b200.h)."""
import ctypes as C
import numpy as np
from . import abi

i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
PLANES = C.POINTER(C.POINTER(C.c_int16))


def declare(lib):
    lib.b200_last_error.restype = C.c_char_p
    lib.b200_version.restype = C.c_char_p
    lib.b200_device_count.restype = C.c_int
    lib.b200_k1_residual.argtypes = [C.POINTER(abi.Geom), PLANES, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.c_void_p, C.c_size_t, C.c_int]
    lib.b200_k1_residual.restype = C.c_int


    lib.b200_lf_deblock.argtypes = [C.POINTER(abi.Geom), PLANES, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                    C.c_void_p, C.c_int]
    lib.b200_lf_deblock.restype = C.c_int


    lib.b200_sao_picture.argtypes = [C.POINTER(abi.Geom), PLANES, PLANES, C.c_void_p, C.c_void_p]
    lib.b200_sao_picture.restype = C.c_int
    lib.b200_alf_picture.argtypes = [C.POINTER(abi.Geom), PLANES, PLANES, C.c_void_p, C.POINTER(abi.AlfTables)]
    lib.b200_alf_picture.restype = C.c_int


    lib.b200_mc_predict.argtypes = [C.POINTER(abi.Geom), PLANES, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    lib.b200_mc_predict.restype = C.c_int


EXPORTS = ["b200_mc_predict", "b200_last_error", "b200_version", "b200_device_count", "b200_k1_residual", "b200_lf_deblock",
           "b200_sao_picture", "b200_alf_picture"]
