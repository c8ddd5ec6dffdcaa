"""vvdec_b200 — B200-native VVC pixel-reconstruction back end (host-side Python binding of the C ABI).

The product is the CUDA library `vvdec_b200/csrc/libvvdec_b200.so` (C ABI in include/vvdec_b200.h).
This package only loads it through ctypes; there is NO CPU fallback: if the library or a CUDA device
is missing, calls raise."""
import os, ctypes as C

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB") or os.path.join(_HERE, "csrc", "libvvdec_b200.so")      # B200_LIB: an instrumented build of the same library (tools/)
_lib = None


class B200Error(RuntimeError):
    pass


def lib():
    """Load the CUDA library (raises if it has not been built: run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} not built — the product path has no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        from . import bindings
        bindings.declare(_lib)
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().b200_last_error().decode()
        raise B200Error(f"vvdec_b200 error {rc}: {msg}")
