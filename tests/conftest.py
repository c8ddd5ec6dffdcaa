import os, sys, ctypes as C
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/libvvdec_ref.so (built from /root/reference)")


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


@pytest.fixture(scope="session")
def oracle():
    from tests import helpers
    return helpers.load_oracle()


@pytest.fixture(scope="session")
def ref():
    from tests import helpers
    lib = helpers.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libvvdec_ref.so not built (needs /root/reference; run `make -C oracle ref`)")
    return lib


@pytest.fixture(scope="session")
def b200():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vvdec_b200
    return vvdec_b200.lib()


@pytest.hookimpl(trylast=True)
def pytest_unconfigure(config):
    """The compiled reference keeps thread pools and recon objects in function-level statics (oracle/ref_seam.h); their destruction order at interpreter exit is
    the loader's and has crashed after an otherwise green run.  Once the session's reporting is done, leave without running static destructors."""
    from tests import helpers
    if getattr(helpers, "_ref_lib", None) is not None or getattr(helpers, "_REF", None) is not None:
        status = getattr(config, "_b200_exitstatus", None)
        if status is not None:
            sys.stdout.flush(); sys.stderr.flush()
            os._exit(int(status))


def pytest_sessionfinish(session, exitstatus):
    session.config._b200_exitstatus = exitstatus
