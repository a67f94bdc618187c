import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMUL_DIR = os.path.join(ROOT, "tests", "emul")
EMUL_LIB = os.path.join(EMUL_DIR, "libqk_emul.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_emul():
    srcs = [os.path.join(EMUL_DIR, f) for f in ("qk_emul.cpp", "qk_host_emul.h")]
    csrc = os.path.join(ROOT, "quokkasim_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(EMUL_LIB) or not os.path.exists(EMUL_LIB.replace("emul.so", "emul_eager.so")) or any(os.path.getmtime(s) > os.path.getmtime(EMUL_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", EMUL_DIR, "-s", "-B", "libqk_emul.so", "libqk_emul_eager.so"])
    return EMUL_LIB


@pytest.fixture(scope="session")
def emul_lib():
    """Host build of the real kernel source (tests/emul): checks kernel LOGIC on CPU. Test-only."""
    return _build_emul()


@pytest.fixture(scope="session")
def emul_eager_lib(emul_lib):
    """Same host build with the REFILL phase taken as soon as one lane waits (QK_REFILL_LANES=1)."""
    return os.path.join(EMUL_DIR, "libqk_emul_eager.so")


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library; GPU tests call through its C ABI."""
    from quokkasim_b200 import _capi

    return _capi.DEFAULT_LIB


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
