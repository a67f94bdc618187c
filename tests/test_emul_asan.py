"""The kernel source compiled for the host under AddressSanitizer (tests/emul): out-of-range state-slot indices, which the
plain host build reads through silently and which compute-sanitizer is not available to find on the GPU box.  (This is how
the BasicEnvironment row read as a process row by qk_needs_refill_now was found: a wild shared-memory index on the GPU.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")
CSRC = os.path.join(ROOT, "quokkasim_b200", "csrc")


def _libasan():
    try:
        p = subprocess.run(["g++", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    except OSError:
        return None
    return p if os.path.isabs(p) and os.path.exists(p) else None


@pytest.mark.skipif(_libasan() is None, reason="g++ has no libasan here")
def test_kernel_source_under_address_sanitizer(tmp_path):
    lib = str(tmp_path / "libqk_emul_asan.so")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address", "-std=c++17", "-fPIC", "-ffp-contract=off",
                           "-Wno-unknown-pragmas", "-pthread", "-x", "c++", "-I" + EMUL, "-I" + CSRC, "-shared", "-o", lib,
                           os.path.join(EMUL, "qk_emul.cpp"), os.path.join(CSRC, "qk_model.cpp")])
    env = dict(os.environ, LD_PRELOAD=_libasan(), ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(EMUL, "asan_sweep.py"), lib], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1200)
    assert r.returncode == 0 and "asan sweep ok" in r.stdout, r.stdout[-3000:]
