"""Builds libquokka_b200.so (the C-ABI shared library with the sm_100a kernels) in-tree with nvcc.

Usage: python -m quokkasim_b200.build [--force]
The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libquokka_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["qk_model.cpp", "qk_batch.cu", "qk_multi.cu", "qk_step_l0f0.cu", "qk_step_l0f1.cu", "qk_step_l1f0.cu", "qk_step_l1f1.cu",
           "qk_step_l0f2.cu", "qk_step_l1f2.cu", "qk_step_l0f3.cu", "qk_step_l1f3.cu", "qk_group_l0f0.cu", "qk_group_l0f1.cu", "qk_group_l0f2.cu",
           "qk_group_l1f0.cu", "qk_group_l1f1.cu", "qk_group_l1f2.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # fp64 parity: no contraction anywhere in the product
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-Xptxas", "-v",
    "--threads", "0",  # the translation units compile in parallel (reproducibly, unlike --split-compile)
    "-shared", "-cudart", "static",
    "-I/usr/include",  # nccl.h (types only: libnccl.so.2 is loaded at run time by qk_multi_create)
    "-ldl",
]


_MARKER = b"QKBUILDID:"


def source_id(extra=()):
    """What the library is built FROM: sha256 over every file of csrc/ and the public header (names and contents, in
    sorted order) and over the compiler flags.  The same string is compiled into the library (qk_build_id())."""
    import hashlib

    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.startswith("."))
    files.append(os.path.join(_HERE, "..", "include", "quokka_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    h.update(" ".join(FLAGS + list(extra) + SOURCES).encode())
    return h.hexdigest()[:24]


def built_id(path=None):
    """The source id compiled into an existing library, read from the file (no dlopen), or None."""
    path = path or LIB
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    k = blob.find(_MARKER)
    if k < 0:
        return None
    end = blob.find(b"\0", k)
    return blob[k + len(_MARKER):end].decode("ascii", "replace")


def _stale(extra=()):
    """Rebuild when the library is missing or was built from other sources / flags than the tree holds now -- decided by
    content, not by file times (a library shipped with a snapshot is newer than every source file, whatever it was
    built from)."""
    return built_id() != source_id(extra)


def build(force=False, verbose=False):
    extra = os.environ.get("QK_EXTRA_NVCC_FLAGS", "").split()  # experiments: -DQK_..., with QK_LIB_OUT for the output
    out = os.environ.get("QK_LIB_OUT", LIB)
    if not force and out == LIB and not _stale(extra):
        return LIB
    cmd = [NVCC] + FLAGS + extra + ['-DQK_BUILD_ID="%s"' % source_id(extra), "-o", out] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(_HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed, see %s" % log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
