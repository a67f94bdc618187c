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
           "qk_step_l0f2.cu", "qk_step_l1f2.cu", "qk_group_l0f0.cu", "qk_group_l0f1.cu", "qk_group_l0f2.cu",
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


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(_HERE, "..", "include", "quokka_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    extra = os.environ.get("QK_EXTRA_NVCC_FLAGS", "").split()  # experiments: -DQK_..., with QK_LIB_OUT for the output
    out = os.environ.get("QK_LIB_OUT", LIB)
    cmd = [NVCC] + FLAGS + extra + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
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
