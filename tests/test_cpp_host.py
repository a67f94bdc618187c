"""The C++ host mirror (include/quokka_b200.hpp) and its discrete_queue example (examples/discrete_queue.cpp, the
reference's quokkasim_examples/src/bin/discrete_queue.rs restated): compiled with g++ and run through the C ABI --
against the CPU build of the kernel source here, against the product library on the GPU."""
import os
import re
import subprocess

import pytest

import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _build(tmp_path, libdir, libname):
    exe = str(tmp_path / "discrete_queue")
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "discrete_queue.cpp"), "-o", exe, "-L" + libdir,
                           "-l:" + libname, "-Wl,-rpath," + libdir])
    return exe


def _events(out):
    return int(re.search(r"events (\d+)", out).group(1))


def test_cpp_example_reproduces_kat0(emul_lib, tmp_path):
    exe = _build(tmp_path, os.path.dirname(emul_lib), os.path.basename(emul_lib))
    out = subprocess.check_output([exe, "--replications", "3", "--out", str(tmp_path), "--kat0"], text=True)
    assert "faulted 0" in out and _events(out) == 3 * 8
    assert open(tmp_path / "QueueLogger.csv").read() == open(os.path.join(GOLDEN, "kat0_StockLogger.csv")).read()
    assert open(tmp_path / "ProcessLogger.csv").read() == open(os.path.join(GOLDEN, "kat0_ProcessLogger.csv")).read()


def test_cpp_and_python_hosts_build_the_same_model(emul_lib, tmp_path):
    """same model rows, same seeds -> the same number of executed actions over the reference's 200 s horizon"""
    exe = _build(tmp_path, os.path.dirname(emul_lib), os.path.basename(emul_lib))
    out = subprocess.check_output([exe, "--replications", "5"], text=True)
    sim = H.discrete_queue().sim(emul_lib, 5, base_seed=1234)
    sim.step_until(200 * 10**9)
    assert _events(out) == sim.summary()["n_events"]
    sim.close()


@pytest.mark.gpu
def test_cpp_example_on_gpu(gpu_lib, tmp_path):
    exe = _build(tmp_path, os.path.dirname(gpu_lib), os.path.basename(gpu_lib))
    out = subprocess.check_output([exe, "--replications", "65536"], text=True)
    sim = H.discrete_queue().sim(gpu_lib, 65536, base_seed=1234)
    sim.step_until(200 * 10**9)
    assert "faulted 0" in out and _events(out) == sim.summary()["n_events"]
    sim.close()
    out = subprocess.check_output([exe, "--replications", "2", "--out", str(tmp_path), "--kat0"], text=True)
    assert open(tmp_path / "QueueLogger.csv").read() == open(os.path.join(GOLDEN, "kat0_StockLogger.csv")).read()
