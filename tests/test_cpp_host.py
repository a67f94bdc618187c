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


def _build(tmp_path, libdir, libname, example="discrete_queue"):
    exe = str(tmp_path / example)
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", example + ".cpp"), "-o", exe, "-L" + libdir,
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


def _check_container_example(lib, tmp_path, n_reps):
    """examples/container_haulage.cpp (continuous + container families through the C++ mirror) against the Python host:
    same clock, same dumped tonnage to the last bit, fleet and mass conserved"""
    exe = _build(tmp_path, os.path.dirname(lib), os.path.basename(lib), "container_haulage")
    out = subprocess.check_output([exe, "--replications", str(n_reps), "--events", "3000"], text=True)
    m = H.container_haulage()
    sim = m.sim(lib, n_reps, base_seed=42)
    sim.step_events(3000)
    s = sim.summary()
    assert "faulted 0" in out and _events(out) == s["n_events"]
    assert int(re.search(r"clock_ns (\d+)", out).group(1)) == s["max_now_ns"]
    dumped = sum(sim.read_vector(m.by_name["Dumped"].component._id, n_reps - 1))
    assert float(re.search(r"dumped (\S+)", out).group(1)) == dumped
    assert "fleet 4" in out
    sim.close()
    # CSV export of the container loggers: the C++ writer and the Python writer (which is pinned by the hand-derived
    # KAT-1 golden files) agree byte for byte on a few thousand records full of fp64 values
    from quokkasim_b200 import csvlog

    n_small = 2
    subprocess.check_output([exe, "--replications", str(n_small), "--events", "2500", "--out", str(tmp_path)], text=True)
    m = H.container_haulage()
    sim = m.sim(lib, n_small, base_seed=42, log_capacity=65536)
    sim.step_events(2500)
    csl, cpl = m.container_loggers
    assert open(tmp_path / "ContainerStockLogger.csv").read() == csvlog.csv_text(csl, sim, n_small - 1)
    assert open(tmp_path / "ContainerProcessLogger.csv").read() == csvlog.csv_text(cpl, sim, n_small - 1)
    sim.close()


def test_cpp_container_example_matches_python_host(emul_lib, tmp_path):
    _check_container_example(emul_lib, tmp_path, 3)


@pytest.mark.gpu
def test_cpp_container_example_on_gpu(gpu_lib, tmp_path):
    _check_container_example(gpu_lib, tmp_path, 4096)
