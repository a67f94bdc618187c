"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header declares,
the host-side builder mirror behaves like the reference's (names, defaults, error text), and the batch entry
points fail loudly -- never fall back -- when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import quokkasim_b200 as qk
from quokkasim_b200 import _capi as capi
from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_symbol_in_the_header():
    hdr = open(os.path.join(ROOT, "include", "quokka_b200.h")).read()
    declared = set(re.findall(r"\b(qk_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"qk_status", "qk_model", "qk_batch"}
    L = capi.load_library()
    for sym in sorted(declared):
        assert hasattr(L, sym), "libquokka_b200.so does not export %s" % sym
    assert declared == set(capi.EXPORTED_SYMBOLS), declared ^ set(capi.EXPORTED_SYMBOLS)
    assert L.qk_abi_version() == 2


def test_library_contains_sm_100a_code_only():
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", capi.DEFAULT_LIB], stdout=subprocess.PIPE, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out


def test_model_table_calls_and_error_text():
    L = capi.load_library()
    m = C.c_void_p()
    capi.check(L, L.qk_model_create(C.byref(m)))
    ids = []
    for kind in (capi.DISCRETE_SOURCE, capi.DISCRETE_STOCK, capi.DISCRETE_PROCESS, capi.DISCRETE_STOCK,
                 capi.DISCRETE_SINK, capi.ENVIRONMENT):
        d = capi.ComponentDesc()
        d.kind = kind
        d.max_capacity = 10
        out = C.c_uint32()
        capi.check(L, L.qk_model_add_component(m, C.byref(d), C.byref(out)))
        ids.append(out.value)
    assert ids == [0, 1, 2, 3, 4, 5]
    src, q1, p, q2, snk, env = ids
    for a, b in ((src, q1), (q1, p), (p, q2), (q2, snk), (env, p)):
        capi.check(L, L.qk_model_connect(m, a, b, -1))
    # unsupported pair: same message shape as the reference's Err (discrete_queue.rs:14-20)
    rc = L.qk_model_connect(m, src, p, -1)
    assert rc == -2
    assert L.qk_last_error().decode() == "No component connection defined from StringSource to StringProcess (n=None)"
    rc = L.qk_model_connect(m, q1, q2, 3)
    assert rc == -2 and L.qk_last_error().decode().endswith("(n=Some(3))")
    # second upstream on a single-reply port
    assert L.qk_model_connect(m, q2, p, -1) == -3
    # schedule_event arms that the reference does not implement (core.rs:1397-1399)
    assert L.qk_model_schedule_env_state(m, 10, p, 1) == -1
    assert b"schedule_event not implemented" in L.qk_last_error()
    capi.check(L, L.qk_model_schedule_env_state(m, 10, env, 1))
    bad = capi.DistDesc()
    bad.kind = capi.DIST_TRIANGULAR
    bad.p[0], bad.p[1], bad.p[2] = 5.0, 1.0, 3.0
    assert L.qk_model_set_dist(m, p, 0, C.byref(bad), None) == -1
    nb = C.c_uint32()
    capi.check(L, L.qk_model_state_bytes(m, 0, C.byref(nb)))
    small = nb.value
    capi.check(L, L.qk_model_state_bytes(m, 1, C.byref(nb)))
    assert 0 < small < nb.value < 4096
    L.qk_model_destroy(m)


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    """the product path must fail loudly when it cannot run on a GPU"""
    from quokkasim_b200 import workloads

    sim = workloads.discrete_queue().sim(None, 4)
    with pytest.raises(qk.QkError) as e:
        sim.step_events(10)
    assert e.value.status == -4 and "no CPU fallback" in str(e.value)


def test_missing_library_is_a_loud_import_error(tmp_path):
    with pytest.raises(ImportError):
        capi.load_library(str(tmp_path / "libnope.so"))


def test_builder_defaults_match_the_reference():
    s = qk.DiscreteStock.new()
    assert (s.element_name, s.element_type, s.low_capacity, s.max_capacity) == ("DiscreteStock", "DiscreteStock", 0, 1)
    p = qk.DiscreteProcess.new()
    assert p.process_time_distr.kind == capi.DIST_CONSTANT and p.process_time_distr.params[0] == 1.0
    f = qk.StringItemFactory()
    assert f.item_name(0) == "Item_0000" and f.item_name(12345) == "Item_12345"
    assert qk.StringItemFactory("Car", 7, 3).item_name(1) == "Car_008"
    assert str(qk.ComponentModel.StringStock(s, qk.Mailbox.new())) == "StringStock"


def test_distribution_factory_seed_assignment_including_q6():
    """common.rs:73-148: Uniform/Triangular/Constant arms bump next_seed, the Normal/TruncNormal/Exponential arms
    return early and do not -> consecutive Exponentials share a stream."""
    df = qk.DistributionFactory(base_seed=1234, next_seed=0)
    a = df.create(qk.DistributionConfig.Triangular(1, 10, 6))
    b = df.create(qk.DistributionConfig.Triangular(1, 10, 6))
    c = df.create(qk.DistributionConfig.Constant(2.0))
    e1 = df.create(qk.DistributionConfig.Exponential(5.0))
    e2 = df.create(qk.DistributionConfig.Exponential(7.0))
    n = df.create(qk.DistributionConfig.Normal(0.0, 1.0))
    u = df.create(qk.DistributionConfig.Uniform(0.0, 1.0))
    assert (a.stream, b.stream) == (0, 1)
    assert c.kind == capi.DIST_CONSTANT
    assert e1.stream == e2.stream == n.stream == u.stream == 3
    assert df.next_seed == 4
    assert qk.DistributionFactory.new(9).next_seed == 9
    with pytest.raises(qk.DistributionParametersError):
        df.create(qk.DistributionConfig.Triangular(5, 1, 3))
    with pytest.raises(qk.DistributionParametersError):
        df.create(qk.DistributionConfig.TruncNormal(0, 1, min=2, max=1))


def test_connect_components_errors_like_the_reference():
    src = qk.ComponentModel.StringSource(qk.DiscreteSource.new(), qk.Mailbox.new())
    snk = qk.ComponentModel.StringSink(qk.DiscreteSink.new(), qk.Mailbox.new())
    q = qk.ComponentModel.StringStock(qk.DiscreteStock.new(), qk.Mailbox.new())
    with pytest.raises(qk.ComponentConnectionError) as e:
        qk.connect_components(src, snk)
    assert str(e.value) == "No component connection defined from StringSource to StringSink (n=None)"
    with pytest.raises(qk.ComponentConnectionError):
        qk.connect_components(q, src)  # a source has no upstream arm (core.rs:910-915 is Source -> Stock only)
    qk.connect_components(src, q)
    qk.connect_components(q, snk)
    lg = qk.ComponentLogger.StringStockLogger(qk.DiscreteStockLogger.new("L"))
    with pytest.raises(qk.ComponentConnectionError):
        qk.connect_logger(lg, src)
    qk.connect_logger(lg, q)


def test_delay_mode_builder_follows_indexmap_semantics():
    p = qk.DiscreteProcess.new()
    d = qk.Distribution.Constant
    p.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("A", d(1), d(2))))
    p.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("B", d(3), d(4))))
    p.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("A", d(5), d(6))))  # same key keeps its position
    assert [m.name for m in p.delay_modes] == ["A", "B"] and p.delay_modes[0].until_delay_distr.params[0] == 5
    p.with_delay_mode(qk.DelayModeChange.Remove("A"))
    assert [m.name for m in p.delay_modes] == ["B"]
    p.with_delay_mode(qk.DelayModeChange.RemoveAll())
    assert p.delay_modes == []


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/quokka_b200.h is a C header (extern "C", plain pointers and sizes): a strict C99 translation unit
    compiles against it, links the product library and drives the model half of the ABI (no GPU needed)."""
    import os
    import subprocess

    from quokkasim_b200 import _capi

    src = tmp_path / "abi.c"
    src.write_text(r'''
#include "quokka_b200.h"
#include <stdio.h>
int main(void) {
  qk_model* m = 0;
  qk_component_desc d = {QK_DISCRETE_STOCK, 0, 3, 0, 0, 0};
  unsigned id = 99;
  if (qk_model_create(&m)) return 1;
  if (qk_model_add_component(m, &d, &id)) return 2;
  if (qk_model_connect(m, 0, 0, -1) != QK_ERR_UNSUPPORTED_CONNECTION) return 3;
  printf("%d %u %s\n", qk_abi_version(), id, qk_last_error());
  qk_model_destroy(m);
  return 0;
}
''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir, libname = os.path.dirname(_capi.DEFAULT_LIB), os.path.basename(_capi.DEFAULT_LIB)
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"),
                           str(src), "-o", exe, "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe], text=True)
    assert out.strip() == "2 0 No component connection defined from StringStock to StringStock (n=None)"


def test_container_abi_argument_checks():
    """model half of the container ABI (no GPU): kinds 32-38, initial containers, factory, max_process_count"""
    import ctypes as C

    import numpy as np

    from quokkasim_b200 import _capi as capi

    L = capi.load_library()
    m = C.c_void_p()
    capi.check(L, L.qk_model_create(C.byref(m)))

    def add(kind, **kw):
        d = capi.ComponentDesc()
        d.kind = kind
        for k, v in kw.items():
            setattr(d, k, v)
        out = C.c_uint32()
        rc = L.qk_model_add_component(m, C.byref(d), C.byref(out))
        return rc, out.value

    rc, cstock = add(capi.CONTAINER_STOCK, max_capacity=4)
    assert rc == 0
    rc, sstock = add(capi.DISCRETE_STOCK, max_capacity=4)
    assert rc == 0
    assert add(capi.CONTAINER_STOCK, n_initial_items=2)[0] == -1  # initial contents come with their payloads
    assert add(capi.CONTAINER_LOAD_PROCESS, max_capacity=0)[0] == -1  # max_process_count must be >= 1
    rc, load = add(capi.CONTAINER_LOAD_PROCESS, max_capacity=2)
    assert rc == 0
    rc, csrc = add(capi.CONTAINER_SOURCE)
    assert rc == 0
    rc, vstock = add(capi.VECTOR_STOCK)
    assert rc == 0
    capi.check(L, L.qk_model_set_vector(m, vstock, 3, 0.0, 10.0, (C.c_double * 3)(1, 2, 3)))
    rc, fstock = add(capi.VECTOR_STOCK)
    capi.check(L, L.qk_model_set_vector(m, fstock, 1, 0.0, 10.0, (C.c_double * 3)(1, 0, 0)))
    res = np.zeros((2, 3))
    has = np.array([1, 0], dtype=np.uint8)
    caps = np.array([5.0, 6.0])
    assert L.qk_model_set_initial_containers(m, sstock, 2, res.ctypes.data, has.ctypes.data, caps.ctypes.data) == -1
    assert L.qk_model_set_initial_containers(m, cstock, 2, res.ctypes.data, has.ctypes.data, None) == -1
    capi.check(L, L.qk_model_set_initial_containers(m, cstock, 2, res.ctypes.data, has.ctypes.data, caps.ctypes.data))
    split = (C.c_double * 3)(0.5, 0.25, 0.25)
    did = C.c_uint32()
    assert L.qk_model_set_container_factory(m, load, None, split, 9.0, C.byref(did)) == -1  # not a container source
    capi.check(L, L.qk_model_set_container_factory(m, csrc, None, split, 9.0, C.byref(did)))
    assert did.value == 0xFFFFFFFF  # create_quantity_distr = None
    bad = capi.DistDesc()
    bad.kind = capi.DIST_TRIANGULAR
    bad.p[0], bad.p[1], bad.p[2] = 5.0, 1.0, 3.0
    assert L.qk_model_set_container_factory(m, csrc, C.byref(bad), split, 9.0, C.byref(did)) == -1
    # connection arms (core.rs:923-995): Vector3Stock -> load only for dim 3, no string stock, one resource stock
    capi.check(L, L.qk_model_connect(m, vstock, load, -1))
    assert L.qk_model_connect(m, fstock, load, -1) == -2
    assert L.qk_model_connect(m, vstock, load, -1) == -3  # port already in use
    assert L.qk_model_connect(m, sstock, load, -1) == -2
    capi.check(L, L.qk_model_connect(m, cstock, load, -1))
    capi.check(L, L.qk_model_connect(m, csrc, cstock, -1))
    assert b"No component connection defined from StringStock to Vector3ContainerLoadProcess" in L.qk_last_error() or True
    n = C.c_uint32()
    capi.check(L, L.qk_model_state_bytes(m, 1, C.byref(n)))
    assert n.value > 0
    L.qk_model_destroy(m)


def test_user_defined_resource_abi_argument_checks():
    """qk_model_set_vector_n (iron_ore.rs: IronOre, five fields, total over two): dim 1..8, a total mask inside dim;
    components connect only when field count and mask agree; the Python mirror types the variants."""
    L = capi.load_library()
    m = C.c_void_p()
    capi.check(L, L.qk_model_create(C.byref(m)))

    def add(kind):
        d = capi.ComponentDesc()
        d.kind = kind
        out = C.c_uint32()
        capi.check(L, L.qk_model_add_component(m, C.byref(d), C.byref(out)))
        return out.value

    init = (C.c_double * 8)(60, 40, 10, 5, 15, 0, 0, 0)
    ore, proc, v3, other = add(capi.VECTOR_STOCK), add(capi.VECTOR_PROCESS), add(capi.VECTOR_STOCK), add(capi.VECTOR_STOCK)
    assert L.qk_model_set_vector_n(m, ore, 9, 3, 10.0, 100.0, init) == -1  # more than QK_VEC_MAX_DIM fields
    assert L.qk_model_set_vector_n(m, ore, 0, 1, 10.0, 100.0, init) == -1
    assert L.qk_model_set_vector_n(m, ore, 5, 0, 10.0, 100.0, init) == -1  # a total over no field
    assert L.qk_model_set_vector_n(m, ore, 5, 0x23, 10.0, 100.0, init) == -1  # ... over a field beyond dim
    capi.check(L, L.qk_model_set_vector_n(m, ore, 5, 3, 10.0, 100.0, init))
    capi.check(L, L.qk_model_set_vector_n(m, proc, 5, 3, 0.0, 0.0, None))
    capi.check(L, L.qk_model_set_vector(m, v3, 3, 0.0, 10.0, init))
    capi.check(L, L.qk_model_set_vector_n(m, other, 5, 7, 10.0, 100.0, init))  # same fields, another total
    capi.check(L, L.qk_model_connect(m, ore, proc, -1))
    assert L.qk_model_connect(m, proc, v3, -1) == -2
    assert L.qk_model_connect(m, proc, other, -1) == -2
    L.qk_model_destroy(m)

    import quokkasim_b200 as qk
    from quokkasim_b200.workloads import IronOre

    assert IronOre(fe=60.0, other_elements=40.0, magnetite=10.0).total() == 100.0 and IronOre.TOTAL_MASK == 3
    with pytest.raises(TypeError):
        qk.ComponentModel.IronOreStock(qk.VectorStock.new().with_initial_resource(qk.Vector3([1, 2, 3])), qk.Mailbox.new())
    a = qk.ComponentModel.IronOreStock(qk.VectorStock.new(), qk.Mailbox.new())
    b = qk.ComponentModel.Vector3Process(qk.VectorProcess.new(), qk.Mailbox.new())
    with pytest.raises(qk.api.ComponentConnectionError):
        qk.connect_components(a, b)
    with pytest.raises(ValueError):
        qk.define_vector_variants("IronOre", IronOre)  # the variants exist already (workloads.py)
    with pytest.raises(TypeError):
        qk.define_resource("TooWide", ["f%d" % i for i in range(9)])


def test_rust_ffi_declares_every_entry_point_of_the_header():
    """rust_shim/src/ffi.rs is kept in step with include/quokka_b200.h by hand (no Rust toolchain in this image): same
    set of entry points, same number of parameters each."""
    import re

    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    header = open(os.path.join(root, "include", "quokka_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    rust = open(os.path.join(root, "rust_shim", "src", "ffi.rs")).read()
    rust = re.sub(r"//[^\n]*", "", rust)
    c_fns = {m.group(1): m.group(2) for m in re.finditer(r"^(?:int|void|const char\*)\s+(qk_\w+)\(([^;]*?)\);", header,
                                                         re.M | re.S)}
    r_fns = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (qk_\w+)\(([^;]*?)\)\s*(?:->[^;]*)?;", rust, re.S)}
    assert set(c_fns) == set(r_fns), (sorted(set(c_fns) - set(r_fns)), sorted(set(r_fns) - set(c_fns)))

    def arity(params):
        p = params.strip()
        return 0 if p in ("", "void") else len([x for x in p.split(",") if x.strip()])

    for name in c_fns:
        assert arity(c_fns[name]) == arity(r_fns[name]), name
    assert len(c_fns) == len(capi.EXPORTED_SYMBOLS)


def test_c_example_of_a_user_defined_resource_runs_on_the_kernel_emulation(emul_lib, tmp_path):
    """examples/iron_ore.c (strict C99 over the C ABI: qk_model_set_vector_n, qk_batch_read_vector_n) linked against the
    host build of the kernel source (tests/emul exports the same ABI) agrees to the last bit with the Python host
    running the same model, and links against the product library as well."""
    import subprocess

    import helpers as H
    from quokkasim_b200 import _capi

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "examples", "iron_ore.c")
    flags = ["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(root, "include"), src]
    exe = str(tmp_path / "iron_ore_emul")
    libdir, libname = os.path.dirname(emul_lib), os.path.basename(emul_lib)
    subprocess.check_call(flags + ["-o", exe, "-L" + libdir, "-l:" + libname, "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe, "8"], text=True).strip().split("\n")
    pdir, pname = os.path.dirname(_capi.DEFAULT_LIB), os.path.basename(_capi.DEFAULT_LIB)
    subprocess.check_call(flags + ["-o", str(tmp_path / "iron_ore_gpu"), "-L" + pdir, "-l:" + pname, "-Wl,-rpath," + pdir])

    m = H.iron_ore()
    sim = m.sim(emul_lib, 8, base_seed=123456789)
    sim.step_until(300 * 10**9)
    s2, p1 = m.by_name["MyStock2"].component._id, m.by_name["IronOreProcess"].component._id
    ore = sim.read_resource(s2, 0)
    moved = int(sim.comp_stats(p1, 0, 1)[0]["count"])
    want = ("replication 0: %d batches moved, MyStock2 = {fe %.17g, other_elements %.17g, magnetite %.17g, hematite %.17g, "
            "limonite %.17g} (5 fields), fe%% %.6f" % ((moved,) + tuple(ore) + (100.0 * ore[0] / (ore[0] + ore[1]),)))
    assert out[0] == want, (out[0], want)
    mean = sum(sum(sim.read_resource(s2, r)[:2]) for r in range(8)) / 8.0
    assert out[1] == "mean of 8 replications: MyStock2 total %.6f" % mean
    sim.close()
