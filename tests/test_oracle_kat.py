"""Pins the CPU oracle against every value-pinning test / published vector available for this path
(SURVEY 8c): the reference's two DelayModes unit tests, Rust std's Duration::try_from_secs_f64 examples,
Random123's Philox known answers, and the hand-derived KAT-0 trace of discrete_queue.rs."""
import ctypes as C
import json
import os
import struct

import numpy as np

from oracle import oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return json.load(open(os.path.join(GOLD, name)))


def _f64(bits_hex):
    return struct.unpack("<d", struct.pack("<Q", int(bits_hex, 16)))[0]


def test_from_secs_f64_rust_doc_examples():
    L = O.lib()
    g = _load("from_secs_f64.json")
    for e in g["ok"]:
        ns = C.c_uint64()
        assert L.orc_from_secs_f64(_f64(e["bits"]), C.byref(ns)) == 0, e
        assert ns.value == e["ns"], (e, ns.value)
    for b in g["error_bits"]:
        ns = C.c_uint64()
        assert L.orc_from_secs_f64(_f64(b), C.byref(ns)) == 2, b  # ORC_FAULT_BAD_DURATION


def test_from_secs_f64_is_round_half_even_of_the_exact_value():
    """property: equals round-half-even(exact binary value * 1e9), checked with Python's exact Fractions."""
    from fractions import Fraction

    L = O.lib()
    rng = np.random.default_rng(0)
    xs = list(rng.uniform(0, 20, 2000)) + list(rng.uniform(0, 1e-6, 500)) + list(10 ** rng.uniform(-12, 9, 2000))
    xs += [k * 2.0**-31 for k in range(1, 40)] + [0.5e-9, 1.5e-9, 2.5e-9, 1e-9, 123456.0000000005]
    for x in xs:
        exact = Fraction(float(x)) * 10**9
        fl = exact.numerator // exact.denominator
        rem = exact - fl
        if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and fl % 2 == 1):
            fl += 1
        if Fraction(float(x)) < Fraction(1, 2**31) and float(x) < 2.0**-31:
            pass  # both give 0: exact*1e9 < 0.47
        ns = C.c_uint64()
        assert L.orc_from_secs_f64(float(x), C.byref(ns)) == 0
        assert ns.value == fl, (x, ns.value, fl)


def test_philox4x32_10_random123_known_answers():
    L = O.lib()
    for v in _load("philox4x32_10.json"):
        ctr = (C.c_uint32 * 4)(*[int(x, 16) for x in v["ctr"]])
        key = (C.c_uint32 * 2)(*[int(x, 16) for x in v["key"]])
        out = (C.c_uint32 * 4)()
        L.orc_philox4x32_10(ctr, key, out)
        assert ["%08x" % x for x in out] == v["out"]


def test_det_ln_accuracy():
    L = O.lib()
    xs = np.concatenate([np.random.default_rng(1).uniform(1e-16, 1.0, 5000), [1.0, 0.5, 2.0**-53, 0.7071067811865476]])
    for x in xs:
        assert abs(L.orc_ln(float(x)) - np.log(x)) <= 4e-16 * max(1.0, abs(np.log(x)))


def test_reference_delay_modes_unit_tests():
    """quokkasim/src/delays.rs:182-242, both #[test]s, replayed against the oracle's DelayModes code."""
    L = O.lib()
    for name, t in _load("delay_modes.json").items():
        h = L.orc_dm_new()
        names = []
        for nm, ud, uf in t["modes"]:
            L.orc_dm_add(h, ud, uf)
            names.append(nm)
        for step in t["steps"]:
            if step[0] == "next":
                d = C.c_uint64()
                assert L.orc_dm_next_event(h, C.byref(d)) >= 0
                assert d.value == step[1] * 10**9, (name, step, d.value)
            else:
                a, b = C.c_int(), C.c_int()
                L.orc_dm_update_state(h, step[1] * 10**9, C.byref(a), C.byref(b))
                frm = None if a.value < 0 else names[a.value]
                to = None if b.value < 0 else names[b.value]
                assert (frm, to) == (step[2], step[3]), (name, step, frm, to)
        L.orc_dm_free(h)


def _discrete_queue_oracle():
    m = O.OracleModel()
    src = m.add_component(O.DISCRETE_SOURCE)
    m.set_dist(src, 0, O.make_dist(O.DIST_CONSTANT, p=[3.0]))
    q1 = m.add_component(O.DISCRETE_STOCK, 0, 10)
    p1 = m.add_component(O.DISCRETE_PROCESS)
    d1 = m.set_dist(p1, 0, O.make_dist(O.DIST_TRIANGULAR, 0, [1, 10, 6]))
    q2 = m.add_component(O.DISCRETE_STOCK, 0, 10)
    p2 = m.add_component(O.DISCRETE_PROCESS)
    d2 = m.set_dist(p2, 0, O.make_dist(O.DIST_TRIANGULAR, 1, [1, 10, 6]))
    q3 = m.add_component(O.DISCRETE_STOCK, 0, 10)
    snk = m.add_component(O.DISCRETE_SINK)
    m.set_dist(snk, 0, O.make_dist(O.DIST_CONSTANT, p=[1.0]))
    for a, b in [(src, q1), (q1, p1), (p1, q2), (q2, p2), (p2, q3), (q3, snk)]:
        m.connect(a, b)
    return m, d1, d2


def test_kat0_discrete_queue_trace():
    """INIT + E1..E8 of discrete_queue.rs under tape P1=[4,5,..], P2=[2.5,..] -> 26 records."""
    g = _load("kat0_discrete_queue.json")
    m, d1, d2 = _discrete_queue_oracle()
    s = O.OracleSim(m)
    s.set_tape(d1, g["tapes"]["Process1"])
    s.set_tape(d2, g["tapes"]["Process2"])
    assert s.init() == 0
    s.step_events(g["n_events"])
    log = s.log()
    got = [[int(r["time_ns"]), int(r["comp"]), int(r["kind"]), int(r["reason"]), int(r["seq"]), int(r["src_comp"]),
            int(r["src_seq"]), int(r["payload"])] for r in log]
    assert got == g["records"]
    assert s.n_events == 8 and s.now == 7 * 10**9 + 2


def test_discrete_queue_reference_horizon_invariants():
    """discrete_queue.rs exactly (native streams, 200 s): conservation and capacity invariants of the oracle."""
    m, _, _ = _discrete_queue_oracle()
    s = O.OracleSim(m, base_seed=1234, rep=0)
    s.init()
    s.step_until(200 * 10**9)
    assert s.status == 0 and s.now == 200 * 10**9
    st = [s.comp_stats(i) for i in range(7)]
    created = int(st[0]["n_a"])
    in_system = sum(len(s.stock_items(i)) for i in (1, 3, 5)) + sum(int(st[i]["n_a"] - st[i]["n_b"]) for i in (0, 2, 4, 6))
    assert created == in_system + int(st[6]["n_b"])
    # at most one item every 3 s; Process1 (mean 5.67 s) is the bottleneck, so Queue1 fills and blocks the source
    assert 30 <= created <= 67
    for i in (1, 3, 5):
        assert int(st[i]["t_b_ns"]) <= 10
    t = s.log()["time_ns"]
    assert (np.diff(t.astype(np.int64)) >= 0).all()


def test_tape_exhaustion_and_zero_duration_faults():
    m, d1, d2 = _discrete_queue_oracle()
    s = O.OracleSim(m)
    s.set_tape(d1, [4.0])
    s.set_tape(d2, [2.0])
    s.init()
    s.step_until(100 * 10**9)
    assert s.status == 3
    s = O.OracleSim(m)
    s.set_tape(d1, [0.0, 1.0])
    s.set_tape(d2, [2.0])
    s.init()
    s.step_until(100 * 10**9)
    assert s.status == 1  # "Time until next event is zero!"
