"""Shared test helpers: model zoo (the reference's example topologies), oracle lowering, parity comparison.

The oracle (oracle/) is used here ONLY as the checker.
"""
import numpy as np

import quokkasim_b200 as qk
from oracle import oracle as O
from quokkasim_b200 import _capi as capi


from quokkasim_b200.workloads import (Model, assembly_line, car_workshop, discrete_queue, haulage,  # noqa: F401
                                      material_blending, scheduled_event_f64, scheduled_process_quantity,
                                      serial_line, source_sink, vector_flow, iron_ore, IronOre,
                                      transport, delay_testing, the_example, container_haulage, container_open)


from oracle.lowering import lower_to_oracle  # noqa: F401,E402
from quokkasim_b200.vector_api import VECTOR_VARIANTS as _VECTOR_VARIANTS  # noqa: E402


def make_tapes(model, n_reps, length, seed, lo=0.5, hi=9.5, dist_scale=None):
    """Pre-drawn variate tapes: one (n_reps, length) array per random Distribution instance."""
    rng = np.random.default_rng(seed)
    tapes = {}
    for d in model.random_dists:
        a, b = lo, hi
        if d.kind == capi.DIST_TRIANGULAR or d.kind == capi.DIST_UNIFORM:
            a, b = d.params[0], d.params[1]
        elif d.kind == capi.DIST_EXPONENTIAL:
            a, b = 0.2 * d.params[0], 2.0 * d.params[0]
        tapes[id(d)] = (d, rng.uniform(a, b, size=(n_reps, length)))
    return tapes


def run_oracle(model, om, dist_ids, rep, t0=0, base_seed=0, tapes=None, n_events=None, t_until=None, log=True):
    s = O.OracleSim(om, base_seed=base_seed, rep=rep, t0_ns=t0, log=log)
    if tapes:
        for d, v in tapes.values():
            for did in dist_ids[id(d)]:
                s.set_tape(did, v[rep_local(rep, tapes)])
    s.init()
    if n_events is not None:
        s.step_events(n_events)
    if t_until is not None:
        s.step_until(t_until)
    return s


def rep_local(rep, tapes):
    return rep


def assert_rep_parity(sim, osim, model, rep_local_idx, check_log=True):
    """Bit-exact comparison of one replication: status, clock, event count, full log, stats, stock contents."""
    st, now, ev = sim.status(rep_local_idx, 1)
    assert int(st[0]) == osim.status, "status rep %d: %d vs oracle %d" % (rep_local_idx, st[0], osim.status)
    assert int(ev[0]) == osim.n_events, "n_events rep %d: %d vs %d" % (rep_local_idx, ev[0], osim.n_events)
    if osim.status == 0:
        assert int(now[0]) == osim.now, "now rep %d: %d vs %d" % (rep_local_idx, now[0], osim.now)
    if check_log:
        glog, total = sim.read_log(rep_local_idx)
        olog = osim.log()
        assert total == len(olog), "log count rep %d: %d vs oracle %d" % (rep_local_idx, total, len(olog))
        assert len(glog) == len(olog), "log ring too small for the test (%d kept of %d)" % (len(glog), total)
        if glog.tobytes() != olog.tobytes():
            for i in range(len(olog)):
                if glog[i].tobytes() != olog[i].tobytes():
                    raise AssertionError("rep %d record %d differs:\n  gpu    %r\n  oracle %r\n  prev   %r" % (
                        rep_local_idx, i, glog[i], olog[i], olog[max(0, i - 3):i]))
    if osim.status != 0:
        return
    for ci, cm in enumerate(model.components):
        gs = sim.comp_stats(ci, rep_local_idx, 1)[0]
        os_ = osim.comp_stats(ci)
        if cm.variant.startswith("Vector3Container"):
            # container family: integer statistics as for the discrete components, and every container held (handle
            # + the bits of what it carries) in order
            gh, gr = sim.read_containers(ci, rep_local_idx)
            oh, orr = osim.containers(ci)
            assert list(gh) == list(oh), "comp %d containers %r vs %r" % (ci, list(gh), list(oh))
            assert gr.tobytes() == orr.tobytes(), "comp %d container resources\n%r\nvs\n%r" % (ci, gr, orr)
            if cm.variant.endswith("Stock"):
                assert int(gs["count"]) == int(os_["n_a"]), "cstock %d n_add" % ci
                assert int(gs["time_ns"]) == int(os_["t_a_ns"]), "cstock %d area" % ci
                assert int(gs["aux"]) == int(os_["t_b_ns"]), "cstock %d max occupancy" % ci
                assert int(gs["level"]) == len(oh)
                assert list(sim.read_stock(ci, rep_local_idx)) == list(oh)
            else:
                assert int(gs["count"]) == int(os_["n_b"]), "cproc %d n_finish %d vs %d" % (ci, gs["count"], os_["n_b"])
                assert int(gs["time_ns"]) == int(os_["t_a_ns"]), "cproc %d busy %d vs %d" % (ci, gs["time_ns"], os_["t_a_ns"])
                assert int(gs["aux"]) == int(os_["t_b_ns"]), "cproc %d delay ns" % ci
        elif cm.variant in _VECTOR_VARIANTS:
            # continuous components: fp64 values are compared BIT-exactly (same operation order, no FMA); the
            # north_star tolerance would be 1e-9 relative
            assert float(gs["fvalue"]).hex() == float(os_["fvalue"]).hex(), "comp %d fvalue %r vs %r" % (
                ci, gs["fvalue"], os_["fvalue"])
            assert float(gs["faux"]).hex() == float(os_["faux"]).hex(), "comp %d faux %r vs %r" % (
                ci, gs["faux"], os_["faux"])
            if cm.variant.endswith("Stock"):
                assert int(gs["count"]) == int(os_["n_a"]), "vstock %d adds" % ci
                got, want = sim.read_resource(ci, rep_local_idx), osim.vector_resource_n(ci)
                assert [x.hex() for x in got] == [x.hex() for x in want], "vstock %d resource %r vs %r" % (ci, got, want)
            else:
                assert int(gs["count"]) == int(os_["n_b"]), "vproc %d successes" % ci
                assert int(gs["time_ns"]) == int(os_["t_a_ns"]), "vproc %d busy" % ci
                assert int(gs["aux"]) == int(os_["t_b_ns"]), "vproc %d delay ns" % ci
                if "Combiner" not in cm.variant:
                    got, want = sim.read_resource(ci, rep_local_idx), osim.vector_resource_n(ci)
                    assert [x.hex() for x in got] == [x.hex() for x in want], "vproc %d in-flight" % ci
        elif cm.variant == "StringStock":
            items = osim.stock_items(ci)
            assert int(gs["count"]) == int(os_["n_a"]), "stock %d n_add" % ci
            assert int(gs["time_ns"]) == int(os_["t_a_ns"]), "stock %d area %d vs %d" % (ci, gs["time_ns"], os_["t_a_ns"])
            assert int(gs["aux"]) == int(os_["t_b_ns"]), "stock %d max occupancy" % ci
            assert int(gs["level"]) == len(items)
            assert list(sim.read_stock(ci, rep_local_idx)) == items, "stock %d contents" % ci
        elif cm.variant == "BasicEnvironment":
            pass
        else:
            assert int(gs["count"]) == int(os_["n_b"]), "proc %d n_finish %d vs %d" % (ci, gs["count"], os_["n_b"])
            assert int(gs["time_ns"]) == int(os_["t_a_ns"]), "proc %d busy %d vs %d" % (ci, gs["time_ns"], os_["t_a_ns"])
            assert int(gs["aux"]) == int(os_["t_b_ns"]), "proc %d delay ns %d vs %d" % (ci, gs["aux"], os_["t_b_ns"])
