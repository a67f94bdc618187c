"""binary log -> the reference's CSV files.  The golden files under tests/golden/kat0_*.csv are the
KAT-0 trace of SURVEY A.8 (discrete_queue.rs, tape P1=[4,5], P2=[2.5], step_until(7.000000002 s)) written out by hand
with the formatting rules of the reference's Serialize impls (discrete.rs:295-318, 608-633) -- the reference itself
cannot run here.  CPU tests drive the kernel source through tests/emul; the GPU test runs the product library."""
import os

import numpy as np
import pytest

import helpers as H
from quokkasim_b200 import csvlog

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _kat0(lib):
    m = H.discrete_queue()
    sim = m.sim(lib, 2, log_capacity=1024)
    tapes = [np.array([[4.0, 5.0] + [6.0] * 20] * 2), np.array([[2.5] + [6.0] * 21] * 2)]
    for d, v in zip(m.random_dists, tapes):
        sim.set_tape(d, v)
    sim.step_until(7_000_000_002)
    return m, sim


def _check_kat0(lib, tmp_path):
    m, sim = _kat0(lib)
    stock_logger, process_logger = m.loggers[0], m.loggers[1]
    for lg, name in ((stock_logger, "kat0_StockLogger.csv"), (process_logger, "kat0_ProcessLogger.csv")):
        want = open(os.path.join(GOLDEN, name)).read()
        assert csvlog.csv_text(lg, sim, replication=1) == want
    # logger.write_csv(dir) -> "{dir}/{name}.csv" (core.rs:388-401)
    path = stock_logger.write_csv(str(tmp_path), sim, replication=0)
    assert path.endswith("StockLogger.csv")
    assert open(path).read() == open(os.path.join(GOLDEN, "kat0_StockLogger.csv")).read()
    sim.close()


def test_kat0_csv_matches_hand_written_reference_format(emul_lib, tmp_path):
    _check_kat0(emul_lib, tmp_path)


@pytest.mark.gpu
def test_kat0_csv_on_gpu(gpu_lib, tmp_path):
    _check_kat0(gpu_lib, tmp_path)


def test_ryu_style_floats():
    cases = {1.0: "1.0", 0.1: "0.1", 100.0: "100.0", 1e21: "1e21", 1e16: "1e16", 1.2345e16: "1.2345e16",
             9.999e15: "9999000000000000.0", 1.5e-7: "1.5e-7", 1e-5: "0.00001", 1.234e-5: "0.00001234",
             123456.789: "123456.789", 5e-324: "5e-324", -2.5: "-2.5", 0.0: "0.0", 2 / 3: "0.6666666666666666",
             1e-6: "1e-6", 12345678901234567.0: "1.2345678901234568e16"}
    for x, want in cases.items():
        assert csvlog.ryu_f64(x) == want, (x, csvlog.ryu_f64(x), want)
        assert float(csvlog.ryu_f64(x)) == x


def test_timestamp_formats():
    assert csvlog.chrono_utc(0) == "1970-01-01 00:00:00 UTC"
    assert csvlog.chrono_utc(3_000_000_001) == "1970-01-01 00:00:03.000000001 UTC"
    assert csvlog.chrono_utc(3_500_000_000) == "1970-01-01 00:00:03.500 UTC"
    assert csvlog.chrono_utc(3_000_250_000) == "1970-01-01 00:00:03.000250 UTC"
    assert csvlog.chrono_utc(1735689600 * 10**9) == "2025-01-01 00:00:00 UTC"
    assert csvlog.tai_display(4 * 10**9) == "1970-01-01 00:00:04"


def test_csv_quoting_and_json_strings():
    assert csvlog.csv_field(None) == ""
    assert csvlog.csv_field("Upstream is empty") == "Upstream is empty"
    assert csvlog.csv_field('"Item_0001"') == '"""Item_0001"""'
    assert csvlog.csv_field("a,b") == '"a,b"'
    assert csvlog.json_string('he said "hi"\n') == '"he said \\"hi\\"\\n"'


def test_vector_and_environment_csv(emul_lib):
    """vector.rs:233-256, 594-720 and core.rs:306-325 column layouts; DelayStart names travel in `reason`"""
    m = H.vector_flow(3)
    sim = m.sim(emul_lib, 1, log_capacity=16384)
    sim.step_until(60 * 10**9)
    sl, pl, el = m.loggers[3]["Vector3"], m.loggers[4]["Vector3"], m.loggers[2]
    stock = csvlog.csv_text(sl, sim).splitlines()
    proc = csvlog.csv_text(pl, sim).splitlines()
    env = csvlog.csv_text(el, sim).splitlines()
    assert stock[0] == "time,event_id,source_event_id,element_name,element_type,log_type,balance,vector_total,vector,new_state"
    assert proc[0] == "time,event_id,source_event_id,element_name,element_type,event_type,total,inflows,outflows,reason"
    assert env == ["time,event_id,source_event_id,element_name,element_type,event",
                   "1970-01-01 00:00:00,ENV_000000,INIT_000000,Env,BasicEnvironment,Normal",
                   "1970-01-01 00:00:40,ENV_000001,SCH_000000,Env,BasicEnvironment,Stopped",
                   "1970-01-01 00:00:55,ENV_000002,SCH_000000,Env,BasicEnvironment,Normal"]
    assert any(",emit_change,,,,Normal" in ln or ",emit_change,,,,Empty" in ln for ln in stock)
    assert any(",add," in ln and '"{""x"":' in ln for ln in stock)
    assert any(",CombineStart," in ln and ln.count('""x""') == 2 for ln in proc)
    assert any(",SplitSuccess," in ln and ln.count('""x""') == 2 for ln in proc)
    assert any(ln.endswith(",ProcessStopped,,,,Stopped by environment") for ln in proc)
    assert any(ln.endswith(",DelayStart,,,,Breakdown") or ln.endswith(",DelayStart,,,,Service") for ln in proc)
    # every row has exactly the header's number of fields
    import csv as pycsv
    import io

    for text in (csvlog.csv_text(sl, sim), csvlog.csv_text(pl, sim)):
        rows = list(pycsv.reader(io.StringIO(text)))
        assert all(len(r) == len(rows[0]) for r in rows)
    sim.close()


def _kat1_model():
    """KAT-1, hand-derived from vector_container.rs:263-424 + discrete.rs:161-233 (the reference cannot run here):
    Bulk = Vector3Stock [60,30,10]; A = container stock holding Truck_0 (no resource, capacity 10, max 5);
    L = ContainerLoadingProcess (code CLP by default), max_process_count 1, time Constant(5), capacity ratio
    Constant(0.5); Q = container stock (max 5).
      t=0       init L: WithdrawRequest, A.remove -> Truck_0 (A flips Normal->Empty: emit at 1 ns); quantity =
                (10 - 0) * 0.5 = 5 of Bulk (total 100): p = 0.05 -> [3, 1.5, 0.5]; ProcessStart; second loop trip:
                no slot left -> "No remaining process slots"; keyed event at 5 s carrying CLP_000002
      t=1ns     A emit_change -> L.update_state(A_000001): "No remaining process slots"; 1ns + (5s - 1ns) = 5 s is not
                sooner than the pending event, which stays (and keeps its source id CLP_000002)
      t=5s      ProcessFinish (source CLP_000002), Q.add (Q flips Empty->Normal), WithdrawRequest, A.remove -> None
      t=5s+1ns  Q emit_change -> L.update_state(Q_000001): WithdrawRequest, A.remove -> None"""
    import quokkasim_b200 as qk
    from quokkasim_b200.workloads import Model, _log_all

    m = Model()
    C = qk.Distribution.Constant
    bulk = m.reg(qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Bulk").with_code("BLK").with_max_capacity(1000.0)
        .with_initial_resource(qk.Vector3([60.0, 30.0, 10.0])), qk.Mailbox.new()))
    a = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("A").with_code("A").with_max_capacity(5)
        .with_initial_resource(qk.ItemDeque([qk.Vector3Container("Truck_0", None, 10.0)])), qk.Mailbox.new()))
    ld = m.reg(qk.ComponentModel.Vector3ContainerLoadProcess(
        qk.ContainerLoadingProcess.new().with_name("L").with_process_time_distr(C(5.0))
        .with_process_capacity_ratio_distr(C(0.5)), qk.Mailbox.new()))
    q = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("Q").with_code("Q").with_max_capacity(5), qk.Mailbox.new()))
    qk.connect_components(bulk, ld)
    qk.connect_components(a, ld)
    qk.connect_components(ld, q)
    _log_all(m)
    return m


def _check_kat1(lib):
    m = _kat1_model()
    sim = m.sim(lib, 2, log_capacity=256)
    sim.step_until(60 * 10**9)
    csl, cpl = m.container_loggers
    for lg, name in ((csl, "kat1_ContainerStockLogger.csv"), (cpl, "kat1_ContainerProcessLogger.csv")):
        assert csvlog.csv_text(lg, sim, replication=1) == open(os.path.join(GOLDEN, name)).read()
    assert sim.read_vector(m.by_name["Bulk"].component._id, 0) == [57.0, 28.5, 9.5]
    assert sim.summary()["n_events"] == 2 * 3
    sim.close()


def test_kat1_container_csv_matches_hand_derived_trace(emul_lib):
    _check_kat1(emul_lib)
    # the oracle reproduces the same hand-derived trace (record for record, through the same decoder)
    m = _kat1_model()
    om, dist_ids = H.lower_to_oracle(m)
    osim = H.run_oracle(m, om, dist_ids, 0, t_until=60 * 10**9)
    sim = m.sim(emul_lib, 1, log_capacity=256)
    sim.step_until(60 * 10**9)
    assert sim.read_log(0)[0].tobytes() == osim.log().tobytes()


@pytest.mark.gpu
def test_kat1_container_csv_on_gpu(gpu_lib):
    _check_kat1(gpu_lib)


def test_combiner_and_splitter_failures_are_process_failure_rows(emul_lib):
    """vector.rs:937-949, 1235-1247: a combiner / splitter that cannot start logs VectorProcessLogType::ProcessFailure,
    which serialises as event_type "ProcessFailure" (vector.rs:621-622) -- only Start and Success carry the Combine /
    Split prefix.  Hand-derived model: the combiner's second upstream is below its low capacity (state Empty) at init,
    the splitter's upstream (what the combiner feeds) is Empty too."""
    import quokkasim_b200 as qk
    from quokkasim_b200.workloads import Model, _log_all

    m = Model()
    c = qk.Distribution.Constant

    def pile(name, low, maxc, init):
        return m.reg(qk.ComponentModel.F64Stock(
            qk.VectorStock.new().with_name(name).with_code(name).with_low_capacity(low).with_max_capacity(maxc)
            .with_initial_resource(init), qk.Mailbox.new()))

    a, b, mid, o1, o2 = pile("A", 1.0, 100.0, 50.0), pile("B", 1.0, 100.0, 0.0), pile("M", 1.0, 100.0, 0.0), \
        pile("O1", 0.0, 100.0, 0.0), pile("O2", 0.0, 100.0, 0.0)
    cmb = m.reg(qk.ComponentModel.F64Combiner2(
        qk.VectorCombiner.new().with_name("Combiner").with_code("CMB").with_process_quantity_distr(c(5.0))
        .with_process_time_distr(c(3.0)), qk.Mailbox.new()))
    spl = m.reg(qk.ComponentModel.F64Splitter2(
        qk.VectorSplitter.new().with_name("Splitter").with_code("SPL").with_split_ratios([0.5, 0.5])
        .with_process_quantity_distr(c(5.0)).with_process_time_distr(c(3.0)), qk.Mailbox.new()))
    qk.connect_components(a, cmb, 0)
    qk.connect_components(b, cmb, 1)
    qk.connect_components(cmb, mid)
    qk.connect_components(mid, spl)
    qk.connect_components(spl, o1, 0)
    qk.connect_components(spl, o2, 1)
    _log_all(m)
    sim = m.sim(emul_lib, 1, log_capacity=256)
    sim.step_until(10 * 10**9)
    rows = csvlog.csv_text(m.loggers[4]["F64"], sim).splitlines()
    assert "Failure" in "\n".join(rows) and "CombineFailure" not in "\n".join(rows) and "SplitFailure" not in "\n".join(rows)
    # Combiner / Splitter init carry their own next id as source (vector.rs:774, 1111)
    assert rows[1] == ("1970-01-01 00:00:00 UTC,CMB_000000,CMB_000000,Combiner,VectorCombiner,ProcessFailure,,,,"
                       "At least one upstream is empty")
    assert rows[2] == "1970-01-01 00:00:00 UTC,SPL_000000,SPL_000000,Splitter,VectorSplitter,ProcessFailure,,,,Upstream is empty"
    sim.close()


def _check_iron_ore_example(lib, tmp_path):
    """examples/iron_ore.py = iron_ore.rs: the user's flat log structs over a five-field resource.  The removed vector
    of the first batch is re-derived here from the logged quantity with the reference's own arithmetic
    (iron_ore.rs:52-73: proportion of the masked total, every field scaled)."""
    import csv
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "examples"))
    import iron_ore

    sim = iron_ore.main(4, str(tmp_path), lib=lib)
    sim.close()
    with open(os.path.join(str(tmp_path), "IronOreProcessLogger.csv")) as f:
        prows = list(csv.reader(f))
    with open(os.path.join(str(tmp_path), "IronOreStockLogger.csv")) as f:
        srows = list(csv.reader(f))
    assert prows[0] == iron_ore.PROCESS_HEADER.split(",") and srows[0] == iron_ore.STOCK_HEADER.split(",")
    assert prows[1][:6] == ["1970-01-01 00:00:00 UTC", "P1_000000", "INIT_000000", "IronOreProcess", "VectorProcess",
                            "WithdrawRequest"]
    assert prows[2][5] == "ProcessStart" and prows[3][5] == "ProcessSuccess" and prows[3][0] == "1970-01-01 00:00:10 UTC"
    q = float(prows[2][6])  # the sampled process quantity, Uniform(2, 8)
    assert 2.0 <= q < 8.0
    s1 = [60.0, 40.0, 10.0, 5.0, 15.0]
    p = q / (s1[0] + s1[1])
    removed = [x * p for x in s1]
    left = [x * (1.0 - p) for x in s1]
    r = srows[1]
    assert r[:6] == ["1970-01-01 00:00:00 UTC", "SP1_000000", "P1_000000", "MyStock1", "IronOreStock", "Remove"]
    assert r[6] == "" and r[7] == csvlog.ryu_f64(left[0] + left[1]) and r[8] == csvlog.ryu_f64(removed[0] + removed[1])
    assert r[9:] == [csvlog.ryu_f64(x) for x in (removed[0], removed[1], removed[0] / (removed[0] + removed[1]),
                                                 removed[2], removed[3], removed[4])]
    assert prows[2][7:13] == r[9:]
    assert srows[2][5:7] == ["StateChange", "Normal"] and srows[2][0] == "1970-01-01 00:00:00.000000001 UTC"
    add = srows[3]
    assert add[1:6] == ["SP2_000000", "P1_000002", "MyStock2", "IronOreStock", "Add"]
    assert add[7] == csvlog.ryu_f64((3.0 + removed[0]) + (2.0 + removed[1])) and add[9:] == r[9:]
    # every replication has its own Philox stream: the CSV of another replication differs in the quantities only
    assert len(srows) > 20 and all(len(x) == 15 for x in srows) and all(len(x) == 14 for x in prows)


def test_user_defined_resource_flat_csv(emul_lib, tmp_path):
    _check_iron_ore_example(emul_lib, tmp_path)


@pytest.mark.gpu
def test_user_defined_resource_flat_csv_on_gpu(gpu_lib, tmp_path):
    _check_iron_ore_example(gpu_lib, tmp_path)
