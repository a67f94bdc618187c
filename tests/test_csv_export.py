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
