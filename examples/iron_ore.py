"""quokkasim_examples/src/bin/iron_ore.rs restated on the batched executor: a user-defined resource (`IronOre`: five
masses, total() = fe + other_elements), the ComponentModel / ComponentLogger variants declared for it, and the user's
own flat log structs (`IronOreStockLog`, `IronOreProcessLog`, iron_ore.rs:98-231) in place of the nested default ones.

    python examples/iron_ore.py [n_replications] [output_dir]      (needs a CUDA device)
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import quokkasim_b200 as qk  # noqa: E402
from quokkasim_b200.csvlog import ryu_f64  # noqa: E402

# struct IronOre + its ResourceAdd / ResourceRemove / ResourceTotal / ResourceMultiply impls (iron_ore.rs:21-96), and the
# define_model_enums! block (iron_ore.rs:258-270) with its connection arms (:272-306)
IronOre = qk.define_vector_variants("IronOreEx", qk.define_resource(
    "IronOre", ("fe", "other_elements", "magnetite", "hematite", "limonite"), total=("fe", "other_elements")))

STOCK_HEADER = ("time,event_id,source_event_id,element_name,element_type,log_type,new_state,balance,vector_total,fe,"
                "other_elements,fe_%,magnetite,hematite,limonite")
PROCESS_HEADER = ("time,event_id,source_event_id,element_name,element_type,event_type,total,fe,other_elements,fe_%,"
                  "magnetite,hematite,limonite,message")


def _f(x):
    return None if x is None else ryu_f64(x)


def _ore_fields(v):
    if v is None:
        return [None] * 6
    return [_f(v.fe), _f(v.other_elements), _f(v.fe / v.total()), _f(v.magnetite), _f(v.hematite), _f(v.limonite)]


def iron_ore_stock_log(ev):
    """impl Serialize for IronOreStockLog (iron_ore.rs:165-207)"""
    head = [ev["time"], ev["event_id"], ev["source_event_id"], ev["element_name"], ev["element_type"]]
    if ev["log_type"] == "StateChange":
        return head + ["StateChange", ev["new_state"], None, None] + _ore_fields(None)
    v = ev["vector"]
    return head + [ev["log_type"], None, _f(ev["balance"]), _f(v.total())] + _ore_fields(v)


def iron_ore_process_log(ev):
    """impl Serialize for IronOreProcessLog (iron_ore.rs:106-138); other events panic there ("Unhandled ...")"""
    head = [ev["time"], ev["event_id"], ev["source_event_id"], ev["element_name"], ev["element_type"]]
    t = ev["event_type"]
    if t in ("ProcessStart", "ProcessSuccess"):
        return head + [t, _f(ev["quantity"])] + _ore_fields(ev["vectors"][0]) + [None]
    if t == "ProcessFailure":
        return head + [t, None] + _ore_fields(None) + [ev["reason"]]
    if t == "WithdrawRequest":
        return head + [t, None] + _ore_fields(None) + [None]
    raise RuntimeError("Unhandled VectorProcessLogType %s" % t)


def build():
    df = qk.DistributionFactory.new(123456789)
    stock1 = qk.ComponentModel.IronOreExStock(
        qk.VectorStock.new().with_name("MyStock1").with_code("SP1").with_type("IronOreStock")
        .with_initial_resource(IronOre(fe=60.0, other_elements=40.0, magnetite=10.0, hematite=5.0, limonite=15.0))
        .with_low_capacity(10.0).with_max_capacity(100.0), qk.Mailbox.new())
    process1 = qk.ComponentModel.IronOreExProcess(
        qk.VectorProcess.new().with_name("MyProcess1").with_code("P1").with_name("IronOreProcess")
        .with_process_quantity_distr(df.create(qk.DistributionConfig.Uniform(2.0, 8.0)))
        .with_process_time_distr(qk.Distribution.Constant(10.0)), qk.Mailbox.new())
    stock2 = qk.ComponentModel.IronOreExStock(
        qk.VectorStock.new().with_name("MyStock2").with_code("SP2").with_type("IronOreStock")
        .with_initial_resource(IronOre(fe=3.0, other_elements=2.0, magnetite=0.5, hematite=0.25, limonite=0.75))
        .with_low_capacity(10.0).with_max_capacity(100.0), qk.Mailbox.new())
    qk.connect_components(stock1, process1)
    qk.connect_components(process1, stock2)
    process_logger = qk.ComponentLogger.IronOreExProcessLogger(qk.VectorProcessLogger.new("IronOreProcessLogger"))
    stock_logger = qk.ComponentLogger.IronOreExStockLogger(qk.VectorStockLogger.new("IronOreStockLogger"))
    qk.connect_logger(stock_logger, stock1)
    qk.connect_logger(process_logger, process1)
    qk.connect_logger(stock_logger, stock2)
    return (stock1, process1, stock2), process_logger, stock_logger


def main(n_replications=1024, output_dir="outputs/iron_ore", lib=None):
    comps, process_logger, stock_logger = build()
    sim_builder = qk.BatchedSimInit.new()
    for c in comps:
        sim_builder = qk.register_component(sim_builder, c)
    simu, _ = sim_builder.init(0, n_replications=n_replications, base_seed=123456789, log_capacity=512, lib=lib)
    simu.step_until(qk.Duration.from_secs(300))
    os.makedirs(output_dir, exist_ok=True)
    process_logger.write_csv(output_dir, simu, 0, record_map=iron_ore_process_log, header=PROCESS_HEADER)
    stock_logger.write_csv(output_dir, simu, 0, record_map=iron_ore_stock_log, header=STOCK_HEADER)
    red = simu.reduce_stats()
    print("%d replications to 300 s; batches moved per replication: mean %.3f" % (
        n_replications, int(red[1]["sum_count"]) / n_replications))
    return simu


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1024, sys.argv[2] if len(sys.argv) > 2 else "outputs/iron_ore")
