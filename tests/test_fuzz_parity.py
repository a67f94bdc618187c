"""Randomised parity: models assembled at random from the discrete and container families (shared stocks, feedback
loops, Constant times for exact ties, breakdown modes, environment stop/resume, initial contents, load/unload against
Vector3Stocks) are run through the kernel source and compared with the oracle record for record.  The CPU test drives
the g++ build of the kernel source (tests/emul); the GPU test drives the product library."""
import random

import numpy as np
import pytest

import helpers as H
import quokkasim_b200 as qk
from quokkasim_b200.workloads import Model, _log_all

CAPACITY_FAULTS = (4, 5, 6, 7)  # executor capacity faults (advisory stock capacity etc.): no counterpart in the oracle


def random_model(seed):
    rng = random.Random(seed)
    m = Model()
    m.log_unloggable = True
    df = qk.DistributionFactory(base_seed=1000 + seed, next_seed=0)
    container = rng.random() < 0.5
    fam = "Vector3Container" if container else "String"

    def dist(scale=1.0):
        k = rng.choice(["const", "const", "uniform", "tri", "exp"])
        a = rng.choice([1.0, 2.0, 3.0, 4.5]) * scale
        if k == "const":
            return qk.Distribution.Constant(a)
        if k == "uniform":
            d = df.create(qk.DistributionConfig.Uniform(0.5 * a, 1.5 * a))
        elif k == "tri":
            d = df.create(qk.DistributionConfig.Triangular(0.5 * a, 2.0 * a, a))
        else:
            d = df.create(qk.DistributionConfig.Exponential(a))
        m.random_dists.append(d)
        return d

    def cm(variant, comp):
        return m.reg(getattr(qk.ComponentModel, variant)(comp, qk.Mailbox.new()))

    n_init_total = [0]

    def stock(i):
        maxc = rng.randint(1, 6)
        low = rng.choice([0, 0, 1])
        n_init = rng.randint(0, maxc)
        s = qk.DiscreteStock.new().with_name("S%d" % i).with_code("S%d" % i).with_low_capacity(low).with_max_capacity(maxc)
        if container:
            items = [qk.Vector3Container("C%d_%d" % (i, k),
                                         qk.Vector3([rng.uniform(0, 3), rng.uniform(0, 3), 1.0]) if rng.random() < 0.6 else None,
                                         rng.choice([8.0, 12.0, 20.0])) for k in range(n_init)]
        else:
            items = ["I%d_%d" % (i, k) for k in range(n_init)]
        n_init_total[0] += n_init
        s = s.with_initial_resource(qk.ItemDeque(items)).with_capacity_hint(64)
        return cm(fam + "Stock", s)

    def maybe_delay(c):
        if rng.random() < 0.3:
            c = c.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("D", dist(20.0), dist(3.0))))
        return c

    n_stages = rng.randint(2, 4)
    stocks = [stock(i) for i in range(n_stages + 1)]
    procs = []
    if container:
        fac = qk.Vector3ContainerFactory("B", 0, 3, dist(2.0) if rng.random() < 0.7 else None, [0.5, 0.25, 0.25],
                                         rng.choice([10.0, 15.0]))
        src = cm(fam + "Source", maybe_delay(qk.DiscreteSource.new().with_name("Src").with_code("SRC")
                                              .with_item_factory(fac).with_process_time_distr(dist(2.0))))
    else:
        src = cm(fam + "Source", maybe_delay(qk.DiscreteSource.new().with_name("Src").with_code("SRC")
                                              .with_process_time_distr(dist(2.0))))
    qk.connect_components(src, stocks[0])
    procs.append(src)
    for i in range(n_stages):
        for k in range(rng.randint(1, 2)):
            name = "P%d_%d" % (i, k)
            kinds = ["Process", "Process", "ParallelProcess"] + (["LoadProcess", "UnloadProcess"] if container else [])
            kind = rng.choice(kinds)
            if kind == "Process":
                p = cm(fam + kind, maybe_delay(qk.DiscreteProcess.new().with_name(name).with_code(name)
                                               .with_process_time_distr(dist())))
            elif kind == "ParallelProcess":
                p = cm(fam + kind, maybe_delay(qk.DiscreteParallelProcess.new().with_name(name).with_code(name)
                                               .with_process_time_distr(dist()).with_capacity_hint(96)))
            elif kind == "LoadProcess":
                p = cm(fam + kind, maybe_delay(qk.ContainerLoadingProcess.new().with_name(name).with_code(name)
                                               .with_max_process_count(rng.randint(1, 3)).with_process_time_distr(dist())
                                               .with_process_capacity_ratio_distr(
                                                   rng.choice([qk.Distribution.Constant(0.5), qk.Distribution.Constant(1.0)]))))
                bulk = cm("Vector3Stock", qk.VectorStock.new().with_name("B" + name).with_code("B" + name)
                          .with_low_capacity(5.0).with_max_capacity(1e6)
                          .with_initial_resource(qk.Vector3([500.0, 300.0, 200.0])))
                qk.connect_components(bulk, p)
            else:
                p = cm(fam + kind, maybe_delay(qk.ContainerUnloadingProcess.new().with_name(name).with_code(name)
                                               .with_max_process_count(rng.randint(1, 3)).with_process_time_distr(dist())
                                               .with_process_quantity_ratio_distr(
                                                   rng.choice([qk.Distribution.Constant(0.4), qk.Distribution.Constant(1.0)]))))
                out = cm("Vector3Stock", qk.VectorStock.new().with_name("O" + name).with_code("O" + name)
                         .with_low_capacity(0.0).with_max_capacity(rng.choice([30.0, 1e6])))
                qk.connect_components(p, out)
            qk.connect_components(stocks[i], p)
            qk.connect_components(p, stocks[i + 1])
            procs.append(p)
    if rng.random() < 0.4:  # feedback loop
        p = cm(fam + "Process", qk.DiscreteProcess.new().with_name("Back").with_code("BK").with_process_time_distr(dist()))
        qk.connect_components(stocks[-1], p)
        qk.connect_components(p, stocks[0])
        procs.append(p)
    snk = cm(fam + "Sink", maybe_delay(qk.DiscreteSink.new().with_name("Snk").with_code("SNK")
                                        .with_process_time_distr(dist(1.5))))
    qk.connect_components(stocks[-1], snk)
    procs.append(snk)
    if rng.random() < 0.5:
        env = cm("BasicEnvironment", qk.BasicEnvironment.new().with_name("Env").with_code("ENV"))
        for p in procs:
            if p.variant != "Vector3ContainerParallelProcess" and rng.random() < 0.6:
                qk.connect_components(env, p)
        t = 0
        for k in range(rng.randint(1, 4)):
            t += rng.randint(5, 40)
            m.ext.append((qk.Duration.from_secs(t), env, k % 2 == 0 and 1 or 0))
    _log_all(m)
    return m


def _run(lib, seed, n_reps, flags=0):
    m = random_model(seed)
    om, dist_ids = H.lower_to_oracle(m)
    sim = m.sim(lib, n_reps, base_seed=seed, log_capacity=16384, flags=flags)
    sim.step_until(120 * 10**9)
    st, _, _ = sim.status()
    compared = 0
    for r in range(n_reps):
        if int(st[r]) in CAPACITY_FAULTS:
            continue
        osim = H.run_oracle(m, om, dist_ids, r, base_seed=seed, t_until=120 * 10**9)
        H.assert_rep_parity(sim, osim, m, r)
        compared += 1
    sim.close()
    return compared


@pytest.mark.parametrize("seed", range(40))
def test_random_models_match_the_oracle(emul_lib, seed):
    _run(emul_lib, seed, 2, flags=seed % 2)  # both state placements


@pytest.mark.parametrize("seed", range(300, 324))
def test_random_models_match_the_oracle_lane_groups(emul_lib, seed):
    """the lane-group kernel (qk_group.cuh), 2 / 4 / 8 lanes per replication"""
    _run(emul_lib, seed, 2, flags=4 | ((2, 4, 8)[seed % 3] << 8))


@pytest.mark.parametrize("seed", range(600, 624))
def test_random_models_match_the_oracle_phased(emul_lib, seed):
    """the phased lane-group kernel (pop + quick by lane groups, handlers one thread per replication)"""
    _run(emul_lib, seed, 3, flags=4 | 8 | ((2, 4, 8)[seed % 3] << 8))
    if seed < 608:
        _run_vector(emul_lib, seed, 2, flags=4 | 8 | ((2, 4, 8)[seed % 3] << 8))


@pytest.mark.parametrize("seed", range(700, 732))
def test_random_models_match_the_oracle_split_placement(emul_lib, seed):
    """QK_BATCH_STATE_SPLIT: scheduler queue + stock words in the hot planes, component records in the cold planes"""
    _run(emul_lib, seed, 2, flags=16)
    if seed < 712:
        _run_vector(emul_lib, seed, 2, flags=16)
    # ... with the two-tier scheduler queue
    _run(emul_lib, seed, 2, flags=16 | 32)
    if seed < 712:
        _run_vector(emul_lib, seed, 2, flags=16 | 32)


def test_random_models_are_mostly_comparable(emul_lib):
    """the generator must not hide behind capacity faults"""
    assert sum(_run(emul_lib, s, 1) for s in range(40, 60)) >= 17


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(100, 112))
def test_random_models_match_the_oracle_on_gpu(gpu_lib, seed):
    assert _run(gpu_lib, seed, 96) >= 0


# ------------------------------------------------------------------------------------------ continuous family
# user-defined resources (iron_ore.rs): 2 to 8 fields, totals over a subset of them -- feature set 3 of the kernels
_WIDE = []
for _name, _n, _counted in (("FzPair", 2, (1,)), ("FzTri", 3, (0, 1)), ("FzQuad", 4, (0, 2)), ("FzOre", 5, (0, 1)),
                            ("FzSept", 7, (0, 1, 2, 3, 4, 5, 6)), ("FzOct", 8, (0, 2, 4, 5, 7))):
    _fields = ["f%d" % k for k in range(_n)]
    _WIDE.append(qk.define_vector_variants(_name, qk.define_resource(_name, _fields, total=[_fields[k] for k in _counted])))


def random_vector_model(seed, wide=False):
    """A random chain of the continuous components: Source -> stocks -> Process / Combiner2 / Splitter2 -> Sink, f64 or
    Vector3 (wide: a user-defined resource of 2..8 fields whose total counts some of them), random capacities around the
    flow so that Full / Empty flips happen, breakdown modes, environment."""
    rng = random.Random(10_000 + seed)
    m = Model()
    df = qk.DistributionFactory(base_seed=77 + seed, next_seed=0)
    dim = rng.choice([1, 3])
    P = "Vector3" if dim == 3 else "F64"
    V = (lambda v: qk.Vector3(v)) if dim == 3 else (lambda v: float(sum(v)))
    if wide:
        R = _WIDE[seed % len(_WIDE)]
        P = R.NAME
        n = len(R.FIELDS)
        counted = [k for k in range(n) if (R.TOTAL_MASK >> k) & 1]

        def V(v, R=R, n=n, counted=counted):  # noqa: E306
            # the three numbers of the plain generator go to the counted fields (so capacities mean what they meant),
            # the others carry their own quantities along
            vals = [0.25 * (k + 1) * v[k % 3] for k in range(n)]
            for j, k in enumerate(counted):
                vals[k] = v[j % 3] if j < 3 else 0.125 * v[j % 3]
            return R(*vals)

    def dist(a):
        k = rng.choice(["const", "uniform", "uniform", "tri"])
        if k == "const":
            return qk.Distribution.Constant(a)
        if k == "uniform":
            d = df.create(qk.DistributionConfig.Uniform(0.5 * a, 1.5 * a))
        else:
            d = df.create(qk.DistributionConfig.Triangular(0.5 * a, 2.0 * a, a))
        m.random_dists.append(d)
        return d

    def cm(kind, comp):
        return m.reg(getattr(qk.ComponentModel, P + kind)(comp, qk.Mailbox.new()))

    def stock(name):
        init = [rng.uniform(0, 4), rng.uniform(0, 4), rng.uniform(0, 4)]
        return cm("Stock", qk.VectorStock.new().with_name(name).with_code(name)
                  .with_low_capacity(rng.choice([0.0, 1.0, 3.0])).with_max_capacity(rng.choice([8.0, 15.0, 40.0]))
                  .with_initial_resource(V(init)))

    def proc(c):
        c = c.with_process_quantity_distr(dist(rng.choice([0.5, 1.0, 2.0]))).with_process_time_distr(dist(rng.choice([1.0, 2.0])))
        if rng.random() < 0.35:
            c = c.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("Brk", dist(15.0), dist(3.0))))
        return c

    src = cm("Source", proc(qk.VectorSource.new().with_name("Src").with_code("SRC").with_source_vector(V([1.0, 2.0, 3.0]))))
    s = stock("S0")
    qk.connect_components(src, s)
    actors = [src]
    for i in range(rng.randint(1, 3)):
        kind = rng.choice(["Process", "Combiner2", "Splitter2"])
        nxt = stock("S%d" % (i + 1))
        if kind == "Process":
            p = cm("Process", proc(qk.VectorProcess.new().with_name("P%d" % i).with_code("P%d" % i)))
            qk.connect_components(s, p)
            qk.connect_components(p, nxt)
        elif kind == "Combiner2":
            side = stock("X%d" % i)
            p = cm("Combiner2", proc(qk.VectorCombiner.new().with_name("C%d" % i).with_code("C%d" % i)))
            qk.connect_components(s, p, 0)
            qk.connect_components(side, p, 1)
            qk.connect_components(p, nxt)
        else:
            side = stock("Y%d" % i)
            r = rng.choice([0.3, 0.5])
            p = cm("Splitter2", proc(qk.VectorSplitter.new().with_name("L%d" % i).with_code("L%d" % i)
                                      .with_split_ratios([r, 1.0 - r])))
            qk.connect_components(s, p)
            qk.connect_components(p, nxt, 0)
            qk.connect_components(p, side, 1)
        actors.append(p)
        s = nxt
    snk = cm("Sink", proc(qk.VectorSink.new().with_name("Snk").with_code("SNK")))
    qk.connect_components(s, snk)
    actors.append(snk)
    if rng.random() < 0.5:
        env = m.reg(qk.ComponentModel.BasicEnvironment(qk.BasicEnvironment.new().with_name("Env").with_code("ENV"),
                                                       qk.Mailbox.new()))
        for a in actors:
            if rng.random() < 0.7:
                qk.connect_components(env, a)
        m.ext.append((qk.Duration.from_secs(rng.randint(5, 30)), env, qk.BasicEnvironmentState.Stopped))
        m.ext.append((qk.Duration.from_secs(rng.randint(31, 60)), env, qk.BasicEnvironmentState.Normal))
    _log_all(m)
    return m


def _run_vector(lib, seed, n_reps, flags=0, wide=False):
    m = random_vector_model(seed, wide)
    om, dist_ids = H.lower_to_oracle(m)
    sim = m.sim(lib, n_reps, base_seed=seed, log_capacity=32768, flags=flags)
    sim.step_until(90 * 10**9)
    st, _, _ = sim.status()
    for r in range(n_reps):
        if int(st[r]) in CAPACITY_FAULTS:
            continue
        H.assert_rep_parity(sim, H.run_oracle(m, om, dist_ids, r, base_seed=seed, t_until=90 * 10**9), m, r)
    sim.close()


@pytest.mark.parametrize("seed", range(24))
def test_random_vector_models_match_the_oracle(emul_lib, seed):
    _run_vector(emul_lib, seed, 2, flags=seed % 2)


@pytest.mark.parametrize("seed", range(600, 624))
def test_random_user_defined_resource_models_match_the_oracle(emul_lib, seed):
    """feature set 3 (qk_model_set_vector_n): on-chip, HBM-plane, split and two-tier placements"""
    _run_vector(emul_lib, seed, 2, flags=(0, 1, 16, 48)[seed % 4], wide=True)


@pytest.mark.parametrize("seed", range(400, 408))
def test_random_vector_models_match_the_oracle_lane_groups(emul_lib, seed):
    _run_vector(emul_lib, seed, 2, flags=4 | ((2, 4, 8)[seed % 3] << 8))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(500, 512))
def test_random_models_match_the_oracle_lane_groups_on_gpu(gpu_lib, seed):
    _run(gpu_lib, seed, 48, flags=4 | ((4, 8, 16)[seed % 3] << 8))
    if seed < 504:
        _run_vector(gpu_lib, seed, 48, flags=4 | ((4, 8, 16)[seed % 3] << 8))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(800, 812))
def test_random_models_match_the_oracle_split_placement_on_gpu(gpu_lib, seed):
    _run(gpu_lib, seed, 96, flags=16)
    _run(gpu_lib, seed, 96, flags=16 | 32)
    if seed < 804:
        _run_vector(gpu_lib, seed, 96, flags=16)
        _run_vector(gpu_lib, seed, 96, flags=16 | 32)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(200, 208))
def test_random_vector_models_match_the_oracle_on_gpu(gpu_lib, seed):
    _run_vector(gpu_lib, seed, 64)
