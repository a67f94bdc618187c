/*
 * quokka_b200.hpp -- C++ host side above the C ABI (include/quokka_b200.h): the model-builder surface of
 * jajetloh/quokkasim v0.2.2 -- the discrete family, the continuous (vector) family, the container family and
 * BasicEnvironment -- restated as plain descriptions that are lowered to C-ABI rows.  Header only, C++17, no
 * dependencies beyond the C header.  (CSV export covers the discrete and container loggers; the continuous ones are
 * in the Python mirror.)
 *
 * The reference is Rust; there is no Rust toolchain in the build image, so this is the compiled-language host mirror
 * of its operator interface (same names, argument meaning and error behaviour), next to the Python mirror
 * (quokkasim_b200/api.py) and the source-only Rust shim.
 *
 *   reference (paths relative to /root/reference/quokkasim/src)                   here (namespace quokka)
 *   ------------------------------------------------------------------------     -------------------------------------
 *   DistributionConfig / DistributionFactory::create      common.rs:37-148       DistributionConfig / DistributionFactory
 *   #[derive(WithMethods)] builders  quokkasim_derive_macros/src/lib.rs:22-94    .with_name() .with_code() ... (by value)
 *   DiscreteStock / Process / Source / Sink / ParallelProcess  components/discrete.rs   DiscreteStock::new_() ...
 *   VectorStock / Process / Source / Sink / Combiner / Splitter  components/vector.rs   VectorStock::new_() ...
 *   Vector3Container, Vector3ContainerFactory, ContainerLoadingProcess,
 *   ContainerUnloadingProcess                      components/vector_container.rs     same names
 *   ComponentModel::StringStock(c, Mailbox::new()) ...    core.rs:505-557        ComponentModel::StringStock(c, Mailbox::new_())
 *   connect_components!(&mut a, &mut b)                   core.rs:565-1178       connect_components(a, b)  (throws the Err text)
 *   connect_logger!(&mut logger, &mut c)                  core.rs:1303-1338      connect_logger(logger, c)
 *   register_component!(sim_init, c)                      core.rs:1424-1428      sim_init = register_component(sim_init, c)
 *   create_scheduled_event!(.., SetEnvironmentState, ..)  core.rs:1377-1401      create_scheduled_event(sim, t, cfg, addr)
 *   sim_init.init(t0) / simu.step_until(t)                discrete_queue.rs:109-111   BatchedSimInit::init / step_until
 *   logger.write_csv(dir)                                 core.rs:388-401        logger.write_csv(dir, sim, replication)
 *
 * All simulation work happens in libquokka_b200.so on the GPU; nothing here can advance a replication.
 */
#ifndef QUOKKA_B200_HPP
#define QUOKKA_B200_HPP

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "quokka_b200.h"

namespace quokka {

/* ---- errors: the reference's `Box<dyn Error>` strings ---------------------------------------------------------- */
struct Error : std::runtime_error {
  int status;
  Error(const std::string& msg, int st) : std::runtime_error(msg), status(st) {}
};
inline void check(int rc) {
  if (rc != QK_OK) throw Error(qk_last_error(), rc);
}

/* ---- time ------------------------------------------------------------------------------------------------------- */
struct Duration {
  static uint64_t from_secs(uint64_t s) { return s * 1000000000ull; }
  static uint64_t from_millis(uint64_t ms) { return ms * 1000000ull; }
  static uint64_t from_nanos(uint64_t ns) { return ns; }
};
struct MonotonicTime {
  static constexpr uint64_t EPOCH = 0; /* absolute ns since the TAI epoch */
};

/* ---- distributions (common.rs:25-148) --------------------------------------------------------------------------- */
struct Distribution {
  qk_dist_desc desc;
  std::shared_ptr<int> identity; /* components that share one Distribution value share its variate tape */
  static Distribution Constant(double v) {
    Distribution d;
    std::memset(&d.desc, 0, sizeof(d.desc));
    d.desc.kind = QK_DIST_CONSTANT;
    d.desc.p[0] = v;
    d.identity = std::make_shared<int>(0);
    return d;
  }
  static Distribution default_() { return Constant(1.0); } /* common.rs:181-185 */
};

struct DistributionConfig {
  uint32_t kind;
  double p[4];
  static DistributionConfig make(uint32_t k, double a = 0, double b = 0, double c = 0, double d = 0) {
    DistributionConfig x;
    x.kind = k;
    x.p[0] = a, x.p[1] = b, x.p[2] = c, x.p[3] = d;
    return x;
  }
  static DistributionConfig Uniform(double min, double max) { return make(QK_DIST_UNIFORM, min, max); }
  static DistributionConfig Triangular(double min, double max, double mode) { return make(QK_DIST_TRIANGULAR, min, max, mode); }
  static DistributionConfig Constant(double v) { return make(QK_DIST_CONSTANT, v); }
  static DistributionConfig Normal(double mean, double std) { return make(QK_DIST_NORMAL, mean, std); }
  static DistributionConfig Exponential(double mean) { return make(QK_DIST_EXPONENTIAL, mean); }
};

/* Seeds increment per created distribution, EXCEPT that the Normal / TruncNormal / Exponential arms of the reference
 * return early and skip the increment (common.rs:98,121,134 vs :145): reproduced, the stream id decides which
 * instances share a variate sequence. */
struct DistributionFactory {
  uint64_t base_seed, next_seed;
  explicit DistributionFactory(uint64_t seed) : base_seed(seed), next_seed(seed) {}
  Distribution create(const DistributionConfig& c) {
    if (c.kind == QK_DIST_CONSTANT) {
      next_seed += 1;
      return Distribution::Constant(c.p[0]);
    }
    if (c.kind == QK_DIST_TRIANGULAR && (!(c.p[1] >= c.p[2] && c.p[2] >= c.p[0]) || !(c.p[1] > c.p[0])))
      throw Error("invalid Triangular parameters", QK_ERR_INVALID_ARG);
    if (c.kind == QK_DIST_EXPONENTIAL && !(c.p[0] > 0)) throw Error("invalid Exponential parameters", QK_ERR_INVALID_ARG);
    Distribution d;
    std::memset(&d.desc, 0, sizeof(d.desc));
    d.desc.kind = c.kind;
    d.desc.stream = (uint32_t)next_seed;
    for (int i = 0; i < 4; ++i) d.desc.p[i] = c.p[i];
    d.identity = std::make_shared<int>(0);
    if (c.kind == QK_DIST_UNIFORM || c.kind == QK_DIST_TRIANGULAR) next_seed += 1;
    return d;
  }
};

struct DelayMode {
  std::string name;
  Distribution until_delay_distr, until_fix_distr;
};

/* discrete.rs:664-686: items are "{prefix}_{index:0>num_digits}" */
struct StringItemFactory {
  std::string prefix = "Item";
  uint64_t next_index = 0;
  size_t num_digits = 4;
  std::string item_name(uint64_t index) const {
    std::string n = std::to_string(next_index + index);
    if (n.size() < num_digits) n = std::string(num_digits - n.size(), '0') + n;
    return prefix + "_" + n;
  }
};

enum class BasicEnvironmentState { Normal = 0, Stopped = 1 };

/* core.rs:47-128 */
struct Vector3 {
  double values[3];
  Vector3(double x = 0, double y = 0, double z = 0) : values{x, y, z} {}
  double total() const { return (values[0] + values[1]) + values[2]; }
};
/* vector_container.rs:71-76 */
struct Vector3Container {
  std::string id;
  bool has_resource = false;
  Vector3 resource;
  double capacity = 0;
  Vector3Container() {}
  Vector3Container(std::string i, double cap) : id(std::move(i)), capacity(cap) {}
  Vector3Container(std::string i, const Vector3& r, double cap) : id(std::move(i)), has_resource(true), resource(r), capacity(cap) {}
};

/* ---- components: every component is one row -------------------------------------------------------------------- */
struct ComponentData {
  uint32_t kind = 0;
  std::string element_name, element_code, element_type;
  uint32_t low_capacity = 0, max_capacity = 1; /* discrete.rs:70-71 */
  std::vector<std::string> resource;            /* ItemDeque<String> */
  Distribution process_time_distr = Distribution::default_();
  std::vector<DelayMode> delay_modes;
  StringItemFactory item_factory;
  BasicEnvironmentState state = BasicEnvironmentState::Normal;
  int id = -1;
  bool logged = false;
  /* continuous family (vector.rs): resource type decided by the ComponentModel variant (dim 1 = f64, 3 = Vector3) */
  uint32_t dim = 0, n_ports = 0;
  double vlow = 0, vmax = 0;        /* VectorStock capacities (vector.rs:70-71) */
  bool vinit_set = false;
  Vector3 vinit;                    /* VectorStock initial resource / VectorSource source_vector */
  std::vector<double> split_ratios; /* VectorSplitter (default 1/N each), VectorCombiner (unused by the reference) */
  Distribution process_quantity_distr = Distribution::default_(); /* also the load / unload ratio distribution */
  /* container family (vector_container.rs) */
  std::vector<Vector3Container> containers; /* ItemDeque<Vector3Container> of a container stock */
  bool cfactory_set = false, cfactory_has_quantity = false;
  Distribution cfactory_quantity = Distribution::default_();
  double cfactory_split[3] = {0, 0, 0}, cfactory_capacity = 0;
  std::string cfactory_prefix;
  uint64_t cfactory_next_index = 0;
  size_t cfactory_num_digits = 0;
  uint32_t max_process_count = 1;
  uint32_t container_kind = 0; /* QK_CONTAINER_* once wrapped in a Vector3Container* variant */
};

/* vector_container.rs:126-156 */
struct Vector3ContainerFactory {
  std::string prefix;
  uint64_t next_index = 0;
  size_t num_digits = 0;
  bool has_quantity = false;
  Distribution create_quantity_distr = Distribution::default_();
  double create_component_split[3] = {0, 0, 0};
  double capacity = 0;
};

class Component {
 public:
  std::shared_ptr<ComponentData> d;
  Component(uint32_t kind, const char* type_name) : d(std::make_shared<ComponentData>()) {
    d->kind = kind;
    d->element_name = type_name;
    d->element_type = type_name;
  }
  Component with_name(const std::string& v) { d->element_name = v; return *this; }
  Component with_code(const std::string& v) { d->element_code = v; return *this; }
  Component with_type(const std::string& v) { d->element_type = v; return *this; }
  Component with_low_capacity(uint32_t v) { d->low_capacity = v; return *this; }
  Component with_max_capacity(uint32_t v) { d->max_capacity = v; return *this; }
  Component with_initial_resource(std::vector<std::string> items) { d->resource = std::move(items); return *this; }
  Component with_process_time_distr(const Distribution& x) { d->process_time_distr = x; return *this; }
  Component with_item_factory(const StringItemFactory& f) { d->item_factory = f; return *this; }
  Component with_state(BasicEnvironmentState s) { d->state = s; return *this; }
  /* continuous family */
  Component with_low_capacity_f(double v) { d->vlow = v; return *this; }
  Component with_max_capacity_f(double v) { d->vmax = v; return *this; }
  Component with_initial_resource(const Vector3& v) { d->vinit = v; d->vinit_set = true; return *this; }
  Component with_initial_resource(double v) { d->vinit = Vector3(v, 0, 0); d->vinit_set = true; return *this; }
  Component with_source_vector(const Vector3& v) { d->vinit = v; d->vinit_set = true; return *this; }
  Component with_source_vector(double v) { d->vinit = Vector3(v, 0, 0); d->vinit_set = true; return *this; }
  Component with_process_quantity_distr(const Distribution& x) { d->process_quantity_distr = x; return *this; }
  Component with_split_ratios(std::vector<double> r) { d->split_ratios = std::move(r); return *this; }
  /* container family */
  Component with_initial_resource(std::vector<Vector3Container> items) { d->containers = std::move(items); return *this; }
  Component with_item_factory(const Vector3ContainerFactory& f) {
    d->cfactory_set = true;
    d->cfactory_has_quantity = f.has_quantity;
    d->cfactory_quantity = f.create_quantity_distr;
    for (int k = 0; k < 3; ++k) d->cfactory_split[k] = f.create_component_split[k];
    d->cfactory_capacity = f.capacity;
    d->cfactory_prefix = f.prefix;
    d->cfactory_next_index = f.next_index;
    d->cfactory_num_digits = f.num_digits;
    return *this;
  }
  Component with_max_process_count(uint32_t n) { d->max_process_count = n; return *this; }
  Component with_process_capacity_ratio_distr(const Distribution& x) { d->process_quantity_distr = x; return *this; }
  Component with_process_quantity_ratio_distr(const Distribution& x) { d->process_quantity_distr = x; return *this; }
  /* DelayModes::modify(Add) (delays.rs:157-164): an existing name keeps its position */
  Component with_delay_mode(const DelayMode& m) {
    for (auto& x : d->delay_modes)
      if (x.name == m.name) {
        x = m;
        return *this;
      }
    d->delay_modes.push_back(m);
    return *this;
  }
};

#define QUOKKA_COMPONENT_TYPE(Name, Kind) \
  struct Name {                           \
    static Component new_() { return Component(Kind, #Name); } \
  };
QUOKKA_COMPONENT_TYPE(DiscreteStock, QK_DISCRETE_STOCK)
QUOKKA_COMPONENT_TYPE(DiscreteProcess, QK_DISCRETE_PROCESS)
QUOKKA_COMPONENT_TYPE(DiscreteSource, QK_DISCRETE_SOURCE)
QUOKKA_COMPONENT_TYPE(DiscreteSink, QK_DISCRETE_SINK)
QUOKKA_COMPONENT_TYPE(DiscreteParallelProcess, QK_DISCRETE_PARALLEL_PROCESS)
QUOKKA_COMPONENT_TYPE(VectorStock, QK_VECTOR_STOCK)
QUOKKA_COMPONENT_TYPE(VectorProcess, QK_VECTOR_PROCESS)
QUOKKA_COMPONENT_TYPE(VectorSource, QK_VECTOR_SOURCE)
QUOKKA_COMPONENT_TYPE(VectorSink, QK_VECTOR_SINK)
QUOKKA_COMPONENT_TYPE(VectorCombiner, QK_VECTOR_COMBINER)
QUOKKA_COMPONENT_TYPE(VectorSplitter, QK_VECTOR_SPLITTER)
QUOKKA_COMPONENT_TYPE(ContainerUnloadingProcess, QK_CONTAINER_UNLOAD_PROCESS)
#undef QUOKKA_COMPONENT_TYPE
struct ContainerLoadingProcess {
  static Component new_() {
    Component c(QK_CONTAINER_LOAD_PROCESS, "ContainerLoadingProcess");
    c.d->element_code = "CLP"; /* vector_container.rs:205 */
    return c;
  }
};
struct BasicEnvironment {
  static Component new_() {
    Component c(QK_ENVIRONMENT, "BasicEnvironment");
    c.d->element_name.clear();
    return c;
  }
};

/* kept so that `ComponentModel::StringStock(component, Mailbox::new_())` reads like the reference; inert */
struct Mailbox {
  static Mailbox new_() { return Mailbox(); }
};

struct Connection {
  std::shared_ptr<ComponentData> a, b;
  int n;
  uint64_t stamp;
};
inline uint64_t& connection_stamp() {
  static uint64_t s = 0;
  return s;
}

/* the generated enum of core.rs:504-562 (discrete family + BasicEnvironment) */
class ComponentModel {
 public:
  std::string variant;
  Component component;
  std::shared_ptr<std::vector<Connection>> conns = std::make_shared<std::vector<Connection>>();
  ComponentModel(std::string v, Component c) : variant(std::move(v)), component(std::move(c)) {}
#define QUOKKA_VARIANT(Name, Kind)                                                          \
  static ComponentModel Name(Component c, Mailbox) {                                         \
    if (c.d->kind != Kind) throw Error("ComponentModel::" #Name " expects another component type", QK_ERR_INVALID_ARG); \
    return ComponentModel(#Name, std::move(c));                                              \
  }
  QUOKKA_VARIANT(StringStock, QK_DISCRETE_STOCK)
  QUOKKA_VARIANT(StringProcess, QK_DISCRETE_PROCESS)
  QUOKKA_VARIANT(StringParallelProcess, QK_DISCRETE_PARALLEL_PROCESS)
  QUOKKA_VARIANT(StringSource, QK_DISCRETE_SOURCE)
  QUOKKA_VARIANT(StringSink, QK_DISCRETE_SINK)
  QUOKKA_VARIANT(BasicEnvironment, QK_ENVIRONMENT)
#undef QUOKKA_VARIANT
  /* continuous family (core.rs:505-533): the variant fixes the resource type and, for combiners / splitters, M / N */
#define QUOKKA_VVARIANT(Name, Kind, Dim, Ports)                                                          \
  static ComponentModel Name(Component c, Mailbox) {                                                     \
    if (c.d->kind != Kind) throw Error("ComponentModel::" #Name " expects another component type", QK_ERR_INVALID_ARG); \
    c.d->dim = Dim;                                                                                      \
    c.d->n_ports = Ports;                                                                                \
    if (Ports && c.d->split_ratios.empty()) c.d->split_ratios.assign(Ports, 1.0 / (Ports ? Ports : 1)); \
    if (Ports && c.d->split_ratios.size() != Ports) throw Error(#Name " needs one split ratio per port", QK_ERR_INVALID_ARG); \
    return ComponentModel(#Name, std::move(c));                                                          \
  }
#define QUOKKA_VFAMILY(P, Dim)                                   \
  QUOKKA_VVARIANT(P##Stock, QK_VECTOR_STOCK, Dim, 0)             \
  QUOKKA_VVARIANT(P##Process, QK_VECTOR_PROCESS, Dim, 0)         \
  QUOKKA_VVARIANT(P##Source, QK_VECTOR_SOURCE, Dim, 0)           \
  QUOKKA_VVARIANT(P##Sink, QK_VECTOR_SINK, Dim, 0)               \
  QUOKKA_VVARIANT(P##Combiner1, QK_VECTOR_COMBINER, Dim, 1)      \
  QUOKKA_VVARIANT(P##Combiner2, QK_VECTOR_COMBINER, Dim, 2)      \
  QUOKKA_VVARIANT(P##Combiner3, QK_VECTOR_COMBINER, Dim, 3)      \
  QUOKKA_VVARIANT(P##Combiner4, QK_VECTOR_COMBINER, Dim, 4)      \
  QUOKKA_VVARIANT(P##Combiner5, QK_VECTOR_COMBINER, Dim, 5)      \
  QUOKKA_VVARIANT(P##Splitter1, QK_VECTOR_SPLITTER, Dim, 1)      \
  QUOKKA_VVARIANT(P##Splitter2, QK_VECTOR_SPLITTER, Dim, 2)      \
  QUOKKA_VVARIANT(P##Splitter3, QK_VECTOR_SPLITTER, Dim, 3)      \
  QUOKKA_VVARIANT(P##Splitter4, QK_VECTOR_SPLITTER, Dim, 4)      \
  QUOKKA_VVARIANT(P##Splitter5, QK_VECTOR_SPLITTER, Dim, 5)
  QUOKKA_VFAMILY(F64, 1)
  QUOKKA_VFAMILY(Vector3, 3)
#undef QUOKKA_VFAMILY
#undef QUOKKA_VVARIANT
  /* container family (core.rs:549-555): the discrete component classes instantiated with Vector3Container */
#define QUOKKA_CVARIANT(Name, BaseKind, CKind)                                                           \
  static ComponentModel Name(Component c, Mailbox) {                                                     \
    if (c.d->kind != BaseKind) throw Error("ComponentModel::" #Name " expects another component type", QK_ERR_INVALID_ARG); \
    c.d->container_kind = CKind;                                                                         \
    c.d->dim = 3;                                                                                        \
    return ComponentModel(#Name, std::move(c));                                                          \
  }
  QUOKKA_CVARIANT(Vector3ContainerStock, QK_DISCRETE_STOCK, QK_CONTAINER_STOCK)
  QUOKKA_CVARIANT(Vector3ContainerProcess, QK_DISCRETE_PROCESS, QK_CONTAINER_PROCESS)
  QUOKKA_CVARIANT(Vector3ContainerParallelProcess, QK_DISCRETE_PARALLEL_PROCESS, QK_CONTAINER_PARALLEL_PROCESS)
  QUOKKA_CVARIANT(Vector3ContainerSource, QK_DISCRETE_SOURCE, QK_CONTAINER_SOURCE)
  QUOKKA_CVARIANT(Vector3ContainerSink, QK_DISCRETE_SINK, QK_CONTAINER_SINK)
  QUOKKA_CVARIANT(Vector3ContainerLoadProcess, QK_CONTAINER_LOAD_PROCESS, QK_CONTAINER_LOAD_PROCESS)
  QUOKKA_CVARIANT(Vector3ContainerUnloadProcess, QK_CONTAINER_UNLOAD_PROCESS, QK_CONTAINER_UNLOAD_PROCESS)
#undef QUOKKA_CVARIANT
};

/* connect_components!(&mut a, &mut b[, n]) (core.rs:565-1178; discrete arms :886-921, environment arms :1129-1148):
 * an undefined pair is the reference's Err("No component connection defined from {} to {} (n={:?})") */
inline void connect_components(ComponentModel& a, ComponentModel& b, int n = -1) {
  const std::string &va = a.variant, &vb = b.variant;
  auto in = [](const std::string& v, std::initializer_list<const char*> l) {
    for (auto x : l)
      if (v == x) return true;
    return false;
  };
  bool ok = (va == "StringStock" && in(vb, {"StringProcess", "StringParallelProcess", "StringSink"})) ||
            (vb == "StringStock" && in(va, {"StringProcess", "StringParallelProcess", "StringSource"})) ||
            (va == "BasicEnvironment" && in(vb, {"StringProcess", "StringParallelProcess", "StringSource", "StringSink"}));
  /* continuous arms (core.rs:572-883): same resource type on both sides; Stock -> CombinerM and SplitterN -> Stock
   * take Some(n) */
  for (const char* P : {"F64", "Vector3"}) {
    const std::string p(P);
    if (va.compare(0, p.size(), p) != 0 || vb.compare(0, p.size(), p) != 0) continue;
    if (va.compare(0, 16, "Vector3Container") == 0 || vb.compare(0, 16, "Vector3Container") == 0) continue;
    const std::string x = va.substr(p.size()), y = vb.substr(p.size());
    const bool xc = x.compare(0, 8, "Combiner") == 0, xs = x.compare(0, 8, "Splitter") == 0;
    const bool yc = y.compare(0, 8, "Combiner") == 0, ys = y.compare(0, 8, "Splitter") == 0;
    if (x == "Stock" && (y == "Process" || y == "Sink" || ys)) ok = true;
    if (x == "Stock" && yc) ok = n >= 0 && n < y.back() - '0';
    if (y == "Stock" && (x == "Process" || x == "Source" || xc)) ok = true;
    if (y == "Stock" && xs) ok = n >= 0 && n < x.back() - '0';
  }
  if (va == "BasicEnvironment" && (vb.compare(0, 3, "F64") == 0 || vb.compare(0, 7, "Vector3") == 0) &&
      vb.compare(0, 16, "Vector3Container") != 0 && vb.find("Stock") == std::string::npos)
    ok = true;
  /* container arms (core.rs:923-995) and their environment arms (core.rs:1149-1172: no ParallelProcess arm) */
  const char* CS = "Vector3ContainerStock";
  if (va == CS && in(vb, {"Vector3ContainerProcess", "Vector3ContainerParallelProcess", "Vector3ContainerSink",
                          "Vector3ContainerLoadProcess", "Vector3ContainerUnloadProcess"})) ok = true;
  if (vb == CS && in(va, {"Vector3ContainerProcess", "Vector3ContainerParallelProcess", "Vector3ContainerSource",
                          "Vector3ContainerLoadProcess", "Vector3ContainerUnloadProcess"})) ok = true;
  if (va == "Vector3Stock" && vb == "Vector3ContainerLoadProcess") ok = true;
  if (va == "Vector3ContainerUnloadProcess" && vb == "Vector3Stock") ok = true;
  if (va == "BasicEnvironment" && in(vb, {"Vector3ContainerProcess", "Vector3ContainerSource", "Vector3ContainerSink",
                                          "Vector3ContainerLoadProcess", "Vector3ContainerUnloadProcess"})) ok = true;
  if (!ok)
    throw Error("No component connection defined from " + va + " to " + vb + " (n=" +
                    (n < 0 ? std::string("None") : "Some(" + std::to_string(n) + ")") + ")",
                QK_ERR_UNSUPPORTED_CONNECTION);
  a.conns->push_back(Connection{a.component.d, b.component.d, n, connection_stamp()++});
}

/* ---- loggers (core.rs:1216-1365) --------------------------------------------------------------------------------- */
class BatchedSimulation;
class ComponentLogger {
 public:
  std::string variant, name;
  std::vector<std::shared_ptr<ComponentData>> components;
  ComponentLogger(std::string v, std::string n) : variant(std::move(v)), name(std::move(n)) {}
  static ComponentLogger StringStockLogger(const std::string& name) { return ComponentLogger("StringStockLogger", name); }
  static ComponentLogger StringProcessLogger(const std::string& name) { return ComponentLogger("StringProcessLogger", name); }
  static ComponentLogger BasicEnvironmentLogger(const std::string& name) { return ComponentLogger("BasicEnvironmentLogger", name); }
  /* records of these are read with BatchedSimulation::read_log; their CSV layouts live in the Python mirror (csvlog.py) */
  static ComponentLogger F64StockLogger(const std::string& name) { return ComponentLogger("F64StockLogger", name); }
  static ComponentLogger F64ProcessLogger(const std::string& name) { return ComponentLogger("F64ProcessLogger", name); }
  static ComponentLogger Vector3StockLogger(const std::string& name) { return ComponentLogger("Vector3StockLogger", name); }
  static ComponentLogger Vector3ProcessLogger(const std::string& name) { return ComponentLogger("Vector3ProcessLogger", name); }
  static ComponentLogger Vector3ContainerStockLogger(const std::string& name) { return ComponentLogger("Vector3ContainerStockLogger", name); }
  static ComponentLogger Vector3ContainerProcessLogger(const std::string& name) { return ComponentLogger("Vector3ContainerProcessLogger", name); }
  /* the CSV text Logger::write_csv (core.rs:388-401) produces for one replication */
  std::string csv_text(BatchedSimulation& sim, uint64_t replication = 0) const;
  std::string write_csv(const std::string& dir, BatchedSimulation& sim, uint64_t replication = 0) const {
    const std::string path = dir + "/" + name + ".csv";
    std::ofstream f(path, std::ios::binary);
    f << csv_text(sim, replication);
    return path;
  }
};

inline void connect_logger(ComponentLogger& logger, ComponentModel& c) {
  const std::string& v = c.variant;
  bool ok = (logger.variant == "StringStockLogger" && v == "StringStock") ||
            (logger.variant == "StringProcessLogger" &&
             (v == "StringProcess" || v == "StringParallelProcess" || v == "StringSource" || v == "StringSink")) ||
            (logger.variant == "BasicEnvironmentLogger" && v == "BasicEnvironment");
  for (const char* P : {"F64", "Vector3"}) { /* core.rs:1311-1322 */
    const std::string p(P);
    if (v.compare(0, p.size(), p) != 0 || v.compare(0, 16, "Vector3Container") == 0) continue;
    if (logger.variant == p + "StockLogger" && v == p + "Stock") ok = true;
    if (logger.variant == p + "ProcessLogger" && v != p + "Stock") ok = true;
  }
  if (logger.variant == "Vector3ContainerStockLogger" && v == "Vector3ContainerStock") ok = true;
  if (logger.variant == "Vector3ContainerProcessLogger" && /* core.rs:1331-1334: no Source / Sink arm */
      (v == "Vector3ContainerProcess" || v == "Vector3ContainerParallelProcess" || v == "Vector3ContainerLoadProcess" ||
       v == "Vector3ContainerUnloadProcess"))
    ok = true;
  if (!ok)
    throw Error("No logger connection defined from " + logger.variant + " to " + v + " (n=None)",
                QK_ERR_UNSUPPORTED_CONNECTION);
  logger.components.push_back(c.component.d);
  c.component.d->logged = true;
}

/* ---- scheduled events (core.rs:1367-1401) ------------------------------------------------------------------------ */
struct ScheduledEventConfig {
  std::string kind;
  int value;
  static ScheduledEventConfig SetEnvironmentState(BasicEnvironmentState s) { return ScheduledEventConfig{"SetEnvironmentState", (int)s}; }
};

/* ---- SimInit / Simulation ----------------------------------------------------------------------------------------- */
struct BatchOptions {
  uint64_t n_replications = 1, base_seed = 0, rep_offset = 0;
  int device = 0;
  uint32_t log_capacity = 0, threads_per_block = 0, flags = 0, lanes_per_replication = 0;
  /* more than one GPU: the replications are sharded over `devices` (or over every visible device) by the library's
   * qk_multi_* plane, and stats() returns the globally reduced summary (NCCL inside the library) */
  std::vector<int32_t> devices;
  bool all_devices = false;
  bool multi() const { return all_devices || devices.size() > 1; }
};

class BatchedSimulation {
 public:
  std::vector<ComponentModel> components;
  uint64_t t0;
  BatchOptions opt;

  BatchedSimulation(std::vector<ComponentModel> comps, uint64_t t0_ns, const BatchOptions& o)
      : components(std::move(comps)), t0(t0_ns), opt(o) {
    for (size_t i = 0; i < components.size(); ++i) {
      auto& d = *components[i].component.d;
      if (d.id >= 0 && d.id != (int)i) throw Error("component " + d.element_name + " registered twice", QK_ERR_INVALID_ARG);
      d.id = (int)i;
    }
  }
  BatchedSimulation(const BatchedSimulation&) = delete;
  BatchedSimulation& operator=(const BatchedSimulation&) = delete;
  BatchedSimulation(BatchedSimulation&& o) noexcept
      : components(std::move(o.components)), t0(o.t0), opt(o.opt), model_(o.model_), batch_(o.batch_), multi_(o.multi_),
        ext_(std::move(o.ext_)), tapes_(std::move(o.tapes_)), dist_ids_(std::move(o.dist_ids_)) {
    o.model_ = nullptr;
    o.batch_ = nullptr;
    o.multi_ = nullptr;
  }
  ~BatchedSimulation() {
    if (multi_) qk_multi_destroy(multi_);
    if (batch_) qk_batch_destroy(batch_);
    if (model_) qk_model_destroy(model_);
  }

  /* create_scheduled_event!(&mut sched, &time, &cfg, &addr, &mut df): only before the first step */
  void schedule_env_state(uint64_t t_ns, const ComponentModel& env, BasicEnvironmentState s) {
    if (batch_) throw Error("scheduled events must be created before the first step of a BatchedSimulation", QK_ERR_STATE);
    if (env.variant != "BasicEnvironment")
      throw Error("schedule_event not implemented for types (SetEnvironmentState, " + env.variant + ")", QK_ERR_INVALID_ARG);
    ext_.push_back({t_ns, env.component.d, (uint32_t)s});
  }
  /* pre-drawn variate tape of one Distribution instance: values[replication * len + i] (parity mode) */
  void set_tape(const Distribution& dist, const std::vector<double>& values, uint64_t per_rep_len) {
    if (batch_) throw Error("set_tape must be called before the first step", QK_ERR_STATE);
    if (values.size() != per_rep_len * opt.n_replications) throw Error("tape must hold n_replications * len values", QK_ERR_INVALID_ARG);
    tapes_.push_back({dist.identity.get(), values, per_rep_len});
  }
  void step_until(uint64_t t_ns) {
    materialize();
    check(multi_ ? qk_multi_step_until(multi_, t_ns) : qk_batch_step_until(batch_, t_ns));
  }
  void step_events(uint64_t n) {
    materialize();
    check(multi_ ? qk_multi_step_events(multi_, n) : qk_batch_step_events(batch_, n));
  }
  /* per-component summaries over ALL replications (on several GPUs: reduced inside the library with NCCL) and the
   * whole-job totals */
  std::vector<qk_comp_summary> stats(qk_batch_summary* total = nullptr) {
    materialize();
    std::vector<qk_comp_summary> out(components.size());
    if (multi_) {
      check(qk_multi_stats(multi_, out.data(), (uint32_t)out.size(), total));
    } else {
      check(qk_batch_reduce_stats(batch_, out.data(), (uint32_t)out.size()));
      if (total) check(qk_batch_summary_get(batch_, total));
    }
    return out;
  }
  /* device-timed duration of the last step call (maximum over the GPUs) */
  float last_step_ms() {
    materialize();
    float ms = 0;
    check(multi_ ? qk_multi_last_step_ms(multi_, &ms) : qk_batch_last_step_ms(batch_, &ms, nullptr));
    return ms;
  }
  uint32_t n_shards() {
    materialize();
    uint32_t n = 1;
    if (multi_) check(qk_multi_n_shards(multi_, &n));
    return n;
  }
  qk_batch_summary summary() {
    materialize();
    qk_batch_summary s;
    if (multi_) {
      std::vector<qk_comp_summary> tmp(components.size());
      check(qk_multi_stats(multi_, tmp.data(), (uint32_t)tmp.size(), &s));
    } else {
      check(qk_batch_summary_get(batch_, &s));
    }
    return s;
  }
  std::vector<qk_comp_stats> comp_stats(const ComponentModel& c, uint64_t rep0, uint64_t n) {
    materialize();
    std::vector<qk_comp_stats> out(n);
    check(qk_batch_comp_stats(batch_, rep0, n, (uint32_t)c.component.d->id, out.data()));
    return out;
  }
  /* records of one replication, oldest first; *total = records ever written */
  std::vector<qk_log_record> read_log(uint64_t rep, uint64_t* total = nullptr) {
    materialize();
    uint64_t n = 0, tot = 0;
    check(qk_batch_read_log(batch_, rep, nullptr, 0, &n, &tot));
    const uint64_t kept = tot < opt.log_capacity ? tot : opt.log_capacity;
    std::vector<qk_log_record> buf(kept);
    if (kept) check(qk_batch_read_log(batch_, rep, buf.data(), kept, &n, &tot));
    if (total) *total = tot;
    return buf;
  }
  /* resource of a vector stock / in-flight vector of a vector process-like component */
  Vector3 read_vector(const ComponentModel& c, uint64_t rep) {
    materialize();
    Vector3 v;
    check(qk_batch_read_vector(batch_, rep, (uint32_t)c.component.d->id, v.values));
    return v;
  }
  /* containers held by a container stock (FIFO order) or a container process-like component: (handle, resource) pairs;
   * has_resource == false for a container that carries nothing.  id / capacity are not device state: the id is the
   * item name of the handle, the capacity is the factory's or the initial item's. */
  std::vector<std::pair<uint32_t, Vector3Container>> read_containers(const ComponentModel& c, uint64_t rep) {
    materialize();
    uint64_t n = 0;
    check(qk_batch_read_containers(batch_, rep, (uint32_t)c.component.d->id, nullptr, nullptr, 0, &n));
    std::vector<uint32_t> h(n ? n : 1);
    std::vector<uint64_t> r(3 * (n ? n : 1));
    check(qk_batch_read_containers(batch_, rep, (uint32_t)c.component.d->id, h.data(), r.data(), h.size(), &n));
    std::vector<std::pair<uint32_t, Vector3Container>> out;
    for (uint64_t i = 0; i < n; ++i) {
      Vector3Container k;
      k.has_resource = r[3 * i] != QK_CONTAINER_NONE_BITS;
      if (k.has_resource) std::memcpy(k.resource.values, &r[3 * i], 24);
      out.push_back({h[i], k});
    }
    return out;
  }
  qk_batch* handle() {
    materialize();
    return batch_;
  }

 private:
  struct Ext {
    uint64_t t;
    std::shared_ptr<ComponentData> env;
    uint32_t state;
  };
  struct Tape {
    const int* identity;
    std::vector<double> values;
    uint64_t len;
  };
  qk_model* model_ = nullptr;
  qk_batch* batch_ = nullptr; /* one GPU: the batch ; several GPUs: shard 0 (per-replication reads address shard 0) */
  qk_multi* multi_ = nullptr;
  std::vector<Ext> ext_;
  std::vector<Tape> tapes_;
  std::vector<std::pair<const int*, uint32_t>> dist_ids_;

  /* component objects -> C-ABI rows; the batch is created lazily so that scheduled events created after init()
   * (assembly_line.rs:240-246) still land in the model tables */
  void materialize() {
    if (batch_ || multi_) return;
    check(qk_model_create(&model_));
    for (auto& cm : components) {
      auto& c = *cm.component.d;
      qk_component_desc d;
      std::memset(&d, 0, sizeof(d));
      d.kind = c.container_kind ? c.container_kind : c.kind;
      d.low_capacity = c.low_capacity;
      d.max_capacity = c.max_capacity;
      d.n_initial_items = (c.kind == QK_DISCRETE_STOCK && !c.container_kind) ? (uint32_t)c.resource.size() : 0;
      d.initial_env_state = (uint32_t)c.state;
      if (c.kind == QK_CONTAINER_LOAD_PROCESS || c.kind == QK_CONTAINER_UNLOAD_PROCESS) d.max_capacity = c.max_process_count;
      uint32_t id = 0;
      check(qk_model_add_component(model_, &d, &id));
      if (c.kind >= QK_VECTOR_STOCK && c.kind <= QK_VECTOR_SPLITTER) {
        if (!c.dim) throw Error("vector component " + c.element_name + " was never wrapped in a ComponentModel variant", QK_ERR_INVALID_ARG);
        check(qk_model_set_vector(model_, id, c.dim, c.vlow, c.vmax, c.vinit.values));
        if (c.n_ports) check(qk_model_set_ports(model_, id, c.n_ports, c.split_ratios.data()));
      }
      if (c.container_kind == QK_CONTAINER_STOCK && !c.containers.empty()) {
        std::vector<double> res, caps;
        std::vector<uint8_t> has;
        for (auto& k : c.containers) {
          for (int j = 0; j < 3; ++j) res.push_back(k.resource.values[j]);
          has.push_back(k.has_resource ? 1 : 0);
          caps.push_back(k.capacity);
        }
        check(qk_model_set_initial_containers(model_, id, (uint32_t)has.size(), res.data(), has.data(), caps.data()));
      }
      if (c.container_kind == QK_CONTAINER_SOURCE) {
        uint32_t did = 0;
        check(qk_model_set_container_factory(model_, id, c.cfactory_has_quantity ? &c.cfactory_quantity.desc : nullptr,
                                             c.cfactory_split, c.cfactory_capacity, &did));
        if (c.cfactory_has_quantity) dist_ids_.push_back({c.cfactory_quantity.identity.get(), did});
      }
      const bool process_like = (c.kind >= QK_DISCRETE_PROCESS && c.kind <= QK_DISCRETE_PARALLEL_PROCESS) ||
                                (c.kind >= QK_VECTOR_PROCESS && c.kind <= QK_VECTOR_SPLITTER) ||
                                c.kind == QK_CONTAINER_LOAD_PROCESS || c.kind == QK_CONTAINER_UNLOAD_PROCESS;
      if (process_like) {
        uint32_t did = 0;
        check(qk_model_set_dist(model_, id, 1, &c.process_quantity_distr.desc, &did));
        dist_ids_.push_back({c.process_quantity_distr.identity.get(), did});
        check(qk_model_set_dist(model_, id, 0, &c.process_time_distr.desc, &did));
        dist_ids_.push_back({c.process_time_distr.identity.get(), did});
        for (auto& m : c.delay_modes) {
          uint32_t a = 0, b = 0;
          check(qk_model_add_delay_mode(model_, id, &m.until_delay_distr.desc, &m.until_fix_distr.desc, &a, &b));
          dist_ids_.push_back({m.until_delay_distr.identity.get(), a});
          dist_ids_.push_back({m.until_fix_distr.identity.get(), b});
        }
      }
      if (c.logged) check(qk_model_set_logged(model_, id, 1));
    }
    /* connections in the order connect_components! was called: subscriber order is broadcast order */
    std::vector<Connection> all;
    for (auto& cm : components) all.insert(all.end(), cm.conns->begin(), cm.conns->end());
    for (size_t i = 1; i < all.size(); ++i)
      for (size_t j = i; j > 0 && all[j].stamp < all[j - 1].stamp; --j) std::swap(all[j], all[j - 1]);
    for (auto& k : all) {
      if (k.a->id < 0 || k.b->id < 0) throw Error("connect_components involves a component that was never registered", QK_ERR_INVALID_ARG);
      check(qk_model_connect(model_, (uint32_t)k.a->id, (uint32_t)k.b->id, k.n));
    }
    for (auto& e : ext_) check(qk_model_schedule_env_state(model_, e.t, (uint32_t)e.env->id, e.state));
    if (opt.multi()) {
      qk_multi_config mc;
      std::memset(&mc, 0, sizeof(mc));
      mc.n_reps = opt.n_replications;
      mc.rep_offset = opt.rep_offset;
      mc.base_seed = opt.base_seed;
      mc.devices = opt.all_devices ? nullptr : opt.devices.data();
      mc.n_devices = opt.all_devices ? 0u : (uint32_t)opt.devices.size();
      mc.log_capacity = opt.log_capacity;
      mc.threads_per_block = opt.threads_per_block;
      mc.flags = opt.flags;
      mc.lanes_per_replication = opt.lanes_per_replication;
      check(qk_multi_create(model_, &mc, &multi_));
      for (auto& t : tapes_)
        for (auto& di : dist_ids_)
          if (di.first == t.identity) check(qk_multi_set_tape(multi_, di.second, t.values.data(), t.len));
      check(qk_multi_init(multi_, t0));
      return;
    }
    qk_batch_config cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.lanes_per_replication = opt.lanes_per_replication;
    cfg.n_reps = opt.n_replications;
    cfg.rep_offset = opt.rep_offset;
    cfg.base_seed = opt.base_seed;
    cfg.device = opt.device;
    cfg.log_capacity = opt.log_capacity;
    cfg.threads_per_block = opt.threads_per_block;
    cfg.flags = opt.flags;
    check(qk_batch_create(model_, &cfg, &batch_));
    for (auto& t : tapes_)
      for (auto& di : dist_ids_)
        if (di.first == t.identity) check(qk_batch_set_tape(batch_, di.second, t.values.data(), t.len));
    check(qk_batch_init(batch_, t0));
  }
};

/* stands where nexosim::SimInit stands in the examples */
class BatchedSimInit {
 public:
  std::vector<ComponentModel> components;
  static BatchedSimInit new_() { return BatchedSimInit(); }
  /* returns the simulation, like SimInit::init returns (Simulation, Scheduler) */
  BatchedSimulation init(uint64_t t0_ns, const BatchOptions& opt = BatchOptions()) { return BatchedSimulation(components, t0_ns, opt); }
};
/* register_component!(sim_init, c): registration order = component id = scheduler origin rank - 1 */
inline BatchedSimInit register_component(BatchedSimInit s, const ComponentModel& c) {
  s.components.push_back(c);
  return s;
}
inline void create_scheduled_event(BatchedSimulation& sim, uint64_t t_ns, const ScheduledEventConfig& cfg, const ComponentModel& addr) {
  if (cfg.kind != "SetEnvironmentState")
    throw Error("schedule_event not implemented for types (" + cfg.kind + ", " + addr.variant + ")", QK_ERR_INVALID_ARG);
  sim.schedule_env_state(t_ns, addr, (BasicEnvironmentState)cfg.value);
}

/* ---- CSV (discrete.rs:295-318, 608-633; core.rs:306-325) ---------------------------------------------------------- */
namespace detail {
inline std::string two(int v) {
  char b[8];
  std::snprintf(b, sizeof(b), "%02d", v);
  return b;
}
inline std::string date_time(uint64_t t_ns, uint32_t* nanos) {
  const uint64_t secs = t_ns / 1000000000ull;
  *nanos = (uint32_t)(t_ns % 1000000000ull);
  /* civil-from-days (proleptic Gregorian), days since 1970-01-01 */
  int64_t z = (int64_t)(secs / 86400) + 719468;
  const int64_t era = z / 146097;
  const unsigned doe = (unsigned)(z - era * 146097);
  const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  int64_t y = (int64_t)yoe + era * 400;
  const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const unsigned mp = (5 * doy + 2) / 153;
  const unsigned d = doy - (153 * mp + 2) / 5 + 1;
  const unsigned m = mp < 10 ? mp + 3 : mp - 9;
  if (m <= 2) y += 1;
  const unsigned sod = (unsigned)(secs % 86400);
  char b[64];
  std::snprintf(b, sizeof(b), "%04lld-%02u-%02u %02u:%02u:%02u", (long long)y, m, d, sod / 3600, (sod / 60) % 60, sod % 60);
  return b;
}
/* MonotonicTime::to_chrono_date_time(0).unwrap().to_string(): "%Y-%m-%d %H:%M:%S%.f UTC" with 0/3/6/9 digits */
inline std::string chrono_utc(uint64_t t_ns) {
  uint32_t nanos;
  std::string s = date_time(t_ns, &nanos);
  char b[16];
  if (nanos) {
    if (nanos % 1000000 == 0) std::snprintf(b, sizeof(b), ".%03u", nanos / 1000000);
    else if (nanos % 1000 == 0) std::snprintf(b, sizeof(b), ".%06u", nanos / 1000);
    else std::snprintf(b, sizeof(b), ".%09u", nanos);
    s += b;
  }
  return s + " UTC";
}
/* tai-time's own Display (core.rs:277) */
inline std::string tai_display(uint64_t t_ns) {
  uint32_t nanos;
  std::string s = date_time(t_ns, &nanos);
  if (nanos) {
    char b[16];
    std::snprintf(b, sizeof(b), ".%09u", nanos);
    s += b;
  }
  return s;
}
inline std::string csv_field(const std::string& s) {
  if (s.find_first_of(",\"\r\n") == std::string::npos) return s;
  std::string o = "\"";
  for (char ch : s) {
    if (ch == '"') o += '"';
    o += ch;
  }
  return o + "\"";
}
inline std::string json_string(const std::string& s) {
  std::string o = "\"";
  for (char ch : s) {
    if (ch == '"') o += "\\\"";
    else if (ch == '\\') o += "\\\\";
    else if (ch == '\n') o += "\\n";
    else if (ch == '\r') o += "\\r";
    else if (ch == '\t') o += "\\t";
    else o += ch;
  }
  return o + "\"";
}
inline const char* reason_text(uint32_t r) {
  static const char* T[] = {"", "Upstream is empty", "Upstream is not connected", "Downstream is full", "Downstream is not connected",
                            "Upstream did not provide resource", "Stopped by environment", "Resumed by environment", "Upstream is full",
                            "At least one upstream is empty", "No upstreams are connected", "At least one downstream is full",
                            "No downstreams are connected", "No remaining process slots", "Downstream container stock is full",
                            "Downstream resource stock is full"};
  return r < sizeof(T) / sizeof(T[0]) ? T[r] : "";
}
/* f64 -> text the way serde_json / ryu print it: shortest round-trip digits, "1.0" not "1", exponent form outside
 * [1e-5, 1e16); non-finite values are JSON null */
inline std::string json_f64(double x) {
  if (!std::isfinite(x)) return "null";
  if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
  char buf[64];
  const auto res = std::to_chars(buf, buf + sizeof(buf), std::fabs(x), std::chars_format::scientific);
  std::string sci(buf, res.ptr); /* d[.ddd]e[+-]XX, shortest digits */
  const size_t e = sci.find('e');
  std::string digits;
  for (size_t i = 0; i < e; ++i)
    if (sci[i] != '.') digits += sci[i];
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int kk = std::atoi(sci.c_str() + e + 1) + 1; /* value = 0.d1d2... * 10^kk */
  const int n = (int)digits.size();
  std::string out = x < 0 ? "-" : "";
  if (0 < kk && kk <= 16 && n <= kk) out += digits + std::string((size_t)(kk - n), '0') + ".0";
  else if (0 < kk && kk <= 16) out += digits.substr(0, (size_t)kk) + "." + digits.substr((size_t)kk);
  else if (-5 < kk && kk <= 0) out += "0." + std::string((size_t)(-kk), '0') + digits;
  else out += (n == 1 ? digits : digits.substr(0, 1) + "." + digits.substr(1)) + "e" + std::to_string(kk - 1);
  return out;
}
}  // namespace detail

inline std::string ComponentLogger::csv_text(BatchedSimulation& sim, uint64_t replication) const {
  uint64_t total = 0;
  const std::vector<qk_log_record> recs = sim.read_log(replication, &total);
  if (total > recs.size()) throw Error("log ring kept fewer records than were written: raise log_capacity for CSV export", QK_ERR_BUFFER_TOO_SMALL);
  std::vector<std::shared_ptr<ComponentData>> by_id;
  std::vector<const ComponentData*> sources;
  std::vector<std::string> initial_items; /* initial stock contents, numbered in registration order of their stocks */
  std::vector<double> initial_caps;
  for (auto& cm : sim.components) {
    const ComponentData& d = *cm.component.d;
    by_id.push_back(cm.component.d);
    if (d.kind == QK_DISCRETE_SOURCE) sources.push_back(&d);
    if (d.kind == QK_DISCRETE_STOCK && d.container_kind)
      for (auto& k : d.containers) {
        initial_items.push_back(k.id);
        initial_caps.push_back(k.capacity);
      }
    else if (d.kind == QK_DISCRETE_STOCK)
      for (auto& k : d.resource) {
        initial_items.push_back(k);
        initial_caps.push_back(0.0);
      }
  }
  auto event_id = [&](uint32_t comp, uint32_t seq) -> std::string {
    if (comp == QK_SRC_INIT) return "INIT_000000";
    if (comp == QK_SRC_SCH) return "SCH_000000";
    char b[16];
    std::snprintf(b, sizeof(b), "_%06u", seq);
    return by_id[comp]->element_code + b;
  };
  auto item_name = [&](uint64_t h) -> std::string {
    const uint32_t ord = QK_ITEM_SOURCE(h), idx = QK_ITEM_INDEX(h);
    if (ord == 63) return initial_items[idx];
    const ComponentData& f = *sources[ord];
    if (!f.container_kind) return f.item_factory.item_name(idx);
    std::string n = std::to_string(f.cfactory_next_index + idx); /* "{prefix}{index:0width$}" vector_container.rs:138 */
    if (n.size() < f.cfactory_num_digits) n = std::string(f.cfactory_num_digits - n.size(), '0') + n;
    return f.cfactory_prefix + n;
  };
  /* serde_json::to_string(item): a String, or a Vector3Container (Serialize impl vector_container.rs:113-124) whose
   * resource arrived in the continuation record `cont` that follows the main record */
  auto item_json = [&](const qk_log_record& r, const qk_log_record* cont) -> std::string {
    if (!by_id[r.comp]->container_kind) return detail::json_string(item_name(r.payload));
    const uint32_t ord = QK_ITEM_SOURCE(r.payload), idx = QK_ITEM_INDEX(r.payload);
    const double cap = ord == 63 ? initial_caps[idx] : sources[ord]->cfactory_capacity;
    std::string res = "null";
    if (cont && cont->time_ns != QK_CONTAINER_NONE_BITS) {
      double v[3];
      std::memcpy(&v[0], &cont->time_ns, 8);
      std::memcpy(&v[1], &cont->src_comp, 8); /* src_comp | aux | src_seq overlay the second double */
      std::memcpy(&v[2], &cont->payload, 8);
      res = "{\"x\":" + detail::json_f64(v[0]) + ",\"y\":" + detail::json_f64(v[1]) + ",\"z\":" + detail::json_f64(v[2]) + "}";
    }
    return "{\"id\":" + detail::json_string(item_name(r.payload)) + ",\"resource\":" + res + ",\"capacity\":" +
           detail::json_f64(cap) + "}";
  };
  auto member = [&](uint32_t comp) {
    for (auto& c : components)
      if (c.get() == by_id[comp].get()) return true;
    return false;
  };
  std::ostringstream out;
  bool any = false;
  const bool env_logger = variant == "BasicEnvironmentLogger";
  if (variant.compare(0, 3, "F64") == 0 || (variant.compare(0, 7, "Vector3") == 0 && variant.compare(0, 16, "Vector3Container") != 0))
    throw Error("CSV export of the continuous loggers is in the Python mirror (quokkasim_b200/csvlog.py)", QK_ERR_INVALID_ARG);
  for (size_t ri = 0; ri < recs.size(); ++ri) {
    const qk_log_record& r = recs[ri];
    if (r.kind == QK_LOG_VEC_DATA || !member(r.comp)) continue;
    const qk_log_record* cont = (ri + 1 < recs.size() && recs[ri + 1].kind == QK_LOG_VEC_DATA && recs[ri + 1].comp == r.comp &&
                                 recs[ri + 1].seq == r.seq) ? &recs[ri + 1] : nullptr;
    const ComponentData& c = *by_id[r.comp];
    if (!any) {
      any = true; /* csv::Writer writes the header with the first record: an empty log is an empty file */
      if (env_logger) out << "time,event_id,source_event_id,element_name,element_type,event\n";
      else if (variant == "StringStockLogger" || variant == "Vector3ContainerStockLogger")
        out << "time,event_id,source_event_id,element_name,element_type,log_type,item,reason\n";
      else out << "time,event_id,source_event_id,element_name,element_type,event_type,item,reason\n";
    }
    std::string type, item, reason;
    switch (r.kind) {
      case QK_LOG_ENV_STATE:
        out << detail::tai_display(r.time_ns) << ',' << event_id(r.comp, r.seq) << ',' << event_id(r.src_comp, r.src_seq) << ','
            << detail::csv_field(c.element_name) << ",BasicEnvironment," << (r.reason == 1 ? "Stopped" : "Normal") << '\n';
        continue;
      case QK_LOG_STOCK_ADD: type = "Add", item = item_json(r, cont); break;
      case QK_LOG_STOCK_REMOVE:
        type = "Remove", item = r.payload == QK_ITEM_NONE ? "null" : item_json(r, cont);
        break;
      case QK_LOG_STOCK_STATE_CHANGE: {
        static const char* S[] = {"Empty", "Normal", "Full"};
        type = "StateChange";
        item = std::string("{\"") + S[r.reason] + "\":{\"occupied\":" + std::to_string(r.payload >> 32) +
               ",\"empty\":" + std::to_string(r.payload & 0xFFFFFFFFull) + "}}";
        break;
      }
      case QK_LOG_PROCESS_START: type = "ProcessStart", item = item_json(r, cont); break;
      case QK_LOG_PROCESS_FINISH: type = "ProcessFinish", item = item_json(r, cont); break;
      case QK_LOG_DELAY_START: type = "DelayStart", item = c.delay_modes[r.payload].name; break;
      case QK_LOG_DELAY_END: type = "DelayEnd", item = c.delay_modes[r.payload].name; break;
      case QK_LOG_PROCESS_CONTINUE: type = "ProcessContinue", reason = detail::reason_text(r.reason); break;
      case QK_LOG_PROCESS_NONSTART: type = "ProcessNonStart", reason = detail::reason_text(r.reason); break;
      case QK_LOG_PROCESS_STOPPED: type = "ProcessStopped", reason = detail::reason_text(r.reason); break;
      case QK_LOG_WITHDRAW_REQUEST: type = "WithdrawRequest"; break;
      default: continue;
    }
    out << detail::chrono_utc(r.time_ns) << ',' << event_id(r.comp, r.seq) << ',' << event_id(r.src_comp, r.src_seq) << ','
        << detail::csv_field(c.element_name) << ',' << detail::csv_field(c.element_type) << ',' << type << ','
        << detail::csv_field(item) << ',' << detail::csv_field(reason) << '\n';
  }
  return out.str();
}

}  // namespace quokka
#endif /* QUOKKA_B200_HPP */
