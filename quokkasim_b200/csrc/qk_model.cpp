/*
 * qk_model.cpp -- model description tables behind the C ABI and their lowering to the device layout.
 *
 * Replaces, for the batched path, what the reference's generated enums do on the host:
 * ComponentModel::register_component / connect_components / connect_logger and
 * ScheduledEventConfig::schedule_event (core.rs:565-1401).  Instead of wiring NeXosim ports with
 * closures, every call appends a row; qk_lower() turns the rows into the flat tables the kernels read.
 */
#include "qk_model.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

static thread_local char g_err[512] = "";

void qk_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int qk_is_process_like(uint32_t k) {
  return k == QK_DISCRETE_PROCESS || k == QK_DISCRETE_SOURCE || k == QK_DISCRETE_SINK ||
         k == QK_DISCRETE_PARALLEL_PROCESS || (k >= QK_VECTOR_PROCESS && k <= QK_VECTOR_SPLITTER) ||
         (k >= QK_CONTAINER_PROCESS && k <= QK_CONTAINER_UNLOAD_PROCESS);
}
static int is_vector_kind(uint32_t k) { return k >= QK_VECTOR_STOCK && k <= QK_VECTOR_SPLITTER; }
static int is_container_kind(uint32_t k) { return k >= QK_CONTAINER_STOCK && k <= QK_CONTAINER_UNLOAD_PROCESS; }
static int is_item_stock(uint32_t k) { return k == QK_DISCRETE_STOCK || k == QK_CONTAINER_STOCK; }
static int is_multi_server(uint32_t k) {
  return k == QK_DISCRETE_PARALLEL_PROCESS || k == QK_CONTAINER_PARALLEL_PROCESS || k == QK_CONTAINER_LOAD_PROCESS ||
         k == QK_CONTAINER_UNLOAD_PROCESS;
}
/* the discrete kind a container component is an instantiation of (core.rs:549-553); Load / Unload have their own */
static uint32_t device_kind(uint32_t k) {
  switch (k) {
    case QK_CONTAINER_STOCK: return QK_DISCRETE_STOCK;
    case QK_CONTAINER_PROCESS: return QK_DISCRETE_PROCESS;
    case QK_CONTAINER_SOURCE: return QK_DISCRETE_SOURCE;
    case QK_CONTAINER_SINK: return QK_DISCRETE_SINK;
    case QK_CONTAINER_PARALLEL_PROCESS: return QK_DISCRETE_PARALLEL_PROCESS;
    default: return k;
  }
}

static const char* kind_name(uint32_t k) {
  switch (k) {
    case QK_DISCRETE_STOCK: return "StringStock";
    case QK_DISCRETE_PROCESS: return "StringProcess";
    case QK_DISCRETE_SOURCE: return "StringSource";
    case QK_DISCRETE_SINK: return "StringSink";
    case QK_DISCRETE_PARALLEL_PROCESS: return "StringParallelProcess";
    case QK_ENVIRONMENT: return "BasicEnvironment";
    case QK_VECTOR_STOCK: return "VectorStock";
    case QK_VECTOR_PROCESS: return "VectorProcess";
    case QK_VECTOR_SOURCE: return "VectorSource";
    case QK_VECTOR_SINK: return "VectorSink";
    case QK_VECTOR_COMBINER: return "VectorCombiner";
    case QK_VECTOR_SPLITTER: return "VectorSplitter";
    case QK_CONTAINER_STOCK: return "Vector3ContainerStock";
    case QK_CONTAINER_PROCESS: return "Vector3ContainerProcess";
    case QK_CONTAINER_SOURCE: return "Vector3ContainerSource";
    case QK_CONTAINER_SINK: return "Vector3ContainerSink";
    case QK_CONTAINER_PARALLEL_PROCESS: return "Vector3ContainerParallelProcess";
    case QK_CONTAINER_LOAD_PROCESS: return "Vector3ContainerLoadProcess";
    case QK_CONTAINER_UNLOAD_PROCESS: return "Vector3ContainerUnloadProcess";
    default: return "?";
  }
}

extern "C" {

const char* qk_last_error(void) { return g_err; }
int qk_abi_version(void) { return QK_ABI_VERSION; }

/* what this library was built from: quokkasim_b200/build.py hashes csrc/, the public header and the compiler flags and
 * passes the result in; build.py finds the marker in the file to decide whether the library in the tree is current */
#ifndef QK_BUILD_ID
#define QK_BUILD_ID "unknown"
#endif
static const char qk_build_marker[] = "QKBUILDID:" QK_BUILD_ID;
const char* qk_build_id(void) { return qk_build_marker + 10; }

int qk_model_create(qk_model** out) {
  if (!out) {
    qk_set_error("qk_model_create: out is NULL");
    return QK_ERR_INVALID_ARG;
  }
  qk_model* m = new qk_model();
  m->n_sources = 0;
  m->n_initial_total = 0;
  *out = m;
  return QK_OK;
}

void qk_model_destroy(qk_model* m) { delete m; }

int qk_model_add_component(qk_model* m, const qk_component_desc* d, uint32_t* out_id) {
  if (!m || !d) {
    qk_set_error("qk_model_add_component: NULL argument");
    return QK_ERR_INVALID_ARG;
  }
  if (!((d->kind >= QK_DISCRETE_STOCK && d->kind <= QK_ENVIRONMENT) || is_vector_kind(d->kind) ||
        is_container_kind(d->kind))) {
    qk_set_error("qk_model_add_component: unknown component kind %u", d->kind);
    return QK_ERR_INVALID_ARG;
  }
  if (m->comps.size() >= QK_MAX_COMPONENTS) {
    qk_set_error("qk_model_add_component: more than %d components", QK_MAX_COMPONENTS);
    return QK_ERR_CAPACITY;
  }
  QkHostComp c;
  c.d = *d;
  c.up = c.down = c.env = -1;
  c.dist_time = c.dist_qty = -1;
  c.logged = false;
  c.src_ord = -1;
  c.dim = is_vector_kind(d->kind) ? 1 : (is_container_kind(d->kind) ? 3 : 0);
  c.dist_cqty = -1;
  c.csplit[0] = c.csplit[1] = c.csplit[2] = 0.0;
  c.ccapacity = 0.0;
  if (d->kind == QK_CONTAINER_STOCK && d->n_initial_items) {
    qk_set_error("qk_model_add_component: a container stock takes its initial contents from "
                 "qk_model_set_initial_containers (n_initial_items must be 0)");
    return QK_ERR_INVALID_ARG;
  }
  if ((d->kind == QK_CONTAINER_LOAD_PROCESS || d->kind == QK_CONTAINER_UNLOAD_PROCESS) &&
      (d->max_capacity == 0 || d->max_capacity > 4096)) {
    qk_set_error("qk_model_add_component: max_process_count (max_capacity) of a load/unload process must be 1..4096");
    return QK_ERR_INVALID_ARG;
  }
  c.n_ports = (d->kind == QK_VECTOR_COMBINER || d->kind == QK_VECTOR_SPLITTER) ? 1 : 0;
  c.vlow = c.vmax = 0.0; /* VectorStock defaults (vector.rs:70-71) */
  for (int k = 0; k < QK_VEC_MAX_DIM; ++k) c.vinit[k] = 0.0;
  c.tmask = 1;
  for (int k = 0; k < 5; ++k) {
    c.ratios[k] = 0.0;
    c.ups[k] = c.downs[k] = -1;
  }
  c.initial_base = 0; /* assigned at lowering: initial items are numbered in registration order of their stocks */
  if (d->kind == QK_DISCRETE_SOURCE || d->kind == QK_CONTAINER_SOURCE) {
    if (m->n_sources >= 63) {
      qk_set_error("qk_model_add_component: at most 63 DiscreteSource components (item handle ordinal)");
      return QK_ERR_CAPACITY;
    }
    c.src_ord = (int)m->n_sources++;
  }
  if (qk_is_process_like(d->kind)) {
    /* Distribution::default() == Constant(1.) (common.rs:181-185) */
    qk_dist_desc one;
    memset(&one, 0, sizeof(one));
    one.kind = QK_DIST_CONSTANT;
    one.p[0] = 1.0;
    c.dist_time = (int)m->dists.size();
    m->dists.push_back(one);
    c.dist_qty = (int)m->dists.size();
    m->dists.push_back(one);
  }
  m->comps.push_back(c);
  if (out_id) *out_id = (uint32_t)m->comps.size() - 1;
  return QK_OK;
}

static int check_dist(const qk_dist_desc* d) {
  if (!d) return 0;
  switch (d->kind) {
    case QK_DIST_CONSTANT:
    case QK_DIST_UNIFORM:
    case QK_DIST_NORMAL:
    case QK_DIST_EXPONENTIAL:
      return 1;
    case QK_DIST_TRIANGULAR: /* rand_distr Triangular::new: max >= mode >= min, max > min */
      return d->p[1] >= d->p[2] && d->p[2] >= d->p[0] && d->p[1] > d->p[0];
    case QK_DIST_TRUNCNORMAL: /* common.rs:108-112 */
      return d->p[2] < d->p[3];
    default:
      return 0;
  }
}

int qk_model_set_dist(qk_model* m, uint32_t comp, uint32_t which, const qk_dist_desc* d, uint32_t* out_id) {
  if (!m || comp >= m->comps.size() || !qk_is_process_like(m->comps[comp].d.kind) || which > 1) {
    qk_set_error("qk_model_set_dist: component %u has no distribution slot %u", comp, which);
    return QK_ERR_INVALID_ARG;
  }
  if (!check_dist(d)) {
    qk_set_error("qk_model_set_dist: invalid distribution parameters");
    return QK_ERR_INVALID_ARG;
  }
  int id = which == 0 ? m->comps[comp].dist_time : m->comps[comp].dist_qty;
  m->dists[(size_t)id] = *d;
  if (out_id) *out_id = (uint32_t)id;
  return QK_OK;
}

int qk_model_add_delay_mode(qk_model* m, uint32_t comp, const qk_dist_desc* ud, const qk_dist_desc* uf,
                            uint32_t* out_ud, uint32_t* out_uf) {
  if (!m || comp >= m->comps.size() || !qk_is_process_like(m->comps[comp].d.kind)) {
    qk_set_error("qk_model_add_delay_mode: component %u cannot carry delay modes", comp);
    return QK_ERR_INVALID_ARG;
  }
  if (!check_dist(ud) || !check_dist(uf)) {
    qk_set_error("qk_model_add_delay_mode: invalid distribution parameters");
    return QK_ERR_INVALID_ARG;
  }
  if (m->comps[comp].dms.size() >= QK_MAX_DELAY_MODES) {
    qk_set_error("qk_model_add_delay_mode: more than %d delay modes on one component", QK_MAX_DELAY_MODES);
    return QK_ERR_CAPACITY;
  }
  DevDelayMode dm;
  dm.until_delay = (uint16_t)m->dists.size();
  m->dists.push_back(*ud);
  dm.until_fix = (uint16_t)m->dists.size();
  m->dists.push_back(*uf);
  m->comps[comp].dms.push_back(dm);
  if (out_ud) *out_ud = dm.until_delay;
  if (out_uf) *out_uf = dm.until_fix;
  return QK_OK;
}

/* connect_components (core.rs:565-1178).  Every supported (a,b) pair wires the same three
 * connections as the reference arm it stands for; unsupported pairs return the reference's
 * "No component connection defined" error (discrete_queue.rs:14-20). */
int qk_model_connect(qk_model* m, uint32_t a, uint32_t b, int32_t port_n) {
  if (!m || a >= m->comps.size() || b >= m->comps.size()) {
    qk_set_error("qk_model_connect: component id out of range");
    return QK_ERR_INVALID_ARG;
  }
  QkHostComp& A = m->comps[a];
  QkHostComp& B = m->comps[b];
  uint32_t ka = A.d.kind, kb = B.d.kind;
  if (ka == QK_DISCRETE_STOCK &&
      (kb == QK_DISCRETE_PROCESS || kb == QK_DISCRETE_SINK || kb == QK_DISCRETE_PARALLEL_PROCESS)) {
    /* core.rs:886-891, 898-903, 916-921: state_emitter->update_state, req_upstream, withdraw_upstream */
    if (B.up >= 0) {
      qk_set_error("qk_model_connect: component %u already has an upstream stock (a second connection on "
                   "req_upstream/withdraw_upstream is not supported)", b);
      return QK_ERR_PORT_IN_USE;
    }
    A.subs.push_back((int)b);
    B.up = (int)a;
    return QK_OK;
  }
  if (kb == QK_DISCRETE_STOCK &&
      (ka == QK_DISCRETE_PROCESS || ka == QK_DISCRETE_SOURCE || ka == QK_DISCRETE_PARALLEL_PROCESS)) {
    /* core.rs:892-897, 904-915: state_emitter->update_state, req_downstream, push_downstream */
    if (A.down >= 0) {
      qk_set_error("qk_model_connect: component %u already has a downstream stock", a);
      return QK_ERR_PORT_IN_USE;
    }
    B.subs.push_back((int)a);
    A.down = (int)b;
    return QK_OK;
  }
  /* container family (core.rs:923-995) */
  if (ka == QK_CONTAINER_STOCK &&
      (kb == QK_CONTAINER_PROCESS || kb == QK_CONTAINER_PARALLEL_PROCESS || kb == QK_CONTAINER_SINK ||
       kb == QK_CONTAINER_LOAD_PROCESS || kb == QK_CONTAINER_UNLOAD_PROCESS)) {
    if (B.up >= 0) {
      qk_set_error("qk_model_connect: component %u already has an upstream container stock", b);
      return QK_ERR_PORT_IN_USE;
    }
    A.subs.push_back((int)b);
    B.up = (int)a;
    return QK_OK;
  }
  if (kb == QK_CONTAINER_STOCK &&
      (ka == QK_CONTAINER_PROCESS || ka == QK_CONTAINER_PARALLEL_PROCESS || ka == QK_CONTAINER_SOURCE ||
       ka == QK_CONTAINER_LOAD_PROCESS || ka == QK_CONTAINER_UNLOAD_PROCESS)) {
    if (A.down >= 0) {
      qk_set_error("qk_model_connect: component %u already has a downstream container stock", a);
      return QK_ERR_PORT_IN_USE;
    }
    B.subs.push_back((int)a);
    A.down = (int)b;
    return QK_OK;
  }
  if (ka == QK_VECTOR_STOCK && A.dim == 3 && A.tmask == 7u && kb == QK_CONTAINER_LOAD_PROCESS) { /* core.rs:960-965 */
    if (B.ups[0] >= 0) {
      qk_set_error("qk_model_connect: load process %u already has an upstream resource stock", b);
      return QK_ERR_PORT_IN_USE;
    }
    A.subs.push_back((int)b);
    B.ups[0] = (int)a;
    return QK_OK;
  }
  if (kb == QK_VECTOR_STOCK && B.dim == 3 && B.tmask == 7u && ka == QK_CONTAINER_UNLOAD_PROCESS) { /* core.rs:990-995 */
    if (A.downs[0] >= 0) {
      qk_set_error("qk_model_connect: unload process %u already has a downstream resource stock", a);
      return QK_ERR_PORT_IN_USE;
    }
    B.subs.push_back((int)a);
    A.downs[0] = (int)b;
    return QK_OK;
  }
  /* core.rs:1129-1172; there is no (BasicEnvironment, Vector3ContainerParallelProcess) arm */
  if (ka == QK_ENVIRONMENT && qk_is_process_like(kb) && kb != QK_CONTAINER_PARALLEL_PROCESS) {
    /* emit_change->update_state, req_environment */
    if (B.env >= 0) {
      qk_set_error("qk_model_connect: component %u already has an environment", b);
      return QK_ERR_PORT_IN_USE;
    }
    A.subs.push_back((int)b);
    B.env = (int)a;
    return QK_OK;
  }
  /* vector arms (core.rs:572-616 for f64, same shape for Vector3; combiner/splitter arms need Some(n)) */
  if (ka == QK_VECTOR_STOCK && A.dim == B.dim && A.tmask == B.tmask) {
    if (kb == QK_VECTOR_PROCESS || kb == QK_VECTOR_SINK || kb == QK_VECTOR_SPLITTER) {
      if (B.up >= 0) {
        qk_set_error("qk_model_connect: component %u already has an upstream stock", b);
        return QK_ERR_PORT_IN_USE;
      }
      A.subs.push_back((int)b);
      B.up = (int)a;
      return QK_OK;
    }
    if (kb == QK_VECTOR_COMBINER && port_n >= 0 && (uint32_t)port_n < B.n_ports) {
      if (B.ups[port_n] >= 0) {
        qk_set_error("qk_model_connect: combiner %u upstream port %d is already connected", b, port_n);
        return QK_ERR_PORT_IN_USE;
      }
      A.subs.push_back((int)b);
      B.ups[port_n] = (int)a;
      return QK_OK;
    }
  }
  if (kb == QK_VECTOR_STOCK && A.dim == B.dim && A.tmask == B.tmask) {
    if (ka == QK_VECTOR_PROCESS || ka == QK_VECTOR_SOURCE || ka == QK_VECTOR_COMBINER) {
      if (A.down >= 0) {
        qk_set_error("qk_model_connect: component %u already has a downstream stock", a);
        return QK_ERR_PORT_IN_USE;
      }
      B.subs.push_back((int)a);
      A.down = (int)b;
      return QK_OK;
    }
    if (ka == QK_VECTOR_SPLITTER && port_n >= 0 && (uint32_t)port_n < A.n_ports) {
      if (A.downs[port_n] >= 0) {
        qk_set_error("qk_model_connect: splitter %u downstream port %d is already connected", a, port_n);
        return QK_ERR_PORT_IN_USE;
      }
      B.subs.push_back((int)a);
      A.downs[port_n] = (int)b;
      return QK_OK;
    }
  }
  if (port_n >= 0)
    qk_set_error("No component connection defined from %s to %s (n=Some(%d))", kind_name(ka), kind_name(kb), port_n);
  else
    qk_set_error("No component connection defined from %s to %s (n=None)", kind_name(ka), kind_name(kb));
  return QK_ERR_UNSUPPORTED_CONNECTION;
}

int qk_model_set_logged(qk_model* m, uint32_t comp, int enabled) {
  if (!m || comp >= m->comps.size()) {
    qk_set_error("qk_model_set_logged: component id out of range");
    return QK_ERR_INVALID_ARG;
  }
  m->comps[comp].logged = enabled != 0;
  return QK_OK;
}

int qk_model_schedule_env_state(qk_model* m, uint64_t t_ns, uint32_t env, uint32_t state) {
  if (!m || env >= m->comps.size() || m->comps[env].d.kind != QK_ENVIRONMENT || state > 1) {
    /* core.rs:1397-1399 */
    qk_set_error("schedule_event not implemented for types (SetEnvironmentState, %s)",
                 (m && env < m->comps.size()) ? kind_name(m->comps[env].d.kind) : "?");
    return QK_ERR_INVALID_ARG;
  }
  DevExt e;
  memset(&e, 0, sizeof(e));
  e.t = t_ns;
  e.type = 0;
  e.target = env;
  e.state = state;
  m->ext.push_back(e);
  return QK_OK;
}

int qk_model_set_vector_n(qk_model* m, uint32_t comp, uint32_t dim, uint32_t total_mask, double low, double max,
                          const double* init) {
  if (!m || comp >= m->comps.size() || !is_vector_kind(m->comps[comp].d.kind) || dim < 1 || dim > QK_VEC_MAX_DIM) {
    qk_set_error("qk_model_set_vector_n: component %u is not a vector component or dim %u is not in 1..%d", comp, dim,
                 QK_VEC_MAX_DIM);
    return QK_ERR_INVALID_ARG;
  }
  if (dim == 1) total_mask = 1; /* f64: the total is the value */
  if (total_mask == 0 || (total_mask >> dim) != 0) {
    qk_set_error("qk_model_set_vector_n: total mask 0x%x names no field, or a field beyond dim %u", total_mask, dim);
    return QK_ERR_INVALID_ARG;
  }
  QkHostComp& c = m->comps[comp];
  c.dim = dim;
  c.tmask = total_mask;
  c.vlow = low;
  c.vmax = max;
  for (uint32_t k = 0; k < QK_VEC_MAX_DIM; ++k) c.vinit[k] = (init && k < dim) ? init[k] : 0.0;
  return QK_OK;
}

int qk_model_set_vector(qk_model* m, uint32_t comp, uint32_t dim, double low, double max, const double* init) {
  if (!(dim == 1 || dim == 3)) {
    qk_set_error("qk_model_set_vector: component %u is not a vector component or dim %u is not 1 or 3", comp, dim);
    return QK_ERR_INVALID_ARG;
  }
  return qk_model_set_vector_n(m, comp, dim, dim == 1 ? 1u : 7u, low, max, init);
}

int qk_model_set_ports(qk_model* m, uint32_t comp, uint32_t n_ports, const double* ratios) {
  if (!m || comp >= m->comps.size() || n_ports < 1 || n_ports > 5 ||
      !(m->comps[comp].d.kind == QK_VECTOR_COMBINER || m->comps[comp].d.kind == QK_VECTOR_SPLITTER)) {
    qk_set_error("qk_model_set_ports: component %u is not a combiner/splitter or n_ports %u not in 1..5", comp, n_ports);
    return QK_ERR_INVALID_ARG;
  }
  QkHostComp& c = m->comps[comp];
  c.n_ports = n_ports;
  for (uint32_t k = 0; k < 5; ++k) c.ratios[k] = (ratios && k < n_ports) ? ratios[k] : 0.0;
  return QK_OK;
}

int qk_model_schedule_low_capacity(qk_model* m, uint64_t t_ns, uint32_t stock, double value) {
  if (!m || stock >= m->comps.size() || m->comps[stock].d.kind != QK_VECTOR_STOCK || m->comps[stock].dim != 1) {
    /* core.rs:1397-1399: only (SetLowCapacity, F64Stock) has an arm */
    qk_set_error("schedule_event not implemented for types (SetLowCapacity, %s)",
                 (m && stock < m->comps.size()) ? kind_name(m->comps[stock].d.kind) : "?");
    return QK_ERR_INVALID_ARG;
  }
  DevExt e;
  memset(&e, 0, sizeof(e));
  e.t = t_ns;
  e.type = 1;
  e.target = stock;
  e.value = value;
  m->ext.push_back(e);
  return QK_OK;
}

/* create_scheduled_event!(SetProcessQuantity(cfg)) on an F64Process (core.rs:1387-1391): df.create(cfg) happens at
 * schedule time (the caller passes the resulting instance), with_process_quantity_distr_inplace + update_state at t */
int qk_model_schedule_process_quantity(qk_model* m, uint64_t t_ns, uint32_t process, const qk_dist_desc* dist,
                                       uint32_t* out_dist_id) {
  if (!m || process >= m->comps.size() || m->comps[process].d.kind != QK_VECTOR_PROCESS || m->comps[process].dim != 1) {
    /* core.rs:1397-1399: only (SetProcessQuantity, F64Process) has an arm */
    qk_set_error("schedule_event not implemented for types (SetProcessQuantity, %s)",
                 (m && process < m->comps.size()) ? kind_name(m->comps[process].d.kind) : "?");
    return QK_ERR_INVALID_ARG;
  }
  if (!dist || !check_dist(dist)) {
    qk_set_error("qk_model_schedule_process_quantity: invalid distribution parameters");
    return QK_ERR_INVALID_ARG;
  }
  DevExt e;
  memset(&e, 0, sizeof(e));
  e.t = t_ns;
  e.type = 2;
  e.target = process;
  e.state = (uint32_t)m->dists.size();
  m->dists.push_back(*dist);
  m->ext.push_back(e);
  if (out_dist_id) *out_dist_id = e.state;
  return QK_OK;
}

int qk_model_set_container_factory(qk_model* m, uint32_t source, const qk_dist_desc* create_quantity,
                                   const double* split3, double capacity, uint32_t* out_dist_id) {
  if (!m || source >= m->comps.size() || m->comps[source].d.kind != QK_CONTAINER_SOURCE) {
    qk_set_error("qk_model_set_container_factory: component %u is not a container source", source);
    return QK_ERR_INVALID_ARG;
  }
  if (create_quantity && !check_dist(create_quantity)) {
    qk_set_error("qk_model_set_container_factory: invalid distribution parameters");
    return QK_ERR_INVALID_ARG;
  }
  QkHostComp& c = m->comps[source];
  if (create_quantity) {
    if (c.dist_cqty < 0) {
      c.dist_cqty = (int)m->dists.size();
      m->dists.push_back(*create_quantity);
    } else {
      m->dists[(size_t)c.dist_cqty] = *create_quantity;
    }
  } else {
    c.dist_cqty = -1;
  }
  for (int k = 0; k < 3; ++k) c.csplit[k] = split3 ? split3[k] : 0.0;
  c.ccapacity = capacity;
  if (out_dist_id) *out_dist_id = c.dist_cqty >= 0 ? (uint32_t)c.dist_cqty : 0xFFFFFFFFu;
  return QK_OK;
}

int qk_model_set_initial_containers(qk_model* m, uint32_t stock, uint32_t n, const double* resources,
                                    const uint8_t* has_resource, const double* capacities) {
  if (!m || stock >= m->comps.size() || m->comps[stock].d.kind != QK_CONTAINER_STOCK || (n && !capacities)) {
    qk_set_error("qk_model_set_initial_containers: component %u is not a container stock (or capacities is NULL)", stock);
    return QK_ERR_INVALID_ARG;
  }
  QkHostComp& c = m->comps[stock];
  c.d.n_initial_items = n;
  c.init_res.assign((size_t)n * 3, 0.0);
  c.init_has.assign(n, 0);
  c.init_cap.assign(n, 0.0);
  for (uint32_t k = 0; k < n; ++k) {
    c.init_has[k] = (has_resource && resources) ? has_resource[k] : 0;
    if (c.init_has[k])
      for (int j = 0; j < 3; ++j) c.init_res[(size_t)k * 3 + j] = resources[(size_t)k * 3 + j];
    c.init_cap[k] = capacities[k];
  }
  return QK_OK;
}

int qk_model_n_components(const qk_model* m, uint32_t* out) {
  if (!m || !out) return QK_ERR_INVALID_ARG;
  *out = (uint32_t)m->comps.size();
  return QK_OK;
}
int qk_model_n_dists(const qk_model* m, uint32_t* out) {
  if (!m || !out) return QK_ERR_INVALID_ARG;
  *out = (uint32_t)m->dists.size();
  return QK_OK;
}
int qk_model_state_bytes(const qk_model* m, int with_log, uint32_t* out) {
  if (!m || !out) return QK_ERR_INVALID_ARG;
  QkLowered l;
  int rc = qk_lower(m, with_log != 0, 0, &l);
  if (rc) return rc;
  *out = l.hdr.n64 * 8 + l.hdr.n32 * 4;
  return QK_OK;
}

} /* extern "C" */

/* ------------------------------------------------------------------------------------------------ */

static uint32_t align8(uint32_t x) { return (x + 7u) & ~7u; }

/* Duration::from_secs_f64 on the host, for Constant process times (resolved once at lowering):
 * round-half-even of the exact binary value * 1e9; >= 2^34 s, negative and NaN are faults.
 * Returns ns or QK_DUR_FAULT_BASE + qk_rep_status. */
static uint64_t host_duration_code(double secs) {
  if (!(secs >= 0.0)) return QK_DUR_FAULT_BASE + QK_REP_BAD_DURATION;
  uint64_t bits;
  memcpy(&bits, &secs, 8);
  const int e = (int)((bits >> 52) & 0x7FF) - 1023;
  const uint64_t mant = (bits & 0xFFFFFFFFFFFFFull) | 0x10000000000000ull;
  if (e < -31) return 0;
  if (e >= 34) return QK_DUR_FAULT_BASE + QK_REP_BAD_DURATION;
  const unsigned __int128 prod = (unsigned __int128)mant * 1000000000ull;
  const int sh = 52 - e; /* 19..83 */
  unsigned __int128 q = prod >> sh;
  const unsigned __int128 rem = prod & (((unsigned __int128)1 << sh) - 1);
  const unsigned __int128 half = (unsigned __int128)1 << (sh - 1);
  if (rem > half || (rem == half && (q & 1))) q += 1;
  return (uint64_t)q;
}

int qk_lower(const qk_model* m, bool log, int placement, QkLowered* out) {
  const bool split = placement != 0;   /* component records in the cold planes */
  const bool qsplit = placement == 2;  /* ... and the process-like fire slots too (two-tier scheduler queue) */
  const size_t nc = m->comps.size();
  if (nc == 0) {
    qk_set_error("qk_lower: model has no components");
    return QK_ERR_INVALID_ARG;
  }
  std::vector<DevComp> dc(nc);
  std::vector<DevDist> dd(m->dists.size());
  std::vector<DevDelayMode> dms;
  std::vector<uint16_t> subs;
  std::vector<uint16_t> sched;
  std::vector<DevVec> vecs;

  for (size_t i = 0; i < m->dists.size(); ++i) {
    memset(&dd[i], 0, sizeof(DevDist));
    dd[i].kind = m->dists[i].kind;
    dd[i].stream = m->dists[i].stream;
    for (int k = 0; k < 4; ++k) dd[i].p[k] = m->dists[i].p[k];
    dd[i].draw_slot = 0xFFFFFFFFu;
  }

  /* schedulable components in rank order */
  for (size_t i = 0; i < nc; ++i) {
    memset(&dc[i], 0, sizeof(DevComp));
    uint32_t k = m->comps[i].d.kind;
    dc[i].fire_idx = 0xFFFF;
    if (is_item_stock(k) || k == QK_VECTOR_STOCK || qk_is_process_like(k)) {
      dc[i].fire_idx = (uint16_t)sched.size();
      sched.push_back((uint16_t)i);
    }
  }
  uint32_t n_sched = (uint32_t)sched.size();
  uint32_t n_stale_words = (n_sched + 31) / 32;
  if (n_stale_words == 0) n_stale_words = 1;
  /* prefetchable distributions: process_time_distr of single-server components when it is not Constant */
  std::vector<uint16_t> pf_comp;
  for (size_t i = 0; i < nc; ++i) {
    uint32_t k = device_kind(m->comps[i].d.kind);
    if ((k == QK_DISCRETE_PROCESS || k == QK_DISCRETE_SOURCE || k == QK_DISCRETE_SINK) &&
        m->dists[(size_t)m->comps[i].dist_time].kind != QK_DIST_CONSTANT)
      pf_comp.push_back((uint16_t)i);
  }
  const uint32_t n_pf = (uint32_t)pf_comp.size();
  const uint32_t n_pf_words = (n_pf + 31) / 32;
  uint32_t n64c = 0, n32c = 0; /* cold slots */
  uint32_t n64 = G64_FIRE0 + ((n_sched + 3u) & ~3u); /* fire[] padded to a multiple of four (scheduler scan) */
  const uint32_t pf32 = G32_STALE0 + n_stale_words;
  uint32_t n32 = pf32 + n_pf_words;
  /* two-tier scheduler queue (placement 2): the hot fire array holds the STOCKS' slots only (their emit_change events
   * fire one nanosecond after they are scheduled: they are the next action almost every time one is pending); the keyed
   * events of the process-like components live in a cold array indexed by rank, summarised on chip by a cached minimum
   * (kernel registers) and a busy mask (hot words after the prefetch mask).  A stock's fire_idx is then its index in
   * the hot array; hot_rank[] maps it back to the rank. */
  std::vector<uint16_t> hot_rank;
  uint32_t cfire0 = 0, busy32 = 0, emask32 = 0, gshift = 0, n_groups = 0, gmin64 = 0, gidx32 = 0, gdirty32 = 0;
  if (qsplit) {
    for (uint32_t r = 0; r < n_sched; ++r) {
      const uint32_t k = m->comps[sched[r]].d.kind;
      if (is_item_stock(k) || k == QK_VECTOR_STOCK) {
        dc[sched[r]].fire_idx = (uint16_t)hot_rank.size();
        hot_rank.push_back((uint16_t)r);
      }
    }
    n64 = G64_FIRE0 + (((uint32_t)hot_rank.size() + 3u) & ~3u);
    cfire0 = n64c;
    n64c += (n_sched + 3u) & ~3u;
    busy32 = n32;
    n32 += n_stale_words;
    emask32 = n32; /* bit j: hot fire entry j holds a pending emit_change */
    n32 += ((uint32_t)hot_rank.size() + 31u) / 32u;
    /* group minima of the cold array: ranks in groups of 2^gshift (16, or more when there are over 512 ranks: the dirty
     * mask is one word), each with its minimum (time: hot64 ; rank: hot32) kept exact on chip, so that re-establishing the
     * overall minimum after a keyed event fired fetches ONE group of the cold array instead of all of it */
    gshift = 4;
    while (((n_sched + (1u << gshift) - 1u) >> gshift) > 32u) ++gshift;
    n_groups = (n_sched + (1u << gshift) - 1u) >> gshift;
    if (n_groups == 0) n_groups = 1;
    gmin64 = n64;
    n64 += n_groups;
    gidx32 = n32;
    n32 += n_groups;
    gdirty32 = n32++;
  }
  /* split placement (qk_batch_create, QK_BATCH_STATE_SPLIT): only what the scheduler pop and the quick path read stays in
   * the hot planes -- the fire slots, the stocks' head|count and pending-emit words, the seq counters -- and every
   * component record (countdowns, flags, items, pre-drawn durations, draw counters, resources) is allocated in the
   * cold planes, i.e. lives in global memory only */
  uint32_t& w64 = split ? n64c : n64;
  uint32_t& w32 = split ? n32c : n32;

  /* capacities that depend on neighbours */
  std::vector<uint32_t> cap(nc, 0);
  for (size_t i = 0; i < nc; ++i) { /* stock rings first */
    const QkHostComp& c = m->comps[i];
    if (!is_item_stock(c.d.kind)) continue;
    uint32_t r;
    if (c.d.capacity_hint) {
      r = c.d.capacity_hint;
    } else {
      uint32_t base = std::max(c.d.max_capacity, c.d.n_initial_items);
      if (base > 4096) {
        qk_set_error("component %zu: max_capacity %u needs an explicit capacity_hint (item-ring slots)", i,
                     c.d.max_capacity);
        return QK_ERR_CAPACITY;
      }
      uint32_t producers = 0;
      for (size_t j = 0; j < nc; ++j)
        if (m->comps[j].down == (int)i && !is_multi_server(m->comps[j].d.kind)) producers++;
      /* capacity is advisory (discrete.rs:181-186 has no check): each single-server producer that saw
       * Empty/Normal when it started may still push when the stock is already at max */
      r = base + producers;
      if (r == 0) r = 1;
    }
    if (r < c.d.n_initial_items || r > 0x7FFF) { /* head and count are 15-bit fields of S32_HC */
      qk_set_error("component %zu: item ring of %u slots is out of range", i, r);
      return QK_ERR_CAPACITY;
    }
    cap[i] = r;
  }
  for (size_t i = 0; i < nc; ++i) {
    const QkHostComp& c = m->comps[i];
    if (!is_multi_server(c.d.kind)) continue;
    uint32_t r = c.d.capacity_hint ? c.d.capacity_hint : (c.up >= 0 ? 2 * cap[(size_t)c.up] + 8 : 8);
    /* Load / Unload are bounded by max_process_count (vector_container.rs:350, 661) */
    if (c.d.kind == QK_CONTAINER_LOAD_PROCESS || c.d.kind == QK_CONTAINER_UNLOAD_PROCESS) r = c.d.max_capacity;
    if (r > 0xFFFF) {
      qk_set_error("component %zu: more than 65535 in-flight slots (capacity_hint / max_process_count %u)", i, r);
      return QK_ERR_CAPACITY;
    }
    cap[i] = r;
  }

  for (size_t i = 0; i < nc; ++i) {
    const QkHostComp& c = m->comps[i];
    DevComp& d = dc[i];
    d.kind = (uint8_t)device_kind(c.d.kind);
    d.n_dm = (uint8_t)c.dms.size();
    d.src_ord = (uint8_t)(c.src_ord >= 0 ? c.src_ord : 0);
    d.up = (int16_t)c.up;
    d.down = (int16_t)c.down;
    d.env = (int16_t)c.env;
    d.low = c.d.low_capacity;
    d.max = c.d.max_capacity;
    d.ring_cap = (uint16_t)cap[i];
    d.dist_time = (uint16_t)(c.dist_time >= 0 ? c.dist_time : 0);
    d.flags = c.logged ? DC_LOGGED : 0;
    const bool is_ctr = is_container_kind(c.d.kind) != 0;
    if (is_ctr) d.flags |= DC_CONTAINER;
    d.o64_ttnpe = 0xFFFF;
    d.o32_next_item = 0xFFFF;
    d.o64_nextdur = 0xFFFF;
    d.pf_idx = 0xFFFF;
    if (c.dist_time >= 0 && m->dists[(size_t)c.dist_time].kind == QK_DIST_CONSTANT)
{
      const uint64_t cd = host_duration_code(m->dists[(size_t)c.dist_time].p[0]);
      d.const_dur_lo = (uint32_t)cd;
      d.const_dur_hi = (uint32_t)(cd >> 32);
    }
    if (c.subs.size() > 255) {
      qk_set_error("component %zu: more than 255 subscribers", i);
      return QK_ERR_CAPACITY;
    }
    d.n_subs = (uint8_t)c.subs.size();
    d.subs_off = (uint16_t)subs.size();
    for (size_t k = 0; k < c.subs.size(); ++k) subs.push_back((uint16_t)c.subs[k]);
    d.dm_off = (uint16_t)dms.size();
    for (size_t k = 0; k < c.dms.size(); ++k) dms.push_back(c.dms[k]);

    uint32_t kind = device_kind(c.d.kind);
    if (kind == QK_DISCRETE_PROCESS || kind == QK_DISCRETE_SOURCE || kind == QK_DISCRETE_SINK) {
      d.o64 = (uint16_t)w64;
      w64 += P64_BASE_COUNT;
      if (kind == QK_DISCRETE_SINK) d.o64_ttnpe = (uint16_t)w64++;
      d.o64_dm = (uint16_t)w64;
      w64 += d.n_dm;
      d.o32 = (uint16_t)w32;
      w32 += P32_BASE_COUNT;
      if (kind == QK_DISCRETE_SOURCE) d.o32_next_item = (uint16_t)w32++;
      for (uint32_t k = 0; k < n_pf; ++k)
        if (pf_comp[k] == i) {
          d.pf_idx = (uint16_t)k;
          d.o64_nextdur = (uint16_t)w64++;
        }
      if (is_ctr) {
        d.ocold64 = (uint16_t)n64c;
        n64c += 3;
        if (kind == QK_DISCRETE_SOURCE) { /* Vector3ContainerFactory (vector_container.rs:126-156) */
          DevVec v;
          memset(&v, 0, sizeof(v));
          v.dim = 3;
          v.tmask = 7;
          for (int k = 0; k < 3; ++k) v.vinit[k] = c.csplit[k]; /* create_component_split */
          v.vmax = c.ccapacity;
          for (int k = 0; k < 5; ++k) v.ups[k] = v.downs[k] = -1;
          v.dist_qty = (uint16_t)(c.dist_cqty >= 0 ? c.dist_cqty : 0xFFFF); /* create_quantity_distr or None */
          if (c.dist_cqty >= 0 && m->dists[(size_t)c.dist_cqty].kind != QK_DIST_CONSTANT)
            dd[(size_t)c.dist_cqty].draw_slot = w32++;
          d.vec_idx = (uint16_t)vecs.size();
          vecs.push_back(v);
        }
      }
    } else if (is_multi_server(kind)) {
      d.o64 = (uint16_t)w64;
      w64 += PP64_BASE_COUNT;
      d.o64_dm = (uint16_t)w64;
      w64 += d.n_dm;
      w64 += d.ring_cap; /* in-progress durations */
      d.o32 = (uint16_t)w32;
      w32 += PP32_BASE_COUNT + 2u * d.ring_cap;
      if (is_ctr) {
        d.ocold64 = (uint16_t)n64c;
        n64c += 6u * d.ring_cap; /* payload of the in-progress slots, then of the complete ring */
      }
      if (kind == QK_CONTAINER_LOAD_PROCESS || kind == QK_CONTAINER_UNLOAD_PROCESS) {
        DevVec v;
        memset(&v, 0, sizeof(v));
        v.dim = 3;
        v.tmask = 7;
        for (int k = 0; k < 5; ++k) {
          v.ups[k] = (int16_t)c.ups[k];     /* Load: ups[0] = the Vector3Stock it withdraws resource from */
          v.downs[k] = (int16_t)c.downs[k]; /* Unload: downs[0] = the Vector3Stock it pushes resource to */
        }
        v.dist_qty = (uint16_t)c.dist_qty; /* process_capacity_ratio_distr / process_quantity_ratio_distr */
        if (m->dists[(size_t)c.dist_qty].kind != QK_DIST_CONSTANT) dd[(size_t)c.dist_qty].draw_slot = w32++;
        d.vec_idx = (uint16_t)vecs.size();
        vecs.push_back(v);
      }
    } else if (kind == QK_DISCRETE_STOCK) {
      d.o64 = (uint16_t)n64;
      n64 += S64_COUNT;
      d.o32 = (uint16_t)n32;
      n32 += S32_COUNT;
      if (log) n32 += 2; /* S32_EMITSRC, S32_HEADITEM */
      d.ocold32 = (uint16_t)n32c;
      n32c += d.ring_cap;
      if (is_ctr) {
        d.ocold64 = (uint16_t)n64c;
        n64c += 3u * d.ring_cap;
      }
      uint32_t depth = 0;
      for (size_t j = 0; j < nc; ++j) {
        const QkHostComp& o = m->comps[j];
        if (o.up == (int)i || o.down == (int)i) depth += 2u * (is_multi_server(o.d.kind) ? cap[j] : 1u);
      }
      if (depth < 2) depth = 2;
      if (depth > 250) depth = 250;
      d.emit_depth = (uint16_t)depth;
      d.dist_time = (uint16_t)c.d.n_initial_items; /* table field reuse: initial item count */
    } else if (kind == QK_VECTOR_STOCK) {
      DevVec v;
      memset(&v, 0, sizeof(v));
      v.dim = (uint16_t)c.dim;
      v.vmax = c.vmax;
      v.vlow = c.vlow;
      v.tmask = (uint16_t)c.tmask;
      for (int k = 0; k < QK_VEC_MAX_DIM; ++k) v.vinit[k] = c.vinit[k];
      for (int k = 0; k < 5; ++k) v.ups[k] = v.downs[k] = -1;
      d.o64 = (uint16_t)w64;
      v.o64_res = (uint16_t)w64;
      w64 += c.dim;
      v.o64_low = (uint16_t)w64++;
      v.o64_last = (uint16_t)w64++;
      d.o32 = (uint16_t)n32;
      n32 += S32_COUNT; /* S32_HC unused, S32_EMIT used */
      uint32_t depth = 0;
      for (size_t j = 0; j < nc; ++j) {
        const QkHostComp& o = m->comps[j];
        if (o.up == (int)i || o.down == (int)i) depth += 2;
        for (int k = 0; k < 5; ++k)
          if (o.ups[k] == (int)i || o.downs[k] == (int)i) depth += 2u * (is_multi_server(o.d.kind) ? cap[j] : 1u);
      }
      depth += 2 * (uint32_t)m->ext.size(); /* SetLowCapacity's add(0.) can flip the state too */
      if (depth < 2) depth = 2;
      if (depth > 250) depth = 250;
      d.emit_depth = (uint16_t)depth;
      d.vec_idx = (uint16_t)vecs.size();
      vecs.push_back(v);
    } else if (is_vector_kind(kind)) { /* vector process-like */
      DevVec v;
      memset(&v, 0, sizeof(v));
      v.dim = (uint16_t)c.dim;
      v.n_ports = (uint16_t)c.n_ports;
      v.tmask = (uint16_t)c.tmask;
      for (int k = 0; k < QK_VEC_MAX_DIM; ++k) v.vinit[k] = c.vinit[k];
      for (int k = 0; k < 5; ++k) {
        v.ratios[k] = c.ratios[k];
        v.ups[k] = (int16_t)c.ups[k];
        v.downs[k] = (int16_t)c.downs[k];
      }
      v.dist_qty = (uint16_t)c.dist_qty;
      v.o32_qty = 0xFFFF;
      d.o64 = (uint16_t)w64;
      w64 += P64_BASE_COUNT;
      d.o64_dm = (uint16_t)w64;
      w64 += d.n_dm;
      v.o64_res = (uint16_t)w64;
      w64 += c.dim + (kind == QK_VECTOR_COMBINER ? 1u : 0u);
      d.o32 = (uint16_t)w32;
      w32 += P32_BASE_COUNT;
      if (m->dists[(size_t)c.dist_qty].kind != QK_DIST_CONSTANT) dd[(size_t)c.dist_qty].draw_slot = w32++;
      for (size_t k = 0; k < m->ext.size(); ++k) /* SetProcessQuantity events: the instance they bring, and where the
                                                    component keeps which instance is current */
        if (m->ext[k].type == 2 && m->ext[k].target == i) {
          if (v.o32_qty == 0xFFFF) v.o32_qty = (uint16_t)w32++;
          if (m->dists[m->ext[k].state].kind != QK_DIST_CONSTANT) dd[m->ext[k].state].draw_slot = w32++;
        }
      d.vec_idx = (uint16_t)vecs.size();
      vecs.push_back(v);
    } else { /* environment */
      d.o32 = (uint16_t)w32;
      w32 += E32_COUNT;
      d.low = c.d.initial_env_state ? 1u : 0u; /* table field reuse: initial BasicEnvironmentState */
    }
    if (qk_is_process_like(kind)) {
      if (m->dists[(size_t)c.dist_time].kind != QK_DIST_CONSTANT) dd[(size_t)c.dist_time].draw_slot = w32++;
      for (size_t k = 0; k < c.dms.size(); ++k) {
        if (m->dists[c.dms[k].until_delay].kind != QK_DIST_CONSTANT) dd[c.dms[k].until_delay].draw_slot = w32++;
        if (m->dists[c.dms[k].until_fix].kind != QK_DIST_CONSTANT) dd[c.dms[k].until_fix].draw_slot = w32++;
      }
    }
    if (log) {
      d.olog32 = (uint16_t)n32;
      n32 += 1; /* L32_SEQ */
      if (kind == QK_DISCRETE_STOCK) n32c += d.emit_depth; /* emit source ring follows the item ring */
      if (kind == QK_VECTOR_STOCK) {
        d.ocold32 = (uint16_t)n32c;
        n32c += d.emit_depth;
        d.olog64 = (uint16_t)n64c; /* occupied (f64) of each pending emit_change */
        n64c += d.emit_depth;
      }
      if (qk_is_process_like(kind)) {
        d.olog64 = (uint16_t)n64c;
        n64c += PL64_COUNT;
      }
    }
    if (n64 > 0xFFF0 || n32 > 0xFFF0 || n64c > 0xFFF0 || n32c > 0xFFF0) {
      qk_set_error("model state exceeds 65520 slots per replication");
      return QK_ERR_CAPACITY;
    }
  }
  uint32_t ident_off = (uint32_t)subs.size();
  for (size_t i = 0; i < nc; ++i) subs.push_back((uint16_t)i);
  /* Model::init list: every process-like component runs update_state(INIT_000000) (discrete.rs:392-398) and
   * BasicEnvironment logs its state (core.rs:229-234); stocks use the default no-op init */
  uint32_t init_off = (uint32_t)subs.size(), n_init = 0;
  for (size_t i = 0; i < nc; ++i)
    if (!is_item_stock(m->comps[i].d.kind) && m->comps[i].d.kind != QK_VECTOR_STOCK) {
      subs.push_back((uint16_t)i);
      n_init++;
    }

  /* externally scheduled events: origin rank 0, FIFO among equal times (stable sort) */
  std::vector<DevExt> ext = m->ext;
  std::stable_sort(ext.begin(), ext.end(), [](const DevExt& a, const DevExt& b) { return a.t < b.t; });

  DevHeader h;
  memset(&h, 0, sizeof(h));
  h.n_comp = (uint32_t)nc;
  h.n_sched = n_sched;
  h.n_dist = (uint32_t)dd.size();
  h.n_dm = (uint32_t)dms.size();
  h.n_subs = (uint32_t)subs.size();
  h.n_ext = (uint32_t)ext.size();
  h.n64 = n64;
  h.n32 = n32;
  h.n64c = n64c;
  h.n32c = n32c;
  h.n_stale_words = n_stale_words;
  h.n_pf = n_pf;
  h.n_pf_words = n_pf_words;
  h.pf32 = pf32;
  h.ident_off = ident_off;
  h.init_off = init_off;
  h.n_init = n_init;
  h.log_enabled = log ? 1 : 0;
  h.split = (uint32_t)placement;
  h.n_hot = (uint32_t)hot_rank.size();
  h.cfire0 = cfire0;
  h.busy32 = busy32;
  h.emask32 = emask32;
  h.gshift = gshift;
  h.n_groups = n_groups;
  h.gmin64 = gmin64;
  h.gidx32 = gidx32;
  h.gdirty32 = gdirty32;
  uint32_t off = align8((uint32_t)sizeof(DevHeader));
  h.off_comps = off;
  off = align8(off + (uint32_t)(nc * sizeof(DevComp)));
  h.off_dists = off;
  off = align8(off + (uint32_t)(dd.size() * sizeof(DevDist)));
  h.off_ext = off;
  off = align8(off + (uint32_t)(ext.size() * sizeof(DevExt)));
  h.off_dms = off;
  off = align8(off + (uint32_t)(dms.size() * sizeof(DevDelayMode)));
  h.off_subs = off;
  off = align8(off + (uint32_t)(subs.size() * sizeof(uint16_t)));
  h.off_sched = off;
  off = align8(off + (uint32_t)(sched.size() * sizeof(uint16_t)));
  h.off_pf = off;
  off = align8(off + (uint32_t)(pf_comp.size() * sizeof(uint16_t)));
  h.off_hot_rank = off;
  off = align8(off + (uint32_t)(hot_rank.size() * sizeof(uint16_t)));
  h.n_vec = (uint32_t)vecs.size();
  h.off_vecs = off;
  off = align8(off + (uint32_t)(vecs.size() * sizeof(DevVec)));
  /* hot rows (see qk_internal.h) */
  std::vector<DevHot> hot(nc);
  h.features = ext.empty() ? 0 : 1;
  for (size_t i = 0; i < nc; ++i) {
    const DevComp& d = dc[i];
    DevHot& q = hot[i];
    memset(&q, 0, sizeof(q));
    q.a.kind = d.kind;
    q.a.n_dm = d.n_dm;
    q.a.n_subs = d.n_subs;
    q.a.src_ord = d.src_ord;
    q.a.o32 = d.o32;
    q.a.o64 = d.o64;
    q.a.fire_idx = d.fire_idx;
    q.a.flags = d.flags;
    const bool single = d.kind == QK_DISCRETE_PROCESS || d.kind == QK_DISCRETE_SOURCE || d.kind == QK_DISCRETE_SINK;
    if (single && d.n_dm == 0 && d.env < 0 && !(d.flags & DC_CONTAINER)) q.a.flags |= DC_SIMPLE;
    if (!(d.kind == QK_DISCRETE_STOCK || (q.a.flags & DC_SIMPLE)) || (d.flags & DC_CONTAINER)) h.features = 1;
    q.a.olog32 = d.olog32;
    q.a.olog64 = d.kind == QK_DISCRETE_STOCK ? d.ocold32 : d.olog64;
    if (qk_is_process_like(d.kind)) {
      q.b.up_hc = d.up >= 0 ? (uint16_t)(dc[(size_t)d.up].o32 + S32_HC) : (uint16_t)QK_HC_NONE;
      q.b.down_hc = d.down >= 0 ? (uint16_t)(dc[(size_t)d.down].o32 + S32_HC) : (uint16_t)QK_HC_NONE;
      q.b.env = d.env;
      q.b.dist_time = d.dist_time;
      q.b.o64_dm = d.o64_dm;
      q.b.dm_off = d.dm_off;
      q.b.o64_nextdur = d.o64_nextdur;
      q.b.pf_idx = d.pf_idx;
    } else {
      q.sb.low = d.low;
      q.sb.max = d.max;
      q.sb.ring_cap = d.ring_cap;
      q.sb.emit_depth = d.emit_depth;
      q.sb.subs_off = d.subs_off;
      q.sb.o32 = d.o32;
    }
    q.c.o64_ttnpe = d.o64_ttnpe;
    q.c.o32_next_item = d.o32_next_item;
    q.c.const_dur_lo = d.const_dur_lo;
    q.c.const_dur_hi = d.const_dur_hi;
    q.c.vec_idx = d.vec_idx;
    q.c.ring_cap = d.ring_cap;
  }
  off = (off + 15u) & ~15u;
  h.off_hot = off;
  off = align8(off + (uint32_t)(nc * sizeof(DevHot)));
  /* quick-path descriptors, one per subs[] entry */
  std::vector<DevSubQ> subq(subs.size());
  for (size_t i = 0; i < subs.size(); ++i) {
    const DevComp& d = dc[subs[i]];
    const DevHot& q = hot[subs[i]];
    DevSubQ& s = subq[i];
    memset(&s, 0, sizeof(s));
    const bool single_k = d.kind == QK_DISCRETE_PROCESS || d.kind == QK_DISCRETE_SOURCE || d.kind == QK_DISCRETE_SINK;
    const bool dm_quick = single_k && d.n_dm > 0 && d.env < 0 && !(d.flags & DC_CONTAINER);
    if (!(q.a.flags & DC_SIMPLE) && !dm_quick) continue; /* flags == 0: always the full handler */
    s.flags = dm_quick ? SQ_DM : SQ_QUICK;
    s.fire_slot = qsplit ? d.fire_idx : (uint16_t)(G64_FIRE0 + d.fire_idx); /* two-tier queue: the rank itself */
    if (d.kind == QK_DISCRETE_SOURCE) s.flags |= SQ_US_FORCED | (1u << SQ_US_SHIFT);
    else if (d.up < 0) s.flags |= SQ_US_FORCED | (3u << SQ_US_SHIFT);
    else s.up_hc = q.b.up_hc;
    if (d.kind == QK_DISCRETE_SINK) s.flags |= SQ_DS_FORCED | (1u << SQ_DS_SHIFT);
    else if (d.down < 0) s.flags |= SQ_DS_FORCED | (3u << SQ_DS_SHIFT);
    else s.down_hc = q.b.down_hc;
  }
  for (size_t i = 0; i < nc; ++i) { /* SQ_DUP: duplicates within one subscriber list */
    const DevComp& d = dc[i];
    for (uint32_t a = 0; a < d.n_subs; ++a)
      for (uint32_t b2 = 0; b2 < d.n_subs; ++b2)
        if (a != b2 && subs[d.subs_off + a] == subs[d.subs_off + b2]) subq[d.subs_off + a].flags |= SQ_DUP;
  }
  h.off_subq = off;
  off = align8(off + (uint32_t)(subq.size() * sizeof(DevSubQ)));
  /* container tables: capacity by handle, resource of the initial containers (numbered in registration order of
   * their stocks, like the handles the init kernel hands out) */
  std::vector<double> ccap(64, 0.0);
  std::vector<uint64_t> cinit;
  {
    uint32_t base = 0;
    for (size_t i = 0; i < nc; ++i) {
      const QkHostComp& c = m->comps[i];
      if (c.d.kind == QK_CONTAINER_SOURCE) ccap[(size_t)c.src_ord] = c.ccapacity;
      if (!is_item_stock(c.d.kind)) continue;
      for (uint32_t k = 0; k < c.d.n_initial_items; ++k) {
        const bool has = c.d.kind == QK_CONTAINER_STOCK && c.init_has[k];
        ccap.push_back(c.d.kind == QK_CONTAINER_STOCK ? c.init_cap[k] : 0.0);
        for (int j = 0; j < 3; ++j) {
          uint64_t w = 0;
          if (has) memcpy(&w, &c.init_res[(size_t)k * 3 + j], 8);
          else if (j == 0) w = QK_CONTAINER_NONE_BITS;
          cinit.push_back(w);
        }
      }
      base += c.d.n_initial_items;
    }
    if (base > 0x3FFFFFFu) {
      qk_set_error("more than 2^26 initial items");
      return QK_ERR_CAPACITY;
    }
  }
  const bool any_ctr = std::any_of(dc.begin(), dc.end(), [](const DevComp& d) { return (d.flags & DC_CONTAINER) != 0; });
  /* a vector resource that is neither f64 nor Vector3 (user-defined: more fields, or a total over a subset) */
  const bool any_wide = std::any_of(m->comps.begin(), m->comps.end(), [](const QkHostComp& c) {
    return is_vector_kind(c.d.kind) && !((c.dim == 1) || (c.dim == 3 && c.tmask == 7u));
  });
  if (any_wide) {
    if (any_ctr) {
      qk_set_error("user-defined vector resources and the Vector3 container family cannot share one model");
      return QK_ERR_INVALID_ARG;
    }
    h.features = 3;
  }
  if (any_ctr) {
    h.features = 2;
    h.off_ccap = off;
    off = align8(off + (uint32_t)(ccap.size() * 8));
    h.off_cinit = off;
    off = align8(off + (uint32_t)(cinit.size() * 8));
  }
  h.total_bytes = off;

  out->blob.assign(off, 0);
  if (any_ctr) {
    memcpy(out->blob.data() + h.off_ccap, ccap.data(), ccap.size() * 8);
    if (!cinit.empty()) memcpy(out->blob.data() + h.off_cinit, cinit.data(), cinit.size() * 8);
  }
  memcpy(out->blob.data() + h.off_hot, hot.data(), nc * sizeof(DevHot));
  memcpy(out->blob.data() + h.off_subq, subq.data(), subq.size() * sizeof(DevSubQ));
  memcpy(out->blob.data(), &h, sizeof(h));
  memcpy(out->blob.data() + h.off_comps, dc.data(), nc * sizeof(DevComp));
  if (!dd.empty()) memcpy(out->blob.data() + h.off_dists, dd.data(), dd.size() * sizeof(DevDist));
  if (!ext.empty()) memcpy(out->blob.data() + h.off_ext, ext.data(), ext.size() * sizeof(DevExt));
  if (!dms.empty()) memcpy(out->blob.data() + h.off_dms, dms.data(), dms.size() * sizeof(DevDelayMode));
  memcpy(out->blob.data() + h.off_subs, subs.data(), subs.size() * sizeof(uint16_t));
  if (!sched.empty()) memcpy(out->blob.data() + h.off_sched, sched.data(), sched.size() * sizeof(uint16_t));
  if (!pf_comp.empty()) memcpy(out->blob.data() + h.off_pf, pf_comp.data(), pf_comp.size() * sizeof(uint16_t));
  if (!hot_rank.empty()) memcpy(out->blob.data() + h.off_hot_rank, hot_rank.data(), hot_rank.size() * sizeof(uint16_t));
  if (!vecs.empty()) memcpy(out->blob.data() + h.off_vecs, vecs.data(), vecs.size() * sizeof(DevVec));
  out->hdr = h;
  out->vecs = vecs;
  out->comps = dc;
  out->dists = dd;
  return QK_OK;
}
