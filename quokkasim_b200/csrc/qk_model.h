/* qk_model.h -- host-side model description and lowering (internal). */
#ifndef QK_MODEL_H
#define QK_MODEL_H
#include <string>
#include <vector>

#include "qk_internal.h"

struct QkHostComp {
  qk_component_desc d;
  int up, down, env;
  std::vector<int> subs;
  int dist_time, dist_qty;
  std::vector<DevDelayMode> dms;
  bool logged;
  int src_ord;
  uint32_t initial_base;
  /* vector components */
  uint32_t dim, n_ports;
  double vlow, vmax, vinit[QK_VEC_MAX_DIM], ratios[5];
  uint32_t tmask; /* fields counted by the resource's total (qk_model_set_vector_n) */
  int ups[5], downs[5];
  /* container family: factory of a container source; initial contents of a container stock */
  int dist_cqty; /* create_quantity_distr instance or -1 (None) */
  double csplit[3], ccapacity;
  std::vector<double> init_res; /* [n][3] */
  std::vector<uint8_t> init_has;
  std::vector<double> init_cap;
};

struct qk_model {
  std::vector<QkHostComp> comps;
  std::vector<qk_dist_desc> dists;
  std::vector<DevExt> ext;
  uint32_t n_sources;
  uint32_t n_initial_total;
};

struct QkLowered {
  std::vector<uint8_t> blob; /* DevHeader + tables */
  DevHeader hdr;
  std::vector<DevComp> comps;
  std::vector<DevDist> dists;
  std::vector<DevVec> vecs;
};

void qk_set_error(const char* fmt, ...);
int qk_is_process_like(uint32_t kind);
/* lowers the model for the given logging mode and state placement (0: everything in the hot planes ; 1: component
 * records in the cold planes ; 2: ... and the keyed events of the process-like components too, see qk_lower); returns
 * QK_OK or an error */
int qk_lower(const qk_model* m, bool log, int placement, QkLowered* out);

#endif
