/*
 * qk_vector.cuh -- continuous (vector) components on the device: VectorStock, VectorProcess, VectorSource,
 * VectorSink, VectorCombiner<M>, VectorSplitter<N> for f64 (dim 1) and Vector3 (dim 3) resources, and -- in the
 * feature-set-3 kernels, VD = QK_VEC_MAX_DIM -- for user-defined resources of up to eight fields whose total counts a
 * subset of them (quokkasim_examples/src/bin/iron_ore.rs:21-96).
 * Included by qk_kernels.cuh (uses its Ctx / S64 / S32 / qk_log / qk_post machinery).
 *
 * Replaces, for the batched path (paths relative to /root/reference/quokkasim/src):
 *   resource algebra                     core.rs:7-128      qk_vtotal / qk_vremove (fp64, one rounding per op)
 *   VectorStock add/remove/emit_change   vector.rs:89-171   qk_vstock_add / qk_vstock_remove / emit in SELECT
 *   VectorProcess::update_state          vector.rs:362-523  qk_update_vector (kind flag)
 *   VectorCombiner / VectorSplitter      vector.rs:822-1001, 1124-1299
 *   VectorSource / VectorSink            vector.rs:1406-1553, 1664-1815
 * All arithmetic uses explicit round-to-nearest intrinsics in the reference's operation order, so stock levels
 * agree with the CPU oracle bit for bit (north_star asks for 1e-9 relative).
 */
#ifndef QK_VECTOR_CUH
#define QK_VECTOR_CUH

/* QK_D2U / QK_U2D: qk_kernels.cuh */

/* continuation record carrying three doubles (see qk_log_record docs in quokka_b200.h) */
template <bool LOG, int SM>
__device__ __forceinline__ void qk_log_vec(Ctx& cx, const DevComp* c, uint32_t comp, uint64_t id, uint32_t idx,
                                           double v0, double v1, double v2) {
  if (!LOG) return;
  if (c->flags & DC_LOGGED) {
    uint4* r = cx.log_base + 2 * (size_t)(cx.log_cnt & cx.log_mask);
    const uint64_t b0 = QK_D2U(v0), b1 = QK_D2U(v1), b2 = QK_D2U(v2);
    uint4 a, b;
    a.x = (uint32_t)b0;
    a.y = (uint32_t)(b0 >> 32);
    a.z = comp | ((uint32_t)QK_LOG_VEC_DATA << 16) | (idx << 24);
    a.w = (uint32_t)id; /* seq of the record it continues */
    b.x = (uint32_t)b1;
    b.y = (uint32_t)(b1 >> 32);
    b.z = (uint32_t)b2;
    b.w = (uint32_t)(b2 >> 32);
    qk_store_record(r, a, b);
    cx.log_cnt += 1;
  }
}

__device__ __forceinline__ double qk_vtotal(const double* v, uint32_t dim) {
  return dim == 1 ? v[0] : __dadd_rn(__dadd_rn(v[0], v[1]), v[2]); /* core.rs:96-100: left-to-right sum */
}
/* ResourceTotal of a resource of up to VD fields: f64 and Vector3 as above (VD == 3: nothing else exists in those
 * kernels); a user-defined resource sums the fields its total mask names, in field order (iron_ore.rs:75-79) */
template <int VD>
__device__ __forceinline__ double qk_vtotaln(const double* v, uint32_t dim, uint32_t mask) {
  if constexpr (VD == 3) {
    return qk_vtotal(v, dim);
  } else {
    if (dim == 1) return v[0];
    double t = 0.0;
    bool first = true;
#pragma unroll
    for (int k = 0; k < VD; ++k)
      if ((uint32_t)k < dim && ((mask >> k) & 1u)) {
        t = first ? v[k] : __dadd_rn(t, v[k]);
        first = false;
      }
    return t;
  }
}

/* fields beyond dim read as 0 */
template <int SM, int VD = 3>
__device__ __forceinline__ void qk_vload(Ctx& cx, uint32_t slot, uint32_t dim, double* v) {
  if constexpr (VD == 3) {
    v[0] = QK_U2D(W64(slot));
    v[1] = dim == 3 ? QK_U2D(W64(slot + 1)) : 0.0;
    v[2] = dim == 3 ? QK_U2D(W64(slot + 2)) : 0.0;
  } else {
#pragma unroll
    for (int k = 0; k < VD; ++k) v[k] = (uint32_t)k < dim ? QK_U2D(W64(slot + k)) : 0.0;
  }
}
template <int SM, int VD = 3>
__device__ __forceinline__ void qk_vstore(Ctx& cx, uint32_t slot, uint32_t dim, const double* v) {
  if constexpr (VD == 3) {
    W64(slot) = QK_D2U(v[0]);
    if (dim == 3) {
      W64(slot + 1) = QK_D2U(v[1]);
      W64(slot + 2) = QK_D2U(v[2]);
    }
  } else {
#pragma unroll
    for (int k = 0; k < VD; ++k)
      if ((uint32_t)k < dim) W64(slot + k) = QK_D2U(v[k]);
  }
}
/* the continuation records of one vector: one per three fields, index + 8 * chunk in the record's index byte */
template <bool LOG, int SM, int VD>
__device__ __forceinline__ void qk_log_vecn(Ctx& cx, const DevComp* c, uint32_t comp, uint64_t id, uint32_t idx,
                                            const double* v, uint32_t dim) {
  if constexpr (VD == 3) {
    qk_log_vec<LOG, SM>(cx, c, comp, id, idx, v[0], v[1], v[2]);
  } else {
#pragma unroll
    for (int ch = 0; 3 * ch < VD; ++ch)
      if ((uint32_t)(3 * ch) < dim)
        qk_log_vec<LOG, SM>(cx, c, comp, id, idx + 8u * ch, v[3 * ch], 3 * ch + 1 < VD ? v[3 * ch + 1] : 0.0,
                            3 * ch + 2 < VD ? v[3 * ch + 2] : 0.0);
  }
}

/* ResourceRemove: f64 `*self -= q; q` (core.rs:13-18); Vector3 proportional (core.rs:73-94) */
struct QkRemoved {
  double self[3], out[3];
};
/* one out-of-line copy (fp64 division and six multiplications; by value, see qk_draw) */
static __device__ __noinline__ QkRemoved qk_vremove3(double s0, double s1, double s2, double q) {
  QkRemoved r;
  const double total = __dadd_rn(__dadd_rn(s0, s1), s2);
  const double ps = __ddiv_rn(q, total);
  const double pr = __dsub_rn(1.0, ps);
  r.out[0] = __dmul_rn(s0, ps);
  r.out[1] = __dmul_rn(s1, ps);
  r.out[2] = __dmul_rn(s2, ps);
  r.self[0] = __dmul_rn(s0, pr);
  r.self[1] = __dmul_rn(s1, pr);
  r.self[2] = __dmul_rn(s2, pr);
  return r;
}
__device__ __forceinline__ void qk_vremove(double* self, double q, uint32_t dim, double* out) {
  if (dim == 1) {
    self[0] = __dsub_rn(self[0], q);
    out[0] = q;
    out[1] = out[2] = 0.0;
    return;
  }
  const QkRemoved r = qk_vremove3(self[0], self[1], self[2], q);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    out[k] = r.out[k];
    self[k] = r.self[k];
  }
}
/* ... and of a user-defined resource: proportional over EVERY field, the proportion taken of the masked total
 * (iron_ore.rs:52-73) */
template <int VD>
__device__ __forceinline__ void qk_vremoven(double* self, double q, uint32_t dim, uint32_t mask, double* out) {
  if constexpr (VD == 3) {
    qk_vremove(self, q, dim, out);
  } else {
    if (dim == 1) {
      self[0] = __dsub_rn(self[0], q);
      out[0] = q;
#pragma unroll
      for (int k = 1; k < VD; ++k) out[k] = 0.0;
      return;
    }
    const double ps = __ddiv_rn(q, qk_vtotaln<VD>(self, dim, mask));
    const double pr = __dsub_rn(1.0, ps);
#pragma unroll
    for (int k = 0; k < VD; ++k) {
      out[k] = (uint32_t)k < dim ? __dmul_rn(self[k], ps) : 0.0;
      if ((uint32_t)k < dim) self[k] = __dmul_rn(self[k], pr);
    }
  }
}

/* VectorStock::get_state (vector.rs:89-99): Full if max - occ <= 0, else Empty if occ < low, else Normal */
__device__ __forceinline__ uint32_t qk_vcat(double occ, double low, double vmax) {
  return __dsub_rn(vmax, occ) <= 0.0 ? 2u : (occ < low ? 0u : 1u);
}

template <int SM, int VD = 3>
__device__ __forceinline__ uint32_t qk_vstock_cat_of(Ctx& cx, int sidx) {
  if (sidx < 0) return 3u;
  const DevComp* c = &cx.comps[sidx];
  const DevVec* v = &cx.vecs[c->vec_idx];
  double r[VD];
  qk_vload<SM, VD>(cx, v->o64_res, v->dim, r);
  return qk_vcat(qk_vtotaln<VD>(r, v->dim, v->tmask), QK_U2D(W64(v->o64_low)), v->vmax);
}

/* cx.schedule_event(now + 1ns, emit_change, (state, id)) -- the emit carries the state computed NOW */
template <bool LOG, int SM>
__device__ __forceinline__ void qk_vemit_push(Ctx& cx, const DevComp* c, uint32_t src_seq, uint32_t cat, double occ) {
  const uint32_t fire = G64_FIRE0 + c->fire_idx;
  uint32_t e = S32(c->o32 + S32_EMIT);
  uint32_t n = e & 0xFFu, split = (e >> 8) & 0xFFu, head = (e >> 16) & 0xFFu;
  const uint64_t T = cx.now + 1;
  if (n == 0) {
    S64(fire) = T;
    if constexpr (qk_is_two_tier(SM)) S32(cx.emask32 + (c->fire_idx >> 5)) |= 1u << (c->fire_idx & 31); /* pending */
    split = 1;
    head = 0;
  } else if (S64(fire) == T) {
    split += 1;
  }
  n += 1;
  if (n > c->emit_depth) {
    cx.status = QK_REP_EMIT_OVERFLOW;
    return;
  }
  if (LOG) {
    uint32_t pos = head + n - 1;
    if (pos >= c->emit_depth) pos -= c->emit_depth;
    C32(c->ocold32 + pos) = src_seq | (cat << 30);
    C64(c->olog64 + pos) = QK_D2U(occ);
  }
  S32(c->o32 + S32_EMIT) = n | (split << 8) | (head << 16);
}

template <int SM>
__device__ __forceinline__ void qk_vstock_area(Ctx& cx, uint32_t sidx, const DevVec* v, double total) {
  const uint64_t last = W64(v->o64_last);
  const double dt = __dmul_rn((double)(cx.now - last), 1e-9);
  qk_red_addf64(&cx.gs64[(size_t)(sidx * ST_PER_COMP + ST64_TIME) * cx.rep_pad], __dmul_rn(total, dt));
  W64(v->o64_last) = cx.now;
}

/* Stock::add (core.rs:177-183; vector.rs:110-135) */
template <bool LOG, int SM, int VD = 3>
__device__ __forceinline__ void qk_vstock_add(Ctx& cx, uint32_t sidx, const double* add, uint64_t src) {
  const DevComp* c = &cx.comps[sidx];
  const DevHotA ha = QK_HOT_A(sidx);
  const DevVec* v = &cx.vecs[c->vec_idx];
  const uint32_t dim = v->dim;
  double r[VD];
  qk_vload<SM, VD>(cx, v->o64_res, dim, r);
  const double low = QK_U2D(W64(v->o64_low));
  const double before = qk_vtotaln<VD>(r, dim, v->tmask);
  const uint32_t prev = qk_vcat(before, low, v->vmax);
  qk_vstock_area<SM>(cx, sidx, v, before);
  if constexpr (VD == 3) {
    for (uint32_t k = 0; k < dim; ++k) r[k] = __dadd_rn(r[k], add[k]);
  } else {
#pragma unroll
    for (int k = 0; k < VD; ++k)
      if ((uint32_t)k < dim) r[k] = __dadd_rn(r[k], add[k]);
  }
  qk_vstore<SM, VD>(cx, v->o64_res, dim, r);
  QK_ST32_ADD(sidx, ST32_COUNT, 1u);
  const double after = qk_vtotaln<VD>(r, dim, v->tmask);
  const uint64_t id = qk_log<LOG, SM>(cx, ha, sidx, src, QK_LOG_VSTOCK_ADD, 0, QK_D2U(after));
  qk_log_vecn<LOG, SM, VD>(cx, c, sidx, id, 0, add, dim);
  const uint32_t cur = qk_vcat(after, low, v->vmax);
  if (prev != cur) qk_vemit_push<LOG, SM>(cx, c, (uint32_t)id, cur, after);
}

/* Stock::remove (core.rs:197-204; vector.rs:137-165) */
template <bool LOG, int SM, int VD = 3>
__device__ __forceinline__ void qk_vstock_remove(Ctx& cx, uint32_t sidx, double q, uint64_t src, double* out) {
  const DevComp* c = &cx.comps[sidx];
  const DevHotA ha = QK_HOT_A(sidx);
  const DevVec* v = &cx.vecs[c->vec_idx];
  const uint32_t dim = v->dim;
  double r[VD];
  qk_vload<SM, VD>(cx, v->o64_res, dim, r);
  const double low = QK_U2D(W64(v->o64_low));
  const double before = qk_vtotaln<VD>(r, dim, v->tmask);
  const uint32_t prev = qk_vcat(before, low, v->vmax);
  qk_vstock_area<SM>(cx, sidx, v, before);
  qk_vremoven<VD>(r, q, dim, v->tmask, out);
  qk_vstore<SM, VD>(cx, v->o64_res, dim, r);
  const double after = qk_vtotaln<VD>(r, dim, v->tmask);
  const uint64_t id = qk_log<LOG, SM>(cx, ha, sidx, src, QK_LOG_VSTOCK_REMOVE, 0, QK_D2U(after));
  qk_log_vecn<LOG, SM, VD>(cx, c, sidx, id, 0, out, dim);
  const uint32_t cur = qk_vcat(after, low, v->vmax);
  if (prev != cur) qk_vemit_push<LOG, SM>(cx, c, (uint32_t)id, cur, after);
}

/* DelayModes::get_next_event (delays.rs:144-155): active fix, else the smallest time-until-delay */
template <int SM>
__device__ __forceinline__ uint64_t qk_dm_next_event(Ctx& cx, const DevComp* c, uint32_t fixmask) {
  if (fixmask) return W64(c->o64_dm + (__ffs((int)fixmask) - 1));
  uint64_t best = QK_TIME_NONE;
  for (int k = 0; k < c->n_dm; ++k) {
    const uint64_t d = W64(c->o64_dm + k);
    if (d < best) best = d;
  }
  return best;
}

/* update_state_impl + post_update_state of the five vector process-like components */
template <bool LOG, int SM, int VD = 3>
__device__ __forceinline__ void qk_update_vector(Ctx& cx, const DevComp* c, uint32_t comp, uint64_t src) {
  const DevHotA ha = QK_HOT_A(comp);
  const DevHotB hb = QK_HOT_B(comp);
  const DevVec* v = &cx.vecs[c->vec_idx];
  const uint32_t kind = c->kind, dim = v->dim;
  uint32_t flags = W32(c->o32 + P32_FLAGS);
  uint64_t left = W64(c->o64 + P64_LEFT);
  const uint64_t dt = cx.now - W64(c->o64 + P64_PREV);
  uint64_t ttnpe = QK_TIME_NONE;
  {
    const bool is_in_delay = ((flags >> PF_FIX_SHIFT) & 0xFFu) != 0;
    const bool has_item0 = (flags & PF_HAS_ITEM) != 0;
    const bool is_in_process = has_item0 && !is_in_delay;
    const bool is_env_blocked = (flags & PF_ENV_STOPPED) != 0;
    if (!(is_in_delay || is_env_blocked) && has_item0) {
      left -= (dt < left ? dt : left);
      if (left == 0) {
        flags &= ~PF_HAS_ITEM;
        double r[VD];
        qk_vload<SM, VD>(cx, v->o64_res, dim, r);
        QK_ST32_ADD(comp, ST32_COUNT, 1u);
        if (kind == QK_VECTOR_COMBINER) {
          /* the running total and quantity were accumulated at CombineStart in the reference's order */
          const double quantity = QK_U2D(W64(v->o64_res + dim));
          src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_VPROC_COMBINE_SUCCESS, 0, QK_D2U(quantity));
          qk_log_vecn<LOG, SM, VD>(cx, c, comp, src, 0, r, dim);
          qk_red_addf64(&cx.gs64[(size_t)(comp * ST_PER_COMP + ST64_FQ) * cx.rep_pad], quantity);
          if (c->down >= 0) qk_vstock_add<LOG, SM, VD>(cx, (uint32_t)c->down, r, src);
        } else if (kind == QK_VECTOR_SPLITTER) {
          const double total = qk_vtotaln<VD>(r, dim, v->tmask);
          double parts[5][VD];
          for (uint32_t i = 0; i < v->n_ports; ++i) { /* vector.rs:1152-1155: remove on a fresh clone */
            double clone[VD];
#pragma unroll
            for (int k = 0; k < VD; ++k) clone[k] = r[k];
            qk_vremoven<VD>(clone, __dmul_rn(total, v->ratios[i]), dim, v->tmask, parts[i]);
          }
          src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_VPROC_SPLIT_SUCCESS, 0, QK_D2U(total));
          for (uint32_t i = 0; i < v->n_ports; ++i) qk_log_vecn<LOG, SM, VD>(cx, c, comp, src, i, parts[i], dim);
          qk_red_addf64(&cx.gs64[(size_t)(comp * ST_PER_COMP + ST64_FQ) * cx.rep_pad], total);
          for (uint32_t i = 0; i < v->n_ports; ++i)
            if (v->downs[i] >= 0) {
              qk_vstock_add<LOG, SM, VD>(cx, (uint32_t)v->downs[i], parts[i], src);
              if (cx.status) return;
            }
        } else {
          const double quantity = qk_vtotaln<VD>(r, dim, v->tmask);
          src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_VPROC_SUCCESS, 0, QK_D2U(quantity));
          qk_log_vecn<LOG, SM, VD>(cx, c, comp, src, 0, r, dim);
          qk_red_addf64(&cx.gs64[(size_t)(comp * ST_PER_COMP + ST64_FQ) * cx.rep_pad], quantity);
          if (kind != QK_VECTOR_SINK && c->down >= 0) qk_vstock_add<LOG, SM, VD>(cx, (uint32_t)c->down, r, src);
        }
        if (cx.status) return;
        double zero[VD];
#pragma unroll
        for (int k = 0; k < VD; ++k) zero[k] = 0.0;
        qk_vstore<SM, VD>(cx, v->o64_res, dim, zero);
        if (kind == QK_VECTOR_COMBINER) W64(v->o64_res + dim) = 0;
      }
    }
    /* VectorSource ticks its delay modes even while the environment has stopped it (vector.rs:1444) */
    const bool tick = kind == QK_VECTOR_SOURCE ? (is_in_delay || is_in_process)
                                               : (!is_env_blocked && (is_in_delay || is_in_process));
    if (c->n_dm && tick) {
      qk_tick_delays<LOG, SM>(cx, ha, hb, comp, &flags, dt, &src);
      if (cx.status) return;
    }
  }
  if (c->env >= 0 || (flags & PF_ENV_STOPPED)) qk_refresh_env<LOG, SM>(cx, ha, c->env, comp, &flags, &src);

  const bool has_item = (flags & PF_HAS_ITEM) != 0;
  const uint32_t fixmask = (flags >> PF_FIX_SHIFT) & 0xFFu;
  const bool is_env_stopped = (flags & PF_ENV_STOPPED) != 0;
  const bool has_active_delay = fixmask != 0 || is_env_stopped;
  if (!has_item && !has_active_delay) {
    bool ok = false;
    uint32_t fail = 0;
    if (kind == QK_VECTOR_PROCESS) {
      const uint32_t us = qk_vstock_cat_of<SM, VD>(cx, c->up), ds = qk_vstock_cat_of<SM, VD>(cx, c->down);
      if ((us == 1u || us == 2u) && (ds == 0u || ds == 1u)) ok = true;
      else if (us == 0u) fail = QK_REASON_UPSTREAM_EMPTY;
      else if (us == 3u) fail = QK_REASON_UPSTREAM_NOT_CONNECTED;
      else if (ds == 3u) fail = QK_REASON_DOWNSTREAM_NOT_CONNECTED;
      else fail = QK_REASON_DOWNSTREAM_FULL;
    } else if (kind == QK_VECTOR_SOURCE) {
      const uint32_t ds = qk_vstock_cat_of<SM, VD>(cx, c->down);
      if (ds == 2u) fail = QK_REASON_DOWNSTREAM_FULL;
      else if (ds == 3u) fail = QK_REASON_DOWNSTREAM_NOT_CONNECTED;
      else ok = true;
    } else if (kind == QK_VECTOR_SINK) {
      const uint32_t us = qk_vstock_cat_of<SM, VD>(cx, c->up);
      if (us == 1u || us == 2u) ok = true;
      else if (us == 0u) fail = QK_REASON_UPSTREAM_EMPTY;
      else fail = QK_REASON_UPSTREAM_NOT_CONNECTED;
    } else if (kind == QK_VECTOR_COMBINER) {
      int all_us = 1; /* Some(true) / 0 Some(false) / -1 None (vector.rs:899-914) */
      for (uint32_t i = 0; i < v->n_ports; ++i) {
        const uint32_t us = qk_vstock_cat_of<SM, VD>(cx, v->ups[i]);
        if (us == 3u) {
          all_us = -1;
          break;
        }
        if (!(us == 1u || us == 2u)) all_us = 0;
      }
      const uint32_t ds = qk_vstock_cat_of<SM, VD>(cx, c->down);
      if (all_us == 1 && (ds == 0u || ds == 1u)) ok = true;
      else if (all_us == 0) fail = QK_REASON_AN_UPSTREAM_EMPTY;
      else if (all_us == -1) fail = QK_REASON_NO_UPSTREAMS;
      else if (ds == 3u) fail = QK_REASON_DOWNSTREAM_NOT_CONNECTED;
      else fail = QK_REASON_DOWNSTREAM_FULL;
    } else { /* splitter (vector.rs:1200-1245) */
      const uint32_t us = qk_vstock_cat_of<SM, VD>(cx, c->up);
      bool all_ds = true;
      for (uint32_t i = 0; i < v->n_ports; ++i) {
        const uint32_t ds = qk_vstock_cat_of<SM, VD>(cx, v->downs[i]);
        if (!(ds == 1u || ds == 0u)) all_ds = false;
      }
      if ((us == 2u || us == 1u) && all_ds) ok = true;
      else if (!all_ds) fail = QK_REASON_A_DOWNSTREAM_FULL;
      else if (us == 3u) fail = QK_REASON_UPSTREAM_NOT_CONNECTED;
      else fail = QK_REASON_UPSTREAM_EMPTY;
    }
    if (ok) {
      /* process_quantity_distr: the model's, or the instance the latest SetProcessQuantity event brought (core.rs:1387-1391) */
      const double q = qk_sample<SM>(cx, v->o32_qty != 0xFFFF ? W32(v->o32_qty) : (uint32_t)v->dist_qty);
      if (cx.status) return;
      double got[5][VD];
      uint32_t ngot = 1, start_kind = QK_LOG_VPROC_START;
      if (kind == QK_VECTOR_SOURCE) {
        const double st = qk_vtotaln<VD>(v->vinit, dim, v->tmask);
        if (!(st > 0.0)) {
          cx.status = QK_REP_SOURCE_VECTOR; /* panic!("Source vector has total 0 or negative") vector.rs:1490 */
          return;
        }
        const double factor = __ddiv_rn(q, st);
#pragma unroll
        for (int k = 0; k < VD; ++k) got[0][k] = (uint32_t)k < dim ? __dmul_rn(v->vinit[k], factor) : 0.0;
      } else {
        src = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_WITHDRAW_REQUEST, 0, 0);
        if (kind == QK_VECTOR_COMBINER) {
          start_kind = QK_LOG_VPROC_COMBINE_START;
          ngot = v->n_ports;
          for (uint32_t i = 0; i < ngot; ++i) { /* the full quantity from EACH upstream (quirk Q7) */
            qk_vstock_remove<LOG, SM, VD>(cx, (uint32_t)v->ups[i], q, src, got[i]);
            if (cx.status) return;
          }
        } else {
          if (kind == QK_VECTOR_SPLITTER) start_kind = QK_LOG_VPROC_SPLIT_START;
          qk_vstock_remove<LOG, SM, VD>(cx, (uint32_t)c->up, q, src, got[0]);
          if (cx.status) return;
        }
      }
      const uint64_t d = qk_sample_duration<SM>(cx, c->dist_time);
      if (cx.status) return;
      left = d;
      flags |= PF_HAS_ITEM;
      QK_ST64_ADD(comp, ST64_TIME, d);
      if (kind == QK_VECTOR_COMBINER) {
        double total[VD];
#pragma unroll
        for (int k = 0; k < VD; ++k) total[k] = 0.0;
        double qsum = 0.0;
        for (uint32_t i = 0; i < ngot; ++i) {
          for (uint32_t k = 0; k < dim; ++k) total[k] = __dadd_rn(total[k], got[i][k]);
          qsum = __dadd_rn(qsum, qk_vtotaln<VD>(got[i], dim, v->tmask));
        }
        qk_vstore<SM, VD>(cx, v->o64_res, dim, total);
        W64(v->o64_res + dim) = QK_D2U(qsum);
      } else {
        qk_vstore<SM, VD>(cx, v->o64_res, dim, got[0]);
      }
      const uint64_t id = qk_log<LOG, SM>(cx, ha, comp, src, start_kind, 0, QK_D2U(q));
      for (uint32_t i = 0; i < ngot; ++i) qk_log_vecn<LOG, SM, VD>(cx, c, comp, id, i, got[i], dim);
      if (kind != QK_VECTOR_SINK) src = id; /* VectorSink drops the returned id (vector.rs:1747) */
      ttnpe = d;
    } else {
      const uint64_t id = qk_log<LOG, SM>(cx, ha, comp, src, QK_LOG_VPROC_FAILURE, fail, 0);
      if (kind != QK_VECTOR_SINK) src = id;
    }
  } else if (has_item && !has_active_delay) {
    ttnpe = left;
  } else if (kind != QK_VECTOR_SINK) {
    ttnpe = fixmask ? W64(c->o64_dm + (__ffs((int)fixmask) - 1)) : QK_TIME_NONE;
  }
  /* flags may have gained PF_HAS_ITEM above */
  const bool has_item2 = (flags & PF_HAS_ITEM) != 0;
  uint64_t ttnde = QK_TIME_NONE;
  if (c->n_dm && (has_item2 || has_active_delay || !is_env_stopped)) ttnde = qk_dm_next_event<SM>(cx, c, fixmask);
  uint64_t next = QK_TIME_NONE;
  if (kind == QK_VECTOR_SINK) { /* vector.rs:1780-1786 */
    const bool active = fixmask != 0;
    if (ttnpe != QK_TIME_NONE && ttnde != QK_TIME_NONE && active) next = ttnpe < ttnde ? ttnpe : ttnde;
    else if (ttnpe != QK_TIME_NONE && ttnde == QK_TIME_NONE) next = ttnpe;
    else if (ttnpe == QK_TIME_NONE && ttnde != QK_TIME_NONE && active) next = ttnde;
  } else {
    next = ttnpe < ttnde ? ttnpe : ttnde; /* min of those that exist (vector.rs:495) */
  }
  W32(c->o32 + P32_FLAGS) = flags;
  W64(c->o64 + P64_LEFT) = left;
  qk_post<LOG, SM>(cx, ha, next, src);
  W64(c->o64 + P64_PREV) = cx.now;
}

#endif
