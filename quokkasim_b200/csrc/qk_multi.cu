/*
 * qk_multi.cu -- the multi-GPU plane behind the C ABI: qk_multi_* runs ONE job of n replications on
 * several GPUs of one box from a single host process, without Python.
 *
 * Replications are independent, so the data path has no collective: device d of D owns the contiguous range
 * [n*d/D, n*(d+1)/D) of replication indices (= its Philox counter range), as a qk_batch on its own stream.  The one
 * exchange is the final reduction of summary statistics in qk_multi_stats: every device reduces its replications to the
 * packed partials (exact u64 words, quokka_b200.h QP_*), ONE ncclAllGather brings all D arrays to every device, a
 * small kernel combines them (unsigned min on the extrema, sum elsewhere), every device bins its replications over the
 * now global ranges and ONE ncclAllReduce(sum) adds the histograms; device 0's copy goes to the host.
 *
 * NCCL is loaded at run time (dlopen "libnccl.so.2") the first time a job with more than one device is created:
 * single-GPU users of the library do not need it, and a process that already carries a libnccl.so.2 (PyTorch's) shares it.
 */
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "qk_model.h"

void qk_fill_comp_summary(const uint64_t* w, const uint64_t* hist, qk_comp_summary* out); /* qk_batch.cu */

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
    api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
#define QK_SYM(field, sym)                                           \
  api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym)); \
  if (!api.field) {                                                  \
    api.lib = nullptr;                                               \
    return nullptr;                                                  \
  }
  QK_SYM(CommInitAll, "ncclCommInitAll")
  QK_SYM(CommDestroy, "ncclCommDestroy")
  QK_SYM(AllGather, "ncclAllGather")
  QK_SYM(AllReduce, "ncclAllReduce")
  QK_SYM(GroupStart, "ncclGroupStart")
  QK_SYM(GroupEnd, "ncclGroupEnd")
  QK_SYM(GetErrorString, "ncclGetErrorString")
#undef QK_SYM
  return &api;
}

}  // namespace

struct qk_multi {
  std::vector<int> devices;
  std::vector<qk_batch*> shards;
  std::vector<uint64_t> rep0, n;
  std::vector<cudaStream_t> streams;
  std::vector<ncclComm_t> comms;
  std::vector<uint64_t*> d_mine, d_gathered, d_both; /* per device: own partials ; all partials ; combined + histograms */
  uint32_t n_comp = 0;
  uint64_t n_reps = 0;
};

#define MULTI_CUDA(expr)                                                                   \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      qk_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return QK_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)
#define MULTI_NCCL(api, expr)                                                                        \
  do {                                                                                               \
    ncclResult_t _r = (expr);                                                                        \
    if (_r != ncclSuccess) {                                                                         \
      qk_set_error("%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), __FILE__, __LINE__);   \
      return QK_ERR_CUDA;                                                                            \
    }                                                                                                \
  } while (0)

extern "C" {

void qk_multi_destroy(qk_multi* mg) {
  if (!mg) return;
  NcclApi* api = mg->comms.empty() ? nullptr : nccl_api();
  for (size_t d = 0; d < mg->devices.size(); ++d) {
    cudaSetDevice(mg->devices[d]);
    if (d < mg->shards.size() && mg->shards[d]) qk_batch_destroy(mg->shards[d]);
    if (d < mg->d_mine.size()) cudaFree(mg->d_mine[d]);
    if (d < mg->d_gathered.size()) cudaFree(mg->d_gathered[d]);
    if (d < mg->d_both.size()) cudaFree(mg->d_both[d]);
    if (api && d < mg->comms.size() && mg->comms[d]) api->CommDestroy(mg->comms[d]);
    if (d < mg->streams.size() && mg->streams[d]) cudaStreamDestroy(mg->streams[d]);
  }
  delete mg;
}

int qk_multi_create(const qk_model* m, const qk_multi_config* cfg, qk_multi** out) {
  if (!m || !cfg || !out || cfg->n_reps == 0) {
    qk_set_error("qk_multi_create: NULL argument or n_reps == 0");
    return QK_ERR_INVALID_ARG;
  }
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
    cudaGetLastError();
    qk_set_error("qk_multi_create: no CUDA device (this library has no CPU fallback)");
    return QK_ERR_NO_DEVICE;
  }
  qk_multi* mg = new qk_multi();
  if (cfg->devices && cfg->n_devices) {
    for (uint32_t i = 0; i < cfg->n_devices; ++i) {
      if (cfg->devices[i] < 0 || cfg->devices[i] >= visible) {
        qk_set_error("qk_multi_create: device %d out of range (%d devices)", cfg->devices[i], visible);
        delete mg;
        return QK_ERR_INVALID_ARG;
      }
      mg->devices.push_back(cfg->devices[i]);
    }
  } else {
    for (int i = 0; i < visible; ++i) mg->devices.push_back(i);
  }
  const size_t D = mg->devices.size();
  if (cfg->n_reps < D) {
    qk_set_error("qk_multi_create: fewer replications (%llu) than devices (%zu)", (unsigned long long)cfg->n_reps, D);
    delete mg;
    return QK_ERR_INVALID_ARG;
  }
  uint32_t n_comp = 0;
  qk_model_n_components(m, &n_comp);
  mg->n_comp = n_comp;
  mg->n_reps = cfg->n_reps;
  mg->shards.assign(D, nullptr);
  mg->streams.assign(D, nullptr);
  mg->rep0.resize(D);
  mg->n.resize(D);
  mg->d_mine.assign(D, nullptr);
  mg->d_gathered.assign(D, nullptr);
  mg->d_both.assign(D, nullptr);
  /* one host thread per GPU for the creation (gigabytes of allocation and the table upload per device) */
  std::vector<int> rcs(D, QK_OK);
  std::vector<std::string> errs(D);
  std::vector<std::thread> th;
  for (size_t d = 0; d < D; ++d)
    th.emplace_back([&, d]() {
      const size_t words = (size_t)n_comp * QK_PARTIAL_WORDS, hist = (size_t)n_comp * 2 * QK_HIST_BINS;
      cudaError_t e = cudaSetDevice(mg->devices[d]);
      if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&mg->streams[d], cudaStreamNonBlocking);
      if (e == cudaSuccess) e = cudaMalloc(&mg->d_mine[d], words * 8);
      if (e == cudaSuccess) e = cudaMalloc(&mg->d_gathered[d], words * 8 * D);
      if (e == cudaSuccess) e = cudaMalloc(&mg->d_both[d], (words + hist) * 8);
      if (e != cudaSuccess) {
        rcs[d] = QK_ERR_CUDA;
        errs[d] = std::string("qk_multi_create: ") + cudaGetErrorString(e);
        return;
      }
      qk_batch_config bc;
      memset(&bc, 0, sizeof(bc));
      mg->rep0[d] = cfg->n_reps * d / D;
      mg->n[d] = cfg->n_reps * (d + 1) / D - mg->rep0[d];
      bc.n_reps = mg->n[d];
      bc.rep_offset = cfg->rep_offset + mg->rep0[d];
      bc.base_seed = cfg->base_seed;
      bc.device = mg->devices[d];
      bc.log_capacity = cfg->log_capacity;
      bc.threads_per_block = cfg->threads_per_block;
      bc.flags = cfg->flags;
      bc.lanes_per_replication = cfg->lanes_per_replication;
      bc.stream = mg->streams[d];
      rcs[d] = qk_batch_create(m, &bc, &mg->shards[d]);
      if (rcs[d]) errs[d] = qk_last_error(); /* thread-local: carried over to the calling thread below */
    });
  for (auto& t : th) t.join();
  for (size_t d = 0; d < D; ++d)
    if (rcs[d]) {
      qk_set_error("device %d: %s", mg->devices[d], errs[d].c_str());
      const int rc = rcs[d];
      qk_multi_destroy(mg);
      return rc;
    }
  if (D > 1) {
    NcclApi* api = nccl_api();
    if (!api) {
      qk_set_error("qk_multi_create: %zu devices need NCCL for the statistics reduction, and libnccl.so.2 could not be "
                   "loaded (%s)", D, dlerror() ? dlerror() : "missing symbol");
      qk_multi_destroy(mg);
      return QK_ERR_STATE;
    }
    mg->comms.assign(D, nullptr);
    ncclResult_t r = api->CommInitAll(mg->comms.data(), (int)D, mg->devices.data());
    if (r != ncclSuccess) {
      qk_set_error("ncclCommInitAll failed: %s", api->GetErrorString(r));
      mg->comms.clear();
      qk_multi_destroy(mg);
      return QK_ERR_CUDA;
    }
  }
  *out = mg;
  return QK_OK;
}

int qk_multi_n_shards(qk_multi* mg, uint32_t* n) {
  if (!mg || !n) return QK_ERR_INVALID_ARG;
  *n = (uint32_t)mg->shards.size();
  return QK_OK;
}

int qk_multi_shard(qk_multi* mg, uint32_t i, qk_batch** batch, uint64_t* rep0, uint64_t* n_reps, int32_t* device) {
  if (!mg || i >= mg->shards.size()) {
    qk_set_error("qk_multi_shard: shard %u out of range", i);
    return QK_ERR_INVALID_ARG;
  }
  if (batch) *batch = mg->shards[i];
  if (rep0) *rep0 = mg->rep0[i];
  if (n_reps) *n_reps = mg->n[i];
  if (device) *device = mg->devices[i];
  return QK_OK;
}

int qk_multi_set_tape(qk_multi* mg, uint32_t dist_id, const double* values, uint64_t per_rep_len) {
  if (!mg || (!values && per_rep_len)) return QK_ERR_INVALID_ARG;
  for (size_t d = 0; d < mg->shards.size(); ++d) {
    const int rc = qk_batch_set_tape(mg->shards[d], dist_id, values ? values + mg->rep0[d] * per_rep_len : nullptr, per_rep_len);
    if (rc) return rc;
  }
  return QK_OK;
}

#define QK_FOR_SHARDS(call)                          \
  if (!mg) return QK_ERR_INVALID_ARG;                \
  for (size_t d = 0; d < mg->shards.size(); ++d) {   \
    const int rc = call;                             \
    if (rc) return rc;                               \
  }                                                  \
  return QK_OK;

/* every call below only ENQUEUES on the devices' streams, so one host thread keeps all GPUs busy */
int qk_multi_init(qk_multi* mg, uint64_t t0_ns) { QK_FOR_SHARDS(qk_batch_init(mg->shards[d], t0_ns)) }
int qk_multi_step_until(qk_multi* mg, uint64_t t_ns) { QK_FOR_SHARDS(qk_batch_step_until(mg->shards[d], t_ns)) }
int qk_multi_step_events(qk_multi* mg, uint64_t n_events) { QK_FOR_SHARDS(qk_batch_step_events(mg->shards[d], n_events)) }
int qk_multi_synchronize(qk_multi* mg) { QK_FOR_SHARDS(qk_batch_synchronize(mg->shards[d])) }

int qk_multi_last_step_ms(qk_multi* mg, float* ms_max) {
  if (!mg || !ms_max) return QK_ERR_INVALID_ARG;
  *ms_max = 0;
  for (size_t d = 0; d < mg->shards.size(); ++d) {
    float ms = 0;
    const int rc = qk_batch_last_step_ms(mg->shards[d], &ms, nullptr);
    if (rc) return rc;
    if (ms > *ms_max) *ms_max = ms; /* device-timed, max over the GPUs */
  }
  return QK_OK;
}

int qk_multi_stats(qk_multi* mg, qk_comp_summary* out, uint32_t n_comp, qk_batch_summary* total) {
  if (!mg || !out || n_comp != mg->n_comp) {
    qk_set_error("qk_multi_stats: invalid argument (n_comp must equal the model's)");
    return QK_ERR_INVALID_ARG;
  }
  const size_t D = mg->shards.size();
  const size_t words = (size_t)n_comp * QK_PARTIAL_WORDS, hist = (size_t)n_comp * 2 * QK_HIST_BINS;
  NcclApi* api = D > 1 ? nccl_api() : nullptr;
  for (size_t d = 0; d < D; ++d) {
    const int rc = qk_batch_reduce_partials_device(mg->shards[d], D > 1 ? mg->d_mine[d] : mg->d_gathered[d], n_comp);
    if (rc) return rc;
  }
  if (D > 1) { /* the one exchange: every device gets every device's partials */
    MULTI_NCCL(api, api->GroupStart());
    for (size_t d = 0; d < D; ++d)
      MULTI_NCCL(api, api->AllGather(mg->d_mine[d], mg->d_gathered[d], words, ncclUint64, mg->comms[d], mg->streams[d]));
    MULTI_NCCL(api, api->GroupEnd());
  }
  for (size_t d = 0; d < D; ++d) {
    int rc = qk_batch_combine_partials_device(mg->shards[d], mg->d_gathered[d], (uint32_t)D, mg->d_both[d], n_comp);
    if (rc) return rc;
    rc = qk_batch_histograms_device(mg->shards[d], mg->d_both[d], mg->d_both[d] + words, n_comp);
    if (rc) return rc;
  }
  if (D > 1) {
    MULTI_NCCL(api, api->GroupStart());
    for (size_t d = 0; d < D; ++d)
      MULTI_NCCL(api, api->AllReduce(mg->d_both[d] + words, mg->d_both[d] + words, hist, ncclUint64, ncclSum, mg->comms[d],
                                     mg->streams[d]));
    MULTI_NCCL(api, api->GroupEnd());
  }
  std::vector<uint64_t> host(words + hist);
  MULTI_CUDA(cudaSetDevice(mg->devices[0]));
  MULTI_CUDA(cudaMemcpyAsync(host.data(), mg->d_both[0], host.size() * 8, cudaMemcpyDeviceToHost, mg->streams[0]));
  for (size_t d = 0; d < D; ++d) {
    MULTI_CUDA(cudaSetDevice(mg->devices[d]));
    MULTI_CUDA(cudaStreamSynchronize(mg->streams[d]));
  }
  for (uint32_t c = 0; c < n_comp; ++c)
    qk_fill_comp_summary(host.data() + (size_t)c * QK_PARTIAL_WORDS, host.data() + words + (size_t)c * 2 * QK_HIST_BINS, &out[c]);
  if (total) {
    memset(total, 0, sizeof(*total));
    total->n_reps = host[QP_NREPS];
    total->n_events = host[QP_NEVENTS];
    total->n_faulted = host[QP_NFAULTED];
    total->min_now_ns = ~0ull;
    for (size_t d = 0; d < D; ++d) { /* clocks and record counts are not part of the partials: small per-shard reads */
      qk_batch_summary s;
      const int rc = qk_batch_summary_get(mg->shards[d], &s);
      if (rc) return rc;
      total->n_log_records += s.n_log_records;
      if (s.min_now_ns < total->min_now_ns) total->min_now_ns = s.min_now_ns;
      if (s.max_now_ns > total->max_now_ns) total->max_now_ns = s.max_now_ns;
    }
  }
  return QK_OK;
}

} /* extern "C" */
