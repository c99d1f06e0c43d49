// pb_comm (one NCCL communicator rank) and pb_group (single-process, one pb_index replica per device)
// halves of the C ABI.  See group.hpp for why NCCL is bound with dlopen.
#include <dlfcn.h>

#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <nccl.h>

#include "common.hpp"
#include "group.hpp"

namespace {

struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  const char* (*GetLastError)(ncclComm_t) = nullptr;
  std::string why;
};

NcclApi* nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {std::getenv("PB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
      api.why = dlerror() ? dlerror() : "dlopen failed";
    }
    if (!api.h) return;
    auto sym = [&](const char* s) { return dlsym(api.h, s); };
#define PB_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(sym(name))
    PB_SYM(GetVersion, "ncclGetVersion");
    PB_SYM(GetUniqueId, "ncclGetUniqueId");
    PB_SYM(CommInitRank, "ncclCommInitRank");
    PB_SYM(CommInitAll, "ncclCommInitAll");
    PB_SYM(CommDestroy, "ncclCommDestroy");
    PB_SYM(CommAbort, "ncclCommAbort");
    PB_SYM(AllGather, "ncclAllGather");
    PB_SYM(GroupStart, "ncclGroupStart");
    PB_SYM(GroupEnd, "ncclGroupEnd");
    PB_SYM(GetErrorString, "ncclGetErrorString");
    PB_SYM(GetLastError, "ncclGetLastError");
#undef PB_SYM
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.CommDestroy || !api.AllGather ||
        !api.GroupStart || !api.GroupEnd || !api.GetErrorString) {
      api.why = "libnccl lacks a required symbol";
      api.h = nullptr;
    }
  });
  return api.h ? &api : nullptr;
}

int need_nccl(NcclApi** out) {
  NcclApi* a = nccl();
  if (!a) {
    pb::set_error("NCCL is not available (libnccl.so.2 could not be loaded); multi-GPU entry points need it");
    return PB_ERR_UNSUPPORTED;
  }
  *out = a;
  return PB_OK;
}

#define NC(x)                                                                                          \
  do {                                                                                                 \
    ncclResult_t r_ = (x);                                                                             \
    if (r_ != ncclSuccess) {                                                                           \
      pb::set_error("%s failed: %s (%s:%d)", #x, api->GetErrorString(r_), __FILE__, __LINE__);         \
      return PB_ERR_CUDA;                                                                              \
    }                                                                                                  \
  } while (0)
#define CUG(x)                                                                                         \
  do {                                                                                                 \
    cudaError_t e_ = (x);                                                                              \
    if (e_ != cudaSuccess) {                                                                           \
      pb::set_error("%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__);          \
      return (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? PB_ERR_NO_DEVICE : PB_ERR_CUDA; \
    }                                                                                                  \
  } while (0)

}  // namespace

struct pb_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  bool owned = true;      // destroyed by pb_comm_destroy (members of a pb_group are destroyed by the group)
};

namespace pbg {

int comm_allgather(pb_comm* c, const void* send, void* recv, size_t bytes, cudaStream_t st) {
  NcclApi* api = nullptr;
  int rc = need_nccl(&api);
  if (rc != PB_OK) return rc;
  NC(api->AllGather(send, recv, bytes, ncclUint8, c->comm, st));
  return PB_OK;
}
int comm_world(const pb_comm* c) { return c->world; }
int comm_rank(const pb_comm* c) { return c->rank; }
int comm_device(const pb_comm* c) { return c->device; }
int nccl_group_start() {
  NcclApi* api = nullptr;
  int rc = need_nccl(&api);
  if (rc != PB_OK) return rc;
  NC(api->GroupStart());
  return PB_OK;
}
int nccl_group_end() {
  NcclApi* api = nullptr;
  int rc = need_nccl(&api);
  if (rc != PB_OK) return rc;
  NC(api->GroupEnd());
  return PB_OK;
}

}  // namespace pbg

static_assert(PB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "pb_comm id size");

extern "C" {

int pb_comm_unique_id(uint8_t* id) {
  if (!id) { pb::set_error("pb_comm_unique_id: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    NcclApi* api = nullptr;
    int rc = need_nccl(&api);
    if (rc != PB_OK) return rc;
    ncclUniqueId u;
    NC(api->GetUniqueId(&u));
    std::memcpy(id, u.internal, PB_COMM_ID_BYTES);
    return PB_OK;
  });
}

int pb_comm_create(const uint8_t* id, int rank, int world, int device, pb_comm** out) {
  if (!id || !out || world < 1 || rank < 0 || rank >= world) { pb::set_error("pb_comm_create: bad argument"); return PB_ERR_INVALID; }
  PB_TRY({
    NcclApi* api = nullptr;
    int rc = need_nccl(&api);
    if (rc != PB_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); pb::set_error("no CUDA device is available"); return PB_ERR_NO_DEVICE; }
    if (device < 0 || device >= ndev) { pb::set_error("pb_comm_create: device %d out of range", device); return PB_ERR_INVALID; }
    CUG(cudaSetDevice(device));
    ncclUniqueId u;
    std::memcpy(u.internal, id, PB_COMM_ID_BYTES);
    pb_comm* c = new pb_comm();
    c->rank = rank; c->world = world; c->device = device;
    ncclResult_t r = api->CommInitRank(&c->comm, world, u, rank);
    if (r != ncclSuccess) {
      pb::set_error("ncclCommInitRank failed: %s", api->GetErrorString(r));
      delete c;
      return PB_ERR_CUDA;
    }
    *out = c;
    return PB_OK;
  });
}

int pb_comm_info(const pb_comm* c, int* rank, int* world, int* nccl_version) {
  if (!c) { pb::set_error("pb_comm_info: null argument"); return PB_ERR_INVALID; }
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (nccl_version) {
    *nccl_version = 0;
    NcclApi* api = nccl();
    if (api && api->GetVersion) api->GetVersion(nccl_version);
  }
  return PB_OK;
}

void pb_comm_destroy(pb_comm* c) {
  if (!c) return;
  NcclApi* api = nccl();
  if (api && c->comm) {
    cudaSetDevice(c->device);
    api->CommDestroy(c->comm);
  }
  delete c;
}

}  // extern "C"

// ==========================================================================================
// pb_group: ONE process, one replica of the image per device (SURVEY §8b "pb_group_create /
// pb_group_query_batch").  A batch is cut into contiguous query blocks, member r runs block r on its
// device from its own host thread, the packed result blocks are all-gathered (ncclCommInitAll
// communicators, one ncclAllGather per member inside one group call) and member 0's gathered copy is
// what the caller receives.
// ==========================================================================================
struct pb_group {
  std::vector<int> devices;
  std::vector<pb_index*> ix;
  std::vector<pb_comm*> comm;
  std::vector<pb_batch*> batch;
  std::mutex mu;
};

extern "C" {

int pb_group_create(const pb_index_image* image, const int* devices, int n, pb_group** out) {
  if (!image || !devices || !out || n < 1) { pb::set_error("pb_group_create: bad argument"); return PB_ERR_INVALID; }
  PB_TRY({
    NcclApi* api = nullptr;
    int rc = need_nccl(&api);
    if (rc != PB_OK) return rc;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < i; ++j)
        if (devices[i] == devices[j]) { pb::set_error("pb_group_create: device %d listed twice", devices[i]); return PB_ERR_INVALID; }
    pb_group* g = new pb_group();
    g->devices.assign(devices, devices + n);
    g->ix.assign(n, nullptr); g->comm.assign(n, nullptr); g->batch.assign(n, nullptr);
    auto fail = [&](int code) { pb_group_destroy(g); return code; };
    // replicas: one upload per device, in parallel host threads (the uploads are independent)
    std::vector<int> rcs(n, PB_OK);
    std::vector<std::string> errs(n);
    {
      std::vector<std::thread> th;
      for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
          rcs[i] = pb_index_create(image, g->devices[i], &g->ix[i]);
          if (rcs[i] != PB_OK) errs[i] = pb::get_error();
        });
      for (auto& t : th) t.join();
    }
    for (int i = 0; i < n; ++i)
      if (rcs[i] != PB_OK) { pb::set_error("pb_group_create: device %d: %s", g->devices[i], errs[i].c_str()); return fail(rcs[i]); }
    std::vector<ncclComm_t> comms(n, nullptr);
    ncclResult_t r = api->CommInitAll(comms.data(), n, g->devices.data());
    if (r != ncclSuccess) { pb::set_error("ncclCommInitAll failed: %s", api->GetErrorString(r)); return fail(PB_ERR_CUDA); }
    for (int i = 0; i < n; ++i) {
      g->comm[i] = new pb_comm();
      g->comm[i]->comm = comms[i]; g->comm[i]->rank = i; g->comm[i]->world = n; g->comm[i]->device = g->devices[i];
      g->comm[i]->owned = false;
    }
    *out = g;
    return PB_OK;
  });
}

void pb_group_destroy(pb_group* g) {
  if (!g) return;
  for (auto* b : g->batch) if (b) pb_batch_destroy(b);
  for (auto* c : g->comm) if (c) pb_comm_destroy(c);
  for (auto* ix : g->ix) if (ix) pb_index_destroy(ix);
  delete g;
}

int pb_group_size(const pb_group* g) { return g ? (int)g->ix.size() : 0; }

int pb_group_set_live_state(pb_group* g, const uint32_t* removed_ords, uint64_t n_removed, uint64_t n_live_docs,
                            const double* field_avg) {
  if (!g) { pb::set_error("pb_group_set_live_state: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    std::lock_guard<std::mutex> lk(g->mu);
    for (auto* ix : g->ix) {
      int rc = pb_index_set_live_state(ix, removed_ords, n_removed, n_live_docs, field_avg);
      if (rc != PB_OK) return rc;
    }
    return PB_OK;
  });
}

int pb_group_query_batch(pb_group* g, const pb_query_batch_desc* q, pb_query_results* out) {
  if (!g || !q || !out) { pb::set_error("pb_group_query_batch: null argument"); return PB_ERR_INVALID; }
  PB_TRY({
    std::lock_guard<std::mutex> lk(g->mu);
    const int n = (int)g->ix.size();
    const uint64_t Q = q->n_queries;
    if (Q && (!q->query_term_off || !q->term_byte_off)) { pb::set_error("pb_group_query_batch: null argument"); return PB_ERR_INVALID; }
    const uint64_t slot = (Q + (uint64_t)n - 1) / (uint64_t)n;      // queries per member (the last block may be short)
    // per-member descriptors: contiguous query blocks with offsets rebased to 0
    struct Shard { std::vector<uint64_t> qoff, toff; pb_query_batch_desc d; };
    std::vector<Shard> sh(n);
    for (int r = 0; r < n; ++r) {
      const uint64_t lo = std::min<uint64_t>(Q, slot * r), hi = std::min<uint64_t>(Q, slot * (r + 1));
      Shard& s = sh[r];
      s.d = *q;
      s.d.n_queries = hi - lo;
      const uint64_t t0 = Q ? q->query_term_off[lo] : 0, t1 = Q ? q->query_term_off[hi] : 0;
      if (t1 < t0) { pb::set_error("query batch: query_term_off not monotone"); return PB_ERR_INVALID; }
      s.qoff.resize(hi - lo + 1);
      for (uint64_t i = lo; i <= hi && Q; ++i) s.qoff[i - lo] = q->query_term_off[i] - t0;
      if (!Q) s.qoff[0] = 0;
      s.toff.resize(t1 - t0 + 1);
      const uint64_t b0 = (t1 > t0) ? q->term_byte_off[t0] : 0;
      for (uint64_t i = t0; i <= t1 && t1 > t0; ++i) s.toff[i - t0] = q->term_byte_off[i] - b0;
      if (t1 == t0) s.toff[0] = 0;
      s.d.query_term_off = s.qoff.data();
      s.d.term_byte_off = s.toff.data();
      s.d.term_bytes = q->term_bytes ? q->term_bytes + b0 : nullptr;
    }
    // stage + run every block from its own host thread; the gather is issued afterwards by THIS thread
    // for all members inside one NCCL group call, so a block that fails cannot leave its peers blocked
    // inside a collective
    std::vector<int> rcs(n, PB_OK);
    std::vector<std::string> errs(n);
    {
      std::vector<std::thread> th;
      for (int r = 0; r < n; ++r)
        th.emplace_back([&, r] {
          int rc = PB_OK;
          if (g->batch[r]) rc = pb_batch_reload(g->batch[r], &sh[r].d);
          else rc = pb_batch_create(g->ix[r], &sh[r].d, &g->batch[r]);
          if (rc == PB_OK) rc = pb_batch_set_gather(g->batch[r], g->comm[r], slot);
          if (rc == PB_OK) rc = pb_batch_run_local(g->batch[r]);
          rcs[r] = rc;
          if (rc != PB_OK) errs[r] = pb::get_error();
        });
      for (auto& t : th) t.join();
    }
    for (int r = 0; r < n; ++r)
      if (rcs[r] != PB_OK) { pb::set_error("pb_group_query_batch: member %d (device %d): %s", r, g->devices[r], errs[r].c_str()); return rcs[r]; }
    int rc = pbg::nccl_group_start();
    if (rc != PB_OK) return rc;
    for (int r = 0; r < n && rc == PB_OK; ++r) rc = pb_batch_gather(g->batch[r]);
    int rc2 = pbg::nccl_group_end();
    if (rc != PB_OK) return rc;
    if (rc2 != PB_OK) return rc2;
    for (int r = 1; r < n; ++r) {
      rc = pb_batch_sync(g->batch[r]);
      if (rc != PB_OK) return rc;
    }
    // member 0's gathered copy, compacted from [n][slot] to [Q]
    return pb_batch_fetch_gathered(g->batch[0], Q, out);
  });
}

int pb_group_member_stats(pb_group* g, int member, pb_batch_stats* out) {
  if (!g || !out || member < 0 || member >= (int)g->batch.size() || !g->batch[member]) { pb::set_error("pb_group_member_stats: bad argument"); return PB_ERR_INVALID; }
  return pb_batch_get_stats(g->batch[member], out);
}

}  // extern "C"
