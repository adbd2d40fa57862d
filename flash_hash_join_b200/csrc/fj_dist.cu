// fj_dist.cu — NCCL (dlopen) wrappers.  No reference counterpart: hash_join.cpp is single-process.
#include "fj_dist.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

namespace fj {

enum { ERR_NCCL = -3 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char* (*GetErrorString)(ncclResult_t);
};

static NcclApi g_api;
static bool g_api_ok = false;
static std::mutex g_api_mu;

static fj_status_t load_api() {
  std::lock_guard<std::mutex> lk(g_api_mu);
  if (g_api_ok) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return set_err(ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define FJ_SYM(field, name)                                                        \
  do {                                                                             \
    *reinterpret_cast<void**>(&g_api.field) = dlsym(h, name);                      \
    if (!g_api.field) return set_err(ERR_NCCL, "libnccl lacks symbol %s", name);   \
  } while (0)
  FJ_SYM(GetUniqueId, "ncclGetUniqueId");
  FJ_SYM(CommInitRank, "ncclCommInitRank");
  FJ_SYM(CommDestroy, "ncclCommDestroy");
  FJ_SYM(Broadcast, "ncclBroadcast");
  FJ_SYM(AllReduce, "ncclAllReduce");
  FJ_SYM(AllGather, "ncclAllGather");
  FJ_SYM(Send, "ncclSend");
  FJ_SYM(Recv, "ncclRecv");
  FJ_SYM(GroupStart, "ncclGroupStart");
  FJ_SYM(GroupEnd, "ncclGroupEnd");
  FJ_SYM(GetErrorString, "ncclGetErrorString");
#undef FJ_SYM
  g_api_ok = true;
  return 0;
}

#define FJ_NCCL(expr)                                                                                        \
  do {                                                                                                       \
    ncclResult_t r__ = (expr);                                                                               \
    if (r__ != ncclSuccess)                                                                                  \
      return set_err(ERR_NCCL, "%s failed: %s", #expr, g_api.GetErrorString ? g_api.GetErrorString(r__) : "?"); \
  } while (0)

static_assert(sizeof(ncclUniqueId) == 128, "fj_comm_unique_id hands out 128 bytes");

fj_status_t dist_unique_id(void* id128) {
  fj_status_t s = load_api();
  if (s) return s;
  ncclUniqueId id;
  FJ_NCCL(g_api.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

fj_status_t dist_init(DistState& d, int rank, int world, const void* id128) {
  fj_status_t s = load_api();
  if (s) return s;
  if (d.ready) dist_destroy(d);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm;
  FJ_NCCL(g_api.CommInitRank(&comm, world, id, rank));
  d.comm = comm;
  d.rank = rank;
  d.world = world;
  d.api = &g_api;
  d.ready = true;
  return 0;
}

void dist_destroy(DistState& d) {
  if (d.ready && d.comm) g_api.CommDestroy(reinterpret_cast<ncclComm_t>(d.comm));
  d.comm = nullptr;
  d.ready = false;
  d.world = 1;
  d.rank = 0;
}

fj_status_t dist_broadcast_u64(DistState& d, void* buf, size_t count, int root, cudaStream_t st) {
  FJ_NCCL(g_api.Broadcast(buf, buf, count, ncclUint64, root, reinterpret_cast<ncclComm_t>(d.comm), st));
  return 0;
}
fj_status_t dist_broadcast_oop_u64(DistState& d, const void* send, void* recv, size_t count, int root, cudaStream_t st) {
  FJ_NCCL(g_api.Broadcast(send, recv, count, ncclUint64, root, reinterpret_cast<ncclComm_t>(d.comm), st));
  return 0;
}
fj_status_t dist_broadcast2_u64(DistState& d, const void* send_a, void* recv_a, const void* send_b, void* recv_b, size_t count,
                                int root, cudaStream_t st) {
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(d.comm);
  FJ_NCCL(g_api.GroupStart());
  FJ_NCCL(g_api.Broadcast(send_a, recv_a, count, ncclUint64, root, comm, st));
  FJ_NCCL(g_api.Broadcast(send_b, recv_b, count, ncclUint64, root, comm, st));
  FJ_NCCL(g_api.GroupEnd());
  return 0;
}
fj_status_t dist_allreduce_sum_u64(DistState& d, const void* send, void* recv, size_t count, cudaStream_t st) {
  FJ_NCCL(g_api.AllReduce(send, recv, count, ncclUint64, ncclSum, reinterpret_cast<ncclComm_t>(d.comm), st));
  return 0;
}
fj_status_t dist_allgather_u64(DistState& d, const void* send, void* recv, size_t count_per_rank, cudaStream_t st) {
  FJ_NCCL(g_api.AllGather(send, recv, count_per_rank, ncclUint64, reinterpret_cast<ncclComm_t>(d.comm), st));
  return 0;
}

fj_status_t dist_exchange(DistState& d, const DistMsg* sends, size_t n_sends, const DistMsg* recvs, size_t n_recvs,
                          cudaStream_t st) {
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(d.comm);
  FJ_NCCL(g_api.GroupStart());
  for (size_t i = 0; i < n_sends; ++i)
    if (sends[i].bytes) FJ_NCCL(g_api.Send(sends[i].ptr, sends[i].bytes, ncclUint8, sends[i].peer, comm, st));
  for (size_t i = 0; i < n_recvs; ++i)
    if (recvs[i].bytes) FJ_NCCL(g_api.Recv(recvs[i].ptr, recvs[i].bytes, ncclUint8, recvs[i].peer, comm, st));
  FJ_NCCL(g_api.GroupEnd());
  return 0;
}

}  // namespace fj
