// fj_dist.h — NCCL plumbing for the one-process-per-GPU multi-GPU joins (internal).
// NCCL is resolved with dlopen at fj_comm_init time so that the single-GPU library has no
// link-time dependency on it (and binds to whichever libnccl.so.2 the process already loaded).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace fj {

typedef int fj_status_t;
fj_status_t set_err(fj_status_t code, const char* fmt, ...);

struct NcclApi;  // resolved function table

struct DistState {
  bool ready = false;
  int rank = 0, world = 1;
  void* comm = nullptr;  // ncclComm_t
  const NcclApi* api = nullptr;
};

fj_status_t dist_unique_id(void* id128);
fj_status_t dist_init(DistState& d, int rank, int world, const void* id128);
void dist_destroy(DistState& d);
// collectives on 64-bit words
fj_status_t dist_broadcast_u64(DistState& d, void* buf, size_t count, int root, cudaStream_t st);
// out of place: root sends from `send`, everybody (root included) receives in `recv`
fj_status_t dist_broadcast_oop_u64(DistState& d, const void* send, void* recv, size_t count, int root, cudaStream_t st);
// two broadcasts in one NCCL group (one launch): root sends from send_a / send_b, everybody receives in recv_a / recv_b
fj_status_t dist_broadcast2_u64(DistState& d, const void* send_a, void* recv_a, const void* send_b, void* recv_b, size_t count,
                                int root, cudaStream_t st);
fj_status_t dist_allreduce_sum_u64(DistState& d, const void* send, void* recv, size_t count, cudaStream_t st);
fj_status_t dist_allgather_u64(DistState& d, const void* send, void* recv, size_t count_per_rank, cudaStream_t st);
// grouped point-to-point exchange: every message of `sends` is ncclSend to its peer, every message of `recvs`
// ncclRecv from its peer, all inside one ncclGroupStart/End.  Messages between the same pair of ranks are
// matched in list order.
struct DistMsg { int peer; void* ptr; uint64_t bytes; };
fj_status_t dist_exchange(DistState& d, const DistMsg* sends, size_t n_sends, const DistMsg* recvs, size_t n_recvs,
                          cudaStream_t st);

}  // namespace fj
