// fj_engine.cu — host side of libflashjoin_b200: the C ABI (include/flashjoin_b200.h), the device
// arena, path selection and the optimistic-attempt driver.
//
// Replaces the reference's L3/L4 host logic in /root/reference/hash_join.cpp:
//   join drivers _hash_join_{scalar,radix}_{count,materialize} (:315-567)  -> Engine::attempt_*
//   adaptive_hash_join_* + RADIX_JOIN_THRESHOLD (:576-594)                 -> Engine::choose_path
//   SimpleTimer (:45-55)                                                   -> CUDA events per phase
//   initialize_memory_system (:596)                                        -> fj_init
// The reference allocates every table/partition/result buffer inside each call (std::vector,
// make_unique through mimalloc); here buffers live in a grow-only device arena reused across calls.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/flashjoin_b200.h"
#include "fj_dist.h"
#include "fj_kernels.h"

namespace fj {

thread_local std::string g_err;

fj_status set_err(fj_status code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define FJ_CUDA(expr)                                                                                     \
  do {                                                                                                    \
    cudaError_t e__ = (expr);                                                                             \
    if (e__ != cudaSuccess) {                                                                             \
      cudaGetLastError();                                                                                 \
      return set_err(e__ == cudaErrorMemoryAllocation ? FJ_ERR_OOM : FJ_ERR_CUDA, "%s failed: %s (%s:%d)", \
                     #expr, cudaGetErrorString(e__), __FILE__, __LINE__);                                 \
    }                                                                                                     \
  } while (0)
#define FJ_TRY(expr)              \
  do {                            \
    fj_status s__ = (expr);       \
    if (s__ != FJ_OK) return s__; \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  fj_status ensure(size_t bytes) {
    if (bytes <= cap) return FJ_OK;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = (bytes + (size_t(2) << 20) - 1) & ~((size_t(2) << 20) - 1);
    FJ_CUDA(cudaMalloc(&p, want));
    cap = want;
    return FJ_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

static double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct RadixPlan {
  bool ok = false;
  bool narrow = false;
  int bits = 0, bits1 = 0, bits2 = 0;
  uint32_t P = 0, F1 = 0, F2 = 0;
  uint64_t cap1_b = 0, cap1_p = 0, cap2_b = 0, cap2_p = 0;
  uint32_t smax = 0, tcap = 0, chunk = 16384, max_chunks = 1;
  bool join3 = false;  // packed rows, two passes: collision-free pipelined k_join3, else k_join
};

static uint64_t round4(uint64_t x) { return (x + 3) & ~uint64_t(3); }
static uint64_t cap_build(uint64_t n, uint64_t parts) {
  const double m = (double)n / (double)parts;
  return round4((uint64_t)std::ceil(m + 6.0 * std::sqrt(m) + 32.0));
}
static uint64_t cap_probe(uint64_t n, uint64_t parts) {
  const double m = (double)n / (double)parts;
  return round4((uint64_t)std::ceil(m * 1.10 + 8.0 * std::sqrt(m) + 64.0));
}

struct Engine {
  std::mutex mu;
  bool inited = false;
  DeviceInfo di;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[16] = {};  // 0-3 attempt phases, 4-5 distributed call, 6-7 fj_timer_*, 8-11 shuffle phases, 12 build | probe partition pass, 13 end of the result exchange
  Ctl* h_ctl = nullptr;  // pinned
  // mapped pinned: PUB_WORDS words of tag << 32 | control-block word, written by k_publish_ctl; d_pub is the device-side address
  unsigned long long* h_pub = nullptr;
  void* d_pub = nullptr;
  unsigned long long pub_seq = 0;
  fj_status fetch_ctl(Ctl* d_ctl);
  // multi-GPU count: the ncclAllReduce of the control block is enqueued right behind the first attempt's kernels
  // (before the host has seen the flags), so a step has ONE host synchronisation instead of two.  The summed
  // flags word tells every rank whether any rank has to retry; only then a second all-reduce follows.
  bool spec_ar = false, spec_done = false;
  unsigned long long* h_spec = nullptr;  // pinned, sizeof(Ctl): element-wise sum of all ranks' control blocks
  fj_status spec_allreduce();
  // host -> device copy of one input column.  Page-locked sources (flash_join.pinned_empty, cudaHostRegister'ed
  // memory) go straight to the DMA engine; large PAGEABLE sources (plain numpy arrays — what a reference user
  // passes) are staged by a few host threads through a ring of pinned buffers, chunk by chunk, overlapped with
  // the DMA (cudaMemcpyAsync from pageable memory stages single-threaded: ~12 GB/s here vs 55 GB/s pinned).
  struct Stager {
    int threads = 0;
    size_t chunk = 0;
    std::vector<char*> bufs;           // [threads][2]
    std::vector<cudaEvent_t> evs;      // [threads][2] buffer free again
    std::vector<cudaStream_t> streams; // [threads]
    std::vector<cudaEvent_t> done;     // [threads]
  } stager;
  fj_status h2d(void* dst, const void* src, size_t bytes);
  // device -> host copy of a result column, the mirror image: large pageable destinations (fresh numpy arrays) are
  // filled by the same host threads from the pinned ring (each thread also takes the page faults of its chunks)
  fj_status d2h(void* dst, const void* src, size_t bytes);
  fj_status stager_setup(int want_threads, size_t chunk);
  void stager_release();
  DevBuf in_bk, in_bv, in_pk, table, bloom, ctl, out_keys, out_vals, out_idx;
  DevBuf part_a_b, part_a_p, part_b_b, part_b_p, cursors, flush, sj_tails, bcast_rows, sel_area;
  uint64_t pairs_n = 0;
  bool pairs_valid = false, pairs_idx = false;
  std::map<std::string, int64_t> cfg;
  DistState dist;
  // peer-memory exchange buffers (CUDA IPC over NVLink), set up at fj_comm_init; see k_count_dense_peer
  struct PeerState {
    bool ready = false;
    void* local = nullptr;
    size_t bytes = 0;
    std::vector<void*> mapped;            // [world], mapped[rank] == local
    unsigned long long** d_ptrs = nullptr;  // the same table in device memory
    unsigned long long step = 0;          // advances in lockstep on every rank (one per peer-path call)
  } peer;
  fj_status peer_setup();
  void peer_teardown();
  // every rank passes the status of its rank-local work; every rank gets FJ_OK only if all did (one 8-byte
  // ncclAllReduce): no rank may walk into a data collective its peers will never enter
  fj_status agree(fj_status local);
  // peer-memory shuffle (dense key domain): every rank's partition buffers + exchange area, mapped by every rank
  struct XPart {
    void* local = nullptr;
    size_t bytes = 0;
    std::vector<void*> mapped;           // [world], mapped[rank] == local
    unsigned long long step = 0;         // advances in lockstep on every rank
    bool meta_valid = false;             // meta[] = (nb, np) of every rank's slice, as of the last size exchange
    unsigned long long meta[16] = {};
    bool broken = false;                 // IPC mapping failed once: the NCCL shuffle answers from now on
  } xp;
  fj_status xpart_ensure(size_t bytes);
  void xpart_teardown();
  // returns FJ_OK with *handled == false when the data / configuration need the general NCCL shuffle
  fj_status join_shuffle_peer(unsigned jflags, const unsigned long long* d_bk, const unsigned long long* d_bv, uint64_t nb,
                              const unsigned long long* d_pk, uint64_t np, fj_stats* s, uint64_t* total, bool* handled);
  fj_status attempt_count_peer(uint64_t dbits, int root, const unsigned long long* bk_root, bool bk_on_device, uint64_t nb,
                               const unsigned long long* pk, uint64_t np, fj_stats* s);

  Engine() {
    cfg["load_pct"] = 50;
    cfg["load_pct_auto"] = 1;
    cfg["bloom_bits_per_key"] = 16;
    // adaptive: global table while its bytes stay below this percentage of L2, radix beyond.  Measured crossover of the
    // general hash paths on B200 (1e8 probe rows, profiles/r02t_sweep_adaptive.jsonl): 8e6 build rows 1.15 ms (table)
    // vs 2.21 ms (radix), 1.6e7 rows 2.30 vs 2.30, 3e7 rows 3.86 vs 2.48 — the table wins until it is about twice L2
    cfg["adaptive_table_l2_pct"] = 200;
    cfg["dense16_min_rows"] = 8192;       // build rows from which the dense16 radix path (k_part + k_sjoin) is planned
    cfg["mapped_result"] = 1;             // the attempt's control block returns through mapped pinned memory (k_publish_ctl) instead of a D2H copy
    cfg["dense16_sel_min_pct"] = 50;      // adaptive materialize, build side small enough for the dense table path: sampled match rate
                                          // (percent, up to 2^17 build rows; x sqrt(2^17 / rows) beyond) below which that path is
                                          // taken instead of dense16 (0 = never sample)
    cfg["dense16_sel_max_rows"] = 1 << 19;  // build rows beyond which the match rate is not sampled (the table path hardly ever wins)
    cfg["dense16_min_probe"] = 1 << 24;   // adaptive materialize: probe rows from which dense16 is preferred to the dense table path
    cfg["radix_sub_rows"] = 0;  // 0 = derive from shared memory
    cfg["radix_optimistic"] = 1;
    cfg["smem_bloom"] = 1;
    cfg["bloom_guard"] = 1;  // no filter when it would be read through L2 next to an L2-resident table
    cfg["probe_ctas_per_sm"] = 0;  // 0 = occupancy-derived
    cfg["narrow"] = 1;
    cfg["join3"] = 1;  // packed rows, two radix passes: collision-free pipelined k_join3 (0 = k_join)
    cfg["chunk_rows"] = 1 << 24;
    cfg["dense"] = 1;               // optimistic dense-key-domain fast paths (exact bitmap count, direct-address radix join)
    cfg["dense_min_rows"] = 1 << 20;  // radix: smallest build side that takes the direct-address join
    cfg["dense_group_mb"] = 8;      // radix: direct-address regions kept L2 resident per pipeline stage (delay_b + delay_p + 1 stages live)
    cfg["dense_ring"] = 4;          // k_djoin: items a CTA's dispatcher may publish ahead of its slowest worker warp
    cfg["dense_batch"] = 2;         // k_djoin: tickets per dispatcher round trip
    cfg["dense_delay_b"] = 1;       // k_djoin: steps between zeroing a group of regions and filling it
    cfg["dense_delay_p"] = 3;       // k_djoin: steps between filling a group and probing it (sweep: profiles/r01f_sweep_djoin.jsonl)
    cfg["stage_threads"] = 8;       // host threads staging large pageable inputs through pinned buffers (0: plain cudaMemcpyAsync)
    cfg["stage_min_mb"] = 64;       // smallest pageable input column that is staged
    // k_part input: per-warp TMA rings (0) or 128-bit loads straight into registers (1).  Measured on C3
    // (profiles/r02w_exp_part_direct.jsonl): rows with values 536 vs 564 us, probe keys 292 vs 294 us, the key-only build
    // pass of a count 323 vs 290 us — only that one takes the direct loads
    cfg["part_direct_kv"] = 0;
    cfg["part_direct_k"] = 0;
    cfg["part_direct_count_build"] = 1;
    cfg["peer_relay_min_rows"] = 1 << 18;  // multi-GPU count over peer memory: build rows from which the key slices / partial bitmaps relay is used
    cfg["dist_peer_reduce"] = 1;    // BROADCAST: 8-byte reductions (failure agreement, global count) over peer memory instead of ncclAllReduce
    cfg["dist_peer_bcast"] = 1;     // BROADCAST: build sides that fit the staging area travel over peer memory (k_peer_bcast), not ncclBroadcast
    cfg["dist_warmup"] = 1;         // fj_comm_init pays NCCL's first-use cost of broadcast and point-to-point channels
    cfg["dist_peer_shuffle"] = 1;   // SHUFFLE on a dense key domain: one partition pass storing straight into the owners' buffers
    cfg["dist_peer"] = 1;           // multi-GPU count over IPC-mapped peer memory (one kernel per GPU, no NCCL in the step)
    cfg["dist_spec_allreduce"] = 1; // multi-GPU count: all-reduce enqueued behind the first attempt (one host sync per step)
    cfg["dense_fused"] = 1;         // bitmap count as one persistent launch (grid barriers) instead of three kernels
    cfg["dense16"] = 1;             // radix path, dense key domain: one high-fan-out pass + shared-memory direct-address join (k_part / k_sjoin)
    cfg["dense16_logp"] = 0;        // 0 = derive the partition count from the build size; else log2(partitions), 8..11
    // FJ_CFG_<KEY>=<integer> in the environment overrides a default (e.g. FJ_CFG_DENSE=0)
    for (auto& kv : cfg) {
      std::string name = "FJ_CFG_";
      for (char ch : kv.first) name += (char)std::toupper((unsigned char)ch);
      if (const char* v = getenv(name.c_str())) kv.second = atoll(v);
    }
    cfg["shuffle_virtual_ranks"] = 1;  // > 1: every rank owns that many shuffle destinations (exercises the
                                       // multi-destination scatter / exchange layout on few GPUs)
  }

  fj_status init(int device);
  void shutdown();
  int choose_path(int algo, unsigned flags, uint64_t nb, bool narrow_guess, const RadixPlan& plan) const;
  RadixPlan plan_radix(uint64_t nb, uint64_t np, bool narrow) const;
  fj_status attempt_scalar(unsigned flags, bool narrow, bool exact, const unsigned long long* bk,
                           const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                           uint64_t idx_base, fj_stats* s);
  // dense key domain, radix path: one scatter pass by the low key bits + direct-address join in L2 (k_djoin)
  struct DensePlan { bool ok = false; uint64_t klimit = 0, rstride = 0, cap_b = 0, cap_p = 0; };
  DensePlan plan_dense(unsigned flags, uint64_t nb, uint64_t np) const;
  fj_status attempt_dense(unsigned flags, const DensePlan& dp, const unsigned long long* bk, const unsigned long long* bv,
                          uint64_t nb, const unsigned long long* pk, uint64_t np, fj_stats* s);
  DevBuf direct;
  // dense key domain, radix path, round 2: ONE partition pass (k_part) + shared-memory direct-address join (k_sjoin)
  struct Dense16Plan { bool ok = false; int logp = 0; uint64_t klimit = 0; uint32_t slots = 0; uint64_t cap_b = 0, cap_p = 0; };
  Dense16Plan plan_dense16(unsigned flags, uint64_t nb, uint64_t np, bool small_ok = false) const;
  // sel_min_pct != 0: k_sel_sample runs first and abandons the attempt (CTL_LOW_SEL) below that match rate
  fj_status attempt_dense16(unsigned flags, const Dense16Plan& dp, const unsigned long long* bk, const unsigned long long* bv,
                            uint64_t nb, const unsigned long long* pk, uint64_t np, fj_stats* s, uint32_t sel_min_pct = 0,
                            uint64_t sel_table_bits = 0);
  // dense key domain, count only: exact membership bitmap in shared memory instead of table + filter
  uint64_t dense_bitmap_bits(uint64_t nb) const;
  fj_status attempt_scalar_dense(unsigned flags, uint64_t dbits, const unsigned long long* bk, const unsigned long long* bv,
                                 uint64_t nb, const unsigned long long* pk, uint64_t np, uint64_t idx_base, fj_stats* s);
  // flat != nullptr: both sides are already in partition-element format (rows received from the multi-GPU
  // shuffle, holes included): flat->b / flat->p replace bk,bv / pk and every pass is a stage-2 pass
  struct FlatInput { const void* b; uint64_t nb; const void* p; uint64_t np; };
  fj_status attempt_radix(unsigned flags, const RadixPlan& pl, const unsigned long long* bk,
                          const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                          fj_stats* s, const FlatInput* flat = nullptr);
  fj_status join_shuffle(int algo, unsigned jflags, const unsigned long long* d_bk, const unsigned long long* d_bv,
                         uint64_t nb, const unsigned long long* d_pk, uint64_t np, fj_stats* s, uint64_t* total);
  fj_status finish_attempt(unsigned flags, fj_stats* s);
  DevBuf send_b, send_p, recv_b, recv_p, shuf_cur, shuf_meta, exp_bk, exp_bv, exp_pk, all_bk, all_bv;
  fj_status join_device(int algo, unsigned flags, const unsigned long long* bk, const unsigned long long* bv,
                        uint64_t nb, const unsigned long long* pk, uint64_t np, uint64_t idx_base, fj_stats* s);
  fj_status join(int algo, unsigned flags, const uint64_t* bk, const uint64_t* bv, size_t nb, const uint64_t* pk,
                 size_t np, uint64_t* out_matches, double* out_seconds, fj_stats* stats);
  fj_status join_dist(int mode, int algo, unsigned flags, int root, const uint64_t* bk, const uint64_t* bv, size_t nb,
                      const uint64_t* pk, size_t np, uint64_t* out_global, uint64_t* out_local, double* out_seconds,
                      fj_stats* stats);
  fj_status ensure_out(unsigned flags, uint64_t np);
  DevBuf dist_scratch;
  float ms(int a, int b) {
    float m = 0.f;
    cudaEventElapsedTime(&m, ev[a], ev[b]);
    return m;
  }
};

static Engine& E() {
  static Engine* e = new Engine();  // intentionally leaked: no destructor order issues at exit
  return *e;
}

fj_status Engine::init(int device) {
  if (inited) {
    if (device >= 0 && device != di.device)
      return set_err(FJ_ERR_STATE, "the engine is already initialised on device %d (fj_shutdown first to move to device %d)", di.device, device);
    return FJ_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return set_err(FJ_ERR_NO_DEVICE, "no CUDA device available (%s); flashjoin_b200 has no CPU fallback",
                   e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  if (device < 0) {  // LOCAL_RANK (one process per GPU), else the calling thread's current device
    const char* lr = getenv("LOCAL_RANK");
    if (lr) device = atoi(lr) % n;
    else if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; }
  }
  if (device >= n) return set_err(FJ_ERR_BAD_ARG, "device %d out of range (%d devices)", device, n);
  FJ_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FJ_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_err(FJ_ERR_NO_DEVICE, "device %d (%s) is sm_%d%d; this library is built for sm_100a only", device,
                   prop.name, prop.major, prop.minor);
  di.device = device;
  di.sms = prop.multiProcessorCount;
  di.l2_bytes = prop.l2CacheSize;
  di.smem_optin = prop.sharedMemPerBlockOptin;
  di.cc_major = prop.major;
  di.cc_minor = prop.minor;
  // tunables from the environment (experiments, sweeps): FJ_CFG_<key>=<integer> for any key of fj_config_set
  for (auto& kv : cfg) {
    const std::string name = "FJ_CFG_" + kv.first;
    if (const char* v = getenv(name.c_str())) kv.second = atoll(v);
  }
  FJ_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (auto& x : ev) FJ_CUDA(cudaEventCreate(&x));
  FJ_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_ctl), sizeof(Ctl)));
  FJ_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_spec), sizeof(Ctl)));
  FJ_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_pub), (PUB_WORDS + 2) * 8, cudaHostAllocMapped));
  FJ_CUDA(cudaHostGetDevicePointer(&d_pub, h_pub, 0));
  memset(h_pub, 0, (PUB_WORDS + 2) * 8);  // tag 0 is never used
  pub_seq = 0;
  FJ_TRY(ctl.ensure(4096));  // [0, 256): Ctl; [256, 272): grid-barrier words of the fused kernels (zero between launches)
  FJ_CUDA(cudaMemset(ctl.p, 0, 4096));
  inited = true;
  return FJ_OK;
}

void Engine::shutdown() {
  if (!inited) return;
  xpart_teardown();
  peer_teardown();
  dist_destroy(dist);
  cudaStreamSynchronize(st);
  for (DevBuf* b : {&in_bk, &in_bv, &in_pk, &table, &bloom, &ctl, &out_keys, &out_vals, &out_idx, &part_a_b,
                    &part_a_p, &part_b_b, &part_b_p, &cursors, &flush, &direct, &dist_scratch, &send_b, &send_p, &recv_b,
                    &recv_p, &shuf_cur, &shuf_meta, &exp_bk, &exp_bv, &exp_pk, &all_bk, &all_bv, &sj_tails, &bcast_rows, &sel_area})
    b->release();
  if (h_ctl) cudaFreeHost(h_ctl);
  h_ctl = nullptr;
  if (h_pub) cudaFreeHost(h_pub);
  h_pub = nullptr;
  d_pub = nullptr;
  if (h_spec) cudaFreeHost(h_spec);
  h_spec = nullptr;
  stager_release();
  for (auto& x : ev) { if (x) cudaEventDestroy(x); x = nullptr; }
  if (st) cudaStreamDestroy(st);
  st = nullptr;
  pairs_valid = false;
  inited = false;
}

// ---- planning ----------------------------------------------------------------------------------
// holes that k_scatter2's 16-byte run padding adds to one output partition fed by `ntiles` input tiles:
// each (tile, partition) run is padded by 0..pad_rows rows, (pad_rows / 2) on average
static uint64_t pad_allow(uint64_t ntiles, uint32_t pad_rows) {
  return pad_rows ? (uint64_t)(0.6 * pad_rows * (double)ntiles) + 16 : 0;
}

RadixPlan Engine::plan_radix(uint64_t nb, uint64_t np, bool narrow) const {
  RadixPlan pl;
  pl.narrow = narrow;
  // two join CTAs per SM: half of the SM's shared memory each, minus the per-CTA reservation
  const size_t budget = std::min<size_t>(di.smem_optin, (228 * 1024 - 2 * 1024) / 2 - 2048);
  const size_t tb = narrow ? 8 : 16;
  uint64_t smax_limit = (budget / (tb + 8)) & ~uint64_t(3);
  if (smax_limit > 65532) smax_limit = 65532;  // tuple index + 1 must fit 16 bits
  const int64_t user = cfg.at("radix_sub_rows");
  if (user > 0 && (uint64_t)user < smax_limit) smax_limit = std::max<uint64_t>(64, (uint64_t)user & ~uint64_t(3));
  const bool allow3 = narrow && cfg.at("join3") != 0;
  const uint32_t tile_b = scatter_tile_rows(true, narrow), tile_p = scatter_tile_rows(false, narrow);
  const uint32_t pad_b = scatter_pad_rows(true, narrow), pad_p = scatter_pad_rows(false, narrow);
  for (int B = 4; B <= 18; ++B) {
    const uint64_t P = 1ull << B;
    RadixPlan c;
    c.narrow = narrow;
    c.bits = B;
    if (B <= 8) { c.bits1 = B; c.bits2 = 0; }
    else { c.bits1 = (B + 1) / 2; c.bits2 = B - c.bits1; }
    if (c.bits1 > 8) break;  // one scatter pass fans out to at most 256 partitions
    // packed rows that need two passes anyway go to k_join3, which wants >= 2^14 partitions (its bitmap
    // covers the 32 - B hash bits the passes did not consume)
    const bool j3 = allow3 && c.bits2 > 0;
    if (j3 && 32 - B > 18) continue;
    c.P = (uint32_t)P;
    c.F1 = 1u << c.bits1;
    c.F2 = 1u << c.bits2;
    if (c.bits2) {
      c.cap1_b = round4(cap_build(nb, c.F1) + 32 + pad_allow(nb / tile_b + 1, pad_b));
      c.cap1_p = round4(cap_probe(np, c.F1) + pad_allow(np / tile_p + 1, pad_p));
      c.cap2_b = round4(cap_build(nb, P) + pad_allow(c.cap1_b / tile_b + 1, pad_b));
      c.cap2_p = round4(cap_probe(np, P) + pad_allow(c.cap1_p / tile_p + 1, pad_p));
    } else {
      c.cap2_b = round4(cap_build(nb, P) + pad_allow(nb / tile_b + 1, pad_b));
      c.cap2_p = round4(cap_probe(np, P) + pad_allow(np / tile_p + 1, pad_p));
    }
    c.smax = (uint32_t)std::min<uint64_t>(c.cap2_b, 0xffffffffu);
    c.join3 = j3;
    if (j3) {
      uint64_t lim = join3_max_build_rows();
      if (user > 0 && (uint64_t)user < lim) lim = std::max<uint64_t>(64, (uint64_t)user & ~uint64_t(3));
      if (c.cap2_b > lim || join3_smem_bytes(c.smax, 32 - B) + 1024 > budget) continue;
      c.tcap = 0;
      c.chunk = join3_probe_chunk();
    } else {
      if (c.cap2_b > smax_limit) continue;
      c.tcap = c.smax * 2;
      c.chunk = 16384;
    }
    c.max_chunks = (uint32_t)((c.cap2_p + c.chunk - 1) / c.chunk);
    if (c.max_chunks == 0) c.max_chunks = 1;
    if ((uint64_t)c.P * c.max_chunks > 0x7fffffffull) break;
    c.ok = true;
    return c;
  }
  return pl;  // too large for two passes: not applicable
}

int Engine::choose_path(int algo, unsigned flags, uint64_t nb, bool narrow_guess, const RadixPlan& plan) const {
  (void)flags;
  if (algo == FJ_ALGO_SCALAR) return FJ_ALGO_SCALAR;
  if (algo == FJ_ALGO_RADIX) return plan.ok ? FJ_ALGO_RADIX : FJ_ALGO_SCALAR;
  // adaptive (replaces RADIX_JOIN_THRESHOLD = 1'000'000 rows, hash_join.cpp:576): keep the global
  // table when it stays L2 resident, partition otherwise.
  const double load = (double)cfg.at("load_pct") / 100.0;
  const double table_bytes = (double)nb / load * (narrow_guess ? 8.0 : 16.0);
  const double l2_budget = (double)di.l2_bytes * (double)cfg.at("adaptive_table_l2_pct") / 100.0;
  if (table_bytes <= l2_budget || !plan.ok) return FJ_ALGO_SCALAR;
  return FJ_ALGO_RADIX;
}

// ---- host -> device input copy -------------------------------------------------------------------
void Engine::stager_release() {
  for (char* b : stager.bufs) if (b) cudaFreeHost(b);
  for (auto e : stager.evs) if (e) cudaEventDestroy(e);
  for (auto e : stager.done) if (e) cudaEventDestroy(e);
  for (auto q : stager.streams) if (q) cudaStreamDestroy(q);
  stager = Stager();
}

fj_status Engine::h2d(void* dst, const void* src, size_t bytes) {
  if (!bytes) return FJ_OK;
  const int want_threads = (int)std::min<int64_t>(32, std::max<int64_t>(0, cfg["stage_threads"]));
  bool staged = want_threads > 0 && bytes >= ((size_t)std::max<int64_t>(1, cfg["stage_min_mb"]) << 20);
  if (staged) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, src) != cudaSuccess) {
      cudaGetLastError();
    } else if (pa.type != cudaMemoryTypeUnregistered) {
      staged = false;  // pinned / registered / managed: the DMA engine reads it directly
    }
  }
  if (!staged) {
    FJ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return FJ_OK;
  }
  const size_t CH = size_t(8) << 20;
  FJ_TRY(stager_setup(want_threads, CH));
  // the copy streams must not overtake work already queued on the engine's stream that still uses `dst`
  FJ_CUDA(cudaEventRecord(stager.done[0], st));
  for (int t = 0; t < want_threads; ++t) FJ_CUDA(cudaStreamWaitEvent(stager.streams[(size_t)t], stager.done[0], 0));
  const size_t nchunks = (bytes + CH - 1) / CH;
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  const int dev = di.device;
  auto worker = [&](int t) {
    if (cudaSetDevice(dev) != cudaSuccess) { failed = 1; return; }
    int b = 0;
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= nchunks || failed.load()) break;
      const size_t off = i * CH, len = std::min(CH, bytes - off);
      char* buf = stager.bufs[(size_t)t * 2 + b];
      cudaEvent_t ev_free = stager.evs[(size_t)t * 2 + b];
      if (cudaEventSynchronize(ev_free) != cudaSuccess) { failed = 1; break; }  // the DMA out of this buffer is over
      memcpy(buf, static_cast<const char*>(src) + off, len);
      if (cudaMemcpyAsync(static_cast<char*>(dst) + off, buf, len, cudaMemcpyHostToDevice, stager.streams[(size_t)t]) != cudaSuccess ||
          cudaEventRecord(ev_free, stager.streams[(size_t)t]) != cudaSuccess) { failed = 1; break; }
      b ^= 1;
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < want_threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  if (failed.load()) {
    cudaGetLastError();
    return set_err(FJ_ERR_CUDA, "staged host->device copy failed");
  }
  // everything the copy streams did happens-before whatever is queued on the engine's stream next
  for (int t = 0; t < want_threads; ++t) {
    FJ_CUDA(cudaEventRecord(stager.done[(size_t)t], stager.streams[(size_t)t]));
    FJ_CUDA(cudaStreamWaitEvent(st, stager.done[(size_t)t], 0));
  }
  return FJ_OK;
}

fj_status Engine::stager_setup(int want_threads, size_t CH) {
  if (stager.threads == want_threads && stager.chunk == CH) return FJ_OK;
  stager_release();
  stager.threads = want_threads;
  stager.chunk = CH;
  stager.bufs.assign((size_t)want_threads * 2, nullptr);
  stager.evs.assign((size_t)want_threads * 2, nullptr);
  stager.streams.assign((size_t)want_threads, nullptr);
  stager.done.assign((size_t)want_threads, nullptr);
  for (auto& b : stager.bufs) FJ_CUDA(cudaMallocHost(reinterpret_cast<void**>(&b), CH));
  for (auto& e : stager.evs) FJ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& e : stager.done) FJ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto& q : stager.streams) FJ_CUDA(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
  return FJ_OK;
}

fj_status Engine::d2h(void* dst, const void* src, size_t bytes) {
  if (!bytes) return FJ_OK;
  const int want_threads = (int)std::min<int64_t>(32, std::max<int64_t>(0, cfg["stage_threads"]));
  bool staged = want_threads > 0 && bytes >= ((size_t)std::max<int64_t>(1, cfg["stage_min_mb"]) << 20);
  if (staged) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, dst) != cudaSuccess) {
      cudaGetLastError();
    } else if (pa.type != cudaMemoryTypeUnregistered) {
      staged = false;  // pinned / registered / managed: the DMA engine writes it directly
    }
  }
  if (!staged) {
    FJ_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
    return FJ_OK;
  }
  const size_t CH = size_t(8) << 20;
  FJ_TRY(stager_setup(want_threads, CH));
  FJ_CUDA(cudaStreamSynchronize(st));  // the pairs are final
  const size_t nchunks = (bytes + CH - 1) / CH;
  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  const int dev = di.device;
  // every thread keeps two chunks in flight: the DMA of chunk i + 1 runs while chunk i is copied out of its pinned buffer
  auto worker = [&](int t) {
    if (cudaSetDevice(dev) != cudaSuccess) { failed = 1; return; }
    cudaStream_t q = stager.streams[(size_t)t];
    size_t idx[2] = {nchunks, nchunks};
    auto start = [&](int b) {
      const size_t i = next.fetch_add(1);
      idx[b] = i;
      if (i >= nchunks) return;
      const size_t off = i * CH, len = std::min(CH, bytes - off);
      if (cudaMemcpyAsync(stager.bufs[(size_t)t * 2 + b], static_cast<const char*>(src) + off, len, cudaMemcpyDeviceToHost, q) != cudaSuccess ||
          cudaEventRecord(stager.evs[(size_t)t * 2 + b], q) != cudaSuccess) failed = 1;
    };
    start(0);
    start(1);
    for (int b = 0; idx[b] < nchunks && !failed.load(); b ^= 1) {
      if (cudaEventSynchronize(stager.evs[(size_t)t * 2 + b]) != cudaSuccess) { failed = 1; break; }
      const size_t off = idx[b] * CH, len = std::min(CH, bytes - off);
      memcpy(static_cast<char*>(dst) + off, stager.bufs[(size_t)t * 2 + b], len);
      start(b);
    }
    cudaStreamSynchronize(q);
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < want_threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  if (failed.load()) {
    cudaGetLastError();
    return set_err(FJ_ERR_CUDA, "staged device->host copy failed");
  }
  return FJ_OK;
}

fj_status Engine::spec_allreduce() {
  if (!spec_ar || spec_done) return FJ_OK;
  spec_done = true;
  static_assert(sizeof(Ctl) % 8 == 0, "Ctl is summed as 64-bit words");
  unsigned long long* d_sum = reinterpret_cast<unsigned long long*>(static_cast<char*>(ctl.p) + 512);
  FJ_TRY(dist_allreduce_sum_u64(dist, ctl.p, d_sum, sizeof(Ctl) / 8, st));
  FJ_CUDA(cudaMemcpyAsync(h_spec, d_sum, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
  return FJ_OK;
}

fj_status Engine::ensure_out(unsigned flags, uint64_t np) {
  if (!(flags & FJ_FLAG_MATERIALIZE)) return FJ_OK;
  const size_t bytes = std::max<uint64_t>(np, 1) * 8;
  FJ_TRY(out_keys.ensure(bytes));
  FJ_TRY(out_vals.ensure(bytes));
  if (flags & FJ_FLAG_PROBE_IDX) FJ_TRY(out_idx.ensure(bytes));
  return FJ_OK;
}

// ---- one attempt on the global-table path ------------------------------------------------------
fj_status Engine::attempt_scalar(unsigned flags, bool narrow, bool exact, const unsigned long long* bk,
                                 const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                                 uint64_t idx_base, fj_stats* s) {
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  const uint64_t spb = narrow ? 4 : 2;
  // load factor: 50 % by default; a table that stays far below L2 size even at 25 % takes 25 %
  // (full home buckets, which force a second dependent sector read, drop from 14 % to 2 %)
  uint64_t load = (uint64_t)std::max<int64_t>(10, std::min<int64_t>(90, cfg["load_pct"]));
  if (cfg["load_pct_auto"] && nb * (narrow ? 8ull : 16ull) * 4ull <= (uint64_t)di.l2_bytes / 4) load = std::min<uint64_t>(load, 25);
  uint64_t nbuckets = (nb * 100 + load * spb - 1) / (load * spb);
  if (nbuckets < 1) nbuckets = 1;
  if (nbuckets > 0xfffffff0ull) return set_err(FJ_ERR_BAD_ARG, "build side too large for one table (%llu rows)", (unsigned long long)nb);
  TableView t;
  t.narrow = narrow;
  t.nbuckets = (uint32_t)nbuckets;
  const size_t table_bytes = (size_t)nbuckets * 32;
  FJ_TRY(table.ensure(table_bytes));
  t.slots = table.as<unsigned long long>();
  bool bloom_smem = false;
  size_t bloom_bytes = 0;
  // Bloom guard: a filter that does not fit shared memory is read through L2 — one more random access per probe row.
  // That pays only when it saves an HBM access, i.e. when the table itself does not stay in L2; next to an
  // L2-resident table it can only lose (1.25e8 x 1e6, 90 % match: 0.91 ms with the filter, 0.52 ms without,
  // profiles/r01i_quick_bench.jsonl).  The *_bloom entry points then run without a filter (bloom_kind 0) — results are
  // identical, as on the reference, where the filter only ever changes the time.
  const bool filter_fits_smem = cfg["smem_bloom"] && round4((nb * 8 + 31) / 32) <= probe_smem_bloom_limit_words(di);
  const bool want_filter = (flags & FJ_FLAG_BLOOM) &&
                           (filter_fits_smem || !cfg["bloom_guard"] || table_bytes > (size_t)di.l2_bytes / 2);
  if (want_filter) {
    const uint64_t bits = (uint64_t)std::max<int64_t>(4, cfg["bloom_bits_per_key"]);
    uint64_t words = round4(std::max<uint64_t>(8, (nb * bits + 31) / 32));
    const uint64_t lim = probe_smem_bloom_limit_words(di);
    if (filter_fits_smem) {  // >= 8 bits/key still fit shared memory
      words = std::min(words, lim);
      bloom_smem = true;
    }
    if (words > 0xfffffff0ull) words = 0xfffffff0ull;
    bloom_bytes = words * 4;
    FJ_TRY(bloom.ensure(bloom_bytes));
    t.bloom = bloom.as<uint32_t>();
    t.bloom_words = (uint32_t)words;
  }
  int launches = 0;
  Ctl* d_ctl = ctl.as<Ctl>();
  FJ_CUDA(cudaEventRecord(ev[0], st));
  launch_prepare(d_ctl, t.slots, table_bytes, t.bloom, t.bloom ? bloom_bytes : 0, di, st);  // ctl + empty table + filter
  ++launches;
  FJ_CUDA(cudaEventRecord(ev[1], st));
  launch_build(t, bk, bv, nb, exact ? 1 : 0, d_ctl, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[2], st));
  ProbeOut po;
  if (mat) {
    po.keys = out_keys.as<unsigned long long>();
    po.vals = out_vals.as<unsigned long long>();
    po.idx = (flags & FJ_FLAG_PROBE_IDX) ? out_idx.as<unsigned long long>() : nullptr;
    po.idx_base = idx_base;
  }
  launch_probe(t, pk, np, bv, mat ? &po : nullptr, bloom_smem, (int)cfg["probe_ctas_per_sm"], d_ctl, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[3], st));
  FJ_TRY(spec_allreduce());  // multi-GPU count only (no-op otherwise)
  FJ_TRY(fetch_ctl(d_ctl));
  s->clear_s += ms(0, 1) * 1e-3;
  s->build_s += ms(1, 2) * 1e-3;
  s->probe_s += ms(2, 3) * 1e-3;
  s->device_s += ms(0, 3) * 1e-3;
  s->kernel_launches += launches;
  s->table_bytes = table_bytes + bloom_bytes;
  s->path = FJ_ALGO_SCALAR;
  s->narrow = narrow ? 1 : 0;
  s->bloom_kind = t.bloom ? (bloom_smem ? 1 : 2) : 0;
  s->dedup_exact = exact ? 1 : 0;
  s->radix_bits1 = s->radix_bits2 = 0;
  return FJ_OK;
}

// ---- dense key domain, count only (global-table path) -----------------------------------------
// Optimistic key bound: the next power of two above 2*nb (h2o keys are 1..1.1*nb, 1..1.9*nb at 10 % match),
// clipped to what fits shared memory; 0 = not applicable.
uint64_t Engine::dense_bitmap_bits(uint64_t nb) const {
  if (!cfg.at("dense") || !cfg.at("narrow") || nb == 0) return 0;
  uint64_t d = 256;
  while (d < 2 * nb) d <<= 1;
  const uint64_t lim = (uint64_t)probe_smem_bitmap_limit_bytes(di) * 8;
  if (d > lim) d = lim & ~uint64_t(127);
  return d >= nb + 1 ? d : 0;
}

fj_status Engine::attempt_scalar_dense(unsigned flags, uint64_t dbits, const unsigned long long* bk,
                                       const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                                       uint64_t idx_base, fj_stats* s) {
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  const size_t bytes = dbits / 8;
  FJ_TRY(bloom.ensure(bytes));
  if (mat) FJ_TRY(table.ensure((size_t)dbits * 8));  // direct-address values: 8 bytes per key of the domain
  Ctl* d_ctl = ctl.as<Ctl>();
  int launches = 0;
  FJ_CUDA(cudaEventRecord(ev[0], st));
  // one persistent launch (prepare | build | probe separated by grid barriers), else the three-kernel sequence
  uint32_t* gsync = reinterpret_cast<uint32_t*>(static_cast<char*>(ctl.p) + 256);  // zeroed at init, kept zero by the kernel
  bool fused = false;
  if (mat) {
    ProbeOut po;
    po.keys = out_keys.as<unsigned long long>();
    po.vals = out_vals.as<unsigned long long>();
    po.idx = (flags & FJ_FLAG_PROBE_IDX) ? out_idx.as<unsigned long long>() : nullptr;
    po.idx_base = idx_base;
    fused = launch_mat_dense_fused(bk, bv, nb, pk, np, bloom.as<uint32_t>(), (uint32_t)(dbits / 32), table.as<unsigned long long>(),
                                   d_ctl, gsync, po, di, st, &launches);
    if (!fused) {  // no co-resident launch configuration: this attempt did not run, the hash path answers
      h_ctl->flags = CTL_NOT_DENSE;
      return FJ_OK;
    }
  } else {
    fused = cfg["dense_fused"] != 0 &&
            launch_count_dense_fused(bk, nb, pk, np, bloom.as<uint32_t>(), (uint32_t)(dbits / 32), d_ctl, gsync, di, st, &launches);
    if (!fused) {
      launch_prepare(d_ctl, nullptr, 0, bloom.p, bytes, di, st);
      ++launches;
      FJ_CUDA(cudaEventRecord(ev[1], st));
      launch_build_bitmap(bloom.as<uint32_t>(), dbits, bk, nb, d_ctl, di, st, &launches);
      FJ_CUDA(cudaEventRecord(ev[2], st));
      launch_probe_count_dense(pk, np, bloom.as<uint32_t>(), (uint32_t)(dbits / 32), d_ctl, di, st, &launches);
    }
  }
  FJ_CUDA(cudaEventRecord(ev[3], st));
  FJ_TRY(spec_allreduce());  // multi-GPU count only (no-op otherwise)
  FJ_TRY(fetch_ctl(d_ctl));
  if (fused) {
    s->probe_s += ms(0, 3) * 1e-3;  // one kernel: clear, build and probe are phases of it
  } else {
    s->clear_s += ms(0, 1) * 1e-3;
    s->build_s += ms(1, 2) * 1e-3;
    s->probe_s += ms(2, 3) * 1e-3;
  }
  s->device_s += ms(0, 3) * 1e-3;
  s->kernel_launches += launches;
  s->table_bytes = bytes + (mat ? (uint64_t)dbits * 8 : 0);
  s->path = FJ_ALGO_SCALAR;
  s->narrow = 1;
  s->bloom_kind = 3;
  s->dense = 1;
  s->dedup_exact = 0;
  s->radix_bits1 = s->radix_bits2 = 0;
  return FJ_OK;
}

// ---- one attempt on the radix path -------------------------------------------------------------
fj_status Engine::attempt_radix(unsigned flags, const RadixPlan& pl, const unsigned long long* bk,
                                const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                                fj_stats* s, const FlatInput* flat) {
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  const bool two = pl.bits2 > 0;
  const size_t tb = radix_elem_bytes(true, pl.narrow), tp = radix_elem_bytes(false, pl.narrow);
  if (two) {
    FJ_TRY(part_a_b.ensure((size_t)pl.F1 * pl.cap1_b * tb));
    FJ_TRY(part_a_p.ensure((size_t)pl.F1 * pl.cap1_p * tp));
  }
  FJ_TRY(part_b_b.ensure((size_t)pl.P * pl.cap2_b * tb));
  FJ_TRY(part_b_p.ensure((size_t)pl.P * pl.cap2_p * tp));
  // cursor layout: [A build F1][A probe F1][B build P][B probe P][flat build count][flat probe count]
  const size_t ncur = 2 * (size_t)pl.F1 + 2 * (size_t)pl.P;
  FJ_TRY(cursors.ensure((ncur + 2) * 4));
  uint32_t* cur_a_b = cursors.as<uint32_t>();
  uint32_t* cur_a_p = cur_a_b + pl.F1;
  uint32_t* cur_b_b = cur_a_p + pl.F1;
  uint32_t* cur_b_p = cur_b_b + pl.P;
  Ctl* d_ctl = ctl.as<Ctl>();
  int launches = 0;

  FJ_CUDA(cudaEventRecord(ev[0], st));
  launch_prepare(d_ctl, nullptr, 0, cursors.p, ncur * 4, di, st);  // ctl + partition cursors
  ++launches;
  FJ_CUDA(cudaEventRecord(ev[1], st));

  ScatterArgs a;
  a.ctl = d_ctl;
  int stage = 1;
  if (flat) {  // rows that arrived from the shuffle: one flat "partition" per side
    if (flat->nb > 0xffffffffull || flat->np > 0xffffffffull)
      return set_err(FJ_ERR_BAD_ARG, "more than 2^32 - 1 rows on one rank after the shuffle");
    uint32_t* flat_cnt = cur_b_p + pl.P;
    const uint32_t hc[2] = {(uint32_t)flat->nb, (uint32_t)flat->np};
    FJ_CUDA(cudaMemcpyAsync(flat_cnt, hc, sizeof(hc), cudaMemcpyHostToDevice, st));
    stage = 2;
    a.in_nparts = 1;
  }
  auto first_pass = [&](bool build, void* out, uint32_t* cur, uint64_t cap, int shift, uint32_t fan) {
    if (flat) {
      a.in_part = build ? flat->b : flat->p;
      a.in_counts = cur_b_p + pl.P + (build ? 0 : 1);
      a.in_cap = a.n_upper = build ? flat->nb : flat->np;
    } else {
      a.in_keys = build ? bk : pk;
      a.in_vals = build ? bv : nullptr;
      a.n = build ? nb : np;
    }
    a.out = out; a.out_cursor = cur; a.out_cap = cap; a.shift = shift; a.fan = fan;
    if (!flat || a.in_cap) launch_scatter(build, pl.narrow, stage, a, di, st, &launches);
  };
  if (!two) {
    first_pass(true, part_b_b.p, cur_b_b, pl.cap2_b, 32 - pl.bits, pl.P);
    first_pass(false, part_b_p.p, cur_b_p, pl.cap2_p, 32 - pl.bits, pl.P);
  } else {
    first_pass(true, part_a_b.p, cur_a_b, pl.cap1_b, 32 - pl.bits1, pl.F1);
    first_pass(false, part_a_p.p, cur_a_p, pl.cap1_p, 32 - pl.bits1, pl.F1);
    ScatterArgs b;
    b.ctl = d_ctl;
    b.shift = 32 - pl.bits; b.fan = pl.F2; b.in_nparts = pl.F1;
    b.in_part = part_a_b.p; b.in_counts = cur_a_b; b.in_cap = pl.cap1_b; b.n_upper = flat ? flat->nb : nb;
    b.out = part_b_b.p; b.out_cursor = cur_b_b; b.out_cap = pl.cap2_b;
    launch_scatter(true, pl.narrow, 2, b, di, st, &launches);
    b.in_part = part_a_p.p; b.in_counts = cur_a_p; b.in_cap = pl.cap1_p; b.n_upper = flat ? flat->np : np;
    b.out = part_b_p.p; b.out_cursor = cur_b_p; b.out_cap = pl.cap2_p;
    launch_scatter(false, pl.narrow, 2, b, di, st, &launches);
  }
  FJ_CUDA(cudaEventRecord(ev[2], st));

  JoinArgs j;
  j.build = part_b_b.p; j.bcnt = cur_b_b; j.cap_b = pl.cap2_b;
  j.probe = part_b_p.p; j.pcnt = cur_b_p; j.cap_p = pl.cap2_p;
  j.smax = pl.smax; j.tcap = pl.tcap; j.chunk = pl.chunk; j.max_chunks = pl.max_chunks; j.nparts = pl.P;
  j.ctl = d_ctl;
  j.out_keys = mat ? out_keys.as<unsigned long long>() : nullptr;
  j.out_vals = mat ? out_vals.as<unsigned long long>() : nullptr;
  // radix + Bloom (hash_join_radix_bloom / hash_join_count_radix_bloom, hash_join.cpp:627, :636): k_join carries a
  // per-partition register-blocked filter in shared memory (16 bits per key) when it fits next to the table; k_join3's
  // collision-free membership bitmap already IS an exact filter in front of its value directory
  int bloom_kind = 0;
  if ((flags & FJ_FLAG_BLOOM) && !pl.join3) {
    const uint32_t bw = std::max<uint32_t>(32u, (pl.smax / 2 + 3u) & ~3u);
    if ((size_t)pl.smax * (pl.narrow ? 8 : 16) + (size_t)pl.tcap * 4 + (size_t)bw * 4 + 2048 <= di.smem_optin) {
      j.bloom_words = bw;
      bloom_kind = 4;
    }
  } else if ((flags & FJ_FLAG_BLOOM) && pl.join3) {
    bloom_kind = 3;
  }
  if (pl.join3) launch_join3(mat, j, 32 - pl.bits, di, st, &launches);
  else launch_join(pl.narrow, mat, j, st, &launches);
  if (!pl.narrow && !flat) launch_emit_sentinel(d_ctl, bv, j.out_keys, j.out_vals, mat, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[3], st));
  FJ_TRY(spec_allreduce());  // multi-GPU count only (no-op otherwise)
  FJ_TRY(fetch_ctl(d_ctl));
  s->clear_s += ms(0, 1) * 1e-3;
  s->partition_s += ms(1, 2) * 1e-3;
  s->probe_s += ms(2, 3) * 1e-3;
  s->device_s += ms(0, 3) * 1e-3;
  s->kernel_launches += launches;
  s->table_bytes = (uint64_t)pl.P * (pl.cap2_b * tb + pl.cap2_p * tp) + (two ? (uint64_t)pl.F1 * (pl.cap1_b * tb + pl.cap1_p * tp) : 0);
  s->path = FJ_ALGO_RADIX;
  s->narrow = pl.narrow ? 1 : 0;
  s->bloom_kind = bloom_kind;
  s->dedup_exact = 0;
  s->radix_bits1 = pl.bits1;
  s->radix_bits2 = pl.bits2;
  return FJ_OK;
}

// ---- one attempt on the dense-key-domain radix path ---------------------------------------------
// Optimistic: every build key < klimit = 2^(ceil(log2 nb) + 1) (h2o ids are 1..1.1*nb) and every build value
// < 2^32 - 1.  ONE scatter pass by the low 8 key bits (k_scatter2 with DomainArgs::ident), then k_djoin.
static uint64_t round8(uint64_t x) { return (x + 7) & ~uint64_t(7); }
Engine::DensePlan Engine::plan_dense(unsigned flags, uint64_t nb, uint64_t np) const {
  DensePlan dp;
  if (!cfg.at("dense") || !cfg.at("narrow") || (flags & (FJ_FLAG_FORCE_WIDE | FJ_FLAG_PROBE_IDX))) return dp;
  if (nb < (uint64_t)std::max<int64_t>(cfg.at("dense_min_rows"), 1024) || nb > (1ull << 31)) return dp;
  uint64_t k = 1024;
  while (k < 2 * nb) k <<= 1;
  if (k > 0xFFFFFFFFull) k = 0xFFFFFFFFull;
  const uint32_t F = djoin_fan();
  const uint32_t tile_b = scatter_tile_rows(true, true), tile_p = scatter_tile_rows(false, true);
  const uint32_t pad_b = scatter_pad_rows(true, true), pad_p = scatter_pad_rows(false, true);
  dp.klimit = k;
  dp.rstride = round4((k + F - 1) / F);
  dp.cap_b = round4(cap_build(nb, F) + 32 + pad_allow(nb / tile_b + 1, pad_b));
  dp.cap_p = round8(cap_probe(np, F) + pad_allow(np / tile_p + 1, pad_p));
  if (dp.cap_b > 0xfffffff0ull || dp.cap_p > 0xfffffff0ull) return dp;
  dp.ok = true;
  return dp;
}

fj_status Engine::attempt_dense(unsigned flags, const DensePlan& dp, const unsigned long long* bk,
                                const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                                fj_stats* s) {
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  const uint32_t F = djoin_fan();
  FJ_TRY(part_a_b.ensure((size_t)F * dp.cap_b * 8));
  FJ_TRY(part_a_p.ensure((size_t)F * dp.cap_p * 4));
  // cursor layout: [build F][probe F][k_djoin sync words]
  const size_t ncur = 2 * (size_t)F + djoin_sync_words();
  FJ_TRY(cursors.ensure(ncur * 4));
  uint32_t* cur_b = cursors.as<uint32_t>();
  uint32_t* cur_p = cur_b + F;
  Ctl* d_ctl = ctl.as<Ctl>();
  int launches = 0;
  FJ_CUDA(cudaEventRecord(ev[0], st));
  launch_prepare(d_ctl, nullptr, 0, cursors.p, ncur * 4, di, st);
  ++launches;
  FJ_CUDA(cudaEventRecord(ev[1], st));
  ScatterArgs a;
  a.ctl = d_ctl; a.shift = 0; a.fan = F;
  a.dom.klimit = dp.klimit; a.dom.vlimit = 0xFFFFFFFEull; a.dom.badflag = CTL_NOT_DENSE; a.dom.ident = 1;
  a.in_keys = bk; a.in_vals = bv; a.n = nb; a.out = part_a_b.p; a.out_cursor = cur_b; a.out_cap = dp.cap_b;
  launch_scatter(true, true, 1, a, di, st, &launches);
  a.in_keys = pk; a.in_vals = nullptr; a.n = np; a.out = part_a_p.p; a.out_cursor = cur_p; a.out_cap = dp.cap_p;
  launch_scatter(false, true, 1, a, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[2], st));
  DjoinArgs j;
  j.build = part_a_b.p; j.bcnt = cur_b; j.cap_b = dp.cap_b;
  j.probe = part_a_p.p; j.pcnt = cur_p; j.cap_p = dp.cap_p;
  j.direct = direct.as<uint32_t>(); j.rstride = dp.rstride;
  j.group_bytes = (uint64_t)std::max<int64_t>(1, cfg["dense_group_mb"]) << 20;
  j.ring = (uint32_t)std::max<int64_t>(1, cfg["dense_ring"]);
  j.batch = (uint32_t)std::max<int64_t>(1, cfg["dense_batch"]);
  j.delay_b = (uint32_t)std::max<int64_t>(1, cfg["dense_delay_b"]);
  j.delay_p = (uint32_t)std::max<int64_t>(1, cfg["dense_delay_p"]);
  j.ctl = d_ctl; j.sync = cur_p + F;
  j.out_keys = mat ? out_keys.as<unsigned long long>() : nullptr;
  j.out_vals = mat ? out_vals.as<unsigned long long>() : nullptr;
  const bool launched = launch_djoin(mat, j, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[3], st));
  if (launched) FJ_TRY(spec_allreduce());  // multi-GPU count only (no-op otherwise)
  FJ_TRY(fetch_ctl(d_ctl));
  if (!launched) h_ctl->flags |= CTL_NOT_DENSE;  // the host falls back to the general radix path
  // fewer non-empty slots than rows stored: some key was stored twice (duplicate build keys)
  if (mat && !(h_ctl->flags & (CTL_NOT_DENSE | CTL_OVERFLOW)) && h_ctl->dense_slots != h_ctl->dense_rows) h_ctl->flags |= CTL_DUP;
  s->clear_s += ms(0, 1) * 1e-3;
  s->partition_s += ms(1, 2) * 1e-3;
  s->probe_s += ms(2, 3) * 1e-3;
  s->device_s += ms(0, 3) * 1e-3;
  s->kernel_launches += launches;
  s->table_bytes = (uint64_t)F * (dp.cap_b * 8 + dp.cap_p * 4 + dp.rstride * 4);
  s->path = FJ_ALGO_RADIX;
  s->narrow = 1;
  s->bloom_kind = 0;
  s->dedup_exact = 0;
  s->radix_bits1 = 8;
  s->radix_bits2 = 0;
  s->dense = 1;
  return FJ_OK;
}

// ---- one attempt on the dense-key-domain radix path, round 2 (k_part + k_sjoin) -------------------
// Optimistic: every build key < klimit (<= 65528 direct-address slots per partition) and, for a materialize, every
// build value <= 65534.  Partition count: enough that a partition's slice of the key domain fits the shared-memory
// region of k_sjoin, and enough partitions to balance the SMs.
static uint64_t round16(uint64_t x) { return (x + 15) & ~uint64_t(15); }
Engine::Dense16Plan Engine::plan_dense16(unsigned flags, uint64_t nb, uint64_t np, bool small_ok) const {
  Dense16Plan dp;
  if (!cfg.at("dense") || !cfg.at("dense16") || !cfg.at("narrow") || (flags & (FJ_FLAG_FORCE_WIDE | FJ_FLAG_PROBE_IDX))) return dp;
  // small build sides only when the caller asked for them (adaptive materialize with a big probe side); an explicit radix
  // request keeps the general radix path below dense_min_rows
  const int64_t min_rows = small_ok ? std::min(cfg.at("dense_min_rows"), cfg.at("dense16_min_rows")) : cfg.at("dense_min_rows");
  if (nb < (uint64_t)std::max<int64_t>(min_rows, 1024) || np == 0) return dp;
  const uint64_t maxslots = sjoin_max_slots(di);
  const uint64_t need = nb + nb / 5 + 1024;  // h2o ids are 1..1.1*n
  int logp = (int)cfg.at("dense16_logp");
  if (logp <= 0) {
    // 1024 partitions whenever the key domain fits them (1e8 probe rows against 1e4 .. 2e6 build rows: 0.63 - 0.65 ms
    // with 1024 partitions, 0.68 - 0.70 with 512, 0.64 - 0.67 with 2048: profiles/r02u_exp_small_dense16.jsonl)
    logp = 10;
    while (logp < 11 && (maxslots << logp) < need) ++logp;
  }
  if (logp < 8 || logp > 11 || (maxslots << logp) < need) return dp;
  if (part_smem_bytes(logp) + 256 > di.smem_optin) return dp;
  const uint64_t P = 1ull << logp;
  uint64_t k = 1024;
  while (k < 4 * nb) k <<= 1;
  dp.klimit = std::min<uint64_t>(maxslots << logp, k);
  dp.slots = (uint32_t)std::min<uint64_t>(maxslots, ((dp.klimit + P - 1) / P + 7) & ~uint64_t(7));
  dp.logp = logp;
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  // every CTA of the pass leaves one (padded) sector per partition behind
  dp.cap_b = round16(cap_build(nb, P) + ((uint64_t)part_grid(mat, nb, di) + 2) * part_sector_elems(mat));
  // a partition holds nb / P distinct keys: with few of them the probe rows per partition follow the number of keys that
  // fell into it, not the number of rows (1e4 build keys, 1024 partitions: 10 +- 2 keys, +-21 % rows) — room for 4.5 sigma
  const double keys_per_part = std::max(1.0, (double)nb / (double)P);
  const uint64_t few_keys = (uint64_t)((double)np / (double)P * std::min(2.0, 4.5 / std::sqrt(keys_per_part)));
  dp.cap_p = round16(cap_probe(np, P) + few_keys + ((uint64_t)part_grid(false, np, di) + 2) * part_sector_elems(false));
  if (dp.cap_b > 0xfffffff0ull || dp.cap_p > 0xfffffff0ull) return dp;
  dp.ok = true;
  return dp;
}

fj_status Engine::attempt_dense16(unsigned flags, const Dense16Plan& dp, const unsigned long long* bk,
                                  const unsigned long long* bv, uint64_t nb, const unsigned long long* pk, uint64_t np,
                                  fj_stats* s, uint32_t sel_min_pct, uint64_t sel_table_bits) {
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  const uint32_t P = 1u << dp.logp;
  const size_t eb = mat ? 4 : 2;
  FJ_TRY(part_a_b.ensure((size_t)P * dp.cap_b * eb));
  FJ_TRY(part_a_p.ensure((size_t)P * dp.cap_p * 2));
  const uint32_t cs = part_cursor_stride();
  FJ_TRY(cursors.ensure(2 * (size_t)P * cs * 4));
  uint32_t* cur_b = cursors.as<uint32_t>();
  uint32_t* cur_p = cur_b + (size_t)P * cs;
  if (mat) {  // k_sjoin reserves output in blocks: room for every CTA's unused block tails
    const size_t ob = (size_t)(np + sjoin_out_slack_pairs(di)) * 8;
    FJ_TRY(out_keys.ensure(ob));
    FJ_TRY(out_vals.ensure(ob));
    FJ_TRY(sj_tails.ensure(sjoin_tail_bytes(di)));
  }
  Ctl* d_ctl = ctl.as<Ctl>();
  int launches = 0;
  FJ_CUDA(cudaEventRecord(ev[0], st));
  if (sel_min_pct) FJ_TRY(sel_area.ensure(sel_sample_bytes(dp.klimit)));
  launch_prepare(d_ctl, sel_min_pct ? sel_area.p : nullptr, sel_min_pct ? sel_sample_bytes(dp.klimit) : 0, cursors.p, 2 * (size_t)P * cs * 4, di, st,
                 cs, P, part_cursor_start(mat, nb, di), part_cursor_start(false, np, di));
  ++launches;
  if (sel_min_pct) launch_sel_sample(d_ctl, bk, nb, pk, np, sel_area.p, dp.klimit, sel_table_bits, sel_min_pct, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[1], st));
  PartArgs a;
  a.ctl = d_ctl; a.klimit = dp.klimit; a.logp = dp.logp;
  a.cursor_stride = cs;
  a.in_keys = bk; a.in_vals = mat ? bv : nullptr; a.n = nb; a.cap = dp.cap_b; a.cursor = cur_b; a.out = part_a_b.p; a.strict = true;
  a.direct_in = (mat ? cfg["part_direct_kv"] : cfg["part_direct_count_build"]) != 0;
  bool launched = launch_part(mat, a, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[12], st));
  a.in_keys = pk; a.in_vals = nullptr; a.n = np; a.cap = dp.cap_p; a.cursor = cur_p; a.out = part_a_p.p; a.strict = false;
  a.direct_in = cfg["part_direct_k"] != 0;
  launched = launched && launch_part(false, a, di, st, &launches);
  FJ_CUDA(cudaEventRecord(ev[2], st));
  if (launched) {
    SjoinArgs j;
    j.build[0] = part_a_b.p; j.bcnt = cur_b; j.cap_b = dp.cap_b;
    j.probe[0] = part_a_p.p; j.pcnt = cur_p; j.cap_p = dp.cap_p;
    j.cnt_stride = 0; j.cursor_stride = cs; j.p_first = 0; j.p_count = P; j.logp = dp.logp; j.nsub = 1; j.slots_alloc = dp.slots;
    j.ctl = d_ctl;
    j.out_keys = mat ? out_keys.as<unsigned long long>() : nullptr;
    j.out_vals = mat ? out_vals.as<unsigned long long>() : nullptr;
    j.tails = mat ? sj_tails.as<unsigned long long>() : nullptr;
    launched = launch_sjoin(mat, j, di, st, &launches);
  }
  FJ_CUDA(cudaEventRecord(ev[3], st));
  if (launched) FJ_TRY(spec_allreduce());  // multi-GPU count only (no-op otherwise)
  FJ_TRY(fetch_ctl(d_ctl));
  if (!launched) h_ctl->flags |= CTL_NOT_DENSE16;  // no launch configuration: the next layout answers
  if (mat && launched && !(h_ctl->flags & (CTL_NOT_DENSE16 | CTL_OVERFLOW))) {
    // out_cursor counts whole reserved blocks; k_pairs_compact has made [0, match_count) dense
    if (h_ctl->out_cursor < h_ctl->match_count || h_ctl->out_cursor - h_ctl->match_count > sjoin_out_slack_pairs(di))
      return set_err(FJ_ERR_STATE, "internal: %llu pairs reserved for %llu matches", (unsigned long long)h_ctl->out_cursor,
                     (unsigned long long)h_ctl->match_count);
    h_ctl->out_cursor = h_ctl->match_count;
  }
  s->clear_s += ms(0, 1) * 1e-3;
  s->partition_s += ms(1, 2) * 1e-3;
  s->probe_s += ms(2, 3) * 1e-3;
  s->device_s += ms(0, 3) * 1e-3;
  s->kernel_launches += launches;
  s->table_bytes = (uint64_t)P * (dp.cap_b * eb + dp.cap_p * 2);
  s->path = FJ_ALGO_RADIX;
  s->narrow = 1;
  s->bloom_kind = 0;
  s->dedup_exact = 0;
  s->radix_bits1 = dp.logp;
  s->radix_bits2 = 0;
  s->dense = 2;
  s->part_build_us = (int32_t)(ms(1, 12) * 1e3f);
  s->part_probe_us = (int32_t)(ms(12, 2) * 1e3f);
  return FJ_OK;
}

// The control block of the attempt just queued, into *h_ctl; returns once the stream has drained up to here.
// (the publishing kernel is not counted in fj_stats.kernel_launches, which keeps meaning "kernels of the join")
fj_status Engine::fetch_ctl(Ctl* d_ctl) {
  if (!cfg["mapped_result"]) {
    FJ_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
    FJ_CUDA(cudaGetLastError());
    return FJ_OK;
  }
  if ((uint32_t)++pub_seq == 0) ++pub_seq;
  const uint32_t tag = (uint32_t)pub_seq;
  launch_publish_ctl(d_ctl, d_pub, tag, st);
  volatile unsigned long long* words = h_pub;
  uint32_t data[PUB_WORDS];
  // a faulted kernel never publishes: look at the stream now and then so that the error surfaces instead of a hang
  for (unsigned spins = 1;; ++spins) {
    bool all = true;
    for (int i = 0; i < PUB_WORDS; ++i) {
      const unsigned long long w = words[i];
      data[i] = (uint32_t)w;
      all &= (uint32_t)(w >> 32) == tag;
    }
    if (all) break;
    if ((spins & 4095u) == 0) {
      const cudaError_t q = cudaStreamQuery(st);
      if (q == cudaErrorNotReady) continue;
      if (q != cudaSuccess) return set_err(FJ_ERR_CUDA, "%s", cudaGetErrorString(q));
      bool late = true;  // the stream has drained: the words are there unless the launch itself was lost
      for (int i = 0; i < PUB_WORDS; ++i) late &= (uint32_t)(words[i] >> 32) == tag;
      if (!late) return set_err(FJ_ERR_STATE, "internal: the stream drained without publishing the control block");
    }
  }
  static_assert(sizeof(Ctl) == sizeof(data), "control block words");
  memcpy(h_ctl, data, sizeof(Ctl));
  FJ_CUDA(cudaGetLastError());
  return FJ_OK;
}

fj_status Engine::finish_attempt(unsigned flags, fj_stats* s) {
  s->matches = h_ctl->match_count;
  if (flags & FJ_FLAG_MATERIALIZE) {
    if (h_ctl->out_cursor != h_ctl->match_count)
      return set_err(FJ_ERR_STATE, "internal: pair cursor %llu != match count %llu", (unsigned long long)h_ctl->out_cursor,
                     (unsigned long long)h_ctl->match_count);
    pairs_valid = true;
    pairs_n = h_ctl->match_count;
    pairs_idx = (flags & FJ_FLAG_PROBE_IDX) != 0;
  }
  return FJ_OK;
}

// ---- the optimistic-attempt driver (inputs resident in HBM) ------------------------------------
fj_status Engine::join_device(int algo, unsigned flags, const unsigned long long* bk, const unsigned long long* bv,
                              uint64_t nb, const unsigned long long* pk, uint64_t np, uint64_t idx_base, fj_stats* s) {
  pairs_valid = false;
  s->matches = 0;
  if (nb == 0 || np == 0) {  // hash_join.cpp: empty build or probe -> (0, t) on every path
    s->path = algo == FJ_ALGO_RADIX ? FJ_ALGO_RADIX : FJ_ALGO_SCALAR;
    s->attempts = 1;
    if (flags & FJ_FLAG_MATERIALIZE) { pairs_valid = true; pairs_n = 0; pairs_idx = (flags & FJ_FLAG_PROBE_IDX) != 0; }
    return FJ_OK;
  }
  FJ_TRY(ensure_out(flags, np));
  const bool mat = flags & FJ_FLAG_MATERIALIZE;
  bool narrow = !(flags & FJ_FLAG_FORCE_WIDE) && cfg["narrow"] != 0;
  bool exact = false;
  RadixPlan plan = plan_radix(nb, np, narrow);
  int path = choose_path(algo, flags, nb, narrow, plan);
  if ((flags & FJ_FLAG_PROBE_IDX) && mat) path = FJ_ALGO_SCALAR;  // radix drops row ids (:254-292)
  // optimistic dense-key-domain fast paths (both leave through CTL_NOT_DENSE when the data says otherwise)
  uint64_t dense_bits = (narrow && path == FJ_ALGO_SCALAR && (!mat || cfg["dense_fused"] != 0)) ? dense_bitmap_bits(nb) : 0;
  // materialize: only while two CTAs (two bitmaps) fit one SM — with a single 1024-thread CTA per SM the
  // direct-address kernel was slower than the hash-table kernel at 90 % match rate (1.30 vs 1.03 ms at 1.25e8 x 1e6,
  // profiles/r01h_quick_bench.jsonl), so larger domains keep the general path
  if (mat && dense_bits / 8 * 2 + 8192 > (uint64_t)di.smem_optin) {
    const uint64_t lim2 = (((uint64_t)di.smem_optin - 8192) / 2 * 8) & ~uint64_t(127);
    dense_bits = lim2 >= nb + 1 ? lim2 : 0;
  }
  DensePlan dplan;
  // the direct-address radix join has no partition-size limit: it also serves build sides that two general scatter
  // passes cannot cut down to shared-memory size (plan.ok == false, > ~2.6e8 rows), which would otherwise fall back
  // to one huge global table
  // adaptive on a dense key domain (measured with 1e8 probe rows, profiles/r02t_sweep_adaptive.jsonl and
  // r02u_exp_small_dense16.jsonl): a materialize is fastest on the dense16 radix path at EVERY build size (0.63 - 0.70 ms
  // for 1e4 .. 2e6 build rows, the dense table path 0.72 - 1.18 ms), a count only once the membership bitmap no longer
  // fits shared memory (2e6 rows: 0.39 vs 0.50 ms; below that the bitmap kernel answers in 0.14 - 0.16 ms).  The attempt
  // is optimistic: keys outside the domain cost one pass over the build side and the next layout answers.
  bool prefer16 = false;
  if (algo == FJ_ALGO_ADAPTIVE && narrow && cfg["dense"] && cfg["dense16"] && !(flags & (FJ_FLAG_FORCE_WIDE | FJ_FLAG_PROBE_IDX))) {
    if (mat) prefer16 = np >= (uint64_t)cfg["dense16_min_probe"];
    else prefer16 = dense_bits == 0 && nb >= (1ull << 20);
  }
  bool radix_wanted =
      path == FJ_ALGO_RADIX || prefer16 ||
      (!plan.ok && !((flags & FJ_FLAG_PROBE_IDX) && mat) &&
       (algo == FJ_ALGO_RADIX || (algo == FJ_ALGO_ADAPTIVE && (double)nb / ((double)cfg["load_pct"] / 100.0) * 8.0 >
                                                                 (double)di.l2_bytes * (double)cfg["adaptive_table_l2_pct"] / 100.0)));
  // with a build side small enough for the dense table path the match rate decides (k_sel_sample); explicit radix requests
  // and build sides beyond the shared-memory bitmaps never sample.  Measured with 1e8 probe rows
  // (profiles/r02O_exp_selectivity.jsonl): up to 1e5 build rows the table path wins below ~50 % matches (10 %: 0.35 vs 0.51 ms,
  // 40 %: 0.46 vs 0.52), at 4e5 rows only below ~27 % (its L2-resident value table grows with the build side) — the
  // threshold falls with the square root of the build size; the sample costs the dense16 path 13 us of 0.63 ms.
  uint32_t sel_min_pct = 0;
  if (prefer16 && mat && path == FJ_ALGO_SCALAR && dense_bits != 0 && nb <= (uint64_t)cfg["dense16_sel_max_rows"]) {
    double t = (double)std::max<int64_t>(0, std::min<int64_t>(100, cfg["dense16_sel_min_pct"]));
    if (nb > 131072) t *= std::sqrt(131072.0 / (double)nb);
    sel_min_pct = (uint32_t)t;
  }
  Dense16Plan d16;
  if (narrow && radix_wanted) {
    d16 = plan_dense16(flags, nb, np, prefer16);
    dplan = plan_dense(flags, nb, np);
  }
  for (int attempt = 1; attempt <= 7; ++attempt) {
    s->attempts = attempt;
    s->dense = 0;
    if (path == FJ_ALGO_RADIX) {
      if (plan.narrow != narrow) plan = plan_radix(nb, np, narrow);
      if (!plan.ok) { path = FJ_ALGO_SCALAR; }
    }
    const bool dense16 = radix_wanted && narrow && d16.ok;
    // the L2-resident direct-address regions of the round-1 path are only allocated when that path is really taken
    if (!dense16 && radix_wanted && narrow && dplan.ok && direct.ensure((size_t)djoin_fan() * dplan.rstride * 4) != FJ_OK) {
      cudaGetLastError();
      dplan.ok = false;  // no room: general path
    }
    const bool dense_radix = !dense16 && radix_wanted && narrow && dplan.ok;
    const bool dense_scalar = path == FJ_ALGO_SCALAR && narrow && !exact && dense_bits != 0;
    if (dense16) FJ_TRY(attempt_dense16(flags, d16, bk, bv, nb, pk, np, s, sel_min_pct, dense_bits));
    else if (dense_radix) FJ_TRY(attempt_dense(flags, dplan, bk, bv, nb, pk, np, s));
    else if (dense_scalar) FJ_TRY(attempt_scalar_dense(flags, dense_bits, bk, bv, nb, pk, np, idx_base, s));
    else if (path == FJ_ALGO_RADIX) FJ_TRY(attempt_radix(flags, plan, bk, bv, nb, pk, np, s));
    else FJ_TRY(attempt_scalar(flags, narrow, exact, bk, bv, nb, pk, np, idx_base, s));
    const unsigned f = h_ctl->flags;
    if (dense16 && (f & (CTL_NOT_DENSE16 | CTL_OVERFLOW))) {  // wider layouts answer
      d16.ok = false;
      if (f & CTL_LOW_SEL) radix_wanted = path == FJ_ALGO_RADIX;  // few probe rows match: the dense table path (dense_bits != 0)
      continue;
    }
    if (f & CTL_NOT_DENSE) { dplan.ok = false; dense_bits = 0; continue; }
    if ((f & CTL_OVERFLOW) && dense_radix) { dplan.ok = false; continue; }  // skewed low key bits: hash partitioning instead
    if ((f & CTL_NEED_WIDE) && narrow) { narrow = false; continue; }
    if ((f & CTL_OVERFLOW) && path == FJ_ALGO_RADIX) { path = FJ_ALGO_SCALAR; continue; }
    if ((f & CTL_DUP) && !exact) { exact = true; narrow = false; path = FJ_ALGO_SCALAR; continue; }
    return finish_attempt(flags, s);
  }
  return set_err(FJ_ERR_STATE, "internal: join did not converge after 7 attempts (flags %u)", h_ctl->flags);
}

fj_status Engine::join(int algo, unsigned flags, const uint64_t* bk, const uint64_t* bv, size_t nb, const uint64_t* pk,
                       size_t np, uint64_t* out_matches, double* out_seconds, fj_stats* stats) {
  if (algo < FJ_ALGO_ADAPTIVE || algo > FJ_ALGO_RADIX) return set_err(FJ_ERR_BAD_ARG, "unknown algo %d", algo);
  if (flags & ~(FJ_FLAG_BLOOM | FJ_FLAG_MATERIALIZE | FJ_FLAG_DEVICE_INPUTS | FJ_FLAG_FORCE_WIDE | FJ_FLAG_PROBE_IDX))
    return set_err(FJ_ERR_BAD_ARG, "unknown flag bits 0x%x", flags);
  if ((nb && (!bk || !bv)) || (np && !pk)) return set_err(FJ_ERR_BAD_ARG, "NULL input pointer with non-zero length");
  if (!out_matches) return set_err(FJ_ERR_BAD_ARG, "out_matches is NULL");
  FJ_TRY(init(-1));
  FJ_CUDA(cudaSetDevice(di.device));
  const double t0 = now_s();
  fj_stats s;
  memset(&s, 0, sizeof(s));
  s.n_gpus = 1;
  const unsigned long long *d_bk, *d_bv, *d_pk;
  if (flags & FJ_FLAG_DEVICE_INPUTS) {
    d_bk = reinterpret_cast<const unsigned long long*>(bk);
    d_bv = reinterpret_cast<const unsigned long long*>(bv);
    d_pk = reinterpret_cast<const unsigned long long*>(pk);
  } else {
    FJ_TRY(in_bk.ensure(std::max<size_t>(nb, 1) * 8));
    FJ_TRY(in_bv.ensure(std::max<size_t>(nb, 1) * 8));
    FJ_TRY(in_pk.ensure(std::max<size_t>(np, 1) * 8));
    const double th = now_s();
    if (nb) {
      FJ_TRY(h2d(in_bk.p, bk, nb * 8));
      FJ_TRY(h2d(in_bv.p, bv, nb * 8));
    }
    if (np) FJ_TRY(h2d(in_pk.p, pk, np * 8));
    FJ_CUDA(cudaStreamSynchronize(st));
    s.h2d_s = now_s() - th;
    s.h2d_bytes = (uint64_t)(2 * nb + np) * 8;
    d_bk = in_bk.as<unsigned long long>();
    d_bv = in_bv.as<unsigned long long>();
    d_pk = in_pk.as<unsigned long long>();
  }
  FJ_TRY(join_device(algo, flags, d_bk, d_bv, nb, d_pk, np, 0, &s));
  s.wall_s = now_s() - t0;
  s.algorithmic_bytes = (flags & FJ_FLAG_MATERIALIZE) ? 16ull * nb + 8ull * np + 16ull * s.matches : 8ull * (nb + np);
  *out_matches = s.matches;
  // host inputs: the copy into HBM is part of the join the caller timed (the reference's seconds cover the whole join
  // from the arrays it was given, hash_join.cpp:319-379); fj_stats keeps device_s and h2d_s apart
  if (out_seconds) *out_seconds = s.device_s + s.h2d_s;
  if (stats) *stats = s;
  return FJ_OK;
}

// ---- SHUFFLE (large build side, BASELINE.json configs[2] at G > 1 and configs[4]) ---------------
// Every rank holds a slice of BOTH sides.  Rows are hash-partitioned by destination rank with the same
// pipelined scatter kernel the local radix passes use (destination digit = low 16 hash bits range-reduced,
// independent of the top bits the local passes consume), already narrowed to the packed partition format
// when the data allows (12 instead of 24 bytes per build+probe row pair cross NVLink), exchanged with one
// grouped ncclSend/ncclRecv all-to-all-v, and joined locally on the receiving rank straight from the
// received partition-format rows.  Equal keys meet on exactly one rank, so local results simply add up.
fj_status Engine::join_shuffle(int algo, unsigned jflags, const unsigned long long* d_bk, const unsigned long long* d_bv,
                               uint64_t nb, const unsigned long long* d_pk, uint64_t np, fj_stats* s, uint64_t* total) {
  const int W = dist.world, R = dist.rank;
  int V = (int)std::max<int64_t>(1, cfg["shuffle_virtual_ranks"]);
  if (W * V > 256) V = std::max(1, 256 / W);
  if (W > 256) return set_err(FJ_ERR_BAD_ARG, "shuffle supports at most 256 ranks");
  const uint32_t F = (uint32_t)(W * V);
  const bool mat = jflags & FJ_FLAG_MATERIALIZE;
  bool narrow = !(jflags & FJ_FLAG_FORCE_WIDE) && cfg["narrow"] != 0;
  Ctl* d_ctl = ctl.as<Ctl>();
  pairs_valid = false;
  const size_t M = 8 + 2 * (size_t)F;  // 64-bit words of per-rank metadata
  std::vector<uint64_t> meta(M), all(M * (size_t)W);
  FJ_TRY(shuf_meta.ensure((M + M * W) * 8));
  uint64_t* d_meta = shuf_meta.as<uint64_t>();
  double comm_ms = 0.0, part_ms = 0.0;

  bool conservative = false;  // send regions sized for ALL local rows (after an overflow of the expected-size regions)
  for (int attempt = 1; attempt <= 4; ++attempt) {
    // ---- 1. scatter the local rows by destination.  A rank-local failure here (allocation, launch) must not leave the
    // peers alone in the collective of step 2: the status travels with the metadata and every rank returns an error.
    const size_t tb = radix_elem_bytes(true, narrow), tp = radix_elem_bytes(false, narrow);
    // destination regions like the local radix pass: expected size + 6 sigma + the per-tile padding; CTL_OVERFLOW on any
    // rank re-runs with regions that hold every local row (heavy skew towards one destination)
    const uint64_t pad_b = (uint64_t)scatter_pad_rows(true, narrow) * (nb / scatter_tile_rows(true, narrow) + 1) + 16;
    const uint64_t pad_p = (uint64_t)scatter_pad_rows(false, narrow) * (np / scatter_tile_rows(false, narrow) + 1) + 16;
    const uint64_t cap_sb = round4((conservative || F == 1 ? nb : std::min<uint64_t>(nb, cap_build(nb, F))) + pad_b);
    const uint64_t cap_sp = round4((conservative || F == 1 ? np : std::min<uint64_t>(np, cap_probe(np, F))) + pad_p);
    std::vector<uint32_t> h_cur(2 * (size_t)F, 0u);
    unsigned long long srow = EMPTY64, sval = 0, my_sent_probes = 0;
    const fj_status st1 = [&]() -> fj_status {
      FJ_TRY(send_b.ensure((size_t)F * cap_sb * tb));
      FJ_TRY(send_p.ensure((size_t)F * cap_sp * tp));
      FJ_TRY(shuf_cur.ensure(2 * (size_t)F * 4));
      uint32_t* cur_sb = shuf_cur.as<uint32_t>();
      uint32_t* cur_sp = cur_sb + F;
      int launches = 0;
      FJ_CUDA(cudaEventRecord(ev[8], st));
      launch_init_ctl(d_ctl, st);
      ++launches;
      FJ_CUDA(cudaMemsetAsync(shuf_cur.p, 0, 2 * (size_t)F * 4, st));
      ScatterArgs a;
      a.ctl = d_ctl; a.shift = -1; a.fan = F;
      if (nb) {
        a.in_keys = d_bk; a.in_vals = d_bv; a.n = nb; a.out = send_b.p; a.out_cursor = cur_sb; a.out_cap = cap_sb;
        launch_scatter(true, narrow, 1, a, di, st, &launches);
      }
      if (np) {
        a.in_keys = d_pk; a.in_vals = nullptr; a.n = np; a.out = send_p.p; a.out_cursor = cur_sp; a.out_cap = cap_sp;
        launch_scatter(false, narrow, 1, a, di, st, &launches);
      }
      FJ_CUDA(cudaEventRecord(ev[9], st));
      FJ_CUDA(cudaMemcpyAsync(h_cur.data(), shuf_cur.p, 2 * (size_t)F * 4, cudaMemcpyDeviceToHost, st));
      FJ_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
      FJ_CUDA(cudaStreamSynchronize(st));
      FJ_CUDA(cudaGetLastError());
      part_ms += ms(8, 9);
      s->kernel_launches += launches;
      // the out-of-band key (wide rows only): first build row that carries it, and its value
      srow = h_ctl->sentinel_row;
      if (!narrow && srow != EMPTY64) FJ_CUDA(cudaMemcpy(&sval, d_bv + srow, 8, cudaMemcpyDeviceToHost));
      my_sent_probes = narrow ? 0ull : h_ctl->sentinel_probes;
      return FJ_OK;
    }();
    const std::string err1 = st1 == FJ_OK ? std::string() : g_err;
    if (st1 != FJ_OK) {
      cudaGetLastError();
      std::fill(h_cur.begin(), h_cur.end(), 0u);
      h_ctl->flags = 0;
    }

    // ---- 2. everybody learns everybody's flags and send counts
    meta[0] = h_ctl->flags;
    meta[1] = (narrow || srow == EMPTY64) ? EMPTY64 : (((uint64_t)R << 40) | srow);  // global keep-first order: (rank, row)
    meta[2] = sval;
    meta[3] = my_sent_probes;
    meta[4] = nb;
    meta[5] = np;
    meta[6] = st1 == FJ_OK ? 0 : 1;  // rank-local failure before the exchange
    meta[7] = 0;
    for (uint32_t d = 0; d < F; ++d) { meta[8 + d] = h_cur[d]; meta[8 + F + d] = h_cur[F + d]; }
    FJ_CUDA(cudaEventRecord(ev[10], st));
    FJ_CUDA(cudaMemcpyAsync(d_meta, meta.data(), M * 8, cudaMemcpyHostToDevice, st));
    FJ_TRY(dist_allgather_u64(dist, d_meta, d_meta + M, M, st));
    FJ_CUDA(cudaMemcpyAsync(all.data(), d_meta + M, M * W * 8, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
    unsigned any = 0;
    bool any_failed = false;
    for (int r = 0; r < W; ++r) {
      any |= (unsigned)all[(size_t)r * M];
      any_failed |= all[(size_t)r * M + 6] != 0;
    }
    if (any_failed) {  // every rank leaves here, with the same verdict
      if (st1 != FJ_OK) { g_err = err1; return st1; }
      return set_err(FJ_ERR_STATE, "shuffle: a peer rank failed before the exchange (see its fj_last_error)");
    }
    if ((any & CTL_OVERFLOW) && !conservative) {  // a destination received far more than its share: regions for every local row
      FJ_CUDA(cudaEventRecord(ev[11], st));
      FJ_CUDA(cudaEventSynchronize(ev[11]));
      comm_ms += ms(10, 11);
      conservative = true;
      continue;
    }
    if ((any & CTL_NEED_WIDE) && narrow) {  // some rank holds a row that does not fit 32|32: everybody re-runs wide
      FJ_CUDA(cudaEventRecord(ev[11], st));
      FJ_CUDA(cudaEventSynchronize(ev[11]));
      comm_ms += ms(10, 11);
      narrow = false;
      continue;
    }
    if (any & CTL_OVERFLOW) return set_err(FJ_ERR_STATE, "internal: shuffle send region overflowed");
    s->attempts = attempt;

    // ---- 3. all-to-all-v of the partition-format rows
    uint64_t nb_recv = 0, np_recv = 0;
    std::vector<DistMsg> sends, recvs;
    for (uint32_t d = 0; d < F; ++d) {
      const int peer = (int)(d / V);
      sends.push_back({peer, send_b.as<char>() + (size_t)d * cap_sb * tb, (uint64_t)h_cur[d] * tb});
    }
    for (uint32_t d = 0; d < F; ++d) {
      const int peer = (int)(d / V);
      sends.push_back({peer, send_p.as<char>() + (size_t)d * cap_sp * tp, (uint64_t)h_cur[F + d] * tp});
    }
    for (int r = 0; r < W; ++r)
      for (int v = 0; v < V; ++v) nb_recv += all[(size_t)r * M + 8 + (size_t)R * V + v];
    for (int r = 0; r < W; ++r)
      for (int v = 0; v < V; ++v) np_recv += all[(size_t)r * M + 8 + F + (size_t)R * V + v];
    FJ_TRY(recv_b.ensure(nb_recv * tb + 64));
    FJ_TRY(recv_p.ensure(np_recv * tp + 64));
    {
      uint64_t ob = 0, op = 0;
      for (int r = 0; r < W; ++r)
        for (int v = 0; v < V; ++v) {
          const uint64_t c = all[(size_t)r * M + 8 + (size_t)R * V + v];
          recvs.push_back({r, recv_b.as<char>() + ob * tb, c * tb});
          ob += c;
        }
      for (int r = 0; r < W; ++r)
        for (int v = 0; v < V; ++v) {
          const uint64_t c = all[(size_t)r * M + 8 + F + (size_t)R * V + v];
          recvs.push_back({r, recv_p.as<char>() + op * tp, c * tp});
          op += c;
        }
    }
    FJ_TRY(dist_exchange(dist, sends.data(), sends.size(), recvs.data(), recvs.size(), st));
    FJ_CUDA(cudaEventRecord(ev[11], st));
    FJ_CUDA(cudaEventSynchronize(ev[11]));
    comm_ms += ms(10, 11);
    s->h2d_bytes += 0;

    // the out-of-band key, resolved globally: the value of the first build row (rank-major) that carries it
    bool sent_exists = false;
    unsigned long long sent_val = 0;
    {
      uint64_t best = EMPTY64;
      for (int r = 0; r < W; ++r) {
        const uint64_t key = all[(size_t)r * M + 1];
        if (key < best) { best = key; sent_val = all[(size_t)r * M + 2]; sent_exists = true; }
      }
    }

    // ---- 4. local join of the received rows
    FJ_TRY(ensure_out(jflags, np_recv + my_sent_probes + 1));
    bool ran = false;
    if (nb_recv && np_recv) {
      RadixPlan plan = plan_radix(nb_recv, np_recv, narrow);
      int path = (algo == FJ_ALGO_SCALAR || !plan.ok) ? FJ_ALGO_SCALAR : FJ_ALGO_RADIX;
      for (int local = 0; local < 2 && !ran; ++local) {
        if (path == FJ_ALGO_RADIX) {
          const FlatInput fl{recv_b.p, nb_recv, recv_p.p, np_recv};
          FJ_TRY(attempt_radix(jflags, plan, nullptr, nullptr, 0, nullptr, 0, s, &fl));
          if (h_ctl->flags & CTL_OVERFLOW) { path = FJ_ALGO_SCALAR; continue; }  // skewed partition: global table instead
          ran = true;
        } else {
          // partition-format rows -> plain columns (holes dropped), then the global-table path
          FJ_TRY(exp_bk.ensure(nb_recv * 8));
          FJ_TRY(exp_bv.ensure(nb_recv * 8));
          FJ_TRY(exp_pk.ensure(np_recv * 8));
          unsigned long long* d_cnt2 = reinterpret_cast<unsigned long long*>(d_meta);
          int l2 = 0;
          FJ_CUDA(cudaMemsetAsync(d_cnt2, 0, 16, st));
          launch_expand(true, narrow, recv_b.p, nb_recv, exp_bk.as<unsigned long long>(), exp_bv.as<unsigned long long>(), d_cnt2,
                        di, st, &l2);
          launch_expand(false, narrow, recv_p.p, np_recv, exp_pk.as<unsigned long long>(), nullptr, d_cnt2 + 1, di, st, &l2);
          unsigned long long h_cnt2[2] = {0, 0};
          FJ_CUDA(cudaMemcpyAsync(h_cnt2, d_cnt2, 16, cudaMemcpyDeviceToHost, st));
          FJ_CUDA(cudaStreamSynchronize(st));
          s->kernel_launches += l2;
          if (h_cnt2[0] && h_cnt2[1]) {
            FJ_TRY(attempt_scalar(jflags & ~FJ_FLAG_PROBE_IDX, narrow, false, exp_bk.as<unsigned long long>(),
                                  exp_bv.as<unsigned long long>(), h_cnt2[0], exp_pk.as<unsigned long long>(), h_cnt2[1], 0, s));
            ran = true;
          }
          break;
        }
      }
    }
    if (!ran) {  // nothing to join on this rank: still a defined control block (the collectives below need every rank)
      int l3 = 0;
      launch_init_ctl(d_ctl, st);
      ++l3;
      s->kernel_launches += l3;
      s->path = FJ_ALGO_RADIX;
      s->narrow = narrow ? 1 : 0;
    }
    if (sent_exists && my_sent_probes) {
      int l4 = 0;
      launch_emit_sentinel_value(d_ctl, sent_val, my_sent_probes, out_keys.as<unsigned long long>(), out_vals.as<unsigned long long>(),
                                 mat, st, &l4);
      s->kernel_launches += l4;
    }
    FJ_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
    FJ_CUDA(cudaGetLastError());

    // ---- 5. duplicates anywhere?  keep-first is defined on the GLOBAL build order (rank-major), which the
    // shuffle does not preserve: gather the whole build side on every rank and run the exact local join of
    // the rank's own probe rows against it.  Slow, rare (every BASELINE.json config has unique build keys).
    uint64_t flag2[2] = {(uint64_t)((h_ctl->flags & CTL_DUP) ? 1 : 0), 0};
    std::vector<uint64_t> all2(2 * (size_t)W);
    FJ_CUDA(cudaEventRecord(ev[10], st));
    FJ_CUDA(cudaMemcpyAsync(d_meta, flag2, 16, cudaMemcpyHostToDevice, st));
    FJ_TRY(dist_allgather_u64(dist, d_meta, d_meta + 2, 2, st));
    FJ_CUDA(cudaMemcpyAsync(all2.data(), d_meta + 2, 16 * (size_t)W, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaEventRecord(ev[11], st));
    FJ_CUDA(cudaStreamSynchronize(st));
    comm_ms += ms(10, 11);
    bool any_dup = false;
    for (int r = 0; r < W; ++r) any_dup |= all2[2 * (size_t)r] != 0;
    if (any_dup) {
      uint64_t nb_all = 0, my_off = 0;
      for (int r = 0; r < W; ++r) { if (r == R) my_off = nb_all; nb_all += all[(size_t)r * M + 4]; }
      FJ_TRY(all_bk.ensure(std::max<uint64_t>(nb_all, 1) * 8));
      FJ_TRY(all_bv.ensure(std::max<uint64_t>(nb_all, 1) * 8));
      FJ_CUDA(cudaEventRecord(ev[10], st));
      if (nb) {
        FJ_CUDA(cudaMemcpyAsync(all_bk.as<unsigned long long>() + my_off, d_bk, nb * 8, cudaMemcpyDeviceToDevice, st));
        FJ_CUDA(cudaMemcpyAsync(all_bv.as<unsigned long long>() + my_off, d_bv, nb * 8, cudaMemcpyDeviceToDevice, st));
      }
      uint64_t off = 0;
      for (int r = 0; r < W; ++r) {
        const uint64_t c = all[(size_t)r * M + 4];
        if (c) {
          FJ_TRY(dist_broadcast_u64(dist, all_bk.as<unsigned long long>() + off, c, r, st));
          FJ_TRY(dist_broadcast_u64(dist, all_bv.as<unsigned long long>() + off, c, r, st));
        }
        off += c;
      }
      FJ_CUDA(cudaEventRecord(ev[11], st));
      FJ_CUDA(cudaEventSynchronize(ev[11]));
      comm_ms += ms(10, 11);
      FJ_TRY(join_device(FJ_ALGO_SCALAR, jflags, all_bk.as<unsigned long long>(), all_bv.as<unsigned long long>(), nb_all, d_pk, np,
                         0, s));
    } else {
      FJ_TRY(finish_attempt(jflags, s));
    }

    // ---- 6. global count
    uint64_t cnt2[2] = {s->matches, 0};
    FJ_CUDA(cudaEventRecord(ev[10], st));
    FJ_CUDA(cudaMemcpyAsync(d_meta, cnt2, 16, cudaMemcpyHostToDevice, st));
    FJ_TRY(dist_allreduce_sum_u64(dist, d_meta, d_meta + 2, 1, st));
    unsigned long long tot = 0;
    FJ_CUDA(cudaMemcpyAsync(&tot, d_meta + 2, 8, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaEventRecord(ev[11], st));
    FJ_CUDA(cudaStreamSynchronize(st));
    comm_ms += ms(10, 11);
    *total = tot;
    s->comm_s += comm_ms * 1e-3;
    s->partition_s += part_ms * 1e-3;
    return FJ_OK;
  }
  return set_err(FJ_ERR_STATE, "internal: shuffle join did not converge");
}

// ---- SHUFFLE over peer memory (dense key domain; BASELINE.json configs[2] at G > 1) --------------------------
// ONE partition pass per side (k_part): partition d = low key bits, owner GPU = top log2(world) bits of d.  Every GPU
// partitions its slice into its OWN IPC-mapped partition buffer at local speed; the owner of a partition then PULLS its
// rows from every source's buffer with the bulk copies (TMA, chunks of up to 31 KB) that feed k_sjoin's input ring, so
// rows cross NVLink once (4 bytes per build row, 2 per probe row), in large transfers, overlapped with the join.
// (Storing the 32-byte sectors straight into the owner's buffer from k_part was measured first: NVLink moved only
// ~200 GB/s per GPU in 32-byte writes and the partition pass did not scale — profiles/r02q_bench_n4_peer_store.json.)
// Two small k_xsync launches replace the collectives: count push + slice-size check + barrier, and the result exchange
// (which doubles as the barrier that frees the partition buffers for the next step).  One host synchronisation per
// step; no NCCL call in the steady state (the slice sizes are exchanged with ncclAllGather only when some rank's
// sizes changed).
fj_status Engine::xpart_ensure(size_t bytes) {
  if (xp.local && xp.bytes >= bytes) return FJ_OK;
  const int W = dist.world, R = dist.rank;
  xpart_teardown();
  const size_t want = bytes + bytes / 8 + (1 << 20);
  bool ok = cudaMalloc(&xp.local, want) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = cudaMemset(xp.local, 0, want) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine, xp.local) == cudaSuccess;
  cudaGetLastError();
  FJ_TRY(dist_scratch.ensure((size_t)(W + 1) * 64 + 64));
  unsigned long long* d_h = dist_scratch.as<unsigned long long>();
  std::vector<cudaIpcMemHandle_t> all((size_t)W);
  FJ_CUDA(cudaMemcpyAsync(d_h, &mine, 64, cudaMemcpyHostToDevice, st));
  FJ_TRY(dist_allgather_u64(dist, d_h, d_h + 8, 8, st));
  FJ_CUDA(cudaMemcpyAsync(all.data(), d_h + 8, (size_t)W * 64, cudaMemcpyDeviceToHost, st));
  FJ_CUDA(cudaStreamSynchronize(st));
  xp.mapped.assign((size_t)W, nullptr);
  xp.mapped[(size_t)R] = xp.local;
  for (int r = 0; r < W && ok; ++r) {
    if (r == R) continue;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
    xp.mapped[(size_t)r] = p;
  }
  // everybody or nobody
  unsigned long long flag = ok ? 1ull : 0ull, sum = 0;
  FJ_CUDA(cudaMemcpyAsync(d_h, &flag, 8, cudaMemcpyHostToDevice, st));
  FJ_TRY(dist_allreduce_sum_u64(dist, d_h, d_h + 1, 1, st));
  FJ_CUDA(cudaMemcpyAsync(&sum, d_h + 1, 8, cudaMemcpyDeviceToHost, st));
  FJ_CUDA(cudaStreamSynchronize(st));
  if (sum != (unsigned long long)W) {
    xpart_teardown();
    xp.broken = true;
    return FJ_OK;
  }
  xp.bytes = want;
  xp.step = 0;  // fresh (zeroed) barrier words on every rank
  return FJ_OK;
}

void Engine::xpart_teardown() {
  if (st) cudaStreamSynchronize(st);
  for (size_t r = 0; r < xp.mapped.size(); ++r)
    if (xp.mapped[r] && xp.mapped[r] != xp.local) cudaIpcCloseMemHandle(xp.mapped[r]);
  xp.mapped.clear();
  if (xp.local) cudaFree(xp.local);
  xp.local = nullptr;
  xp.bytes = 0;
  cudaGetLastError();
}

fj_status Engine::join_shuffle_peer(unsigned jflags, const unsigned long long* d_bk, const unsigned long long* d_bv, uint64_t nb,
                                    const unsigned long long* d_pk, uint64_t np, fj_stats* s, uint64_t* total, bool* handled) {
  *handled = false;
  const int W = dist.world, R = dist.rank;
  // decided from what every rank knows identically
  if (!cfg["dist_peer_shuffle"] || !cfg["dense"] || !cfg["dense16"] || !cfg["narrow"] || xp.broken) return FJ_OK;
  if (jflags & (FJ_FLAG_FORCE_WIDE | FJ_FLAG_PROBE_IDX)) return FJ_OK;
  if (W < 2 || W > 8 || (W & (W - 1))) return FJ_OK;
  const bool mat = jflags & FJ_FLAG_MATERIALIZE;
  int lw = 0;
  while ((1 << lw) < W) ++lw;
  Ctl* d_ctl = ctl.as<Ctl>();
  FJ_TRY(dist_scratch.ensure(4096));
  pairs_valid = false;

  for (int attempt = 0; attempt < 3; ++attempt) {
    // ---- slice sizes of every rank (exchanged only when the device-side check of the previous step said so)
    if (!xp.meta_valid) {
      unsigned long long mine[2] = {nb, np};
      unsigned long long* d_m = dist_scratch.as<unsigned long long>();
      FJ_CUDA(cudaMemcpyAsync(d_m, mine, 16, cudaMemcpyHostToDevice, st));
      FJ_TRY(dist_allgather_u64(dist, d_m, d_m + 2, 2, st));
      FJ_CUDA(cudaMemcpyAsync(xp.meta, d_m + 2, 16 * (size_t)W, cudaMemcpyDeviceToHost, st));
      FJ_CUDA(cudaStreamSynchronize(st));
      xp.meta_valid = true;
    }
    uint64_t nb_tot = 0, np_tot = 0, nb_max = 0, np_max = 0;
    for (int r = 0; r < W; ++r) {
      nb_tot += xp.meta[2 * r]; np_tot += xp.meta[2 * r + 1];
      nb_max = std::max<uint64_t>(nb_max, xp.meta[2 * r]); np_max = std::max<uint64_t>(np_max, xp.meta[2 * r + 1]);
    }
    if (nb_tot == 0 || np_tot == 0) return FJ_OK;  // the general path knows how to answer (0, t)
    // ---- plan (identical on every rank): partitions from the global build side, capacities from the largest slice
    const Dense16Plan gp = plan_dense16(jflags, nb_tot, np_tot);
    if (!gp.ok || gp.logp < lw + 4) return FJ_OK;
    const uint32_t P = 1u << gp.logp, lpo = (uint32_t)(gp.logp - lw), ppo = 1u << lpo;
    const size_t eb = mat ? 4 : 2;
    const uint64_t cap_b = round16(cap_build(std::max<uint64_t>(nb_max, 1), P) + ((uint64_t)part_grid(mat, nb_max, di) + 2) * part_sector_elems(mat));
    const uint64_t cap_p = round16(cap_probe(std::max<uint64_t>(np_max, 1), P) + ((uint64_t)part_grid(false, np_max, di) + 2) * part_sector_elems(false));
    const size_t ctrl_bytes = (xsync_ctrl_bytes(P) + 255) & ~size_t(255);
    const size_t b_bytes = ((size_t)P * cap_b * eb + 255) & ~size_t(255);   // ppo partitions x W sources
    const size_t p_bytes = ((size_t)P * cap_p * 2 + 255) & ~size_t(255);
    FJ_TRY(xpart_ensure(ctrl_bytes + b_bytes + p_bytes));
    if (xp.broken) return FJ_OK;
    const uint32_t cs = part_cursor_stride();
    FJ_TRY(cursors.ensure(2 * (size_t)P * cs * 4));
    uint32_t* cur_b = cursors.as<uint32_t>();
    uint32_t* cur_p = cur_b + (size_t)P * cs;
    // pairs of my partitions: at most every probe row of the job (skew), expected np_tot / W: allocate for the expected
    // share with slack and let k_sjoin's block reservation fail loudly beyond it
    const uint64_t out_rows = np_tot / (uint64_t)W + np_tot / (uint64_t)(4 * W) + (1u << 20);
    if (mat) {
      const size_t ob = (size_t)(out_rows + sjoin_out_slack_pairs(di)) * 8;
      FJ_TRY(out_keys.ensure(ob));
      FJ_TRY(out_vals.ensure(ob));
      FJ_TRY(sj_tails.ensure(sjoin_tail_bytes(di)));
    }
    unsigned long long* d_res = dist_scratch.as<unsigned long long>() + 64;
    unsigned long long* h_res = reinterpret_cast<unsigned long long*>(h_spec);  // pinned scratch (8 words)

    int launches = 0;
    const unsigned long long step = xp.step++;
    XsyncArgs xa;
    for (int r = 0; r < W; ++r) xa.ctrl[r] = xp.mapped[(size_t)r];
    xa.rank = R; xa.world = W; xa.ctl = d_ctl; xa.nb = nb; xa.np = np;
    for (int i = 0; i < 2 * W; ++i) xa.meta[i] = xp.meta[i];
    xa.cur_b = cur_b; xa.cur_p = cur_p; xa.cursor_stride = cs; xa.P = P; xa.lpo = lpo; xa.result = d_res;

    FJ_CUDA(cudaEventRecord(ev[0], st));
    launch_prepare(d_ctl, nullptr, 0, cursors.p, 2 * (size_t)P * cs * 4, di, st, cs, P, part_cursor_start(mat, nb, di), part_cursor_start(false, np, di));
    ++launches;
    FJ_CUDA(cudaEventRecord(ev[1], st));
    PartArgs a;
    a.ctl = d_ctl; a.klimit = gp.klimit; a.logp = gp.logp;
    a.cursor_stride = cs;
    bool launched = true;
    if (nb) {
      a.out = static_cast<char*>(xp.local) + ctrl_bytes;
      a.in_keys = d_bk; a.in_vals = mat ? d_bv : nullptr; a.n = nb; a.cap = cap_b; a.cursor = cur_b; a.strict = true;
      a.direct_in = (mat ? cfg["part_direct_kv"] : cfg["part_direct_count_build"]) != 0;
      launched = launch_part(mat, a, di, st, &launches);
    }
    FJ_CUDA(cudaEventRecord(ev[12], st));
    if (np && launched) {
      a.out = static_cast<char*>(xp.local) + ctrl_bytes + b_bytes;
      a.in_keys = d_pk; a.in_vals = nullptr; a.n = np; a.cap = cap_p; a.cursor = cur_p; a.strict = false;
      a.direct_in = cfg["part_direct_k"] != 0;
      launched = launch_part(false, a, di, st, &launches);
    }
    xa.phase = 1; xa.seq = 3 * step + 2;
    launch_xsync(xa, st, &launches);
    FJ_CUDA(cudaEventRecord(ev[2], st));
    if (launched) {
      const uint32_t* cnt = reinterpret_cast<const uint32_t*>(static_cast<char*>(xp.local) + xsync_count_offset_bytes());
      SjoinArgs j;
      for (int r = 0; r < W; ++r) {  // source r keeps its rows of my partitions in ITS buffer: pulled over NVLink by k_sjoin's bulk copies
        j.build[r] = static_cast<char*>(xp.mapped[(size_t)r]) + ctrl_bytes;
        j.probe[r] = static_cast<char*>(xp.mapped[(size_t)r]) + ctrl_bytes + b_bytes;
      }
      j.cap_b = cap_b; j.cap_p = cap_p;
      j.p_first = (uint32_t)R * ppo; j.p_count = ppo; j.logp = gp.logp; j.nsub = W; j.slots_alloc = gp.slots; j.rot = (uint32_t)R + 1u;
      // count arrays [source][local partition]; k_sjoin indexes them with the GLOBAL partition id
      j.cnt_stride = ppo; j.cursor_stride = 1;
      j.bcnt = cnt - j.p_first; j.pcnt = cnt + (size_t)W * ppo - j.p_first;
      j.ctl = d_ctl;
      j.out_keys = mat ? out_keys.as<unsigned long long>() : nullptr;
      j.out_vals = mat ? out_vals.as<unsigned long long>() : nullptr;
      j.tails = mat ? sj_tails.as<unsigned long long>() : nullptr;
      launched = launch_sjoin(mat, j, di, st, &launches);
    }
    FJ_CUDA(cudaEventRecord(ev[3], st));
    xa.phase = 2; xa.seq = 3 * step + 3;
    launch_xsync(xa, st, &launches);
    FJ_CUDA(cudaMemcpyAsync(h_res, d_res, 32, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaEventRecord(ev[13], st));
    FJ_CUDA(cudaStreamSynchronize(st));
    FJ_CUDA(cudaGetLastError());
    s->kernel_launches += launches;
    if (!launched) return set_err(FJ_ERR_STATE, "internal: no launch configuration for the peer-memory shuffle");
    if (h_ctl->flags & CTL_PEER_TIMEOUT) return set_err(FJ_ERR_NCCL, "peer-memory shuffle timed out (a rank did not join the step)");
    const unsigned gflags = (unsigned)h_res[1];
    if (gflags & CTL_PEER_TIMEOUT) return set_err(FJ_ERR_NCCL, "peer-memory shuffle timed out on a peer");
    if (gflags & CTL_META_CHANGED) {  // same verdict on every rank: exchange the sizes and plan again
      xp.meta_valid = false;
      continue;
    }
    if (gflags & (CTL_NOT_DENSE16 | CTL_OVERFLOW | CTL_DUP)) return FJ_OK;  // same verdict on every rank: the general shuffle answers
    if (mat) {
      if (h_ctl->out_cursor < h_ctl->match_count || h_ctl->out_cursor - h_ctl->match_count > sjoin_out_slack_pairs(di))
        return set_err(FJ_ERR_STATE, "internal: %llu pairs reserved for %llu matches", (unsigned long long)h_ctl->out_cursor,
                       (unsigned long long)h_ctl->match_count);
      if (h_ctl->out_cursor > out_rows + sjoin_out_slack_pairs(di))
        return set_err(FJ_ERR_STATE, "peer-memory shuffle: %llu pairs on this rank exceed the output arena (skewed partitions)",
                       (unsigned long long)h_ctl->out_cursor);
      h_ctl->out_cursor = h_ctl->match_count;
    }
    s->attempts = attempt + 1;
    s->clear_s += ms(0, 1) * 1e-3;
    s->partition_s += ms(1, 2) * 1e-3;
    s->probe_s += ms(2, 3) * 1e-3;
    s->comm_s += ms(3, 13) * 1e-3;
    s->table_bytes = (uint64_t)b_bytes + p_bytes;
    s->path = FJ_ALGO_RADIX;
    s->narrow = 1;
    s->bloom_kind = 0;
    s->dedup_exact = 0;
    s->radix_bits1 = gp.logp;
    s->radix_bits2 = 0;
    s->dense = 2;
    s->part_build_us = (int32_t)(ms(1, 12) * 1e3f);
    s->part_probe_us = (int32_t)(ms(12, 2) * 1e3f);
    FJ_TRY(finish_attempt(jflags, s));
    *total = h_res[0];
    *handled = true;
    return FJ_OK;
  }
  return set_err(FJ_ERR_STATE, "internal: peer-memory shuffle did not settle on the slice sizes");
}

fj_status Engine::agree(fj_status local) {
  const std::string err = local == FJ_OK ? std::string() : g_err;
  if (local != FJ_OK) cudaGetLastError();
  if (dist_scratch.ensure(4096) != FJ_OK) return local != FJ_OK ? local : FJ_ERR_OOM;  // nothing collective has been started yet
  unsigned long long* d_w = dist_scratch.as<unsigned long long>() + 128;
  unsigned long long mine = local == FJ_OK ? 0ull : 1ull, sum = 0;
  if (peer.ready && dist.world <= 32 && cfg["dist_peer_reduce"]) {  // one small kernel over peer memory instead of an NCCL call
    int l0 = 0;
    uint32_t perr = 0;
    uint32_t* d_err = reinterpret_cast<uint32_t*>(static_cast<char*>(ctl.p) + 384);
    launch_peer_reduce(peer.d_ptrs, dist.rank, dist.world, ++peer.step, nullptr, mine, nullptr, d_w, d_err, st, &l0);
    FJ_CUDA(cudaMemcpyAsync(&sum, d_w, 8, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaMemcpyAsync(&perr, d_err, 4, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
    if (perr) return set_err(FJ_ERR_NCCL, "peer-memory reduction timed out (a rank did not join the call)");
  } else {
    FJ_CUDA(cudaMemcpyAsync(d_w, &mine, 8, cudaMemcpyHostToDevice, st));
    FJ_TRY(dist_allreduce_sum_u64(dist, d_w, d_w + 1, 1, st));
    FJ_CUDA(cudaMemcpyAsync(&sum, d_w + 1, 8, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
  }
  if (local != FJ_OK) { g_err = err; return local; }
  if (sum) return set_err(FJ_ERR_STATE, "a peer rank failed before the collective (see its fj_last_error)");
  return FJ_OK;
}

// ---- peer-memory exchange (CUDA IPC) ------------------------------------------------------------
// Every rank allocates one exchange buffer, publishes its IPC handle with an ncclAllGather and maps everybody
// else's buffer.  All ranks must agree on whether the peer path exists, so the per-rank outcome is summed.
fj_status Engine::peer_setup() {
  peer_teardown();
  const int W = dist.world, R = dist.rank;
  if (!dist.ready || W < 2 || W > 32 || !cfg["dist_peer"]) return FJ_OK;
  const size_t bytes = peer_buffer_bytes();  // control words + staging area (2 Mi build keys) + partial bitmap
  bool ok = cudaMalloc(&peer.local, bytes) == cudaSuccess;
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = cudaMemset(peer.local, 0, bytes) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine, peer.local) == cudaSuccess;
  cudaGetLastError();
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle travels as 8 64-bit words");
  FJ_TRY(dist_scratch.ensure((size_t)(W + 1) * 64 + 64));
  unsigned long long* d_h = dist_scratch.as<unsigned long long>();
  std::vector<cudaIpcMemHandle_t> all((size_t)W);
  FJ_CUDA(cudaMemcpyAsync(d_h, &mine, 64, cudaMemcpyHostToDevice, st));
  FJ_TRY(dist_allgather_u64(dist, d_h, d_h + 8, 8, st));
  FJ_CUDA(cudaMemcpyAsync(all.data(), d_h + 8, (size_t)W * 64, cudaMemcpyDeviceToHost, st));
  FJ_CUDA(cudaStreamSynchronize(st));
  peer.mapped.assign((size_t)W, nullptr);
  peer.mapped[(size_t)R] = peer.local;
  for (int r = 0; r < W && ok; ++r) {
    if (r == R) continue;
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
    peer.mapped[(size_t)r] = p;
  }
  if (ok) ok = cudaMalloc(reinterpret_cast<void**>(&peer.d_ptrs), (size_t)W * 8) == cudaSuccess &&
               cudaMemcpy(peer.d_ptrs, peer.mapped.data(), (size_t)W * 8, cudaMemcpyHostToDevice) == cudaSuccess;
  cudaGetLastError();
  // everybody or nobody
  unsigned long long flag = ok ? 1ull : 0ull, sum = 0;
  FJ_CUDA(cudaMemcpyAsync(d_h, &flag, 8, cudaMemcpyHostToDevice, st));
  FJ_TRY(dist_allreduce_sum_u64(dist, d_h, d_h + 1, 1, st));
  FJ_CUDA(cudaMemcpyAsync(&sum, d_h + 1, 8, cudaMemcpyDeviceToHost, st));
  FJ_CUDA(cudaStreamSynchronize(st));
  if (sum != (unsigned long long)W) {
    peer_teardown();
    return FJ_OK;  // the NCCL path stays
  }
  peer.bytes = bytes;
  peer.step = 0;
  peer.ready = true;
  return FJ_OK;
}

void Engine::peer_teardown() {
  if (st) cudaStreamSynchronize(st);
  for (size_t r = 0; r < peer.mapped.size(); ++r)
    if (peer.mapped[r] && peer.mapped[r] != peer.local) cudaIpcCloseMemHandle(peer.mapped[r]);
  peer.mapped.clear();
  if (peer.d_ptrs) cudaFree(peer.d_ptrs);
  if (peer.local) cudaFree(peer.local);
  peer.d_ptrs = nullptr;
  peer.local = nullptr;
  peer.bytes = 0;
  peer.ready = false;
  cudaGetLastError();
}

// One step of the multi-GPU count over peer memory: (root) copy the build keys into the staging area, then one
// launch of k_count_dense_peer on every rank.  h_ctl->global_count is the answer unless h_ctl->flags says otherwise.
fj_status Engine::attempt_count_peer(uint64_t dbits, int root, const unsigned long long* bk_root, bool bk_on_device, uint64_t nb,
                                     const unsigned long long* pk, uint64_t np, fj_stats* s) {
  const size_t bytes = dbits / 8;
  FJ_TRY(bloom.ensure(bytes));
  Ctl* d_ctl = ctl.as<Ctl>();
  uint32_t* gsync = reinterpret_cast<uint32_t*>(static_cast<char*>(ctl.p) + 256);
  int launches = 0;
  const unsigned long long step = ++peer.step;
  FJ_CUDA(cudaEventRecord(ev[0], st));
  if (dist.rank == root)
    FJ_CUDA(cudaMemcpyAsync(static_cast<char*>(peer.local) + peer_staging_offset_bytes(), bk_root, nb * 8,
                            bk_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
  const bool relay = nb >= (uint64_t)std::max<int64_t>(1, cfg["peer_relay_min_rows"]);
  if (!launch_count_dense_peer(nb, pk, np, bloom.as<uint32_t>(), (uint32_t)(dbits / 32), d_ctl, gsync, peer.d_ptrs, dist.rank,
                               dist.world, root, step, relay, di, st, &launches))
    return set_err(FJ_ERR_CUDA, "k_count_dense_peer: no co-resident launch configuration");
  FJ_CUDA(cudaEventRecord(ev[3], st));
  FJ_TRY(fetch_ctl(d_ctl));
  if (h_ctl->flags & CTL_PEER_TIMEOUT) return set_err(FJ_ERR_NCCL, "peer-memory exchange timed out (a rank did not join the step)");
  s->probe_s += ms(0, 3) * 1e-3;
  s->device_s += ms(0, 3) * 1e-3;
  s->kernel_launches += launches;
  s->table_bytes = bytes;
  s->path = FJ_ALGO_SCALAR;
  s->narrow = 1;
  s->bloom_kind = 3;
  s->dense = 1;
  s->attempts = 1;
  s->matches = h_ctl->match_count;
  return FJ_OK;
}

// ---- multi-GPU drivers (one process per GPU) ---------------------------------------------------
// BROADCAST (small build side, BASELINE.json configs[3]): the build rows given on `root` are
// ncclBroadcast to every rank (16*nb bytes — cheaper than shipping the >= 2x larger table), every
// rank builds its own table and probes its own probe slice; the per-rank counts are summed with
// ncclAllReduce.  Materialized pairs stay sharded on the rank that produced them.
fj_status Engine::join_dist(int mode, int algo, unsigned flags, int root, const uint64_t* bk, const uint64_t* bv,
                            size_t nb, const uint64_t* pk, size_t np, uint64_t* out_global, uint64_t* out_local,
                            double* out_seconds, fj_stats* stats) {
  if (!dist.ready) return set_err(FJ_ERR_STATE, "fj_comm_init has not been called");
  if (mode != FJ_DIST_BROADCAST && mode != FJ_DIST_SHUFFLE) return set_err(FJ_ERR_BAD_ARG, "unknown dist mode %d", mode);
  if (!out_global) return set_err(FJ_ERR_BAD_ARG, "out_matches_global is NULL");
  if (root < 0 || root >= dist.world) return set_err(FJ_ERR_BAD_ARG, "root %d out of range", root);
  if (algo < FJ_ALGO_ADAPTIVE || algo > FJ_ALGO_RADIX) return set_err(FJ_ERR_BAD_ARG, "unknown algo %d", algo);
  if (np && !pk) return set_err(FJ_ERR_BAD_ARG, "NULL probe pointer with non-zero length");
  FJ_CUDA(cudaSetDevice(di.device));
  const double t0 = now_s();
  fj_stats s;
  memset(&s, 0, sizeof(s));
  s.n_gpus = dist.world;
  const bool dev_in = flags & FJ_FLAG_DEVICE_INPUTS;
  const unsigned jflags = flags & ~FJ_FLAG_DEVICE_INPUTS;
  FJ_TRY(dist_scratch.ensure(64));
  unsigned long long* d_cnt = dist_scratch.as<unsigned long long>();

  if (mode == FJ_DIST_BROADCAST) {
    const bool is_root = dist.rank == root;
    // count on a dense key domain with the global-table path: one kernel per GPU over peer memory, no NCCL call.
    // The decision uses only what every rank knows identically (flags, nb, configuration).
    {
      const bool narrow_cfg = !(jflags & FJ_FLAG_FORCE_WIDE) && cfg["narrow"] != 0;
      const double table_bytes = (double)nb / ((double)cfg["load_pct"] / 100.0) * 8.0;
      const bool scalar_path = algo == FJ_ALGO_SCALAR ||
                               (algo == FJ_ALGO_ADAPTIVE && table_bytes <= (double)di.l2_bytes * (double)cfg["adaptive_table_l2_pct"] / 100.0);
      const uint64_t dbits = (peer.ready && cfg["dist_peer"] && cfg["dense_fused"] && !(jflags & FJ_FLAG_MATERIALIZE) && narrow_cfg &&
                              scalar_path && nb > 0 && nb * 8 <= peer_staging_bytes())
                                 ? dense_bitmap_bits(nb) : 0;
      if (dbits) {
        const unsigned long long* d_pk;
        if (dev_in) {
          d_pk = reinterpret_cast<const unsigned long long*>(pk);
        } else {
          FJ_TRY(in_pk.ensure(std::max<size_t>(np, 1) * 8));
          const double th = now_s();
          if (np) FJ_TRY(h2d(in_pk.p, pk, np * 8));
          FJ_CUDA(cudaStreamSynchronize(st));
          s.h2d_s = now_s() - th;
          s.h2d_bytes = (uint64_t)((is_root ? nb : 0) + np) * 8;
          d_pk = in_pk.as<unsigned long long>();
        }
        pairs_valid = false;
        FJ_TRY(attempt_count_peer(dbits, root, reinterpret_cast<const unsigned long long*>(bk), dev_in, nb, d_pk, np, &s));
        if (!(h_ctl->flags & CTL_NOT_DENSE)) {  // same answer on every rank: the flag depends on the build side only
          *out_global = h_ctl->global_count;
          if (out_local) *out_local = s.matches;
          s.wall_s = now_s() - t0;
          s.algorithmic_bytes = 8ull * (nb + np);
          if (out_seconds) *out_seconds = s.device_s + s.h2d_s;
          if (stats) *stats = s;
          return FJ_OK;
        }
        // a build key outside the optimistic domain: every rank falls through to the general (NCCL) path
        s.attempts = 0;
      }
    }
    // stage inputs in HBM.  Rank-local failures (the root's missing build side, an allocation, a host->device copy) are
    // agreed on before the broadcast: every rank returns an error, nobody waits in a collective
    const bool mat = jflags & FJ_FLAG_MATERIALIZE;
    const unsigned long long* d_pk = nullptr;
    const void *src_bk = nullptr, *src_bv = nullptr;  // what root broadcasts from
    const double th = now_s();
    FJ_TRY(agree([&]() -> fj_status {
      if (is_root && nb && (!bk || (mat && !bv))) return set_err(FJ_ERR_BAD_ARG, "root rank must supply the build side");
      FJ_TRY(in_bk.ensure(std::max<size_t>(nb, 1) * 8));
      FJ_TRY(in_bv.ensure(std::max<size_t>(nb, 1) * 8));
      src_bk = in_bk.p;
      src_bv = in_bv.p;
      if (dev_in) {
        d_pk = reinterpret_cast<const unsigned long long*>(pk);
        if (is_root) { src_bk = bk; src_bv = bv; }  // straight out of the caller's HBM arrays
      } else {
        FJ_TRY(in_pk.ensure(std::max<size_t>(np, 1) * 8));
        if (is_root && nb) {
          FJ_TRY(h2d(in_bk.p, bk, nb * 8));
          if (mat) FJ_TRY(h2d(in_bv.p, bv, nb * 8));
        }
        if (np) FJ_TRY(h2d(in_pk.p, pk, np * 8));
        FJ_CUDA(cudaStreamSynchronize(st));
        s.h2d_s = now_s() - th;
        s.h2d_bytes = (uint64_t)((is_root ? (mat ? 2 : 1) * nb : 0) + np) * 8;
        d_pk = in_pk.as<unsigned long long>();
      }
      return FJ_OK;
    }()));
    // broadcast the raw build rows (keys and values in one NCCL group; a count never reads the values, so only
    // the keys travel then)
    FJ_CUDA(cudaEventRecord(ev[4], st));
    bool peer_bcast_used = false;
    const unsigned long long* d_bkeys = in_bk.as<unsigned long long>();
    const unsigned long long* d_bvals = mat ? in_bv.as<unsigned long long>() : in_bk.as<unsigned long long>();
    if (nb) {
      // Build sides that fit the exchange buffer's staging area travel over peer memory in two hops (k_peer_bcast: the
      // root sends every row once, the ranks gather the slices from each other) — decided from nb and the configuration,
      // which every rank knows identically.  Keys and values land in ONE buffer (keys first).
      const uint64_t nb2 = (nb + 1) & ~uint64_t(1);  // both halves 16-byte aligned
      const uint64_t words = mat ? 2 * nb2 : nb2;
      bool done = false;
      if (peer.ready && cfg["dist_peer_bcast"] && words * 8 <= peer_staging_bytes()) {
        FJ_TRY(bcast_rows.ensure(words * 8));
        unsigned long long* stage = reinterpret_cast<unsigned long long*>(static_cast<char*>(peer.local) + peer_staging_offset_bytes());
        if (is_root) {
          FJ_CUDA(cudaMemcpyAsync(stage, src_bk, nb * 8, cudaMemcpyDeviceToDevice, st));
          if (mat) FJ_CUDA(cudaMemcpyAsync(stage + nb2, src_bv, nb * 8, cudaMemcpyDeviceToDevice, st));
        }
        int l0 = 0;
        uint32_t* gsync = reinterpret_cast<uint32_t*>(static_cast<char*>(ctl.p) + 256);
        uint32_t* d_err = reinterpret_cast<uint32_t*>(static_cast<char*>(ctl.p) + 384);  // zero since fj_init; read back with the result
        const unsigned long long step = ++peer.step;
        if (!launch_peer_bcast(peer.d_ptrs, dist.rank, dist.world, root, step, words, bcast_rows.as<unsigned long long>(), d_err, gsync, di, st,
                               &l0))
          return set_err(FJ_ERR_CUDA, "k_peer_bcast: no co-resident launch configuration");
        s.kernel_launches += l0;
        peer_bcast_used = true;
        d_bkeys = bcast_rows.as<unsigned long long>();
        d_bvals = mat ? d_bkeys + nb2 : d_bkeys;
        done = true;
      }
      if (!done) {
        if (mat) FJ_TRY(dist_broadcast2_u64(dist, src_bk, in_bk.p, src_bv, in_bv.p, nb, root, st));
        else FJ_TRY(dist_broadcast_oop_u64(dist, src_bk, in_bk.p, nb, root, st));
      }
    }
    FJ_CUDA(cudaEventRecord(ev[5], st));
    // count: the all-reduce of the control block rides behind the first attempt's kernels (one host sync per step)
    spec_ar = !mat && cfg["dist_spec_allreduce"] != 0;
    spec_done = false;
    fj_status js = join_device(algo, jflags, d_bkeys, d_bvals, nb, d_pk, np, 0, &s);
    const bool spec = spec_ar;
    spec_ar = false;
    if (js != FJ_OK) return js;
    unsigned long long total = 0;
    bool need_plain = true;
    if (spec) {
      if (!spec_done) {  // no kernel ran on this rank (empty side): contribute an all-zero control block
        spec_ar = true;
        launch_init_ctl(ctl.as<Ctl>(), st);
        fj_status hs = spec_allreduce();
        spec_ar = false;
        if (hs != FJ_OK) return hs;
        FJ_CUDA(cudaStreamSynchronize(st));
      }
      // word 4 = flags | pad << 32 summed over the ranks: zero <=> nobody had to retry, word 0 is the global count
      if (h_spec[4] == 0) {
        total = h_spec[0];
        need_plain = false;
      }
    }
    FJ_CUDA(cudaEventRecord(ev[0], st));
    if (need_plain) {
      // global count: all-reduce straight from the control block the probe kernel wrote
      const void* cnt_src = &ctl.as<Ctl>()->match_count;
      if (nb == 0 || np == 0) {  // no kernel ran on this rank
        FJ_CUDA(cudaMemsetAsync(d_cnt, 0, 8, st));
        cnt_src = d_cnt;
      }
      if (peer.ready && dist.world <= 32 && cfg["dist_peer_reduce"]) {
        int l0 = 0;
        launch_peer_reduce(peer.d_ptrs, dist.rank, dist.world, ++peer.step, static_cast<const unsigned long long*>(cnt_src), 0, nullptr, d_cnt + 1,
                           reinterpret_cast<uint32_t*>(static_cast<char*>(ctl.p) + 384), st, &l0);
        s.kernel_launches += l0;
        peer_bcast_used = true;  // (the error word is read back below)
      } else {
        FJ_TRY(dist_allreduce_sum_u64(dist, cnt_src, d_cnt + 1, 1, st));
      }
      FJ_CUDA(cudaMemcpyAsync(&total, d_cnt + 1, 8, cudaMemcpyDeviceToHost, st));
    }
    FJ_CUDA(cudaEventRecord(ev[1], st));
    uint32_t bcast_err = 0;
    if (peer_bcast_used)
      FJ_CUDA(cudaMemcpyAsync(&bcast_err, static_cast<char*>(ctl.p) + 384, 4, cudaMemcpyDeviceToHost, st));
    FJ_CUDA(cudaStreamSynchronize(st));
    if (bcast_err) return set_err(FJ_ERR_NCCL, "peer-memory broadcast timed out (a rank did not join the step)");
    s.comm_s = (ms(4, 5) + ms(0, 1)) * 1e-3;
    s.device_s += s.comm_s;
    *out_global = total;
  } else {
    if ((nb && (!bk || !bv))) return set_err(FJ_ERR_BAD_ARG, "NULL build pointer with non-zero length");
    if (flags & FJ_FLAG_PROBE_IDX) return set_err(FJ_ERR_BAD_ARG, "FJ_FLAG_PROBE_IDX is not available in shuffle mode (rows move between ranks)");
    const unsigned long long *d_bk, *d_bv, *d_pk;
    if (dev_in) {
      d_bk = reinterpret_cast<const unsigned long long*>(bk);
      d_bv = reinterpret_cast<const unsigned long long*>(bv);
      d_pk = reinterpret_cast<const unsigned long long*>(pk);
    } else {
      FJ_TRY(in_bk.ensure(std::max<size_t>(nb, 1) * 8));
      FJ_TRY(in_bv.ensure(std::max<size_t>(nb, 1) * 8));
      FJ_TRY(in_pk.ensure(std::max<size_t>(np, 1) * 8));
      const double th = now_s();
      if (nb) {
        FJ_TRY(h2d(in_bk.p, bk, nb * 8));
        FJ_TRY(h2d(in_bv.p, bv, nb * 8));
      }
      if (np) FJ_TRY(h2d(in_pk.p, pk, np * 8));
      FJ_CUDA(cudaStreamSynchronize(st));
      s.h2d_s = now_s() - th;
      s.h2d_bytes = (uint64_t)(2 * nb + np) * 8;
      d_bk = in_bk.as<unsigned long long>();
      d_bv = in_bv.as<unsigned long long>();
      d_pk = in_pk.as<unsigned long long>();
    }
    FJ_CUDA(cudaEventRecord(ev[4], st));
    uint64_t tot = 0;
    bool handled = false;
    if (algo != FJ_ALGO_SCALAR) FJ_TRY(join_shuffle_peer(jflags, d_bk, d_bv, nb, d_pk, np, &s, &tot, &handled));
    if (!handled) FJ_TRY(join_shuffle(algo, jflags, d_bk, d_bv, nb, d_pk, np, &s, &tot));
    FJ_CUDA(cudaEventRecord(ev[5], st));
    FJ_CUDA(cudaEventSynchronize(ev[5]));
    s.device_s = ms(4, 5) * 1e-3;  // whole distributed join on this rank, exchange included
    *out_global = tot;
  }
  if (out_local) *out_local = s.matches;
  s.wall_s = now_s() - t0;
  s.algorithmic_bytes = (jflags & FJ_FLAG_MATERIALIZE) ? 16ull * nb + 8ull * np + 16ull * s.matches : 8ull * (nb + np);
  if (out_seconds) *out_seconds = s.device_s + s.h2d_s;
  if (stats) *stats = s;
  return FJ_OK;
}

}  // namespace fj

// ================================================================================================
// C ABI
// ================================================================================================
using namespace fj;

extern "C" {

FJ_API const char* fj_last_error(void) { return g_err.c_str(); }
FJ_API const char* fj_version(void) { return "flashjoin_b200 0.1.0 (sm_100a)"; }

FJ_API fj_status fj_init(int device) {
  std::lock_guard<std::mutex> lk(E().mu);
  return E().init(device);
}
FJ_API fj_status fj_shutdown(void) {
  std::lock_guard<std::mutex> lk(E().mu);
  E().shutdown();
  return FJ_OK;
}
FJ_API fj_status fj_device_count(int* count) {
  if (!count) return set_err(FJ_ERR_BAD_ARG, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
  *count = n;
  return FJ_OK;
}

FJ_API fj_status fj_join_u64(int algo, unsigned flags, const uint64_t* bk, const uint64_t* bv, size_t nb,
                             const uint64_t* pk, size_t np, uint64_t* out_matches, double* out_seconds,
                             fj_stats* stats) {
  std::lock_guard<std::mutex> lk(E().mu);
  return E().join(algo, flags, bk, bv, nb, pk, np, out_matches, out_seconds, stats);
}

FJ_API fj_status fj_pairs_count(uint64_t* n) {
  std::lock_guard<std::mutex> lk(E().mu);
  if (!n) return set_err(FJ_ERR_BAD_ARG, "n is NULL");
  if (!E().pairs_valid) return set_err(FJ_ERR_STATE, "no materialized pairs: the last join was not a materialize call");
  *n = E().pairs_n;
  return FJ_OK;
}
FJ_API fj_status fj_pairs_fetch(uint64_t* keys, uint64_t* values, uint64_t* probe_idx_or_null, size_t capacity) {
  std::lock_guard<std::mutex> lk(E().mu);
  Engine& e = E();
  if (!e.pairs_valid) return set_err(FJ_ERR_STATE, "no materialized pairs: the last join was not a materialize call");
  if (capacity < e.pairs_n) return set_err(FJ_ERR_BAD_ARG, "capacity %zu < %llu pairs", capacity, (unsigned long long)e.pairs_n);
  if (e.pairs_n == 0) return FJ_OK;
  if (!keys || !values) return set_err(FJ_ERR_BAD_ARG, "NULL output pointer");
  if (probe_idx_or_null && !e.pairs_idx) return set_err(FJ_ERR_STATE, "probe indices were not requested (FJ_FLAG_PROBE_IDX)");
  FJ_CUDA(cudaSetDevice(e.di.device));
  FJ_TRY(e.d2h(keys, e.out_keys.p, e.pairs_n * 8));
  FJ_TRY(e.d2h(values, e.out_vals.p, e.pairs_n * 8));
  if (probe_idx_or_null) FJ_TRY(e.d2h(probe_idx_or_null, e.out_idx.p, e.pairs_n * 8));
  return FJ_OK;
}
FJ_API fj_status fj_pairs_device(const uint64_t** keys, const uint64_t** values, const uint64_t** probe_idx, uint64_t* n) {
  std::lock_guard<std::mutex> lk(E().mu);
  Engine& e = E();
  if (!e.pairs_valid) return set_err(FJ_ERR_STATE, "no materialized pairs: the last join was not a materialize call");
  if (keys) *keys = e.out_keys.as<uint64_t>();
  if (values) *values = e.out_vals.as<uint64_t>();
  if (probe_idx) *probe_idx = e.pairs_idx ? e.out_idx.as<uint64_t>() : nullptr;
  if (n) *n = e.pairs_n;
  return FJ_OK;
}

FJ_API fj_status fj_config_set(const char* key, int64_t value) {
  std::lock_guard<std::mutex> lk(E().mu);
  if (!key) return set_err(FJ_ERR_BAD_ARG, "key is NULL");
  auto it = E().cfg.find(key);
  if (it == E().cfg.end()) return set_err(FJ_ERR_BAD_ARG, "unknown config key '%s'", key);
  it->second = value;
  return FJ_OK;
}
FJ_API fj_status fj_config_get(const char* key, int64_t* value) {
  std::lock_guard<std::mutex> lk(E().mu);
  if (!key || !value) return set_err(FJ_ERR_BAD_ARG, "NULL argument");
  auto it = E().cfg.find(key);
  if (it == E().cfg.end()) return set_err(FJ_ERR_BAD_ARG, "unknown config key '%s'", key);
  *value = it->second;
  return FJ_OK;
}

FJ_API fj_status fj_dev_alloc(void** ptr, size_t bytes) {
  if (!ptr) return set_err(FJ_ERR_BAD_ARG, "ptr is NULL");
  { std::lock_guard<std::mutex> lk(E().mu); FJ_TRY(E().init(-1)); FJ_CUDA(cudaSetDevice(E().di.device)); }
  FJ_CUDA(cudaMalloc(ptr, bytes ? bytes : 1));
  return FJ_OK;
}
static void select_engine_device() {
  std::lock_guard<std::mutex> lk(E().mu);
  if (E().inited) cudaSetDevice(E().di.device);
}
FJ_API fj_status fj_dev_free(void* ptr) {
  select_engine_device();
  if (ptr) FJ_CUDA(cudaFree(ptr));
  return FJ_OK;
}
FJ_API fj_status fj_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes) {
  select_engine_device();
  if (bytes) FJ_CUDA(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
  return FJ_OK;
}
FJ_API fj_status fj_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes) {
  select_engine_device();
  if (bytes) FJ_CUDA(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
  return FJ_OK;
}
FJ_API fj_status fj_host_alloc_pinned(void** ptr, size_t bytes) {
  if (!ptr) return set_err(FJ_ERR_BAD_ARG, "ptr is NULL");
  { std::lock_guard<std::mutex> lk(E().mu); FJ_TRY(E().init(-1)); FJ_CUDA(cudaSetDevice(E().di.device)); }
  FJ_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
  return FJ_OK;
}
FJ_API fj_status fj_host_free_pinned(void* ptr) {
  if (ptr) FJ_CUDA(cudaFreeHost(ptr));
  return FJ_OK;
}
FJ_API fj_status fj_device_synchronize(void) {
  select_engine_device();
  FJ_CUDA(cudaDeviceSynchronize());
  return FJ_OK;
}

FJ_API fj_status fj_generate_g2(int side, uint64_t n_total, uint64_t ny, int match_pct, uint64_t seed, uint64_t start,
                                uint64_t count, uint64_t* keys_dev, uint64_t* values_dev_or_null) {
  (void)n_total;
  std::lock_guard<std::mutex> lk(E().mu);
  if (side != 0 && side != 1) return set_err(FJ_ERR_BAD_ARG, "side must be 0 (build) or 1 (probe)");
  if (!keys_dev && count) return set_err(FJ_ERR_BAD_ARG, "keys_dev is NULL");
  if (ny == 0 || match_pct < 0 || match_pct > 100) return set_err(FJ_ERR_BAD_ARG, "bad generator parameters");
  FJ_TRY(E().init(-1));
  FJ_CUDA(cudaSetDevice(E().di.device));
  // mirrors flash_hash_join_b200/datagen.py:_perm_index
  const uint64_t c = ny * (uint64_t)match_pct / 100, U = 2 * ny - c;
  uint64_t a = 0x9E3779B1ull | 1ull;
  auto gcd = [](uint64_t x, uint64_t y) { while (y) { uint64_t t = x % y; x = y; y = t; } return x; };
  while (gcd(a, U) != 1) a += 2;
  const uint64_t b = (uint64_t)(((unsigned __int128)seed * 0x85EBCA6Bull + 12345ull) % U);
  launch_generate_g2(side, ny, c, U, a % U, b, seed, start, count, reinterpret_cast<unsigned long long*>(keys_dev),
                     reinterpret_cast<unsigned long long*>(values_dev_or_null), E().st);
  FJ_CUDA(cudaStreamSynchronize(E().st));
  FJ_CUDA(cudaGetLastError());
  return FJ_OK;
}

FJ_API fj_status fj_flush_l2(void) {
  std::lock_guard<std::mutex> lk(E().mu);
  Engine& e = E();
  FJ_TRY(e.init(-1));
  FJ_CUDA(cudaSetDevice(e.di.device));
  const size_t bytes = std::max<size_t>((size_t)e.di.l2_bytes * 2, size_t(256) << 20);
  FJ_TRY(e.flush.ensure(bytes));
  FJ_CUDA(cudaMemsetAsync(e.flush.p, 0x5a, bytes, e.st));
  FJ_CUDA(cudaStreamSynchronize(e.st));
  return FJ_OK;
}

FJ_API fj_status fj_timer_start(void) {
  std::lock_guard<std::mutex> lk(E().mu);
  Engine& e = E();
  FJ_TRY(e.init(-1));
  FJ_CUDA(cudaSetDevice(e.di.device));
  FJ_CUDA(cudaEventRecord(e.ev[6], e.st));
  return FJ_OK;
}
FJ_API fj_status fj_timer_stop(double* seconds) {
  std::lock_guard<std::mutex> lk(E().mu);
  Engine& e = E();
  if (!seconds) return set_err(FJ_ERR_BAD_ARG, "seconds is NULL");
  if (!e.inited) return set_err(FJ_ERR_STATE, "fj_timer_start has not been called");
  FJ_CUDA(cudaSetDevice(e.di.device));
  FJ_CUDA(cudaEventRecord(e.ev[7], e.st));
  FJ_CUDA(cudaEventSynchronize(e.ev[7]));
  float m = 0.f;
  FJ_CUDA(cudaEventElapsedTime(&m, e.ev[6], e.ev[7]));
  *seconds = (double)m * 1e-3;
  return FJ_OK;
}

// ---- multi-GPU ---------------------------------------------------------------------------------
FJ_API fj_status fj_comm_unique_id(void* id128) {
  if (!id128) return set_err(FJ_ERR_BAD_ARG, "id128 is NULL");
  return dist_unique_id(id128);
}
FJ_API fj_status fj_comm_init(int rank, int world, const void* id128) {
  std::lock_guard<std::mutex> lk(E().mu);
  if (!id128 || world < 1 || rank < 0 || rank >= world) return set_err(FJ_ERR_BAD_ARG, "bad rank/world/id");
  FJ_TRY(E().init(-1));
  FJ_CUDA(cudaSetDevice(E().di.device));
  FJ_TRY(dist_init(E().dist, rank, world, id128));
  FJ_TRY(E().peer_setup());  // (its all-gather and all-reduce also pay NCCL's first-use cost of those collectives)
  // first use of a broadcast and of point-to-point channels costs 0.5 - 3 s (connection setup): pay it here, not in
  // the first join (profiles/r01m4_dist_check_2gpu.log)
  Engine& e = E();
  if (world > 1 && e.cfg["dist_warmup"]) {
    FJ_TRY(e.dist_scratch.ensure(4096));
    unsigned long long* w = e.dist_scratch.as<unsigned long long>();
    FJ_CUDA(cudaMemsetAsync(w, 0, 4096, e.st));
    FJ_TRY(dist_broadcast_u64(e.dist, w, 1, 0, e.st));
    std::vector<DistMsg> sends, recvs;
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      sends.push_back({r, reinterpret_cast<char*>(w + 8), 8});
      recvs.push_back({r, reinterpret_cast<char*>(w + 16 + r), 8});
    }
    FJ_TRY(dist_exchange(e.dist, sends.data(), sends.size(), recvs.data(), recvs.size(), e.st));
    FJ_CUDA(cudaStreamSynchronize(e.st));
  }
  return FJ_OK;
}
FJ_API fj_status fj_comm_destroy(void) {
  std::lock_guard<std::mutex> lk(E().mu);
  E().xpart_teardown();
  E().xp = Engine::XPart();
  E().peer_teardown();
  dist_destroy(E().dist);
  return FJ_OK;
}
FJ_API fj_status fj_join_dist_u64(int mode, int algo, unsigned flags, int root, const uint64_t* bk, const uint64_t* bv,
                                  size_t nb, const uint64_t* pk, size_t np, uint64_t* out_matches_global,
                                  uint64_t* out_matches_local, double* out_seconds, fj_stats* stats) {
  std::lock_guard<std::mutex> lk(E().mu);
  return E().join_dist(mode, algo, flags, root, bk, bv, nb, pk, np, out_matches_global, out_matches_local, out_seconds,
                       stats);
}

}  // extern "C"
