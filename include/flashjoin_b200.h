/*
 * flashjoin_b200.h — C ABI of the B200-native equi-join engine (libflashjoin_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference has no C ABI of its
 * own: its only interface is the pybind11 module `flash_join` (/root/reference/hash_join.cpp:598-640),
 * whose 12 join entry points are instantiations of four driver templates.  Every entry point
 * below names the reference function(s) it replaces.  A reference maintainer binds these with
 * pybind11 / ctypes / cgo exactly as shown in INTEGRATION.md; the repo's own binding is
 * flash_hash_join_b200/csrc/flash_join_py.cpp.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no C++/CUDA/torch types cross the boundary.
 *   - every function returns an fj_status (0 = ok, negative = error class); the message of the
 *     last error on the calling thread is available from fj_last_error().
 *   - keys and values are 64-bit words (the reference takes py::array_t<uint64_t>,
 *     hash_join.cpp:316; int64 inputs are the same bits).
 *   - there is NO CPU fallback: without a usable sm_100 device every join call fails with
 *     FJ_ERR_NO_DEVICE / FJ_ERR_CUDA.
 *   - calls on one engine are serialised by an internal mutex (the reference is serialised by the
 *     GIL, hash_join.cpp has no gil_scoped_release).
 */
#ifndef FLASHJOIN_B200_H_
#define FLASHJOIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FJ_API __attribute__((visibility("default")))
#else
#define FJ_API
#endif

typedef int fj_status;
enum {
  FJ_OK = 0,
  FJ_ERR_BAD_ARG = -1,   /* NULL pointer with non-zero size, unknown algo/flag, ...            */
  FJ_ERR_CUDA = -2,      /* a CUDA runtime call failed (message has the CUDA error string)    */
  FJ_ERR_NCCL = -3,      /* NCCL missing or a collective failed                               */
  FJ_ERR_OOM = -4,       /* device or pinned-host allocation failed                           */
  FJ_ERR_NO_DEVICE = -5, /* no CUDA device, or the device is not compute capability 10.x      */
  FJ_ERR_STATE = -6      /* call order (e.g. fj_pairs_fetch before any materialize join)      */
};

/* Join strategy.  Replaces the choice between the reference's driver templates:
 *   FJ_ALGO_SCALAR   -> _hash_join_scalar_count / _hash_join_scalar_materialize (hash_join.cpp:536-567, :383-496)
 *   FJ_ALGO_RADIX    -> _hash_join_radix_count  / _hash_join_radix_materialize  (hash_join.cpp:498-534, :315-381)
 *   FJ_ALGO_ADAPTIVE -> adaptive_hash_join_count / adaptive_hash_join_materialize (hash_join.cpp:578-594);
 *                       the small/large switch is re-derived from the device's L2 size, not 1'000'000 rows. */
enum { FJ_ALGO_ADAPTIVE = 0, FJ_ALGO_SCALAR = 1, FJ_ALGO_RADIX = 2 };

/* Flags (bit-or). */
enum {
  FJ_FLAG_BLOOM = 1u << 0,         /* the *_bloom entry points: FlashHashTable<true> (hash_join.cpp:75, :185-189) */
  FJ_FLAG_MATERIALIZE = 1u << 1,   /* produce (probe key, build value) pairs (hash_join.cpp:351-352, :435-436, :466-467) */
  FJ_FLAG_DEVICE_INPUTS = 1u << 2, /* bk/bv/pk are device pointers already resident in HBM (no H2D inside the call) */
  FJ_FLAG_FORCE_WIDE = 1u << 3,    /* disable the optimistic packed 32-bit slot path; always use 16-byte slots */
  FJ_FLAG_PROBE_IDX = 1u << 4      /* materialize also records the probe row index of every pair (scalar path only) */
};

/* Per-call statistics (all times in seconds; device times are CUDA-event times on the engine's stream). */
typedef struct fj_stats {
  double h2d_s;        /* host->device copy of the inputs (0 with FJ_FLAG_DEVICE_INPUTS)          */
  double clear_s;      /* hash-table / bloom / counter initialisation                              */
  double build_s;      /* build-side insert kernels (scalar path)                                  */
  double partition_s;  /* radix histogram + scatter kernels (radix path)                           */
  double probe_s;      /* probe / partition-join kernels, including pair compaction                */
  double comm_s;       /* NCCL broadcast / all-to-all / all-reduce (distributed calls)             */
  double device_s;     /* whole join on the device (SimpleTimer scope of hash_join.cpp:319-379 etc.  */
                       /* minus host<->device copies); the `seconds` returned = device_s + h2d_s     */
  double wall_s;       /* host wall clock of the whole call, copies included                       */
  uint64_t matches;
  uint64_t table_bytes;       /* hash table (+bloom) footprint in HBM                              */
  uint64_t algorithmic_bytes; /* count: 8(nb+np); materialize: 16nb + 8np + 16m    */
  uint64_t h2d_bytes;
  int32_t path;        /* FJ_ALGO_SCALAR or FJ_ALGO_RADIX actually taken                           */
  int32_t narrow;      /* 1 = packed 32-bit key|value slots were used, 0 = 16-byte slots           */
  int32_t bloom_kind;  /* 0 none, 1 shared-memory resident, 2 global (L2) resident,                */
                       /* 3 exact membership bitmap (dense key domain / k_join3's partition bitmap), */
                       /* 4 per-partition filter in shared memory (radix + Bloom, k_join)          */
  int32_t attempts;    /* 1 normally; >1 when an optimistic attempt was abandoned and re-run       */
  int32_t dedup_exact; /* 1 = duplicate build keys were seen and the keep-first slow path ran      */
  int32_t kernel_launches; /* number of this library's join kernels launched by the call (the      */
                           /* small kernel that hands the verdict to the host is not counted)      */
  int32_t radix_bits1, radix_bits2; /* fan-out of the radix passes (0 = pass not run)              */
  int32_t n_gpus;
  int32_t dense;       /* 1 = a dense-key-domain fast path produced the result (bitmap count or     */
                       /* L2-resident direct-address radix join), 2 = the one-pass partition +     */
                       /* shared-memory direct-address join (k_part / k_sjoin); 0 = general path   */
  int32_t part_build_us; /* radix path: device time of the build-side partition pass(es), microseconds         */
  int32_t part_probe_us; /* radix path: device time of the probe-side partition pass(es), microseconds         */
  int32_t reserved[4];
} fj_stats;

/* ---- lifecycle -------------------------------------------------------------------------------
 * Replaces initialize_memory_system() / flash_join.initialize() (hash_join.cpp:596, :639): creates
 * the CUDA context on `device` (-1 = $LOCAL_RANK when set — one process per GPU — else the calling
 * thread's current CUDA device), the engine stream and the reusable device arena.  Idempotent for
 * the same device; FJ_ERR_STATE when the engine already lives on another device.  fj_shutdown
 * releases everything. */
FJ_API fj_status fj_init(int device);
FJ_API fj_status fj_shutdown(void);
FJ_API fj_status fj_device_count(int* count);
FJ_API const char* fj_last_error(void);
FJ_API const char* fj_version(void);

/* ---- the join ---------------------------------------------------------------------------------
 * One call = one of the reference's 12 entry points (hash_join.cpp:603-637), selected by
 * (algo, flags & FJ_FLAG_BLOOM, flags & FJ_FLAG_MATERIALIZE):
 *     num_matches = |{ j : probe_keys[j] in set(build_keys) }|   (build side de-duplicated on key,
 *     keep-first: hash_join.cpp:125; each probe row matches at most once: :176)
 * bk/bv have nb elements, pk has np elements (host pointers unless FJ_FLAG_DEVICE_INPUTS).
 * *out_matches receives the count, *out_seconds the time of the join: device time plus, for host
 * inputs, the host->device copy of the columns (fj_stats.device_s + .h2d_s; the reference's seconds
 * likewise cover the whole join from the arrays it was given); both mirror the reference's return
 * tuple (py::int_ total_results, double core_duration_sec).
 * With FJ_FLAG_MATERIALIZE the pairs stay in HBM until the next join call and can be read with
 * fj_pairs_*.  `stats` may be NULL. */
FJ_API fj_status fj_join_u64(int algo, unsigned flags,
                             const uint64_t* bk, const uint64_t* bv, size_t nb,
                             const uint64_t* pk, size_t np,
                             uint64_t* out_matches, double* out_seconds, fj_stats* stats);

/* Materialized pairs of the last FJ_FLAG_MATERIALIZE call (the reference computes result_keys /
 * result_values and drops them, hash_join.cpp:365-380; here they can be fetched).  Order is
 * unspecified; parity is on the sorted multiset of (probe key, build value). */
FJ_API fj_status fj_pairs_count(uint64_t* n);
FJ_API fj_status fj_pairs_fetch(uint64_t* keys, uint64_t* values, uint64_t* probe_idx_or_null, size_t capacity);
FJ_API fj_status fj_pairs_device(const uint64_t** keys, const uint64_t** values, const uint64_t** probe_idx, uint64_t* n);

/* ---- tunables (runtime counterparts of the reference's compile-time constants,
 * hash_join.cpp:38, :79, :99, :302, :393, :576) ------------------------------------------------
 * keys: "load_pct" (table load factor, %), "load_pct_auto" (1: small tables drop to 25 %), "bloom_bits_per_key", "adaptive_table_l2_pct",
 *       "radix_sub_rows" (target build rows per shared-memory partition), "radix_optimistic",
 *       "smem_bloom" (0/1), "join3" (0/1: collision-free pipelined partition join for packed rows), "probe_ctas_per_sm", "chunk_rows" (host-input pipelining chunk),
 *       "dense" (0/1: optimistic dense-key-domain fast paths), "dense_min_rows" (smallest build side that takes the
 *       direct-address radix join), "dense_group_mb" (MB of direct-address regions per L2 pipeline stage),
 *       "dense_ring" / "dense_batch" (k_djoin: per-CTA item look-ahead), "dense_delay_b" / "dense_delay_p" (k_djoin: steps
 *       between zeroing, filling and probing a group of regions), "dense_fused" (0/1: bitmap count / materialize as one
 *       persistent launch), "dist_peer" (0/1: multi-GPU count over IPC-mapped peer memory instead of NCCL; read at
 *       fj_comm_init and per call), "dist_spec_allreduce" (0/1: NCCL count all-reduce enqueued behind the first attempt).
 *       "stage_threads" / "stage_min_mb" (host threads that stage pageable input columns of at least that many MB through
 *       pinned buffers; 0 threads = plain cudaMemcpyAsync).
 *       Round 2: "dense16" (0/1: the one-pass dense-key-domain radix join k_part + k_sjoin), "dense16_logp" (partitions,
 *       0 = derive), "dense16_min_rows" / "dense16_min_probe" (adaptive materialize: smallest build / probe side that prefers
 *       dense16), "dense16_sel_min_pct" / "dense16_sel_max_rows" (adaptive materialize with a small build side: sampled match
 *       rate below which the dense table path is taken, and the largest build side that is sampled), "bloom_guard" (0/1: skip a
 *       filter that spills out of shared memory next to an L2-resident table), "mapped_result" (0/1: the attempt's control
 *       block returns through mapped pinned memory instead of a device->host copy), "dist_peer_shuffle" / "dist_peer_bcast" /
 *       "dist_peer_reduce" (0/1: SHUFFLE / build broadcast / 8-byte reductions over IPC-mapped peer memory instead of NCCL),
 *       "peer_relay_min_rows" (build rows from which k_count_dense_peer relays key slices and partial bitmaps).
 *       Every key can also be preset from the environment as FJ_CFG_<KEY>=<integer>. */
FJ_API fj_status fj_config_set(const char* key, int64_t value);
FJ_API fj_status fj_config_get(const char* key, int64_t* value);

/* ---- device / pinned memory helpers for callers that keep inputs resident in HBM ------------- */
FJ_API fj_status fj_dev_alloc(void** ptr, size_t bytes);
FJ_API fj_status fj_dev_free(void* ptr);
FJ_API fj_status fj_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes);
FJ_API fj_status fj_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes);
FJ_API fj_status fj_host_alloc_pinned(void** ptr, size_t bytes);
FJ_API fj_status fj_host_free_pinned(void* ptr);
FJ_API fj_status fj_device_synchronize(void);
/* Fill device arrays with generator G2 (flash_hash_join_b200/datagen.py:g2_slice) directly in HBM:
 * side 0 = build (keys + values), side 1 = probe (keys; values_or_null ignored). */
FJ_API fj_status fj_generate_g2(int side, uint64_t n_total, uint64_t ny, int match_pct, uint64_t seed,
                                uint64_t start, uint64_t count, uint64_t* keys_dev, uint64_t* values_dev_or_null);
/* Write `bytes` of device memory (> L2) so that the next timed call starts with a cold L2. */
FJ_API fj_status fj_flush_l2(void);
/* CUDA-event stopwatch on the engine's own stream (the stream every kernel of this library is
 * launched on): fj_timer_start records an event, fj_timer_stop records a second one, waits for it
 * and returns the elapsed device time between the two, idle gaps between calls included.
 * Replaces SimpleTimer (hash_join.cpp:45-55) for callers timing several joins back to back. */
FJ_API fj_status fj_timer_start(void);
FJ_API fj_status fj_timer_stop(double* seconds);

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ----------------------------------------
 * No reference counterpart (hash_join.cpp is single-process, std::thread only).  The caller
 * distributes a 128-byte NCCL unique id produced by rank 0 (any transport: torch.distributed,
 * MPI, a file) and every rank calls fj_comm_init.  fj_comm_init also maps a 16 MB exchange buffer of every
 * peer GPU through CUDA IPC (GPUs of one node); a broadcast-mode COUNT on a dense key domain then runs as one
 * kernel per GPU over that peer memory (build keys read from the root GPU, counts exchanged with system-scope
 * loads/stores over NVLink) with no NCCL call in the join.  Like a collective, a distributed join blocks until
 * every rank has made the same call.
 *   fj_join_dist_u64, FJ_DIST_BROADCAST: the build side lives on rank `root` (every rank passes the same
 *     nb; bk/bv are only read on root and may be NULL elsewhere) and is ncclBroadcast to all ranks, each
 *     rank builds locally and probes ITS OWN probe slice (pk/np are rank-local); counts are summed with
 *     ncclAllReduce.  *out_matches_global = global count, *out_matches_local = this rank's.
 *   FJ_DIST_SHUFFLE: every rank passes its slice of BOTH sides; rows are hash-partitioned by
 *     destination rank (already narrowed to the packed partition format when the data allows),
 *     exchanged with one grouped ncclSend/ncclRecv all-to-all-v and joined locally from the received
 *     rows.  Duplicate build keys anywhere fall back to gathering the build side on every rank
 *     (keep-first is defined on the rank-major global row order).  Materialized pairs stay on the
 *     rank that produced them (fj_pairs_*).  FJ_FLAG_PROBE_IDX is not available in this mode. */
enum { FJ_DIST_BROADCAST = 0, FJ_DIST_SHUFFLE = 1 };
FJ_API fj_status fj_comm_unique_id(void* id128);
FJ_API fj_status fj_comm_init(int rank, int world, const void* id128);
FJ_API fj_status fj_comm_destroy(void);
FJ_API fj_status fj_join_dist_u64(int mode, int algo, unsigned flags, int root,
                                  const uint64_t* bk, const uint64_t* bv, size_t nb,
                                  const uint64_t* pk, size_t np,
                                  uint64_t* out_matches_global, uint64_t* out_matches_local,
                                  double* out_seconds, fj_stats* stats);

#ifdef __cplusplus
}
#endif
#endif /* FLASHJOIN_B200_H_ */
