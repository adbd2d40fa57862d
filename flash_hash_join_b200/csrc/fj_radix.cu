// fj_radix.cu — the radix-partitioned join path: one or two scatter passes that split both sides
// into shared-memory-sized partitions, then one CTA per partition builds its hash table in shared
// memory and streams the partition's probe keys through it.
//
// Replaces, from /root/reference/hash_join.cpp:
//   get_partition_idx (:209), parallel_radix_partition_kv / _k (:210-292)   -> k_scatter<...>
//   per-partition build_local + probe_vectorized loops of _hash_join_radix_count (:515-525)
//   and _hash_join_radix_materialize (:340-356), incl. the result gather (:362-378) -> k_join<...>
//
// Differences by design (B200-first):
//   * no histogram pre-pass: partitions are fixed-capacity regions (expected size + slack) and a
//     tile reserves its run inside a region with one global atomicAdd per (tile, partition); an
//     overflowing region raises CTL_OVERFLOW and the host re-runs the join on the global-table path.
//   * fan-out is sized so a partition's build side fits shared memory (2 CTAs / SM), not "L2-ish"
//     256 partitions; with > 2^8 partitions the split is done in two passes of <= 2^9 each.
//   * ranking inside a tile uses shared-memory atomicAdd (measured 1.8 T keys/s on B200, tools/ubench.cu)
//     and the tile is written out partition-contiguous (coalesced runs).
//   * rows are narrowed while partitioning when the data allows: build tuple = key32<<32|value32,
//     probe key = 32 bits (optimistic; CTL_NEED_WIDE abandons the attempt).
#include <type_traits>

#include "fj_kernels.h"

namespace fj {

// ------------------------------------------------------------------------------------------------
// element formats after the first scatter pass
template <bool BUILD, bool NARROW> struct Elem;
template <> struct Elem<true, true> {    // packed key32|value32
  using T = unsigned long long;
  static __device__ __forceinline__ unsigned long long key(const T& e) { return e >> 32; }
};
template <> struct Elem<true, false> {   // {key64, value64}
  using T = ulonglong2;
  static __device__ __forceinline__ unsigned long long key(const T& e) { return e.x; }
};
template <> struct Elem<false, true> {   // key32
  using T = uint32_t;
  static __device__ __forceinline__ unsigned long long key(const T& e) { return e; }
};
template <> struct Elem<false, false> {  // key64
  using T = unsigned long long;
  static __device__ __forceinline__ unsigned long long key(const T& e) { return e; }
};

constexpr int SC_THREADS = 512;
constexpr int SC_IPT = 8;
constexpr int SC_TILE = SC_THREADS * SC_IPT;  // 4096 rows per tile
constexpr int SC_FMAX = 512;                  // max fan-out of one pass
constexpr long long SC_POISON = (long long)0x7fffffffffffffffLL;

// exclusive scan of one value per thread across a 512-thread block
__device__ __forceinline__ uint32_t block_excl_scan_512(uint32_t v, uint32_t* s_warp /*[16]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  uint32_t pre = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < SC_THREADS / 32; ++w) {
    const uint32_t x = s_warp[w];
    if (w < warp) pre += x;
    tot += x;
  }
  total = tot;
  __syncthreads();
  return pre + incl - v;
}

// digit of one row.  shift >= 0: radix digit = `fan` (power of two) hash bits starting at `shift` (top bits
// first, so consecutive passes refine a partition).  shift < 0: DESTINATION mode of the multi-GPU shuffle —
// the low 16 hash bits range-reduced to [0, fan), fan arbitrary (<= 512): independent of the top bits that
// the local radix passes and the table index consume afterwards.
__host__ __device__ __forceinline__ uint32_t scatter_digit(uint32_t h, int shift, uint32_t fan) {
  return shift >= 0 ? (h >> shift) & (fan - 1) : ((h & 0xffffu) * fan) >> 16;
}
uint32_t shuffle_dest_host(uint64_t key, uint32_t fan) { return scatter_digit(hash32(key), -1, fan); }

// STAGE 1: input = raw 64-bit columns (keys[, vals]); STAGE 2: input = stage-1 partitions.
template <bool BUILD, bool NARROW, int STAGE>
__global__ void __launch_bounds__(SC_THREADS, 2)
    k_scatter(const unsigned long long* __restrict__ in_keys, const unsigned long long* __restrict__ in_vals,
              uint64_t n,                                              // stage 1
              const typename Elem<BUILD, NARROW>::T* __restrict__ in_part,  // stage 2
              const uint32_t* __restrict__ in_counts, uint32_t in_nparts, uint64_t in_cap,
              typename Elem<BUILD, NARROW>::T* __restrict__ out, uint32_t* __restrict__ out_cursor, uint64_t out_cap,
              int shift, uint32_t fan, Ctl* __restrict__ ctl, uint64_t row_base, DomainArgs dom) {
  using E = Elem<BUILD, NARROW>;
  using T = typename E::T;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stage = reinterpret_cast<T*>(smem_raw);
  uint16_t* sdig = reinterpret_cast<uint16_t*>(smem_raw + sizeof(T) * SC_TILE);
  __shared__ uint32_t s_hist[SC_FMAX];
  __shared__ uint32_t s_off[SC_FMAX];
  __shared__ long long s_gdelta[SC_FMAX];
  __shared__ uint32_t s_tpref[SC_FMAX + 1];
  __shared__ uint32_t s_warp[SC_THREADS / 32];
  __shared__ uint32_t s_total;

  const int tid = threadIdx.x;
  uint64_t ntiles;
  if (STAGE == 1) {
    ntiles = (n + SC_TILE - 1) / SC_TILE;
  } else {
    uint32_t t = 0;
    if (tid < (int)in_nparts) {
      uint64_t c = in_counts[tid];
      if (c > in_cap) c = in_cap;
      t = (uint32_t)((c + SC_TILE - 1) / SC_TILE);
    }
    uint32_t total;
    const uint32_t pre = block_excl_scan_512(t, s_warp, total);
    if (tid < (int)in_nparts) s_tpref[tid] = pre;
    if (tid == 0) s_tpref[in_nparts] = total;
    __syncthreads();
    ntiles = total;
  }

  unsigned long long sentinel_local = 0;
  uint32_t kmax = 0;

  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    // ---- locate the tile
    uint64_t in_base;
    uint32_t count, p1 = 0;
    if (STAGE == 1) {
      in_base = tile * SC_TILE;
      count = (uint32_t)((n - in_base) < (uint64_t)SC_TILE ? (n - in_base) : SC_TILE);
    } else {
      uint32_t lo = 0, hi = in_nparts;  // largest p with s_tpref[p] <= tile
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_tpref[mid] <= (uint32_t)tile) lo = mid; else hi = mid;
      }
      p1 = lo;
      uint64_t c = in_counts[p1];
      if (c > in_cap) c = in_cap;
      const uint64_t within = (tile - s_tpref[p1]) * (uint64_t)SC_TILE;
      in_base = (uint64_t)p1 * in_cap + within;
      count = (uint32_t)((c - within) < (uint64_t)SC_TILE ? (c - within) : SC_TILE);
    }
    if (tid < (int)fan) s_hist[tid] = 0;
    __syncthreads();

    // ---- load, convert, rank (shared-memory atomicAdd returns the rank inside (tile, digit))
    T elem[SC_IPT];
    uint32_t dr[SC_IPT];  // digit << 16 | rank   (0xffffffff = dropped row)
#pragma unroll
    for (int i = 0; i < SC_IPT; ++i) {
      const uint32_t e = i * SC_THREADS + tid;
      dr[i] = 0xffffffffu;
      if (e < count) {
        bool ok = true;
        unsigned long long k;
        if constexpr (STAGE == 1) {
          k = ld_stream1(in_keys + in_base + e);
          if constexpr (BUILD) {
            const unsigned long long v = ld_stream1(in_vals + in_base + e);
            if constexpr (NARROW) {
              const unsigned long long packed = (k << 32) | (v & 0xffffffffull);
              if (!((k < dom.klimit) & (v <= dom.vlimit))) { atomicOr(&ctl->flags, dom.badflag); ok = false; }
              else kmax = max(kmax, (uint32_t)k);
              elem[i] = packed;
            } else {
              if (k == EMPTY64) { atomicMin(&ctl->sentinel_row, (unsigned long long)(row_base + in_base + e)); ok = false; }
              elem[i].x = k; elem[i].y = v;
            }
          } else {
            if constexpr (NARROW) {
              if (k >= dom.klimit) ok = false;  // cannot match a packed build side
              elem[i] = (uint32_t)k;
            } else {
              if (k == EMPTY64) { ++sentinel_local; ok = false; }
              elem[i] = k;
            }
          }
        } else {
          elem[i] = in_part[in_base + e];
          k = E::key(elem[i]);
        }
        if (ok) {
          const uint32_t d = scatter_digit(dom.ident ? (uint32_t)k : hash32(k), shift, fan);
          const uint32_t r = atomicAdd(&s_hist[d], 1u);
          dr[i] = (d << 16) | r;
        }
      }
    }
    __syncthreads();

    // ---- per-digit: smem offset (block scan) and global run reservation (one atomic per digit)
    {
      const uint32_t c = tid < (int)fan ? s_hist[tid] : 0u;
      uint32_t total;
      const uint32_t off = block_excl_scan_512(c, s_warp, total);
      if (tid < (int)fan) {
        s_off[tid] = off;
        long long gd = SC_POISON;
        if (c) {
          const uint32_t outp = (STAGE == 1) ? (uint32_t)tid : p1 * fan + (uint32_t)tid;
          const uint32_t g = atomicAdd(out_cursor + outp, c);
          if ((uint64_t)g + c > out_cap) atomicOr(&ctl->flags, CTL_OVERFLOW);
          else gd = (long long)((uint64_t)outp * out_cap + g) - (long long)off;
        }
        s_gdelta[tid] = gd;
      }
      if (tid == 0) s_total = total;
    }
    __syncthreads();

    // ---- reorder the tile in shared memory: partition-contiguous
#pragma unroll
    for (int i = 0; i < SC_IPT; ++i) {
      if (dr[i] != 0xffffffffu) {
        const uint32_t d = dr[i] >> 16;
        const uint32_t pos = s_off[d] + (dr[i] & 0xffffu);
        stage[pos] = elem[i];
        sdig[pos] = (uint16_t)d;
      }
    }
    __syncthreads();

    // ---- write out: consecutive threads -> consecutive addresses inside each partition run
    const uint32_t total = s_total;
    for (uint32_t j = tid; j < total; j += SC_THREADS) {
      const long long gd = s_gdelta[sdig[j]];
      if (gd != SC_POISON) out[gd + (long long)j] = stage[j];
    }
    // next iteration's first __syncthreads (after zeroing s_hist) orders these reads before reuse
  }

  if (BUILD && NARROW && STAGE == 1 && dom.ident) {
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if ((tid & 31) == 0 && kmax) atomicMax(&ctl->max_key, (unsigned long long)kmax);
  }
  if (!BUILD && !NARROW && STAGE == 1) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sentinel_local += __shfl_xor_sync(0xffffffffu, sentinel_local, d);
    if ((tid & 31) == 0 && sentinel_local) atomicAdd(&ctl->sentinel_probes, sentinel_local);
  }
}

template <bool BUILD, bool NARROW, int STAGE>
static void launch_scatter_inst(const ScatterArgs& a, const DeviceInfo& di, cudaStream_t st) {
  using T = typename Elem<BUILD, NARROW>::T;
  auto kern = k_scatter<BUILD, NARROW, STAGE>;
  const size_t smem = (sizeof(T) + sizeof(uint16_t)) * SC_TILE;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  uint64_t max_tiles;
  if (STAGE == 1) max_tiles = (a.n + SC_TILE - 1) / SC_TILE;
  else max_tiles = (a.n_upper + SC_TILE - 1) / SC_TILE + a.in_nparts;
  uint64_t grid = (uint64_t)di.sms * 2;
  if (grid > max_tiles) grid = max_tiles;
  if (grid == 0) return;
  kern<<<(unsigned)grid, SC_THREADS, smem, st>>>(a.in_keys, a.in_vals, a.n, reinterpret_cast<const T*>(a.in_part),
                                                 a.in_counts, a.in_nparts, a.in_cap, reinterpret_cast<T*>(a.out),
                                                 a.out_cursor, a.out_cap, a.shift, a.fan, a.ctl, a.row_base, a.dom);
}

size_t radix_elem_bytes(bool build, bool narrow) {
  return build ? (narrow ? 8 : 16) : (narrow ? 4 : 8);
}

// ================================================================================= scatter, TMA pipelined
// k_scatter2: the production partition kernel.  Everything that touches HBM is asynchronous:
//   in : cp.async.bulk global -> shared into a 2-deep ring of input tiles, completion on an mbarrier;
//        the copy of tile t+2 is issued as soon as tile t has been consumed
//   out: every (tile, partition) run is ONE cp.async.bulk shared -> global from a double-buffered staging
//        area in which the tile has been regrouped by partition.  Bulk stores need 16-byte aligned
//        runs, so a run is padded to a multiple of 16 bytes with all-ones elements ("holes"; an
//        all-ones element can never be a row — see Elem below) and reservations are made in
//        padded units.  Consumers (the next scatter pass, k_join, k_expand) skip holes.
// The threads only ever touch shared memory: LDS row -> hash -> ATOMS rank -> STS staged element.
// Tile = 4096 rows (2048 for 16-byte elements): 8 (4) rows per thread; two CTAs per SM.
constexpr int S2_THREADS = 512;
constexpr int S2_FMAX = 256;  // max fan-out of one pass (digits fit 8 bits)

template <bool BUILD, bool NARROW> struct Hole;
template <> struct Hole<true, true> {
  static __device__ __forceinline__ unsigned long long make() { return EMPTY64; }
  static __device__ __forceinline__ bool is(unsigned long long e) { return (uint32_t)(e >> 32) == 0xFFFFFFFFu; }
};
template <> struct Hole<true, false> {  // 16-byte elements never need padding; the out-of-band key never enters
  static __device__ __forceinline__ ulonglong2 make() { return make_ulonglong2(EMPTY64, EMPTY64); }
  static __device__ __forceinline__ bool is(const ulonglong2& e) { return e.x == EMPTY64; }
};
template <> struct Hole<false, true> {
  static __device__ __forceinline__ uint32_t make() { return 0xFFFFFFFFu; }
  static __device__ __forceinline__ bool is(uint32_t e) { return e == 0xFFFFFFFFu; }
};
template <> struct Hole<false, false> {
  static __device__ __forceinline__ unsigned long long make() { return EMPTY64; }
  static __device__ __forceinline__ bool is(unsigned long long e) { return e == EMPTY64; }
};

template <class T> struct S2Tile { static constexpr int ROWS = sizeof(T) == 16 ? 2048 : 4096; };
constexpr int S2_CHUNKS = 2048 + S2_FMAX;  // 16-byte chunks of the staging buffer (<= 32 KB of rows), padding included

template <bool BUILD, bool NARROW, int STAGE, bool DEST>
__global__ void __launch_bounds__(S2_THREADS, 2)
    k_scatter2(const unsigned long long* __restrict__ in_keys, const unsigned long long* __restrict__ in_vals,
               uint64_t n_tiles1,                                            // stage 1: number of FULL tiles
               const typename Elem<BUILD, NARROW>::T* __restrict__ in_part,  // stage 2
               const uint32_t* __restrict__ in_counts, uint32_t in_nparts, uint64_t in_cap, int merge,
               typename Elem<BUILD, NARROW>::T* __restrict__ out, uint32_t* __restrict__ out_cursor, uint64_t out_cap,
               int shift, uint32_t fan, Ctl* __restrict__ ctl, uint64_t row_base, DomainArgs dom) {
  using E = Elem<BUILD, NARROW>;
  using T = typename E::T;
  using H = Hole<BUILD, NARROW>;
  constexpr int TILE = S2Tile<T>::ROWS;
  constexpr int IPT = TILE / S2_THREADS;
  constexpr int PADN = 16 / (int)sizeof(T) > 1 ? 16 / (int)sizeof(T) : 1;  // rows per 16-byte chunk
  constexpr int IN_ROW = STAGE == 1 ? 8 : (int)sizeof(T);                  // stage 1: the ring holds the KEYS only
  constexpr int RING = TILE * IN_ROW;                                       // bytes per ring stage
  constexpr int POISON = 0x7fffffff;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* ring = smem_raw;
  T* stg = reinterpret_cast<T*>(smem_raw + 2 * RING);  // TILE rows + room for the padding of `fan` runs
  __shared__ uint32_t s_hist[S2_FMAX + 1];      // rows per digit of the current tile; [fan] = bin of the dropped rows
  __shared__ uint32_t s_cnt[S2_FMAX];           // copy of s_hist taken by the scan
  __shared__ uint32_t s_off[S2_FMAX + 1];       // staging offset (rows) of the digit's padded run; [fan] = spare slot
  __shared__ int s_delta[S2_FMAX];              // global chunk index - staging chunk index of the staged tile
  __shared__ uint8_t s_cdig[S2_CHUNKS];         // digit of every 16-byte chunk of the staged tile
  __shared__ uint32_t s_nchunk[2];
  __shared__ uint32_t s_tpref[S2_FMAX + 1];
  __shared__ uint32_t s_warp[S2_THREADS / 32];
  __shared__ __align__(8) uint64_t s_full[2];
  __shared__ unsigned long long s_tbase[2];     // per ring stage: input offset (elements) of the tile in flight
  __shared__ uint32_t s_tcount[2], s_tp1[2];

  const int tid = threadIdx.x, lane = tid & 31;
  uint64_t ntiles;
  if (STAGE == 1) {
    ntiles = n_tiles1;
  } else {
    uint32_t t = 0;
    if (tid < (int)in_nparts) {
      uint64_t c = in_counts[tid];
      if (c > in_cap) c = in_cap;
      t = (uint32_t)((c + TILE - 1) / TILE);
    }
    uint32_t total;
    const uint32_t pre = block_excl_scan_512(t, s_warp, total);
    if (tid < (int)in_nparts) s_tpref[tid] = pre;
    if (tid == 0) s_tpref[in_nparts] = total;
    __syncthreads();
    ntiles = total;
  }
  if (tid <= S2_FMAX) s_hist[tid] = 0;

  // producer (thread 0): locate tile -> (input offset, rows, source partition), publish it for the consumers
  // and bulk-copy the tile into ring stage s
  auto issue = [&](uint64_t tile, int s) {
    uint64_t in_base;
    uint32_t count, p1 = 0;
    if (STAGE == 1) {
      in_base = tile * TILE;
      count = TILE;
    } else {
      uint32_t lo = 0, hi = in_nparts;  // largest p with s_tpref[p] <= tile
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s_tpref[mid] <= (uint32_t)tile) lo = mid; else hi = mid;
      }
      p1 = lo;
      uint64_t c = in_counts[p1];
      if (c > in_cap) c = in_cap;
      const uint64_t within = (tile - s_tpref[p1]) * (uint64_t)TILE;
      in_base = (uint64_t)p1 * in_cap + within;
      count = (uint32_t)((c - within) < (uint64_t)TILE ? (c - within) : TILE);
    }
    s_tbase[s] = in_base;
    s_tcount[s] = count;
    s_tp1[s] = p1;
    unsigned char* dst = ring + s * RING;
    if (STAGE == 1) {
      mbar_expect_tx(&s_full[s], TILE * 8u);
      bulk_g2s(dst, in_keys + in_base, TILE * 8u, &s_full[s]);
    } else {
      const uint32_t bytes = (count * (uint32_t)sizeof(T) + 15u) & ~15u;  // region capacity is a multiple of 16 B
      mbar_expect_tx(&s_full[s], bytes);
      bulk_g2s(dst, in_part + in_base, bytes, &s_full[s]);
    }
  };

  if (tid == 0) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    if (blockIdx.x < ntiles) issue(blockIdx.x, 0);
    if (blockIdx.x + (uint64_t)gridDim.x < ntiles) issue(blockIdx.x + (uint64_t)gridDim.x, 1);
  }

  // step 4: copy the staged tile out, 16 bytes per thread and iteration; the chunks of a run are contiguous
  // both in the staging buffer and in global memory, so a warp writes (a few) contiguous pieces
  // Inside the tile loop warp 0 does not take part: it turns the next histogram into staging offsets meanwhile
  // (step 2), which used to make it the straggler of the barrier that follows.
  auto copy_out = [&](uint32_t nchunk, bool with_warp0) {
    const uint4* src16 = reinterpret_cast<const uint4*>(stg);
    uint4* out16 = reinterpret_cast<uint4*>(out);
    if (!with_warp0 && tid < 32) return;
    const uint32_t first = with_warp0 ? (uint32_t)tid : (uint32_t)tid - 32u;
    const uint32_t step = with_warp0 ? (uint32_t)S2_THREADS : (uint32_t)S2_THREADS - 32u;
    for (uint32_t c = first; c < nchunk; c += step) {
      const int dl = s_delta[s_cdig[c]];
      if (dl != POISON) out16[(long long)dl + (long long)c] = src16[c];
    }
  };

  const uint32_t dpl = (fan + 31u) >> 5;  // digits per lane in the single-warp scan
  unsigned long long sentinel_local = 0;
  uint32_t kmax = 0;  // packed stage-1 build: largest staged key (the dense-domain join sizes its regions by it)
  // digit threads (tid < fan): the run reserved for this thread's digit in the tile staged last
  uint32_t my_g = 0, my_pc = 0, my_off = 0, my_outp = 0;
  uint32_t it = 0;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const int s = it & 1;
    mbar_wait(&s_full[s], (it >> 1) & 1u);  // acquire: also makes s_tbase/s_tcount/s_tp1[s] visible
    const uint64_t in_base = s_tbase[s];
    const uint32_t count = s_tcount[s], p1 = s_tp1[s];

    // ---- 1. rows out of the ring; rank inside (tile, digit) from a shared-memory atomic.  Stage-1 build
    // values are only needed when the row is staged (step 3): they come straight from HBM into registers.
    T elem[IPT];
    uint32_t dr[IPT];
    unsigned long long bval[(STAGE == 1 && BUILD) ? IPT : 1];
    const unsigned char* src = ring + s * RING;
    if constexpr (STAGE == 1 && BUILD) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) bval[i] = ld_stream1(in_vals + in_base + i * S2_THREADS + tid);
    }
    // Branch-free on purpose: a dropped row (hole, past the end, out-of-band key) is ranked into the spare bin
    // s_hist[fan] instead of being skipped, so the 8 rows of a thread are 8 independent instruction streams.
    bool bad = false;                          // packed build: a row that does not fit 32|32
    unsigned long long sent_row = EMPTY64;     // wide build: first row with the out-of-band key
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t e = i * S2_THREADS + tid;
      bool ok = e < count;
      unsigned long long k = 0;
      if constexpr (STAGE == 1) {
        k = reinterpret_cast<const unsigned long long*>(src)[e];
        if constexpr (BUILD) {
          if constexpr (NARROW) {
            const bool fits = k < dom.klimit;  // general: < 2^32 - 1; dense domain: < the optimistic key bound
            bad |= ok & !fits;
            ok &= fits;
            kmax = max(kmax, ok ? (uint32_t)k : 0u);
            elem[i] = k << 32;
          } else {
            const bool oob = k == EMPTY64;
            sent_row = (ok & oob & (sent_row == EMPTY64)) ? (unsigned long long)(row_base + in_base + e) : sent_row;
            ok &= !oob;
            elem[i].x = k;
          }
        } else {
          if constexpr (NARROW) {
            ok &= k < dom.klimit;  // else: cannot match a packed build side
            elem[i] = (uint32_t)k;
          } else {
            const bool oob = k == EMPTY64;
            sentinel_local += (ok & oob) ? 1ull : 0ull;
            ok &= !oob;
            elem[i] = k;
          }
        }
      } else {
        elem[i] = reinterpret_cast<const T*>(src)[e];
        ok &= !H::is(elem[i]);
        k = E::key(elem[i]);
      }
      const uint32_t h = (STAGE == 1 && NARROW && !DEST && dom.ident) ? (uint32_t)k : hash32(k);
      const uint32_t d = ok ? (DEST ? ((h & 0xffffu) * fan) >> 16 : (h >> shift) & (fan - 1)) : fan;
      const uint32_t r = atomicAdd(&s_hist[d], 1u);
      dr[i] = ok ? ((d << 16) | r) : (fan << 16);  // dropped rows are staged into one spare slot
    }
    if constexpr (STAGE == 1 && BUILD && NARROW) {
      if (bad) atomicOr(&ctl->flags, dom.badflag);  // the attempt is abandoned by the host
    }
    if constexpr (STAGE == 1 && BUILD && !NARROW) {
      if (sent_row != EMPTY64) atomicMin(&ctl->sentinel_row, sent_row);
    }
    // the reservation made for the previous tile has long returned: publish where its runs go
    if (it > 0 && tid < (int)fan) {
      int dl = POISON;
      if (my_pc) {
        if ((uint64_t)my_g + my_pc > out_cap) atomicOr(&ctl->flags, CTL_OVERFLOW);
        else dl = (int)((long long)(((uint64_t)my_outp * out_cap + my_g) / PADN) - (long long)(my_off / PADN));
      }
      s_delta[tid] = dl;
    }
    __syncthreads();  // ring stage s consumed, histogram complete, previous tile staged and its s_delta published

    // ---- 4 (of the PREVIOUS tile): copy it out.  Deferred to here so that the latency of the reserving
    // atomicAdd (issued in step 3) is covered by step 1 of this tile.  Thread 0 refills the ring first.
    if (tid == 0) {
      const uint64_t nxt = tile + 2ull * gridDim.x;
      if (nxt < ntiles) issue(nxt, s);
    }
    // ---- 2. warp 0 turns the histogram into padded staging offsets while the other warps copy out
    if (tid < 32) {
      uint32_t sum = 0;
      for (uint32_t j = 0; j < dpl; ++j) {
        const uint32_t d = lane * dpl + j;
        if (d < fan) sum += (s_hist[d] + PADN - 1) / PADN * PADN;
      }
      uint32_t incl = sum;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
      }
      uint32_t run = incl - sum;
      for (uint32_t j = 0; j < dpl; ++j) {
        const uint32_t d = lane * dpl + j;
        if (d < fan) {
          const uint32_t c = s_hist[d];
          s_hist[d] = 0;
          s_cnt[d] = c;
          s_off[d] = run;
          run += (c + PADN - 1) / PADN * PADN;
        }
      }
      if (lane == 31) {
        s_nchunk[s] = incl / PADN;
        s_hist[fan] = 0;
        s_off[fan] = ((uint32_t)TILE + fan * (PADN - 1) + 7u) & ~7u;  // first row past the staged tile
      }
    }
    if (it > 0) copy_out(s_nchunk[s ^ 1], false);
    __syncthreads();

    // ---- 3. regroup the tile by digit in the staging buffer; the digit threads reserve the global runs
    bool bad3 = false;
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if constexpr (STAGE == 1 && BUILD) {
        if constexpr (NARROW) {
          bad3 |= (bval[i] > dom.vlimit) & ((dr[i] >> 16) != fan);
          elem[i] |= bval[i] & 0xffffffffull;
        } else {
          elem[i].y = bval[i];
        }
      }
      const uint32_t pos = s_off[dr[i] >> 16] + (dr[i] & 0xffffu);
      stg[pos] = elem[i];
      s_cdig[pos / PADN] = (uint8_t)(dr[i] >> 16);  // digit of the 16-byte chunk (every row of the chunk writes the same
                                                     // byte; a dropped row writes past the staged chunks)
    }
    if constexpr (STAGE == 1 && BUILD && NARROW) {
      if (bad3) atomicOr(&ctl->flags, dom.badflag);
    }
    if (tid < (int)fan) {
      const uint32_t my_c = s_cnt[tid];
      my_pc = (my_c + PADN - 1) / PADN * PADN;
      my_off = s_off[tid];
      if (my_c) {
        my_outp = (STAGE == 1 || merge) ? (uint32_t)tid : p1 * fan + (uint32_t)tid;
        my_g = atomicAdd(out_cursor + my_outp, my_pc);  // consumed after step 1 of the next tile
        for (uint32_t j = my_c; j < my_pc; ++j) stg[my_off + j] = H::make();  // holes pad the run to 16 bytes
      }
    }
  }
  // drain: the last tile staged by this CTA
  if (it > 0) {
    if (tid < (int)fan) {
      int dl = POISON;
      if (my_pc) {
        if ((uint64_t)my_g + my_pc > out_cap) atomicOr(&ctl->flags, CTL_OVERFLOW);
        else dl = (int)((long long)(((uint64_t)my_outp * out_cap + my_g) / PADN) - (long long)(my_off / PADN));
      }
      s_delta[tid] = dl;
    }
    __syncthreads();
    copy_out(s_nchunk[(it - 1) & 1], true);
  }

  if (BUILD && NARROW && STAGE == 1 && dom.ident) {
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0 && kmax) atomicMax(&ctl->max_key, (unsigned long long)kmax);
  }
  if (!BUILD && !NARROW && STAGE == 1) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sentinel_local += __shfl_xor_sync(0xffffffffu, sentinel_local, d);
    if ((tid & 31) == 0 && sentinel_local) atomicAdd(&ctl->sentinel_probes, sentinel_local);
  }
}

template <bool BUILD, bool NARROW, int STAGE, bool DEST>
static void launch_scatter2_inst(const ScatterArgs& a, const DeviceInfo& di, cudaStream_t st) {
  using T = typename Elem<BUILD, NARROW>::T;
  constexpr int TILE = S2Tile<T>::ROWS;
  constexpr int IN_ROW = STAGE == 1 ? 8 : (int)sizeof(T);
  auto kern = k_scatter2<BUILD, NARROW, STAGE, DEST>;
  constexpr uint32_t PADN = 16 / sizeof(T) > 1 ? 16 / sizeof(T) : 1;
  const uint32_t stg_elems = ((uint32_t)TILE + a.fan * (PADN - 1) + 7u) & ~7u;  // every run may carry PADN-1 holes
  const size_t smem = 2 * (size_t)TILE * IN_ROW + ((size_t)stg_elems + 1) * sizeof(T);  // + the spare slot of dropped rows
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  uint64_t max_tiles, n_tiles1 = 0;
  if (STAGE == 1) max_tiles = n_tiles1 = a.n / TILE;
  else max_tiles = (a.n_upper + TILE - 1) / TILE + a.in_nparts;
  uint64_t grid = (uint64_t)di.sms * 2;
  if (grid > max_tiles) grid = max_tiles;
  if (grid == 0) return;
  kern<<<(unsigned)grid, S2_THREADS, smem, st>>>(a.in_keys, a.in_vals, n_tiles1, reinterpret_cast<const T*>(a.in_part),
                                                 a.in_counts, a.in_nparts, a.in_cap, a.merge ? 1 : 0,
                                                 reinterpret_cast<T*>(a.out), a.out_cursor, a.out_cap, a.shift, a.fan, a.ctl,
                                                 a.row_base, a.dom);
}

// rows per tile of the pipelined scatter for one element format
uint32_t scatter_tile_rows(bool build, bool narrow) { return radix_elem_bytes(build, narrow) == 16 ? 2048u : 4096u; }
uint32_t scatter_pad_rows(bool build, bool narrow) {
  const size_t e = radix_elem_bytes(build, narrow);
  return e >= 16 ? 0u : (uint32_t)(16 / e - 1);
}

// Stage 1: full tiles of 16-byte aligned inputs go through k_scatter2, the ragged tail (and inputs that
// are only 8-byte aligned) through k_scatter.  Stage 2 (partition -> partition) is always k_scatter2.
void launch_scatter(bool build, bool narrow, int stage, const ScatterArgs& a, const DeviceInfo& di, cudaStream_t st,
                    int* launches) {
#define FJ_SC2(B, N, S)                                               \
  do {                                                                \
    if (S == 1 && a.shift < 0) launch_scatter2_inst<B, N, 1, true>(a, di, st); \
    else launch_scatter2_inst<B, N, S, false>(a, di, st);             \
  } while (0)
#define FJ_DISPATCH(M, S)                                              \
  do {                                                                 \
    if (build) { if (narrow) M(true, true, S); else M(true, false, S); } \
    else { if (narrow) M(false, true, S); else M(false, false, S); }   \
  } while (0)
  if (stage == 2) {
    FJ_DISPATCH(FJ_SC2, 2);
    ++*launches;
    return;
  }
  const uint32_t tile = scatter_tile_rows(build, narrow);
  const bool aligned = (reinterpret_cast<uintptr_t>(a.in_keys) & 15u) == 0 &&
                       (!build || (reinterpret_cast<uintptr_t>(a.in_vals) & 15u) == 0);
  const uint64_t full = aligned ? a.n / tile * tile : 0;
  if (full) {
    FJ_DISPATCH(FJ_SC2, 1);
    ++*launches;
  }
  if (full < a.n) {
    ScatterArgs t = a;
    t.in_keys = a.in_keys + full;
    t.in_vals = a.in_vals ? a.in_vals + full : nullptr;
    t.n = a.n - full;
    t.row_base = a.row_base + full;
#define FJ_SC1(B, N, S) launch_scatter_inst<B, N, 1>(t, di, st)
    FJ_DISPATCH(FJ_SC1, 1);
#undef FJ_SC1
    ++*launches;
  }
#undef FJ_DISPATCH
#undef FJ_SC2
}


// ================================================================================= partition join
// One CTA per (partition, probe chunk).  Shared memory: the partition's build tuples (TMA bulk
// load) + an index table of 32-bit slots: fingerprint16 << 16 | (tuple index + 1), 0 = empty.
// Slots are claimed with 32-bit shared-memory atomicCAS (64-bit CAS in shared memory is ~15x
// slower on B200, tools/ubench.cu).
constexpr int JN_THREADS = 512;
constexpr int JN_WARPS = JN_THREADS / 32;

template <bool NARROW> struct ProbeVec;
template <> struct ProbeVec<true> { static constexpr int K = 4; };   // uint4 = 4 x key32
template <> struct ProbeVec<false> { static constexpr int K = 2; };  // 2 x key64

// BLOOM: the per-partition table carries a register-blocked Bloom filter in shared memory (one 32-bit word per key,
// fj_common.cuh), filled during the build and checked before the table is probed — the counterpart of the
// FlashHashTable<true> the reference builds per partition for hash_join_radix_bloom / hash_join_count_radix_bloom
// (hash_join.cpp:344, :518, :627, :636).  As there, it changes the time only, never the result.
template <bool NARROW, bool MAT, bool BLOOM>
__global__ void __launch_bounds__(JN_THREADS, 2)
    k_join(const typename Elem<true, NARROW>::T* __restrict__ build, const uint32_t* __restrict__ bcnt, uint64_t cap_b,
           const typename Elem<false, NARROW>::T* __restrict__ probe, const uint32_t* __restrict__ pcnt, uint64_t cap_p,
           uint32_t smax, uint32_t tcap, uint32_t bwords, uint32_t chunk, uint32_t max_chunks, Ctl* __restrict__ ctl,
           unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_vals) {
  using TB = typename Elem<true, NARROW>::T;
  using TP = typename Elem<false, NARROW>::T;
  constexpr int K = ProbeVec<NARROW>::K;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TB* tuples = reinterpret_cast<TB*>(smem_raw);
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw + (size_t)smax * sizeof(TB));
  uint32_t* bf = table + tcap;  // BLOOM: bwords filter words
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_wcnt[MAT ? JN_WARPS * K : 1];
  __shared__ unsigned long long s_base;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t p = blockIdx.x / max_chunks;
  const uint32_t ck = blockIdx.x - p * max_chunks;
  uint64_t npp = pcnt[p];
  if (npp > cap_p) npp = cap_p;
  const uint64_t pstart = (uint64_t)ck * chunk;
  if (pstart >= npp) return;
  uint32_t nbp = bcnt[p];
  if (nbp > cap_b) nbp = (uint32_t)cap_b;  // region overflowed: CTL_OVERFLOW already raised by the scatter
  if (nbp == 0) return;
  if (nbp > smax) {
    if (tid == 0) atomicOr(&ctl->flags, CTL_OVERFLOW);
    return;
  }

  // ---- stage the build tuples (TMA bulk copy) while the table is cleared
  if (tid == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)(((size_t)nbp * sizeof(TB) + 15) & ~(size_t)15);
    mbar_expect_tx(&s_bar, bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(build + (uint64_t)p * cap_b);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t nn = bytes - off < 32768u ? bytes - off : 32768u;
      bulk_g2s(smem_raw + off, src + off, nn, &s_bar);
    }
  }
  for (uint32_t i = tid; i < tcap + (BLOOM ? bwords : 0u); i += JN_THREADS) table[i] = 0u;  // table and filter are adjacent
  __syncthreads();
  mbar_wait(&s_bar, 0);

  // ---- build: claim a slot per tuple with 32-bit CAS
  for (uint32_t i = tid; i < nbp; i += JN_THREADS) {
    if (Hole<true, NARROW>::is(tuples[i])) continue;  // padding written by k_scatter2
    const unsigned long long key = Elem<true, NARROW>::key(tuples[i]);
    if constexpr (BLOOM) {
      const uint32_t bh = bloom_hash(key);
      atomicOr(bf + bloom_word(bh, bwords), bloom_mask(bh));
    }
    const uint32_t g = hash32(key) * 0x9E3779B1u;
    const uint32_t fp = g & 0xffffu;
    const uint32_t mine = (fp << 16) | (i + 1u);
    uint32_t s = __umulhi(g, tcap);
    for (uint32_t it = 0; it < tcap; ++it) {
      const uint32_t old = atomicCAS(table + s, 0u, mine);
      if (old == 0u) break;
      if ((old >> 16) == fp) {
        const uint32_t j = (old & 0xffffu) - 1u;
        if (Elem<true, NARROW>::key(tuples[j]) == key) { atomicOr(&ctl->flags, CTL_DUP); break; }
      }
      if (++s == tcap) s = 0;
    }
  }
  __syncthreads();

  // ---- probe this chunk of the partition's probe keys
  const TP* pk = probe + (uint64_t)p * cap_p;
  uint64_t pend = pstart + chunk;
  if (pend > npp) pend = npp;
  unsigned long long local_count = 0;
  for (uint64_t tb = pstart; tb < pend; tb += (uint64_t)JN_THREADS * K) {
    const uint64_t e0 = tb + (uint64_t)tid * K;
    unsigned long long key[K];
    bool valid[K];
    if (e0 + K <= pend) {
      if constexpr (NARROW) {
        const uint4 v = *reinterpret_cast<const uint4*>(pk + e0);  // region base and chunk are 16 B aligned
        key[0] = v.x; key[1] = v.y; key[2] = v.z; key[3] = v.w;
      } else {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(pk + e0);
        key[0] = v.x; key[1] = v.y;
      }
#pragma unroll
      for (int q = 0; q < K; ++q) valid[q] = true;
    } else {
#pragma unroll
      for (int q = 0; q < K; ++q) {
        valid[q] = e0 + q < pend;
        key[q] = valid[q] ? (unsigned long long)pk[e0 + q] : 0ull;
      }
    }
    unsigned long long val[K];
    uint32_t hitmask = 0;
#pragma unroll
    for (int q = 0; q < K; ++q) {
      if (!valid[q]) continue;
      if constexpr (BLOOM) {
        const uint32_t bh = bloom_hash(key[q]);
        const uint32_t m = bloom_mask(bh);
        if ((bf[bloom_word(bh, bwords)] & m) != m) continue;  // definitely not in this partition's build side
      }
      const uint32_t g = hash32(key[q]) * 0x9E3779B1u;
      const uint32_t fp = g & 0xffffu;
      uint32_t s = __umulhi(g, tcap);
      for (uint32_t it = 0; it < tcap; ++it) {
        const uint32_t slot = table[s];
        if (slot == 0u) break;
        if ((slot >> 16) == fp) {
          const TB t = tuples[(slot & 0xffffu) - 1u];
          if (Elem<true, NARROW>::key(t) == key[q]) {
            if constexpr (NARROW) val[q] = t & 0xffffffffull;
            else val[q] = t.y;
            hitmask |= 1u << q;
            break;
          }
        }
        if (++s == tcap) s = 0;
      }
    }
    if (!MAT) {
      local_count += __popc(hitmask);
    } else {
      uint32_t rank[K];
#pragma unroll
      for (int q = 0; q < K; ++q) {
        const unsigned bal = __ballot_sync(0xffffffffu, (hitmask >> q) & 1u);
        rank[q] = __popc(bal & lanemask_lt());
        if (lane == 0) s_wcnt[warp * K + q] = __popc(bal);
      }
      __syncthreads();
      if (warp == 0) {
        constexpr int PER = (JN_WARPS * K + 31) / 32;
        uint32_t c[PER];
        uint32_t sum = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int e = lane * PER + u;
          c[u] = e < JN_WARPS * K ? s_wcnt[e] : 0u;
          sum += c[u];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int e = lane * PER + u;
          if (e < JN_WARPS * K) s_wcnt[e] = run;
          run += c[u];
        }
        if (lane == 31) {
          s_base = incl ? atomicAdd(&ctl->out_cursor, (unsigned long long)incl) : 0ull;
          local_count += incl;
        }
      }
      __syncthreads();
      const unsigned long long base = s_base;
#pragma unroll
      for (int q = 0; q < K; ++q) {
        if ((hitmask >> q) & 1u) {
          const unsigned long long pos = base + s_wcnt[warp * K + q] + rank[q];
          st_stream(out_keys + pos, key[q]);
          st_stream(out_vals + pos, val[q]);
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
  if (lane == 0 && local_count) atomicAdd(&ctl->match_count, local_count);
}

void launch_join(bool narrow, bool mat, const JoinArgs& a, cudaStream_t st, int* launches) {
  const bool bloom = a.bloom_words != 0;
  const size_t smem = (size_t)a.smax * (narrow ? 8 : 16) + (size_t)a.tcap * 4 + (size_t)a.bloom_words * 4;
  const uint64_t grid = (uint64_t)a.nparts * a.max_chunks;
  if (grid == 0) return;
#define FJ_JN(N, M, B)                                                                                        \
  do {                                                                                                        \
    auto kern = k_join<N, M, B>;                                                                              \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
    kern<<<(unsigned)grid, JN_THREADS, smem, st>>>(                                                           \
        reinterpret_cast<const typename Elem<true, N>::T*>(a.build), a.bcnt, a.cap_b,                         \
        reinterpret_cast<const typename Elem<false, N>::T*>(a.probe), a.pcnt, a.cap_p, a.smax, a.tcap, a.bloom_words, \
        a.chunk, a.max_chunks, a.ctl, a.out_keys, a.out_vals);                                                \
  } while (0)
#define FJ_JN2(N, M) do { if (bloom) FJ_JN(N, M, true); else FJ_JN(N, M, false); } while (0)
  if (narrow) { if (mat) FJ_JN2(true, true); else FJ_JN2(true, false); }
  else { if (mat) FJ_JN2(false, true); else FJ_JN2(false, false); }
#undef FJ_JN2
#undef FJ_JN
  ++*launches;
}

// ================================================================================= partition join, pipelined
// k_join3 (packed 32|32 rows, >= 2^14 partitions): collision-free join in shared memory.
//
// For a packed row the key is a 32-bit value and hash32 is a BIJECTION on 32-bit values; the radix passes
// consumed the top `bits` hash bits, so inside partition p a key is identified by the remaining
// rbits = 32 - bits hash bits ("rem").  The partition's build side is therefore a SET of rems in a universe
// of 2^rbits (<= 2^18) values: a bitmap (<= 32 KB) answers membership exactly — no key comparison, no
// collision chain, no divergent loop — and a rank directory (16-bit running popcount per bitmap word) maps a
// member to the slot of its value in a dense value array:  value = vals[prefix[w] + popc(word & below(bit))].
// Build = atomicOr (a bit that was already set is a duplicate build key -> CTL_DUP), a block scan of the word
// popcounts, and one value store per row; probe = three shared-memory loads, branch-free.
//
// Persistent CTAs walk the (partition, probe chunk) items; all HBM input is prefetched one item ahead
// with TMA bulk copies (build tuples into a staging area, probe keys into a double buffer).  Every warp
// reserves the output range of its own matches with one global atomic and writes it as contiguous runs.
constexpr int J3_THREADS = 512;
constexpr int J3_WARPS = J3_THREADS / 32;
constexpr int J3_IPT = 8;
constexpr int J3_PCH = J3_THREADS * J3_IPT;  // probe rows per chunk
constexpr int J3_BPT = 8;                    // build rows per thread kept in registers (smax <= J3_THREADS * J3_BPT)

template <bool MAT>
__global__ void __launch_bounds__(J3_THREADS, 2)
    k_join3(const unsigned long long* __restrict__ build, const uint32_t* __restrict__ bcnt, uint64_t cap_b,
            const uint32_t* __restrict__ probe, const uint32_t* __restrict__ pcnt, uint64_t cap_p, uint32_t smax,
            int rbits, uint32_t max_chunks, uint64_t nitems, Ctl* __restrict__ ctl,
            unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_vals) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t nwords = 1u << (rbits - 5);
  unsigned long long* staging = reinterpret_cast<unsigned long long*>(smem_raw);                 // smax tuples
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(smem_raw + (size_t)smax * 8);                    // nwords
  uint16_t* prefix = reinterpret_cast<uint16_t*>(smem_raw + (size_t)smax * 8 + (size_t)nwords * 4);  // nwords
  uint32_t* vals = reinterpret_cast<uint32_t*>(smem_raw + (size_t)smax * 8 + (size_t)nwords * 6);    // smax
  uint32_t* pbuf = vals + smax;                                                                   // 2 x J3_PCH
  __shared__ __align__(8) uint64_t s_bar_t, s_bar_p[2];
  __shared__ uint32_t s_nb[2], s_np[2];
  __shared__ uint32_t s_wsum[J3_WARPS];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rmask = (1u << rbits) - 1u;

  auto item_meta = [&](uint64_t idx, uint32_t& p, uint32_t& nbp, uint32_t& npc, uint64_t& pstart) {
    p = (uint32_t)(idx / max_chunks);
    const uint32_t c = (uint32_t)(idx - (uint64_t)p * max_chunks);
    uint64_t nb = bcnt[p];
    if (nb > cap_b) nb = cap_b;  // region overflowed: CTL_OVERFLOW already raised by the scatter
    uint64_t np = pcnt[p];
    if (np > cap_p) np = cap_p;
    pstart = (uint64_t)c * J3_PCH;
    npc = pstart < np ? (uint32_t)((np - pstart) < (uint64_t)J3_PCH ? (np - pstart) : J3_PCH) : 0u;
    nbp = (uint32_t)nb;
    if (nbp > smax) {
      atomicOr(&ctl->flags, CTL_OVERFLOW);
      nbp = 0;
    }
    if (nbp == 0 || npc == 0) nbp = npc = 0;
  };
  auto bulk_in = [&](void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    mbar_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
      bulk_g2s(static_cast<unsigned char*>(dst) + off, static_cast<const unsigned char*>(src) + off, n, bar);
    }
  };
  auto issue_tuples = [&](uint64_t idx, uint32_t k) {  // thread 0
    uint32_t p, nbp, npc;
    uint64_t pstart;
    item_meta(idx, p, nbp, npc, pstart);
    s_nb[k & 1] = nbp;
    bulk_in(staging, build + (uint64_t)p * cap_b, (nbp * 8u + 15u) & ~15u, &s_bar_t);
  };
  auto issue_probe = [&](uint64_t idx, uint32_t k) {  // thread 0
    uint32_t p, nbp, npc;
    uint64_t pstart;
    item_meta(idx, p, nbp, npc, pstart);
    s_np[k & 1] = npc;
    bulk_in(pbuf + (size_t)(k & 1) * J3_PCH, probe + (uint64_t)p * cap_p + pstart, (npc * 4u + 15u) & ~15u, &s_bar_p[k & 1]);
  };

  if (tid == 0) {
    mbar_init(&s_bar_t, 1);
    mbar_init(&s_bar_p[0], 1);
    mbar_init(&s_bar_p[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0 && blockIdx.x < nitems) {
    issue_tuples(blockIdx.x, 0);
    issue_probe(blockIdx.x, 0);
  }
  __syncthreads();

  unsigned long long local_count = 0;
  uint32_t k = 0;
  for (uint64_t idx = blockIdx.x; idx < nitems; idx += gridDim.x, ++k) {
    const uint64_t nxt = idx + gridDim.x;
    if (tid == 0 && nxt < nitems) issue_probe(nxt, k + 1);  // that buffer was released by the barrier ending item k-1
    const uint32_t nbp = s_nb[k & 1], npc = s_np[k & 1];

    // ---- A. clear the bitmap while the tuples land
    if (nbp) {
      uint4* bm4 = reinterpret_cast<uint4*>(bitmap);
      for (uint32_t i = tid; i < nwords / 4; i += J3_THREADS) bm4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(&s_bar_t, k & 1u);
    __syncthreads();

    if (nbp) {  // block-uniform
      // ---- B. set the member bits; keep (rem, value) of this thread's rows in registers.  Branch-free: a
      // row that is padding or past the end ORs a zero into word 0, so the rows of a thread are independent
      // instruction streams instead of a chain of divergent regions.
      uint32_t rem[J3_BPT], bval[J3_BPT];
      bool dup = false;
#pragma unroll
      for (int j = 0; j < J3_BPT; ++j) {
        rem[j] = 0xFFFFFFFFu;
        bval[j] = 0;
        if (j * J3_THREADS < (int)nbp) {  // block-uniform: whole slots past the partition cost nothing
          const uint32_t i = j * J3_THREADS + tid;
          const unsigned long long t = staging[i < nbp ? i : 0];
          const uint32_t key = (uint32_t)(t >> 32);
          const bool ok = (i < nbp) & (key != 0xFFFFFFFFu);  // 0xFFFFFFFF: padding written by k_scatter2
          const uint32_t rm = hash32(key) & rmask;
          rem[j] = ok ? rm : 0xFFFFFFFFu;
          bval[j] = (uint32_t)t;
          const uint32_t bit = ok ? (1u << (rm & 31u)) : 0u;
          const uint32_t old = atomicOr(bitmap + (ok ? (rm >> 5) : (uint32_t)lane), bit);  // dropped rows OR a zero
          dup |= (old & bit) != 0u;  // same key twice (hash32 is a bijection on 32-bit keys)
        }
      }
      if (dup) atomicOr(&ctl->flags, CTL_DUP);
      __syncthreads();
      // ---- C. rank directory: running popcount before every bitmap word.  Thread t owns the wpt consecutive
      // words [t * wpt, (t + 1) * wpt); with wpt == 8 (2^17 hash values per partition, the 1e8-row case) they
      // are read as two 128-bit loads and the eight 16-bit prefixes leave as one 128-bit store.
      {
        const uint32_t wpt = nwords / J3_THREADS;  // words per thread: 1 .. 16 (nwords is a power of two >= 512)
        uint32_t pc[8];
        uint32_t sum = 0;
        if (wpt == 8) {
          const uint4 a = reinterpret_cast<const uint4*>(bitmap)[tid * 2], b = reinterpret_cast<const uint4*>(bitmap)[tid * 2 + 1];
          pc[0] = __popc(a.x); pc[1] = __popc(a.y); pc[2] = __popc(a.z); pc[3] = __popc(a.w);
          pc[4] = __popc(b.x); pc[5] = __popc(b.y); pc[6] = __popc(b.z); pc[7] = __popc(b.w);
#pragma unroll
          for (int w = 0; w < 8; ++w) sum += pc[w];
        } else {
          for (uint32_t w = 0; w < wpt; ++w) sum += __popc(bitmap[tid * wpt + w]);
        }
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        // every warp sums the totals of the warps before it itself (one REDUX) instead of waiting for a
        // second barrier behind a single scanning warp
        static_assert(J3_WARPS <= 32, "one lane per warp total");
        const uint32_t wpre = __reduce_add_sync(0xffffffffu, lane < warp ? s_wsum[lane] : 0u);
        uint32_t run = wpre + incl - sum;
        if (wpt == 8) {
          uint32_t q[8];
#pragma unroll
          for (int w = 0; w < 8; ++w) { q[w] = run; run += pc[w]; }
          reinterpret_cast<uint4*>(prefix)[tid] =
              make_uint4(q[0] | (q[1] << 16), q[2] | (q[3] << 16), q[4] | (q[5] << 16), q[6] | (q[7] << 16));
        } else {
          for (uint32_t w = 0; w < wpt; ++w) {
            prefix[tid * wpt + w] = (uint16_t)run;
            run += __popc(bitmap[tid * wpt + w]);
          }
        }
      }
      __syncthreads();
      // ---- D. place the values by rank
#pragma unroll
      for (int j = 0; j < J3_BPT; ++j) {
        if (j * J3_THREADS < (int)nbp) {
          const bool ok = rem[j] != 0xFFFFFFFFu;
          const uint32_t w = ok ? (rem[j] >> 5) : (uint32_t)lane;
          const uint32_t r = prefix[w] + __popc(bitmap[w] & ((1u << (rem[j] & 31u)) - 1u));
          if (ok) vals[r] = bval[j];
        }
      }
    }
    __syncthreads();
    if (tid == 0 && nxt < nitems) issue_tuples(nxt, k + 1);  // staging is free again

    // ---- probe: membership bit, then the value by rank.  No loops, no key comparison.
    mbar_wait(&s_bar_p[k & 1], (k >> 1) & 1u);
    const uint32_t* pk = pbuf + (size_t)(k & 1) * J3_PCH;
    uint32_t key[J3_IPT], val[J3_IPT];
    uint32_t hitmask = 0;
    if (npc) {
#pragma unroll
      for (int i = 0; i < J3_IPT; ++i) {
        const uint32_t r = i * J3_THREADS + tid;
        key[i] = r < npc ? pk[r] : 0xFFFFFFFFu;
      }
#pragma unroll
      for (int i = 0; i < J3_IPT; ++i) {
        val[i] = 0;
        if (i * J3_THREADS >= (int)npc) continue;  // block-uniform
        const uint32_t rm = hash32(key[i]) & rmask;
        const uint32_t w = rm >> 5, sh = rm & 31u;
        const uint32_t word = bitmap[w];
        const bool hit = (key[i] != 0xFFFFFFFFu) & ((word >> sh) & 1u);  // 0xFFFFFFFF: hole / past the end
        // unconditional: for a miss the rank is still <= the partition's row count, i.e. inside vals[0..smax]
        val[i] = MAT ? vals[prefix[w] + __popc(word & ((1u << sh) - 1u))] : 0u;
        hitmask |= hit ? (1u << i) : 0u;
      }
    }
    if (!MAT) {
      local_count += __popc(hitmask);
    } else if (npc) {  // block-uniform
      // Every warp reserves the output range of its own matches (<= 256 pairs: 2 KB per column, written as 8
      // contiguous runs) with one global atomic: no block-wide scan and no barrier between probing and storing.
      uint32_t off[J3_IPT];
      uint32_t wtot = 0;
#pragma unroll
      for (int i = 0; i < J3_IPT; ++i) {
        const unsigned bal = __ballot_sync(0xffffffffu, (hitmask >> i) & 1u);
        off[i] = wtot + __popc(bal & lanemask_lt());
        wtot += __popc(bal);
      }
      unsigned long long base = 0;
      if (lane == 0 && wtot) {
        base = atomicAdd(&ctl->out_cursor, (unsigned long long)wtot);
        local_count += wtot;
      }
      base = __shfl_sync(0xffffffffu, base, 0);
      unsigned long long* ok = out_keys + base;
      unsigned long long* ov = out_vals + base;
#pragma unroll
      for (int i = 0; i < J3_IPT; ++i) {
        if ((hitmask >> i) & 1u) {
          st_stream(ok + off[i], (unsigned long long)key[i]);
          st_stream(ov + off[i], (unsigned long long)val[i]);
        }
      }
    }
    __syncthreads();  // pbuf[k & 1], the bitmap and vals are reused by the next items
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
  if (lane == 0 && local_count) atomicAdd(&ctl->match_count, local_count);
}

size_t join3_smem_bytes(uint32_t smax, int rbits) {
  return (size_t)smax * 12 + (size_t)(1u << (rbits - 5)) * 6 + 2 * (size_t)J3_PCH * 4;
}
uint32_t join3_probe_chunk() { return J3_PCH; }
uint32_t join3_max_build_rows() { return J3_THREADS * J3_BPT; }
int join3_min_rbits() { return 14; }  // bitmap words >= J3_THREADS

void launch_join3(bool mat, const JoinArgs& a, int rbits, const DeviceInfo& di, cudaStream_t st, int* launches) {
  const size_t smem = join3_smem_bytes(a.smax, rbits);
  const uint64_t nitems = (uint64_t)a.nparts * a.max_chunks;
  if (nitems == 0) return;
  uint64_t grid = (uint64_t)di.sms * 2;
  if (grid > nitems) grid = nitems;
#define FJ_J3(M)                                                                                              \
  do {                                                                                                        \
    auto kern = k_join3<M>;                                                                                   \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
    kern<<<(unsigned)grid, J3_THREADS, smem, st>>>(reinterpret_cast<const unsigned long long*>(a.build), a.bcnt, a.cap_b, \
                                                   reinterpret_cast<const uint32_t*>(a.probe), a.pcnt, a.cap_p, a.smax, \
                                                   rbits, a.max_chunks, nitems, a.ctl, a.out_keys, a.out_vals); \
  } while (0)
  if (mat) FJ_J3(true); else FJ_J3(false);
#undef FJ_J3
  ++*launches;
}

// ================================================================================= dense key domain: direct-address join
// k_djoin (packed rows whose keys lie in a dense domain [0, klimit)): ONE scatter pass
// by the low 8 key bits, then no hashing at all — partition p owns a direct-address region
// region_p[key >> 8] = value + 1 (0 = absent) of 4-byte slots in HBM that is zeroed, filled and probed while
// it is L2 RESIDENT: the kernel walks the 256 partitions in groups sized so that three groups of regions
// (one being zeroed, one being filled, one being probed) fit the L2 budget.  This replaces the second
// scatter pass of both sides and the shared-memory partition join (k_scatter2 x2 + k_join3) by one
// random 4-byte L2 store per build row and one random 4-byte L2 load per probe row.
//
// Persistent CTAs (all co-resident) pull work items in a fixed global order from an atomic ticket (one
// dispatcher warp per CTA resolves tickets and publishes items to a ring; 15 worker warps consume them
// independently, without block barriers):
//   step s:  Z(group s) zero regions | B(group s-1) store build rows | P(group s-2) probe | C(group s-2) count
// B(p, .) waits until all Z(p, .) are done, P/C(p, .) until all B(p, .) are done (per-partition completion
// counters; a waiter only ever waits for items with smaller tickets, which are running or finished, so the
// spin cannot deadlock).  Duplicate build keys are found without atomics: C counts the non-empty slots and
// the host compares with the number of rows stored (fewer slots than rows <=> some key was stored twice ->
// CTL_DUP, the exact keep-first path runs instead).  The count-only variant skips C (a duplicate does not
// change the key set).
constexpr int DJ_THREADS = 512;
constexpr int DJ_WARPS = DJ_THREADS / 32;
constexpr int DJ_F = 256;                 // partitions = low 8 key bits
constexpr int DJ_IPT = 8;
constexpr int DJ_WT = DJ_THREADS - 32;    // worker threads (warp 0 dispatches)
constexpr int DJ_ROWS = DJ_WT * DJ_IPT;   // rows per build / probe item
constexpr int DJ_RING_MAX = 16;           // published-item ring: `ring` slots in use (power of two, runtime knob)
constexpr int DJ_BATCH_MAX = 8;           // tickets per dispatcher round trip (runtime knob)
constexpr int DJ_DELAY_MAX = 6;           // largest Z->B plus B->P distance in steps
constexpr int DJ_ZSLOTS = 16384;          // direct-address slots per zero / count item (64 KB)
constexpr int DJ_MAXSEG = 4 * (DJ_F + DJ_DELAY_MAX);
enum { DJ_Z = 0, DJ_B = 1, DJ_P = 2, DJ_C = 3, DJ_DONE = 4 };

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_cg_u128(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// one 32-byte sector of a read-once stream: no L1 allocation, first in line for L2 eviction (LDG.E.NA.EFL2.256) so
// that the stream does not push the direct-address regions out of L2
__device__ __forceinline__ void ld_stream256(const void* p, uint32_t (&w)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}

template <bool MAT>
__global__ void __launch_bounds__(DJ_THREADS, 2)
    k_djoin(const unsigned long long* __restrict__ build, const uint32_t* __restrict__ bcnt, uint64_t cap_b,
            const uint32_t* __restrict__ probe, const uint32_t* __restrict__ pcnt, uint64_t cap_p,
            uint32_t* __restrict__ direct, uint64_t rstride /*slots per region, multiple of 4*/, uint64_t group_bytes,
            Ctl* __restrict__ ctl, uint32_t* __restrict__ sync /*[0] ticket, [4..] zdone[256], bdone[256]*/,
            unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_vals,
            uint32_t tune /*ring | batch << 8 | Z->B steps << 16 | B->P steps << 24*/) {
  __shared__ uint32_t s_seg[DJ_MAXSEG + 1];
  __shared__ uint32_t s_preb[DJ_F + 1], s_prep[DJ_F + 1];
  __shared__ uint32_t s_warp[DJ_WARPS];
  __shared__ volatile uint32_t s_kind[DJ_RING_MAX], s_part[DJ_RING_MAX], s_chunk[DJ_RING_MAX];
  __shared__ volatile uint32_t s_seq[DJ_RING_MAX];  // n + 1 once the CTA's n-th item is published in slot n % ring
  __shared__ uint32_t s_done[DJ_RING_MAX];          // worker-warp completions per slot (monotonic)
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t ring = tune & 0xffu, batch = (tune >> 8) & 0xffu, dly_b = (tune >> 16) & 0xffu, dly_p = dly_b + (tune >> 24);
  if (tid < DJ_RING_MAX) {
    s_seq[tid] = 0;
    s_done[tid] = 0;
  }
  uint32_t* ticket = sync;
  uint32_t* zdone = sync + 4;
  uint32_t* bdone = sync + 4 + DJ_F;

  // an earlier kernel of this attempt gave up (key outside the dense domain, partition overflow): nothing to do
  if (*reinterpret_cast<volatile unsigned int*>(&ctl->flags) & (CTL_NOT_DENSE | CTL_OVERFLOW)) return;
  uint64_t reff64 = ((*reinterpret_cast<volatile unsigned long long*>(&ctl->max_key) >> 8) + 4) & ~3ull;  // slots in use per region
  if (reff64 > rstride) reff64 = rstride;
  const uint32_t reff = (uint32_t)reff64;
  const uint32_t nz = (reff + DJ_ZSLOTS - 1) / DJ_ZSLOTS;
  uint32_t gp = (uint32_t)(group_bytes / ((uint64_t)reff * 4));
  gp = gp < 1 ? 1 : (gp > DJ_F ? DJ_F : gp);
  const uint32_t ng = (DJ_F + gp - 1) / gp;

  // ---- chunks per partition and their prefix sums
  {
    uint32_t cb = 0, cp = 0;
    if (tid < DJ_F) {
      uint64_t nb = bcnt[tid], np = pcnt[tid];
      if (nb > cap_b) nb = cap_b;
      if (np > cap_p) np = cap_p;
      cb = (uint32_t)((nb + DJ_ROWS - 1) / DJ_ROWS);
      cp = nb ? (uint32_t)((np + DJ_ROWS - 1) / DJ_ROWS) : 0u;  // no build rows: nothing can match
    }
    uint32_t total;
    uint32_t pre = block_excl_scan_512(cb, s_warp, total);
    if (tid < DJ_F) s_preb[tid] = pre;
    if (tid == 0) s_preb[DJ_F] = total;
    pre = block_excl_scan_512(cp, s_warp, total);
    if (tid < DJ_F) s_prep[tid] = pre;
    if (tid == 0) s_prep[DJ_F] = total;
  }
  __syncthreads();
  // ---- the item order: 4 segments per step (Z, B, P, C); thread t sizes segments 3t .. 3t+2
  const uint32_t nseg = 4 * (ng + dly_p);
  auto seg_count = [&](uint32_t seg) -> uint32_t {
    if (seg >= nseg) return 0u;
    const uint32_t kind = seg & 3u, step = seg >> 2;
    const uint32_t delay = kind == DJ_Z ? 0u : (kind == DJ_B ? dly_b : dly_p);
    if (step < delay || step - delay >= ng) return 0u;
    const uint32_t lo = (step - delay) * gp, hi = lo + gp < DJ_F ? lo + gp : DJ_F;
    if (kind == DJ_Z) return (hi - lo) * nz;
    if (kind == DJ_C) return MAT ? (hi - lo) * nz : 0u;
    if (kind == DJ_B) return s_preb[hi] - s_preb[lo];
    return s_prep[hi] - s_prep[lo];
  };
  uint32_t total_items;
  {
    const uint32_t c0 = seg_count(3 * tid), c1 = seg_count(3 * tid + 1), c2 = seg_count(3 * tid + 2);
    const uint32_t pre = block_excl_scan_512(c0 + c1 + c2, s_warp, total_items);
    if (3u * tid < nseg + 1) s_seg[3 * tid] = pre;
    if (3u * tid + 1 < nseg + 1) s_seg[3 * tid + 1] = pre + c0;
    if (3u * tid + 2 < nseg + 1) s_seg[3 * tid + 2] = pre + c0 + c1;
  }
  __syncthreads();

  // ---- warp roles.  Warp 0 is the DISPATCHER: it pulls tickets in batches, resolves each ticket to an item
  // (kind, partition, chunk), waits for the item's dependencies and publishes it into a ring of `ring` slots.
  // Warps 1..15 are WORKERS: every warp walks the ring on its own (no block barrier per item — in the first
  // version 48 % of all warp stall samples sat in the two __syncthreads around thread 0's serial section,
  // profiles/r01c_c3_radix_ncu_summary.txt); the last warp to finish a Z / B item sends the completion signal.
  const uint32_t warp = tid >> 5;
  unsigned long long local_count = 0;
  uint32_t rows_stored = 0, slots_set = 0;
  if (warp == 0) {
    uint32_t n = 0;  // items published by this CTA so far
    bool finished = false;
    while (!finished) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(ticket, batch);
      base = __shfl_sync(0xffffffffu, base, 0);
      // lanes 0 .. batch-1 resolve one ticket each
      uint32_t kind = DJ_DONE, p = 0, c = 0;
      const uint32_t t = base + lane;
      if ((uint32_t)lane < batch && t < total_items) {
        uint32_t lo = 0, hi = nseg;  // largest seg with s_seg[seg] <= t
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s_seg[mid] <= t) lo = mid; else hi = mid;
        }
        kind = lo & 3u;
        const uint32_t step = lo >> 2;
        const uint32_t g = step - (kind == DJ_Z ? 0u : (kind == DJ_B ? dly_b : dly_p));
        const uint32_t p0 = g * gp, p1 = p0 + gp < DJ_F ? p0 + gp : DJ_F;
        const uint32_t j = t - s_seg[lo];
        if (kind == DJ_Z || kind == DJ_C) {
          p = p0 + j / nz;
          c = j - (p - p0) * nz;
        } else {
          const uint32_t* pre = kind == DJ_B ? s_preb : s_prep;
          const uint32_t want = pre[p0] + j;
          uint32_t a = p0, b = p1;  // largest p with pre[p] <= want
          while (b - a > 1) {
            const uint32_t mid = (a + b) >> 1;
            if (pre[mid] <= want) a = mid; else b = mid;
          }
          p = a;
          c = want - pre[p];
        }
      }
      auto deps_ok = [&]() -> bool {  // every dependency has a smaller ticket: it is running or finished
        if (kind == DJ_DONE || kind == DJ_Z) return true;
        if (ld_acquire_u32(zdone + p) < nz) return false;
        if (kind == DJ_B) return true;
        return ld_acquire_u32(bdone + p) >= s_preb[p + 1] - s_preb[p];
      };
      bool ok = deps_ok();  // first poll of the whole batch in parallel
#pragma unroll 1
      for (int j = 0; j < (int)batch; ++j) {
        while (!__shfl_sync(0xffffffffu, ok, j)) {  // warp-uniform loop: no lane waits at a reconvergence point
          if (lane == j) {
            __nanosleep(64);
            ok = deps_ok();
          }
        }
        const uint32_t slot = n & (ring - 1);
        // the slot is free once every worker warp has finished the item published `ring` items ago
        const uint32_t freed = (uint32_t)(DJ_WARPS - 1) * (n / ring);
        while (*reinterpret_cast<volatile uint32_t*>(&s_done[slot]) < freed) __nanosleep(32);
        if (lane == j) {
          s_kind[slot] = kind;
          s_part[slot] = p;
          s_chunk[slot] = c;
          __threadfence_block();
          s_seq[slot] = n + 1;
        }
        __syncwarp();
        ++n;
        if (__shfl_sync(0xffffffffu, kind, j) == DJ_DONE) {
          finished = true;
          break;
        }
      }
    }
  } else {
    const uint32_t wt = tid - 32;  // worker thread 0 .. DJ_WT-1
    for (uint32_t n = 0;; ++n) {
      const uint32_t slot = n & (ring - 1);
      while (s_seq[slot] != n + 1) {
      }
      __syncwarp();
      const uint32_t kind = s_kind[slot], p = s_part[slot], c = s_chunk[slot];
      if (kind == DJ_DONE) break;
      uint32_t* region = direct + (uint64_t)p * rstride;

      if (kind == DJ_Z) {
        uint4* dst = reinterpret_cast<uint4*>(region + (uint64_t)c * DJ_ZSLOTS);
        const uint32_t n4 = ((reff - c * DJ_ZSLOTS) < (uint32_t)DJ_ZSLOTS ? (reff - c * DJ_ZSLOTS) : (uint32_t)DJ_ZSLOTS) / 4;
        for (uint32_t i = wt; i < n4; i += DJ_WT) dst[i] = make_uint4(0u, 0u, 0u, 0u);
      } else if (kind == DJ_C) {
        const uint4* src = reinterpret_cast<const uint4*>(region + (uint64_t)c * DJ_ZSLOTS);
        const uint32_t n4 = ((reff - c * DJ_ZSLOTS) < (uint32_t)DJ_ZSLOTS ? (reff - c * DJ_ZSLOTS) : (uint32_t)DJ_ZSLOTS) / 4;
        for (uint32_t i = wt; i < n4; i += DJ_WT) {
          const uint4 v = ld_cg_u128(src + i);
          slots_set += (v.x != 0u) + (v.y != 0u) + (v.z != 0u) + (v.w != 0u);
        }
      } else if (kind == DJ_B) {
        uint64_t nbp = bcnt[p];
        if (nbp > cap_b) nbp = cap_b;
        const uint32_t cnt = (uint32_t)((nbp - (uint64_t)c * DJ_ROWS) < (uint64_t)DJ_ROWS ? (nbp - (uint64_t)c * DJ_ROWS) : DJ_ROWS);
        const unsigned long long* rows = build + (uint64_t)p * cap_b + (uint64_t)c * DJ_ROWS;
        uint32_t w[2][8];  // two 32-byte units per thread: rows 4u .. 4u + 3 as {value, key} word pairs
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t u = i * DJ_WT + wt;
          if (4u * u < cnt) {
            ld_stream256(rows + 4u * u, w[i]);
          } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) w[i][r] = 0xFFFFFFFFu;
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const uint32_t u = i * DJ_WT + wt;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const uint32_t k = w[i][2 * r + 1];  // packed row = key32 << 32 | value32 (little endian)
            const bool ok = (4u * u + r < cnt) & (k != 0xFFFFFFFFu);  // 0xFFFFFFFF: padding written by k_scatter2
            if (ok) region[k >> 8] = w[i][2 * r] + 1u;
            rows_stored += ok ? 1u : 0u;
          }
        }
      } else {  // DJ_P
        uint64_t npp = pcnt[p];
        if (npp > cap_p) npp = cap_p;
        const uint32_t cnt = (uint32_t)((npp - (uint64_t)c * DJ_ROWS) < (uint64_t)DJ_ROWS ? (npp - (uint64_t)c * DJ_ROWS) : DJ_ROWS);
        const uint32_t* rows = probe + (uint64_t)p * cap_p + (uint64_t)c * DJ_ROWS;
        uint32_t key[DJ_IPT], val[DJ_IPT];
        if (8u * wt < cnt) {  // one 32-byte unit per thread: rows 8 wt .. 8 wt + 7
          ld_stream256(rows + 8u * wt, key);
        } else {
#pragma unroll
          for (int i = 0; i < DJ_IPT; ++i) key[i] = 0xFFFFFFFFu;
        }
#pragma unroll
        for (int i = 0; i < DJ_IPT; ++i) key[i] = 8u * wt + i < cnt ? key[i] : 0xFFFFFFFFu;
        uint32_t hitmask = 0;
#pragma unroll
        for (int i = 0; i < DJ_IPT; ++i) {  // all 8 L2 gathers in flight
          const uint32_t idx = key[i] >> 8;
          const bool ok = (key[i] != 0xFFFFFFFFu) & (idx < reff);  // hole / past the end / beyond every build key
          val[i] = ok ? ld_cg_u32(region + idx) : 0u;
        }
#pragma unroll
        for (int i = 0; i < DJ_IPT; ++i) hitmask |= val[i] ? (1u << i) : 0u;
        if (!MAT) {
          local_count += __popc(hitmask);
        } else {
          uint32_t off[DJ_IPT];
          uint32_t wtot = 0;
#pragma unroll
          for (int i = 0; i < DJ_IPT; ++i) {
            const unsigned bal = __ballot_sync(0xffffffffu, (hitmask >> i) & 1u);
            off[i] = wtot + __popc(bal & lanemask_lt());
            wtot += __popc(bal);
          }
          unsigned long long base = 0;
          if (lane == 0 && wtot) {
            base = atomicAdd(&ctl->out_cursor, (unsigned long long)wtot);
            local_count += wtot;
          }
          base = __shfl_sync(0xffffffffu, base, 0);
#pragma unroll
          for (int i = 0; i < DJ_IPT; ++i) {
            if ((hitmask >> i) & 1u) {
              st_stream(out_keys + base + off[i], (unsigned long long)key[i]);
              st_stream(out_vals + base + off[i], (unsigned long long)(val[i] - 1u));
            }
          }
        }
      }
      __syncwarp();  // the warp's share of the item is complete
      if (lane == 0) {
        // release chain: this warp's stores -> (cta-scope fence + shared-memory atomic) -> the last warp of the item
        // -> (gpu-scope fence + global atomic) -> the waiting dispatcher's ld.acquire.  Fences are cumulative, so
        // only the last warp pays for a gpu-scope fence (one per item instead of one per warp: 18 % of the stall
        // samples of the first dispatcher version were MEMBAR.GPU, profiles/r01d_c3_radix_ncu_summary.txt)
        const bool signals = kind == DJ_Z || kind == DJ_B;
        if (signals) __threadfence_block();
        const uint32_t old = atomicAdd(&s_done[slot], 1u);
        if (signals && old + 1 == (uint32_t)(DJ_WARPS - 1) * (n / ring + 1)) {  // last warp of the item
          __threadfence();
          atomicAdd((kind == DJ_Z ? zdone : bdone) + p, 1u);
        }
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
    rows_stored += __shfl_xor_sync(0xffffffffu, rows_stored, d);
    slots_set += __shfl_xor_sync(0xffffffffu, slots_set, d);
  }
  if (lane == 0) {
    if (local_count) atomicAdd(&ctl->match_count, local_count);
    if (rows_stored) atomicAdd(&ctl->dense_rows, (unsigned long long)rows_stored);
    if (slots_set) atomicAdd(&ctl->dense_slots, (unsigned long long)slots_set);
  }
}

size_t djoin_sync_words() { return 4 + 2 * DJ_F; }
uint32_t djoin_fan() { return DJ_F; }

bool launch_djoin(bool mat, const DjoinArgs& a, const DeviceInfo& di, cudaStream_t st, int* launches) {
  int occ = 0;
  cudaError_t e = mat ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_djoin<true>, DJ_THREADS, 0)
                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_djoin<false>, DJ_THREADS, 0);
  if (e != cudaSuccess || occ < 1) return false;  // cannot guarantee co-residency: the caller takes the general path
  const unsigned grid = (unsigned)(di.sms * occ);  // every CTA must be resident: the items synchronise by spinning
  uint32_t ring = 1;
  while (ring * 2 <= (uint32_t)DJ_RING_MAX && ring * 2 <= a.ring) ring *= 2;
  const uint32_t batch = a.batch < 1 ? 1u : (a.batch > (uint32_t)DJ_BATCH_MAX ? (uint32_t)DJ_BATCH_MAX : a.batch);
  uint32_t db = a.delay_b < 1 ? 1u : a.delay_b, dp = a.delay_p < 1 ? 1u : a.delay_p;
  if (db > (uint32_t)DJ_DELAY_MAX - 1) db = DJ_DELAY_MAX - 1;
  if (db + dp > (uint32_t)DJ_DELAY_MAX) dp = DJ_DELAY_MAX - db;
  const uint32_t tune = ring | (batch << 8) | (db << 16) | (dp << 24);
  // cooperative: the work items synchronise by spinning on each other, every CTA must be resident
  const bool ok = mat ? launch_coop(k_djoin<true>, grid, DJ_THREADS, 0, st, reinterpret_cast<const unsigned long long*>(a.build), a.bcnt, a.cap_b,
                                    reinterpret_cast<const uint32_t*>(a.probe), a.pcnt, a.cap_p, a.direct, a.rstride, a.group_bytes, a.ctl,
                                    a.sync, a.out_keys, a.out_vals, tune)
                      : launch_coop(k_djoin<false>, grid, DJ_THREADS, 0, st, reinterpret_cast<const unsigned long long*>(a.build), a.bcnt, a.cap_b,
                                    reinterpret_cast<const uint32_t*>(a.probe), a.pcnt, a.cap_p, a.direct, a.rstride, a.group_bytes, a.ctl,
                                    a.sync, a.out_keys, a.out_vals, tune);
  if (!ok) return false;
  ++*launches;
  return true;
}

// probe rows whose key is the out-of-band sentinel (wide path only)
__global__ void __launch_bounds__(1024) k_emit_sentinel(Ctl* __restrict__ ctl, const unsigned long long* __restrict__ bv,
                                                        unsigned long long* __restrict__ out_keys,
                                                        unsigned long long* __restrict__ out_vals, int mat) {
  __shared__ unsigned long long s_base;
  const unsigned long long row = ctl->sentinel_row;
  const unsigned long long n = ctl->sentinel_probes;
  if (row == EMPTY64 || n == 0) return;
  if (threadIdx.x == 0) {
    atomicAdd(&ctl->match_count, n);
    s_base = mat ? atomicAdd(&ctl->out_cursor, n) : 0ull;
  }
  __syncthreads();
  if (!mat) return;
  const unsigned long long v = bv[row];
  for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
    out_keys[s_base + i] = EMPTY64;
    out_vals[s_base + i] = v;
  }
}
// out-of-band key joined with an explicit (value, probe count): the multi-GPU shuffle resolves the key's
// build row across ranks on the host
__global__ void __launch_bounds__(1024) k_emit_sentinel_value(Ctl* __restrict__ ctl, unsigned long long value,
                                                              unsigned long long n, unsigned long long* __restrict__ out_keys,
                                                              unsigned long long* __restrict__ out_vals, int mat) {
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) {
    atomicAdd(&ctl->match_count, n);
    s_base = mat ? atomicAdd(&ctl->out_cursor, n) : 0ull;
  }
  __syncthreads();
  if (!mat) return;
  for (unsigned long long i = threadIdx.x; i < n; i += blockDim.x) {
    out_keys[s_base + i] = EMPTY64;
    out_vals[s_base + i] = value;
  }
}
void launch_emit_sentinel_value(Ctl* ctl, unsigned long long value, unsigned long long n, unsigned long long* out_keys,
                                unsigned long long* out_vals, bool mat, cudaStream_t st, int* launches) {
  if (n == 0) return;
  k_emit_sentinel_value<<<1, 1024, 0, st>>>(ctl, value, n, out_keys, out_vals, mat ? 1 : 0);
  ++*launches;
}

// partition-element rows -> raw 64-bit columns, holes dropped (fallback of the shuffle path onto the
// global-table join).  *cursor receives the number of rows written.
template <bool BUILD, bool NARROW>
__global__ void __launch_bounds__(256) k_expand(const typename Elem<BUILD, NARROW>::T* __restrict__ in, uint64_t n,
                                                unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals,
                                                unsigned long long* __restrict__ cursor) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t n_round = (n + 31) & ~uint64_t(31);
  const unsigned lane = threadIdx.x & 31;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_round; i += stride) {
    typename Elem<BUILD, NARROW>::T e = Hole<BUILD, NARROW>::make();
    if (i < n) e = in[i];
    const bool ok = !Hole<BUILD, NARROW>::is(e);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    unsigned long long base = 0;
    if (lane == 0 && bal) base = atomicAdd(cursor, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (ok) {
      const unsigned long long pos = base + __popc(bal & lanemask_lt());
      keys[pos] = Elem<BUILD, NARROW>::key(e);
      if constexpr (BUILD) {
        if constexpr (NARROW) vals[pos] = e & 0xffffffffull;
        else vals[pos] = e.y;
      }
    }
  }
}
void launch_expand(bool build, bool narrow, const void* in, uint64_t n, unsigned long long* keys, unsigned long long* vals,
                   unsigned long long* cursor, const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (n == 0) return;
  uint64_t want = (n + 255) / 256;
  const uint64_t cap = (uint64_t)di.sms * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
#define FJ_EX(B, N) \
  k_expand<B, N><<<grid, 256, 0, st>>>(reinterpret_cast<const typename Elem<B, N>::T*>(in), n, keys, vals, cursor)
  if (build) { if (narrow) FJ_EX(true, true); else FJ_EX(true, false); }
  else { if (narrow) FJ_EX(false, true); else FJ_EX(false, false); }
#undef FJ_EX
  ++*launches;
}

void launch_emit_sentinel(Ctl* ctl, const unsigned long long* bv, unsigned long long* out_keys,
                          unsigned long long* out_vals, bool mat, cudaStream_t st, int* launches) {
  k_emit_sentinel<<<1, 1024, 0, st>>>(ctl, bv, out_keys, out_vals, mat ? 1 : 0);
  ++*launches;
}

}  // namespace fj
