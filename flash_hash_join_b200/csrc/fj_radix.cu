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
              int shift, uint32_t fan, Ctl* __restrict__ ctl) {
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
              if (!narrow_ok(k, v)) { atomicOr(&ctl->flags, CTL_NEED_WIDE); ok = false; }
              elem[i] = packed;
            } else {
              if (k == EMPTY64) { atomicMin(&ctl->sentinel_row, (unsigned long long)(in_base + e)); ok = false; }
              elem[i].x = k; elem[i].y = v;
            }
          } else {
            if constexpr (NARROW) {
              if ((k >> 32) != 0 || (uint32_t)k == 0xFFFFFFFFu) ok = false;  // cannot match a packed build side
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
          const uint32_t d = scatter_digit(hash32(k), shift, fan);
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
                                                 a.out_cursor, a.out_cap, a.shift, a.fan, a.ctl);
}

void launch_scatter(bool build, bool narrow, int stage, const ScatterArgs& a, const DeviceInfo& di, cudaStream_t st,
                    int* launches) {
#define FJ_SC(B, N)                                               \
  do {                                                            \
    if (stage == 1) launch_scatter_inst<B, N, 1>(a, di, st);      \
    else launch_scatter_inst<B, N, 2>(a, di, st);                 \
  } while (0)
  if (build) { if (narrow) FJ_SC(true, true); else FJ_SC(true, false); }
  else { if (narrow) FJ_SC(false, true); else FJ_SC(false, false); }
#undef FJ_SC
  ++*launches;
}

size_t radix_elem_bytes(bool build, bool narrow) {
  return build ? (narrow ? 8 : 16) : (narrow ? 4 : 8);
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

template <bool NARROW, bool MAT>
__global__ void __launch_bounds__(JN_THREADS, 2)
    k_join(const typename Elem<true, NARROW>::T* __restrict__ build, const uint32_t* __restrict__ bcnt, uint64_t cap_b,
           const typename Elem<false, NARROW>::T* __restrict__ probe, const uint32_t* __restrict__ pcnt, uint64_t cap_p,
           uint32_t smax, uint32_t tcap, uint32_t chunk, uint32_t max_chunks, Ctl* __restrict__ ctl,
           unsigned long long* __restrict__ out_keys, unsigned long long* __restrict__ out_vals) {
  using TB = typename Elem<true, NARROW>::T;
  using TP = typename Elem<false, NARROW>::T;
  constexpr int K = ProbeVec<NARROW>::K;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TB* tuples = reinterpret_cast<TB*>(smem_raw);
  uint32_t* table = reinterpret_cast<uint32_t*>(smem_raw + (size_t)smax * sizeof(TB));
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
  for (uint32_t i = tid; i < tcap; i += JN_THREADS) table[i] = 0u;
  __syncthreads();
  mbar_wait(&s_bar, 0);

  // ---- build: claim a slot per tuple with 32-bit CAS
  for (uint32_t i = tid; i < nbp; i += JN_THREADS) {
    const unsigned long long key = Elem<true, NARROW>::key(tuples[i]);
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
  const size_t smem = (size_t)a.smax * (narrow ? 8 : 16) + (size_t)a.tcap * 4;
  const uint64_t grid = (uint64_t)a.nparts * a.max_chunks;
  if (grid == 0) return;
#define FJ_JN(N, M)                                                                                           \
  do {                                                                                                        \
    auto kern = k_join<N, M>;                                                                                 \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
    kern<<<(unsigned)grid, JN_THREADS, smem, st>>>(                                                           \
        reinterpret_cast<const typename Elem<true, N>::T*>(a.build), a.bcnt, a.cap_b,                         \
        reinterpret_cast<const typename Elem<false, N>::T*>(a.probe), a.pcnt, a.cap_p, a.smax, a.tcap, a.chunk, \
        a.max_chunks, a.ctl, a.out_keys, a.out_vals);                                                         \
  } while (0)
  if (narrow) { if (mat) FJ_JN(true, true); else FJ_JN(true, false); }
  else { if (mat) FJ_JN(false, true); else FJ_JN(false, false); }
#undef FJ_JN
  ++*launches;
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

// partition-element rows -> raw 64-bit columns (fallback of the shuffle path onto the global-table join)
template <bool BUILD, bool NARROW>
__global__ void __launch_bounds__(256) k_expand(const typename Elem<BUILD, NARROW>::T* __restrict__ in, uint64_t n,
                                                unsigned long long* __restrict__ keys, unsigned long long* __restrict__ vals) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    const typename Elem<BUILD, NARROW>::T e = in[i];
    keys[i] = Elem<BUILD, NARROW>::key(e);
    if constexpr (BUILD) {
      if constexpr (NARROW) vals[i] = e & 0xffffffffull;
      else vals[i] = e.y;
    }
  }
}
void launch_expand(bool build, bool narrow, const void* in, uint64_t n, unsigned long long* keys, unsigned long long* vals,
                   const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (n == 0) return;
  uint64_t want = (n + 255) / 256;
  const uint64_t cap = (uint64_t)di.sms * 16;
  const unsigned grid = (unsigned)(want < cap ? want : cap);
#define FJ_EX(B, N) k_expand<B, N><<<grid, 256, 0, st>>>(reinterpret_cast<const typename Elem<B, N>::T*>(in), n, keys, vals)
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
