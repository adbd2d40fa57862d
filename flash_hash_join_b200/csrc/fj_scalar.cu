// fj_scalar.cu — the non-partitioned ("scalar" in the reference's vocabulary) join path:
// one open-addressing, linear-probing table in HBM/L2, built with 64-bit atomicCAS, probed by a
// persistent streaming kernel.
//
// Replaces, from /root/reference/hash_join.cpp:
//   FlashHashTable ctor/clear (:98-110)          -> cudaMemsetAsync(0xFF) of a 32 B-bucket table
//   insert_concurrent / build_concurrent (:130-151, :193-203) -> k_build<NARROW, MODE>
//   probe_vectorized (:153-182) + _hash_join_scalar_count (:536-567)        -> k_probe<.., MAT=false>
//   probe_vectorized + _hash_join_scalar_materialize (:383-496)             -> k_probe<.., MAT=true>
//   bloom set/test (:122/:142, :185-189)         -> fused into k_build / k_probe
//
// Table layout (HBM, 32-byte sector == one bucket; a probe reads exactly one sector per step):
//   narrow: 4 packed words per bucket, word = key32 << 32 | value32, EMPTY = ~0
//   wide  : 2 slots per bucket, slot = {key64, value64}, EMPTY key = ~0 (that key value itself is
//           kept out of band in Ctl::sentinel_row)
#include <algorithm>
#include <cstdio>
#include <type_traits>

#include "fj_kernels.h"

namespace fj {

// =================================================================================== ctl init
__global__ void k_init_ctl(Ctl* ctl) {
  ctl->match_count = 0;
  ctl->out_cursor = 0;
  ctl->sentinel_row = EMPTY64;
  ctl->sentinel_probes = 0;
  ctl->flags = 0;
  ctl->pad = 0;
  ctl->max_key = 0;
  ctl->dense_rows = 0;
  ctl->dense_slots = 0;
  ctl->global_count = 0;
}
void launch_init_ctl(Ctl* ctl, cudaStream_t st) { k_init_ctl<<<1, 1, 0, st>>>(ctl); }

// The attempt's verdict goes back through MAPPED pinned memory: thread t stores the t-th 32-bit word of the control block
// together with the call's 32-bit sequence tag as ONE 64-bit store (single-copy atomic, so the host can never see a torn
// word), and the host spins until every word carries the tag.  No system-scope fence is needed — it alone cost ~5 us in
// a one-thread version that stored the block, fenced and then stored a sequence word (profiles/r02fin ncu: membar 100 %).
// Against cudaMemcpyAsync + cudaStreamSynchronize the turn-around of a short call drops from 20 to 14 us
// (profiles/r02S_ubench3.jsonl) — it shows on the small joins (C1).
__global__ void k_publish_ctl(const Ctl* __restrict__ ctl, volatile unsigned long long* dst, unsigned int tag) {
  static_assert(sizeof(Ctl) == PUB_WORDS * 4, "one 64-bit store per 32-bit word of the control block");
  const unsigned int t = threadIdx.x;
  if (t < (unsigned int)PUB_WORDS) dst[t] = ((unsigned long long)tag << 32) | reinterpret_cast<const unsigned int*>(ctl)[t];
}
void launch_publish_ctl(const Ctl* ctl, void* mapped_dst, unsigned int tag, cudaStream_t st) {
  k_publish_ctl<<<1, 32, 0, st>>>(ctl, static_cast<volatile unsigned long long*>(mapped_dst), tag);
}

// One launch instead of {k_init_ctl, cudaMemsetAsync(table), cudaMemsetAsync(bloom | cursors)}: the control
// block, a 16-byte-granular region filled with all-ones (the empty table, FlashHashTable ctor :98-110) and a
// 16-byte-granular region of zeros (the Bloom filter, or the partition cursors of the radix path).
// Optionally the zeros region is an array of strided 32-bit counters (k_part's reservation cursors, one every `cs` words,
// cs a multiple of 4): counter d starts at va (d < half) or vb instead of 0.
__global__ void __launch_bounds__(512) k_prepare(Ctl* __restrict__ ctl, uint4* __restrict__ ones, uint64_t n_ones,
                                                 uint4* __restrict__ zeros, uint64_t n_zeros, uint32_t cs, uint32_t half, uint32_t va,
                                                 uint32_t vb) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t == 0) {
    ctl->match_count = 0;
    ctl->out_cursor = 0;
    ctl->sentinel_row = EMPTY64;
    ctl->sentinel_probes = 0;
    ctl->flags = 0;
    ctl->pad = 0;
    ctl->max_key = 0;
    ctl->dense_rows = 0;
    ctl->dense_slots = 0;
    ctl->global_count = 0;
  }
  const uint4 f = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu), z = make_uint4(0u, 0u, 0u, 0u);
  for (uint64_t i = t; i < n_ones; i += stride) ones[i] = f;
  if (cs == 0) {
    for (uint64_t i = t; i < n_zeros; i += stride) zeros[i] = z;
  } else {
    const uint32_t q = cs / 4;  // 16-byte pieces per counter
    for (uint64_t i = t; i < n_zeros; i += stride) {
      const uint64_t d = i / q;
      zeros[i] = (i % q == 0) ? make_uint4(d < half ? va : vb, 0u, 0u, 0u) : z;
    }
  }
}
// ones_bytes / zeros_bytes are rounded UP to 16 bytes: the buffers behind them come from the engine's arena,
// whose allocations are 256-byte granular
void launch_prepare(Ctl* ctl, void* ones, size_t ones_bytes, void* zeros, size_t zeros_bytes, const DeviceInfo& di,
                    cudaStream_t st, uint32_t counter_stride, uint32_t counters_half, uint32_t value_a, uint32_t value_b) {
  const uint64_t n1 = (ones_bytes + 15) / 16, n0 = (zeros_bytes + 15) / 16;
  const uint64_t want = (std::max(n1, n0) + 511) / 512;
  const uint64_t cap = (uint64_t)di.sms * 8;
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min(want, cap));
  k_prepare<<<grid, 512, 0, st>>>(ctl, static_cast<uint4*>(ones), n1, static_cast<uint4*>(zeros), n0, counter_stride & ~3u, counters_half, value_a,
                                  value_b);
}

// =================================================================================== build
// MODE 0: fast path.  MODE 1: exact keep-first (wide only) — the value word accumulates the
// minimum build row index with atomicMin; k_fixup_values then replaces it by bv[row].
template <bool NARROW, int MODE, bool BLOOM>
__global__ void __launch_bounds__(256) k_build(unsigned long long* __restrict__ slots, uint32_t nbuckets,
                                               uint32_t* __restrict__ bloom, uint32_t bloom_words,
                                               const unsigned long long* __restrict__ bk,
                                               const unsigned long long* __restrict__ bv, uint64_t nb,
                                               Ctl* __restrict__ ctl) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nb; i += stride) {
    const unsigned long long k = bk[i];
    const unsigned long long v = bv[i];
    if (NARROW) {
      const unsigned long long packed = (k << 32) | (v & 0xffffffffull);
      if (!narrow_ok(k, v)) {
        atomicOr(&ctl->flags, CTL_NEED_WIDE);  // attempt is abandoned by the host
        return;
      }
      const uint32_t h = hash32(k);
      uint32_t b = reduce32(h, nbuckets);
      bool done = false;
      for (uint32_t it = 0; it < nbuckets && !done; ++it) {
        unsigned long long* bp = slots + (size_t)b * 4;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (done) break;
          unsigned long long cur = bp[s];
          if (cur == EMPTY64) {
            cur = atomicCAS(bp + s, EMPTY64, packed);
            if (cur == EMPTY64) { done = true; break; }
          }
          if ((uint32_t)(cur >> 32) == (uint32_t)k) {  // same key already present
            atomicOr(&ctl->flags, CTL_DUP);
            done = true;
          }
        }
        if (++b == nbuckets) b = 0;
      }
      if (BLOOM) { const uint32_t x = bloom_hash(k); atomicOr(bloom + bloom_word(x, bloom_words), bloom_mask(x)); }
    } else {
      if (k == EMPTY64) {  // the one key that collides with the empty marker: keep out of band
        atomicMin(&ctl->sentinel_row, (unsigned long long)i);
        continue;
      }
      const uint32_t h = hash32(k);
      uint32_t b = reduce32(h, nbuckets);
      bool done = false;
      for (uint32_t it = 0; it < nbuckets && !done; ++it) {
        unsigned long long* bp = slots + (size_t)b * 4;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (done) break;
          unsigned long long cur = bp[2 * s];
          if (cur == EMPTY64) {
            cur = atomicCAS(bp + 2 * s, EMPTY64, k);
            if (cur == EMPTY64) {
              if (MODE == 0) bp[2 * s + 1] = v;
              else atomicMin(bp + 2 * s + 1, (unsigned long long)i);
              done = true;
              break;
            }
          }
          if (cur == k) {
            if (MODE == 0) atomicOr(&ctl->flags, CTL_DUP);
            else atomicMin(bp + 2 * s + 1, (unsigned long long)i);
            done = true;
          }
        }
        if (++b == nbuckets) b = 0;
      }
      if (BLOOM) { const uint32_t x = bloom_hash(k); atomicOr(bloom + bloom_word(x, bloom_words), bloom_mask(x)); }
    }
  }
}

// exact path: value word holds a build row index -> replace by the row's value
__global__ void __launch_bounds__(256) k_fixup_values(unsigned long long* __restrict__ slots, uint64_t nslots,
                                                      const unsigned long long* __restrict__ bv) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < nslots; s += stride) {
    if (slots[2 * s] != EMPTY64) slots[2 * s + 1] = bv[slots[2 * s + 1]];
  }
}

void launch_build(const TableView& t, const unsigned long long* bk, const unsigned long long* bv, uint64_t nb,
                  int mode, Ctl* ctl, const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (nb == 0) return;
  const int threads = 256;
  uint64_t want = (nb + threads - 1) / threads;
  const uint64_t cap = (uint64_t)di.sms * 16;
  const int grid = (int)(want < cap ? want : cap);
  const bool bloom = t.bloom != nullptr;
#define FJ_BUILD(N, M, B) \
  k_build<N, M, B><<<grid, threads, 0, st>>>(t.slots, t.nbuckets, t.bloom, t.bloom_words, bk, bv, nb, ctl)
  if (t.narrow) {
    if (bloom) FJ_BUILD(true, 0, true); else FJ_BUILD(true, 0, false);
  } else if (mode == 0) {
    if (bloom) FJ_BUILD(false, 0, true); else FJ_BUILD(false, 0, false);
  } else {
    if (bloom) FJ_BUILD(false, 1, true); else FJ_BUILD(false, 1, false);
  }
#undef FJ_BUILD
  ++*launches;
  if (!t.narrow && mode == 1) {
    const uint64_t nslots = (uint64_t)t.nbuckets * 2;
    uint64_t w2 = (nslots + threads - 1) / threads;
    const int g2 = (int)(w2 < cap ? w2 : cap);
    k_fixup_values<<<g2, threads, 0, st>>>(t.slots, nslots, bv);
    ++*launches;
  }
}

// =================================================================================== probe
// Lookup of one key.  The home bucket (one 32-byte sector, loaded by the caller so that several
// loads are in flight) is resolved branch-free; the linear continuation past a FULL home bucket
// (rare at load factor 0.5) lives in a separate non-inlined function.  Buckets fill in slot order
// (k_build claims the first empty slot), so "last slot empty" <=> "bucket not full".
template <bool NARROW>
__device__ __noinline__ bool probe_chain(unsigned long long key, uint32_t b, const unsigned long long* __restrict__ slots,
                                         uint32_t nbuckets, unsigned long long& value) {
  for (uint32_t it = 1; it < nbuckets; ++it) {
    if (++b == nbuckets) b = 0;
    unsigned long long s0, s1, s2, s3;
    ld_sector(slots + (size_t)b * 4, s0, s1, s2, s3);
    if (NARROW) {
      const uint32_t k32 = (uint32_t)key;
      if ((uint32_t)(s0 >> 32) == k32) { value = (uint32_t)s0; return true; }
      if ((uint32_t)(s1 >> 32) == k32) { value = (uint32_t)s1; return true; }
      if ((uint32_t)(s2 >> 32) == k32) { value = (uint32_t)s2; return true; }
      if ((uint32_t)(s3 >> 32) == k32) { value = (uint32_t)s3; return true; }
      if ((uint32_t)(s3 >> 32) == 0xFFFFFFFFu) return false;
    } else {
      if (s0 == key) { value = s1; return true; }
      if (s2 == key) { value = s3; return true; }
      if (s2 == EMPTY64) return false;
    }
  }
  return false;
}

template <bool NARROW>
__device__ __forceinline__ bool probe_home(unsigned long long key, uint32_t b, unsigned long long s0,
                                           unsigned long long s1, unsigned long long s2, unsigned long long s3,
                                           const unsigned long long* __restrict__ slots, uint32_t nbuckets,
                                           unsigned long long& value) {
  if (NARROW) {
    const uint32_t k32 = (uint32_t)key;
    const bool ok = ((key >> 32) == 0) & (k32 != 0xFFFFFFFFu);  // else: cannot be in a packed table
    const bool m0 = (uint32_t)(s0 >> 32) == k32, m1 = (uint32_t)(s1 >> 32) == k32;
    const bool m2 = (uint32_t)(s2 >> 32) == k32, m3 = (uint32_t)(s3 >> 32) == k32;
    const bool hit = ok & (m0 | m1 | m2 | m3);
    value = m0 ? (uint32_t)s0 : m1 ? (uint32_t)s1 : m2 ? (uint32_t)s2 : (uint32_t)s3;
    const bool full = (uint32_t)(s3 >> 32) != 0xFFFFFFFFu;
    if (ok & !hit & full) return probe_chain<true>(key, b, slots, nbuckets, value);
    return hit;
  } else {
    const bool ok = key != EMPTY64;  // the sentinel key is joined out of band by the caller
    const bool m0 = s0 == key, m1 = s2 == key;
    const bool hit = ok & (m0 | m1);
    value = m0 ? s1 : s3;
    if (ok & !hit & (s2 != EMPTY64)) return probe_chain<false>(key, b, slots, nbuckets, value);
    return hit;
  }
}

constexpr int PROBE_KPT = 8;   // probe keys per thread per tile (4 x 128-bit loads)
constexpr int PROBE_QCAP = 64; // per-warp survivor queue (Bloom variants), entries

// load one tile of probe keys: 128-bit coalesced loads when aligned and full, guarded otherwise.
// Returns the mask of valid key slots (0xff for a full tile).
template <int THREADS>
__device__ __forceinline__ uint32_t load_tile(const unsigned long long* __restrict__ pk, uint64_t np, uint64_t tbase, bool vec,
                                              unsigned long long (&key)[PROBE_KPT]) {
  const int tid = threadIdx.x;
  if (vec) {
#pragma unroll
    for (int r = 0; r < PROBE_KPT / 2; ++r) {
      const uint64_t e = tbase + 2ull * ((uint64_t)r * THREADS + tid);
      ld_stream2(pk + e, key[2 * r], key[2 * r + 1]);
    }
    return 0xffu;
  }
  uint32_t valid = 0;
#pragma unroll
  for (int q = 0; q < PROBE_KPT; ++q) {
    const uint64_t e = tbase + (uint64_t)q * THREADS + tid;
    const bool ok = e < np;
    key[q] = ok ? ld_stream1(pk + e) : 0ull;
    valid |= ok ? (1u << q) : 0u;
  }
  return valid;
}

// stage the whole Bloom filter in shared memory with TMA bulk copies (cp.async.bulk + mbarrier).  Only issues
// the copies: the caller waits on `bar` (phase 0) right before the first filter word is read, so the first
// tile of probe keys is already in flight while the filter lands.
__device__ __forceinline__ void stage_bloom(unsigned char* smem_dst, const uint32_t* __restrict__ bloom, uint32_t bloom_words,
                                            uint64_t* bar) {
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = bloom_words * 4u;
    mbar_expect_tx(bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
      bulk_g2s(smem_dst + off, reinterpret_cast<const unsigned char*>(bloom) + off, n, bar);
    }
  }
}

// ------------------------------------------------------------------------------------ count
// Every probe row first goes through a dense, branch-free stage with all lanes active (Bloom test,
// or the home-bucket test when there is no filter).  Rows that still need table work afterwards
// ("survivors": filter hits, or keys whose home bucket was full without a match) are a sparse
// subset; they are compacted into a per-warp shared-memory queue — one warp scan per tile, not one
// ballot per row — and the table is probed only when a full warp of survivors is available, so the
// sparse divergent work becomes dense warp work again.  When a tile has too many survivors for the
// queue (filter useless, e.g. ~100 % match rate) the warp resolves that tile directly instead.
template <bool NARROW, int BLOOM /*0 none, 1 smem, 2 global*/, int THREADS>
__global__ void __launch_bounds__(THREADS, BLOOM == 1 ? 1 : 2)
    k_probe_count(const unsigned long long* __restrict__ pk, uint64_t np, const unsigned long long* __restrict__ slots,
                  uint32_t nbuckets, const uint32_t* __restrict__ bloom, uint32_t bloom_words, Ctl* __restrict__ ctl,
                  int vec_ok) {
  constexpr uint32_t TILE = THREADS * PROBE_KPT;
  constexpr uint32_t QMASK = PROBE_QCAP - 1;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  const uint32_t* sbloom = reinterpret_cast<const uint32_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned long long* queue =
      reinterpret_cast<unsigned long long*>(smem_raw + (BLOOM == 1 ? (size_t)bloom_words * 4 : 0)) + warp * PROBE_QCAP;

  if (BLOOM == 1) stage_bloom(smem_raw, bloom, bloom_words, &s_bar);
  bool sent_present = false;
  if (!NARROW) sent_present = ctl->sentinel_row != EMPTY64;

  // full lookup of one key, starting `skip` buckets after its home bucket
  auto lookup = [&](unsigned long long k, uint32_t skip) -> bool {
    uint32_t b = reduce32(hash32(k), nbuckets) + skip;
    if (b >= nbuckets) b -= nbuckets;
    unsigned long long s0, s1, s2, s3, v;
    ld_sector(slots + (size_t)b * 4, s0, s1, s2, s3);
    bool hit = probe_home<NARROW>(k, b, s0, s1, s2, s3, slots, nbuckets, v);
    if (!NARROW) hit |= (k == EMPTY64) & sent_present;
    return hit;
  };
  constexpr uint32_t SKIP = BLOOM == 0 ? 1u : 0u;  // without a filter the queue holds keys past their home bucket

  uint32_t cnt = 0, qh = 0, qt = 0;
  const uint64_t ntiles = (np + TILE - 1) / TILE;

  // one tile: dense branch-free stage over all 8 keys of the thread, then the survivors go to the warp queue
  auto do_tile = [&](const unsigned long long (&key)[PROBE_KPT], const uint32_t valid) {
    uint32_t surv = 0;  // bit q set: key[q] needs (more) table work
    if (BLOOM == 0) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        unsigned long long s[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // 4 sector loads in flight per thread
          const uint32_t b = reduce32(hash32(key[half * 4 + j]), nbuckets);
          ld_sector(slots + (size_t)b * 4, s[j][0], s[j][1], s[j][2], s[j][3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = half * 4 + j;
          const unsigned long long k = key[q];
          const bool vq = (valid >> q) & 1u;
          bool hit, more;
          if (NARROW) {
            const uint32_t k32 = (uint32_t)k;
            const bool ok = vq & ((k >> 32) == 0) & (k32 != 0xFFFFFFFFu);
            hit = ok & (((uint32_t)(s[j][0] >> 32) == k32) | ((uint32_t)(s[j][1] >> 32) == k32) |
                        ((uint32_t)(s[j][2] >> 32) == k32) | ((uint32_t)(s[j][3] >> 32) == k32));
            more = ok & !hit & ((uint32_t)(s[j][3] >> 32) != 0xFFFFFFFFu);
          } else {
            const bool ok = vq & (k != EMPTY64);
            hit = ok & ((s[j][0] == k) | (s[j][2] == k));
            more = ok & !hit & (s[j][2] != EMPTY64);
            hit |= vq & (k == EMPTY64) & sent_present;
          }
          cnt += hit ? 1u : 0u;
          surv |= more ? (1u << q) : 0u;
        }
      }
    } else {
      uint32_t x[PROBE_KPT], w[PROBE_KPT];
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {  // all filter words requested before any is tested
        // packed table: a key with a non-zero high word cannot be present, so only the low word is hashed
        // (the build side hashes the same way: its high words are zero)
        x[q] = NARROW ? bloom_hash((uint32_t)key[q]) : bloom_hash(key[q]);
        const uint32_t wi = bloom_word(x[q], bloom_words);
        w[q] = (BLOOM == 1) ? sbloom[wi] : __ldg(bloom + wi);
      }
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        const uint32_t m = bloom_mask(x[q]);
        bool pass = (~w[q] & m) == 0u;
        if (NARROW) pass &= (key[q] >> 32) == 0;
        else pass |= key[q] == EMPTY64;  // the out-of-band key is not in the filter
        surv |= pass ? (1u << q) : 0u;
      }
      surv &= valid;
    }
    // ---- compact the tile's survivors into the warp queue (one scan per tile)
    const uint32_t c = __popc(surv);
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;  // warp-uniform
    if (qt - qh + total <= (uint32_t)PROBE_QCAP) {
      uint32_t pos = qt + incl - c;
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        if ((surv >> q) & 1u) { queue[pos & QMASK] = key[q]; ++pos; }
      }
      qt += total;
      while (qt - qh >= 32u) {
        __syncwarp();
        const unsigned long long k = queue[(qh + lane) & QMASK];
        qh += 32u;
        cnt += lookup(k, SKIP) ? 1u : 0u;
      }
      __syncwarp();
    } else {
      // dense tile: resolve the survivors in place
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        if ((surv >> q) & 1u) cnt += lookup(key[q], SKIP) ? 1u : 0u;
      }
    }
  };

  // register double buffering, unrolled by two so that the buffers never have to be copied: the next
  // tile's keys are requested before the current tile is processed, so the HBM latency of the key
  // stream overlaps the filter / table work
  unsigned long long ka[PROBE_KPT], kb[PROBE_KPT];
  uint32_t va = 0, vb = 0;
  auto fetch = [&](uint64_t tile, unsigned long long (&k)[PROBE_KPT]) -> uint32_t {
    const uint64_t tb = tile * TILE;
    return load_tile<THREADS>(pk, np, tb, vec_ok && tb + TILE <= np, k);
  };
  uint64_t tile = blockIdx.x;
  if (tile < ntiles) va = fetch(tile, ka);
  if (BLOOM == 1) mbar_wait(&s_bar, 0);
  while (tile < ntiles) {
    const uint64_t t1 = tile + gridDim.x;
    if (t1 < ntiles) vb = fetch(t1, kb);
    do_tile(ka, va);
    if (t1 >= ntiles) break;
    const uint64_t t2 = t1 + gridDim.x;
    if (t2 < ntiles) va = fetch(t2, ka);
    do_tile(kb, vb);
    tile = t2;
  }
  // drain the remainder (< 32 survivors)
  __syncwarp();
  if ((uint32_t)lane < qt - qh) cnt += lookup(queue[(qh + lane) & QMASK], SKIP) ? 1u : 0u;
  unsigned long long total = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
  if (lane == 0 && total) atomicAdd(&ctl->match_count, total);
}

// ------------------------------------------------------------------------------------ materialize
// Pairs are compacted per tile: per (warp, key slot) ballot -> block scan of the WARPS*KPT counts
// -> one bump of the global cursor per tile -> each (warp, slot) writes one contiguous run.
template <bool NARROW, int BLOOM, bool IDX, int THREADS>
__global__ void __launch_bounds__(THREADS, BLOOM == 1 ? 1 : 3)
    k_probe_mat(const unsigned long long* __restrict__ pk, uint64_t np, const unsigned long long* __restrict__ slots,
                uint32_t nbuckets, const uint32_t* __restrict__ bloom, uint32_t bloom_words,
                const unsigned long long* __restrict__ bv, Ctl* __restrict__ ctl, unsigned long long* __restrict__ out_keys,
                unsigned long long* __restrict__ out_vals, unsigned long long* __restrict__ out_idx,
                unsigned long long idx_base, int vec_ok) {
  constexpr int WARPS = THREADS / 32;
  constexpr uint32_t TILE = THREADS * PROBE_KPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_wcnt[WARPS * PROBE_KPT];
  __shared__ unsigned long long s_base;
  __shared__ __align__(8) uint64_t s_bar;
  const uint32_t* sbloom = reinterpret_cast<const uint32_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (BLOOM == 1) { stage_bloom(smem_raw, bloom, bloom_words, &s_bar); mbar_wait(&s_bar, 0); }
  bool sent_present = false;
  unsigned long long sent_value = 0;
  if (!NARROW) {
    const unsigned long long sr = ctl->sentinel_row;
    if (sr != EMPTY64) { sent_present = true; sent_value = bv[sr]; }
  }

  unsigned long long local_count = 0;
  const uint64_t ntiles = (np + TILE - 1) / TILE;
  for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const uint64_t tbase = tile * TILE;
    unsigned long long key[PROBE_KPT];
    bool valid[PROBE_KPT];
    const bool vec = vec_ok && tbase + TILE <= np;
    const uint32_t vmask = load_tile<THREADS>(pk, np, tbase, vec, key);
#pragma unroll
    for (int q = 0; q < PROBE_KPT; ++q) valid[q] = (vmask >> q) & 1u;

    using val_t = typename std::conditional<NARROW, uint32_t, unsigned long long>::type;
    val_t val[PROBE_KPT];
    uint32_t hitmask = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t b[4];
      bool need[4];
      unsigned long long s[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = half * 4 + j;
        b[j] = reduce32(hash32(key[q]), nbuckets);
        bool go = valid[q];
        if (BLOOM != 0) {
          const uint32_t x = bloom_hash(key[q]);
          const uint32_t wi = bloom_word(x, bloom_words);
          const uint32_t w = (BLOOM == 1) ? sbloom[wi] : __ldg(bloom + wi);
          const uint32_t m = bloom_mask(x);
          go = go & ((w & m) == m);
        }
        need[j] = go;
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = EMPTY64;
        if (go) ld_sector(slots + (size_t)b[j] * 4, s[j][0], s[j][1], s[j][2], s[j][3]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int q = half * 4 + j;
        unsigned long long v;
        bool hit = need[j] & probe_home<NARROW>(key[q], b[j], s[j][0], s[j][1], s[j][2], s[j][3], slots, nbuckets, v);
        if (!NARROW && valid[q] && key[q] == EMPTY64 && sent_present) { hit = true; v = sent_value; }
        val[q] = (val_t)v;
        if (hit) hitmask |= 1u << q;
      }
    }

    uint32_t rank[PROBE_KPT];
#pragma unroll
    for (int q = 0; q < PROBE_KPT; ++q) {
      const unsigned bal = __ballot_sync(0xffffffffu, (hitmask >> q) & 1u);
      rank[q] = __popc(bal & lanemask_lt());
      if (lane == 0) s_wcnt[warp * PROBE_KPT + q] = __popc(bal);
    }
    __syncthreads();
    if (warp == 0) {
      constexpr int PER = (WARPS * PROBE_KPT + 31) / 32;
      uint32_t c[PER];
      uint32_t sum = 0;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        const int e = lane * PER + u;
        c[u] = e < WARPS * PROBE_KPT ? s_wcnt[e] : 0u;
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
        if (e < WARPS * PROBE_KPT) s_wcnt[e] = run;
        run += c[u];
      }
      if (lane == 31) {
        s_base = incl ? atomicAdd(&ctl->out_cursor, (unsigned long long)incl) : 0ull;
        local_count += incl;  // counted once per tile by this lane
      }
    }
    __syncthreads();
    const unsigned long long base = s_base;
#pragma unroll
    for (int q = 0; q < PROBE_KPT; ++q) {
      if ((hitmask >> q) & 1u) {
        const unsigned long long pos = base + s_wcnt[warp * PROBE_KPT + q] + rank[q];
        st_stream(out_keys + pos, key[q]);
        st_stream(out_vals + pos, (unsigned long long)val[q]);
        if (IDX) {
          const uint64_t row = vec ? tbase + 2ull * ((uint64_t)(q >> 1) * THREADS + tid) + (q & 1)
                                   : tbase + (uint64_t)q * THREADS + tid;
          st_stream(out_idx + pos, idx_base + row);
        }
      }
    }
    __syncthreads();  // s_wcnt / s_base are reused by the next tile
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
  if (lane == 0 && local_count) atomicAdd(&ctl->match_count, local_count);
}

size_t probe_smem_bloom_limit_words(const DeviceInfo& di) {
  // the count kernel runs 32 warps next to the filter: leave room for their survivor queues (16 KB)
  // and the kernel's static shared memory
  const size_t reserve = (size_t)32 * PROBE_QCAP * 8 + 4096;
  const size_t bytes = di.smem_optin > reserve + 4096 ? di.smem_optin - reserve : 0;
  return (bytes / 16) * 4;
}

template <class K>
static uint64_t persistent_grid(K kern, int threads, size_t smem, uint64_t ntiles, int ctas_per_sm, const DeviceInfo& di) {
  if (ctas_per_sm <= 0) {  // auto: fill every SM to the kernel's occupancy limit
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem) != cudaSuccess || occ < 1) occ = 1;
    ctas_per_sm = occ;
  }
  uint64_t grid = (uint64_t)di.sms * ctas_per_sm;
  return grid > ntiles ? ntiles : grid;
}

// ------------------------------------------------------------------------------------ dense key domain
// Data-dependent fast path of the count entry points: when every build key is
// smaller than `dbits` (optimistic; a key outside raises CTL_NOT_DENSE and the host re-runs the attempt on
// the general path) the build side is an exact membership bitmap of dbits bits — a Bloom filter without
// false positives, so the count needs no table at all: num_matches = |{j : bit[pk[j]]}|
// (hash_join.cpp:536-567 counts exactly the probe rows whose key is in the build set; duplicates in the
// build side do not change that set).  The bitmap is staged in shared memory by TMA like the Bloom filter.
__global__ void __launch_bounds__(256) k_build_bitmap(uint32_t* __restrict__ bitmap, unsigned long long dbits,
                                                      const unsigned long long* __restrict__ bk, uint64_t nb,
                                                      Ctl* __restrict__ ctl) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nb; i += stride) {
    const unsigned long long k = bk[i];
    if (k >= dbits) {
      atomicOr(&ctl->flags, CTL_NOT_DENSE);  // attempt is abandoned by the host
      return;
    }
    atomicOr(bitmap + (uint32_t)(k >> 5), 1u << ((uint32_t)k & 31u));
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
    k_probe_count_dense(const unsigned long long* __restrict__ pk, uint64_t np, const uint32_t* __restrict__ bitmap,
                        uint32_t dwords /*multiple of 4*/, Ctl* __restrict__ ctl, int vec_ok) {
  constexpr uint32_t TILE = THREADS * PROBE_KPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar;
  const uint32_t* sbm = reinterpret_cast<const uint32_t*>(smem_raw);
  const int lane = threadIdx.x & 31;
  if (*reinterpret_cast<volatile unsigned int*>(&ctl->flags) & CTL_NOT_DENSE) return;  // block-uniform, before any TMA
  stage_bloom(smem_raw, bitmap, dwords, &s_bar);
  const unsigned long long dbits = (unsigned long long)dwords * 32ull;

  uint32_t cnt = 0;
  const uint64_t ntiles = (np + TILE - 1) / TILE;
  auto do_tile = [&](const unsigned long long (&key)[PROBE_KPT], const uint32_t valid) {
    uint32_t w[PROBE_KPT];
#pragma unroll
    for (int q = 0; q < PROBE_KPT; ++q) {
      const bool in = key[q] < dbits;
      w[q] = sbm[in ? (uint32_t)(key[q] >> 5) : 0u];
      w[q] = in ? w[q] : 0u;
    }
    uint32_t hits = 0;
#pragma unroll
    for (int q = 0; q < PROBE_KPT; ++q) hits |= ((w[q] >> ((uint32_t)key[q] & 31u)) & 1u) << q;
    cnt += __popc(hits & valid);
  };
  unsigned long long ka[PROBE_KPT], kb[PROBE_KPT];
  uint32_t va = 0, vb = 0;
  auto fetch = [&](uint64_t tile, unsigned long long (&k)[PROBE_KPT]) -> uint32_t {
    const uint64_t tb = tile * TILE;
    return load_tile<THREADS>(pk, np, tb, vec_ok && tb + TILE <= np, k);
  };
  uint64_t tile = blockIdx.x;
  if (tile < ntiles) va = fetch(tile, ka);
  mbar_wait(&s_bar, 0);
  while (tile < ntiles) {
    const uint64_t t1 = tile + gridDim.x;
    if (t1 < ntiles) vb = fetch(t1, kb);
    do_tile(ka, va);
    if (t1 >= ntiles) break;
    const uint64_t t2 = t1 + gridDim.x;
    if (t2 < ntiles) va = fetch(t2, ka);
    do_tile(kb, vb);
    tile = t2;
  }
  unsigned long long total = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
  if (lane == 0 && total) atomicAdd(&ctl->match_count, total);
}

// The same count as ONE launch: {k_prepare, k_build_bitmap, k_probe_count_dense} cost ~7 us each for the two
// small kernels (launch + drain), 10 % of the whole C2 step.  Here the persistent grid (all CTAs co-resident)
// zeroes the bitmap and the control block, meets at a grid barrier, inserts the build keys, meets again, and
// then every CTA copies the finished bitmap into its shared memory and streams its probe tiles; the first
// probe tile is already in flight across both barriers.  The three barrier words are self-cleaning: the last
// CTA to leave resets them, so no host-side state or memset is needed between calls.
__device__ __forceinline__ uint32_t ld_acquire_gpu_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(uint32_t* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    while (ld_acquire_gpu_u32(counter) < gridDim.x) {
    }
  }
  __syncthreads();
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
    k_count_dense_fused(const unsigned long long* __restrict__ bk, uint64_t nb, const unsigned long long* __restrict__ pk,
                        uint64_t np, uint32_t* __restrict__ bitmap, uint32_t dwords /*multiple of 4*/, Ctl* __restrict__ ctl,
                        uint32_t* __restrict__ gsync /*[0], [1] barriers, [2] exit count; all zero between launches*/,
                        int vec_ok) {
  constexpr uint32_t TILE = THREADS * PROBE_KPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint32_t* sbm = reinterpret_cast<uint32_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t gtid = blockIdx.x * (uint64_t)THREADS + tid, gthreads = (uint64_t)gridDim.x * THREADS;
  const unsigned long long dbits = (unsigned long long)dwords * 32ull;

  // ---- phase 0: control block + empty bitmap; first probe tile in flight
  if (gtid == 0) {
    ctl->match_count = 0;
    ctl->out_cursor = 0;
    ctl->sentinel_row = EMPTY64;
    ctl->sentinel_probes = 0;
    ctl->flags = 0;
    ctl->pad = 0;
    ctl->max_key = 0;
    ctl->dense_rows = 0;
    ctl->dense_slots = 0;
    ctl->global_count = 0;
  }
  for (uint64_t i = gtid; i < dwords / 4; i += gthreads) reinterpret_cast<uint4*>(bitmap)[i] = make_uint4(0u, 0u, 0u, 0u);
  const uint64_t ntiles = (np + TILE - 1) / TILE;
  unsigned long long ka[PROBE_KPT], kb[PROBE_KPT];
  uint32_t va = 0, vb = 0;
  auto fetch = [&](uint64_t tile, unsigned long long (&k)[PROBE_KPT]) -> uint32_t {
    const uint64_t tb = tile * TILE;
    return load_tile<THREADS>(pk, np, tb, vec_ok && tb + TILE <= np, k);
  };
  uint64_t tile = blockIdx.x;
  if (tile < ntiles) va = fetch(tile, ka);
  grid_barrier(gsync + 0);

  // ---- phase 1: build side -> bitmap bits (a key outside the optimistic domain abandons the attempt)
  {
    bool bad = false;
    for (uint64_t i = gtid; i < nb; i += gthreads) {
      const unsigned long long k = bk[i];
      if (k >= dbits) bad = true;
      else atomicOr(bitmap + (uint32_t)(k >> 5), 1u << ((uint32_t)k & 31u));
    }
    if (bad) atomicOr(&ctl->flags, CTL_NOT_DENSE);
  }
  grid_barrier(gsync + 1);

  // ---- phase 2: every CTA takes a private copy of the bitmap and streams its probe tiles
  uint32_t cnt = 0;
  const bool dense = !(*reinterpret_cast<volatile unsigned int*>(&ctl->flags) & CTL_NOT_DENSE);  // same answer in every CTA
  if (dense) {
    for (uint32_t i = tid; i < dwords / 4; i += THREADS) {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(reinterpret_cast<const uint4*>(bitmap) + i));
      reinterpret_cast<uint4*>(sbm)[i] = v;
    }
    __syncthreads();
    auto do_tile = [&](const unsigned long long (&key)[PROBE_KPT], const uint32_t valid) {
      uint32_t w[PROBE_KPT];
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        const bool in = key[q] < dbits;
        w[q] = sbm[in ? (uint32_t)(key[q] >> 5) : 0u];
        w[q] = in ? w[q] : 0u;
      }
      uint32_t hits = 0;
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) hits |= ((w[q] >> ((uint32_t)key[q] & 31u)) & 1u) << q;
      cnt += __popc(hits & valid);
    };
    while (tile < ntiles) {
      const uint64_t t1 = tile + gridDim.x;
      if (t1 < ntiles) vb = fetch(t1, kb);
      do_tile(ka, va);
      if (t1 >= ntiles) break;
      const uint64_t t2 = t1 + gridDim.x;
      if (t2 < ntiles) va = fetch(t2, ka);
      do_tile(kb, vb);
      tile = t2;
    }
  }
  unsigned long long total = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
  if (lane == 0 && total) atomicAdd(&ctl->match_count, total);

  // ---- leave the barrier words clean for the next launch
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(gsync + 2, 1u) == gridDim.x - 1) {
      gsync[0] = 0;
      gsync[1] = 0;
      gsync[2] = 0;
    }
  }
}

template <int THREADS>
static bool launch_count_dense_fused_inst(const unsigned long long* bk, uint64_t nb, const unsigned long long* pk, uint64_t np,
                                          uint32_t* bitmap, uint32_t dwords, Ctl* ctl, uint32_t* gsync, const DeviceInfo& di,
                                          cudaStream_t st) {
  auto kern = k_count_dense_fused<THREADS>;
  const size_t smem = (size_t)dwords * 4;
  static size_t smem_set = 0;   // per instantiation: attribute and occupancy are looked up once per size and device
  static int occ_cached = 0, dev_cached = -1;
  if (smem != smem_set || occ_cached == 0 || dev_cached != di.device) {
    dev_cached = di.device;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) return false;
    smem_set = smem;
    occ_cached = occ;
  }
  // every CTA must be resident (the phases meet at spinning grid barriers): never more than sms * occupancy
  const uint64_t tile = (uint64_t)THREADS * PROBE_KPT;
  uint64_t grid = (uint64_t)di.sms * occ_cached;
  const uint64_t ntiles = (np + tile - 1) / tile;
  if (grid > ntiles) grid = ntiles ? ntiles : 1;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(pk) & 15u) == 0) ? 1 : 0;
  return launch_coop(kern, (unsigned)grid, THREADS, smem, st, bk, nb, pk, np, bitmap, dwords, ctl, gsync, vec_ok);
}
bool launch_count_dense_fused(const unsigned long long* bk, uint64_t nb, const unsigned long long* pk, uint64_t np,
                              uint32_t* bitmap, uint32_t dwords, Ctl* ctl, uint32_t* gsync, const DeviceInfo& di, cudaStream_t st,
                              int* launches) {
  bool ok;
  if ((size_t)dwords * 4 * 2 + 4096 <= di.smem_optin)
    ok = launch_count_dense_fused_inst<512>(bk, nb, pk, np, bitmap, dwords, ctl, gsync, di, st);
  else
    ok = launch_count_dense_fused_inst<1024>(bk, nb, pk, np, bitmap, dwords, ctl, gsync, di, st);
  if (ok) ++*launches;
  return ok;
}

// ---- the multi-GPU count as ONE kernel per GPU, no NCCL call in the step ------------------------------------------
// BROADCAST mode (small build side replicated, probe side split over the GPUs; BASELINE.json configs[3] and the
// contract bench at N > 1): every rank owns a small cudaMalloc'd exchange buffer that every other rank has mapped
// through CUDA IPC at fj_comm_init (NVLink / NVSwitch peer memory):
//     word 0            ready : the step number whose build keys the staging area holds (written by the root's kernel)
//     words 8 + 8 r ..  slot r: { seq, count[2] } written by rank r's kernel (count indexed by step parity)
//     word 8192 ..      staging area for the build keys (root only)
// The kernel is k_count_dense_fused plus two exchanges done with system-scope loads / stores over NVLink:
//   * instead of ncclBroadcast, every rank reads the build keys straight out of the ROOT's staging area;
//   * instead of ncclAllReduce, the last CTA to finish posts the rank's count into every peer's slot, waits until
//     all ranks' slots carry this step, and leaves the sum in Ctl::global_count.
// A rank can be at most one step ahead of another (it needs everybody's slot of step k to finish step k), so
// two count entries per slot are enough.  Spins give up after 10 s (CTL_PEER_TIMEOUT).
constexpr int PEER_READY_WORD = 0;
constexpr int PEER_SLOT_WORD = 8;       // + 8 * rank
constexpr int PEER_STAGING_WORD = 8192;  // 64 KB into the buffer
// Relay (large build sides): every rank pulls only ITS slice of the build keys from the root and sets their bits in a
// partial bitmap that lives in its exchange buffer (PEER_PARTIAL bytes behind the staging area); after a cross-GPU
// barrier every rank ORs the partial bitmaps of all ranks into its own full bitmap.  The root then sends every key once
// (nb * 8 bytes leave it instead of (world - 1) * nb * 8: at 1e6 keys and 8 GPUs the root's NVLink egress made the
// count step 0.30 ms against 0.21 ms on one GPU, profiles/r02A_bench_n8.json); a bitmap is 2 bits per key.
constexpr int PEER_PARTIAL_WORD = 4096;            // + rank: step | bad << 63 posted by `rank` when its partial bitmap is complete
constexpr int PEER_BCAST_WORD = 4160;              // + rank: step posted by `rank` when its slice of the broadcast rows has arrived
constexpr int PEER_RED_WORD = 4224;                // + 8 * rank: { step, -, sum[2], or[2] } of the small reductions (k_peer_reduce), by step parity
constexpr size_t PEER_STAGING_BYTES = size_t(16) << 20;
constexpr size_t PEER_PARTIAL_BYTES = size_t(256) << 10;
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= want (system scope); false after 10 s
__device__ __forceinline__ bool wait_sys_ge(const unsigned long long* p, unsigned long long want) {
  if (ld_acquire_sys_u64(p) >= want) return true;
  const unsigned long long t0 = globaltimer_ns();
  for (;;) {
    if (ld_acquire_sys_u64(p) >= want) return true;
    __nanosleep(200);
    if (globaltimer_ns() - t0 > 10000000000ull) return false;
  }
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
    k_count_dense_peer(uint64_t nb, const unsigned long long* __restrict__ pk, uint64_t np, uint32_t* __restrict__ bitmap,
                       uint32_t dwords /*multiple of 4*/, Ctl* __restrict__ ctl, uint32_t* __restrict__ gsync, int vec_ok,
                       unsigned long long* const* __restrict__ peers /*[world] exchange buffers, peers[rank] is local*/,
                       int rank, int world, int root, unsigned long long step, int relay) {
  constexpr uint32_t TILE = THREADS * PROBE_KPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ int s_flag;
  uint32_t* sbm = reinterpret_cast<uint32_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t gtid = blockIdx.x * (uint64_t)THREADS + tid, gthreads = (uint64_t)gridDim.x * THREADS;
  const unsigned long long dbits = (unsigned long long)dwords * 32ull;
  unsigned long long* const root_buf = peers[root];
  unsigned long long* const my_buf = peers[rank];

  // ---- phase 0: control block + empty bitmap; first probe tile in flight; the root announces its staging area
  // (the copy into it preceded this kernel in the root's stream)
  if (gtid == 0) {
    ctl->match_count = 0;
    ctl->out_cursor = 0;
    ctl->sentinel_row = EMPTY64;
    ctl->sentinel_probes = 0;
    ctl->flags = 0;
    ctl->pad = 0;
    ctl->max_key = 0;
    ctl->dense_rows = 0;
    ctl->dense_slots = 0;
    ctl->global_count = 0;
    if (rank == root) st_release_sys_u64(root_buf + PEER_READY_WORD, step);
  }
  // relay: the bits go into the partial bitmap in my exchange buffer (peers read it), the full bitmap is assembled later
  uint32_t* const partial = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(my_buf) + (size_t)PEER_STAGING_WORD * 8 + PEER_STAGING_BYTES);
  uint32_t* const target = relay ? partial : bitmap;
  for (uint64_t i = gtid; i < dwords / 4; i += gthreads) reinterpret_cast<uint4*>(target)[i] = make_uint4(0u, 0u, 0u, 0u);
  const uint64_t ntiles = (np + TILE - 1) / TILE;
  unsigned long long ka[PROBE_KPT], kb[PROBE_KPT];
  uint32_t va = 0, vb = 0;
  auto fetch = [&](uint64_t tile, unsigned long long (&k)[PROBE_KPT]) -> uint32_t {
    const uint64_t tb = tile * TILE;
    return load_tile<THREADS>(pk, np, tb, vec_ok && tb + TILE <= np, k);
  };
  uint64_t tile = blockIdx.x;
  if (tile < ntiles) va = fetch(tile, ka);
  grid_barrier(gsync + 0);

  // ---- phase 1: build keys come straight out of the root's staging area (peer memory for every other rank)
  {
    if (rank != root) {
      if (tid == 0) s_flag = wait_sys_ge(root_buf + PEER_READY_WORD, step) ? 1 : 0;
      __syncthreads();
      if (!s_flag && tid == 0) atomicOr(&ctl->flags, CTL_PEER_TIMEOUT);
    }
    const bool have = rank == root || s_flag;
    const unsigned long long* bk = root_buf + PEER_STAGING_WORD;
    bool bad = false;
    if (have) {
      // pairs of keys with 16-byte system-scope loads, four loads in flight per thread before the first bitmap update:
      // a load from the root's memory is an NVLink round trip (one 8-byte load per loop iteration, each waited for, cost
      // 13 serial round trips per thread at 1e6 build keys: C4 count at 8 GPUs 0.30 ms vs 0.21 ms on one GPU,
      // profiles/r02s_bench_n8.json)
      auto put = [&](unsigned long long k) {
        if (k >= dbits) bad = true;
        else atomicOr(target + (uint32_t)(k >> 5), 1u << ((uint32_t)k & 31u));
      };
      const uint64_t npair = nb / 2;  // the staging area is 16-byte aligned
      // relay: pairs [p0, p1) are this rank's slice
      const uint64_t per = relay ? (npair + (uint64_t)world - 1) / (uint64_t)world : npair;
      const uint64_t p0 = relay ? (uint64_t)rank * per : 0, p1 = p0 + per < npair ? p0 + per : npair;
      for (uint64_t i0 = p0 + gtid; i0 < p1; i0 += 4 * gthreads) {
        unsigned long long a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint64_t i = i0 + (uint64_t)u * gthreads;
          a[u] = b[u] = ~0ull;
          if (i < p1) asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(a[u]), "=l"(b[u]) : "l"(bk + 2 * i));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (i0 + (uint64_t)u * gthreads < p1) {
            put(a[u]);
            put(b[u]);
          }
        }
      }
      if ((nb & 1ull) && gtid == 0 && (!relay || rank == world - 1)) put(ld_relaxed_sys_u64(bk + nb - 1));
    }
    if (bad) atomicOr(&ctl->flags, CTL_NOT_DENSE);
  }
  grid_barrier(gsync + 1);
  if (relay) {
    // ---- phase 1b: my partial bitmap is complete: tell every rank (with my verdict on the keys I saw), wait for all
    // partial bitmaps, OR them into my full bitmap
    if (gtid == 0) {
      __threadfence_system();
      const unsigned long long bad_here = (*reinterpret_cast<volatile unsigned int*>(&ctl->flags) & CTL_NOT_DENSE) ? (1ull << 63) : 0ull;
      for (int r = 0; r < world; ++r) st_release_sys_u64(peers[r] + PEER_PARTIAL_WORD + rank, step | bad_here);
    }
    if (tid == 0) {
      int ok = 1, anybad = 0;
      for (int r = 0; r < world && ok; ++r) {
        const unsigned long long* wd = my_buf + PEER_PARTIAL_WORD + r;
        unsigned long long v = ld_acquire_sys_u64(wd);
        if ((v & ~(1ull << 63)) < step) {
          const unsigned long long t0 = globaltimer_ns();
          for (;;) {
            v = ld_acquire_sys_u64(wd);
            if ((v & ~(1ull << 63)) >= step) break;
            __nanosleep(200);
            if (globaltimer_ns() - t0 > 10000000000ull) { ok = 0; break; }
          }
        }
        anybad |= (int)(v >> 63);
      }
      if (!ok) atomicOr(&ctl->flags, CTL_PEER_TIMEOUT);
      if (anybad) atomicOr(&ctl->flags, CTL_NOT_DENSE);  // some rank saw a key outside the domain: everybody falls back
      s_flag = ok;
    }
    __syncthreads();
    if (s_flag) {
      for (uint64_t i = gtid; i < dwords / 4; i += gthreads) {
        uint4 acc = make_uint4(0u, 0u, 0u, 0u);
        for (int r = 0; r < world; ++r) {
          const uint4* pp = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(peers[r]) + (size_t)PEER_STAGING_WORD * 8 + PEER_STAGING_BYTES) + i;
          uint4 v;
          asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(pp));
          acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
        }
        reinterpret_cast<uint4*>(bitmap)[i] = acc;
      }
    }
    grid_barrier(gsync + 3);  // (a barrier word serves once per launch: the fourth word is this kernel's own)
  }

  // ---- phase 2: private copy of the bitmap, stream the probe tiles
  uint32_t cnt = 0;
  const bool dense = !(*reinterpret_cast<volatile unsigned int*>(&ctl->flags) & (CTL_NOT_DENSE | CTL_PEER_TIMEOUT));
  if (dense) {
    for (uint32_t i = tid; i < dwords / 4; i += THREADS) {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(reinterpret_cast<const uint4*>(bitmap) + i));
      reinterpret_cast<uint4*>(sbm)[i] = v;
    }
    __syncthreads();
    auto do_tile = [&](const unsigned long long (&key)[PROBE_KPT], const uint32_t valid) {
      uint32_t w[PROBE_KPT];
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        const bool in = key[q] < dbits;
        w[q] = sbm[in ? (uint32_t)(key[q] >> 5) : 0u];
        w[q] = in ? w[q] : 0u;
      }
      uint32_t hits = 0;
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) hits |= ((w[q] >> ((uint32_t)key[q] & 31u)) & 1u) << q;
      cnt += __popc(hits & valid);
    };
    while (tile < ntiles) {
      const uint64_t t1 = tile + gridDim.x;
      if (t1 < ntiles) vb = fetch(t1, kb);
      do_tile(ka, va);
      if (t1 >= ntiles) break;
      const uint64_t t2 = t1 + gridDim.x;
      if (t2 < ntiles) va = fetch(t2, ka);
      do_tile(kb, vb);
      tile = t2;
    }
  }
  unsigned long long total = cnt;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
  if (lane == 0 && total) atomicAdd(&ctl->match_count, total);

  // ---- exit: the last CTA of this rank exchanges the count with every peer and sums (replaces ncclAllReduce)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_flag = atomicAdd(gsync + 2, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (s_flag && tid < 32) {  // warp 0 of the last CTA: lane r talks to rank r (world <= 32)
    __threadfence();
    const unsigned long long mine = atomicAdd(&ctl->match_count, 0ull);
    const int par = (int)(step & 1ull);
    if (lane < world) {
      unsigned long long* slot = peers[lane] + PEER_SLOT_WORD + 8 * rank;  // my slot in rank `lane`'s buffer
      st_relaxed_sys_u64(slot + 1 + par, mine);
      st_release_sys_u64(slot, step);
    }
    unsigned long long got = 0;
    bool ok = true;
    if (lane < world) {
      const unsigned long long* slot = my_buf + PEER_SLOT_WORD + 8 * lane;  // rank `lane`'s slot in my buffer
      ok = wait_sys_ge(slot, step);
      got = ok ? ld_relaxed_sys_u64(slot + 1 + par) : 0ull;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) got += __shfl_xor_sync(0xffffffffu, got, d);
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
      if (!all_ok) atomicOr(&ctl->flags, CTL_PEER_TIMEOUT);
      ctl->global_count = got;
      gsync[0] = 0;
      gsync[1] = 0;
      gsync[2] = 0;
      gsync[3] = 0;
    }
  }
}

template <int THREADS>
static bool launch_count_dense_peer_inst(uint64_t nb, const unsigned long long* pk, uint64_t np, uint32_t* bitmap, uint32_t dwords,
                                         Ctl* ctl, uint32_t* gsync, unsigned long long* const* peers, int rank, int world, int root,
                                         unsigned long long step, int relay, const DeviceInfo& di, cudaStream_t st) {
  auto kern = k_count_dense_peer<THREADS>;
  const size_t smem = (size_t)dwords * 4;
  static size_t smem_set = 0;
  static int occ_cached = 0, dev_cached = -1;
  if (smem != smem_set || occ_cached == 0 || dev_cached != di.device) {
    dev_cached = di.device;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) return false;
    smem_set = smem;
    occ_cached = occ;
  }
  const uint64_t tile = (uint64_t)THREADS * PROBE_KPT;
  uint64_t grid = (uint64_t)di.sms * occ_cached;  // all CTAs resident (grid barriers)
  const uint64_t ntiles = (np + tile - 1) / tile;
  if (grid > ntiles) grid = ntiles ? ntiles : 1;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(pk) & 15u) == 0) ? 1 : 0;
  return launch_coop(kern, (unsigned)grid, THREADS, smem, st, nb, pk, np, bitmap, dwords, ctl, gsync, vec_ok, peers, rank, world, root, step, relay);
}
// Broadcast of the root's build rows to every rank over peer memory (replaces ncclBroadcast for build sides that fit the
// staging area; BASELINE.json configs[3]: 1e6 rows x 16 bytes to 8 GPUs took 0.2 ms with NCCL).  The root's rows lie in
// its staging area (keys, then values: `words` 64-bit words).  Two hops, like an all-gather: every rank first pulls ITS
// 1/world slice from the root into the same place of its own staging area, the ranks meet at a cross-GPU barrier, then
// every rank pulls the other slices from the ranks that hold them into `out` — the root sends every word once, every
// link carries words / world, and the pulls of the second hop come from world - 1 different sources at once.
__global__ void __launch_bounds__(512) k_peer_bcast(unsigned long long* const* __restrict__ peers, int rank, int world, int root,
                                                     unsigned long long step, uint64_t words /*multiple of 2*/,
                                                     unsigned long long* __restrict__ out, uint32_t* __restrict__ err /*set on a timeout*/,
                                                     uint32_t* __restrict__ gsync) {
  __shared__ int s_flag;
  const int tid = threadIdx.x;
  const uint64_t gtid = blockIdx.x * (uint64_t)blockDim.x + tid, gthreads = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long* const my_stage = peers[rank] + PEER_STAGING_WORD;
  const unsigned long long* const root_stage = peers[root] + PEER_STAGING_WORD;
  const uint64_t pairs = words / 2, per = (pairs + (uint64_t)world - 1) / (uint64_t)world;
  auto copy_pairs = [&](const unsigned long long* src, unsigned long long* dst, uint64_t p0, uint64_t p1, bool remote) {
    for (uint64_t i0 = p0 + gtid; i0 < p1; i0 += 4 * gthreads) {  // four 16-byte loads in flight per thread
      unsigned long long a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t i = i0 + (uint64_t)u * gthreads;
        if (i < p1) {
          if (remote) asm volatile("ld.relaxed.sys.global.v2.u64 {%0,%1}, [%2];" : "=l"(a[u]), "=l"(b[u]) : "l"(src + 2 * i));
          else { a[u] = src[2 * i]; b[u] = src[2 * i + 1]; }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t i = i0 + (uint64_t)u * gthreads;
        if (i < p1) { dst[2 * i] = a[u]; dst[2 * i + 1] = b[u]; }
      }
    }
  };
  // ---- hop 1: my slice, from the root (the root's copy into its staging area preceded this kernel in its stream)
  if (rank == root) {
    if (gtid == 0) st_release_sys_u64(peers[root] + PEER_READY_WORD, step);
  } else {
    if (tid == 0) s_flag = wait_sys_ge(root_stage - PEER_STAGING_WORD + PEER_READY_WORD, step) ? 1 : 0;
    __syncthreads();
    if (!s_flag) {
      if (tid == 0) atomicOr(err, 1u);
    } else {
      const uint64_t p0 = (uint64_t)rank * per, p1 = p0 + per < pairs ? p0 + per : pairs;
      if (p0 < p1) copy_pairs(root_stage, my_stage, p0, p1, true);
    }
  }
  grid_barrier(gsync + 0);
  // ---- my slice is in place: tell everybody, wait for everybody
  if (gtid == 0) {
    __threadfence_system();
    for (int r = 0; r < world; ++r) st_release_sys_u64(peers[r] + PEER_BCAST_WORD + rank, step);
  }
  if (tid == 0) {
    int ok = 1;
    for (int r = 0; r < world && ok; ++r) ok = wait_sys_ge(peers[rank] + PEER_BCAST_WORD + r, step) ? 1 : 0;
    if (!ok) atomicOr(err, 1u);
    s_flag = ok;
  }
  __syncthreads();
  // ---- hop 2: every slice from the rank that holds it (my own and, on the root, all of them: local copies)
  if (s_flag) {
    for (int q = 0; q < world; ++q) {
      const int src_rank = (rank + q) % world;  // start with the local slice, then a different peer on every rank
      const uint64_t p0 = (uint64_t)src_rank * per, p1 = p0 + per < pairs ? p0 + per : pairs;
      if (p0 >= p1) continue;
      const bool local = src_rank == rank || rank == root;
      copy_pairs((local ? my_stage : peers[src_rank] + PEER_STAGING_WORD), out, p0, p1, !local);
    }
  }
  // the barrier word cleans itself: the last CTA to get here resets both
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(gsync + 2, 1u) == gridDim.x - 1) {
      gsync[0] = 0;
      gsync[2] = 0;
    }
  }
}
// Sum / OR of two words over all ranks through peer memory (replaces an 8-byte ncclAllReduce: the agreement on
// rank-local failures and the global match count of the broadcast path).  One CTA; two entries per rank by step parity
// (a rank can be at most one reduction ahead of another).  sum_src / or_src may be null (then sum_imm / 0 are used).
__global__ void __launch_bounds__(32) k_peer_reduce(unsigned long long* const* __restrict__ peers, int rank, int world,
                                                     unsigned long long step, const unsigned long long* __restrict__ sum_src,
                                                     unsigned long long sum_imm, const unsigned int* __restrict__ or_src,
                                                     unsigned long long* __restrict__ result /*[2]: sum, or*/, uint32_t* __restrict__ err) {
  const int lane = threadIdx.x;
  const int par = (int)(step & 1ull);
  const unsigned long long mine = sum_src ? *sum_src : sum_imm;
  const unsigned long long mor = or_src ? (unsigned long long)*or_src : 0ull;
  if (lane < world) {
    unsigned long long* slot = peers[lane] + PEER_RED_WORD + 8 * rank;
    st_relaxed_sys_u64(slot + 2 + par, mine);
    st_relaxed_sys_u64(slot + 4 + par, mor);
    st_release_sys_u64(slot, step);
  }
  unsigned long long gs = 0, go = 0;
  bool ok = true;
  if (lane < world) {
    const unsigned long long* slot = peers[rank] + PEER_RED_WORD + 8 * lane;
    ok = wait_sys_ge(slot, step);
    if (ok) {
      gs = ld_relaxed_sys_u64(slot + 2 + par);
      go = ld_relaxed_sys_u64(slot + 4 + par);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    gs += __shfl_xor_sync(0xffffffffu, gs, d);
    go |= __shfl_xor_sync(0xffffffffu, go, d);
  }
  const bool all_ok = __all_sync(0xffffffffu, ok);
  if (lane == 0) {
    if (!all_ok) atomicOr(err, 1u);
    result[0] = gs;
    result[1] = go;
  }
}
void launch_peer_reduce(unsigned long long* const* peers, int rank, int world, unsigned long long step, const unsigned long long* sum_src,
                        unsigned long long sum_imm, const unsigned int* or_src, unsigned long long* result, uint32_t* err, cudaStream_t st,
                        int* launches) {
  k_peer_reduce<<<1, 32, 0, st>>>(peers, rank, world, step, sum_src, sum_imm, or_src, result, err);
  ++*launches;
}

bool launch_peer_bcast(unsigned long long* const* peers, int rank, int world, int root, unsigned long long step, uint64_t words,
                       unsigned long long* out, uint32_t* err, uint32_t* gsync, const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (words == 0 || (words & 1ull) || words * 8 > PEER_STAGING_BYTES) return false;
  const uint64_t want = (words / 2 + 511) / 512;
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)di.sms));  // one CTA per SM: co-resident
  if (!launch_coop(k_peer_bcast, grid, 512u, 0, st, peers, rank, world, root, step, words, out, err, gsync)) return false;
  ++*launches;
  return true;
}

size_t peer_staging_offset_bytes() { return (size_t)PEER_STAGING_WORD * 8; }
size_t peer_staging_bytes() { return PEER_STAGING_BYTES; }
size_t peer_buffer_bytes() { return (size_t)PEER_STAGING_WORD * 8 + PEER_STAGING_BYTES + PEER_PARTIAL_BYTES; }
bool launch_count_dense_peer(uint64_t nb, const unsigned long long* pk, uint64_t np, uint32_t* bitmap, uint32_t dwords, Ctl* ctl,
                             uint32_t* gsync, unsigned long long* const* peers, int rank, int world, int root,
                             unsigned long long step, bool relay, const DeviceInfo& di, cudaStream_t st, int* launches) {
  bool ok;
  const int rl = relay && world > 1 && (size_t)dwords * 4 <= PEER_PARTIAL_BYTES ? 1 : 0;
  if ((size_t)dwords * 4 * 2 + 4096 <= di.smem_optin)
    ok = launch_count_dense_peer_inst<512>(nb, pk, np, bitmap, dwords, ctl, gsync, peers, rank, world, root, step, rl, di, st);
  else
    ok = launch_count_dense_peer_inst<1024>(nb, pk, np, bitmap, dwords, ctl, gsync, peers, rank, world, root, step, rl, di, st);
  if (ok) ++*launches;
  return ok;
}

// Materialize on a dense key domain (global-table path, hash_join.cpp:383-496): the same persistent launch,
// with a direct-address value table next to the bitmap: direct[key] = build value (8 bytes, L2 resident:
// 2 MB at C2, 14 MB at 1e6 build rows).  The exact bitmap in shared memory answers "does this probe row
// match" without touching memory; only matching rows gather their value from L2, and pairs are compacted per
// tile like k_probe_mat (one cursor bump per tile).  A duplicate build key shows up as an already-set bit
// (atomicOr returns it) -> CTL_DUP: keep-first needs row order, the host re-runs on the exact path.
template <bool IDX, int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 512 ? 2 : 1)
    k_mat_dense_fused(const unsigned long long* __restrict__ bk, const unsigned long long* __restrict__ bv, uint64_t nb,
                      const unsigned long long* __restrict__ pk, uint64_t np, uint32_t* __restrict__ bitmap,
                      uint32_t dwords /*multiple of 4*/, unsigned long long* __restrict__ direct, Ctl* __restrict__ ctl,
                      uint32_t* __restrict__ gsync, unsigned long long* __restrict__ out_keys,
                      unsigned long long* __restrict__ out_vals, unsigned long long* __restrict__ out_idx,
                      unsigned long long idx_base, int vec_ok) {
  // the CTA shares one bitmap but probes as NG independent groups of 256 threads (own tiles, own compaction scratch,
  // own named barrier): a 1024-thread CTA whose 32 warps met at three block barriers per tile ran 30 % slower
  // than the hash-table kernel at 90 % match rate (C4 shape, profiles/r01g_quick_bench.jsonl)
  constexpr int GT = 256, NG = THREADS / GT, WARPS = GT / 32;
  constexpr uint32_t TILE = GT * PROBE_KPT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t s_wcnt_all[NG][WARPS * PROBE_KPT];
  __shared__ unsigned long long s_base_all[NG];
  uint32_t* sbm = reinterpret_cast<uint32_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, grp = tid / GT, gtid_l = tid % GT, warp = gtid_l >> 5;
  uint32_t* s_wcnt = s_wcnt_all[grp];
  unsigned long long& s_base = s_base_all[grp];
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(GT) : "memory"); };
  const uint64_t gtid = blockIdx.x * (uint64_t)THREADS + tid, gthreads = (uint64_t)gridDim.x * THREADS;
  const unsigned long long dbits = (unsigned long long)dwords * 32ull;

  // ---- phase 0
  if (gtid == 0) {
    ctl->match_count = 0;
    ctl->out_cursor = 0;
    ctl->sentinel_row = EMPTY64;
    ctl->sentinel_probes = 0;
    ctl->flags = 0;
    ctl->pad = 0;
    ctl->max_key = 0;
    ctl->dense_rows = 0;
    ctl->dense_slots = 0;
    ctl->global_count = 0;
  }
  for (uint64_t i = gtid; i < dwords / 4; i += gthreads) reinterpret_cast<uint4*>(bitmap)[i] = make_uint4(0u, 0u, 0u, 0u);
  grid_barrier(gsync + 0);

  // ---- phase 1: bitmap bits + direct-address values
  {
    unsigned bad = 0;
    for (uint64_t i = gtid; i < nb; i += gthreads) {
      const unsigned long long k = bk[i];
      if (k >= dbits) {
        bad |= CTL_NOT_DENSE;
      } else {
        const uint32_t bit = 1u << ((uint32_t)k & 31u);
        const uint32_t old = atomicOr(bitmap + (uint32_t)(k >> 5), bit);
        if (old & bit) bad |= CTL_DUP;
        else direct[k] = bv[i];
      }
    }
    if (bad) atomicOr(&ctl->flags, bad);
  }
  grid_barrier(gsync + 1);

  // ---- phase 2
  unsigned long long local_count = 0;
  const bool go = !(*reinterpret_cast<volatile unsigned int*>(&ctl->flags) & (CTL_NOT_DENSE | CTL_DUP));  // grid-uniform
  if (go) {
    for (uint32_t i = tid; i < dwords / 4; i += THREADS) {
      uint4 v;
      asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                   : "l"(reinterpret_cast<const uint4*>(bitmap) + i));
      reinterpret_cast<uint4*>(sbm)[i] = v;
    }
    __syncthreads();
    const uint64_t ntiles = (np + TILE - 1) / TILE;
    for (uint64_t tile = (uint64_t)blockIdx.x * NG + grp; tile < ntiles; tile += (uint64_t)gridDim.x * NG) {
      const uint64_t tbase = tile * TILE;
      unsigned long long key[PROBE_KPT], val[PROBE_KPT];
      const bool vec = vec_ok && tbase + TILE <= np;
      uint32_t vmask = 0;
      if (vec) {
#pragma unroll
        for (int r = 0; r < PROBE_KPT / 2; ++r) ld_stream2(pk + tbase + 2ull * ((uint64_t)r * GT + gtid_l), key[2 * r], key[2 * r + 1]);
        vmask = 0xffu;
      } else {
#pragma unroll
        for (int q = 0; q < PROBE_KPT; ++q) {
          const uint64_t e = tbase + (uint64_t)q * GT + gtid_l;
          const bool ok = e < np;
          key[q] = ok ? ld_stream1(pk + e) : 0ull;
          vmask |= ok ? (1u << q) : 0u;
        }
      }
      uint32_t hitmask = 0;
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        const bool in = key[q] < dbits;
        const uint32_t w = sbm[in ? (uint32_t)(key[q] >> 5) : 0u];
        const bool hit = in & ((vmask >> q) & 1u) & ((w >> ((uint32_t)key[q] & 31u)) & 1u);
        hitmask |= hit ? (1u << q) : 0u;
      }
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {  // all gathers of the thread in flight (L2-resident table)
        val[q] = 0ull;
        if ((hitmask >> q) & 1u)
          asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(val[q]) : "l"(direct + key[q]));
      }
      uint32_t rank[PROBE_KPT];
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        const unsigned bal = __ballot_sync(0xffffffffu, (hitmask >> q) & 1u);
        rank[q] = __popc(bal & lanemask_lt());
        if (lane == 0) s_wcnt[warp * PROBE_KPT + q] = __popc(bal);
      }
      group_sync();
      if (warp == 0) {
        constexpr int PER = (WARPS * PROBE_KPT + 31) / 32;
        uint32_t c[PER];
        uint32_t sum = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          const int e = lane * PER + u;
          c[u] = e < WARPS * PROBE_KPT ? s_wcnt[e] : 0u;
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
          if (e < WARPS * PROBE_KPT) s_wcnt[e] = run;
          run += c[u];
        }
        if (lane == 31) {
          s_base = incl ? atomicAdd(&ctl->out_cursor, (unsigned long long)incl) : 0ull;
          local_count += incl;  // counted once per tile by this lane
        }
      }
      group_sync();
      const unsigned long long base = s_base;
#pragma unroll
      for (int q = 0; q < PROBE_KPT; ++q) {
        if ((hitmask >> q) & 1u) {
          const unsigned long long pos = base + s_wcnt[warp * PROBE_KPT + q] + rank[q];
          st_stream(out_keys + pos, key[q]);
          st_stream(out_vals + pos, val[q]);
          if (IDX) {
            const uint64_t row = vec ? tbase + 2ull * ((uint64_t)(q >> 1) * GT + gtid_l) + (q & 1)
                                     : tbase + (uint64_t)q * GT + gtid_l;
            st_stream(out_idx + pos, idx_base + row);
          }
        }
      }
      group_sync();  // s_wcnt / s_base are reused by the group's next tile
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
  if (lane == 0 && local_count) atomicAdd(&ctl->match_count, local_count);

  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(gsync + 2, 1u) == gridDim.x - 1) {
      gsync[0] = 0;
      gsync[1] = 0;
      gsync[2] = 0;
    }
  }
}

template <bool IDX, int THREADS>
static bool launch_mat_dense_fused_inst(const unsigned long long* bk, const unsigned long long* bv, uint64_t nb,
                                        const unsigned long long* pk, uint64_t np, uint32_t* bitmap, uint32_t dwords,
                                        unsigned long long* direct, Ctl* ctl, uint32_t* gsync, const ProbeOut& po,
                                        const DeviceInfo& di, cudaStream_t st) {
  auto kern = k_mat_dense_fused<IDX, THREADS>;
  const size_t smem = (size_t)dwords * 4;
  static size_t smem_set = 0;
  static int occ_cached = 0, dev_cached = -1;
  if (smem != smem_set || occ_cached == 0 || dev_cached != di.device) {
    dev_cached = di.device;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem) != cudaSuccess || occ < 1) return false;
    smem_set = smem;
    occ_cached = occ;
  }
  const uint64_t tile = (uint64_t)THREADS * PROBE_KPT;  // rows per CTA and round (NG group tiles of 256 * PROBE_KPT rows)
  uint64_t grid = (uint64_t)di.sms * occ_cached;  // all CTAs resident: the phases meet at spinning grid barriers
  const uint64_t ntiles = (np + tile - 1) / tile;
  if (grid > ntiles) grid = ntiles ? ntiles : 1;
  const int vec_ok = ((reinterpret_cast<uintptr_t>(pk) & 15u) == 0) ? 1 : 0;
  return launch_coop(kern, (unsigned)grid, THREADS, smem, st, bk, bv, nb, pk, np, bitmap, dwords, direct, ctl, gsync, po.keys, po.vals, po.idx,
                     po.idx_base, vec_ok);
}
bool launch_mat_dense_fused(const unsigned long long* bk, const unsigned long long* bv, uint64_t nb, const unsigned long long* pk,
                            uint64_t np, uint32_t* bitmap, uint32_t dwords, unsigned long long* direct, Ctl* ctl, uint32_t* gsync,
                            const ProbeOut& po, const DeviceInfo& di, cudaStream_t st, int* launches) {
  const bool two = (size_t)dwords * 4 * 2 + 8192 <= di.smem_optin;
  bool ok;
  if (po.idx) {
    ok = two ? launch_mat_dense_fused_inst<true, 512>(bk, bv, nb, pk, np, bitmap, dwords, direct, ctl, gsync, po, di, st)
             : launch_mat_dense_fused_inst<true, 1024>(bk, bv, nb, pk, np, bitmap, dwords, direct, ctl, gsync, po, di, st);
  } else {
    ok = two ? launch_mat_dense_fused_inst<false, 512>(bk, bv, nb, pk, np, bitmap, dwords, direct, ctl, gsync, po, di, st)
             : launch_mat_dense_fused_inst<false, 1024>(bk, bv, nb, pk, np, bitmap, dwords, direct, ctl, gsync, po, di, st);
  }
  if (ok) ++*launches;
  return ok;
}

size_t probe_smem_bitmap_limit_bytes(const DeviceInfo& di) {
  return di.smem_optin > 8192 ? ((di.smem_optin - 8192) / 16) * 16 : 0;
}

void launch_build_bitmap(uint32_t* bitmap, uint64_t dbits, const unsigned long long* bk, uint64_t nb, Ctl* ctl,
                         const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (nb == 0) return;
  const int threads = 256;
  const uint64_t want = (nb + threads - 1) / threads, cap = (uint64_t)di.sms * 16;
  k_build_bitmap<<<(int)(want < cap ? want : cap), threads, 0, st>>>(bitmap, dbits, bk, nb, ctl);
  ++*launches;
}

template <int THREADS>
static void launch_count_dense_inst(const unsigned long long* pk, uint64_t np, const uint32_t* bitmap, uint32_t dwords,
                                    Ctl* ctl, const DeviceInfo& di, cudaStream_t st) {
  auto kern = k_probe_count_dense<THREADS>;
  const size_t smem = (size_t)dwords * 4;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const uint64_t tile = (uint64_t)THREADS * PROBE_KPT;
  const uint64_t grid = persistent_grid(kern, THREADS, smem, (np + tile - 1) / tile, 0, di);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(pk) & 15u) == 0) ? 1 : 0;
  kern<<<(unsigned)grid, THREADS, smem, st>>>(pk, np, bitmap, dwords, ctl, vec_ok);
}
void launch_probe_count_dense(const unsigned long long* pk, uint64_t np, const uint32_t* bitmap, uint32_t dwords, Ctl* ctl,
                              const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (np == 0) return;
  // two 512-thread CTAs per SM while two bitmaps fit next to each other, one 1024-thread CTA otherwise
  if ((size_t)dwords * 4 * 2 + 4096 <= di.smem_optin) launch_count_dense_inst<512>(pk, np, bitmap, dwords, ctl, di, st);
  else launch_count_dense_inst<1024>(pk, np, bitmap, dwords, ctl, di, st);
  ++*launches;
}

template <bool NARROW, int BLOOM, int THREADS>
static void launch_count_inst(const TableView& t, const unsigned long long* pk, uint64_t np, int ctas_per_sm, Ctl* ctl,
                              const DeviceInfo& di, cudaStream_t st) {
  auto kern = k_probe_count<NARROW, BLOOM, THREADS>;
  size_t smem = 0;
  if (BLOOM == 1) smem += (size_t)t.bloom_words * 4;
  smem += (size_t)(THREADS / 32) * PROBE_QCAP * 8;
  if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const uint64_t tile = (uint64_t)THREADS * PROBE_KPT;
  const uint64_t grid = persistent_grid(kern, THREADS, smem, (np + tile - 1) / tile, BLOOM == 1 ? 0 : ctas_per_sm, di);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(pk) & 15u) == 0) ? 1 : 0;
  kern<<<(unsigned)grid, THREADS, smem, st>>>(pk, np, t.slots, t.nbuckets, t.bloom, t.bloom_words, ctl, vec_ok);
}

template <bool NARROW, int BLOOM, bool IDX, int THREADS>
static void launch_mat_inst(const TableView& t, const unsigned long long* pk, uint64_t np, const unsigned long long* bv,
                            const ProbeOut* out, int ctas_per_sm, Ctl* ctl, const DeviceInfo& di, cudaStream_t st) {
  auto kern = k_probe_mat<NARROW, BLOOM, IDX, THREADS>;
  size_t smem = 0;
  if (BLOOM == 1) {
    smem = (size_t)t.bloom_words * 4;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  const uint64_t tile = (uint64_t)THREADS * PROBE_KPT;
  const uint64_t grid = persistent_grid(kern, THREADS, smem, (np + tile - 1) / tile, BLOOM == 1 ? 0 : ctas_per_sm, di);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(pk) & 15u) == 0) ? 1 : 0;
  kern<<<(unsigned)grid, THREADS, smem, st>>>(pk, np, t.slots, t.nbuckets, t.bloom, t.bloom_words, bv, ctl, out->keys,
                                             out->vals, out->idx, out->idx_base, vec_ok);
}

void launch_probe(const TableView& t, const unsigned long long* pk, uint64_t np, const unsigned long long* bv,
                  const ProbeOut* out, bool bloom_in_smem, int ctas_per_sm, Ctl* ctl, const DeviceInfo& di,
                  cudaStream_t st, int* launches) {
  if (np == 0) return;
  const int bloom = t.bloom == nullptr ? 0 : (bloom_in_smem ? 1 : 2);
  const bool mat = out != nullptr;
  const bool idx = mat && out->idx != nullptr;
  // (shared-memory filter: ONE 1024-thread CTA per SM.  Two 512-thread CTAs with a filter of 8 bits per key instead of
  // 16 were measured at C2: 0.477 vs 0.185 ms — more false positives reach the table stage and every SM holds the
  // filter twice: profiles/r02D_exp_bloom.jsonl)
  if (!mat) {
#define FJ_CNT(N)                                                                       \
  do {                                                                                  \
    if (bloom == 0) launch_count_inst<N, 0, 512>(t, pk, np, ctas_per_sm, ctl, di, st);  \
    else if (bloom == 1) launch_count_inst<N, 1, 1024>(t, pk, np, ctas_per_sm, ctl, di, st); \
    else launch_count_inst<N, 2, 512>(t, pk, np, ctas_per_sm, ctl, di, st);             \
  } while (0)
    if (t.narrow) FJ_CNT(true); else FJ_CNT(false);
#undef FJ_CNT
  } else {
#define FJ_MAT(N, I)                                                                               \
  do {                                                                                             \
    if (bloom == 0) launch_mat_inst<N, 0, I, 256>(t, pk, np, bv, out, ctas_per_sm, ctl, di, st);   \
    else if (bloom == 1) launch_mat_inst<N, 1, I, 512>(t, pk, np, bv, out, ctas_per_sm, ctl, di, st); \
    else launch_mat_inst<N, 2, I, 256>(t, pk, np, bv, out, ctas_per_sm, ctl, di, st);              \
  } while (0)
    if (t.narrow) { if (idx) FJ_MAT(true, true); else FJ_MAT(true, false); }
    else { if (idx) FJ_MAT(false, true); else FJ_MAT(false, false); }
#undef FJ_MAT
  }
  ++*launches;
}

// =================================================================================== data generator G2
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull;
  x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull;
  x ^= x >> 33;
  return x;
}
// mirrors flash_hash_join_b200/datagen.py:g2_slice bit for bit
__global__ void __launch_bounds__(256) k_generate_g2(int side, uint64_t ny, uint64_t c, uint64_t U, uint64_t a, uint64_t b,
                                                     uint64_t seed, uint64_t start, uint64_t count,
                                                     unsigned long long* __restrict__ keys,
                                                     unsigned long long* __restrict__ vals) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  int kbits = 2;
  while (kbits < 64 && ((ny - 1) >> kbits) != 0) ++kbits;
  const uint64_t sh_mask = kbits >= 64 ? ~0ull : (1ull << kbits) - 1ull;
  const uint64_t sh_shift = (uint64_t)((kbits + 1) / 2);
  const uint64_t sh_add = (seed * 0x9E3779B97F4A7C15ull + 0x7F4A7C15ull) & sh_mask;
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < count; j += stride) {
    const uint64_t idx = start + j;
    uint64_t id;
    if (side == 0) {
      // build row idx holds build id jj = shuffle(idx): a bijective mixer on k bits, cycle-walked into [0, ny)
      // (datagen.py::_shuffle_index)
      uint64_t jj = idx;
      do {
        jj = (jj + sh_add) & sh_mask;
        jj ^= jj >> sh_shift;
        jj = (jj * 0xD6E8FEB86659FD93ull) & sh_mask;
        jj ^= jj >> sh_shift;
        jj = (jj * 0xCA5A826395121157ull) & sh_mask;
        jj ^= jj >> sh_shift;
      } while (jj >= ny);
      id = jj < c ? jj : jj - c + ny;
      if (vals) vals[j] = mix64(jj + seed * 0x9E3779B97F4A7C15ull) % 100ull;
    } else {
      const unsigned long long r = mix64((idx + 1ull) * 0x9E3779B97F4A7C15ull + seed);
      id = r % ny;
    }
    keys[j] = ((id * a) % U + b) % U + 1ull;
  }
}
void launch_generate_g2(int side, uint64_t ny, uint64_t c, uint64_t U, uint64_t a_mod_u, uint64_t b, uint64_t seed,
                        uint64_t start, uint64_t count, unsigned long long* keys, unsigned long long* vals,
                        cudaStream_t st) {
  if (count == 0) return;
  uint64_t want = (count + 255) / 256;
  const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
  k_generate_g2<<<grid, 256, 0, st>>>(side, ny, c, U, a_mod_u, b, seed, start, count, keys, vals);
}

}  // namespace fj
