// fj_part.cu — dense key domain, radix path, round 2: ONE high-fan-out partition pass per side with per-SM
// write-combining sector buffers in shared memory (k_part), then a direct-address join whose random accesses
// live in SHARED MEMORY (k_sjoin).
//
// Replaces, from /root/reference/hash_join.cpp (for build keys in a dense domain, e.g. the h2o ids 1..1.1*n):
//   get_partition_idx (:209), parallel_radix_partition_kv / _k (:210-292)                      -> k_part<VAL>
//   per-partition FlashHashTable ctor + build_local (:96-128, :191), probe_vectorized (:153-182) and the result
//   gather of _hash_join_radix_materialize / _count (:340-378, :515-531)                        -> k_sjoin<MAT>
//
// Why (profiles/r01i_c3_radix_ncu_summary.txt): the round-1 direct-address join (k_djoin) kept a partition's
// region in L2 and was bound by L2 tag lookups (2e8 random 4-byte accesses, lts 74 %, DRAM 33 %), and the
// 256-way scatter moved 8-byte build rows and 4-byte probe keys.  Here
//   * the digit is the low `logp` key bits (P = 2^logp <= 2048 partitions) and a row keeps only what the partition
//     does not imply: idx = key >> logp (16 bits).  Build row = idx | value << 16 (4 bytes, values < 65535, else the
//     attempt is abandoned), probe row = idx (2 bytes): the partition traffic of C3 (1e8 x 1e8) drops from
//     2.4 GB to 1.2 GB (written once, read once);
//   * a partition's direct-address region (2 bytes per key of its slice of the domain, <= 128 KB) is zeroed, filled
//     and probed in shared memory: no random access ever reaches L2 or HBM.
//
// k_part: persistent, one CTA of 16 independent warps per SM, no block barrier in the main loop.  Shared memory holds,
// for EVERY partition, a 64-byte ring of two 32-byte sectors and a tail | count word: a row claims a ticket with one
// shared-memory atomicAdd and stores its element into the slot the ticket names; the row that takes a sector's last
// ticket lists the sector; a listed sector leaves as a full, aligned 32-byte sector — in ticket order per partition,
// once no slot holds the hole marker any more — to a position reserved IN ADVANCE with a global atomicAdd (the next
// reservation is issued while this sector is stored, its answer is consumed a batch later).  Every warp streams its own
// batches through a private two-slot TMA (cp.async.bulk + mbarrier) input ring.  The protocol is described at the kernel.
//
// k_sjoin: persistent, one CTA per SM; warp 0 is the TMA producer, warps 1..31 consume.  The producer streams the
// chunks of [build rows of p][probe rows of p][build rows of p'] ... — from the local partition buffer or, in the
// multi-GPU shuffle, from every peer's buffer over NVLink — through a 3-deep ring with full/empty mbarriers, independent
// of the consumers' phase (fill | probe | clear), so HBM never idles at a phase change.  k_pairs_compact closes the
// tails of the output blocks; k_xsync carries the cross-GPU steps of the shuffle.
#include <type_traits>

#include "fj_kernels.h"

namespace fj {

// bounded spin on an mbarrier: a lost TMA transaction must never hang the GPU (trap -> the host sees an error)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ================================================================================= k_part
constexpr int PT_RING_BYTES = 65536;          // input rings of a CTA: 16 warps x 2 slots x 2 KB, or 32 warps x 2 slots x 1 KB
constexpr int PT_WSLOTS = 2;                  // input ring slots per warp
constexpr int PT_SECTOR = 32;                 // bytes per flush
constexpr int PT_RINGB = 2 * PT_SECTOR;       // bytes of staging per partition
constexpr int PT_MAXP = 2048;                 // flush-list entries keep the partition in 11 bits
constexpr int PT_MAXW = 8;                    // GPUs of a peer-memory shuffle (sources a partition is pulled from)
constexpr int PT_KEEP = 32;                   // flush-list entries a warp may carry into its next batch
constexpr int PT_LIST_BYTES = 20480;          // flush lists of a CTA: warps x (32 x rows per lane and batch + PT_KEEP) x 4 bytes
constexpr int PT_NQ = 4;                      // sector reservations in flight per lane pair
constexpr uint32_t PT_NOPLACE = 0xFFFFu;      // nextg: the reservation lies beyond the region
constexpr uint32_t PT_INFLIGHT = 0xFFFEu;     // nextg: the reservation has been issued, its answer is not published yet

struct PartParams {
  const unsigned long long* in_keys;
  const unsigned long long* in_vals;
  uint64_t n;
  uint32_t klimit;     // keys >= klimit are outside the domain (klimit <= 2^32 - 1)
  uint32_t cap;        // elements per (partition, sub-region); multiple of 16
  uint32_t* cursor;    // [P * cstride] elements reserved per partition (this source)
  uint32_t cstride;    // 32-bit words between two cursors (see fj_kernels.h: part_cursor_stride)
  Ctl* ctl;
  void* out;           // partition buffer: partition d occupies elements [d * cap, d * cap + cursor[d])
  int logp;            // log2(partitions)
};

__device__ __forceinline__ uint32_t lds_v32(const void* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_v16(const void* p) {
  uint16_t v;
  asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=h"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ uint4 lds_v128(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}
// does a 16-byte piece of a staged sector still hold an unwritten (all-ones) element?
template <bool VAL>
__device__ __forceinline__ bool piece_has_hole(const uint4& v) {
  if constexpr (VAL) {
    return max(max(v.x, v.y), max(v.z, v.w)) == 0xFFFFFFFFu;
  } else {
    const uint32_t m = __vmaxu2(__vmaxu2(v.x, v.y), __vmaxu2(v.z, v.w));
    return (m & 0xFFFFu) == 0xFFFFu || (m >> 16) == 0xFFFFu;
  }
}

// VAL: rows carry a value (element = idx | value << 16, 4 bytes) else keys only (element = idx, 2 bytes)
// STRICT: build side — a key >= klimit or a value > 65534 abandons the attempt (CTL_NOT_DENSE16); else such rows are dropped
//
// Staging protocol (no block barrier in the main loop; the 16 warps run independently):
//   * w[d] = tail << 16 | count.  tail = elements of partition d flushed so far (mod 2^16, a multiple of the sector size),
//     count = elements claimed and not flushed.  A row claims ticket T = tail + count with ONE shared-memory atomicAdd and
//     owns ring slot T mod 16|32 (two sectors).  count < two sectors: the slot is free, the row is stored at once;
//     otherwise the row keeps its ticket and is stored as soon as the tail has advanced far enough (`pending`).
//   * A free slot holds the all-ones hole marker (no valid element equals it).  The row whose ticket is the LAST of a
//     sector (the closer) puts (d, sector number) on its warp's flush list.  A listed sector leaves when it is the oldest
//     of its partition (tail == its first ticket: sectors of a partition leave in order), holds no hole marker any more
//     (every claimed row has been stored) and the partition's next global sector is known.  The flusher resets the
//     sector to hole markers, THEN advances the tail (one atomicAdd: tail += sector, count -= sector), and issues the
//     reservation of the partition's next global sector, whose answer is published a batch later.
//   * Entries that cannot leave yet stay on the list; a warp never blocks (on input data, on a pending row) without
//     servicing its list, so the oldest sector of every partition can always make progress.
// DIRECT: a warp's batches go from global memory straight into registers (128-bit streaming loads, the next batch in
// flight while the current one is placed and flushed) instead of through the warp's TMA input ring
template <bool VAL, bool STRICT, bool DIRECT>
__global__ void __launch_bounds__(512, 1) k_part(const PartParams a) {
  constexpr int NW = 16;  // (32 warps with 1 KB batches were slower on both sides: profiles/r02m_exp_part.jsonl)
  constexpr int PT_THREADS = NW * 32;
  constexpr int PT_WARPS = NW;
  constexpr int PT_SLOT_BYTES = PT_RING_BYTES / (NW * PT_WSLOTS);  // one batch of one warp
  constexpr int PT_WCAP = 32 * (PT_SLOT_BYTES / (VAL ? 16 : 8) / 32) + PT_KEEP;  // every row of a batch may complete a sector
  static_assert(NW * PT_WCAP * 4 <= PT_LIST_BYTES, "flush lists do not fit");
  using ET = std::conditional_t<VAL, uint32_t, uint16_t>;
  constexpr uint32_t ES = sizeof(ET);
  constexpr uint32_t EPS = PT_SECTOR / ES;            // elements per sector: 8 | 16
  constexpr uint32_t LOG_EPS = VAL ? 3 : 4;
  constexpr uint32_t SLOTS = 2 * EPS;
  constexpr uint32_t BROWS = PT_SLOT_BYTES / (VAL ? 16 : 8);  // rows per warp batch: 128 | 256
  constexpr int IPT = BROWS / 32;                     // rows per thread and batch: 4 | 8
  constexpr uint32_t FLUSH_ADD = (EPS << 16) - EPS;   // tail += EPS, count -= EPS (count >= EPS: no borrow)
  constexpr uint32_t SPIN_LIMIT = 1u << 22;           // a protocol error must trap, not hang the GPU
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t P = 1u << a.logp;
  unsigned char* buf = smem;                                               // P x 64 B: two sectors per partition
  uint32_t* w = reinterpret_cast<uint32_t*>(smem + (size_t)P * PT_RINGB);  // P: tail << 16 | count
  uint16_t* nextg = reinterpret_cast<uint16_t*>(w + P);                    // P: sector (in the sub-region) reserved for the next flush
  uint32_t* wl = reinterpret_cast<uint32_t*>(nextg + P) + (threadIdx.x >> 5) * PT_WCAP;  // this warp's flush list: d | sector number << 11
  // every warp streams its own batches through its own PT_WSLOTS x 2 KB input ring (a ring shared by the CTA is refilled
  // only when the slowest warp has read a stage: the warps starved, profiles/r02h_c3_dense16_ncu_summary.txt)
  unsigned char* ring = smem + (size_t)P * (PT_RINGB + 6) + PT_LIST_BYTES + (size_t)(threadIdx.x >> 5) * PT_WSLOTS * PT_SLOT_BYTES;
  __shared__ __align__(8) uint64_t s_full[PT_WARPS * PT_WSLOTS];

  const uint32_t tid = threadIdx.x, lane = tid & 31u, wv = tid >> 5, h = tid & 1u;
  // the build-side pass of this attempt found a row outside the domain: the attempt is abandoned, spare the probe side's pass
  if (*reinterpret_cast<volatile unsigned int*>(&a.ctl->flags) & CTL_NOT_DENSE16) return;
  const uint32_t GW = gridDim.x * PT_WARPS;          // warps of the grid
  const uint32_t gw = blockIdx.x * PT_WARPS + wv;    // this warp: batches gw, gw + GW, gw + 2 GW, ...
  const uint32_t nbatch = (uint32_t)((a.n + BROWS - 1) / BROWS);
  const bool aligned = ((reinterpret_cast<uintptr_t>(a.in_keys) | (VAL ? reinterpret_cast<uintptr_t>(a.in_vals) : 0)) & 15u) == 0;
  const uint32_t nfull = aligned ? (uint32_t)(a.n / BROWS) : 0u;  // batches below nfull arrive through the TMA ring
  uint64_t* const my_full = s_full + wv * PT_WSLOTS;
  const uint32_t capsec = a.cap >> LOG_EPS;
  const uint32_t pmask = P - 1u;

  if (!DIRECT && tid == 0) {
    for (int s = 0; s < PT_WARPS * PT_WSLOTS; ++s) mbar_init(&s_full[s], 1);
    mbar_fence_init();
  }
  // reservation -> sector index inside the (partition, source) sub-region; PT_NOPLACE when it does not fit
  auto to_sector = [&](uint32_t g) -> uint16_t { return g + EPS <= a.cap ? (uint16_t)(g >> LOG_EPS) : (uint16_t)PT_NOPLACE; };
  // every slot free; every (CTA, partition) holds one sector reserved in advance: sector blockIdx.x to begin with (the
  // cursors start at gridDim.x sectors, part_cursor_start: no reservation round trips in the prologue)
  for (uint32_t i = tid; i < P * (PT_RINGB / 16); i += PT_THREADS) reinterpret_cast<uint4*>(buf)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
  for (uint32_t d = tid; d < P; d += PT_THREADS) {
    w[d] = 0;
    nextg[d] = to_sector(blockIdx.x * EPS);
  }
  __syncthreads();

  // this warp's batch number kk (batch gw + kk * GW of the input) -> ring slot kk % PT_WSLOTS
  auto issue = [&](uint32_t kk) {  // one lane
    const uint32_t B = gw + kk * GW;
    if (B >= nfull) return;
    const uint32_t s = kk % PT_WSLOTS;
    const uint64_t row0 = (uint64_t)B * BROWS;
    unsigned char* dst = ring + s * PT_SLOT_BYTES;
    mbar_expect_tx(&my_full[s], PT_SLOT_BYTES);
    bulk_g2s(dst, a.in_keys + row0, BROWS * 8u, &my_full[s]);
    if constexpr (VAL) bulk_g2s(dst + BROWS * 8u, a.in_vals + row0, BROWS * 8u, &my_full[s]);
  };
  if (!DIRECT && lane == 0) {
#pragma unroll
    for (uint32_t kk = 0; kk < PT_WSLOTS; ++kk) issue(kk);
  }

  unsigned char* const out0 = static_cast<unsigned char*>(a.out);
  // half hh (16 bytes) of one 32-byte sector of partition d -> sector gs of its sub-region in the owner's buffer.  Two
  // lanes share a sector, so a warp moves 16 sectors with ONE 128-bit store.
  auto store_half = [&](uint32_t d, uint32_t gs, uint32_t hh, const uint4& v) {
    if (gs == PT_NOPLACE) {
      if (hh == 0) atomicOr(&a.ctl->flags, CTL_OVERFLOW);
      return;
    }
    unsigned char* dst = out0 + (((uint64_t)d * capsec + gs) << 5);
    reinterpret_cast<uint4*>(dst)[hh] = v;
  };

  constexpr uint32_t NOROW = 0xFFFFFFFFu;  // element of a row that is dropped (no valid element equals it)
  struct Rows {
    uint32_t d[IPT];
    uint32_t e[IPT];
    uint32_t pend;  // rows that hold a ticket and are not stored yet
    bool inv;       // some row of the batch is dropped (outside the domain, or beyond the end of the input)
  };
  bool bad = false;
  // digit, element and validity of one row
  auto decode = [&](uint32_t klo, uint32_t khi, uint32_t vlo, uint32_t vhi, bool ok, int i, Rows& r) {
    bool in = (khi == 0u) & (klo < a.klimit);
    if constexpr (VAL) in &= (vhi == 0u) & (vlo <= 65534u);
    if constexpr (STRICT) bad |= ok & !in;
    ok &= in;
    r.d[i] = klo & pmask;
    r.e[i] = ok ? ((klo >> a.logp) | (VAL ? vlo << 16 : 0u)) : NOROW;
    r.inv |= !ok;
  };
  // has batch B (this warp's batch number kk) arrived?  (warp-uniform)
  auto batch_ready = [&](uint32_t B, uint32_t kk) -> bool {
    if (DIRECT || B >= nfull) return true;
    return __all_sync(0xffffffffu, mbar_try_wait(&my_full[kk % PT_WSLOTS], (kk / PT_WSLOTS) & 1u));
  };
  // the rows of batch B.  Full batches of 16-byte aligned inputs come out of the warp's ring slot (which must be ready):
  // a lane reads pairs of adjacent rows with 128-bit loads; the slot is refilled as soon as it has been read.  The
  // ragged tail / unaligned inputs are loaded directly.
  // DIRECT: the loads of a full batch, issued one batch ahead of their use
  struct Raw {
    unsigned long long k[IPT], v[VAL ? IPT : 1];
  };
  auto issue_raw = [&](uint32_t B, Raw& x) {
    if (B >= nfull) return;
    const unsigned long long* kp = a.in_keys + (uint64_t)B * BROWS + 2u * lane;
#pragma unroll
    for (int q = 0; q < IPT / 2; ++q) ld_stream2(kp + q * 64, x.k[2 * q], x.k[2 * q + 1]);
    if constexpr (VAL) {
      const unsigned long long* vp = a.in_vals + (uint64_t)B * BROWS + 2u * lane;
#pragma unroll
      for (int q = 0; q < IPT / 2; ++q) ld_stream2(vp + q * 64, x.v[2 * q], x.v[2 * q + 1]);
    }
  };
  auto load_rows = [&](uint32_t B, uint32_t kk, Rows& r, const Raw* x = nullptr) {
    r.pend = 0;
    r.inv = false;
    if (DIRECT && B < nfull) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const unsigned long long v64 = VAL ? x->v[i] : 0ull;
        decode((uint32_t)x->k[i], (uint32_t)(x->k[i] >> 32), (uint32_t)v64, (uint32_t)(v64 >> 32), true, i, r);
      }
    } else if (B < nfull) {
      const uint4* st = reinterpret_cast<const uint4*>(ring + (kk % PT_WSLOTS) * PT_SLOT_BYTES);
      uint4 k2[IPT / 2], v2[IPT / 2];
#pragma unroll
      for (int q = 0; q < IPT / 2; ++q) {
        k2[q] = st[q * 32 + lane];
        v2[q] = VAL ? st[BROWS / 2 + q * 32 + lane] : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int q = 0; q < IPT / 2; ++q) {
        decode(k2[q].x, k2[q].y, v2[q].x, v2[q].y, true, 2 * q, r);
        decode(k2[q].z, k2[q].w, v2[q].z, v2[q].w, true, 2 * q + 1, r);
      }
      __syncwarp();  // every lane's loads have returned
      if (lane == 0) issue(kk + PT_WSLOTS);
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) {
        const uint64_t row = (uint64_t)B * BROWS + (uint64_t)i * 32u + lane;
        const bool ok = row < a.n;
        unsigned long long k64 = ~0ull, v64 = 0;
        if (ok) {
          k64 = ld_stream1(a.in_keys + row);
          if constexpr (VAL) v64 = ld_stream1(a.in_vals + row);
        }
        decode((uint32_t)k64, (uint32_t)(k64 >> 32), (uint32_t)v64, (uint32_t)(v64 >> 32), ok, i, r);
      }
    }
  };

  uint32_t pd[PT_NQ], pg[PT_NQ];  // reservations in flight: partition, reserved element offset
  uint32_t pvalid = 0;
  uint32_t wn = 0;                // entries on this warp's flush list (warp-uniform)

  // ---- place: one shared-memory atomic hands out the ticket
  // ALLV (a literal at both call sites): no row of the warp's batch is dropped, no per-row predicate
  auto place = [&](const bool ALLV, Rows& r, uint32_t (&tk)[IPT]) {
    uint32_t old[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i)  // a dropped row adds 0 (a branch around every ATOMS otherwise)
      old[i] = atomicAdd(&w[r.d[i]], ALLV ? 1u : (r.e[i] != NOROW ? 1u : 0u));
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const bool v = ALLV || r.e[i] != NOROW;
      const uint32_t cnt = old[i] & 0xFFFFu;
      const uint32_t T = ((old[i] >> 16) + cnt) & 0xFFFFu;
      tk[i] = T;
      const bool room = cnt < SLOTS;  // (one predicated store and a select: nested branches cost 11 instructions per row here)
      if (v && room) reinterpret_cast<ET*>(buf)[r.d[i] * SLOTS + (T & (SLOTS - 1u))] = (ET)r.e[i];
      r.pend |= (v && !room) ? (1u << i) : 0u;
      const bool closer = v && (T & (EPS - 1u)) == EPS - 1u;
      const unsigned m = __ballot_sync(0xffffffffu, closer);
      if (closer) wl[wn + __popc(m & lanemask_lt())] = r.d[i] | ((T >> LOG_EPS) << 11);
      wn += __popc(m);
    }
  };
  // rows that found their slot occupied: has the tail advanced far enough?
  auto retry = [&](Rows& r, const uint32_t (&tk)[IPT]) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      if ((r.pend >> i) & 1u) {
        const uint32_t tail = lds_v32(&w[r.d[i]]) >> 16;
        if (((tk[i] - tail) & 0xFFFFu) < SLOTS) {
          __threadfence_block();  // the flusher's reset of the slot is ordered before its tail update
          reinterpret_cast<ET*>(buf)[r.d[i] * SLOTS + (tk[i] & (SLOTS - 1u))] = (ET)r.e[i];
          r.pend &= ~(1u << i);
        }
      }
    }
  };
  // the reservations issued by the previous flush have returned by now: publish them
  auto publish = [&]() {
#pragma unroll
    for (int q = 0; q < PT_NQ; ++q)
      if ((pvalid >> q) & 1u) nextg[pd[q]] = to_sector(pg[q]);
    pvalid = 0;
  };
  // sixteen list entries, a pair of lanes each (one half of the sector per lane)
  auto flush_step = [&](const int slot, const uint32_t q, uint32_t& keep, const uint32_t nvis) {
    const uint32_t j = q * 16u + (lane >> 1);
    const bool act = j < nvis;
    const unsigned am = __ballot_sync(0xffffffffu, act);
    const uint32_t ent = act ? wl[j] : 0u;
    bool ok = false;
    if (act) {
      const uint32_t d = ent & (PT_MAXP - 1u), S = ent >> 11;
      unsigned char* src = buf + d * PT_RINGB + (S & 1u) * PT_SECTOR + h * 16u;
      const uint32_t ww = lds_v32(&w[d]);
      const uint32_t gs = lds_v16(&nextg[d]);
      const uint4 v = lds_v128(src);
      bool hole = piece_has_hole<VAL>(v);
      hole |= __shfl_xor_sync(am, (int)hole, 1) != 0;
      ok = ((ww >> (16 + LOG_EPS)) == S) && !hole && gs != PT_INFLIGHT;
      if (ok) {
        store_half(d, gs, h, v);
        *reinterpret_cast<uint4*>(src) = make_uint4(~0u, ~0u, ~0u, ~0u);
        if (h == 0) {
          nextg[d] = (uint16_t)PT_INFLIGHT;
          __threadfence_block();  // both halves are free again before the tail moves
          atomicAdd(&w[d], FLUSH_ADD);
          // The reservation's round trip through L2 (1 - 2 us under load) is never waited for here: the atomic writes
          // straight into the register that publish() reads a batch later
          asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(pg[slot]) : "l"(a.cursor + d * a.cstride), "r"(EPS) : "memory");
          pd[slot] = d;
          pvalid |= 1u << slot;
        }
      }
    }
    const bool fail = act && !ok && h == 0;
    const unsigned m = __ballot_sync(0xffffffffu, fail);
    if (fail) wl[keep + __popc(m & lanemask_lt())] = ent;  // positions below the entries still to be read
    keep += __popc(m);
  };
  // all == false: only whole steps of sixteen entries (the last few entries of the list wait for the next batch: a
  // step costs the same for one entry as for sixteen)
  auto flush = [&](bool all) {
    __syncwarp();
    publish();
    uint32_t keep = 0;
    const uint32_t nsteps = all ? (wn + 15u) >> 4 : wn >> 4;
    const uint32_t done = nsteps << 4 < wn ? nsteps << 4 : wn;  // entries visited
    for (uint32_t q0 = 0; q0 < nsteps; q0 += PT_NQ) {
      if (q0) publish();
#pragma unroll
      for (int s = 0; s < PT_NQ; ++s)
        if (q0 + s < nsteps) flush_step(s, q0 + s, keep, done);
    }
    // the entries not visited move down behind the ones that stay
    if (done < wn) {
      if (keep < done) {
        for (uint32_t j = done + lane; j < wn; j += 32u) wl[keep + (j - done)] = wl[j];  // fewer than 16: one pass, reads before writes
        __syncwarp();
      }
      keep += wn - done;
    }
    wn = keep;
  };

  Rows cur, nxt;
  uint32_t tk[IPT];
  uint32_t k = 0;
  uint32_t B = gw;
  Raw raw;
  if (B < nbatch) {
    if constexpr (DIRECT) issue_raw(B, raw);
    for (uint32_t spin = 0; !batch_ready(B, 0); ++spin)
      if (spin > (1u << 24)) __trap();
    load_rows(B, 0, cur, &raw);
  }
  while (B < nbatch) {
    const uint32_t Bn = B + GW;
    if constexpr (DIRECT) {
      if (Bn < nbatch) issue_raw(Bn, raw);  // in flight while this batch is placed and flushed
    }
    if (__any_sync(0xffffffffu, cur.inv)) place(false, cur, tk);
    else place(true, cur, tk);
    // the next batch is fetched and decoded before the flush when it has arrived already
    bool have = false;
    if (!DIRECT && Bn < nbatch && batch_ready(Bn, k + 1)) {
      load_rows(Bn, k + 1, nxt);
      have = true;
    }
    bool anyp = __any_sync(0xffffffffu, cur.pend != 0u);
    flush(anyp);
    for (uint32_t spin = 0; anyp || wn > (uint32_t)PT_KEEP; ++spin) {  // rare: a ring was full, or many sectors cannot leave yet
      retry(cur, tk);
      flush(true);
      anyp = __any_sync(0xffffffffu, cur.pend != 0u);
      if (spin > SPIN_LIMIT) __trap();
    }
    if (Bn < nbatch && !have) {
      for (uint32_t spin = 0; !batch_ready(Bn, k + 1); ++spin) {
        if (wn) flush(true);  // never wait for data while other warps may wait for a sector on this list
        if (spin > SPIN_LIMIT) __trap();
      }
      load_rows(Bn, k + 1, nxt, &raw);
    }
    cur = nxt;
    B = Bn;
    ++k;
  }
  for (uint32_t spin = 0; wn; ++spin) {
    flush(true);
    if (spin > SPIN_LIMIT) __trap();
  }
  publish();
  __syncthreads();

  // ---- drain: every partition's oldest sector (partially filled, the rest hole markers) goes into the sector held in reserve
  for (uint32_t d = tid; d < P; d += PT_THREADS) {
    const uint32_t sec = (w[d] >> (16 + LOG_EPS)) & 1u;
    const uint4* src = reinterpret_cast<const uint4*>(buf + d * PT_RINGB + sec * PT_SECTOR);
    const uint32_t gs = nextg[d];
    store_half(d, gs, 0u, src[0]);
    store_half(d, gs, 1u, src[1]);
  }
  if constexpr (STRICT) {
    if (bad) atomicOr(&a.ctl->flags, CTL_NOT_DENSE16);
  }
}

size_t part_smem_bytes(int logp) {
  return ((size_t)1 << logp) * (PT_RINGB + 6) + PT_LIST_BYTES + PT_RING_BYTES;
}
uint32_t part_sector_elems(bool val) { return val ? 8u : 16u; }
// One cursor per 256 bytes: 2048 adjacent 4-byte cursors live in 64 cache lines, i.e. on a handful of L2 slices, and the
// 1.2e7 reservation atomics of a 1e8-row pass saturated exactly those (lts__throughput max 105 %, avg 33 %;
// lts__d_atomic_input_cycles_active max 58 %: profiles/r02e_c3_dense16_ncu_summary.txt)
uint32_t part_cursor_stride() { return 64u; }
uint32_t part_grid(bool val, uint64_t n, const DeviceInfo& di) {
  const uint64_t round = val ? 2048 : 4096;  // one batch per warp
  const uint64_t rounds = (n + round - 1) / round;
  return (uint32_t)(rounds < (uint64_t)di.sms ? (rounds ? rounds : 1) : (uint64_t)di.sms);
}

uint32_t part_cursor_start(bool val, uint64_t n, const DeviceInfo& di) { return n ? part_grid(val, n, di) * part_sector_elems(val) : 0u; }

bool launch_part(bool val, const PartArgs& x, const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (x.logp < 4 || (1 << x.logp) > PT_MAXP || x.n == 0) return false;
  if (x.klimit > 0xFFFFFFFFull || x.cap > 0xFFFFFFF0ull || (x.cap & 15u) || (x.cap >> (val ? 3 : 4)) >= (uint64_t)PT_INFLIGHT) return false;
  if ((x.n + 127) / 128 > 0xFFFFFFF0ull) return false;
  PartParams a;
  a.in_keys = x.in_keys; a.in_vals = x.in_vals; a.n = x.n; a.klimit = (uint32_t)x.klimit; a.cap = (uint32_t)x.cap; a.cursor = x.cursor; a.cstride = x.cursor_stride;
  a.ctl = x.ctl; a.out = x.out; a.logp = x.logp;
  const size_t smem = part_smem_bytes(x.logp);
  if (smem + 256 > di.smem_optin) return false;
  const uint32_t grid = part_grid(val, x.n, di);
#define FJ_PART4(V, S, D)                                                                             \
  do {                                                                                                \
    static size_t smem_set = 0; /* the attribute call costs microseconds: once per size and device */ \
    static int dev_set = -1;                                                                          \
    if (smem_set != smem || dev_set != di.device) {                                                   \
      cudaFuncSetAttribute(k_part<V, S, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      smem_set = smem;                                                                                \
      dev_set = di.device;                                                                            \
    }                                                                                                 \
    k_part<V, S, D><<<grid, 512, smem, st>>>(a);                                                     \
  } while (0)
#define FJ_PART(V, S) do { if (x.direct_in) FJ_PART4(V, S, true); else FJ_PART4(V, S, false); } while (0)
  if (val) FJ_PART(true, true);  // rows with values are a build side
  else if (x.strict) FJ_PART(false, true);
  else FJ_PART(false, false);
#undef FJ_PART
#undef FJ_PART4
  ++*launches;
  return true;
}

// ================================================================================= k_sel_sample
// Match-rate sample ahead of a dense16 attempt.  With a small build side the dense table path (bitmap in shared memory,
// values in an L2-resident table) only touches the probe rows that hit, while dense16 partitions every probe row: below
// ~50 % match rate the table path wins, above it dense16 does (profiles/r02N_exp_selectivity.jsonl).  The rate is unknown
// up front, so the adaptive materialize measures it: every CTA clears the membership bits of its slice of the build keys
// in a global bitmap (all-ones = empty, prepared by k_prepare together with the two words behind it), and the last CTA to
// arrive tests SS_SAMPLES probe keys spread evenly (with a hashed offset inside each stride) over the probe side, all of
// a thread's loads in flight at once.  A rate below min_pct raises CTL_NOT_DENSE16 | CTL_LOW_SEL — unless a build key
// lies outside the table path's own domain (table_bits), where that path could not answer: the k_part / k_sjoin launches
// queued behind return at once and the host takes the table path.  A few microseconds in front of a 0.6 ms join.
constexpr int SS_THREADS = 512;
constexpr int SS_PER_THREAD = 16;
constexpr uint32_t SS_SAMPLES = SS_THREADS * SS_PER_THREAD;
__global__ void __launch_bounds__(SS_THREADS) k_sel_sample(Ctl* __restrict__ ctl, const unsigned long long* __restrict__ bk, uint64_t nb,
                                                           const unsigned long long* __restrict__ pk, uint64_t np,
                                                           uint32_t* __restrict__ bitmap, uint64_t bits, uint64_t table_bits,
                                                           uint32_t* __restrict__ words, uint32_t min_pct) {
  __shared__ uint32_t s_last, s_hits;
  const uint64_t stride = (uint64_t)gridDim.x * SS_THREADS;
  bool outside = false;
  for (uint64_t i = blockIdx.x * (uint64_t)SS_THREADS + threadIdx.x; i < nb; i += 4 * stride) {  // four loads in flight
    unsigned long long k[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) k[r] = i + r * stride < nb ? bk[i + r * stride] : ~0ull;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (k[r] < bits) atomicAnd(&bitmap[k[r] >> 5], ~(1u << (k[r] & 31)));
      outside |= (k[r] != ~0ull) & (k[r] >= table_bits);
    }
  }
  if (__syncthreads_or(outside) && threadIdx.x == 0) atomicAnd(&words[1], 0u);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    s_last = (atomicAdd(&words[0], 1u) + 1u) == gridDim.x - 1u;  // the arrival counter starts at 0xFFFFFFFF
    s_hits = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const uint32_t ns = (uint32_t)(np < SS_SAMPLES ? np : SS_SAMPLES);
  const uint64_t step = np / ns;
  unsigned long long k[SS_PER_THREAD];
#pragma unroll
  for (int r = 0; r < SS_PER_THREAD; ++r) {
    const uint32_t j = r * SS_THREADS + threadIdx.x;
    k[r] = j < ns ? __ldg(&pk[j * step + (hash32(j) % step)]) : ~0ull;
  }
  uint32_t w[SS_PER_THREAD];
#pragma unroll
  for (int r = 0; r < SS_PER_THREAD; ++r) w[r] = k[r] < bits ? __ldcg(&bitmap[k[r] >> 5]) : 0xFFFFFFFFu;
  uint32_t hits = 0;
#pragma unroll
  for (int r = 0; r < SS_PER_THREAD; ++r) hits += ((w[r] >> (k[r] & 31)) & 1u) ^ 1u;
  hits = __reduce_add_sync(0xffffffffu, hits);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_hits, hits);
  __syncthreads();
  if (threadIdx.x == 0 && __ldcg(&words[1]) != 0u && (uint64_t)s_hits * 100u < (uint64_t)min_pct * ns)
    atomicOr(&ctl->flags, CTL_NOT_DENSE16 | CTL_LOW_SEL);
}
size_t sel_sample_bytes(uint64_t bits) { return (size_t)((bits + 127) / 128 * 16 + 16); }
void launch_sel_sample(Ctl* ctl, const unsigned long long* bk, uint64_t nb, const unsigned long long* pk, uint64_t np, void* area,
                       uint64_t bits, uint64_t table_bits, uint32_t min_pct, const DeviceInfo& di, cudaStream_t st, int* launches) {
  uint32_t* bitmap = static_cast<uint32_t*>(area);
  uint32_t* words = bitmap + (bits + 127) / 128 * 4;  // arrival counter, "every build key inside the table domain"
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(2ull * di.sms, (nb + SS_THREADS - 1) / SS_THREADS));
  k_sel_sample<<<grid, SS_THREADS, 0, st>>>(ctl, bk, nb, pk, np, bitmap, bits, table_bits, words, min_pct);
  if (launches) ++*launches;
}

// ================================================================================= k_xsync
// Multi-GPU shuffle over peer memory (one process per GPU, every rank's exchange area mapped by every other rank through
// CUDA IPC): the cross-GPU steps around k_part and k_sjoin, each ONE small launch on the rank's own stream.
//   phase 1  count push + size check + barrier: my reservation cursors of the partitions a peer owns go into that
//            peer's count array (k_sjoin's bcnt / pcnt, indexed [source][local partition]) and (nb, np) of my slice
//            into every peer's area; after the barrier every rank's partition pass is complete and visible, so k_sjoin
//            may pull partition rows from every peer, and every rank has compared the sizes with the ones the plan
//            (capacities, partition count) was made for (CTL_META_CHANGED: every rank sees the same vector, so every
//            rank abandons the step and re-plans; the partition pass of a stale plan only ever wrote local memory)
//   phase 2  result exchange: (matches, flags, nb, np) of every rank -> sum / or on every rank (replaces
//            ncclAllReduce).  It is also the barrier after which no rank reads any partition buffer of this step any
//            more, so the next step's partition pass needs no entry barrier.
// Barrier words carry a sequence number (3 * step + phase + 1) and are never reset.  Spins give up after 10 s.
constexpr int XS_BAR = 0;          // + rank: barrier sequence number posted by `rank`
constexpr int XS_META = 16;        // + 2 * rank: nb, np of `rank` (phase 0)
constexpr int XS_RED = 64;         // + 4 * rank: matches, flags, nb, np of `rank` (phase 2)
constexpr int XS_CNT = 128;        // 64-bit word offset of the count arrays: uint32 [2 sides][world][P / world]
__device__ __forceinline__ unsigned long long xs_ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long xs_ld_relaxed(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void xs_st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void xs_st_relaxed(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool xs_wait_ge(const unsigned long long* p, unsigned long long want) {
  if (xs_ld_acquire(p) >= want) return true;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    if (xs_ld_acquire(p) >= want) return true;
    __nanosleep(100);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) return false;
  }
}
struct XsyncParams {
  unsigned long long* ctrl[PT_MAXW];  // every rank's exchange area (ctrl[rank] is the local one)
  int rank, world, phase;
  unsigned long long seq;
  Ctl* ctl;
  unsigned long long nb, np;              // this rank's slice
  unsigned long long meta[2 * PT_MAXW];   // phase 0: the sizes the plan was made for
  const uint32_t* cur_b;                  // phase 1: this rank's reservation cursors
  const uint32_t* cur_p;
  uint32_t cstride, P, lpo;
  unsigned long long* result;             // phase 2: [4] matches, flags, nb, np over all ranks (local memory)
};
__global__ void __launch_bounds__(1024) k_xsync(const XsyncParams a) {
  const int tid = threadIdx.x;
  const int W = a.world;
  unsigned long long* const mine = a.ctrl[a.rank];
  if (a.phase == 1) {
    if (tid < W) {
      xs_st_relaxed(a.ctrl[tid] + XS_META + 2 * a.rank, a.nb);
      xs_st_relaxed(a.ctrl[tid] + XS_META + 2 * a.rank + 1, a.np);
    }
    const uint32_t ppo = 1u << a.lpo;  // partitions per owner
    for (uint32_t d = tid; d < a.P; d += blockDim.x) {
      uint32_t* cnt = reinterpret_cast<uint32_t*>(a.ctrl[d >> a.lpo] + XS_CNT);
      const uint32_t at = (uint32_t)a.rank * ppo + (d & (ppo - 1u));
      cnt[at] = a.cur_b[(size_t)d * a.cstride];
      cnt[(uint32_t)W * ppo + at] = a.cur_p[(size_t)d * a.cstride];
    }
  } else {
    if (tid < W) {
      unsigned long long* slot = a.ctrl[tid] + XS_RED + 4 * a.rank;
      xs_st_relaxed(slot, a.ctl->match_count);
      xs_st_relaxed(slot + 1, (unsigned long long)a.ctl->flags);
      xs_st_relaxed(slot + 2, a.nb);
      xs_st_relaxed(slot + 3, a.np);
    }
  }
  __threadfence_system();  // my stores (and the partition rows of the kernels before this one) before the signal
  __syncthreads();
  __shared__ int s_ok;
  if (tid == 0) s_ok = 1;
  __syncthreads();
  if (tid < W) {
    xs_st_release(a.ctrl[tid] + XS_BAR + a.rank, a.seq);
    if (!xs_wait_ge(mine + XS_BAR + tid, a.seq)) s_ok = 0;
  }
  __syncthreads();
  if (tid == 0) {
    if (!s_ok) {
      atomicOr(&a.ctl->flags, CTL_PEER_TIMEOUT);
    } else if (a.phase == 1) {
      bool same = true;
      for (int r = 0; r < W; ++r)
        same &= xs_ld_relaxed(mine + XS_META + 2 * r) == a.meta[2 * r] && xs_ld_relaxed(mine + XS_META + 2 * r + 1) == a.meta[2 * r + 1];
      if (!same) atomicOr(&a.ctl->flags, CTL_META_CHANGED);
    } else if (a.phase == 2) {
      unsigned long long m = 0, f = 0, nb = 0, np = 0;
      for (int r = 0; r < W; ++r) {
        const unsigned long long* slot = mine + XS_RED + 4 * r;
        m += xs_ld_relaxed(slot);
        f |= xs_ld_relaxed(slot + 1);
        nb += xs_ld_relaxed(slot + 2);
        np += xs_ld_relaxed(slot + 3);
      }
      a.result[0] = m;
      a.result[1] = f;
      a.result[2] = nb;
      a.result[3] = np;
    }
  }
}
size_t xsync_ctrl_bytes(uint32_t P) { return (size_t)XS_CNT * 8 + (size_t)2 * P * 4 + 256; }
size_t xsync_count_offset_bytes() { return (size_t)XS_CNT * 8; }
void launch_xsync(const XsyncArgs& x, cudaStream_t st, int* launches) {
  XsyncParams a;
  for (int i = 0; i < PT_MAXW; ++i) a.ctrl[i] = i < x.world ? static_cast<unsigned long long*>(x.ctrl[i]) : nullptr;
  a.rank = x.rank; a.world = x.world; a.phase = x.phase; a.seq = x.seq; a.ctl = x.ctl; a.nb = x.nb; a.np = x.np;
  for (int i = 0; i < 2 * PT_MAXW; ++i) a.meta[i] = x.meta[i];
  a.cur_b = x.cur_b; a.cur_p = x.cur_p; a.cstride = x.cursor_stride; a.P = x.P; a.lpo = x.lpo; a.result = x.result;
  k_xsync<<<1, x.phase == 1 ? 1024 : 32, 0, st>>>(a);
  ++*launches;
}

// ================================================================================= k_sjoin
constexpr int SJ_THREADS = 1024;
constexpr int SJ_CONS = SJ_THREADS - 32;  // consumer threads (warps 1..31)
constexpr int SJ_CWARPS = SJ_CONS / 32;
constexpr int SJ_NPC = 2;                 // 16-byte pieces per consumer thread and chunk
constexpr int SJ_CH = SJ_CONS * 16 * SJ_NPC;  // bytes per ring stage
constexpr int SJ_STAGES = 3;
constexpr uint32_t SJ_SLOTS = 65536;      // direct-address slots: every 16-bit index has one (0xFFFF = hole: never set)
constexpr int SJ_LOG_BLK = 11;            // pairs per output block
constexpr uint32_t SJ_BLK = 1u << SJ_LOG_BLK;
constexpr int SJ_NBLK = 8;                // output-block table entries (blocks b - 6 .. b + 1 around the newest one)
constexpr int SJ_TAIL_WORDS = 4;          // per-CTA record for k_pairs_compact: partial block base, pairs in it, unused block base, -

struct SjoinParams {
  const unsigned char* build[PT_MAXW];  // per source: partition p at elements [p * cap_b, +count) (MAT: 4 bytes idx | value << 16; count: 2 bytes idx)
  const uint32_t* bcnt;        // elements written: bcnt[sub * cnt_stride + p]
  uint64_t cap_b;
  const unsigned char* probe[PT_MAXW];  // per source: partition p at 2-byte elements [p * cap_p, +count)
  const uint32_t* pcnt;
  uint64_t cap_p;
  uint32_t cnt_stride, cstride;  // cursor of (sub, p) = cnt[(sub * cnt_stride + p) * cstride]
  uint32_t p_first, p_count;   // partitions joined here: global ids p_first .. p_first + p_count - 1
  int logp, nsub;
  uint32_t rot;                // rotation of the source order (this GPU's rank + 1)
  uint32_t slots;              // slots a build row can address (multiple of 8): only these are cleared between partitions
  Ctl* ctl;
  unsigned long long* out_keys;
  unsigned long long* out_vals;
  unsigned long long* tails;   // MAT: [gridDim.x * SJ_TAIL_WORDS]
};

__device__ __forceinline__ void sj_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"r"(SJ_CONS) : "memory"); }
__device__ __forceinline__ unsigned long long lds_v64(const void* p) {
  unsigned long long v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void sts_v64(void* p, unsigned long long v) {
  asm volatile("st.volatile.shared.u64 [%0], %1;" ::"r"(smem_u32(p)), "l"(v) : "memory");
}
// sum of the two 16-bit halves of x, added to acc (one IDP.2A)
__device__ __forceinline__ uint32_t add_halves(uint32_t x, uint32_t acc) {
  uint32_t r;
  asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0x0101u), "r"(acc));
  return r;
}

// One partition at a time per CTA: fill the direct-address region (plain 16-bit stores: region[idx] = value + 1), stream
// the probe rows ONCE, then clear the region again.
//   * Duplicate build keys are found without atomics: every row contributes value + 1 >= 1 to a row sum, every slot its
//     content to a slot sum when the region is cleared; a key stored twice loses one contribution (CTL_DUP).
//   * MAT: pairs go to output blocks of SJ_BLK pairs.  A CTA numbers its pairs with a shared-memory cursor; the warp
//     that is first to enter block b reserves block b + 1 with ONE global atomic (2.2e4 for 9e7 pairs; a reservation
//     per warp serialised in L2, a reservation per partition needed a second pass over the probe rows:
//     profiles/r02a / r02i _c3_dense16_ncu_summary.txt).  The unused tail of every CTA's last block (and its block
//     reserved ahead) is a hole that k_pairs_compact fills with the pairs beyond the final count.
//   * The region has a slot for EVERY 16-bit index, so a lookup needs no bounds check: slot 0xFFFF (the hole marker of
//     the partition buffers) and the slots beyond the build side's domain stay zero.
template <bool MAT>
__global__ void __launch_bounds__(SJ_THREADS, 1) k_sjoin(const SjoinParams a) {
  constexpr uint32_t EB = MAT ? 4u : 2u;  // bytes per build element
  extern __shared__ __align__(128) unsigned char smem[];
  uint16_t* region16 = reinterpret_cast<uint16_t*>(smem);
  unsigned char* ring = smem + (size_t)SJ_SLOTS * 2;
  __shared__ __align__(8) uint64_t s_full[SJ_STAGES], s_empty[SJ_STAGES];
  __shared__ __align__(8) unsigned long long s_blk[SJ_NBLK];  // block b: first pair index | (b + 1) << 40
  __shared__ unsigned long long s_rowsum[2], s_slotsum[2];
  __shared__ uint32_t s_vcur;                                  // pairs emitted by this CTA

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // an earlier kernel of this attempt gave up: nothing to do (uniform; before any copy is in flight)
  if (*reinterpret_cast<volatile unsigned int*>(&a.ctl->flags) & (CTL_NOT_DENSE16 | CTL_OVERFLOW | CTL_META_CHANGED | CTL_PEER_TIMEOUT)) {
    if (MAT && tid == 0) a.tails[blockIdx.x * SJ_TAIL_WORDS] = ~0ull;
    return;
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < SJ_STAGES; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], SJ_CWARPS);
    }
    mbar_fence_init();
#pragma unroll
    for (int i = 0; i < SJ_NBLK; ++i) s_blk[i] = 0;
    s_rowsum[0] = s_rowsum[1] = s_slotsum[0] = s_slotsum[1] = 0;
    s_vcur = 0;
  }
  for (uint32_t i = (uint32_t)tid; i < SJ_SLOTS / 8u; i += SJ_THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();

  // The sources of a partition are visited in an order rotated by GPU and CTA: with the same order everywhere, every
  // SM of every GPU would pull from source 0 first, then from source 1, ... — one GPU's NVLink egress at a time
  // (C3 at 8 GPUs: k_sjoin 0.186 ms, 350 GB/s per GPU, profiles/r02s_bench_n8.json)
  auto sub_at = [&](int si) -> int { return (int)(((uint32_t)si + a.rot + blockIdx.x) % (uint32_t)a.nsub); };
  // elements of one side of partition l (local index), summed over the sub-regions
  auto side_total = [&](const uint32_t* cnt, uint64_t cap, uint32_t l) -> uint64_t {
    uint64_t t = 0;
    for (int sub = 0; sub < a.nsub; ++sub) {
      uint64_t c = cnt[((size_t)sub * a.cnt_stride + a.p_first + l) * a.cstride];
      t += c < cap ? c : cap;
    }
    return t;
  };

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    if (lane != 0) return;
    uint32_t it = 0;
    for (uint32_t l = blockIdx.x; l < a.p_count; l += gridDim.x) {
      if (side_total(a.bcnt, a.cap_b, l) == 0 || side_total(a.pcnt, a.cap_p, l) == 0) continue;
      for (int side = 0; side < 2; ++side) {
        const uint32_t* cnt = side ? a.pcnt : a.bcnt;
        const uint64_t cap = side ? a.cap_p : a.cap_b;
        const uint32_t eb = side ? 2u : EB;
        // The rows of partition p on this side = the pieces every source keeps in ITS partition buffer (local, or a
        // peer's over NVLink), concatenated in the rotated source order.  The stream is cut into ring stages of SJ_CH
        // bytes; a stage is filled by one bulk copy per piece (or part of a piece) it covers, so the number of stages —
        // and of producer / consumer handshakes — does not grow with the number of sources.
        uint64_t stream = 0;
        for (int si = 0; si < a.nsub; ++si) {
          uint64_t c = cnt[((size_t)sub_at(si) * a.cnt_stride + a.p_first + l) * a.cstride];
          stream += (c < cap ? c : cap) * eb;
        }
        int si = 0;
        uint64_t piece_left = 0;
        const unsigned char* piece = nullptr;
        for (uint64_t off = 0; off < stream; off += SJ_CH) {
          const int s = it % SJ_STAGES;
          mbar_wait_bounded(&s_empty[s], ((it / SJ_STAGES) & 1u) ^ 1u);
          const uint32_t bytes = (uint32_t)(stream - off < (uint64_t)SJ_CH ? stream - off : (uint64_t)SJ_CH);
          mbar_expect_tx(&s_full[s], bytes);
          uint32_t filled = 0;
          while (filled < bytes) {
            while (piece_left == 0) {  // next source with rows of this partition
              const int sub = sub_at(si++);
              uint64_t c = cnt[((size_t)sub * a.cnt_stride + a.p_first + l) * a.cstride];
              if (c > cap) c = cap;
              piece_left = c * eb;  // multiple of 32 (sectors)
              piece = (side ? a.probe[sub] : a.build[sub]) + (uint64_t)(a.p_first + l) * cap * eb;
            }
            // (splitting a piece into several smaller bulk copies does not pull faster over NVLink: 2 KB .. 8 KB copies
            // and whole pieces all gave 0.307 - 0.310 ms at 2 GPUs)
            const uint32_t n = (uint32_t)(piece_left < (uint64_t)(bytes - filled) ? piece_left : (uint64_t)(bytes - filled));
            bulk_g2s(ring + (size_t)s * SJ_CH + filled, piece, n, &s_full[s]);
            piece += n;
            piece_left -= n;
            filled += n;
          }
          ++it;
        }
      }
    }
    return;
  }

  // ------------------------------------------------------------------ consumers
  const uint32_t ct = (uint32_t)tid - 32u;
  unsigned long long local_count = 0;
  uint32_t it = 0, par = 0;
  const uint4 HOLES = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  // the eight 16-bit indices of one 16-byte piece of a probe chunk -> direct-address values (0 = no match)
  auto lookup8 = [&](const uint4& v, uint32_t (&idx)[8], uint32_t (&val)[8]) {
    const uint32_t e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      idx[r] = (r & 1) ? (e[r >> 1] >> 16) : (e[r >> 1] & 0xffffu);
      val[r] = region16[idx[r]];
    }
  };
  // reserve output block b (one lane): SJ_BLK pairs from the global cursor, published as base | (b + 1) << 40
  auto reserve_block = [&](uint32_t b) {
    const unsigned long long g = atomicAdd(&a.ctl->out_cursor, (unsigned long long)SJ_BLK);
    sts_v64(&s_blk[b % SJ_NBLK], g | ((unsigned long long)(b + 1u) << 40));
  };
  // first pair index of block b (waits for the reservation, which was issued a whole block earlier)
  auto block_base = [&](uint32_t b) -> unsigned long long {
    unsigned long long e = lds_v64(&s_blk[b % SJ_NBLK]);
    for (uint32_t spin = 0; (uint32_t)(e >> 40) != b + 1u; ++spin) {
      if (spin > (1u << 24)) __trap();
      e = lds_v64(&s_blk[b % SJ_NBLK]);
    }
    return e & ((1ull << 40) - 1ull);
  };

  for (uint32_t l = blockIdx.x; l < a.p_count; l += gridDim.x) {
    if (side_total(a.bcnt, a.cap_b, l) == 0 || side_total(a.pcnt, a.cap_p, l) == 0) continue;
    const uint32_t plow = a.p_first + l;  // the key bits the partition implies
    // ---- fill: region[idx] = value + 1 (count: 1)
    unsigned long long rowsum = 0;
    {
      const uint64_t bytes_total = side_total(a.bcnt, a.cap_b, l) * EB;  // the concatenated pieces of every source
      for (uint64_t off = 0; off < bytes_total; off += SJ_CH) {
        const int s = it % SJ_STAGES;
        mbar_wait_bounded(&s_full[s], (it / SJ_STAGES) & 1u);
        const uint32_t bytes = (uint32_t)(bytes_total - off < (uint64_t)SJ_CH ? bytes_total - off : (uint64_t)SJ_CH);
        const uint4* st = reinterpret_cast<const uint4*>(ring + (size_t)s * SJ_CH);
        uint4 v[SJ_NPC];
#pragma unroll
        for (int q = 0; q < SJ_NPC; ++q) {
          const uint32_t piece = (uint32_t)q * SJ_CONS + ct;
          v[q] = piece * 16u < bytes ? st[piece] : HOLES;
        }
        uint32_t part = 0;
#pragma unroll
        for (int q = 0; q < SJ_NPC; ++q) {
          const uint32_t e[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
          if constexpr (MAT) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              // a hole (idx 0xFFFF, value field 0xFFFF) stores 0 into slot 0xFFFF and adds 0 to the row sum
              const uint32_t val1 = ((e[r] >> 16) + 1u) & 0xffffu;
              region16[e[r] & 0xffffu] = (uint16_t)val1;
              part += val1;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const uint32_t i0 = e[r] & 0xffffu, i1 = e[r] >> 16;
              region16[i0] = i0 != 0xffffu;
              region16[i1] = i1 != 0xffffu;
            }
          }
        }
        rowsum += part;
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[s]);
        ++it;
      }
    }
    if constexpr (MAT) {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) rowsum += __shfl_xor_sync(0xffffffffu, rowsum, d);
      if (lane == 0 && rowsum) atomicAdd(&s_rowsum[par], rowsum);
    }
    sj_bar_consumers();
    // ---- probe: one pass
    uint32_t mine = 0;
    {
      const uint64_t bytes_total = side_total(a.pcnt, a.cap_p, l) * 2u;
      for (uint64_t off = 0; off < bytes_total; off += SJ_CH) {
        const int s = it % SJ_STAGES;
        mbar_wait_bounded(&s_full[s], (it / SJ_STAGES) & 1u);
        const uint32_t bytes = (uint32_t)(bytes_total - off < (uint64_t)SJ_CH ? bytes_total - off : (uint64_t)SJ_CH);
        const uint4* st = reinterpret_cast<const uint4*>(ring + (size_t)s * SJ_CH);
        uint4 v[SJ_NPC];
#pragma unroll
        for (int q = 0; q < SJ_NPC; ++q) {
          const uint32_t piece = (uint32_t)q * SJ_CONS + ct;
          v[q] = piece * 16u < bytes ? st[piece] : HOLES;
        }
#pragma unroll
        for (int q = 0; q < SJ_NPC; ++q) {
          uint32_t idx[8], val[8];
          lookup8(v[q], idx, val);
          if (q == SJ_NPC - 1) {  // every piece of the stage is in registers (the lookups depend on them)
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
          }
          if constexpr (!MAT) {
#pragma unroll
            for (int r = 0; r < 8; ++r) mine += val[r] != 0u;
          } else {
            uint32_t offs[8];
            uint32_t wtot = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              const unsigned bal = __ballot_sync(0xffffffffu, val[r] != 0u);
              offs[r] = wtot + __popc(bal & lanemask_lt());
              wtot += __popc(bal);
            }
            if (wtot) {  // warp-uniform
              uint32_t v0 = 0;
              if (lane == 0) {
                v0 = atomicAdd(&s_vcur, wtot);
                const uint32_t bl = (v0 + wtot - 1u) >> SJ_LOG_BLK;
                if (v0 == 0u) {  // the CTA's first pairs
                  reserve_block(0u);
                  reserve_block(1u);
                } else if (bl != ((v0 - 1u) >> SJ_LOG_BLK)) {  // first to enter block bl
                  reserve_block(bl + 1u);
                }
              }
              v0 = __shfl_sync(0xffffffffu, v0, 0);
              const uint32_t b0 = v0 >> SJ_LOG_BLK;
              // pair number pv of block b lies at base(b) + (pv - b * BLK): fold the block start into the pointers
              const unsigned long long base0 = block_base(b0) - ((unsigned long long)b0 << SJ_LOG_BLK);
              unsigned long long* const kp0 = a.out_keys + base0;
              unsigned long long* const vp0 = a.out_vals + base0;
              if (((v0 + wtot - 1u) >> SJ_LOG_BLK) == b0) {  // warp-uniform, 15 cases of 16: one block
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                  if (val[r] != 0u) {
                    const uint32_t pv = v0 + offs[r];
                    st_stream(kp0 + pv, (unsigned long long)((idx[r] << a.logp) | plow));  // keys of this path are < 2^32
                    st_stream(vp0 + pv, (unsigned long long)(val[r] - 1u));
                  }
                }
              } else {
                const unsigned long long base1 = block_base(b0 + 1u) - ((unsigned long long)(b0 + 1u) << SJ_LOG_BLK);
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                  if (val[r] != 0u) {
                    const uint32_t pv = v0 + offs[r];
                    const unsigned long long o = ((pv >> SJ_LOG_BLK) == b0 ? base0 : base1) + pv;
                    st_stream(a.out_keys + o, (unsigned long long)((idx[r] << a.logp) | plow));
                    st_stream(a.out_vals + o, (unsigned long long)(val[r] - 1u));
                  }
                }
              }
            }
          }
        }
        ++it;
      }
    }
    local_count += mine;
    sj_bar_consumers();  // every lookup of this partition is done
    // ---- clear the addressable part of the region (MAT: and sum its contents)
    {
      uint4* r4 = reinterpret_cast<uint4*>(smem);
      const uint32_t n4 = a.slots / 8u;
      uint32_t part = 0;
      for (uint32_t i = ct; i < n4; i += SJ_CONS) {
        if constexpr (MAT) {
          const uint4 x = r4[i];
          part = add_halves(x.x, add_halves(x.y, add_halves(x.z, add_halves(x.w, part))));
        }
        r4[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      if constexpr (MAT) {
        unsigned long long slotsum = part;  // <= 67 slots x 65535 per thread
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) slotsum += __shfl_xor_sync(0xffffffffu, slotsum, d);
        if (lane == 0 && slotsum) atomicAdd(&s_slotsum[par], slotsum);
      }
    }
    sj_bar_consumers();
    if constexpr (MAT) {
      if (ct == 0) {  // the sums of this parity are next touched two partitions (four barriers) from now
        if (s_rowsum[par] != s_slotsum[par]) atomicOr(&a.ctl->flags, CTL_DUP);
        s_rowsum[par] = 0;
        s_slotsum[par] = 0;
      }
      par ^= 1u;
    }
  }
  if constexpr (MAT) {
    sj_bar_consumers();
    if (ct == 0) {
      const uint32_t total = s_vcur;
      unsigned long long* t = a.tails + (size_t)blockIdx.x * SJ_TAIL_WORDS;
      if (total == 0u) {
        t[0] = ~0ull;
      } else {
        const uint32_t bl = (total - 1u) >> SJ_LOG_BLK;
        t[0] = block_base(bl);
        t[1] = total - (bl << SJ_LOG_BLK);
        t[2] = block_base(bl + 1u);
        atomicAdd(&a.ctl->match_count, (unsigned long long)total);
      }
    }
  } else {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) local_count += __shfl_xor_sync(0xffffffffu, local_count, d);
    if (lane == 0 && local_count) atomicAdd(&a.ctl->match_count, local_count);
  }
}

// The pairs k_sjoin<true> wrote occupy [0, R) (R = Ctl::out_cursor, whole blocks) with one or two holes per CTA: the
// unused tail of its last block and the block it reserved ahead.  Move the pairs at positions >= M (M = Ctl::match_count)
// into the holes below M, so that the pairs are the dense range [0, M) the C ABI hands out.  Every CTA derives the same
// (small) tables — holes sorted by position, prefix sums of the hole positions below M (destinations) and of the filled
// positions at or above M (sources; as many as destinations) — then fills its share of the holes.
constexpr int PC_MAXH = 512;  // holes: two per k_sjoin CTA
constexpr int PC_THREADS = 512;
static_assert(PC_MAXH + 1 <= 2 * PC_THREADS, "pc_scan2 handles two elements per thread");
// inclusive prefix sums of a[0, n) and b[0, n) in place, n <= 2 * PC_THREADS: two elements per thread, warp shuffles, one
// pass over the 16 warp totals (three block barriers instead of the four per doubling step of a ping-pong scan)
__device__ __forceinline__ void pc_scan2(unsigned long long* a, unsigned long long* b, unsigned long long* wsum /* [2 * PC_THREADS / 32] */,
                                         uint32_t n) {
  constexpr int NW = PC_THREADS / 32;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t i0 = 2u * threadIdx.x, i1 = i0 + 1u;
  const unsigned long long a0 = i0 < n ? a[i0] : 0ull, a1 = i1 < n ? a[i1] : 0ull;
  const unsigned long long b0 = i0 < n ? b[i0] : 0ull, b1 = i1 < n ? b[i1] : 0ull;
  unsigned long long sa = a0 + a1, sb = b0 + b1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long oa = __shfl_up_sync(0xffffffffu, sa, d), ob = __shfl_up_sync(0xffffffffu, sb, d);
    if ((int)lane >= d) { sa += oa; sb += ob; }
  }
  if (lane == 31u) { wsum[warp] = sa; wsum[NW + warp] = sb; }
  __syncthreads();
  if (warp == 0) {
    unsigned long long va = lane < (uint32_t)NW ? wsum[lane] : 0ull, vb = lane < (uint32_t)NW ? wsum[NW + lane] : 0ull;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      const unsigned long long oa = __shfl_up_sync(0xffffffffu, va, d), ob = __shfl_up_sync(0xffffffffu, vb, d);
      if ((int)lane >= d) { va += oa; vb += ob; }
    }
    if (lane < (uint32_t)NW) { wsum[lane] = va; wsum[NW + lane] = vb; }
  }
  __syncthreads();
  const unsigned long long ea = (warp ? wsum[warp - 1] : 0ull) + sa - (a0 + a1);  // sum of everything in front of this thread's pair
  const unsigned long long eb = (warp ? wsum[NW + warp - 1] : 0ull) + sb - (b0 + b1);
  if (i0 < n) { a[i0] = ea + a0; b[i0] = eb + b0; }
  if (i1 < n) { a[i1] = ea + a0 + a1; b[i1] = eb + b0 + b1; }
  __syncthreads();
}
__global__ void __launch_bounds__(PC_THREADS) k_pairs_compact(const Ctl* ctl, const unsigned long long* tails, uint32_t ncta,
                                                               unsigned long long* out_keys, unsigned long long* out_vals) {
  __shared__ unsigned long long us[PC_MAXH], ue[PC_MAXH];  // holes as recorded
  __shared__ unsigned long long hs[PC_MAXH], he[PC_MAXH];  // holes sorted by start, clipped to [0, R]
  __shared__ unsigned long long dsum[PC_MAXH + 1];         // inclusive prefix sums: hole positions below M
  __shared__ unsigned long long fs[PC_MAXH + 1], fsum[PC_MAXH + 1];  // filled segments at or above M: start, inclusive prefix sums of lengths
  __shared__ unsigned long long wsum[2 * PC_THREADS / 32];
  // the attempt was abandoned (k_sjoin returned at once, or overflowed): nothing to compact
  if (ctl->flags & (CTL_NOT_DENSE16 | CTL_OVERFLOW | CTL_META_CHANGED | CTL_PEER_TIMEOUT)) return;
  const unsigned long long R = ctl->out_cursor, M = ctl->match_count;
  const uint32_t nh = 2u * ncta;
  for (uint32_t i = threadIdx.x; i < nh; i += PC_THREADS) {
    const unsigned long long* t = tails + (size_t)(i >> 1) * SJ_TAIL_WORDS;
    unsigned long long s = R, e = R;
    if (t[0] != ~0ull) {
      if (i & 1u) { s = t[2]; e = t[2] + SJ_BLK; }
      else { s = t[0] + t[1]; e = t[0] + SJ_BLK; }
      if (s >= e) s = e = R;
    }
    us[i] = s;
    ue[i] = e;
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nh; i += PC_THREADS) {  // rank sort (starts are distinct unless the hole is empty)
    const unsigned long long s = us[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nh; ++j) rank += (us[j] < s) || (us[j] == s && j < i);
    hs[rank] = s;
    he[rank] = ue[i];
  }
  __syncthreads();
  // holes are disjoint and sorted: the filled segment in front of hole k starts where hole k - 1 ends (not below M)
  for (uint32_t k = threadIdx.x; k <= nh; k += PC_THREADS) {
    const unsigned long long prev_end = k ? (he[k - 1] > M ? he[k - 1] : M) : M;
    const unsigned long long s = k < nh ? hs[k] : R, e = k < nh ? he[k] : R;
    const unsigned long long s_hi = s > M ? s : M;
    fs[k] = prev_end;
    fsum[k] = s_hi > prev_end ? s_hi - prev_end : 0ull;
    dsum[k] = k < nh ? (e < M ? e : M) - (s < M ? s : M) : 0ull;
  }
  __syncthreads();
  pc_scan2(dsum, fsum, wsum, nh + 1);
  // hole k receives the sources number [dsum[k - 1], dsum[k]) (numbered along the filled segments); every thread
  // has all its loads in flight before the first store
  for (uint32_t k = blockIdx.x; k < nh; k += gridDim.x) {
    const unsigned long long t0 = k ? dsum[k - 1] : 0ull;
    const unsigned long long len = dsum[k] - t0;
    const unsigned long long dst0 = hs[k];
    for (unsigned long long base = 0; base < len; base += 8ull * PC_THREADS) {
      unsigned long long kk[8], vv[8];
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const unsigned long long e = base + (unsigned long long)m * PC_THREADS + threadIdx.x;
        if (e < len) {
          const unsigned long long t = t0 + e;
          uint32_t lo = 0, hi = nh + 1;  // smallest j with fsum[j] > t: the segment that holds source t
          while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            if (fsum[mid] > t) hi = mid; else lo = mid + 1;
          }
          const unsigned long long src = fs[lo] + (t - (lo ? fsum[lo - 1] : 0ull));
          kk[m] = out_keys[src];
          vv[m] = out_vals[src];
        }
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const unsigned long long e = base + (unsigned long long)m * PC_THREADS + threadIdx.x;
        if (e < len) {
          out_keys[dst0 + e] = kk[m];
          out_vals[dst0 + e] = vv[m];
        }
      }
    }
  }
}

size_t sjoin_smem_bytes(uint32_t) { return (size_t)SJ_SLOTS * 2 + (size_t)SJ_STAGES * SJ_CH; }
uint32_t sjoin_max_slots(const DeviceInfo&) { return 65528; }  // idx 0xFFFF is the hole marker
uint32_t sjoin_grid(uint32_t p_count, const DeviceInfo& di) { return p_count < (uint32_t)di.sms ? p_count : (uint32_t)di.sms; }
size_t sjoin_tail_bytes(const DeviceInfo& di) { return (size_t)di.sms * SJ_TAIL_WORDS * 8; }
uint64_t sjoin_out_slack_pairs(const DeviceInfo& di) { return (uint64_t)di.sms * 2 * SJ_BLK; }

bool launch_sjoin(bool mat, const SjoinArgs& x, const DeviceInfo& di, cudaStream_t st, int* launches) {
  if (x.p_count == 0) return false;
  SjoinParams a;
  if (x.nsub < 1 || x.nsub > PT_MAXW) return false;
  for (int i = 0; i < PT_MAXW; ++i) {
    a.build[i] = static_cast<const unsigned char*>(i < x.nsub ? x.build[i] : nullptr);
    a.probe[i] = static_cast<const unsigned char*>(i < x.nsub ? x.probe[i] : nullptr);
  }
  a.bcnt = x.bcnt; a.cap_b = x.cap_b; a.pcnt = x.pcnt; a.cap_p = x.cap_p;
  a.cnt_stride = x.cnt_stride; a.cstride = x.cursor_stride; a.p_first = x.p_first; a.p_count = x.p_count; a.logp = x.logp; a.nsub = x.nsub; a.rot = x.rot;
  a.slots = x.slots_alloc; a.ctl = x.ctl; a.out_keys = x.out_keys; a.out_vals = x.out_vals; a.tails = x.tails;
  const size_t smem = sjoin_smem_bytes(x.slots_alloc);
  if (smem + 256 > di.smem_optin || (x.slots_alloc & 7u) || x.slots_alloc > 65528u) return false;
  if (x.cap_b * 4 > 0xFFFFFFF0ull || x.cap_p * 2 > 0xFFFFFFF0ull) return false;
  const uint32_t grid = sjoin_grid(x.p_count, di);
  if (mat) {
    if (!x.tails || 2u * grid > (uint32_t)PC_MAXH) return false;
    static int attr_mat_dev = -1;
    if (attr_mat_dev != di.device) { cudaFuncSetAttribute(k_sjoin<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_mat_dev = di.device; }
    k_sjoin<true><<<grid, SJ_THREADS, smem, st>>>(a);
    k_pairs_compact<<<di.sms, PC_THREADS, 0, st>>>(x.ctl, x.tails, grid, x.out_keys, x.out_vals);
    ++*launches;
  } else {
    static int attr_cnt_dev = -1;
    if (attr_cnt_dev != di.device) { cudaFuncSetAttribute(k_sjoin<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_cnt_dev = di.device; }
    k_sjoin<false><<<grid, SJ_THREADS, smem, st>>>(a);
  }
  ++*launches;
  return true;
}

}  // namespace fj
