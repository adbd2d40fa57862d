// fj_common.cuh — device-side primitives shared by the join kernels (sm_100a only).
//
// Replaces the reference's L1 primitives (/root/reference/hash_join.cpp:35-204): hash64/Hasher
// (:40-59), the Slot layout and EMPTY_TAG (:78-85), get_bloom_tag/check_bloom_filter (:183-189).
// The hash function is not part of the observable contract (counts and the sorted pair multiset
// are hash independent), so a GPU-friendly 32-bit mixer replaces CRC32C.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace fj {

constexpr unsigned long long EMPTY64 = ~0ull;  // empty slot / empty key sentinel

// control block living in device memory, one per join attempt
struct Ctl {
  unsigned long long match_count;   // number of probe rows that matched
  unsigned long long out_cursor;    // bump cursor into the materialized pair arrays
  unsigned long long sentinel_row;  // min build row index whose key == EMPTY64 (~0 = none); out-of-band key
  unsigned long long sentinel_probes;  // radix path: probe rows whose key == EMPTY64 (joined out of band)
  unsigned int flags;               // CTL_* bits raised by kernels
  unsigned int pad;
  // dense-key-domain paths (direct addressing instead of hashing)
  unsigned long long max_key;       // largest build key the stage-1 scatter staged
  unsigned long long dense_rows;    // build rows stored into the direct-address regions
  unsigned long long dense_slots;   // non-empty direct-address slots afterwards (< dense_rows <=> duplicate build keys)
  // multi-GPU count over peer memory (k_count_dense_peer): sum of every rank's match_count
  unsigned long long global_count;
};
enum : unsigned {
  CTL_NEED_WIDE = 1u,  // a build key or value does not fit the packed 32|32 slot
  CTL_DUP = 2u,        // duplicate build keys seen: keep-first needs the exact path
  CTL_OVERFLOW = 4u,   // an optimistic fixed-capacity partition buffer overflowed
  CTL_NOT_DENSE = 8u,  // a build key (or value) lies outside the optimistic dense key domain
  CTL_PEER_TIMEOUT = 16u,  // a peer GPU did not show up in the peer-memory exchange (k_count_dense_peer)
  CTL_NOT_DENSE16 = 32u,   // k_part: a build key outside the 16-bit-index dense domain, or a build value > 65534
  CTL_META_CHANGED = 64u,  // k_xsync: some rank joined the step with other slice sizes than the plan was made for
  CTL_LOW_SEL = 128u,      // k_sel_sample: too few probe rows match for dense16 to beat the dense table path (with CTL_NOT_DENSE16)
};

// ---- hashing -----------------------------------------------------------------------------------
// lowbias32 finaliser (a bijection on 32 bits) over lo ^ hi*odd.  ~8 integer instructions.
__host__ __device__ __forceinline__ uint32_t hash32(uint64_t key) {
  uint32_t x = (uint32_t)key ^ ((uint32_t)(key >> 32) * 0x9E3779B1u);
  x ^= x >> 16; x *= 0x7feb352dU;
  x ^= x >> 15; x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
// multiply-shift range reduction: [0, 2^32) -> [0, n)   (no power-of-two rounding of tables, cf. :96-99)
__host__ __device__ __forceinline__ uint32_t reduce32(uint32_t h, uint32_t n) {
#ifdef __CUDA_ARCH__
  return __umulhi(h, n);
#else
  return (uint32_t)(((uint64_t)h * n) >> 32);
#endif
}

// ---- Bloom filter: sectorised, register-blocked ------------------------------------------------
// One 32-bit word per key ("register-blocked": both bits of a key live in a single word, so a
// check is one 4-byte load and one AND/compare); the filter is a dense array of 32-byte sectors
// and, when it fits, is staged in shared memory by the probe kernel.  Replaces the
// 16-bit-per-slot directory of hash_join.cpp:60-74, :183-189.  The filter has its own, cheaper
// hash (5 integer instructions): it runs for EVERY probe row, the table hash only for survivors.
__host__ __device__ __forceinline__ uint32_t bloom_hash(uint64_t key) {
  uint32_t x = (uint32_t)key * 0x9E3779B1u + (uint32_t)(key >> 32) * 0x85EBCA77u;
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  return x;
}
__host__ __device__ __forceinline__ uint32_t bloom_mask(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __funnelshift_l(0u, 1u, x) | __funnelshift_l(0u, 1u, x >> 5);  // 1 << (x & 31) | 1 << ((x >> 5) & 31)
#else
  return (1u << (x & 31)) | (1u << ((x >> 5) & 31));
#endif
}
__host__ __device__ __forceinline__ uint32_t bloom_word(uint32_t x, uint32_t nwords) { return reduce32(x, nwords); }

// packed ("narrow") rows: key and value both fit 32 bits and key != 0xFFFFFFFF, so that a packed
// slot key32<<32|value32 can never equal the empty marker and "high word == 0xFFFFFFFF" means empty.
__host__ __device__ __forceinline__ bool narrow_ok(unsigned long long k, unsigned long long v) {
  return ((k | v) >> 32) == 0 && (uint32_t)k != 0xFFFFFFFFu;
}

#ifdef __CUDACC__
// ---- loads -------------------------------------------------------------------------------------
// one whole 32-byte sector in a single LDG.256 (sm_100 has 256-bit global loads)
__device__ __forceinline__ void ld_sector(const unsigned long long* p, unsigned long long& a,
                                          unsigned long long& b, unsigned long long& c,
                                          unsigned long long& d) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];"
               : "=l"(a), "=l"(b), "=l"(c), "=l"(d)
               : "l"(p));
}
// streaming 128-bit load of two probe keys (read once: do not allocate in L1)
__device__ __forceinline__ void ld_stream2(const unsigned long long* p, unsigned long long& a,
                                           unsigned long long& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
}
__device__ __forceinline__ unsigned long long ld_stream1(const unsigned long long* p) {
  unsigned long long a;
  asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(a) : "l"(p));
  return a;
}
__device__ __forceinline__ void st_stream(unsigned long long* p, unsigned long long v) {
  asm volatile("st.global.cs.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---- TMA bulk copy (cp.async.bulk, non-tensor) + mbarrier --------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned; completes on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global bulk copy (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy before a bulk store reads them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}
#endif  // __CUDACC__

}  // namespace fj
