/*
 * oracle/join_oracle.c — CPU restatement of the reference join algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in flash_hash_join_b200/ may import, link or call this
 * file; it exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can check
 * the CUDA engine.  Parity status: PINNED — this restatement is checked (tests/test_oracle.py)
 * against the compiled, unmodified reference (oracle/_ref, built by oracle/build_ref.sh) and
 * against the golden vectors in tests/golden/ that were produced by that reference binary.
 *
 * Each function cites the lines of /root/reference/hash_join.cpp it restates.  The restatement
 * is single-threaded on purpose: the reference's scalar (non-partitioned) build is a race for
 * duplicate build keys (hash_join.cpp:130-151), while its radix path and its 1-thread behaviour
 * are "keep the first occurrence in input order" (hash_join.cpp:112-128, :191, :226-234).  The
 * engine's contract is keep-first, which is what this file computes on every path.
 *
 * Plain C11, no dependencies.  Build: gcc -O2 -shared -fPIC join_oracle.c -o _build/libjoin_oracle.so
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FJO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * hash64 — hash_join.cpp:40-44.  _mm_crc32_u64(seed, key) is the CRC32C (Castagnoli, reflected
 * polynomial 0x82F63B78) update of the 32-bit state `seed` with the 8 little-endian bytes of
 * `key`, with no initial or final inversion.  The 64-bit hash is crc * 0x8648DBDB00000001.
 * ------------------------------------------------------------------------------------------ */
static uint32_t crc32c_table[8][256];
static int crc32c_ready = 0;

static void crc32c_init(void) {
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : (c >> 1);
    crc32c_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = crc32c_table[0][i];
    for (int t = 1; t < 8; ++t) {
      c = crc32c_table[0][c & 0xffu] ^ (c >> 8);
      crc32c_table[t][i] = c;
    }
  }
  crc32c_ready = 1;
}

static inline uint32_t crc32c_u64(uint32_t crc, uint64_t v) {
  /* slicing-by-8 over one 64-bit little-endian word */
  uint64_t x = v ^ (uint64_t)crc;
  return crc32c_table[7][x & 0xff] ^ crc32c_table[6][(x >> 8) & 0xff] ^
         crc32c_table[5][(x >> 16) & 0xff] ^ crc32c_table[4][(x >> 24) & 0xff] ^
         crc32c_table[3][(x >> 32) & 0xff] ^ crc32c_table[2][(x >> 40) & 0xff] ^
         crc32c_table[1][(x >> 48) & 0xff] ^ crc32c_table[0][(x >> 56) & 0xff];
}

FJO_API uint64_t fjo_hash64(uint64_t key, uint32_t seed) {
  if (!crc32c_ready) crc32c_init();
  const uint64_t k = 0x8648DBDBull;
  uint64_t crc = crc32c_u64(seed, key); /* zero-extended 32-bit result, as the intrinsic returns */
  return crc * ((k << 32) + 1);
}

/* Hasher — hash_join.cpp:56-59 */
static inline uint64_t hasher(uint64_t key) { return fjo_hash64(key, 0xAAAAAAAAu); }

/* ------------------------------------------------------------------------------------------
 * Bloom tag table — hash_join.cpp:60-74 (create_tags_table) and :183 (get_bloom_tag).
 * ------------------------------------------------------------------------------------------ */
#define TAGS_TABLE_SIZE 2048
static uint16_t tags_table[TAGS_TABLE_SIZE];
static int tags_ready = 0;

static void tags_init(void) {
  for (uint32_t i = 0; i < TAGS_TABLE_SIZE; ++i) {
    uint32_t h = i * 0x9E3779B9u;
    uint16_t b1 = (uint16_t)(1u << ((h >> 0) & 15));
    uint16_t b2 = (uint16_t)(1u << ((h >> 8) & 15));
    uint16_t b3 = (uint16_t)(1u << ((h >> 16) & 15));
    uint16_t b4 = (uint16_t)(1u << ((h >> 24) & 15));
    tags_table[i] = (uint16_t)(b1 | b2 | b3 | b4);
  }
  tags_ready = 1;
}

FJO_API uint16_t fjo_bloom_tag(uint64_t hash) {
  if (!tags_ready) tags_init();
  return tags_table[((uint32_t)hash) >> (32 - 11)];
}

/* ------------------------------------------------------------------------------------------
 * FlashHashTable — hash_join.cpp:75-204.
 * Slot = {tag, key, value}; EMPTY_TAG 0xFF; capacity = next_pow2(size_t(build_size*1.5 + 32)).
 * ------------------------------------------------------------------------------------------ */
#define EMPTY_TAG 0xFFu
#define SIMD_WIDTH 32

typedef struct {
  uint8_t* tag;
  uint64_t* key;
  uint64_t* value;
  uint16_t* bloom; /* NULL when UseBloomFilter == false */
  size_t capacity, mask;
} fjo_table;

/* calculate_power_of_2 — hash_join.cpp:96 */
static size_t pow2_ceil(size_t n) {
  if (n == 0) return 1;
  if (n == 1) return 1; /* 1UL << (64 - clzll(0)) is UB in the reference; never reached (n >= 32) */
  return (size_t)1 << (64 - __builtin_clzll((unsigned long long)(n - 1)));
}

FJO_API size_t fjo_table_capacity(size_t build_size) {
  /* hash_join.cpp:99 — double arithmetic then truncation to size_t */
  return pow2_ceil((size_t)((double)build_size * 1.5 + (double)SIMD_WIDTH));
}

/* ctor — hash_join.cpp:98-110 */
static int table_init(fjo_table* t, size_t build_size, int use_bloom) {
  if (!tags_ready) tags_init();
  if (!crc32c_ready) crc32c_init();
  t->capacity = fjo_table_capacity(build_size);
  t->mask = t->capacity - 1;
  t->tag = (uint8_t*)malloc(t->capacity);
  t->key = (uint64_t*)malloc(t->capacity * sizeof(uint64_t));
  t->value = (uint64_t*)malloc(t->capacity * sizeof(uint64_t));
  t->bloom = use_bloom ? (uint16_t*)calloc(t->capacity, sizeof(uint16_t)) : NULL;
  if (!t->tag || !t->key || !t->value || (use_bloom && !t->bloom)) return -1;
  memset(t->tag, EMPTY_TAG, t->capacity);
  return 0;
}

static void table_free(fjo_table* t) {
  free(t->tag); free(t->key); free(t->value); free(t->bloom);
  memset(t, 0, sizeof(*t));
}

/* insert_local — hash_join.cpp:112-128: first empty slot in the probe sequence, or return when
 * the key is already present (keep-first).  Note the reference compares keys without checking the
 * tag here (:125); an occupied slot always holds a real key so the result is the same. */
static void table_insert(fjo_table* t, uint64_t key, uint64_t value) {
  uint64_t hash = hasher(key);
  uint8_t tag = (uint8_t)(hash >> 56);
  if (tag == EMPTY_TAG) tag = 0;
  size_t pos = hash & t->mask;
  const size_t initial_pos = pos;
  do {
    if (t->tag[pos] == EMPTY_TAG) {
      t->key[pos] = key;
      t->value[pos] = value;
      t->tag[pos] = tag;
      if (t->bloom) t->bloom[initial_pos] |= fjo_bloom_tag(hash);
      return;
    }
    if (t->key[pos] == key) return;
    pos = (pos + 1) & t->mask;
  } while (pos != initial_pos);
}

/* check_bloom_filter — hash_join.cpp:185-189 */
static inline int table_bloom_pass(const fjo_table* t, uint64_t hash) {
  uint16_t entry = t->bloom[hash & t->mask];
  uint16_t m = fjo_bloom_tag(hash);
  return (m & entry) == m;
}

/* one key of probe_vectorized — hash_join.cpp:163-179 (the prefetch at :157-162 has no effect on
 * results).  Returns 1 and *value on the first slot whose tag and key match; stops at an empty tag. */
static inline int table_probe(const fjo_table* t, uint64_t key, uint64_t* value) {
  uint64_t hash = hasher(key);
  if (t->bloom && !table_bloom_pass(t, hash)) return 0;
  uint8_t tag = (uint8_t)(hash >> 56);
  if (tag == EMPTY_TAG) tag = 0;
  size_t pos = hash & t->mask;
  const size_t initial_pos = pos;
  do {
    uint8_t cur = t->tag[pos];
    if (cur == EMPTY_TAG) return 0;
    if (cur == tag && t->key[pos] == key) { *value = t->value[pos]; return 1; }
    pos = (pos + 1) & t->mask;
  } while (pos != initial_pos);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Radix partitioning — hash_join.cpp:206-292.  256 partitions on the top 8 hash bits (:209);
 * histogram + prefix + scatter; with one thread the (partition, thread) prefix of :226-234 is a
 * plain stable counting sort, and for T threads over contiguous chunks it is the same order.
 * ------------------------------------------------------------------------------------------ */
#define RADIX_BITS 8
#define NUM_PARTITIONS (1u << RADIX_BITS)

FJO_API uint32_t fjo_partition_idx(uint64_t key) { return (uint32_t)(hasher(key) >> (64 - RADIX_BITS)); }

static int radix_partition(const uint64_t* keys, const uint64_t* values, size_t n,
                           uint64_t** out_keys, uint64_t** out_values, size_t* offsets /*257*/) {
  size_t hist[NUM_PARTITIONS];
  memset(hist, 0, sizeof(hist));
  uint8_t* pid = (uint8_t*)malloc(n ? n : 1);
  *out_keys = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  *out_values = values ? (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t)) : NULL;
  if (!pid || !*out_keys || (values && !*out_values)) { free(pid); return -1; }
  for (size_t j = 0; j < n; ++j) { pid[j] = (uint8_t)fjo_partition_idx(keys[j]); hist[pid[j]]++; }
  offsets[0] = 0;
  for (size_t p = 0; p < NUM_PARTITIONS; ++p) offsets[p + 1] = offsets[p] + hist[p];
  size_t pos[NUM_PARTITIONS];
  memcpy(pos, offsets, sizeof(pos));
  for (size_t j = 0; j < n; ++j) {
    size_t w = pos[pid[j]]++;
    (*out_keys)[w] = keys[j];
    if (values) (*out_values)[w] = values[j];
  }
  free(pid);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Join drivers — hash_join.cpp:315-567 and the adaptive dispatch :576-594.
 *   algo: 0 = adaptive, 1 = scalar (hash_join*), 2 = radix (hash_join*_radix)
 *   out_keys/out_vals: NULL for count; otherwise arrays of capacity np receiving the
 *   (probe key, build value) pairs (:351-352, :435-436, :466-467) in the reference's output
 *   order: probe order for the scalar path, partition-major for the radix path.
 * Returns the match count, or -1 on allocation failure / bad algo.
 * ------------------------------------------------------------------------------------------ */
#define RADIX_JOIN_THRESHOLD 1000000u /* hash_join.cpp:576 */

static int64_t join_scalar(int bloom, const uint64_t* bk, const uint64_t* bv, size_t nb,
                           const uint64_t* pk, size_t np, uint64_t* out_keys, uint64_t* out_vals) {
  /* _hash_join_scalar_count :536-567 / _hash_join_scalar_materialize :383-496 (both branches of
   * SMALL_TABLE_THRESHOLD :393 produce the same pairs in probe order) */
  fjo_table t;
  if (table_init(&t, nb, bloom)) return -1;
  for (size_t i = 0; i < nb; ++i) table_insert(&t, bk[i], bv[i]);
  int64_t m = 0;
  for (size_t j = 0; j < np; ++j) {
    uint64_t v;
    if (table_probe(&t, pk[j], &v)) {
      if (out_keys) { out_keys[m] = pk[j]; out_vals[m] = v; }
      ++m;
    }
  }
  table_free(&t);
  return m;
}

static int64_t join_radix(int bloom, const uint64_t* bk, const uint64_t* bv, size_t nb,
                          const uint64_t* pk, size_t np, uint64_t* out_keys, uint64_t* out_vals) {
  /* _hash_join_radix_materialize :315-381 / _hash_join_radix_count :498-534 */
  uint64_t *pbk = NULL, *pbv = NULL, *ppk = NULL, *unused = NULL;
  size_t boff[NUM_PARTITIONS + 1], poff[NUM_PARTITIONS + 1];
  if (radix_partition(bk, bv, nb, &pbk, &pbv, boff)) return -1;
  if (radix_partition(pk, NULL, np, &ppk, &unused, poff)) { free(pbk); free(pbv); return -1; }
  int64_t m = 0;
  for (size_t p = 0; p < NUM_PARTITIONS; ++p) {
    size_t bsz = boff[p + 1] - boff[p], psz = poff[p + 1] - poff[p];
    if (bsz == 0 || psz == 0) continue; /* :343 / :518 */
    fjo_table t;
    if (table_init(&t, bsz, bloom)) { m = -1; break; }
    for (size_t i = 0; i < bsz; ++i) table_insert(&t, pbk[boff[p] + i], pbv[boff[p] + i]);
    for (size_t j = 0; j < psz; ++j) {
      uint64_t key = ppk[poff[p] + j], v;
      if (table_probe(&t, key, &v)) {
        if (out_keys) { out_keys[m] = key; out_vals[m] = v; }
        ++m;
      }
    }
    table_free(&t);
  }
  free(pbk); free(pbv); free(ppk);
  return m;
}

FJO_API int64_t fjo_join(int algo, int bloom, const uint64_t* bk, const uint64_t* bv, size_t nb,
                         const uint64_t* pk, size_t np, uint64_t* out_keys, uint64_t* out_vals) {
  if ((out_keys == NULL) != (out_vals == NULL)) return -1;
  if (algo == 0) algo = (nb < RADIX_JOIN_THRESHOLD) ? 1 : 2; /* :580 / :589 */
  if (algo == 1) return join_scalar(bloom, bk, bv, nb, pk, np, out_keys, out_vals);
  if (algo == 2) return join_radix(bloom, bk, bv, nb, pk, np, out_keys, out_vals);
  return -1;
}

/* Which path adaptive takes (1 scalar / 2 radix) — hash_join.cpp:578-594 */
FJO_API int fjo_adaptive_path(size_t nb) { return nb < RADIX_JOIN_THRESHOLD ? 1 : 2; }

/* Stable 256-way partition exposed for tests of the partition step alone (:210-292). */
FJO_API int fjo_radix_partition(const uint64_t* keys, const uint64_t* values, size_t n,
                                uint64_t* out_keys, uint64_t* out_values, uint64_t* offsets257) {
  uint64_t *k = NULL, *v = NULL;
  size_t off[NUM_PARTITIONS + 1];
  if (radix_partition(keys, values, n, &k, &v, off)) return -1;
  memcpy(out_keys, k, n * sizeof(uint64_t));
  if (values && out_values) memcpy(out_values, v, n * sizeof(uint64_t));
  for (size_t p = 0; p <= NUM_PARTITIONS; ++p) offsets257[p] = off[p];
  free(k); free(v);
  return 0;
}
