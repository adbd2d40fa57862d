// fj_kernels.h — host-callable launchers of the CUDA kernels (internal; not part of the C ABI).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "fj_common.cuh"

#ifdef __CUDACC__
#include <tuple>
#include <utility>
#endif

namespace fj {

#ifdef __CUDACC__
// Launch a kernel whose CTAs synchronise with each other by spinning (grid barriers, producer / consumer tickets) as a
// COOPERATIVE launch: the driver then refuses the launch (an error return, and the caller takes its multi-kernel
// fallback) when the grid cannot be fully co-resident — MPS limits, green contexts, a persistent kernel of another
// stream — instead of starting a grid that would spin forever.
template <class... KArgs, size_t... I>
inline cudaError_t launch_coop_impl(void (*kern)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t st,
                                    std::tuple<KArgs...>& t, std::index_sequence<I...>) {
  void* p[] = {static_cast<void*>(&std::get<I>(t))...};
  return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(kern), dim3(grid), dim3(threads), p, smem, st);
}
template <class... KArgs, class... Args>
inline bool launch_coop(void (*kern)(KArgs...), unsigned grid, unsigned threads, size_t smem, cudaStream_t st, Args... args) {
  std::tuple<KArgs...> t(static_cast<KArgs>(args)...);
  if (launch_coop_impl(kern, grid, threads, smem, st, t, std::index_sequence_for<KArgs...>{}) == cudaSuccess) return true;
  cudaGetLastError();  // not sticky: the caller falls back
  return false;
}
#endif

struct DeviceInfo {
  int device = -1;
  int sms = 0;
  int l2_bytes = 0;
  size_t smem_optin = 0;
  int cc_major = 0, cc_minor = 0;
};

// ---------------------------------------------------------------- global ("scalar") table path
struct TableView {
  unsigned long long* slots = nullptr;  // narrow: nbuckets*4 packed words; wide: nbuckets*2 {key,value}
  uint32_t nbuckets = 0;                // 32-byte buckets
  uint32_t* bloom = nullptr;            // nullptr = no Bloom filter
  uint32_t bloom_words = 0;             // multiple of 4 (16-byte granularity for the bulk copy)
  bool narrow = false;
};

struct ProbeOut {
  unsigned long long* keys = nullptr;  // materialize: (probe key, build value[, probe row idx])
  unsigned long long* vals = nullptr;
  unsigned long long* idx = nullptr;
  unsigned long long idx_base = 0;     // added to the local probe row index (distributed slices)
};

void launch_init_ctl(Ctl* ctl, cudaStream_t st);
// publishes *ctl into mapped pinned memory: PUB_WORDS 64-bit words, word t = tag << 32 | t-th 32-bit word of the block
constexpr int PUB_WORDS = 18;
void launch_publish_ctl(const Ctl* ctl, void* mapped_dst, unsigned int tag, cudaStream_t st);
// control block reset + `ones` filled with 0xFF (empty table) + `zeros` cleared, in one launch; sizes are
// rounded up to 16 bytes (either region may be empty)
// counter_stride != 0: the zeros region is an array of 32-bit counters, one every counter_stride words (a multiple of
// 4); counters [0, counters_half) start at value_a, the others at value_b
void launch_prepare(Ctl* ctl, void* ones, size_t ones_bytes, void* zeros, size_t zeros_bytes, const DeviceInfo& di,
                    cudaStream_t st, uint32_t counter_stride = 0, uint32_t counters_half = 0, uint32_t value_a = 0, uint32_t value_b = 0);
// build kernels: mode 0 = fast (CAS on key, plain value store, raises CTL_DUP on duplicates),
//                mode 1 = exact keep-first (wide only: value word holds min row index, then fix-up)
void launch_build(const TableView& t, const unsigned long long* bk, const unsigned long long* bv, uint64_t nb,
                  int mode, Ctl* ctl, const DeviceInfo& di, cudaStream_t st, int* launches);
// probe: count only (out == nullptr) or materialize.  bloom_in_smem selects the shared-memory
// resident filter (requires bloom_words*4 + slack <= smem_optin).
void launch_probe(const TableView& t, const unsigned long long* pk, uint64_t np, const unsigned long long* bv,
                  const ProbeOut* out, bool bloom_in_smem, int ctas_per_sm, Ctl* ctl, const DeviceInfo& di,
                  cudaStream_t st, int* launches);
size_t probe_smem_bloom_limit_words(const DeviceInfo& di);
// dense key domain (count only): exact membership bitmap of `dbits` bits (multiple of 128) instead of table + filter;
// a build key >= dbits raises CTL_NOT_DENSE
size_t probe_smem_bitmap_limit_bytes(const DeviceInfo& di);
void launch_build_bitmap(uint32_t* bitmap, uint64_t dbits, const unsigned long long* bk, uint64_t nb, Ctl* ctl,
                         const DeviceInfo& di, cudaStream_t st, int* launches);
void launch_probe_count_dense(const unsigned long long* pk, uint64_t np, const uint32_t* bitmap, uint32_t dwords, Ctl* ctl,
                              const DeviceInfo& di, cudaStream_t st, int* launches);
// the three steps above (control block + empty bitmap, build, probe) as ONE persistent launch with two grid
// barriers; gsync = 3 words that are zero between launches (the kernel leaves them zero).  Initialises *ctl
// itself.  false: the launch configuration cannot guarantee co-residency -> use the three-kernel sequence
bool launch_count_dense_fused(const unsigned long long* bk, uint64_t nb, const unsigned long long* pk, uint64_t np,
                              uint32_t* bitmap, uint32_t dwords, Ctl* ctl, uint32_t* gsync, const DeviceInfo& di, cudaStream_t st,
                              int* launches);

// multi-GPU count over peer memory (one kernel per GPU, no NCCL in the step): `peers` = device array of every rank's
// IPC-mapped exchange buffer; the root's build keys must already lie at peer_staging_offset_bytes() of ITS buffer
// (stream-ordered before the launch); the sum of all ranks' counts arrives in Ctl::global_count
// relay: every rank pulls only its slice of the keys and the ranks exchange partial bitmaps (large build sides)
size_t peer_staging_offset_bytes();
size_t peer_staging_bytes();
size_t peer_buffer_bytes();  // exchange buffer every rank allocates: control words, key staging area, partial bitmap
bool launch_count_dense_peer(uint64_t nb, const unsigned long long* pk, uint64_t np, uint32_t* bitmap, uint32_t dwords, Ctl* ctl,
                             uint32_t* gsync, unsigned long long* const* peers, int rank, int world, int root,
                             unsigned long long step, bool relay, const DeviceInfo& di, cudaStream_t st, int* launches);
// dense key domain, materialize: bitmap + direct-address value table direct[key] (8 bytes per key of the domain,
// L2 resident), one persistent launch; a duplicate build key raises CTL_DUP, a key >= dbits CTL_NOT_DENSE
bool launch_mat_dense_fused(const unsigned long long* bk, const unsigned long long* bv, uint64_t nb, const unsigned long long* pk,
                            uint64_t np, uint32_t* bitmap, uint32_t dwords, unsigned long long* direct, Ctl* ctl, uint32_t* gsync,
                            const ProbeOut& po, const DeviceInfo& di, cudaStream_t st, int* launches);

// sum / OR of two words over all ranks through peer memory (k_peer_reduce): result[0] = sum of *sum_src (or sum_imm when
// sum_src is null), result[1] = OR of *or_src; world <= 32
void launch_peer_reduce(unsigned long long* const* peers, int rank, int world, unsigned long long step, const unsigned long long* sum_src,
                        unsigned long long sum_imm, const unsigned int* or_src, unsigned long long* result, uint32_t* err, cudaStream_t st,
                        int* launches);
// broadcast of `words` 64-bit words (a multiple of 2, at most peer_staging_bytes()) lying in the ROOT's staging area to
// `out` on every rank, over peer memory in two hops (k_peer_bcast); gsync words 0 and 2 must be zero (left zero); *err is
// set to 1 when a peer did not show up within 10 s
bool launch_peer_bcast(unsigned long long* const* peers, int rank, int world, int root, unsigned long long step, uint64_t words,
                       unsigned long long* out, uint32_t* err, uint32_t* gsync, const DeviceInfo& di, cudaStream_t st, int* launches);

// ---------------------------------------------------------------- radix-partitioned path
// key / value domain of the packed (narrow) stage-1 scatter.  General packed rows: keys < 2^32 - 1, values < 2^32,
// digit from hash32, a row outside raises CTL_NEED_WIDE.  Dense key domain: keys < the optimistic bound `klimit`,
// values < 2^32 - 1 (the direct-address slot stores value + 1), digit = the low key bits themselves (`ident`), a row
// outside raises CTL_NOT_DENSE and the largest staged key is recorded in Ctl::max_key.
struct DomainArgs {
  unsigned long long klimit = 0xFFFFFFFFull;  // packed rows need key < klimit ...
  unsigned long long vlimit = 0xFFFFFFFFull;  // ... and value <= vlimit
  unsigned badflag = CTL_NEED_WIDE;
  int ident = 0;
};
struct ScatterArgs {
  // stage 1 input: raw 64-bit columns
  const unsigned long long* in_keys = nullptr;
  const unsigned long long* in_vals = nullptr;
  uint64_t n = 0;
  uint64_t row_base = 0;  // index of in_keys[0] in the caller's column (for Ctl::sentinel_row)
  // stage 2 input: the fixed-capacity partitions written by stage 1
  const void* in_part = nullptr;
  const uint32_t* in_counts = nullptr;
  uint32_t in_nparts = 0;
  uint64_t in_cap = 0;
  uint64_t n_upper = 0;  // upper bound on the rows stage 2 will see (for grid sizing)
  bool merge = false;    // stage 2: output partition = digit only (input partitions are merged, not refined)
  // output partitions: partition p occupies [p*out_cap, p*out_cap + cursor[p])
  void* out = nullptr;
  uint32_t* out_cursor = nullptr;
  uint64_t out_cap = 0;
  int shift = 0;      // >= 0: digit = (hash32(key) >> shift) & (fan - 1), fan a power of two
                      // < 0 : shuffle destination = ((hash32(key) & 0xffff) * fan) >> 16, any fan
  uint32_t fan = 1;   // <= 256
  Ctl* ctl = nullptr;
  DomainArgs dom;
};
struct JoinArgs {
  const void* build = nullptr;
  const uint32_t* bcnt = nullptr;
  uint64_t cap_b = 0;
  const void* probe = nullptr;
  const uint32_t* pcnt = nullptr;
  uint64_t cap_p = 0;
  uint32_t smax = 0;        // build tuples that fit the shared-memory staging area (multiple of 4)
  uint32_t tcap = 0;        // slots of the shared-memory index table
  uint32_t bloom_words = 0; // k_join: words of the per-partition Bloom filter behind the table (0 = no filter)
  uint32_t chunk = 0;       // probe rows per CTA (multiple of 4)
  uint32_t max_chunks = 1;  // ceil(cap_p / chunk)
  uint32_t nparts = 0;
  Ctl* ctl = nullptr;
  unsigned long long* out_keys = nullptr;
  unsigned long long* out_vals = nullptr;
};
size_t radix_elem_bytes(bool build, bool narrow);
uint32_t scatter_tile_rows(bool build, bool narrow);  // rows per tile of the pipelined scatter
uint32_t scatter_pad_rows(bool build, bool narrow);   // max holes per (tile, partition) run
void launch_scatter(bool build, bool narrow, int stage, const ScatterArgs& a, const DeviceInfo& di, cudaStream_t st,
                    int* launches);
void launch_join(bool narrow, bool mat, const JoinArgs& a, cudaStream_t st, int* launches);
// collision-free pipelined join of packed (narrow) partitions (bitmap + rank directory over the 32 - bits
// hash bits the radix passes did not consume).  JoinArgs::tcap / ::chunk are ignored; ::max_chunks =
// ceil(cap_p / join3_probe_chunk()); requires join3_min_rbits() <= rbits and smax <= join3_max_build_rows()
void launch_join3(bool mat, const JoinArgs& a, int rbits, const DeviceInfo& di, cudaStream_t st, int* launches);
size_t join3_smem_bytes(uint32_t smax, int rbits);
uint32_t join3_probe_chunk();
uint32_t join3_max_build_rows();
int join3_min_rbits();
// dense key domain: direct-address join of the 256 stage-1 partitions made with DomainArgs::ident (k_djoin).
// `direct` holds 256 regions of `rstride` 4-byte slots; `sync` = djoin_sync_words() zeroed words; returns false
// when the kernel cannot be made fully co-resident (nothing is launched then).  cap_p must be a multiple of 8.
struct DjoinArgs {
  const void* build = nullptr;
  const uint32_t* bcnt = nullptr;
  uint64_t cap_b = 0;
  const void* probe = nullptr;
  const uint32_t* pcnt = nullptr;
  uint64_t cap_p = 0;
  uint32_t* direct = nullptr;
  uint64_t rstride = 0;
  uint64_t group_bytes = 0;  // bytes of regions per pipeline stage (delay_b + delay_p + 1 stages are live at once)
  uint32_t ring = 4, batch = 2;        // per-CTA item look-ahead: published-item ring slots, tickets per dispatcher round trip
  uint32_t delay_b = 1, delay_p = 1;   // pipeline distance in steps: zero -> fill, fill -> probe
  Ctl* ctl = nullptr;
  uint32_t* sync = nullptr;
  unsigned long long* out_keys = nullptr;
  unsigned long long* out_vals = nullptr;
};
size_t djoin_sync_words();
uint32_t djoin_fan();
bool launch_djoin(bool mat, const DjoinArgs& a, const DeviceInfo& di, cudaStream_t st, int* launches);
// dense key domain, round 2 (fj_part.cu): ONE partition pass by the low `logp` key bits with per-SM write-combining
// sector buffers (k_part), rows reduced to idx = key >> logp (build: idx | value << 16, 4 bytes; probe: idx, 2 bytes),
// then a direct-address join whose region lives in shared memory (k_sjoin).  The same kernels serve the multi-GPU
// shuffle: every GPU partitions its slice into its OWN (IPC-mapped) partition buffer, and the owner of a partition
// pulls its rows from every source's buffer with the bulk copies that feed k_sjoin's input ring.
struct PartArgs {
  const unsigned long long* in_keys = nullptr;
  const unsigned long long* in_vals = nullptr;  // val == true only
  uint64_t n = 0;
  uint64_t klimit = 0;   // keys >= klimit are outside the domain (strict: CTL_NOT_DENSE16; else the row is dropped)
  uint64_t cap = 0;      // elements per partition, multiple of 16
  uint32_t* cursor = nullptr;  // [2^logp * cursor_stride], starting at part_cursor_start(): elements reserved per partition
  uint32_t cursor_stride = 1;  // 32-bit words between two cursors
  Ctl* ctl = nullptr;
  void* out = nullptr;   // partition buffer: partition d at elements [d * cap, d * cap + cursor[d])
  int logp = 11;         // log2(partitions)
  bool strict = false;
  bool direct_in = false;   // input rows by 128-bit loads straight into registers instead of the per-warp TMA rings
};
size_t part_smem_bytes(int logp);
uint32_t part_sector_elems(bool val);
uint32_t part_cursor_stride();
uint32_t part_grid(bool val, uint64_t n, const DeviceInfo& di);
// every CTA of a pass starts with sector number blockIdx.x of every partition: the cursors must start at this value
// (0 when the pass is not launched at all, n == 0)
uint32_t part_cursor_start(bool val, uint64_t n, const DeviceInfo& di);
bool launch_part(bool val, const PartArgs& a, const DeviceInfo& di, cudaStream_t st, int* launches);
// match-rate sample in front of a dense16 attempt (k_sel_sample): `area` holds sel_sample_bytes(bits) bytes prepared as
// all-ones; fewer than min_pct % sampled probe keys in the build set raise CTL_NOT_DENSE16 | CTL_LOW_SEL, provided every
// build key is below table_bits (the domain of the path taken instead)
size_t sel_sample_bytes(uint64_t bits);
void launch_sel_sample(Ctl* ctl, const unsigned long long* bk, uint64_t nb, const unsigned long long* pk, uint64_t np, void* area,
                       uint64_t bits, uint64_t table_bits, uint32_t min_pct, const DeviceInfo& di, cudaStream_t st, int* launches);
struct SjoinArgs {
  const void* build[8] = {};       // [nsub] every source's partition buffer (peer mapped for remote sources): partition p at
                                   // elements [p * cap_b, +bcnt); mat: 4-byte idx | value << 16; count: 2-byte idx
  const uint32_t* bcnt = nullptr;  // bcnt[sub * cnt_stride + p]
  uint64_t cap_b = 0;
  const void* probe[8] = {};       // 2-byte idx
  const uint32_t* pcnt = nullptr;
  uint64_t cap_p = 0;
  uint32_t cnt_stride = 0;
  uint32_t cursor_stride = 1;         // cursor of (sub, p) = cnt[(sub * cnt_stride + p) * cursor_stride]
  uint32_t p_first = 0, p_count = 0;  // global ids of the partitions joined by this GPU
  int logp = 11, nsub = 1;
  uint32_t rot = 0;                   // rotation of the order in which the sources of a partition are visited (rank + 1)
  uint32_t slots_alloc = 0;           // shared-memory direct-address slots (>= klimit >> logp), multiple of 8, <= sjoin_max_slots
  Ctl* ctl = nullptr;
  unsigned long long* out_keys = nullptr;
  unsigned long long* out_vals = nullptr;
  unsigned long long* tails = nullptr;  // mat: sjoin_tail_bytes() of scratch (per-CTA output tails for the pair compaction)
};
size_t sjoin_smem_bytes(uint32_t slots_alloc);
uint32_t sjoin_max_slots(const DeviceInfo& di);
size_t sjoin_tail_bytes(const DeviceInfo& di);
// a materializing k_sjoin reserves output in blocks: out_keys / out_vals need room for this many pairs beyond np; after
// launch_sjoin the pairs are the dense range [0, Ctl::match_count) (Ctl::out_cursor counts the blocks reserved)
uint64_t sjoin_out_slack_pairs(const DeviceInfo& di);
bool launch_sjoin(bool mat, const SjoinArgs& a, const DeviceInfo& di, cudaStream_t st, int* launches);

// cross-GPU steps of the peer-memory shuffle (k_xsync, fj_part.cu): phase 1 count push + slice-size check + barrier,
// phase 2 result exchange.  ctrl[r] = rank r's exchange area (peer mapped), whose count
// arrays (uint32 [2 sides][world][P / world]) start at xsync_count_offset_bytes()
struct XsyncArgs {
  void* ctrl[8] = {};
  int rank = 0, world = 1, phase = 0;
  unsigned long long seq = 0;
  Ctl* ctl = nullptr;
  unsigned long long nb = 0, np = 0;
  unsigned long long meta[16] = {};
  const uint32_t* cur_b = nullptr;
  const uint32_t* cur_p = nullptr;
  uint32_t cursor_stride = 1, P = 0, lpo = 0;
  unsigned long long* result = nullptr;
};
size_t xsync_ctrl_bytes(uint32_t P);
size_t xsync_count_offset_bytes();
void launch_xsync(const XsyncArgs& a, cudaStream_t st, int* launches);

void launch_emit_sentinel(Ctl* ctl, const unsigned long long* bv, unsigned long long* out_keys,
                          unsigned long long* out_vals, bool mat, cudaStream_t st, int* launches);

void launch_emit_sentinel_value(Ctl* ctl, unsigned long long value, unsigned long long n, unsigned long long* out_keys,
                                unsigned long long* out_vals, bool mat, cudaStream_t st, int* launches);
// partition-element rows (radix_elem_bytes format) -> raw 64-bit columns, holes dropped; *cursor (zeroed by the
// caller) receives the number of rows written
void launch_expand(bool build, bool narrow, const void* in, uint64_t n, unsigned long long* keys, unsigned long long* vals,
                   unsigned long long* cursor, const DeviceInfo& di, cudaStream_t st, int* launches);
// destination digit of a key in the multi-GPU shuffle (host mirror of the device function)
uint32_t shuffle_dest_host(uint64_t key, uint32_t fan);

// ---------------------------------------------------------------- synthetic data + utilities
void launch_generate_g2(int side, uint64_t ny, uint64_t c, uint64_t U, uint64_t a_mod_u, uint64_t b, uint64_t seed,
                        uint64_t start, uint64_t count, unsigned long long* keys, unsigned long long* vals,
                        cudaStream_t st);

}  // namespace fj
