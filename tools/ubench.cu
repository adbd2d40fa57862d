// ubench.cu — B200 design-input microbenchmarks for the join engine.
// Not product code: measures the primitive rates the kernel designs in DESIGN.md are
// budgeted against (HBM stream, random 32 B sector reads from L2/HBM, 64-bit atomicCAS
// inserts into L2/HBM tables, shared-memory CAS / store-verify build rates, warp match
// ranking rates).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo
// Run (GPU box): ./ubench > gpurun_out/ubench.jsonl
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a)); }
  float stop() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

// ---------------------------------------------------------------- A: stream
__global__ void k_copy(const uint4* __restrict__ in, uint4* __restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void k_read(const uint4* __restrict__ in, uint64_t* sink, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (; i + 3 * stride < n; i += 4 * stride) {
    uint4 a = in[i], b = in[i + stride], c = in[i + 2 * stride], d = in[i + 3 * stride];
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  for (; i < n; i += stride) acc += in[i].x;
  if (acc == 0x12345678) *sink = acc;
}

// ---------------------------------------------------------------- B: random sector reads
// mode 0: 8 B load, mode 1: 32 B (256-bit) load of the whole sector
template <int MODE>
__global__ void k_rand_read(const uint64_t* __restrict__ tab, uint64_t nsect, uint64_t nacc, uint64_t* sink) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
  for (; i < nacc; i += stride) {
    uint64_t h = mix64(i + 0x9e3779b97f4a7c15ULL);
    uint64_t s = __umul64hi(h, nsect);
    const uint64_t* p = tab + s * 4;
    if (MODE == 0) {
      acc += __ldg(p);
    } else {
      uint64_t a, b, c, d;
      asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
      acc += a ^ b ^ c ^ d;
    }
  }
  if (acc == 0x12345678) *sink = acc;
}

// ---------------------------------------------------------------- C: global CAS inserts
__global__ void k_cas_insert(unsigned long long* tab, uint64_t nslots, uint64_t nkeys, unsigned long long* fails) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long nf = 0;
  for (; i < nkeys; i += stride) {
    uint64_t key = i + 1;
    uint64_t h = mix64(key);
    uint64_t s = __umul64hi(h, nslots);
    while (true) {
      unsigned long long old = atomicCAS(tab + s, ~0ULL, (unsigned long long)key);
      if (old == ~0ULL) break;
      ++nf;
      if (++s == nslots) s = 0;
    }
  }
  if (nf) atomicAdd(fails, nf);
}
// 16 B slot: CAS on key then store value
__global__ void k_cas_insert16(unsigned long long* tab, uint64_t nslots, uint64_t nkeys) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < nkeys; i += stride) {
    uint64_t key = i + 1;
    uint64_t h = mix64(key);
    uint64_t s = __umul64hi(h, nslots);
    while (true) {
      unsigned long long old = atomicCAS(tab + 2 * s, ~0ULL, (unsigned long long)key);
      if (old == ~0ULL) { tab[2 * s + 1] = key * 3; break; }
      if (++s == nslots) s = 0;
    }
  }
}

// ---------------------------------------------------------------- D: shared memory build variants
// Each CTA repeatedly builds a table of NSLOT 8-byte slots from NKEY keys.
constexpr int SM_SLOTS = 20480;   // 160 KB
constexpr int SM_KEYS = 10240;
__global__ void __launch_bounds__(1024, 1) k_smem_cas(int reps, unsigned long long* sink) {
  extern __shared__ unsigned long long tab[];
  unsigned long long acc = 0;
  for (int r = 0; r < reps; ++r) {
    for (int i = threadIdx.x; i < SM_SLOTS; i += blockDim.x) tab[i] = ~0ULL;
    __syncthreads();
    for (int i = threadIdx.x; i < SM_KEYS; i += blockDim.x) {
      uint32_t key = (blockIdx.x * 131071u + r * 8191u) * 16384u + i + 1;
      uint32_t h = mix32(key);
      uint32_t s = __umulhi(h, SM_SLOTS);
      unsigned long long kv = ((unsigned long long)key << 32) | i;
      while (true) {
        unsigned long long old = atomicCAS(tab + s, ~0ULL, kv);
        if (old == ~0ULL) break;
        if (++s == SM_SLOTS) s = 0;
      }
    }
    __syncthreads();
    acc += tab[threadIdx.x];
    __syncthreads();
  }
  if (acc == 0x12345) *sink = acc;
}
// store-verify rounds (no atomics): store if empty, barrier, verify, barrier
__global__ void __launch_bounds__(1024, 1) k_smem_storeverify(int reps, unsigned long long* sink, unsigned* rounds_out) {
  extern __shared__ unsigned long long tab[];
  __shared__ int pending;
  unsigned long long acc = 0;
  constexpr int PER = SM_KEYS / 1024;  // 10
  unsigned total_rounds = 0;
  for (int r = 0; r < reps; ++r) {
    for (int i = threadIdx.x; i < SM_SLOTS; i += blockDim.x) tab[i] = ~0ULL;
    uint32_t slot[PER]; unsigned long long kv[PER]; bool done[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      int i = threadIdx.x + j * 1024;
      uint32_t key = (blockIdx.x * 131071u + r * 8191u) * 16384u + i + 1;
      slot[j] = __umulhi(mix32(key), SM_SLOTS);
      kv[j] = ((unsigned long long)key << 32) | i;
      done[j] = false;
    }
    __syncthreads();
    while (true) {
      if (threadIdx.x == 0) pending = 0;
      bool wrote[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        wrote[j] = false;
        if (!done[j]) { if (tab[slot[j]] == ~0ULL) { tab[slot[j]] = kv[j]; wrote[j] = true; } }
      }
      __syncthreads();
      int mypend = 0;
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        if (!done[j]) {
          if (wrote[j] && tab[slot[j]] == kv[j]) done[j] = true;
          else { if (++slot[j] == SM_SLOTS) slot[j] = 0; mypend = 1; }
        }
      }
      if (__any_sync(0xffffffffu, mypend) && (threadIdx.x & 31) == 0) pending = 1;
      __syncthreads();
      ++total_rounds;
      if (!pending) break;
      __syncthreads();
    }
    acc += tab[threadIdx.x];
    __syncthreads();
  }
  if (acc == 0x12345) *sink = acc;
  if (blockIdx.x == 0 && threadIdx.x == 0) *rounds_out = total_rounds;
}
// random 8 B LDS probes into the smem table
__global__ void __launch_bounds__(1024, 1) k_smem_probe(int reps, unsigned long long* sink) {
  extern __shared__ unsigned long long tab[];
  for (int i = threadIdx.x; i < SM_SLOTS; i += blockDim.x) tab[i] = mix32(i);
  __syncthreads();
  unsigned long long acc = 0;
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (int i = threadIdx.x; i < SM_KEYS; i += blockDim.x) {
      uint32_t key = (blockIdx.x * 131071u + r * 8191u) * 16384u + i + 1;
      uint32_t s = __umulhi(mix32(key), SM_SLOTS);
      acc += tab[s];
    }
  }
  if (acc == 0x12345) *sink = acc;
}
// smem atomicAdd histogram rank (256 bins), 16 items per thread
__global__ void __launch_bounds__(512, 2) k_smem_hist_atomic(int reps, unsigned* sink) {
  __shared__ unsigned hist[256];
  unsigned acc = 0;
  for (int r = 0; r < reps; ++r) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t key = (blockIdx.x * 977u + r) * 8192u + j * 512 + threadIdx.x;
      uint32_t d = mix32(key) >> 24;
      acc += atomicAdd(&hist[d], 1u);
    }
    __syncthreads();
  }
  if (acc == 0x12345) *sink = acc;
}
// match_any ranking with warp-private counters
__global__ void __launch_bounds__(512, 2) k_rank_match(int reps, unsigned* sink) {
  __shared__ unsigned cnt[16][256];
  unsigned acc = 0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < reps; ++r) {
    for (int i = lane; i < 256; i += 32) cnt[w][i] = 0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t key = (blockIdx.x * 977u + r) * 8192u + j * 512 + threadIdx.x;
      uint32_t d = mix32(key) >> 24;
      unsigned peers = __match_any_sync(0xffffffffu, d);
      unsigned lt = peers & ((1u << lane) - 1);
      int leader = __ffs(peers) - 1;
      unsigned base = 0;
      if (lane == leader) { base = cnt[w][d]; cnt[w][d] = base + __popc(peers); }
      base = __shfl_sync(0xffffffffu, base, leader);
      acc += base + __popc(lt);
      __syncwarp();
    }
  }
  if (acc == 0x12345) *sink = acc;
}
// ballot-emulated match (8 ballots)
__global__ void __launch_bounds__(512, 2) k_rank_ballot(int reps, unsigned* sink) {
  __shared__ unsigned cnt[16][256];
  unsigned acc = 0;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = 0; r < reps; ++r) {
    for (int i = lane; i < 256; i += 32) cnt[w][i] = 0;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t key = (blockIdx.x * 977u + r) * 8192u + j * 512 + threadIdx.x;
      uint32_t d = mix32(key) >> 24;
      unsigned peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1);
        peers &= ((d >> b) & 1) ? bal : ~bal;
      }
      unsigned lt = peers & ((1u << lane) - 1);
      int leader = __ffs(peers) - 1;
      unsigned base = 0;
      if (lane == leader) { base = cnt[w][d]; cnt[w][d] = base + __popc(peers); }
      base = __shfl_sync(0xffffffffu, base, leader);
      acc += base + __popc(lt);
      __syncwarp();
    }
  }
  if (acc == 0x12345) *sink = acc;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("{\"bench\":\"device\",\"name\":\"%s\",\"sms\":%d,\"l2_bytes\":%d,\"smem_optin\":%zu,\"clock_khz\":%d}\n",
         prop.name, sms, prop.l2CacheSize, prop.sharedMemPerBlockOptin, prop.clockRate);
  Timer t;
  uint64_t* sink; CK(cudaMalloc(&sink, 64)); CK(cudaMemset(sink, 0, 64));

  // A: stream
  {
    size_t bytes = 2ull << 30; uint4 *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    size_t n = bytes / 16;
    for (int it = 0; it < 3; ++it) {
      t.start(); k_copy<<<sms * 16, 512>>>(a, b, n); float ms = t.stop();
      if (it) printf("{\"bench\":\"copy\",\"GBps\":%.1f,\"ms\":%.4f}\n", 2.0 * bytes / ms * 1e-6, ms);
      t.start(); k_read<<<sms * 16, 512>>>(a, sink, n); ms = t.stop();
      if (it) printf("{\"bench\":\"read\",\"GBps\":%.1f,\"ms\":%.4f}\n", 1.0 * bytes / ms * 1e-6, ms);
    }
    CK(cudaFree(a)); CK(cudaFree(b));
  }
  // B: random sector reads
  {
    size_t sizes_mb[] = {2, 8, 32, 64, 96, 256, 4096};
    uint64_t* tab; CK(cudaMalloc(&tab, 4096ull << 20)); CK(cudaMemset(tab, 3, 4096ull << 20));
    uint64_t nacc = 200000000ull;
    for (size_t mb : sizes_mb) {
      uint64_t nsect = (mb << 20) / 32;
      for (int mode = 0; mode < 2; ++mode) {
        float best = 1e9;
        for (int it = 0; it < 3; ++it) {
          t.start();
          if (mode == 0) k_rand_read<0><<<sms * 8, 512>>>(tab, nsect, nacc, sink);
          else k_rand_read<1><<<sms * 8, 512>>>(tab, nsect, nacc, sink);
          float ms = t.stop(); if (ms < best) best = ms;
        }
        printf("{\"bench\":\"rand_read\",\"table_mb\":%zu,\"load_bytes\":%d,\"Gacc_per_s\":%.2f,\"ms\":%.4f}\n",
               mb, mode ? 32 : 8, nacc / best * 1e-6, best);
      }
    }
    CK(cudaFree(tab));
  }
  // C: CAS inserts (8 B slots), load factor 0.5
  {
    size_t sizes_mb[] = {8, 32, 64, 256, 2048};
    unsigned long long* tab; CK(cudaMalloc(&tab, 2048ull << 20));
    unsigned long long* fails; CK(cudaMalloc(&fails, 8));
    for (size_t mb : sizes_mb) {
      uint64_t nslots = (mb << 20) / 8, nkeys = nslots / 2;
      float best = 1e9, best_clear = 1e9; unsigned long long nf = 0;
      for (int it = 0; it < 3; ++it) {
        t.start(); CK(cudaMemsetAsync(tab, 0xff, mb << 20)); float msc = t.stop(); if (msc < best_clear) best_clear = msc;
        CK(cudaMemset(fails, 0, 8));
        t.start(); k_cas_insert<<<sms * 8, 512>>>(tab, nslots, nkeys, fails); float ms = t.stop(); if (ms < best) best = ms;
        CK(cudaMemcpy(&nf, fails, 8, cudaMemcpyDeviceToHost));
      }
      printf("{\"bench\":\"cas_insert8\",\"table_mb\":%zu,\"nkeys\":%llu,\"Gins_per_s\":%.2f,\"ms\":%.4f,\"clear_ms\":%.4f,\"retries\":%llu}\n",
             mb, (unsigned long long)nkeys, nkeys / best * 1e-6, best, best_clear, nf);
      uint64_t nslots16 = (mb << 20) / 16, nkeys16 = nslots16 / 2; best = 1e9;
      for (int it = 0; it < 3; ++it) {
        CK(cudaMemset(tab, 0xff, mb << 20));
        t.start(); k_cas_insert16<<<sms * 8, 512>>>(tab, nslots16, nkeys16); float ms = t.stop(); if (ms < best) best = ms;
      }
      printf("{\"bench\":\"cas_insert16\",\"table_mb\":%zu,\"nkeys\":%llu,\"Gins_per_s\":%.2f,\"ms\":%.4f}\n",
             mb, (unsigned long long)nkeys16, nkeys16 / best * 1e-6, best);
    }
    CK(cudaFree(tab)); CK(cudaFree(fails));
  }
  // D: shared memory
  {
    size_t smem = SM_SLOTS * 8;
    CK(cudaFuncSetAttribute(k_smem_cas, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_storeverify, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int reps = 20; unsigned* rounds; CK(cudaMalloc(&rounds, 4));
    for (int it = 0; it < 2; ++it) {
      t.start(); k_smem_cas<<<sms, 1024, smem>>>(reps, (unsigned long long*)sink); float ms = t.stop();
      if (it) printf("{\"bench\":\"smem_cas64_build\",\"keys_per_cta\":%d,\"us_per_table\":%.3f,\"Gins_per_s_chip\":%.2f}\n",
                     SM_KEYS, ms * 1e3 / reps, (double)SM_KEYS * reps * sms / ms * 1e-6);
      t.start(); k_smem_storeverify<<<sms, 1024, smem>>>(reps, (unsigned long long*)sink, rounds); ms = t.stop();
      unsigned hr; CK(cudaMemcpy(&hr, rounds, 4, cudaMemcpyDeviceToHost));
      if (it) printf("{\"bench\":\"smem_storeverify_build\",\"keys_per_cta\":%d,\"us_per_table\":%.3f,\"Gins_per_s_chip\":%.2f,\"rounds_per_table\":%.1f}\n",
                     SM_KEYS, ms * 1e3 / reps, (double)SM_KEYS * reps * sms / ms * 1e-6, hr / (double)reps);
      t.start(); k_smem_probe<<<sms, 1024, smem>>>(reps * 10, (unsigned long long*)sink); ms = t.stop();
      if (it) printf("{\"bench\":\"smem_probe_lds64\",\"us_per_10k\":%.3f,\"Gprobe_per_s_chip\":%.2f}\n",
                     ms * 1e3 / (reps * 10), (double)SM_KEYS * reps * 10 * sms / ms * 1e-6);
    }
    int reps2 = 200;
    for (int it = 0; it < 2; ++it) {
      t.start(); k_smem_hist_atomic<<<sms * 2, 512>>>(reps2, (unsigned*)sink); float ms = t.stop();
      if (it) printf("{\"bench\":\"rank_smem_atomic\",\"Gkeys_per_s_chip\":%.2f}\n", 8192.0 * reps2 * sms * 2 / ms * 1e-6);
      t.start(); k_rank_match<<<sms * 2, 512>>>(reps2, (unsigned*)sink); ms = t.stop();
      if (it) printf("{\"bench\":\"rank_match_any\",\"Gkeys_per_s_chip\":%.2f}\n", 8192.0 * reps2 * sms * 2 / ms * 1e-6);
      t.start(); k_rank_ballot<<<sms * 2, 512>>>(reps2, (unsigned*)sink); ms = t.stop();
      if (it) printf("{\"bench\":\"rank_ballot8\",\"Gkeys_per_s_chip\":%.2f}\n", 8192.0 * reps2 * sms * 2 / ms * 1e-6);
    }
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
