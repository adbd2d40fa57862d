// ubench2.cu — microbenchmarks behind DESIGN.md §8 item 1 (C3 on one GPU: where should the random accesses of the
// partition join live?).  Not product code.  Three candidate homes for a partition's direct-address region:
//   E  L2       : random 4-byte stores / loads into an L2-resident region (what k_djoin does today; bound = tag lookups)
//   F  DSMEM    : the region spread over the shared memory of an 8-CTA thread-block cluster (8 x ~200 KB = one C3 region),
//                 random 4-byte remote stores / loads through the SM-to-SM network
//   G  smem     : the region in ONE CTA's shared memory (needs 2^11..2^12 partitions): random 4-byte stores / loads
// plus H, the cost side of G: ranking rows into 2048 / 4096 bins with shared-memory atomics (one high-fan-out pass).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/ubench2.cu -o ubench2
// Run (GPU box): ./ubench2 > gpurun_out/ubench2.jsonl        (one JSON line per measurement, rates in G accesses/s)
#include <cooperative_groups.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a)); }
  float stop() { CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }
};

constexpr int IPT = 8;  // independent accesses in flight per thread, as in k_djoin

// ---------------------------------------------------------------- E: random 4-byte accesses, L2-resident region
template <bool STORE>
__global__ void __launch_bounds__(512, 2) k_l2_rand(uint32_t* __restrict__ region, uint32_t nslots, uint64_t nacc, uint32_t* sink) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * IPT;
  uint32_t acc = 0;
  for (uint64_t base = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * IPT; base < nacc; base += stride) {
    uint32_t idx[IPT], v[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) idx[i] = (uint32_t)(((uint64_t)mix32((uint32_t)(base + i) * 2654435761u + 12345u) * nslots) >> 32);
    if (STORE) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) region[idx[i]] = idx[i] + 1u;
    } else {
#pragma unroll
      for (int i = 0; i < IPT; ++i) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v[i]) : "l"(region + idx[i]));
#pragma unroll
      for (int i = 0; i < IPT; ++i) acc += v[i];
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// ---------------------------------------------------------------- F: random 4-byte accesses, region in cluster DSMEM
constexpr int CL = 8;                   // CTAs per cluster
constexpr int F_SLOTS = 49152;          // 4-byte slots per CTA (192 KB) -> 1.5 MB region per cluster
template <bool STORE>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(1024, 1) k_dsmem_rand(int reps, uint32_t* sink) {
  extern __shared__ uint32_t s_region[];
  cg::cluster_group cluster = cg::this_cluster();
  for (int i = threadIdx.x; i < F_SLOTS; i += blockDim.x) s_region[i] = 0u;
  cluster.sync();
  const uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 40503u;
  uint32_t acc = 0;
  for (int rep = 0; rep < reps; ++rep) {
    uint32_t v[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t h = mix32(seed + rep * IPT + i);
      const uint32_t slot = (uint32_t)(((uint64_t)h * (uint32_t)(CL * F_SLOTS)) >> 32);  // uniform over the cluster's region
      uint32_t* p = cluster.map_shared_rank(s_region + slot % F_SLOTS, slot / F_SLOTS);  // mapa: address arithmetic only
      if (STORE) *p = h | 1u;
      else v[i] = *reinterpret_cast<volatile uint32_t*>(p);
    }
    if (!STORE) {
#pragma unroll
      for (int i = 0; i < IPT; ++i) acc += v[i];
    }
  }
  cluster.sync();  // no CTA may exit while others still access its shared memory
  if (acc == 0x12345678u) *sink = acc;
}

// ---------------------------------------------------------------- G: random 4-byte accesses, region in the CTA's own smem
template <bool STORE>
__global__ void __launch_bounds__(512, 2) k_smem_rand(int reps, uint32_t nslots, uint32_t* sink) {
  extern __shared__ uint32_t s_region[];
  for (uint32_t i = threadIdx.x; i < nslots; i += blockDim.x) s_region[i] = 0u;
  __syncthreads();
  const uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 40503u;
  uint32_t acc = 0;
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      const uint32_t h = mix32(seed + rep * IPT + i);
      const uint32_t slot = (uint32_t)(((uint64_t)h * nslots) >> 32);
      if (STORE) s_region[slot] = h | 1u;
      else acc += reinterpret_cast<volatile uint32_t*>(s_region)[slot];
    }
  }
  __syncthreads();
  if (acc == 0x12345678u) *sink = acc;
}

// ---------------------------------------------------------------- H: ranking rows into FAN bins (smem atomics) + staged scatter
template <int FAN>
__global__ void __launch_bounds__(1024, 1) k_rank_hf(int tiles, uint32_t* sink) {
  extern __shared__ uint32_t s_mem[];  // [FAN] histogram, then the staging area
  uint32_t* s_hist = s_mem;
  uint32_t* s_stage = s_mem + FAN;
  constexpr int ROWS = 1024 * IPT;  // rows per tile
  const uint32_t seed = (blockIdx.x * blockDim.x + threadIdx.x) * 40503u;
  uint32_t acc = 0;
  for (int t = 0; t < tiles; ++t) {
    for (int i = threadIdx.x; i < FAN; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    uint32_t d[IPT], r[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
      d[i] = mix32(seed + t * IPT + i) & (FAN - 1);
      r[i] = atomicAdd(&s_hist[d[i]], 1u);
    }
    __syncthreads();
    // (the exclusive scan over FAN bins is left out: it is O(FAN) per tile, not per row) — staged write by rank
#pragma unroll
    for (int i = 0; i < IPT; ++i) s_stage[(d[i] * (ROWS / FAN) + r[i]) & (ROWS - 1)] = d[i];
    __syncthreads();
    acc += s_stage[threadIdx.x];
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  uint32_t* sink;
  CK(cudaMalloc(&sink, 4));
  Timer tm;

  // E
  for (size_t mb : {8, 16, 32, 48, 64, 96}) {
    const uint32_t nslots = (uint32_t)((mb << 20) / 4);
    uint32_t* region;
    CK(cudaMalloc(&region, (size_t)nslots * 4));
    CK(cudaMemset(region, 0, (size_t)nslots * 4));
    const uint64_t nacc = 400000000ull;
    for (int store = 0; store < 2; ++store) {
      float best = 1e30f;
      for (int it = 0; it < 3; ++it) {
        tm.start();
        if (store) k_l2_rand<true><<<sms * 2, 512>>>(region, nslots, nacc, sink);
        else k_l2_rand<false><<<sms * 2, 512>>>(region, nslots, nacc, sink);
        const float ms = tm.stop();
        CK(cudaGetLastError());
        if (ms < best) best = ms;
      }
      printf("{\"bench\": \"E_l2_rand4\", \"op\": \"%s\", \"region_mb\": %zu, \"ms\": %.4f, \"G_per_s\": %.1f}\n", store ? "store" : "load", mb,
             best, nacc / best * 1e-6);
    }
    CK(cudaFree(region));
  }

  // F
  {
    const size_t smem = (size_t)F_SLOTS * 4;
    CK(cudaFuncSetAttribute(k_dsmem_rand<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_dsmem_rand<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (sms / CL) * CL;
    const int reps = 2000;
    for (int store = 0; store < 2; ++store) {
      float best = 1e30f;
      for (int it = 0; it < 3; ++it) {
        tm.start();
        if (store) k_dsmem_rand<true><<<grid, 1024, smem>>>(reps, sink);
        else k_dsmem_rand<false><<<grid, 1024, smem>>>(reps, sink);
        const float ms = tm.stop();
        CK(cudaGetLastError());
        if (ms < best) best = ms;
      }
      const double nacc = (double)grid * 1024 * IPT * reps;
      printf("{\"bench\": \"F_dsmem_rand4\", \"op\": \"%s\", \"cluster\": %d, \"ctas\": %d, \"ms\": %.4f, \"G_per_s\": %.1f}\n", store ? "store" : "load",
             CL, grid, best, nacc / best * 1e-6);
    }
  }

  // G
  for (uint32_t kb : {32u, 64u, 100u}) {
    const uint32_t nslots = kb * 256u;
    const size_t smem = (size_t)nslots * 4;
    CK(cudaFuncSetAttribute(k_smem_rand<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_smem_rand<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int reps = 4000;
    for (int store = 0; store < 2; ++store) {
      float best = 1e30f;
      for (int it = 0; it < 3; ++it) {
        tm.start();
        if (store) k_smem_rand<true><<<sms * 2, 512, smem>>>(reps, nslots, sink);
        else k_smem_rand<false><<<sms * 2, 512, smem>>>(reps, nslots, sink);
        const float ms = tm.stop();
        CK(cudaGetLastError());
        if (ms < best) best = ms;
      }
      const double nacc = (double)sms * 2 * 512 * IPT * reps;
      printf("{\"bench\": \"G_smem_rand4\", \"op\": \"%s\", \"region_kb\": %u, \"ms\": %.4f, \"G_per_s\": %.1f}\n", store ? "store" : "load", kb,
             best, nacc / best * 1e-6);
    }
  }

  // H
  {
    const int tiles = 400;
    {
      const size_t smem = (size_t)(2048 + 1024 * IPT) * 4;
      CK(cudaFuncSetAttribute(k_rank_hf<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tm.start();
      k_rank_hf<2048><<<sms, 1024, smem>>>(tiles, sink);
      const float ms = tm.stop();
      CK(cudaGetLastError());
      printf("{\"bench\": \"H_rank_hf\", \"fan\": 2048, \"ms\": %.4f, \"G_rows_per_s\": %.1f}\n", ms, (double)sms * 1024 * IPT * tiles / ms * 1e-6);
    }
    {
      const size_t smem = (size_t)(4096 + 1024 * IPT) * 4;
      CK(cudaFuncSetAttribute(k_rank_hf<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tm.start();
      k_rank_hf<4096><<<sms, 1024, smem>>>(tiles, sink);
      const float ms = tm.stop();
      CK(cudaGetLastError());
      printf("{\"bench\": \"H_rank_hf\", \"fan\": 4096, \"ms\": %.4f, \"G_rows_per_s\": %.1f}\n", ms, (double)sms * 1024 * IPT * tiles / ms * 1e-6);
    }
  }
  return 0;
}
