// Host turn-around of one short call: how the result word gets back to the host.
//   (a) cudaMemcpyAsync of an 80-byte control block + cudaStreamSynchronize (what the engine does)
//   (b) a one-thread kernel stores the block into mapped pinned memory, the host spins on a sequence word
//   (c) the same store, the host calls cudaStreamSynchronize
// around a kernel of ~10 us, with the two cudaEventRecord calls the engine brackets an attempt with.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/ubench3.cu -o ubench3 && ./ubench3
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
struct Blk { unsigned long long w[10]; };
__global__ void k_work(unsigned long long* p, int spin) {
  unsigned long long t0 = clock64();
  while (clock64() - t0 < (unsigned long long)spin) {}
  if (threadIdx.x == 0 && blockIdx.x == 0) p[0] += 1;
}
__global__ void k_publish(const Blk* src, volatile Blk* dst, unsigned long long seq) {
  for (int i = 1; i < 10; ++i) dst->w[i] = src->w[i];
  __threadfence_system();
  dst->w[0] = seq;
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  Blk *d, *h, *hm, *dm;
  cudaMalloc(&d, sizeof(Blk)); cudaMemset(d, 0, sizeof(Blk));
  cudaMallocHost(&h, sizeof(Blk));
  cudaHostAlloc(&hm, sizeof(Blk), cudaHostAllocMapped); cudaHostGetDevicePointer(&dm, hm, 0);
  hm->w[0] = 0;
  const int reps = 2000, spin = 20000;  // ~10 us at 1.9 GHz
  for (int mode = 0; mode < 3; ++mode) {
    for (int pass = 0; pass < 2; ++pass) {
      cudaStreamSynchronize(st);
      const double t0 = now();
      float dev_ms = 0.f;
      for (int i = 0; i < reps; ++i) {
        const unsigned long long seq = (unsigned long long)(mode * 2 + pass) * reps + i + 1;
        cudaEventRecord(e0, st);
        k_work<<<148, 256, 0, st>>>(d->w, spin);
        cudaEventRecord(e1, st);
        if (mode == 0) {
          cudaMemcpyAsync(h, d, sizeof(Blk), cudaMemcpyDeviceToHost, st);
          cudaStreamSynchronize(st);
        } else if (mode == 1) {
          k_publish<<<1, 1, 0, st>>>(d, dm, seq);
          while (reinterpret_cast<volatile Blk*>(hm)->w[0] != seq) {}
        } else {
          k_publish<<<1, 1, 0, st>>>(d, dm, seq);
          cudaStreamSynchronize(st);
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) { cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
        dev_ms += ms;
      }
      const double wall = (now() - t0) / reps;
      if (pass == 1)
        printf("{\"mode\": \"%s\", \"wall_us_per_call\": %.2f, \"kernel_us\": %.2f, \"turnaround_us\": %.2f}\n",
               mode == 0 ? "memcpyAsync + streamSynchronize" : mode == 1 ? "publish kernel + host spin on mapped word" : "publish kernel + streamSynchronize",
               wall * 1e6, dev_ms / reps * 1e3, wall * 1e6 - dev_ms / reps * 1e3);
    }
  }
  return 0;
}
