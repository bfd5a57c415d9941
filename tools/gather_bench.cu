// Microbenchmark: how fast can ONE SM pull scattered 128-byte row segments into shared memory?
// Variants: 16-byte cp.async (.cg / .ca), in-flight depth (commit groups), CTAs per SM, locality of the rows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_bench tools/gather_bench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

template <int CA>
__device__ __forceinline__ void cp16(uint32_t dst, const void* src) {
  if (CA) asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// each iteration: the CTA's 256 threads copy 32 rows x 128 B (8 lanes per row) = 4 KB; DEPTH groups in flight
template <int CA, int DEPTH>
__global__ void gather_kernel(const char* __restrict__ x, int64_t row_bytes, const int* __restrict__ rows, int iters,
                              int nrows_tab) {
  extern __shared__ __align__(128) char sm[];
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  const int q = threadIdx.x & 7, r = threadIdx.x >> 3;     // 32 rows per iteration
  int pos = (blockIdx.x * 7919) % nrows_tab;
  for (int it = 0; it < iters; ++it) {
    // row index computed arithmetically (no dependent table load): nrows_tab < 0 selects the pattern
    int row;
    if (nrows_tab == -1) row = (int)(((unsigned)(it * 32 + r) * 9973u + blockIdx.x * 7919u) % 900000u);        // scattered
    else if (nrows_tab == -2) row = (int)((blockIdx.x * 4000u + (unsigned)(it * 32 + r)) % 900000u);            // streaming
    else { row = rows[(pos + r) % nrows_tab]; pos = (pos + 32) % nrows_tab; }
    cp16<CA>(sbase + ((it % DEPTH) * 4096) + r * 128 + q * 16, x + (int64_t)row * row_bytes + q * 16);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
  const int64_t nrows = 900000, row_bytes = 256;            // the d6 C=128 bf16 tensor
  char* x; cudaMalloc(&x, nrows * row_bytes); cudaMemset(x, 1, nrows * row_bytes);
  const int ntab = 1 << 22;
  std::vector<int> h(ntab);
  int* d; cudaMalloc(&d, ntab * sizeof(int));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int pattern = 3; pattern < 5; ++pattern) {
    srand(1);
    for (int i = 0; i < ntab; ++i) {
      if (pattern == 0) h[i] = (int)(((int64_t)rand() * 7 + rand()) % nrows);                    // uniformly random rows
      else if (pattern == 1) h[i] = (int)(((i / 7) + (rand() % 64) - 32 + nrows) % nrows);      // local: +-32 rows of a walking base
      else h[i] = i % nrows;                                                                     // sequential
    }
    cudaMemcpy(d, h.data(), ntab * sizeof(int), cudaMemcpyHostToDevice);
    const int tabarg = pattern == 3 ? -1 : (pattern == 4 ? -2 : ntab);
    for (int cpb = 1; cpb <= 4; cpb *= 2) {
      for (int variant = 0; variant < 6; ++variant) {
        const int iters = 4000;
        const int grid = sms * cpb;
        float ms = 0;
        int depth = 0, ca = 0;
#define RUN(CA, DEPTH)                                                                                 \
  {                                                                                                    \
    depth = DEPTH; ca = CA;                                                                            \
    cudaFuncSetAttribute(gather_kernel<CA, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * 4096); \
    gather_kernel<CA, DEPTH><<<grid, 256, DEPTH * 4096>>>(x, row_bytes, d, 100, tabarg);                 \
    cudaEventRecord(e0);                                                                               \
    gather_kernel<CA, DEPTH><<<grid, 256, DEPTH * 4096>>>(x, row_bytes, d, iters, tabarg);               \
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);                  \
  }
        if (variant == 0) RUN(0, 2) else if (variant == 1) RUN(0, 4) else if (variant == 2) RUN(0, 8)
        else if (variant == 3) RUN(0, 12) else if (variant == 4) RUN(1, 8) else RUN(1, 12)
        if (cudaGetLastError() != cudaSuccess) { printf("launch error\n"); continue; }
        const double bytes = (double)grid * iters * 4096.0;
        printf("pattern %d  CTAs/SM %d  %s depth %2d (%2d KB in flight/CTA): %7.1f GB/s total, %6.2f KB/us per SM\n", pattern,
               cpb, ca ? ".ca" : ".cg", depth, depth * 4, bytes / ms / 1e6, bytes / ms / 1e6 / sms * 1e3 / 1e3);
      }
    }
  }
  return 0;
}
