// Microbenchmark: tcgen05.ld / tcgen05.st throughput as a function of the number of warps issuing them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_bench tools/tmem_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R16(v, o) "=r"(v[o+0]), "=r"(v[o+1]), "=r"(v[o+2]), "=r"(v[o+3]), "=r"(v[o+4]), "=r"(v[o+5]), "=r"(v[o+6]), "=r"(v[o+7]), \
                  "=r"(v[o+8]), "=r"(v[o+9]), "=r"(v[o+10]), "=r"(v[o+11]), "=r"(v[o+12]), "=r"(v[o+13]), "=r"(v[o+14]), "=r"(v[o+15])
#define I16(v, o) "r"(v[o+0]), "r"(v[o+1]), "r"(v[o+2]), "r"(v[o+3]), "r"(v[o+4]), "r"(v[o+5]), "r"(v[o+6]), "r"(v[o+7]), \
                  "r"(v[o+8]), "r"(v[o+9]), "r"(v[o+10]), "r"(v[o+11]), "r"(v[o+12]), "r"(v[o+13]), "r"(v[o+14]), "r"(v[o+15])

__global__ void __launch_bounds__(512, 1) bench(int iters, int mode, int wait_each, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t taddr = tmem_base_s + (((warp & 3) * 32) << 16) + (warp >> 2) * 32;  // own lane quarter, own 32 columns
  uint32_t v[32];
  for (int j = 0; j < 32; ++j) v[j] = threadIdx.x + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                   "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : R16(v, 0), R16(v, 16) : "r"(taddr) : "memory");
      if (wait_each) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), I16(v, 0) : "memory");
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr + 16), I16(v, 16) : "memory");
      if (wait_each) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  const long long t1 = clock64();
  uint32_t s = 0;
  for (int j = 0; j < 32; ++j) s += v[j];
  sink[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512) : "memory");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMallocManaged(&out, 16); cudaMalloc(&sink, 4096);
  const int iters = 256;
  printf("%-5s %-6s %-9s %14s %12s\n", "op", "warps", "wait/op", "cyc per op", "B/cyc/SM");
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 4, 8, 16})
      for (int wait_each = 0; wait_each < 2; ++wait_each) {
        for (int rep = 0; rep < 2; ++rep) { bench<<<1, warps * 32, 0>>>(iters, mode, wait_each, out, sink); if (cudaDeviceSynchronize() != cudaSuccess) { printf("err %s\n", cudaGetErrorString(cudaGetLastError())); return 1; } }
        const double cyc = out[0] / (double)iters;  // every warp does `iters` ops of 32 lanes x 32 columns x 4 B = 4 KiB concurrently
        printf("%-5s %-6d %-9d %14.1f %12.1f\n", mode ? "st" : "ld", warps, wait_each, cyc, warps * 4096.0 / cyc);
      }
  return 0;
}
