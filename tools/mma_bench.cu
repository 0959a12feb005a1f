// Microbenchmark: cost of back-to-back tcgen05.mma (M=128, K=16, kind::f16) issued by one thread, as a function
// of N, of the number of independent accumulators they rotate over, and of where A lives (TMEM / SMEM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_bench tools/mma_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) bench(int n_mma, int N, int n_acc, int a_in_smem, int dep_a, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
    const uint32_t lbo = (uint32_t)N * 16;
    const uint64_t bdesc = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    const uint64_t adesc = (uint64_t)(((smem_u32(smem) + 32768) >> 4) & 0x3FFF) | ((uint64_t)((2048 >> 4) & 0x3FFF) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    const int acc_stride = n_acc > 1 ? (448 / n_acc) & ~15 : 0;
    const long long t0 = clock64();
    // unrolled by 8, no integer division on the issue path: what one thread can sustain
    uint32_t dd[8], aa[8];
    for (int j = 0; j < 8; ++j) { dd[j] = tb + (uint32_t)((j % n_acc) * acc_stride); aa[j] = tb + 480 + (dep_a ? 0 : 8 * (j & 1)); }
    for (int i = 0; i < n_mma; i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (a_in_smem)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(dd[j]), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(i + j >= n_acc ? 1 : 0) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(dd[j]), "r"(aa[j]), "l"(bdesc), "r"(idesc), "r"(i + j >= n_acc ? 1 : 0) : "memory");
      }
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  const int n_mma = 96;
  printf("%-6s %-5s %-6s %-6s %12s %12s\n", "A", "N", "n_acc", "sameA", "issue/mma", "total/mma");
  for (int a_smem = 0; a_smem < 2; ++a_smem)
    for (int N : {16, 64, 112, 208, 256})
      for (int n_acc : {1, 2, 4}) {
        if (n_acc * N > 448) continue;
        for (int rep = 0; rep < 2; ++rep) {
          bench<<<1, 128, 96 * 1024>>>(n_mma, N, n_acc, a_smem, 1, out);
          if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        }
        printf("%-6s %-5d %-6d %-6d %12.1f %12.1f\n", a_smem ? "smem" : "tmem", N, n_acc, 1, out[0] / (double)n_mma, out[1] / (double)n_mma);
      }
  return 0;
}
