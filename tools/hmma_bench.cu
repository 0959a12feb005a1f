// Microbenchmarks behind the design of bb_train_tc.cu (B200):
//   1. mma.sync m16n8k16 (HMMA) latency and per-SM throughput
//   2. L2 -> shared memory streaming of one shared 584 KB weight image by G CTAs: cp.async.cg 16 B per thread versus
//      cp.async.bulk (one thread, mbarrier), CTAs reading the same bytes at the same time or rotated
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/hmma_bench tools/hmma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NACC>
__global__ void hmma_kernel(int iters, float* out, long long* cyc) {
  float acc[NACC][4];
  uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
  uint32_t b0 = 0x3c003c00u + threadIdx.x, b1 = 0x3c003c00u;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma16816(acc[j], a, b0, b1);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += acc[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int STAGE = 32768, RING = 4;
// mode 0: cp.async.cg 16 B per thread; mode 1: cp.async.bulk by thread 0; rot: start offset rotated per CTA
__global__ void __launch_bounds__(512, 1) stream_kernel(const uint4* src, int n_chunks, int reps, int mode, int rot, long long* cyc, float* sink) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RING * STAGE);
  const int tid = threadIdx.x;
  if (mode == 1 && tid == 0) {
    for (int i = 0; i < RING; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + i)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int total = n_chunks * reps;
  const int shift = rot ? (int)(((long long)blockIdx.x * n_chunks) / gridDim.x) : 0;
  auto issue = [&](int c) {
    if (c < total) {
      const uint4* s = src + (size_t)((c + shift) % n_chunks) * (STAGE / 16);
      unsigned char* d = smem + (c % RING) * STAGE;
      if (mode == 0) {
        for (int i = tid; i < STAGE / 16; i += 512)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(d + i * 16)), "l"(s + i) : "memory");
      } else if (tid == 0) {
        const uint32_t bar = smem_u32(bars + (c % RING));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(d)),
                     "l"(s), "r"(STAGE), "r"(bar) : "memory");
      }
    }
    if (mode == 0) asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue(0); issue(1); issue(2);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int c = 0; c < total; ++c) {
    if (mode == 0) {
      asm volatile("cp.async.wait_group 2;" ::: "memory");
    } else {
      const uint32_t bar = smem_u32(bars + (c % RING)), parity = (c / RING) & 1;
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
    __syncthreads();
    issue(c + 3);
    acc += reinterpret_cast<const float*>(smem + (c % RING) * STAGE)[tid];
  }
  const long long t1 = clock64();
  if (mode == 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
  sink[blockIdx.x * 512 + tid] = acc;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  printf("HMMA m16n8k16 f16->f32: cycles per MMA per warp (block 0)\n");
#define RUN(NACC, WARPS) { hmma_kernel<NACC><<<148, WARPS * 32>>>(iters, out, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("  %2d warps/SM x %d independent accumulators: %.2f cycles/MMA/warp -> %.2f MMA/cycle/SM (%.0f dense FLOP/clk/SM)\n", WARPS, NACC, \
           (double)h / (iters * NACC), (double)WARPS * NACC * iters / h, 4096.0 * WARPS * NACC * iters / h); }
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(1, 4) RUN(4, 4) RUN(4, 8) RUN(4, 16) RUN(8, 16) RUN(2, 16)
  uint4* src; cudaMalloc(&src, 18 * STAGE); cudaMemset(src, 0, 18 * STAGE);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RING * STAGE + 64);
  printf("L2 -> smem streaming of one 576 KB image (18 x 32 KB chunks), 512 threads, 4-stage ring\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int rot = 0; rot < 2; ++rot)
      for (int grid : {1, 32, 90, 148}) {
        for (int w = 0; w < 2; ++w) stream_kernel<<<grid, 512, RING * STAGE + 64>>>(src, 18, 20, mode, rot, cyc, out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %s %s grid %3d: %.0f cycles per 32 KB chunk = %.1f B/clk/SM  (%s)\n", mode ? "cp.async.bulk" : "cp.async.cg16 ",
               rot ? "rotated" : "lockstep", grid, (double)h / 360, 32768.0 * 360 / h, cudaGetErrorString(e));
      }
  return 0;
}
