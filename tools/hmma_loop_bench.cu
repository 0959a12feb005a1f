// What limits the phase-1 MMA loop of bb_train_tc.cu?  One CTA per SM, 8 warps, one layer pass = KS k-steps x NT n-tiles,
// operands in shared memory.  Variants add the loop's ingredients one by one.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/hmma_loop_bench tools/hmma_loop_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)) : "memory");
}

constexpr int KS = 13, NT = 14, LDA = KS * 16 + 8, NW = 8, MAXJ = 4;
constexpr int STAGE = 32768;

__device__ __forceinline__ uint32_t rnd(uint32_t& st) { st = st * 1664525u + 1013904223u; return st; }
// data: 0 = constants, 1 = random normal-range fp16, 2 = random with many subnormals / tiny values
// the same pass while a 9th warp streams 32 KB bulk copies (TMA 1-D) from global memory into a shared-memory ring
__global__ void __launch_bounds__((NW + 1) * 32, 1) stream_loop_kernel(int passes, float* out, long long* cyc, const uint4* src, int do_stream) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint4* B = reinterpret_cast<uint4*>(smem);
  __half* AH = reinterpret_cast<__half*>(smem + KS * NT * 512);
  __half* AL = AH + 16 * LDA;
  unsigned char* ring = smem + KS * NT * 512 + 2 * 16 * LDA * 2 + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + 3 * STAGE);
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < KS * NT * 32; i += (NW + 1) * 32) B[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x14001400u, 0x14001400u);
  for (int i = tid; i < 2 * 16 * LDA / 2; i += (NW + 1) * 32) reinterpret_cast<uint32_t*>(AH)[i] = 0x2c002c00u;
  if (tid == 0) {
    done = 0;
    for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bars + i)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == NW) {
    if (lane == 0 && do_stream) {
      int c = 0;
      while (!done) {
        const uint32_t bar = smem_u32(bars + c % 3);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + (c % 3) * STAGE)),
                     "l"(src + (size_t)(c % 16) * (STAGE / 16)), "r"(STAGE), "r"(bar) : "memory");
        uint32_t ok = 0;
        while (!ok) {
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok) : "r"(bar), "r"((c / 3) & 1) : "memory");
          if (!ok) __nanosleep(100);
        }
        ++c;
      }
    }
    return;
  }
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
  float acc[MAXJ][2][4];
  for (int j = 0; j < MAXJ; ++j) for (int q = 0; q < 2; ++q) for (int i = 0; i < 4; ++i) acc[j][q][i] = 0.f;
  uint32_t ah[4], al[4];
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    if (warp < NT) {
      for (int ks = 0; ks < KS; ++ks) {
        ldsm_x4(ah, AH + a_row * LDA + ks * 16 + a_col); ldsm_x4(al, AL + a_row * LDA + ks * 16 + a_col);
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          const int nt = warp + NW * j;
          if (nt < NT) {
            const uint4 b = B[(ks * NT + nt) * 32 + lane];
            float tmp[4] = {0.f, 0.f, 0.f, 0.f};
            mma16816(tmp, ah, b.x, b.y);
            mma16816(acc[j][1], ah, b.z, b.w);
            mma16816(acc[j][1], al, b.x, b.y);
            for (int i = 0; i < 4; ++i) acc[j][0][i] += tmp[i];
          }
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
  }
  const long long t1 = clock64();
  if (tid == 0) done = 1;
  float s = 0.f;
  for (int j = 0; j < MAXJ; ++j) for (int q = 0; q < 2; ++q) for (int i = 0; i < 4; ++i) s += acc[j][q][i];
  out[blockIdx.x * NW * 32 + tid] = s;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int V>
__global__ void __launch_bounds__(NW * 32, 1) loop_kernel(int passes, float* out, long long* cyc, int data) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint4* B = reinterpret_cast<uint4*>(smem);                       // [KS][NT][32]
  __half* AH = reinterpret_cast<__half*>(smem + KS * NT * 512);    // [16][LDA]
  __half* AL = AH + 16 * LDA;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < KS * NT * 32; i += NW * 32) B[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x14001400u, 0x14001400u);
  for (int i = tid; i < 2 * 16 * LDA / 2; i += NW * 32) reinterpret_cast<uint32_t*>(AH)[i] = 0x2c002c00u;
  if (data) {
    uint32_t st = 12345u + tid * 977u;
    for (int i = tid; i < KS * NT * 32 * 4; i += NW * 32) {
      const uint32_t r = rnd(st);
      // fp16 pairs: sign | exponent (data 1: 10..16, data 2: 0..3) | mantissa
      const uint32_t e = data == 1 ? (10 + (r >> 28) % 6) : ((r >> 28) & 3);
      reinterpret_cast<uint32_t*>(B)[i] = ((r & 0x83ff83ffu) | (e << 10) | (e << 26));
    }
    for (int i = tid; i < 2 * 16 * LDA / 2; i += NW * 32) {
      const uint32_t r = rnd(st);
      const uint32_t e = data == 1 ? (10 + (r >> 28) % 6) : ((r >> 28) & 3);
      reinterpret_cast<uint32_t*>(AH)[i] = ((r & 0x83ff83ffu) | (e << 10) | (e << 26));
    }
  }
  __syncthreads();
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_col = (lane >> 4) * 8;
  float acc[MAXJ][2][4];
  for (int j = 0; j < MAXJ; ++j) for (int q = 0; q < 2; ++q) for (int i = 0; i < 4; ++i) acc[j][q][i] = 0.f;
  uint32_t ah[4] = {0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u}, al[4] = {0x2c002c00u, 0x2c002c00u, 0x2c002c00u, 0x2c002c00u};
  uint4 breg = make_uint4(0x3c003c00u, 0x3c003c00u, 0x14001400u + tid, 0x14001400u);
  __syncthreads();
  const long long t0 = clock64();
  for (int p = 0; p < passes; ++p) {
    if (warp < NT) {
      for (int ks = 0; ks < KS; ++ks) {
        if (V >= 2) { ldsm_x4(ah, AH + a_row * LDA + ks * 16 + a_col); ldsm_x4(al, AL + a_row * LDA + ks * 16 + a_col); }
#pragma unroll
        for (int j = 0; j < MAXJ; ++j) {
          const int nt = warp + NW * j;
          if (nt < NT) {
            uint4 b = breg;
            if (V >= 1) b = B[(ks * NT + nt) * 32 + lane];
            if (V >= 3) {
              float tmp[4] = {0.f, 0.f, 0.f, 0.f};
              mma16816(tmp, ah, b.x, b.y);
              mma16816(acc[j][1], ah, b.z, b.w);
              mma16816(acc[j][1], al, b.x, b.y);
              for (int i = 0; i < 4; ++i) acc[j][0][i] += tmp[i];
            } else {
              mma16816(acc[j][0], ah, b.x, b.y);
              mma16816(acc[j][1], ah, b.z, b.w);
              mma16816(acc[j][1], al, b.x, b.y);
            }
          }
        }
      }
    }
    if (V >= 4) __syncthreads();
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < MAXJ; ++j) for (int q = 0; q < 2; ++q) for (int i = 0; i < 4; ++i) s += acc[j][q][i];
  out[blockIdx.x * NW * 32 + tid] = s;
  if (tid == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8);
  const int smem = KS * NT * 512 + 2 * 16 * LDA * 2 + 256, passes = 200;
  const char* names[] = {"HMMA only (operands in registers)", "+ B fragments by LDS.128", "+ A fragments by ldmatrix", "+ zero-accumulator hi*hi and fp32 add", "+ __syncthreads per pass"};
#define RUN(V) for (int data = 0; data < 3; ++data) { cudaFuncSetAttribute(loop_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
    for (int w = 0; w < 2; ++w) loop_kernel<V><<<32, NW * 32, smem>>>(passes, out, cyc, data); \
    cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("  V%d data %d %-45s %8.0f cycles per pass (%d HMMA: %.2f cycles/HMMA/SM) %s\n", V, data, names[V], (double)h / passes, KS * NT * 3, (double)h / passes / (KS * NT * 3), cudaGetErrorString(e)); }
  printf("pass of KS=%d k-steps x NT=%d n-tiles, 8 warps, 32 CTAs\n", KS, NT);
  RUN(0) RUN(2) RUN(3)
  {
    uint4* src; cudaMalloc(&src, 16 * STAGE); cudaMemset(src, 0, 16 * STAGE);
    const int smem2 = smem + 3 * STAGE + 64;
    cudaFuncSetAttribute(stream_loop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    for (int st = 0; st < 2; ++st) {
      for (int w = 0; w < 2; ++w) stream_loop_kernel<<<32, (NW + 1) * 32, smem2>>>(passes, out, cyc, src, st);
      cudaError_t e = cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      printf("  V3 + 9th warp %s: %8.0f cycles per pass  %s\n", st ? "streaming 32 KB bulk copies into shared memory" : "idle", (double)h / passes, cudaGetErrorString(e));
    }
  }
  return 0;
}
