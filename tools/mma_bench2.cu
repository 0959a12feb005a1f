// Microbenchmark 2: what bounds a stream of small tcgen05.mma (kind::f16, K = 16) on one SM?
//   * issue by a CONVERGED warp (elect.sync inside), so the issue path is the tight one the product kernel uses
//   * sweeps N, M (64 / 128), A source (TMEM / SMEM no-swizzle / SMEM 128B-swizzle), B swizzle, collector hints,
//     .ws form, number of MMAs (separates fixed latency from the per-MMA cost)
//   * optional interference: 4 other warps run tcgen05.ld + wait::ld in a loop while the MMA stream runs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_bench2 tools/mma_bench2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
  int n_mma, M, N, n_acc;
  int a_src;      // 0 TMEM, 1 SMEM no swizzle, 2 SMEM 128B swizzle
  int b_swz;      // 0 none, 1 128B
  int mode;       // 0 plain, 1 collector a: fill,use,lastuse per group of 3 (same A), 2 .ws plain, 3 .ws with b0 fill/use/lastuse
  int ld_warps;   // warps 1..ld_warps loop on tcgen05.ld while the MMAs run
  int batch;      // > 0: commit + wait after every `batch` MMAs (bounded queue depth)
  int tm_op;      // what the interfering warps do: 0 ld.x32, 1 st.x16, 2 ld.x16 + st.x16 + both waits (the epilogue pattern), 3 same + ~70 ALU ops
  int tm_gap;     // extra clock-spin between interfering ops (cycles)
};

#define R16(v, o) "=r"(v[o+0]), "=r"(v[o+1]), "=r"(v[o+2]), "=r"(v[o+3]), "=r"(v[o+4]), "=r"(v[o+5]), "=r"(v[o+6]), "=r"(v[o+7]), \
                  "=r"(v[o+8]), "=r"(v[o+9]), "=r"(v[o+10]), "=r"(v[o+11]), "=r"(v[o+12]), "=r"(v[o+13]), "=r"(v[o+14]), "=r"(v[o+15])

__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (spins > 4000000u) __trap();
  }
}

template <int MODE, int A_SMEM>
__device__ __forceinline__ void issue3(uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d, uint32_t idesc, uint32_t acc, int j) {
  // j = position in a group of 3 (collector hints)
#define MMA_TS(SUFFIX) asm volatile("{\n\t.reg .pred e, p;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@e tcgen05.mma" SUFFIX " [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc) : "memory")
#define MMA_SS(SUFFIX) asm volatile("{\n\t.reg .pred e, p;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t@e tcgen05.mma" SUFFIX " [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc) : "memory")
  if constexpr (MODE == 0) {
    if constexpr (A_SMEM) MMA_SS(".cta_group::1.kind::f16"); else MMA_TS(".cta_group::1.kind::f16");
  } else if constexpr (MODE == 1) {
    if (j == 0) { if constexpr (A_SMEM) MMA_SS(".cta_group::1.kind::f16.collector::a::fill"); else MMA_TS(".cta_group::1.kind::f16.collector::a::fill"); }
    else if (j == 1) { if constexpr (A_SMEM) MMA_SS(".cta_group::1.kind::f16.collector::a::use"); else MMA_TS(".cta_group::1.kind::f16.collector::a::use"); }
    else { if constexpr (A_SMEM) MMA_SS(".cta_group::1.kind::f16.collector::a::lastuse"); else MMA_TS(".cta_group::1.kind::f16.collector::a::lastuse"); }
  } else if constexpr (MODE == 2) {
    if constexpr (A_SMEM) MMA_SS(".ws.cta_group::1.kind::f16"); else MMA_TS(".ws.cta_group::1.kind::f16");
  } else {
    if (j == 0) { if constexpr (A_SMEM) MMA_SS(".ws.cta_group::1.kind::f16.collector::b0::fill"); else MMA_TS(".ws.cta_group::1.kind::f16.collector::b0::fill"); }
    else if (j == 1) { if constexpr (A_SMEM) MMA_SS(".ws.cta_group::1.kind::f16.collector::b0::use"); else MMA_TS(".ws.cta_group::1.kind::f16.collector::b0::use"); }
    else { if constexpr (A_SMEM) MMA_SS(".ws.cta_group::1.kind::f16.collector::b0::lastuse"); else MMA_TS(".ws.cta_group::1.kind::f16.collector::b0::lastuse"); }
  }
}

template <int MODE, int A_SMEM>
__global__ void __launch_bounds__(544, 1) bench(const Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_s;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) {
    stop_s = 0;
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
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    const int N = c.N;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(c.M >> 4) << 24);
    const uint32_t sb = smem_u32(smem);
    uint64_t bdesc, adesc;
    if (c.b_swz) bdesc = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    else bdesc = (uint64_t)((sb >> 4) & 0x3FFF) | ((uint64_t)(((uint32_t)N * 16) >> 4) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    const uint32_t sa = sb + 49152;
    if (c.a_src == 2) adesc = (uint64_t)((sa >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    else adesc = (uint64_t)((sa >> 4) & 0x3FFF) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)8 << 32) | (1ull << 46);
    const int acc_stride = c.n_acc > 1 ? (448 / c.n_acc) & ~15 : 0;
    uint32_t dd[6];
    for (int j = 0; j < 6; ++j) dd[j] = tb + (uint32_t)((j % c.n_acc) * acc_stride);
    const uint32_t a_t = tb + 480;
    uint32_t par = 0;
    __syncwarp();
    const long long t0 = clock64();
    int since = 0;
    for (int i = 0; i < c.n_mma; i += 6) {
#pragma unroll
      for (int j = 0; j < 6; ++j) issue3<MODE, A_SMEM>(dd[j], a_t + 8 * (j & 1), adesc, bdesc, idesc, i + j >= c.n_acc ? 1u : 0u, j % 3);
      since += 6;
      if (c.batch > 0 && since >= c.batch) {
        since = 0;
        asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
        wait_bar(smem_u32(&bar), par);
        par ^= 1u;
      }
    }
    const long long t1 = clock64();
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    wait_bar(smem_u32(&bar), par);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    stop_s = 1;
  } else if (warp <= c.ld_warps) {
    // own lane quarter (warp % 4), columns 448..479 (untouched by the MMAs' accumulators when n_acc * N <= 448)
    const uint32_t taddr = tb + (((uint32_t)(warp & 3) * 32u) << 16) + 448u;
    uint32_t v[32];
    long long tot = 0, mx = 0;
    int n = 0;
    while (!stop_s && n < 100000) {
      const long long a = clock64();
      if (c.tm_op == 0) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                     "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : R16(v, 0), R16(v, 16) : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      } else {
        if (c.tm_op >= 2) {
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                       : R16(v, 0) : "r"(taddr) : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        }
        if (c.tm_op == 3) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]) * 1.01f, 0.5f)) & 0xffffe000u; v[j] ^= v[(j + 1) & 15] >> 3; v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.25f)); }
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      }
      const long long b = clock64();
      if (c.tm_gap) { while (clock64() - b < c.tm_gap) {} }
      tot += b - a; mx = b - a > mx ? b - a : mx; ++n;
    }
    uint32_t s = 0;
    for (int j = 0; j < 32; ++j) s += v[j];
    if ((threadIdx.x & 31) == 0) { out[2 + 3 * (warp - 1)] = tot; out[3 + 3 * (warp - 1)] = n + (s == 0x12345 ? 1 : 0); out[4 + 3 * (warp - 1)] = mx; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(512) : "memory");
}

static long long* out;
template <int MODE, int A_SMEM>
static int run1(const Cfg& c) {
  auto k = bench<MODE, A_SMEM>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
  for (int rep = 0; rep < 2; ++rep) {
    for (int i = 0; i < 64; ++i) out[i] = 0;
    k<<<1, 32 * (1 + (c.ld_warps > 4 ? c.ld_warps : 4)), 96 * 1024>>>(c, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  }
  return 0;
}
static int run(const Cfg& c) {
  int rc;
  const bool as = c.a_src != 0;
  switch (c.mode) {
    case 0: rc = as ? run1<0, 1>(c) : run1<0, 0>(c); break;
    case 1: rc = as ? run1<1, 1>(c) : run1<1, 0>(c); break;
    case 2: rc = as ? run1<2, 1>(c) : run1<2, 0>(c); break;
    default: rc = as ? run1<3, 1>(c) : run1<3, 0>(c); break;
  }
  if (rc) return rc;
  printf("M=%-3d N=%-3d nmma=%-4d nacc=%d A=%s Bswz=%d mode=%d batch=%-2d ldw=%d | issue/mma %6.1f total/mma %6.1f", c.M, c.N, c.n_mma, c.n_acc,
         c.a_src == 0 ? "tmem" : c.a_src == 1 ? "smem" : "sswz", c.b_swz, c.mode, c.batch, c.ld_warps, out[0] / (double)c.n_mma, out[1] / (double)c.n_mma);
  if (c.ld_warps) printf(" | tm_op %d gap %d: avg %6.1f max %lld (n=%lld)", c.tm_op, c.tm_gap, out[2] / (double)(out[3] ? out[3] : 1), out[4], out[3]);
  printf("\n");
  fflush(stdout);
  return 0;
}

int main(int argc, char** argv) {
  cudaMallocManaged(&out, 64 * sizeof(long long));
  // MMA stream (N = 112 and 64, one accumulator region) next to 0 / 4 / 8 / 16 warps doing TMEM traffic of several kinds
  for (int N : {112, 64})
    for (int op : {0, 1, 2, 3})
      for (int ldw : {0, 4, 8, 16})
        for (int gap : {0, 200}) {
          if (ldw == 0 && (op > 0 || gap > 0)) continue;
          if (run({768, 128, N, 2, 0, 0, 0, ldw, 0, op, gap})) return 1;
        }
  return 0;
}
