// Microbenchmark: issue cost (cycles per warp instruction per SM sub-partition) of the instructions the tcgen05
// chain epilogue is made of, with 1 / 2 / 4 warps per sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/alu_bench tools/alu_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void bench(int iters, long long* out, uint32_t* sink) {
  uint32_t r[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) r[j] = threadIdx.x * 2654435761u + j * 40503u + 0x3f800000u;
  const uint64_t c = 0x3f8000013f800001ull;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      if (OP == 0) {  // F2FP.F16.F32.PACK_AB
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[j]) : "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r[j + 1]) : "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j])));
      } else if (OP == 1) {  // FMNMX
        asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&r[j]) : "f"(__uint_as_float(r[j + 1])));
        asm volatile("max.f32 %0, %0, %1;" : "+f"(*(float*)&r[j + 1]) : "f"(__uint_as_float(r[j])));
      } else if (OP == 2) {  // LOP3
        asm volatile("and.b32 %0, %0, %1;" : "+r"(r[j]) : "r"(r[j + 1]));
        asm volatile("and.b32 %0, %0, %1;" : "+r"(r[j + 1]) : "r"(r[j]));
      } else if (OP == 3) {  // FMUL2
        uint64_t v = (uint64_t)r[j] | ((uint64_t)r[j + 1] << 32);
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(c));
        asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(c));
        r[j] = (uint32_t)v; r[j + 1] = (uint32_t)(v >> 32);
      } else if (OP == 4) {  // FADD2
        uint64_t v = (uint64_t)r[j] | ((uint64_t)r[j + 1] << 32);
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(c));
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(c));
        r[j] = (uint32_t)v; r[j + 1] = (uint32_t)(v >> 32);
      } else if (OP == 5) {  // FMUL (register operands)
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(*(float*)&r[j]) : "f"(__uint_as_float(r[j + 1])));
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(*(float*)&r[j + 1]) : "f"(__uint_as_float(r[j])));
      } else if (OP == 6) {  // FFMA with |x| modifier
        asm volatile("{.reg .f32 t; abs.f32 t, %1; fma.rn.f32 %0, t, %2, %0;}" : "+f"(*(float*)&r[j]) : "f"(__uint_as_float(r[j + 1])), "f"(0.495f));
        asm volatile("{.reg .f32 t; abs.f32 t, %1; fma.rn.f32 %0, t, %2, %0;}" : "+f"(*(float*)&r[j + 1]) : "f"(__uint_as_float(r[j])), "f"(0.495f));
      } else if (OP == 7) {  // cvt.f32.f16 (HADD2.F32)
        asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(*(float*)&r[j]) : "r"(r[j + 1]));
        asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, hi;}" : "=f"(*(float*)&r[j + 1]) : "r"(r[j]));
      } else if (OP == 8) {  // PRMT
        asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(r[j]) : "r"(r[j + 1]));
        asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(r[j + 1]) : "r"(r[j]));
      } else if (OP == 9) {  // cvt.rz.f16x2.f32
        asm volatile("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(r[j]) : "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])));
        asm volatile("cvt.rz.f16x2.f32 %0, %1, %2;" : "=r"(r[j + 1]) : "f"(__uint_as_float(r[j + 1])), "f"(__uint_as_float(r[j])));
      } else if (OP == 10) {  // FFMA2
        uint64_t v = (uint64_t)r[j] | ((uint64_t)r[j + 1] << 32);
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(c));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(c));
        r[j] = (uint32_t)v; r[j + 1] = (uint32_t)(v >> 32);
      } else if (OP == 11) {  // HFMA2 (fp16x2 fma)
        asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(r[j]) : "r"(r[j + 1]));
        asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(r[j + 1]) : "r"(r[j]));
      }
    }
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) s ^= r[j];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, long long* out, uint32_t* sink) {
  const int iters = 512;
  printf("%-22s", name);
  for (int wps : {1, 2, 4}) {  // warps per sub-partition
    for (int rep = 0; rep < 2; ++rep) { bench<OP><<<1, 128 * wps>>>(iters, out, sink); cudaDeviceSynchronize(); }
    printf("  wps=%d: %6.2f cyc/instr/SMSP", wps, out[0] / (double)(iters * 16 * wps));
  }
  printf("\n");
}

int main() {
  long long* out; uint32_t* sink;
  cudaMallocManaged(&out, 16); cudaMalloc(&sink, 4096);
  run<0>("F2FP.PACK_AB rn", out, sink);
  run<9>("F2FP.PACK_AB rz", out, sink);
  run<1>("FMNMX", out, sink);
  run<2>("LOP3", out, sink);
  run<8>("PRMT", out, sink);
  run<3>("FMUL2", out, sink);
  run<4>("FADD2", out, sink);
  run<10>("FFMA2", out, sink);
  run<5>("FMUL", out, sink);
  run<6>("FFMA |x|", out, sink);
  run<7>("cvt.f32.f16", out, sink);
  run<11>("HFMA2", out, sink);
  return 0;
}
