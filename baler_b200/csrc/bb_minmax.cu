// Per-column min / max of a row-major n x c float32 table in one streaming pass.
//
// Replaces data_processing.find_minmax (reference baler/modules/data_processing.py:113-130), which
// walks the table column by column in Python.  HBM-bound: 4*c bytes read per row, nothing written.
//
// Every thread reads 128-bit vectors at a grid stride G chosen so that 4*G is a multiple of c:
// a thread then always sees the same 4 columns and keeps their running min/max in registers.
// Block results are merged with shared-memory atomics on an order-preserving integer encoding of
// the floats, blocks are merged with global atomics on the same encoding, and the last block to
// finish decodes the result.  NaNs propagate (numpy's min/max semantics): a per-column flag.
#include "bb_common.cuh"

namespace {

__device__ __forceinline__ unsigned enc(float f) {
  const unsigned b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// acc layout (unsigned, global): [min c][max c][nan c][counter]
template <int VEC>
__global__ void __launch_bounds__(1024)
colminmax_kernel(const float* __restrict__ x, const int64_t n_elems, const int c,
                 unsigned* __restrict__ acc, float* __restrict__ min_out, float* __restrict__ max_out) {
  extern __shared__ unsigned sm[];  // [min c][max c][nan c]
  for (int i = threadIdx.x; i < 3 * c; i += blockDim.x) sm[i] = (i < c) ? 0xffffffffu : 0u;
  __syncthreads();

  const int64_t G = (int64_t)gridDim.x * blockDim.x;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float mn[VEC], mx[VEC];
  bool nan = false;
#pragma unroll
  for (int q = 0; q < VEC; ++q) { mn[q] = INFINITY; mx[q] = -INFINITY; }
  const int64_t n_vec = n_elems / VEC;
  if constexpr (VEC == 4) {
    const float4* xv = reinterpret_cast<const float4*>(x);
    int64_t v = g;
    for (; v + 3 * G < n_vec; v += 4 * G) {  // 4 independent 16-byte loads in flight per thread
      const float4 a = ld_stream(xv + v), b = ld_stream(xv + v + G), cc = ld_stream(xv + v + 2 * G),
                   dd = ld_stream(xv + v + 3 * G);
      const float e[4][4] = {{a.x, a.y, a.z, a.w}, {b.x, b.y, b.z, b.w}, {cc.x, cc.y, cc.z, cc.w}, {dd.x, dd.y, dd.z, dd.w}};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          mn[q] = fminf(mn[q], e[u][q]); mx[q] = fmaxf(mx[q], e[u][q]); nan |= (e[u][q] != e[u][q]);
        }
    }
    for (; v < n_vec; v += G) {
      const float4 a = ld_stream(xv + v);
      const float e[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { mn[q] = fminf(mn[q], e[q]); mx[q] = fmaxf(mx[q], e[q]); nan |= (e[q] != e[q]); }
    }
  } else {
    for (int64_t v = g; v < n_vec; v += G) {
      const float e = __ldg(x + v);
      mn[0] = fminf(mn[0], e); mx[0] = fmaxf(mx[0], e); nan |= (e != e);
    }
  }
  const int col0 = (int)((g * VEC) % c);
#pragma unroll
  for (int q = 0; q < VEC; ++q) {
    const int col = (col0 + q) % c;
    if (mn[q] <= mx[q]) { atomicMin(&sm[col], enc(mn[q])); atomicMax(&sm[c + col], enc(mx[q])); }
  }
  if (nan) {  // which of the thread's columns held the NaN is not tracked: flag all (NaN rows are degenerate)
#pragma unroll
    for (int q = 0; q < VEC; ++q) sm[2 * c + (col0 + q) % c] = 1u;
  }
  if (VEC == 4 && g == 0) {  // scalar tail (n_elems % 4 elements)
    for (int64_t e = n_vec * 4; e < n_elems; ++e) {
      const float f = x[e];
      const int col = (int)(e % c);
      if (f != f) sm[2 * c + col] = 1u;
      else { atomicMin(&sm[col], enc(f)); atomicMax(&sm[c + col], enc(f)); }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    atomicMin(&acc[i], sm[i]);
    atomicMax(&acc[c + i], sm[c + i]);
    if (sm[2 * c + i]) atomicOr(&acc[2 * c + i], 1u);
  }
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&acc[3 * c], 1u) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      const bool isnan = __ldcg(&acc[2 * c + i]) != 0u;
      min_out[i] = isnan ? __int_as_float(0x7fc00000) : dec(__ldcg(&acc[i]));
      max_out[i] = isnan ? __int_as_float(0x7fc00000) : dec(__ldcg(&acc[c + i]));
    }
    if (threadIdx.x == 0) acc[3 * c] = 0u;  // ready for an accumulating follow-up call
  }
}

// Wide tables (c > 1024: flattened 2-D snapshots, e.g. 50 x 50 = 2500 per-pixel columns of CFD_dense_AE): one thread per
// column, blockIdx.y strides over rows; consecutive threads read consecutive columns (coalesced), same accumulator layout.
__global__ void __launch_bounds__(256)
colminmax_wide_kernel(const float* __restrict__ x, const int64_t n_rows, const int c, unsigned* __restrict__ acc,
                      float* __restrict__ min_out, float* __restrict__ max_out) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < c) {
    float mn = INFINITY, mx = -INFINITY;
    bool nan = false;
    for (int64_t r = blockIdx.y; r < n_rows; r += gridDim.y) {
      const float e = __ldg(x + r * c + col);
      mn = fminf(mn, e); mx = fmaxf(mx, e); nan |= (e != e);
    }
    if (mn <= mx) { atomicMin(&acc[col], enc(mn)); atomicMax(&acc[c + col], enc(mx)); }
    if (nan) atomicOr(&acc[2 * c + col], 1u);
  }
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&acc[3 * c], 1u) == gridDim.x * gridDim.y - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      const bool isnan = __ldcg(&acc[2 * c + i]) != 0u;
      min_out[i] = isnan ? __int_as_float(0x7fc00000) : dec(__ldcg(&acc[i]));
      max_out[i] = isnan ? __int_as_float(0x7fc00000) : dec(__ldcg(&acc[c + i]));
    }
    if (threadIdx.x == 0) acc[3 * c] = 0u;
  }
}

int gcd(int a, int b) { return b ? gcd(b, a % b) : a; }

}  // namespace

// reset != 0: start a new reduction; reset == 0: fold this table into the running min/max (chunked input)
int bb_colminmax_launch(bb_ctx* ctx, const float* x, int64_t n_rows, int n_cols, float* min_dev,
                        float* max_dev, int reset, cudaStream_t stream) {
  if (n_cols <= 0 || n_rows < 0) return BB_ERR_INVALID;
  const size_t need = (size_t)(3 * n_cols + 1) * sizeof(unsigned);
  if (ctx->minmax_scratch_bytes < need) {
    if (ctx->minmax_scratch) cudaFree(ctx->minmax_scratch);
    BB_CUDA(cudaMalloc(&ctx->minmax_scratch, need));
    ctx->minmax_scratch_bytes = need;
    reset = 1;
  }
  unsigned* acc = reinterpret_cast<unsigned*>(ctx->minmax_scratch);
  if (reset) {
    BB_CUDA(cudaMemsetAsync(acc, 0xff, (size_t)n_cols * sizeof(unsigned), stream));
    BB_CUDA(cudaMemsetAsync(acc + n_cols, 0, (size_t)(2 * n_cols + 1) * sizeof(unsigned), stream));
  }
  if (n_rows == 0) return BB_OK;
  if (n_cols > 1024) {
    const int gx = (n_cols + 255) / 256;
    int64_t gy = (int64_t)ctx->sm_count * 8 / gx;
    gy = gy < 1 ? 1 : (gy > n_rows ? n_rows : gy);
    colminmax_wide_kernel<<<dim3(gx, (unsigned)gy), 256, 0, stream>>>(x, n_rows, n_cols, acc, min_dev, max_dev);
    return (int)cudaGetLastError();
  }
  const int64_t n_elems = n_rows * n_cols;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15u) == 0;
  const int lanes = vec ? n_cols / gcd(n_cols, 4) : n_cols;  // thread period that keeps columns fixed
  int threads = lanes * (384 / lanes > 0 ? 384 / lanes : 1);
  if (threads > 1024) return BB_ERR_UNSUPPORTED;
  int64_t want = (n_elems / (vec ? 4 : 1) + threads - 1) / threads;
  int grid = (int)(want < (int64_t)ctx->sm_count * 4 ? (want > 0 ? want : 1) : ctx->sm_count * 4);
  const size_t smem = (size_t)3 * n_cols * sizeof(unsigned);
  if (vec) colminmax_kernel<4><<<grid, threads, smem, stream>>>(x, n_elems, n_cols, acc, min_dev, max_dev);
  else colminmax_kernel<1><<<grid, threads, smem, stream>>>(x, n_elems, n_cols, acc, min_dev, max_dev);
  return (int)cudaGetLastError();
}

namespace {
// elementwise (x - min) / range or y * range + min; 8 bytes of HBM traffic per element
template <bool INVERSE>
__global__ void __launch_bounds__(256) colscale_kernel(const float* __restrict__ x, const int64_t n_elems, const int c,
                                                       const float* __restrict__ mn, const float* __restrict__ rg,
                                                       float* __restrict__ out) {
  const int64_t G = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_elems; e += G) {
    const int col = (int)(e % c);
    const float v = x[e];
    out[e] = INVERSE ? fmaf(v, __ldg(rg + col), __ldg(mn + col)) : __fdiv_rn(__fsub_rn(v, __ldg(mn + col)), __ldg(rg + col));
  }
}
}  // namespace

extern "C" {
int bb_normalize_f32(bb_ctx* ctx, const float* x_dev, int64_t n_rows, int n_cols, const float* min_dev,
                     const float* range_dev, float* out_dev, bb_stream_t stream) {
  if (!ctx || n_cols <= 0 || n_rows < 0 || !min_dev || !range_dev || ((!x_dev || !out_dev) && n_rows)) return BB_ERR_INVALID;
  if (n_rows == 0) return BB_OK;
  const int64_t n = n_rows * n_cols;
  const int grid = (int)((n + 255) / 256 < (int64_t)ctx->sm_count * 8 ? (n + 255) / 256 : ctx->sm_count * 8);
  colscale_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, n, n_cols, min_dev, range_dev, out_dev);
  return (int)cudaGetLastError();
}
int bb_renormalize_f32(bb_ctx* ctx, const float* y_dev, int64_t n_rows, int n_cols, const float* min_dev,
                       const float* range_dev, float* out_dev, bb_stream_t stream) {
  if (!ctx || n_cols <= 0 || n_rows < 0 || !min_dev || !range_dev || ((!y_dev || !out_dev) && n_rows)) return BB_ERR_INVALID;
  if (n_rows == 0) return BB_OK;
  const int64_t n = n_rows * n_cols;
  const int grid = (int)((n + 255) / 256 < (int64_t)ctx->sm_count * 8 ? (n + 255) / 256 : ctx->sm_count * 8);
  colscale_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(y_dev, n, n_cols, min_dev, range_dev, out_dev);
  return (int)cudaGetLastError();
}
}
