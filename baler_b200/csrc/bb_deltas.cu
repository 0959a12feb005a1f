// Error-bounded deltas (reference helper.save_error_bounded_requirement, helper.py:442-470): which elements of the
// reconstruction miss the relative-error bound, and the float16 correction stored for them.  One streaming pass over
// the raw rows and their decoded (normalised) reconstruction: 8 B read per element, hits compacted through one global
// counter (warp-aggregated by the compiler); the host sorts the hits into the reference's row-major order.
#include "bb_common.cuh"

namespace {

__global__ void __launch_bounds__(256)
eb_delta_scan_kernel(const float* __restrict__ x, const float* __restrict__ y, const int64_t n_rows, const int c,
                     const float* __restrict__ mn, const float* __restrict__ rg, const double bound, const int64_t row0,
                     const unsigned long long capacity, unsigned long long* __restrict__ count, long long* __restrict__ rows_out,
                     int* __restrict__ cols_out, __half* __restrict__ deltas_out) {
  const int64_t total = n_rows * c, G = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += G) {
    const int64_t r = e / c;
    const int col = (int)(e - r * c);
    float xn = x[e];
    if (mn != nullptr) xn = __fdiv_rn(__fsub_rn(xn, __ldg(mn + col)), __ldg(rg + col));  // numpy float32 (x - min) / range
    const float yv = y[e];
    // percent error in float64, as upstream computes it on float64 tensors; data == 0 gives +-inf (set to 0 upstream,
    // i.e. never a hit) or NaN (never exceeds)
    const double err = ((double)yv - (double)xn) / (double)xn * 100.0;
    if (fabs(err) > bound && fabs(err) <= 1.79e308) {
      const unsigned long long slot = atomicAdd(count, 1ull);
      if (slot < capacity) {
        rows_out[slot] = row0 + r;
        cols_out[slot] = col;
        // np.subtract(decoded, data, dtype=float16): both operands rounded to float16, subtracted in float16
        deltas_out[slot] = __hsub(__float2half_rn(yv), __float2half_rn(xn));
      }
    }
  }
}

}  // namespace

extern "C" int bb_error_bounded_deltas_f32(bb_ctx* ctx, const float* x_dev, const float* y_dev, int64_t n_rows, int n_cols,
                                           const float* min_dev, const float* range_dev, double bound_percent, int64_t row0,
                                           int64_t capacity, unsigned long long* count_dev, long long* rows_out_dev,
                                           int* cols_out_dev, void* deltas_f16_out_dev, bb_stream_t stream) {
  if (!ctx || n_rows < 0 || n_cols < 1 || capacity < 0 || !count_dev || ((!x_dev || !y_dev) && n_rows)) return BB_ERR_INVALID;
  if ((min_dev == nullptr) != (range_dev == nullptr)) return BB_ERR_INVALID;
  if (capacity && (!rows_out_dev || !cols_out_dev || !deltas_f16_out_dev)) return BB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  BB_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), s));
  if (n_rows == 0) return BB_OK;
  eb_delta_scan_kernel<<<ctx->sm_count * 8, 256, 0, s>>>(x_dev, y_dev, n_rows, n_cols, min_dev, range_dev, bound_percent, row0,
                                                        (unsigned long long)capacity, count_dev, rows_out_dev, cols_out_dev,
                                                        reinterpret_cast<__half*>(deltas_f16_out_dev));
  return (int)cudaGetLastError();
}
