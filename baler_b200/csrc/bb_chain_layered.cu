// Layer-by-layer dense chain for shapes whose weights do not fit the fused kernels' shared memory
// (Conv_AE as its dense-equivalent chain with the 128 -> 2000 -> z Linears, models.py:316-407; CFD_dense_AE on
// 2500-feature snapshots, models.py:186-226).  Each layer is one fp32 GEMM launch
//     Y[n x N] = act(X[n x K] . W[N x K]^T + b)        (W row-major (out, in), exactly as nn.Linear stores it)
// with activations ping-ponging through global scratch in row chunks.  Both operands are K-major, tiles of
// 128 x 64 x 16 are staged through shared memory (transposed on the way in so the inner product loop reads
// 128-bit vectors), 8 x 4 outputs per thread, register prefetch of the next K slab.  All leading dimensions are
// padded to multiples of 4 floats (zero weights / zero bias in the padding) so every global access is a 16-byte
// vector.  Normalisation, dtype conversion and un-normalisation run as thin elementwise kernels around the GEMMs.
#include "bb_common.cuh"

namespace {

constexpr int BM = 128, BN = 64, BK = 16, LT = 256;

__device__ __forceinline__ float act_l(float v, int act) {
  if (act == BB_ACT_LEAKY) return v > 0.f ? v : BB_LEAKY * v;
  if (act == BB_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__global__ void __launch_bounds__(LT)
dense_layer_kernel(const float* __restrict__ X, const int ldx, const float* __restrict__ W, const int ldw,
                   const float* __restrict__ bias, float* __restrict__ Y, const int ldy, const int n, const int K,
                   const int Np, const int act) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  // loader mapping: one float4 (4 consecutive k) per thread and slab; A needs two rows per thread
  const int lr = tid >> 2, lk = (tid & 3) << 2;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  auto load = [&](int k0, float4& a0, float4& a1, float4& b0) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool kv = k0 + lk < K;  // K is padded to a multiple of 4, so a float4 is all-in or all-out
    a0 = (kv && m0 + lr < n) ? __ldg(reinterpret_cast<const float4*>(X + (size_t)(m0 + lr) * ldx + k0 + lk)) : z;
    a1 = (kv && m0 + lr + 64 < n) ? __ldg(reinterpret_cast<const float4*>(X + (size_t)(m0 + lr + 64) * ldx + k0 + lk)) : z;
    b0 = (kv && n0 + lr < Np) ? __ldg(reinterpret_cast<const float4*>(W + (size_t)(n0 + lr) * ldw + k0 + lk)) : z;
  };
  auto stash = [&](int buf, const float4& a0, const float4& a1, const float4& b0) {
    As[buf][lk + 0][lr] = a0.x; As[buf][lk + 1][lr] = a0.y; As[buf][lk + 2][lr] = a0.z; As[buf][lk + 3][lr] = a0.w;
    As[buf][lk + 0][lr + 64] = a1.x; As[buf][lk + 1][lr + 64] = a1.y; As[buf][lk + 2][lr + 64] = a1.z; As[buf][lk + 3][lr + 64] = a1.w;
    Bs[buf][lk + 0][lr] = b0.x; Bs[buf][lk + 1][lr] = b0.y; Bs[buf][lk + 2][lr] = b0.z; Bs[buf][lk + 3][lr] = b0.w;
  };
  float4 a0, a1, b0;
  load(0, a0, a1, b0);
  stash(0, a0, a1, b0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    const bool more = k0 + BK < K;
    if (more) load(k0 + BK, a0, a1, b0);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 x0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
      const float wv[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], wv[j], acc[i][j]);
    }
    if (more) {
      stash(buf ^ 1, a0, a1, b0);
      __syncthreads();
      buf ^= 1;
    }
  }
  const int c = n0 + tx * 4;
  if (c < Np) {
    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = m0 + ty * 8 + i;
      if (r < n) {
        float4 o;
        o.x = act_l(acc[i][0] + b.x, act); o.y = act_l(acc[i][1] + b.y, act);
        o.z = act_l(acc[i][2] + b.z, act); o.w = act_l(acc[i][3] + b.w, act);
        *reinterpret_cast<float4*>(Y + (size_t)r * ldy + c) = o;
      }
    }
  }
}

// in (f32 | f16, compact rows of `dim`) -> scratch rows of pitch `ld` (zero padded), optionally (x - min) / range
__global__ void __launch_bounds__(256) stage_in_kernel(const void* __restrict__ in, const int in_dtype, const int64_t n,
                                                       const int dim, const int ld, const float* __restrict__ mn,
                                                       const float* __restrict__ rg, float* __restrict__ out) {
  const int64_t total = n * ld, G = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += G) {
    const int64_t r = e / ld;
    const int c = (int)(e - r * ld);
    float v = 0.f;
    if (c < dim) {
      v = in_dtype == BB_F16 ? __half2float(reinterpret_cast<const __half*>(in)[r * dim + c]) : reinterpret_cast<const float*>(in)[r * dim + c];
      if (mn != nullptr) v = __fdiv_rn(__fsub_rn(v, __ldg(mn + c)), __ldg(rg + c));
    }
    out[e] = v;
  }
}

// scratch rows of pitch `ld` -> compact rows of `dim` (f32 | f16), optionally y * range + min
__global__ void __launch_bounds__(256) stage_out_kernel(const float* __restrict__ in, const int64_t n, const int dim,
                                                        const int ld, const float* __restrict__ mn,
                                                        const float* __restrict__ rg, void* __restrict__ out,
                                                        const int out_dtype) {
  const int64_t total = n * dim, G = (int64_t)gridDim.x * blockDim.x;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += G) {
    const int64_t r = e / dim;
    const int c = (int)(e - r * dim);
    float v = in[r * ld + c];
    if (mn != nullptr) v = fmaf(v, __ldg(rg + c), __ldg(mn + c));
    if (out_dtype == BB_F16) reinterpret_cast<__half*>(out)[e] = __float2half_rn(v);
    else reinterpret_cast<float*>(out)[e] = v;
  }
}

inline int pad4(int x) { return (x + 3) & ~3; }
constexpr int64_t CHUNK_ROWS = 1 << 15;

}  // namespace

// uploads W (pitch padded to 4, extra zero rows up to a multiple of 4) and bias per layer
int bb_chain_layered_prepare(bb_ctx*, Chain* c) {
  const ChainDesc& d = c->desc;
  size_t total = 0;
  int max_ld = pad4(d.in_dim);
  for (int l = 0; l < d.n_layers; ++l) {
    total += (size_t)pad4(d.layer[l].N) * pad4(d.layer[l].K) + pad4(d.layer[l].N);
    max_ld = pad4(d.layer[l].N) > max_ld ? pad4(d.layer[l].N) : max_ld;
  }
  std::vector<float> blob(total, 0.f);
  size_t off = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const int K = d.layer[l].K, N = d.layer[l].N, Kp = pad4(K), Np = pad4(N);
    c->lay_w_off[l] = off;
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) blob[off + (size_t)n * Kp + k] = (float)c->w_host[l][(size_t)n * K + k];
    off += (size_t)Np * Kp;
    c->lay_b_off[l] = off;
    for (int n = 0; n < N; ++n) blob[off + n] = (float)c->b_host[l][n];
    off += Np;
  }
  if (c->lay_blob_dev) cudaFree(c->lay_blob_dev);
  BB_CUDA(cudaMalloc(&c->lay_blob_dev, blob.size() * sizeof(float)));
  BB_CUDA(cudaMemcpy(c->lay_blob_dev, blob.data(), blob.size() * sizeof(float), cudaMemcpyHostToDevice));
  c->lay_max_ld = max_ld;
  c->lay_ok = true;
  return BB_OK;
}

int bb_chain_layered_launch(bb_ctx* ctx, const Chain* c, const void* in, int in_dtype, int64_t n_rows,
                            const float* pre_min, const float* pre_range, const float* post_min,
                            const float* post_range, void* out, int out_dtype, cudaStream_t stream) {
  if (!c->lay_ok) return BB_ERR_UNSUPPORTED;
  if (n_rows == 0) return BB_OK;
  const ChainDesc& d = c->desc;
  const int64_t chunk = n_rows < CHUNK_ROWS ? n_rows : CHUNK_ROWS;
  const size_t need = 2 * (size_t)chunk * c->lay_max_ld * sizeof(float);
  if (ctx->lay_scratch_bytes < need) {
    // grown only (never shrunk); stream-ordered work that still uses the old buffer has been enqueued before the free
    if (ctx->lay_scratch) BB_CUDA(cudaFree(ctx->lay_scratch));
    ctx->lay_scratch = nullptr;
    BB_CUDA(cudaMalloc(&ctx->lay_scratch, need));
    ctx->lay_scratch_bytes = need;
  }
  float* buf[2] = {ctx->lay_scratch, ctx->lay_scratch + (size_t)chunk * c->lay_max_ld};
  const size_t in_esz = in_dtype == BB_F16 ? 2 : 4, out_esz = out_dtype == BB_F16 ? 2 : 4;
  const int ew_grid = ctx->sm_count * 8;
  for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
    const int64_t rows = n_rows - r0 < chunk ? n_rows - r0 : chunk;
    int ld = pad4(d.in_dim);
    stage_in_kernel<<<ew_grid, 256, 0, stream>>>((const char*)in + (size_t)r0 * d.in_dim * in_esz, in_dtype, rows, d.in_dim, ld,
                                                 pre_min, pre_range, buf[0]);
    int cur = 0;
    for (int l = 0; l < d.n_layers; ++l) {
      const int K = d.layer[l].K, N = d.layer[l].N, Kp = pad4(K), Np = pad4(N);
      const dim3 grid((Np + BN - 1) / BN, (unsigned)((rows + BM - 1) / BM));
      dense_layer_kernel<<<grid, LT, 0, stream>>>(buf[cur], ld, c->lay_blob_dev + c->lay_w_off[l], Kp,
                                                 c->lay_blob_dev + c->lay_b_off[l], buf[cur ^ 1], Np, (int)rows, Kp, Np,
                                                 d.layer[l].act);
      cur ^= 1;
      ld = Np;
    }
    stage_out_kernel<<<ew_grid, 256, 0, stream>>>(buf[cur], rows, d.out_dim, ld, post_min, post_range,
                                                  (char*)out + (size_t)r0 * d.out_dim * out_esz, out_dtype);
  }
  return (int)cudaGetLastError();
}
