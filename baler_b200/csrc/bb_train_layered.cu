// Layer-by-layer fp32 trainer for dense autoencoders whose weight matrices do not fit the fused training kernels'
// shared memory: `CFD_dense_AE` on 2500-feature snapshots (models.py:186-226; W1 and W8 are 200 x 2500), or any dense
// chain of up to BB_MAX_LAYERS Linears.  Same arithmetic contract as bb_train.cu (training.py:31-101: forward,
// sum-MSE / n_columns, backward, Adam with the reference's defaults, loss accumulated on the device), one GEMM launch
// per matrix product instead of one fused kernel:
//     forward   A_{l+1} = act(A_l W_l^T + b_l)                       gemm<NT> + bias + activation epilogue
//     loss      sum((A_L - x)^2) / C,   dA_L = 2 (A_L - x) / C       elementwise + block reduction
//     backward  dZ_l = dA_{l+1} * act'(A_{l+1});  dW_l = dZ_l^T A_l;  db_l = colsum(dZ_l);  dA_l = dZ_l W_l
//     Adam      torch.optim.Adam single-tensor update (training.py:266)
// Shapes here are small-batch x wide (60 x 2500): the GEMMs are latency-bound either way, so the kernel is a plain
// 64 x 64 x 16 shared-memory tile with arbitrary strides (one kernel serves NT, NN and TN), not a tuned one.
#include <cmath>
#include <vector>

#include "bb_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16, GT = 256;

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == BB_ACT_LEAKY) return v > 0.f ? v : BB_LEAKY * v;
  if (act == BB_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// C[m][n] = sum_k A(m, k) * B(k, n) with A(m, k) = A[m * sam + k * sak], B(k, n) = B[k * sbk + n * sbn];
// epilogue: + bias[n], activation; fixed summation order (k ascending), one thread per 4 x 4 outputs
__global__ void __launch_bounds__(GT)
gemm_strided_kernel(const float* __restrict__ A, const int64_t sam, const int64_t sak, const float* __restrict__ B,
                    const int64_t sbk, const int64_t sbn, float* __restrict__ C, const int64_t ldc, const int M, const int N,
                    const int K, const float* __restrict__ bias, const int act) {
  __shared__ float As[TK][TM + 1];
  __shared__ float Bs[TK][TN + 1];
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int e = tid; e < TK * TM; e += GT) {
      const int kk = e / TM, mm = e - kk * TM;
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? __ldg(A + (int64_t)m * sam + (int64_t)k * sak) : 0.f;
    }
    for (int e = tid; e < TK * TN; e += GT) {
      const int kk = e / TN, nn = e - kk * TN;
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? __ldg(B + (int64_t)k * sbk + (int64_t)n * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[(int64_t)m * ldc + n] = act_fwd(acc[i][j] + (bias ? __ldg(bias + n) : 0.f), act);
    }
  }
}

// dA_L = 2 (recon - x) / C and the batch loss sum((recon - x)^2) / C (one partial per block, fixed order)
__global__ void __launch_bounds__(256) loss_seed_kernel(const float* __restrict__ recon, const float* __restrict__ x,
                                                        const int64_t n, const float inv_c, float* __restrict__ dA,
                                                        float* __restrict__ loss_part) {
  __shared__ float red[256];
  float s = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const float d = recon[e] - x[e];
    s += d * d * inv_c;
    if (dA) dA[e] = 2.f * d * inv_c;
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int w = 128; w > 0; w >>= 1) {
    if (threadIdx.x < w) red[threadIdx.x] += red[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_part[blockIdx.x] = red[0];
}

// dZ = dA * act'(A) in place (A is the post-activation value: its sign is the pre-activation's)
__global__ void __launch_bounds__(256) act_bwd_kernel(float* __restrict__ dA, const float* __restrict__ A_out, const int64_t n,
                                                      const int act) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    if (act == BB_ACT_LEAKY) dA[e] *= A_out[e] > 0.f ? 1.f : BB_LEAKY;
    else if (act == BB_ACT_RELU) dA[e] *= A_out[e] > 0.f ? 1.f : 0.f;
  }
}

// db[n] = sum_m dZ[m][n] (rows ascending: reproducible)
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ dZ, const int M, const int N, float* __restrict__ db) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += dZ[(int64_t)m * N + n];
  db[n] = s;
}

// mode 0: Adam update + fold the loss partials into grads[n_params] and *loss_accum; 1: only the loss fold (phase 1 of a
// data-parallel step / validation); 2: Adam update with the (all-reduced) grads, loss from grads[n_params]
__global__ void __launch_bounds__(256) ladam_kernel(const int n_params, const int mode, float* __restrict__ grads,
                                                    float* __restrict__ params, float* __restrict__ m, float* __restrict__ v,
                                                    const float lr_bc1, const float inv_sqrt_bc2, const float beta1,
                                                    const float beta2, const float eps, const float* __restrict__ loss_part,
                                                    const int n_loss_parts, double* __restrict__ loss_accum) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_params && mode != 1) {
    const float g = grads[p];
    const float mn = m[p] + (g - m[p]) * (1.f - beta1);           // exp_avg.lerp_(g, 1 - beta1)
    const float vn = v[p] * beta2 + (1.f - beta2) * g * g;        // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
    m[p] = mn; v[p] = vn;
    params[p] -= lr_bc1 * mn / (sqrtf(vn) * inv_sqrt_bc2 + eps);  // p.addcdiv_(m, sqrt(v) / sqrt(bc2) + eps, -lr / bc1)
  }
  if (p == 0) {
    float loss;
    if (mode == 2) loss = grads[n_params];
    else {
      loss = 0.f;
      for (int i = 0; i < n_loss_parts; ++i) loss += loss_part[i];
      grads[n_params] = loss;
    }
    if (mode != 1 && loss_accum) *loss_accum += (double)loss;
  }
}

constexpr int LOSS_BLOCKS = 64;

}  // namespace

struct bb_ltrainer {
  bb_ctx* ctx = nullptr;
  int n_layers = 0, max_batch = 0, n_params = 0, max_dim = 0;
  int dims[BB_MAX_LAYERS + 1] = {0};
  int acts[BB_MAX_LAYERS] = {0};
  int w_off[BB_MAX_LAYERS] = {0}, b_off[BB_MAX_LAYERS] = {0};
  size_t a_off[BB_MAX_LAYERS + 1] = {0};  // float offsets of A_0 .. A_L in `act` (max_batch rows each)
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr, *act = nullptr, *dz[2] = {nullptr, nullptr};
  float* loss_part = nullptr;
  double* loss_accum = nullptr;
  long long step = 0;
};

namespace {

void lgemm(cudaStream_t s, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
           int M, int N, int K, const float* bias, int act) {
  const dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM);
  gemm_strided_kernel<<<grid, GT, 0, s>>>(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, bias, act);
}

// forward (+ loss, + backward into grads when `backward`)
int lforward_backward(bb_ltrainer* t, const float* x, int rows, bool backward, cudaStream_t s) {
  const int L = t->n_layers;
  BB_CUDA(cudaMemcpyAsync(t->act + t->a_off[0], x, sizeof(float) * (size_t)rows * t->dims[0], cudaMemcpyDeviceToDevice, s));
  for (int l = 0; l < L; ++l) {
    const int K = t->dims[l], N = t->dims[l + 1];
    // A_{l+1}[rows x N] = act(A_l[rows x K] . W_l[N x K]^T + b_l)
    lgemm(s, t->act + t->a_off[l], K, 1, t->params + t->w_off[l], 1, K, t->act + t->a_off[l + 1], N, rows, N, K,
          t->params + t->b_off[l], t->acts[l]);
  }
  const int C = t->dims[L];
  float* dA = t->dz[0];
  loss_seed_kernel<<<LOSS_BLOCKS, 256, 0, s>>>(t->act + t->a_off[L], x, (int64_t)rows * C, 1.f / C, backward ? dA : nullptr, t->loss_part);
  if (!backward) return (int)cudaGetLastError();
  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    const int K = t->dims[l], N = t->dims[l + 1];
    float* dZ = t->dz[cur];
    act_bwd_kernel<<<t->ctx->sm_count * 2, 256, 0, s>>>(dZ, t->act + t->a_off[l + 1], (int64_t)rows * N, t->acts[l]);
    // dW_l[N x K] = dZ^T[N x rows] . A_l[rows x K]
    lgemm(s, dZ, 1, N, t->act + t->a_off[l], K, 1, t->grads + t->w_off[l], K, N, K, rows, nullptr, BB_ACT_NONE);
    colsum_kernel<<<(N + 255) / 256, 256, 0, s>>>(dZ, rows, N, t->grads + t->b_off[l]);
    if (l > 0)  // dA_l[rows x K] = dZ[rows x N] . W_l[N x K]
      lgemm(s, dZ, N, 1, t->params + t->w_off[l], K, 1, t->dz[cur ^ 1], K, rows, K, N, nullptr, BB_ACT_NONE);
    cur ^= 1;
  }
  return (int)cudaGetLastError();
}

}  // namespace

extern "C" {

int bb_ltrainer_create(bb_ctx* ctx, int n_layers, const int* dims, const int* acts, const double* const* weights_host,
                       const double* const* biases_host, int max_batch, bb_ltrainer** out) {
  if (!ctx || !dims || !acts || !weights_host || !biases_host || !out || n_layers < 1 || n_layers > BB_MAX_LAYERS || max_batch < 1)
    return BB_ERR_INVALID;
  if (dims[0] != dims[n_layers]) return BB_ERR_INVALID;  // an autoencoder: the loss compares the output with the input
  BB_CUDA(cudaSetDevice(ctx->device));
  bb_ltrainer* t = new (std::nothrow) bb_ltrainer();
  if (!t) return BB_ERR_NOMEM;
  t->ctx = ctx; t->n_layers = n_layers; t->max_batch = max_batch;
  int p = 0;
  size_t a = 0;
  for (int l = 0; l <= n_layers; ++l) {
    t->dims[l] = dims[l];
    if (dims[l] < 1) { delete t; return BB_ERR_INVALID; }
    t->max_dim = dims[l] > t->max_dim ? dims[l] : t->max_dim;
    t->a_off[l] = a;
    a += (size_t)max_batch * dims[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    t->acts[l] = acts[l];
    t->w_off[l] = p; p += dims[l] * dims[l + 1];
    t->b_off[l] = p; p += dims[l + 1];
  }
  t->n_params = p;
  std::vector<float> hp(p);
  for (int l = 0; l < n_layers; ++l) {
    const int K = dims[l], N = dims[l + 1];
    for (int i = 0; i < N * K; ++i) hp[t->w_off[l] + i] = (float)weights_host[l][i];
    for (int n = 0; n < N; ++n) hp[t->b_off[l] + n] = (float)biases_host[l][n];
  }
  int rc = (int)cudaMalloc(&t->params, sizeof(float) * p);
  if (rc == BB_OK) rc = (int)cudaMalloc(&t->grads, sizeof(float) * (p + 1));
  if (rc == BB_OK) rc = (int)cudaMalloc(&t->m, sizeof(float) * p);
  if (rc == BB_OK) rc = (int)cudaMalloc(&t->v, sizeof(float) * p);
  if (rc == BB_OK) rc = (int)cudaMalloc(&t->act, sizeof(float) * a);
  for (int i = 0; i < 2 && rc == BB_OK; ++i) rc = (int)cudaMalloc(&t->dz[i], sizeof(float) * (size_t)max_batch * t->max_dim);
  if (rc == BB_OK) rc = (int)cudaMalloc(&t->loss_part, sizeof(float) * LOSS_BLOCKS);
  if (rc == BB_OK) rc = (int)cudaMalloc(&t->loss_accum, sizeof(double));
  if (rc == BB_OK) rc = (int)cudaMemcpy(t->params, hp.data(), sizeof(float) * p, cudaMemcpyHostToDevice);
  if (rc == BB_OK) rc = (int)cudaMemset(t->m, 0, sizeof(float) * p);
  if (rc == BB_OK) rc = (int)cudaMemset(t->v, 0, sizeof(float) * p);
  if (rc == BB_OK) rc = (int)cudaMemset(t->grads, 0, sizeof(float) * (p + 1));
  if (rc == BB_OK) rc = (int)cudaMemset(t->loss_accum, 0, sizeof(double));
  if (rc != BB_OK) { bb_ltrainer_destroy(t); return rc; }
  *out = t;
  return BB_OK;
}

int bb_ltrainer_destroy(bb_ltrainer* t) {
  if (!t) return BB_OK;
  void* ptrs[] = {t->params, t->grads, t->m, t->v, t->act, t->dz[0], t->dz[1], t->loss_part, t->loss_accum};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete t;
  return BB_OK;
}

int bb_ltrainer_param_count(const bb_ltrainer* t) { return t ? t->n_params : 0; }
float* bb_ltrainer_params_dev(bb_ltrainer* t) { return t ? t->params : nullptr; }
float* bb_ltrainer_grads_dev(bb_ltrainer* t) { return t ? t->grads : nullptr; }

int bb_ltrainer_get_params(bb_ltrainer* t, double* const* weights_host, double* const* biases_host) {
  if (!t || !weights_host || !biases_host) return BB_ERR_INVALID;
  std::vector<float> hp(t->n_params);
  BB_CUDA(cudaMemcpy(hp.data(), t->params, sizeof(float) * hp.size(), cudaMemcpyDeviceToHost));
  for (int l = 0; l < t->n_layers; ++l) {
    const int K = t->dims[l], N = t->dims[l + 1];
    for (int i = 0; i < N * K; ++i) weights_host[l][i] = (double)hp[t->w_off[l] + i];
    for (int n = 0; n < N; ++n) biases_host[l][n] = (double)hp[t->b_off[l] + n];
  }
  return BB_OK;
}

int bb_ltrainer_step(bb_ltrainer* t, const float* x_dev, int batch_rows, const bb_train_hyper* h, int phase,
                     double* loss_accum_dev, bb_stream_t stream) {
  if (!t || !h || !x_dev || batch_rows < 1 || batch_rows > t->max_batch || phase < 0 || phase > 2) return BB_ERR_INVALID;
  if (h->l1) return BB_ERR_UNSUPPORTED;  // the reference never trains the L1 term (training.py:83-89); fused kernels have the opt-in
  cudaStream_t s = (cudaStream_t)stream;
  if (phase != 2) {
    const int rc = lforward_backward(t, x_dev, batch_rows, true, s);
    if (rc != BB_OK) return rc;
  }
  float lr_bc1 = 0.f, inv_sqrt_bc2 = 0.f;
  if (phase != 1) {
    t->step += 1;
    lr_bc1 = (float)(h->lr / (1.0 - std::pow(h->beta1, (double)t->step)));
    inv_sqrt_bc2 = (float)(1.0 / std::sqrt(1.0 - std::pow(h->beta2, (double)t->step)));
  }
  ladam_kernel<<<(t->n_params + 255) / 256, 256, 0, s>>>(t->n_params, phase, t->grads, t->params, t->m, t->v, lr_bc1, inv_sqrt_bc2,
                                                        (float)h->beta1, (float)h->beta2, (float)h->eps, t->loss_part, LOSS_BLOCKS,
                                                        loss_accum_dev);
  return (int)cudaGetLastError();
}

int bb_ltrainer_epoch(bb_ltrainer* t, const float* x_dev, int64_t n_rows, int batch, const bb_train_hyper* h,
                      double* epoch_loss_host, bb_stream_t stream) {
  if (!t || !h || !x_dev || n_rows < 1 || batch < 1 || batch > t->max_batch || !epoch_loss_host) return BB_ERR_INVALID;
  if (h->world_size > 1) return BB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  BB_CUDA(cudaMemsetAsync(t->loss_accum, 0, sizeof(double), s));
  int64_t n_batches = 0;
  for (int64_t r0 = 0; r0 < n_rows; r0 += batch, ++n_batches) {
    const int rows = (int)(n_rows - r0 < batch ? n_rows - r0 : batch);
    const int rc = bb_ltrainer_step(t, x_dev + (size_t)r0 * t->dims[0], rows, h, 0, t->loss_accum, s);
    if (rc != BB_OK) return rc;
  }
  double total = 0.0;
  BB_CUDA(cudaMemcpyAsync(&total, t->loss_accum, sizeof(double), cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaStreamSynchronize(s));
  *epoch_loss_host = total / (double)n_batches;
  return BB_OK;
}

int bb_ltrainer_validate(bb_ltrainer* t, const float* x_dev, int64_t n_rows, int batch, double* epoch_loss_host,
                         bb_stream_t stream) {
  if (!t || !x_dev || n_rows < 1 || batch < 1 || batch > t->max_batch || !epoch_loss_host) return BB_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  BB_CUDA(cudaMemsetAsync(t->loss_accum, 0, sizeof(double), s));
  int64_t n_batches = 0;
  for (int64_t r0 = 0; r0 < n_rows; r0 += batch, ++n_batches) {
    const int rows = (int)(n_rows - r0 < batch ? n_rows - r0 : batch);
    int rc = lforward_backward(t, x_dev + (size_t)r0 * t->dims[0], rows, false, s);
    if (rc != BB_OK) return rc;
    // mode 1 folds the loss partials into grads[n_params]; mode 2 with zero parameters adds that slot to the accumulator
    ladam_kernel<<<1, 256, 0, s>>>(t->n_params, 1, t->grads, nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 0.f, t->loss_part,
                                   LOSS_BLOCKS, nullptr);
    ladam_kernel<<<1, 256, 0, s>>>(0, 2, t->grads + t->n_params, nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, 0.f, nullptr, 0,
                                   t->loss_accum);
    BB_CUDA(cudaGetLastError());
  }
  double total = 0.0;
  BB_CUDA(cudaMemcpyAsync(&total, t->loss_accum, sizeof(double), cudaMemcpyDeviceToHost, s));
  BB_CUDA(cudaStreamSynchronize(s));
  *epoch_loss_host = total / (double)n_batches;
  return BB_OK;
}

}  // extern "C"
