// C-ABI entry points (include/baler_b200.h): context, model packing, encode / decode dispatch and
// the host-buffer compress / decompress pipelines.
#include <algorithm>
#include <cstring>
#include <new>
#include <thread>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <pthread.h>
#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

#include <cstdlib>

#include "bb_common.cuh"

extern "C" {

int bb_version(void) { return BB_VERSION; }

const char* bb_strerror(int code) {
  switch (code) {
    case BB_OK: return "ok";
    case BB_ERR_INVALID: return "baler_b200: invalid argument";
    case BB_ERR_UNSUPPORTED: return "baler_b200: shape or mode not supported by the CUDA kernels";
    case BB_ERR_NOMEM: return "baler_b200: out of memory";
    case BB_ERR_NODEVICE: return "baler_b200: no CUDA device (there is no CPU fallback)";
    case BB_ERR_OVERFLOW: return "baler_b200: fp16 split range guard tripped; use BB_PREC_FP32";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "baler_b200: unknown error";
  }
}

int bb_ctx_create(int device, bb_ctx** out) {
  if (!out) return BB_ERR_INVALID;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return BB_ERR_NODEVICE;
  if (device < 0 || device >= n) return BB_ERR_INVALID;
  BB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return BB_ERR_UNSUPPORTED;  // sm_100a SASS only
  bb_ctx* c = new (std::nothrow) bb_ctx();
  if (!c) return BB_ERR_NOMEM;
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->smem_optin = prop.sharedMemPerBlockOptin;
  *out = c;
  return BB_OK;
}

int bb_ctx_destroy(bb_ctx* ctx) {
  if (!ctx) return BB_OK;
  if (ctx->minmax_scratch) cudaFree(ctx->minmax_scratch);
  delete ctx;
  return BB_OK;
}

int bb_ctx_sm_count(const bb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

}  // extern "C"

namespace {

int fill_chain(Chain* c, int n_layers, const int* dims, const int* acts, const double* const* w,
               const double* const* b) {
  if (n_layers < 1 || n_layers > BB_MAX_LAYERS || !dims || !acts || !w || !b) return BB_ERR_INVALID;
  ChainDesc& d = c->desc;
  memset(&d, 0, sizeof(d));
  d.n_layers = n_layers;
  d.in_dim = dims[0];
  d.out_dim = dims[n_layers];
  for (int l = 0; l < n_layers; ++l) {
    if (dims[l] < 1 || dims[l + 1] < 1 || acts[l] < BB_ACT_NONE || acts[l] > BB_ACT_RELU || !w[l] || !b[l])
      return BB_ERR_INVALID;
    d.layer[l].K = dims[l];
    d.layer[l].N = dims[l + 1];
    d.layer[l].act = acts[l];
    c->w_host[l].assign(w[l], w[l] + (size_t)dims[l] * dims[l + 1]);
    c->b_host[l].assign(b[l], b[l] + dims[l + 1]);
  }
  return BB_OK;
}

void free_chain(Chain* c) {
  // (pointers are cleared: bb_model_destroy also runs model_trim, which releases the scratch of a live model)
  auto release = [](auto*& p) {
    if (p) cudaFree(p);
    p = nullptr;
  };
  release(c->blob_dev);
  release(c->tc_blob_dev);
  release(c->lay_blob_dev);
  release(c->lay_scratch);
  c->lay_scratch_bytes = 0;
  release(c->lay_tc_blob_dev);
  release(c->g5_blob_dev);
  release(c->g5_bias_dev);
  bb_tc_release(c);
}

// the chain has a tensor-core form: one of the fused tcgen05 kernels, or the layered mma.sync GEMMs for shapes too
// large for any fused kernel
bool chain_has_tc(const Chain* c) { return c->tc_ok || (!c->f32_ok && c->lay_tc_ok); }

int resolve_precision(const bb_model* m, const Chain* c, int precision) {
  if (precision == BB_PREC_AUTO) return chain_has_tc(c) ? BB_PREC_SPLIT16 : BB_PREC_FP32;
  return precision;
}

int run_chain(bb_model* m, const Chain* c, const void* in, int in_dtype, int64_t n_rows,
              const float* pre_min, const float* pre_range, const float* post_min,
              const float* post_range, void* out, int out_dtype, int precision, cudaStream_t stream) {
  if (!m || (!in && n_rows) || (!out && n_rows) || n_rows < 0) return BB_ERR_INVALID;
  if ((pre_min == nullptr) != (pre_range == nullptr) || (post_min == nullptr) != (post_range == nullptr))
    return BB_ERR_INVALID;
  if ((in_dtype != BB_F32 && in_dtype != BB_F16) || (out_dtype != BB_F32 && out_dtype != BB_F16)) return BB_ERR_INVALID;
  const int p = resolve_precision(m, c, precision);
  if (p == BB_PREC_FP32) {
    if (!c->f32_ok)  // weights do not fit shared memory: one GEMM launch per layer
      return bb_chain_layered_launch(m->ctx, c, in, in_dtype, n_rows, pre_min, pre_range, post_min, post_range, out,
                                     out_dtype, 0, nullptr, stream);
    return bb_chain_f32_launch(m->ctx, c, in, in_dtype, n_rows, pre_min, pre_range, post_min, post_range,
                               out, out_dtype, stream);
  }
  if (p == BB_PREC_SPLIT16 && !c->tc_ok && chain_has_tc(c))
    return bb_chain_layered_launch(m->ctx, c, in, in_dtype, n_rows, pre_min, pre_range, post_min, post_range, out, out_dtype, 1,
                                   m->flag_dev, stream);
  if (p == BB_PREC_SPLIT16 || p == BB_PREC_FAST16) {
    if (!c->tc_ok) return BB_ERR_UNSUPPORTED;
    return bb_tc_launch(m->ctx, c, in, in_dtype, n_rows, pre_min, pre_range, post_min, post_range, out,
                        out_dtype, p == BB_PREC_FAST16, m->flag_dev, stream);
  }
  return BB_ERR_INVALID;
}

size_t dtype_size(int dt) { return dt == BB_F64 ? 8 : (dt == BB_F16 ? 2 : 4); }

// host threads this process may use for the scan: the CPUs it is allowed on (a launcher may have bound the rank to its
// GPU's NUMA node), shared with the other ranks of the node (LOCAL_WORLD_SIZE, set by torchrun), at most 16
unsigned host_scan_threads() {
  unsigned hw = std::max(1u, std::thread::hardware_concurrency());
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) hw = std::min<unsigned>(hw, (unsigned)CPU_COUNT(&set));
#endif
  unsigned ranks = 1;
  if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = (unsigned)std::max(1, atoi(e));
  return std::max(1u, std::min(16u, hw / ranks));
}

// fp32 <-> fp64 conversion of a chunk on host threads (the reference stores float64 latents and reconstructions:
// helper.py:565).  AVX2 with non-temporal stores where the destination allows it: the output is written once and not
// read here, so the cache lines need not be fetched first (a plain store of a 3.8 GB reconstruction reads 3.8 GB too).
#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) void widen_range_avx2(const float* src, double* dst, size_t lo, size_t hi) {
  size_t i = lo;
  for (; i < hi && (reinterpret_cast<uintptr_t>(dst + i) & 31u); ++i) dst[i] = (double)src[i];
  for (; i + 8 <= hi; i += 8) {
    const __m256 v = _mm256_loadu_ps(src + i);
    _mm256_stream_pd(dst + i, _mm256_cvtps_pd(_mm256_castps256_ps128(v)));
    _mm256_stream_pd(dst + i + 4, _mm256_cvtps_pd(_mm256_extractf128_ps(v, 1)));
  }
  for (; i < hi; ++i) dst[i] = (double)src[i];
  _mm_sfence();
}
__attribute__((target("avx2"))) void narrow_range_avx2(const double* src, float* dst, size_t lo, size_t hi) {
  size_t i = lo;
  for (; i < hi && (reinterpret_cast<uintptr_t>(dst + i) & 31u); ++i) dst[i] = (float)src[i];
  for (; i + 8 <= hi; i += 8) {
    const __m128 a = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i)), b = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4));
    _mm256_stream_ps(dst + i, _mm256_set_m128(b, a));
  }
  for (; i < hi; ++i) dst[i] = (float)src[i];
  _mm_sfence();
}
#endif

// Persistent worker threads for the per-chunk conversions.  Spawning and joining 16 std::threads costs ~0.45 ms, and a
// float64 round trip converts three arrays per chunk (48 dispatches for 16 chunks: a sixth of its wall time); the pool
// hands a job to sleeping workers in a few microseconds.  One job at a time (callers queue on run_mu); the caller works
// too.  Never destroyed (no join against a library being unloaded at exit); a forked child starts a pool of its own,
// because the parent's workers do not exist there and its mutexes may be held.
class HostPool {
 public:
  static HostPool& get(unsigned want) {
    std::lock_guard<std::mutex> lk(singleton_mu());
    static std::once_flag fork_once;
    std::call_once(fork_once, [] { pthread_atfork(nullptr, nullptr, [] { slot() = nullptr; new (&singleton_mu()) std::mutex(); }); });
    HostPool*& p = slot();
    if (p == nullptr) p = new HostPool();
    p->grow(want);
    return *p;
  }
  // fn(t) for t in [0, n_tasks): tasks are claimed one by one by the workers and by the calling thread
  void run(unsigned n_tasks, const std::function<void(unsigned)>& fn) {
    if (n_tasks == 0) return;
    std::lock_guard<std::mutex> one_job(run_mu_);
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = &fn; n_tasks_ = n_tasks; next_ = 0; left_ = n_tasks;
      ++generation_;
    }
    cv_work_.notify_all();
    work_on_current_job();
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return left_ == 0; });
    job_ = nullptr;
  }

 private:
  static HostPool*& slot() { static HostPool* p = nullptr; return p; }
  static std::mutex& singleton_mu() { static std::mutex* m = new std::mutex(); return *m; }
  void grow(unsigned want) {  // (under singleton_mu) workers = want - 1: the caller is the last one
    while (workers_ + 1 < want) {
      std::thread([this] { worker(); }).detach();
      ++workers_;
    }
  }
  void work_on_current_job() {
    for (;;) {
      unsigned t;
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (job_ == nullptr || next_ >= n_tasks_) return;
        t = next_++;
      }
      (*job_)(t);
      std::lock_guard<std::mutex> lk(mu_);
      if (--left_ == 0) cv_done_.notify_all();
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_work_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
      }
      work_on_current_job();
    }
  }
  std::mutex mu_, run_mu_;
  std::condition_variable cv_work_, cv_done_;
  const std::function<void(unsigned)>* job_ = nullptr;
  unsigned n_tasks_ = 0, next_ = 0, left_ = 0, workers_ = 0;
  uint64_t generation_ = 0;
};

template <typename Fn>
void host_parallel_ranges(size_t n, Fn fn) {
  const unsigned hw = host_scan_threads();
  const size_t per = ((n + hw - 1) / hw + 7) & ~(size_t)7;
  if (n < (1u << 16) || hw == 1) {
    fn((size_t)0, n);
    return;
  }
  const unsigned n_tasks = (unsigned)((n + per - 1) / per);
  const std::function<void(unsigned)> task = [&](unsigned t) {
    const size_t lo = (size_t)t * per, hi = std::min(n, lo + per);
    if (lo < hi) fn(lo, hi);
  };
  HostPool::get(hw).run(n_tasks, task);
}

void widen_f32_f64(const float* src, double* dst, size_t n) {
  host_parallel_ranges(n, [=](size_t lo, size_t hi) {
#if defined(__x86_64__) && defined(__GNUC__)
    if (__builtin_cpu_supports("avx2")) { widen_range_avx2(src, dst, lo, hi); return; }
#endif
    for (size_t i = lo; i < hi; ++i) dst[i] = (double)src[i];
  });
}

void narrow_f64_f32(const double* src, float* dst, size_t n) {
  host_parallel_ranges(n, [=](size_t lo, size_t hi) {
#if defined(__x86_64__) && defined(__GNUC__)
    if (__builtin_cpu_supports("avx2")) { narrow_range_avx2(src, dst, lo, hi); return; }
#endif
    for (size_t i = lo; i < hi; ++i) dst[i] = (float)src[i];
  });
}

constexpr int64_t PIPE_CHUNK_ROWS = 1 << 21;  // rows per pipeline chunk (2 Mi rows: 192 MiB in, 120 MiB out for 24 -> 15)
constexpr int64_t PIPE_MIN_CHUNK_ROWS = 1 << 18;

// Chunking of a host table: at least ~16 chunks so that filling and draining the H2D / kernel / D2H pipeline stays a small
// part of the call (a 12.5M-row shard cut into 2Mi-row chunks spends a third of its time in fill and drain), but never
// below 256 Ki rows (per-chunk launch and event overhead) nor above 2 Mi rows (staging memory)
inline int64_t pipe_chunk_rows(int64_t n_rows) {
  int64_t c = (n_rows + 15) / 16;
  c = std::max<int64_t>(c, PIPE_MIN_CHUNK_ROWS);
  c = std::min<int64_t>(c, PIPE_CHUNK_ROWS);
  c = (c + 127) & ~(int64_t)127;  // whole 128-row tiles: chunk boundaries stay 16-byte aligned for any row width
  return std::max<int64_t>(1, std::min<int64_t>(n_rows, c));
}

int pipe_init(bb_model* m, size_t in_bytes, size_t out_bytes) {
  if (!m->s_compute) {
    BB_CUDA(cudaStreamCreateWithFlags(&m->s_copy_in, cudaStreamNonBlocking));
    BB_CUDA(cudaStreamCreateWithFlags(&m->s_compute, cudaStreamNonBlocking));
    BB_CUDA(cudaStreamCreateWithFlags(&m->s_copy_out, cudaStreamNonBlocking));
    for (int s = 0; s < 2; ++s) {
      BB_CUDA(cudaEventCreateWithFlags(&m->ev_in[s], cudaEventDisableTiming));
      BB_CUDA(cudaEventCreateWithFlags(&m->ev_done[s], cudaEventDisableTiming));
      BB_CUDA(cudaEventCreateWithFlags(&m->ev_out[s], cudaEventDisableTiming));
    }
    BB_CUDA(cudaMalloc(&m->feat_dev, 3 * sizeof(float) * std::max(m->enc.desc.in_dim, m->dec.desc.out_dim)));
  }
  const size_t need[2] = {in_bytes, out_bytes};
  for (int io = 0; io < 2; ++io) {
    if (m->stage_bytes[io] < need[io]) {
      for (int s = 0; s < 2; ++s) {
        if (m->stage_dev[io][s]) cudaFree(m->stage_dev[io][s]);
        m->stage_dev[io][s] = nullptr;
        BB_CUDA(cudaMalloc(&m->stage_dev[io][s], need[io]));
      }
      m->stage_bytes[io] = need[io];
    }
  }
  return BB_OK;
}

// pinned float32 bounce buffers (two slots) of the float64 host paths, grown on demand and kept
int pinned_init(bb_model* m, int io, size_t bytes) {
  if (m->pinned_bytes[io] >= bytes) return BB_OK;
  for (int s = 0; s < 2; ++s) {
    if (m->pinned_f32[io][s]) cudaFreeHost(m->pinned_f32[io][s]);
    m->pinned_f32[io][s] = nullptr;
  }
  m->pinned_bytes[io] = 0;
  for (int s = 0; s < 2; ++s)
    if (cudaMallocHost(&m->pinned_f32[io][s], bytes) != cudaSuccess) { cudaGetLastError(); return BB_ERR_NOMEM; }
  m->pinned_bytes[io] = bytes;
  return BB_OK;
}

// Column min / max of a host table on host threads (data_processing.find_minmax, data_processing.py:113-130), for the
// resident path of bb_compress_host: the scan reads host memory at several times the PCIe rate, so the features are known
// long before the upload of the table ends and the latent download overlaps the rest of the upload (full duplex) instead
// of following it.  Same results as colminmax_kernel: order-preserving integer keys (-0 < +0), NaN propagates per column.
// key(u) = u ^ ((u >> 31) & 0x7fffffff) as a SIGNED int is monotonic in the float value.
inline int32_t mm_key(uint32_t u) { return (int32_t)(u ^ ((uint32_t)((int32_t)u >> 31) & 0x7fffffffu)); }
inline float mm_unkey(int32_t k) {
  const uint32_t u = k >= 0 ? (uint32_t)k : ((uint32_t)k ^ 0x7fffffffu);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// scalar scan of elements [e0, e1) of the flat table; column = element index % F
void mm_scan_scalar(const uint32_t* p, int64_t e0, int64_t e1, int F, int32_t* mn, int32_t* mx, uint32_t* nan) {
  int c = (int)(e0 % F);
  for (int64_t e = e0; e < e1; ++e) {
    const uint32_t u = p[e];
    if ((u & 0x7fffffffu) > 0x7f800000u) {
      nan[c] = 1u;
    } else {
      const int32_t k = mm_key(u);
      mn[c] = k < mn[c] ? k : mn[c];
      mx[c] = k > mx[c] ? k : mx[c];
    }
    if (++c == F) c = 0;
  }
}

#if defined(__x86_64__) && defined(__GNUC__)
// AVX2: the column pattern of 8 consecutive elements repeats every L = lcm(F, 8) elements, so L / 8 vector accumulators
// cover a period; lane j of accumulator v belongs to column (8 v + j) % F
// (accumulators as plain int arrays of 8 L8 entries; kept in registers for the scan when L8 <= 4: the CMS table has L8 = 3)
__attribute__((target("avx2"))) void mm_scan_avx2(const uint32_t* p, int64_t n_periods, int L8, int32_t* mn, int32_t* mx, int32_t* nanv) {
  const __m256i mag = _mm256_set1_epi32(0x7fffffff), inf = _mm256_set1_epi32(0x7f800000);
  const __m256i imax = _mm256_set1_epi32(0x7fffffff), imin = _mm256_set1_epi32((int)0x80000000u);
#define BB_MM_STEP(U, MN, MX, NV)                                                              \
  do {                                                                                         \
    const __m256i key_ = _mm256_xor_si256(U, _mm256_and_si256(_mm256_srai_epi32(U, 31), mag)); \
    const __m256i isnan_ = _mm256_cmpgt_epi32(_mm256_and_si256(U, mag), inf);                  \
    NV = _mm256_or_si256(NV, isnan_);                                                          \
    MN = _mm256_min_epi32(MN, _mm256_blendv_epi8(key_, imax, isnan_));                         \
    MX = _mm256_max_epi32(MX, _mm256_blendv_epi8(key_, imin, isnan_));                         \
  } while (0)
  if (L8 <= 4) {
    __m256i rmn[4], rmx[4], rnv[4];
    for (int v = 0; v < L8; ++v) {
      rmn[v] = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(mn + 8 * v));
      rmx[v] = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(mx + 8 * v));
      rnv[v] = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(nanv + 8 * v));
    }
    for (int64_t q = 0; q < n_periods; ++q, p += (size_t)L8 * 8)
      for (int v = 0; v < L8; ++v) {
        const __m256i u = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 8 * v));
        BB_MM_STEP(u, rmn[v], rmx[v], rnv[v]);
      }
    for (int v = 0; v < L8; ++v) {
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(mn + 8 * v), rmn[v]);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(mx + 8 * v), rmx[v]);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(nanv + 8 * v), rnv[v]);
    }
    return;
  }
  for (int64_t q = 0; q < n_periods; ++q, p += (size_t)L8 * 8)
    for (int v = 0; v < L8; ++v) {
      const __m256i u = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + 8 * v));
      __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(mn + 8 * v));
      __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(mx + 8 * v));
      __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(nanv + 8 * v));
      BB_MM_STEP(u, a, b, c);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(mn + 8 * v), a);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(mx + 8 * v), b);
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(nanv + 8 * v), c);
    }
#undef BB_MM_STEP
}
#endif

void host_colminmax(const float* x, int64_t n_rows, int F, float* mn_out, float* mx_out) {
  const unsigned hw = host_scan_threads();
  int g = F, b = 8;
  while (b) { const int r = g % b; g = b; b = r; }
  const int period_rows = 8 / g, L8 = F / g;  // rows / 8-lane vectors of one period
  bool avx2 = false;
#if defined(__x86_64__) && defined(__GNUC__)
  avx2 = __builtin_cpu_supports("avx2") && L8 <= 1024;
#endif
  const int64_t periods = n_rows / period_rows;
  const int64_t per = (periods + hw - 1) / hw;
  std::vector<std::vector<int32_t>> pmn(hw, std::vector<int32_t>((size_t)F, 0x7fffffff)), pmx(hw, std::vector<int32_t>((size_t)F, (int32_t)0x80000000u));
  std::vector<std::vector<uint32_t>> pnan(hw, std::vector<uint32_t>((size_t)F, 0u));
  const uint32_t* p = reinterpret_cast<const uint32_t*>(x);
  auto work = [&](unsigned t) {
    const int64_t q0 = std::min<int64_t>(periods, (int64_t)t * per), q1 = std::min<int64_t>(periods, q0 + per);
    int32_t* mn = pmn[t].data();
    int32_t* mx = pmx[t].data();
    uint32_t* nan = pnan[t].data();
    int64_t e0 = q0 * period_rows * F;
    const int64_t e1 = (t + 1 == hw ? n_rows : q1 * period_rows) * (int64_t)F;  // the last thread also takes the ragged tail
#if defined(__x86_64__) && defined(__GNUC__)
    if (avx2 && q1 > q0) {
      std::vector<int32_t> acc(3 * (size_t)L8 * 8);
      int32_t* vmn = acc.data();
      int32_t* vmx = vmn + (size_t)L8 * 8;
      int32_t* vnan = vmx + (size_t)L8 * 8;
      for (int i = 0; i < L8 * 8; ++i) { vmn[i] = 0x7fffffff; vmx[i] = (int32_t)0x80000000u; vnan[i] = 0; }
      mm_scan_avx2(p + e0, q1 - q0, L8, vmn, vmx, vnan);
      for (int i = 0; i < L8 * 8; ++i) {
        const int c = i % F;
        mn[c] = vmn[i] < mn[c] ? vmn[i] : mn[c];
        mx[c] = vmx[i] > mx[c] ? vmx[i] : mx[c];
        nan[c] |= vnan[i] ? 1u : 0u;
      }
      e0 = q1 * period_rows * F;
    }
#endif
    mm_scan_scalar(p, e0, e1, F, mn, mx, nan);
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < hw; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& h : th) h.join();
  for (int c = 0; c < F; ++c) {
    int32_t mn = 0x7fffffff, mx = (int32_t)0x80000000u;
    uint32_t nan = 0u;
    for (unsigned t = 0; t < hw; ++t) {
      mn = std::min(mn, pmn[t][c]); mx = std::max(mx, pmx[t][c]); nan |= pnan[t][c];
    }
    const uint32_t qnan = 0x7fc00000u;
    float fn;
    memcpy(&fn, &qnan, 4);
    mn_out[c] = nan ? fn : mm_unkey(mn);
    mx_out[c] = nan ? fn : mm_unkey(mx);
  }
}

// device copy of the whole table between the two passes of bb_compress_host, when it fits in half of the free memory
float* resident_get(bb_model* m, size_t bytes) {
  if (getenv("BALER_B200_NO_RESIDENT")) return nullptr;  // test hook: take the two-pass streaming path of tables that do not fit
  if (m->resident_bytes >= bytes) return m->resident_dev;
  if (m->resident_dev) { cudaFree(m->resident_dev); m->resident_dev = nullptr; m->resident_bytes = 0; }
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || bytes >= free_b / 2) return nullptr;
  if (cudaMalloc(&m->resident_dev, bytes) != cudaSuccess) { cudaGetLastError(); m->resident_dev = nullptr; return nullptr; }
  m->resident_bytes = bytes;
  return m->resident_dev;
}

void model_trim(bb_model* m) {
  if (m->resident_dev) cudaFree(m->resident_dev);
  m->resident_dev = nullptr; m->resident_bytes = 0;
  for (int io = 0; io < 2; ++io) {
    for (int s = 0; s < 2; ++s) {
      if (m->pinned_f32[io][s]) cudaFreeHost(m->pinned_f32[io][s]);
      m->pinned_f32[io][s] = nullptr;
    }
    m->pinned_bytes[io] = 0;
  }
  // the activation scratch of the layered GEMM path (up to 2 x 1.24 GB per direction for 2000-wide layers; cudaFree waits
  // for the device, so work still using it on any stream has finished)
  for (Chain* c : {&m->enc, &m->dec}) {
    if (c->lay_scratch) cudaFree(c->lay_scratch);
    c->lay_scratch = nullptr; c->lay_scratch_bytes = 0;
  }
}

// range = max - min on device (float32 subtraction, like numpy on a float32 table)
__global__ void range_kernel(const float* mn, const float* mx, float* rg, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < c) rg[i] = __fsub_rn(mx[i], mn[i]);
}

}  // namespace

extern "C" {

int bb_model_range_flag(bb_model* m, int reset, int* out);

int bb_model_create_dense(bb_ctx* ctx, int n_enc_layers, const int* enc_dims, const int* enc_acts,
                          const double* const* enc_w, const double* const* enc_b, int n_dec_layers,
                          const int* dec_dims, const int* dec_acts, const double* const* dec_w,
                          const double* const* dec_b, bb_model** out) {
  if (!ctx || !out) return BB_ERR_INVALID;
  BB_CUDA(cudaSetDevice(ctx->device));
  bb_model* m = new (std::nothrow) bb_model();
  if (!m) return BB_ERR_NOMEM;
  m->ctx = ctx;
  int rc = fill_chain(&m->enc, n_enc_layers, enc_dims, enc_acts, enc_w, enc_b);
  if (rc == BB_OK) rc = fill_chain(&m->dec, n_dec_layers, dec_dims, dec_acts, dec_w, dec_b);
  if (rc == BB_OK && m->enc.desc.out_dim != m->dec.desc.in_dim) rc = BB_ERR_INVALID;
  for (Chain* c : {&m->enc, &m->dec}) {
    if (rc != BB_OK) break;
    rc = bb_chain_f32_prepare(ctx, c);
    c->f32_ok = rc == BB_OK;
    if (rc == BB_ERR_UNSUPPORTED) rc = bb_chain_layered_prepare(ctx, c);  // too big for the fused kernels
  }
  if (rc == BB_OK) rc = (int)cudaMalloc(&m->flag_dev, sizeof(int));
  if (rc == BB_OK) rc = (int)cudaMemset(m->flag_dev, 0, sizeof(int));
  if (rc == BB_OK) {
    // the tensor-core path is optional per shape; failure to prepare it is not an error
    int t = bb_tc_prepare(ctx, &m->enc);
    if (t != BB_OK && t != BB_ERR_UNSUPPORTED) rc = t;
    t = bb_tc_prepare(ctx, &m->dec);
    if (t != BB_OK && t != BB_ERR_UNSUPPORTED) rc = t;
  }
  if (rc != BB_OK) {
    bb_model_destroy(m);
    return rc;
  }
  *out = m;
  return BB_OK;
}

int bb_model_trim(bb_model* m) {
  if (!m) return BB_ERR_INVALID;
  cudaSetDevice(m->ctx->device);
  if (m->s_compute) { cudaStreamSynchronize(m->s_copy_in); cudaStreamSynchronize(m->s_compute); cudaStreamSynchronize(m->s_copy_out); }
  model_trim(m);
  return BB_OK;
}

int bb_model_destroy(bb_model* m) {
  if (!m) return BB_OK;
  free_chain(&m->enc);
  free_chain(&m->dec);
  if (m->flag_dev) cudaFree(m->flag_dev);
  if (m->feat_dev) cudaFree(m->feat_dev);
  model_trim(m);
  for (int io = 0; io < 2; ++io)
    for (int s = 0; s < 2; ++s)
      if (m->stage_dev[io][s]) cudaFree(m->stage_dev[io][s]);
  for (int s = 0; s < 2; ++s) {
    if (m->ev_in[s]) cudaEventDestroy(m->ev_in[s]);
    if (m->ev_done[s]) cudaEventDestroy(m->ev_done[s]);
    if (m->ev_out[s]) cudaEventDestroy(m->ev_out[s]);
  }
  if (m->s_copy_in) cudaStreamDestroy(m->s_copy_in);
  if (m->s_compute) cudaStreamDestroy(m->s_compute);
  if (m->s_copy_out) cudaStreamDestroy(m->s_copy_out);
  delete m;
  return BB_OK;
}

// Test hook (not part of the public header): run one direction on the tcgen05 path and dump the scaled
// accumulator of program step `dbg_step` (before activation) to dbg_out_dev [n_rows x step width].
int bb_debug_tc_chain(bb_model* m, int decode, const void* in_dev, int64_t n_rows, void* out_dev, int fast,
                      int dbg_step, float* dbg_out_dev, int force_groups, bb_stream_t stream) {
  if (!m) return BB_ERR_INVALID;
  const Chain* c = decode ? &m->dec : &m->enc;
  return bb_tc_launch_dbg(m->ctx, c, in_dev, BB_F32, n_rows, nullptr, nullptr, nullptr, nullptr, out_dev, BB_F32, fast,
                          m->flag_dev, dbg_step, dbg_out_dev, force_groups, (cudaStream_t)stream);
}

int bb_model_range_flag(bb_model* m, int reset, int* out) {
  if (!m || !out) return BB_ERR_INVALID;
  BB_CUDA(cudaMemcpy(out, m->flag_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (reset && *out) BB_CUDA(cudaMemset(m->flag_dev, 0, sizeof(int)));
  return BB_OK;
}

int bb_model_n_features(const bb_model* m) { return m ? m->enc.desc.in_dim : 0; }
int bb_model_z_dim(const bb_model* m) { return m ? m->enc.desc.out_dim : 0; }
int bb_model_chain_precision(const bb_model* m, int direction) {
  if (!m || direction < 0 || direction > 1) return BB_ERR_INVALID;
  return chain_has_tc(direction == 0 ? &m->enc : &m->dec) ? BB_PREC_SPLIT16 : BB_PREC_FP32;
}

int bb_model_auto_precision(const bb_model* m) {
  return (m && chain_has_tc(&m->enc) && chain_has_tc(&m->dec)) ? BB_PREC_SPLIT16 : BB_PREC_FP32;
}

int bb_colminmax_f32(bb_ctx* ctx, const float* x_dev, int64_t n_rows, int n_cols, float* min_dev,
                     float* max_dev, bb_stream_t stream) {
  if (!ctx || (!x_dev && n_rows) || !min_dev || !max_dev) return BB_ERR_INVALID;
  return bb_colminmax_launch(ctx, x_dev, n_rows, n_cols, min_dev, max_dev, 1, (cudaStream_t)stream);
}

int bb_encode_f32(bb_model* m, const float* x_dev, int64_t n_rows, const float* min_dev,
                  const float* range_dev, void* z_dev, int z_dtype, int precision, bb_stream_t stream) {
  if (!m) return BB_ERR_INVALID;
  return run_chain(m, &m->enc, x_dev, BB_F32, n_rows, min_dev, range_dev, nullptr, nullptr, z_dev, z_dtype,
                   precision, (cudaStream_t)stream);
}

int bb_decode_f32(bb_model* m, const void* z_dev, int z_dtype, int64_t n_rows, const float* min_dev,
                  const float* range_dev, float* y_dev, int precision, bb_stream_t stream) {
  if (!m) return BB_ERR_INVALID;
  return run_chain(m, &m->dec, z_dev, z_dtype, n_rows, nullptr, nullptr, min_dev, range_dev, y_dev, BB_F32,
                   precision, (cudaStream_t)stream);
}

int bb_host_convert(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n) {
  if (n < 0 || ((!src || !dst) && n)) return BB_ERR_INVALID;
  if (src_dtype == BB_F32 && dst_dtype == BB_F64) widen_f32_f64((const float*)src, (double*)dst, (size_t)n);
  else if (src_dtype == BB_F64 && dst_dtype == BB_F32) narrow_f64_f32((const double*)src, (float*)dst, (size_t)n);
  else return BB_ERR_INVALID;
  return BB_OK;
}

int bb_host_colminmax_f32(const float* x_host, int64_t n_rows, int n_cols, float* min_out, float* max_out) {
  if (!x_host || !min_out || !max_out || n_rows < 1 || n_cols < 1) return BB_ERR_INVALID;
  host_colminmax(x_host, n_rows, n_cols, min_out, max_out);
  return BB_OK;
}

int bb_compress_host(bb_model* m, const float* x_host, int64_t n_rows, float* features_host,
                     int recompute_minmax, void* z_host, int z_dtype, int precision) {
  if (!m || (!x_host && n_rows) || (!z_host && n_rows) || n_rows < 0) return BB_ERR_INVALID;
  if (z_dtype != BB_F32 && z_dtype != BB_F16 && z_dtype != BB_F64) return BB_ERR_INVALID;
  if (recompute_minmax && !features_host) return BB_ERR_INVALID;
  BB_CUDA(cudaSetDevice(m->ctx->device));
  const int F = m->enc.desc.in_dim, Z = m->enc.desc.out_dim;
  const int dev_dtype = (z_dtype == BB_F16) ? BB_F16 : BB_F32;
  const int64_t chunk = pipe_chunk_rows(n_rows);
  const int64_t n_chunks = n_rows ? (n_rows + chunk - 1) / chunk : 0;
  const size_t in_b = (size_t)chunk * F * sizeof(float), out_b = (size_t)chunk * Z * dtype_size(dev_dtype);
  int rc = pipe_init(m, in_b, out_b);
  if (rc != BB_OK) return rc;
  float* fmin = m->feat_dev;
  float* frange = m->feat_dev + F;
  float* fmax = m->feat_dev + 2 * F;
  const bool norm = features_host != nullptr;

  // Column statistics of THIS table (helper.py:500-502): either a first streaming pass that keeps
  // the table resident when it fits, or the features handed in.
  float* resident = nullptr;
  std::vector<cudaEvent_t> ev_up;  // resident path: upload of chunk k complete
  if (norm && recompute_minmax && n_rows) {
    resident = resident_get(m, (size_t)n_rows * F * sizeof(float));
    // the host scan pays when 16 host threads are free for it (measured on a 32-CPU node: 2 ranks 483 M rows/s against 414
    // with the device pass, 4 ranks of 8 threads each 385 against 557: the scans then compete with four uploads for the host's
    // memory bandwidth; the device pass below costs nothing on the host)
    if (resident && host_scan_threads() >= 16 && !getenv("BALER_B200_DEVICE_MINMAX")) {
      // the table fits: queue the whole upload now, find the column min / max on host threads meanwhile, and let the
      // encode + latent download of the chunks that have landed run against the rest of the upload
      // (the scan starts first, on its own threads: queuing copies from pageable memory blocks the calling thread)
      std::vector<float> mm(2 * (size_t)F);
      std::thread scan(host_colminmax, x_host, n_rows, F, mm.data(), mm.data() + F);
      ev_up.resize((size_t)n_chunks, nullptr);
      cudaError_t up_rc = cudaSuccess;
      for (int64_t k = 0; k < n_chunks && up_rc == cudaSuccess; ++k) {
        const int64_t r0 = k * chunk, rows = std::min(chunk, n_rows - r0);
        up_rc = cudaMemcpyAsync(resident + (size_t)r0 * F, x_host + (size_t)r0 * F, (size_t)rows * F * sizeof(float), cudaMemcpyHostToDevice, m->s_copy_in);
        if (up_rc == cudaSuccess) up_rc = cudaEventCreateWithFlags(&ev_up[(size_t)k], cudaEventDisableTiming);
        if (up_rc == cudaSuccess) up_rc = cudaEventRecord(ev_up[(size_t)k], m->s_copy_in);
      }
      scan.join();
      if (up_rc != cudaSuccess) {
        for (cudaEvent_t e : ev_up) if (e) cudaEventDestroy(e);
        return (int)up_rc;
      }
      BB_CUDA(cudaMemcpyAsync(fmin, mm.data(), F * sizeof(float), cudaMemcpyHostToDevice, m->s_compute));
      BB_CUDA(cudaMemcpyAsync(fmax, mm.data() + F, F * sizeof(float), cudaMemcpyHostToDevice, m->s_compute));
      range_kernel<<<(F + 127) / 128, 128, 0, m->s_compute>>>(fmin, fmax, frange, F);
      BB_CUDA(cudaMemcpyAsync(features_host, fmin, 2 * F * sizeof(float), cudaMemcpyDeviceToHost, m->s_compute));
      BB_CUDA(cudaStreamSynchronize(m->s_compute));  // (mm goes out of scope; features_host is final)
    } else {
      // a first streaming pass for the statistics on the device (column min / max kernel per chunk as it lands), which
      // keeps the table resident when it fits; otherwise the second pass below uploads it again
      for (int64_t k = 0; k < n_chunks; ++k) {
        const int64_t r0 = k * chunk, rows = std::min(chunk, n_rows - r0);
        const int s = (int)(k & 1);
        float* dst = resident ? resident + (size_t)r0 * F : (float*)m->stage_dev[0][s];
        if (!resident && k >= 2) BB_CUDA(cudaStreamWaitEvent(m->s_copy_in, m->ev_done[s], 0));
        BB_CUDA(cudaMemcpyAsync(dst, x_host + (size_t)r0 * F, (size_t)rows * F * sizeof(float), cudaMemcpyHostToDevice, m->s_copy_in));
        BB_CUDA(cudaEventRecord(m->ev_in[s], m->s_copy_in));
        BB_CUDA(cudaStreamWaitEvent(m->s_compute, m->ev_in[s], 0));
        rc = bb_colminmax_launch(m->ctx, dst, rows, F, fmin, fmax, k == 0, m->s_compute);
        if (rc != BB_OK) return rc;
        BB_CUDA(cudaEventRecord(m->ev_done[s], m->s_compute));
      }
      range_kernel<<<(F + 127) / 128, 128, 0, m->s_compute>>>(fmin, fmax, frange, F);
      BB_CUDA(cudaMemcpyAsync(features_host, fmin, 2 * F * sizeof(float), cudaMemcpyDeviceToHost, m->s_compute));
      BB_CUDA(cudaStreamSynchronize(m->s_compute));
    }
  } else if (norm) {
    BB_CUDA(cudaMemcpyAsync(fmin, features_host, 2 * F * sizeof(float), cudaMemcpyHostToDevice, m->s_compute));
  }
  struct EvGuard {  // the per-chunk upload events live for this call only
    std::vector<cudaEvent_t>& v;
    ~EvGuard() { for (cudaEvent_t e : v) if (e) cudaEventDestroy(e); }
  } ev_guard{ev_up};

  std::vector<float> widen_tmp;
  float* pinned_out[2] = {nullptr, nullptr};
  if (z_dtype == BB_F64 && n_rows) {
    if (pinned_init(m, 1, (size_t)chunk * Z * sizeof(float)) != BB_OK) return BB_ERR_NOMEM;
    pinned_out[0] = m->pinned_f32[1][0]; pinned_out[1] = m->pinned_f32[1][1];
  }
  auto finish_chunk = [&](int64_t k) -> int {  // host side of chunk k: wait for its D2H, widen if needed
    const int s = (int)(k & 1);
    BB_CUDA(cudaEventSynchronize(m->ev_out[s]));
    if (z_dtype == BB_F64) {
      const int64_t r0 = k * chunk, rows = std::min(chunk, n_rows - r0);
      widen_f32_f64(pinned_out[s], (double*)z_host + (size_t)r0 * Z, (size_t)rows * Z);
    }
    return BB_OK;
  };
  rc = BB_OK;
  for (int64_t k = 0; k < n_chunks && rc == BB_OK; ++k) {
    const int64_t r0 = k * chunk, rows = std::min(chunk, n_rows - r0);
    const int s = (int)(k & 1);
    const float* src;
    if (resident) {
      src = resident + (size_t)r0 * F;
      if (!ev_up.empty()) BB_CUDA(cudaStreamWaitEvent(m->s_compute, ev_up[(size_t)k], 0));
    } else {
      if (k >= 2) BB_CUDA(cudaStreamWaitEvent(m->s_copy_in, m->ev_done[s], 0));  // slot's previous kernel finished
      BB_CUDA(cudaMemcpyAsync(m->stage_dev[0][s], x_host + (size_t)r0 * F, (size_t)rows * F * sizeof(float), cudaMemcpyHostToDevice, m->s_copy_in));
      BB_CUDA(cudaEventRecord(m->ev_in[s], m->s_copy_in));
      BB_CUDA(cudaStreamWaitEvent(m->s_compute, m->ev_in[s], 0));
      src = (const float*)m->stage_dev[0][s];
    }
    if (k >= 2) {
      rc = finish_chunk(k - 2);  // frees output slot s on the host side (and its pinned buffer)
      if (rc != BB_OK) break;
      BB_CUDA(cudaStreamWaitEvent(m->s_compute, m->ev_out[s], 0));
    }
    rc = run_chain(m, &m->enc, src, BB_F32, rows, norm ? fmin : nullptr, norm ? frange : nullptr, nullptr, nullptr,
                   m->stage_dev[1][s], dev_dtype, precision, m->s_compute);
    if (rc != BB_OK) break;
    BB_CUDA(cudaEventRecord(m->ev_done[s], m->s_compute));
    BB_CUDA(cudaStreamWaitEvent(m->s_copy_out, m->ev_done[s], 0));
    void* dst = (z_dtype == BB_F64) ? (void*)pinned_out[s] : (void*)((char*)z_host + (size_t)r0 * Z * dtype_size(z_dtype));
    BB_CUDA(cudaMemcpyAsync(dst, m->stage_dev[1][s], (size_t)rows * Z * dtype_size(dev_dtype), cudaMemcpyDeviceToHost, m->s_copy_out));
    BB_CUDA(cudaEventRecord(m->ev_out[s], m->s_copy_out));
  }
  for (int64_t k = std::max<int64_t>(0, n_chunks - 2); k < n_chunks && rc == BB_OK; ++k) rc = finish_chunk(k);
  cudaStreamSynchronize(m->s_compute);
  cudaStreamSynchronize(m->s_copy_out);
  if (rc == BB_OK && precision == BB_PREC_AUTO && chain_has_tc(&m->enc)) {
    int tripped = 0;  // fp16 range guard of the split path: redo on the fp32 kernel (features are already known)
    rc = bb_model_range_flag(m, 1, &tripped);
    if (rc == BB_OK && tripped) return bb_compress_host(m, x_host, n_rows, features_host, 0, z_host, z_dtype, BB_PREC_FP32);
  }
  return rc;
}

int bb_decompress_host(bb_model* m, const void* z_host, int z_dtype, int64_t n_rows,
                       const float* features_host, void* y_host, int y_dtype, int precision) {
  if (!m || (!z_host && n_rows) || (!y_host && n_rows) || n_rows < 0) return BB_ERR_INVALID;
  if (z_dtype != BB_F32 && z_dtype != BB_F16 && z_dtype != BB_F64) return BB_ERR_INVALID;
  if (y_dtype != BB_F32 && y_dtype != BB_F64) return BB_ERR_INVALID;
  BB_CUDA(cudaSetDevice(m->ctx->device));
  const int F = m->dec.desc.out_dim, Z = m->dec.desc.in_dim;
  const int dev_in = (z_dtype == BB_F16) ? BB_F16 : BB_F32;
  const int64_t chunk = pipe_chunk_rows(n_rows);
  const int64_t n_chunks = n_rows ? (n_rows + chunk - 1) / chunk : 0;
  const size_t in_b = (size_t)chunk * Z * dtype_size(dev_in), out_b = (size_t)chunk * F * sizeof(float);
  int rc = pipe_init(m, in_b, out_b);
  if (rc != BB_OK) return rc;
  float* fmin = m->feat_dev;
  float* frange = m->feat_dev + F;
  const bool norm = features_host != nullptr;
  if (norm) BB_CUDA(cudaMemcpyAsync(fmin, features_host, 2 * F * sizeof(float), cudaMemcpyHostToDevice, m->s_compute));

  float* pin_in[2] = {nullptr, nullptr};
  float* pin_out[2] = {nullptr, nullptr};
  auto cleanup = [&]() {};  // the bounce buffers belong to the model
  if (n_rows) {
    if (z_dtype == BB_F64) {
      if (pinned_init(m, 0, (size_t)chunk * Z * sizeof(float)) != BB_OK) return BB_ERR_NOMEM;
      pin_in[0] = m->pinned_f32[0][0]; pin_in[1] = m->pinned_f32[0][1];
    }
    if (y_dtype == BB_F64) {
      if (pinned_init(m, 1, (size_t)chunk * F * sizeof(float)) != BB_OK) return BB_ERR_NOMEM;
      pin_out[0] = m->pinned_f32[1][0]; pin_out[1] = m->pinned_f32[1][1];
    }
  }
  auto finish_chunk = [&](int64_t k) -> int {
    const int s = (int)(k & 1);
    BB_CUDA(cudaEventSynchronize(m->ev_out[s]));
    if (y_dtype == BB_F64) {
      const int64_t r0 = k * chunk, rows = std::min(chunk, n_rows - r0);
      widen_f32_f64(pin_out[s], (double*)y_host + (size_t)r0 * F, (size_t)rows * F);
    }
    return BB_OK;
  };
  rc = BB_OK;
  for (int64_t k = 0; k < n_chunks && rc == BB_OK; ++k) {
    const int64_t r0 = k * chunk, rows = std::min(chunk, n_rows - r0);
    const int s = (int)(k & 1);
    if (k >= 2) {
      BB_CUDA(cudaEventSynchronize(m->ev_in[s]));  // pinned input slot consumed by its H2D copy
      BB_CUDA(cudaStreamWaitEvent(m->s_copy_in, m->ev_done[s], 0));
    }
    const void* src = (const char*)z_host + (size_t)r0 * Z * dtype_size(z_dtype);
    if (z_dtype == BB_F64) {
      narrow_f64_f32((const double*)src, pin_in[s], (size_t)rows * Z);
      src = pin_in[s];
    }
    BB_CUDA(cudaMemcpyAsync(m->stage_dev[0][s], src, (size_t)rows * Z * dtype_size(dev_in), cudaMemcpyHostToDevice, m->s_copy_in));
    BB_CUDA(cudaEventRecord(m->ev_in[s], m->s_copy_in));
    BB_CUDA(cudaStreamWaitEvent(m->s_compute, m->ev_in[s], 0));
    if (k >= 2) {
      rc = finish_chunk(k - 2);
      if (rc != BB_OK) break;
      BB_CUDA(cudaStreamWaitEvent(m->s_compute, m->ev_out[s], 0));
    }
    rc = run_chain(m, &m->dec, m->stage_dev[0][s], dev_in, rows, nullptr, nullptr, norm ? fmin : nullptr,
                   norm ? frange : nullptr, m->stage_dev[1][s], BB_F32, precision, m->s_compute);
    if (rc != BB_OK) break;
    BB_CUDA(cudaEventRecord(m->ev_done[s], m->s_compute));
    BB_CUDA(cudaStreamWaitEvent(m->s_copy_out, m->ev_done[s], 0));
    void* dst = (y_dtype == BB_F64) ? (void*)pin_out[s] : (void*)((float*)y_host + (size_t)r0 * F);
    BB_CUDA(cudaMemcpyAsync(dst, m->stage_dev[1][s], (size_t)rows * F * sizeof(float), cudaMemcpyDeviceToHost, m->s_copy_out));
    BB_CUDA(cudaEventRecord(m->ev_out[s], m->s_copy_out));
  }
  for (int64_t k = std::max<int64_t>(0, n_chunks - 2); k < n_chunks && rc == BB_OK; ++k) rc = finish_chunk(k);
  cudaStreamSynchronize(m->s_compute);
  cudaStreamSynchronize(m->s_copy_out);
  cleanup();
  if (rc == BB_OK && precision == BB_PREC_AUTO && chain_has_tc(&m->dec)) {
    int tripped = 0;
    rc = bb_model_range_flag(m, 1, &tripped);
    if (rc == BB_OK && tripped) return bb_decompress_host(m, z_host, z_dtype, n_rows, features_host, y_host, y_dtype, BB_PREC_FP32);
  }
  return rc;
}

}  // extern "C"
