// tcgen05 GEMM for the layer-by-layer path (bb_chain_layered.cu): Y = act(X W^T + b) per layer on the 5th-generation
// tensor cores, for chains whose weights do not fit the fused kernels (Conv_AE as its dense-equivalent chain with the
// 128 -> 2000 -> z Linears, models.py:316-407; CFD_dense_AE on 2500-feature snapshots, models.py:186-226).
//
// Arithmetic: the fp16 hi / lo split of the fused kernels (x = hi + lo; hi*hi + hi*lo + lo*hi: three kind::f16 MMAs per
// k-step into one fp32 TMEM accumulator; weights scaled by a power of two per layer so that their lo parts stay normal,
// undone in the epilogue).
// Layout: both operands are read by the tensor core straight from shared memory (no ldmatrix: the mma.sync version of
// this GEMM is bound by shared-memory bandwidth), in the canonical no-swizzle K-major core-matrix layout.  They are kept
// in that layout in global memory, so a stage is four plain bulk copies:
//   activations  [128-row tile][k / 8][row 128][8 halves]      hi array, lo array      (written by the producing layer)
//   weights      [column tile][k / 8][column NTW][8 halves]    hi image, lo image      (packed once per model)
// Kernel: persistent, one CTA per SM, warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2..17 = epilogue (TMEM lane
// quarter = warp % 4, column group = (warp - 2) / 4: 16-column chunks g, g + 4, ...).  Two 96 KB stages of 64 k columns, two accumulators of up to 256
// TMEM columns (the epilogue of one tile runs under the main loop of the next).  The epilogue adds the bias, applies the
// activation and writes the next layer's operand directly in the layout above (16-byte pieces, 512 contiguous bytes per
// warp), or fp32 rows for the last layer.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bb_common.cuh"

namespace {

constexpr int GM = 128, GK = 64, GSTAGES = 4, G_EPI_GROUPS = 4, GTHREADS = (2 + 4 * G_EPI_GROUPS) * 32;  // GSTAGES: most stages a layer may use (as many as fit)
constexpr int G_SMEM_MAX = 200 * 1024;
constexpr int G_SCR_LD = 12;                                   // floats per row of an epilogue warp's transpose patch
constexpr int G_SCR_BYTES = 16 * 32 * G_SCR_LD * 4;            // ... of all 16 epilogue warps (last layer only)
constexpr int G_A_BYTES = GM * GK * 2;  // one image (hi or lo) of an A stage
// Activations travel multiplied by 2^8: the lo half of a small activation (|x| < 0.1: lo < 6e-5) would be a subnormal fp16
// with an absolute step of 6e-8, i.e. 1e-5 relative at |x| = 0.01 (measured: latent error 1.3e-5 without the scaling).
// Values beyond 65504 / 256 raise the range flag (BB_PREC_AUTO callers re-run in fp32).
constexpr float G_ACT_SCALE = 256.f;

__device__ __forceinline__ uint32_t g_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void g_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void g_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void g_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool g_mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void g_mbar_wait(uint32_t bar, uint32_t parity) {
  // a protocol bug must not hang the GPU: trap (-> launch error) after ~1 s of polling
  for (uint32_t spins = 0; !g_mbar_try(bar, parity); ++spins) {
    if (spins > 64) __nanosleep(64);
    if (spins > (1u << 23)) __trap();
  }
}
__device__ __forceinline__ bool g_elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void g_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void g_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void g_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr) : "memory");
}
// shared-memory matrix descriptor, no swizzle, K-major: core matrices of 8 rows x 16 bytes; LBO = bytes between core
// matrices adjacent in K, SBO = bytes between core matrices adjacent in M / N (128: rows are contiguous), version 1
__device__ __forceinline__ uint64_t g_desc(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t g_idesc(int n) {
  // c_format F32 (1) @4, a / b format F16 (0) @7 / @10, a / b K-major, N >> 3 @17, M >> 4 @24
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(GM >> 4) << 24);
}
__device__ __forceinline__ void g_mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void g_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// two fp32 values -> packed fp16 hi (top 11 significant bits, truncated: x - hi is exact) and lo = fp16_rn(x - hi)
__device__ __forceinline__ void g_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u), bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(ah, bh);
  const __half2 l = __floats2half2_rn(a - ah, b - bh);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct G5Layer {
  const __half* a_hi;   // packed activations in: [m_tiles][kp / 8][128][8]; lo array follows at a_lo
  const __half* a_lo;
  const __half* w_hi;   // packed weights [n_tiles][kp / 8][ntw][8]; lo image at w_lo
  const __half* w_lo;
  const float* bias;    // [n_tiles * ntw], zero padded
  __half* y_hi;         // packed output (next layer's operand): [m_tiles][kp_next / 8][128][8]; nullptr for the last layer
  __half* y_lo;
  float* y_f32;         // last layer: fp32 rows of pitch ldy
  int ldy;
  int kp, ntw, n_tiles, m_tiles, n, kp_next;
  int n_split, n_bufs;  // accumulators per tile (2: hi * hi | cross products), accumulator sets in flight
  int n_stages;         // operand stages in shared memory
  int rows;             // valid rows of the chunk
  int act;
  float unscale;        // 2^-sw of the layer's weight scaling x 1 / G_ACT_SCALE of its input
  int* flag;
};

__global__ void __launch_bounds__(GTHREADS, 1) gemm_tc5_kernel(const __grid_constant__ G5Layer L) {
  extern __shared__ __align__(128) uint8_t gsm[];
  __shared__ __align__(8) uint64_t bars[4 * GSTAGES];  // full[2] | empty[2] | acc_full[2] | acc_empty[2]
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int lane = tid & 31;
  const uint32_t sbase = g_smem_u32(gsm);
  const uint32_t bar0 = g_smem_u32(&bars[0]);
  const uint32_t b_bytes = (uint32_t)L.ntw * GK * 2;               // one image of a B stage
  const uint32_t stage_bytes = 2 * G_A_BYTES + 2 * b_bytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (GSTAGES + s); };
  auto bar_accf = [&](int a) { return bar0 + 8u * (2 * GSTAGES + a); };
  auto bar_acce = [&](int a) { return bar0 + 8u * (3 * GSTAGES + a); };
  if (tid == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      g_mbar_init(bar_full(s), 1);
      g_mbar_init(bar_empty(s), 1);
      g_mbar_init(bar_accf(s), 1);
      g_mbar_init(bar_acce(s), 4 * G_EPI_GROUPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(g_smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  g_fence_before();
  __syncthreads();
  g_fence_after();
  const uint32_t tmem0 = tmem_base_s;
  const int n_items = L.m_tiles * L.n_tiles;
  const int n_kc = L.kp / GK;

  if (warp == 0) {
    // ---- producer: four bulk copies per stage
    if (g_elect_one()) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int mt = item / L.n_tiles, nt = item - mt * L.n_tiles;
        const size_t a_tile = (size_t)mt * (L.kp / 8) * GM * 8, w_tile = (size_t)nt * (L.kp / 8) * L.ntw * 8;
        for (int kc = 0; kc < n_kc; ++kc, ++it) {
          const int s = it % L.n_stages;
          g_mbar_wait(bar_empty(s), ((it / L.n_stages) & 1u) ^ 1u);
          g_mbar_expect_tx(bar_full(s), stage_bytes);
          const uint32_t dst = sbase + s * stage_bytes;
          const size_t a_off = a_tile + (size_t)kc * (GK / 8) * GM * 8, w_off = w_tile + (size_t)kc * (GK / 8) * L.ntw * 8;
          g_bulk_g2s(dst, L.a_hi + a_off, G_A_BYTES, bar_full(s));
          g_bulk_g2s(dst + G_A_BYTES, L.a_lo + a_off, G_A_BYTES, bar_full(s));
          g_bulk_g2s(dst + 2 * G_A_BYTES, L.w_hi + w_off, b_bytes, bar_full(s));
          g_bulk_g2s(dst + 2 * G_A_BYTES + b_bytes, L.w_lo + w_off, b_bytes, bar_full(s));
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer
    const uint32_t idesc = g_idesc(L.ntw);
    const uint32_t lbo_a = GM * 16, lbo_b = (uint32_t)L.ntw * 16;
    uint32_t it = 0, tile_i = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tile_i) {
      const int a = tile_i % L.n_bufs;
      g_mbar_wait(bar_acce(a), ((tile_i / L.n_bufs) & 1u) ^ 1u);  // the epilogue has drained this accumulator set
      g_fence_after();
      const uint32_t d0 = tmem0 + (uint32_t)(a * L.n_split * L.ntw);
      for (int kc = 0; kc < n_kc; ++kc, ++it) {
        const int s = it % L.n_stages;
        g_mbar_wait(bar_full(s), (it / L.n_stages) & 1u);
        g_fence_after();
        if (g_elect_one()) {
          const uint32_t st = sbase + s * stage_bytes;
          // Long contractions keep the hi * hi products and the two cross products in separate accumulators: the tensor
          // core truncates when it adds to the fp32 accumulator, a bias that grows with the number of additions
          // (measured 1.2e-5 on the latent for K = 2000 with all 375 additions in one accumulator); the cross products
          // are 2^-11 of the sum, so their additions (two thirds of all) need not touch it.  The epilogue adds the two
          // with a rounded fp32 add.
          const uint32_t d_hh = d0, d_cr = d0 + (uint32_t)((L.n_split - 1) * L.ntw);
#pragma unroll
          for (int ks = 0; ks < GK / 16; ++ks) {
            const uint64_t ah = g_desc(st + ks * 2 * lbo_a, lbo_a), al = g_desc(st + G_A_BYTES + ks * 2 * lbo_a, lbo_a);
            const uint64_t bh = g_desc(st + 2 * G_A_BYTES + ks * 2 * lbo_b, lbo_b);
            const uint64_t bl = g_desc(st + 2 * G_A_BYTES + b_bytes + ks * 2 * lbo_b, lbo_b);
            const uint32_t first = (kc > 0 || ks > 0) ? 1u : 0u;
            g_mma(d_hh, ah, bh, idesc, first);
            g_mma(d_cr, ah, bl, idesc, L.n_split > 1 ? first : 1u);
            g_mma(d_cr, al, bh, idesc, 1u);
          }
          g_commit(bar_empty(s));                      // the stage is free once these MMAs have read it
          if (kc == n_kc - 1) g_commit(bar_accf(a));   // ... and the accumulators are complete
        }
        __syncwarp();
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> bias, activation -> next operand (packed hi | lo) or fp32 rows
    const int q = warp & 3, h = (warp - 2) >> 2;        // TMEM lane quarter, column group (16-column chunks h, h + groups, ...)
    const int row = q * 32 + lane;
    uint32_t tile_i = 0;
    bool bad = false;
    const float slope = L.act == BB_ACT_LEAKY ? BB_LEAKY : (L.act == BB_ACT_RELU ? 0.f : 1.f);
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++tile_i) {
      const int mt = item / L.n_tiles, nt = item - mt * L.n_tiles;
      const int a = tile_i % L.n_bufs;
      g_mbar_wait(bar_accf(a), (tile_i / L.n_bufs) & 1u);
      g_fence_after();
      const uint32_t tbase = tmem0 + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * L.n_split * L.ntw);
      const int n_part = L.n_split;  // hi * hi | cross products, or everything in one
      // columns of this tile that exist (whole 16-column chunks): the last tile can be wider than what is left of the layer
      const int c_end = min(L.ntw, (L.n - nt * L.ntw + 15) & ~15);
      // the TMEM load and the bias of the next 16 columns are in flight while the current ones are converted and stored
      // (the bias used to be loaded where it is added: a third of the kernel's stall samples sat on that first FFMA)
      uint32_t vn[2][16];
      float4 bn[4];
      const float4* bias4 = reinterpret_cast<const float4*>(L.bias + nt * L.ntw);  // 16-byte aligned: ntw is a multiple of 32
      if (h * 16 < c_end) {
        g_tmem_ld16(tbase + (uint32_t)(h * 16), vn[0]);
        if (n_part > 1) g_tmem_ld16(tbase + (uint32_t)(L.ntw + h * 16), vn[1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) bn[i] = __ldg(bias4 + h * 4 + i);
      }
      bool released = false;
      for (int c0 = h * 16; c0 < c_end; c0 += 16 * G_EPI_GROUPS) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(vn[0][j]) + (n_part > 1 ? __uint_as_float(vn[1][j]) : 0.f);
        const bool more = c0 + 16 * G_EPI_GROUPS < c_end;
        if (more) {
          g_tmem_ld16(tbase + (uint32_t)(c0 + 16 * G_EPI_GROUPS), vn[0]);
          if (n_part > 1) g_tmem_ld16(tbase + (uint32_t)(L.ntw + c0 + 16 * G_EPI_GROUPS), vn[1]);
        } else {
          // this warp has read its last columns of the accumulator set: the MMA warp may refill it while they are stored
          g_fence_before();
          __syncwarp();
          if (lane == 0) g_mbar_arrive(bar_acce(a));
          released = true;
        }
        const int col = nt * L.ntw + c0;  // global output column of acc[0]
        float o[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          o[4 * i + 0] = fmaf(acc[4 * i + 0], L.unscale, bn[i].x);
          o[4 * i + 1] = fmaf(acc[4 * i + 1], L.unscale, bn[i].y);
          o[4 * i + 2] = fmaf(acc[4 * i + 2], L.unscale, bn[i].z);
          o[4 * i + 3] = fmaf(acc[4 * i + 3], L.unscale, bn[i].w);
        }
        if (more) {
#pragma unroll
          for (int i = 0; i < 4; ++i) bn[i] = __ldg(bias4 + (c0 + 16 * G_EPI_GROUPS) / 4 + i);
        }
        // activation, branch-free: slope 1 (none), 0 (ReLU), BB_LEAKY - max(v, 0) + slope * min(v, 0) is v, relu(v), leaky(v)
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = fmaf(slope, fminf(o[j], 0.f), fmaxf(o[j], 0.f));
        if (L.y_hi != nullptr) {
#pragma unroll
          for (int k8 = 0; k8 < 2; ++k8) {
            if (col + 8 * k8 < L.kp_next) {
              uint4 hi, lo;
              g_split2(o[8 * k8 + 0], o[8 * k8 + 1], hi.x, lo.x);
              g_split2(o[8 * k8 + 2], o[8 * k8 + 3], hi.y, lo.y);
              g_split2(o[8 * k8 + 4], o[8 * k8 + 5], hi.z, lo.z);
              g_split2(o[8 * k8 + 6], o[8 * k8 + 7], hi.w, lo.w);
#pragma unroll
              for (int j = 0; j < 8; ++j) bad = bad || !(fabsf(o[8 * k8 + j]) <= 65504.f);
              const size_t off = (((size_t)mt * (L.kp_next / 8) + (col >> 3) + k8) * GM + row) * 8;
              *reinterpret_cast<uint4*>(L.y_hi + off) = hi;
              *reinterpret_cast<uint4*>(L.y_lo + off) = lo;
            }
          }
        } else {
          // fp32 rows: a thread holds 16 consecutive values of ONE row, so storing them directly makes every store
          // instruction touch 32 rows (4 useful bytes per sector: 145 us for the 2000 -> 250 Linear of Conv_AE against 61
          // for the same contraction with the packed output).  Eight columns at a time go through a 32 x 8 patch of shared
          // memory (48-byte row stride) and leave as 16-byte stores, two lanes per row.
#pragma unroll
          for (int j = 0; j < 16; ++j) bad = bad || !(fabsf(o[j]) <= 3.0e38f);
          float* scr = reinterpret_cast<float*>(gsm + (size_t)L.n_stages * stage_bytes) + (warp - 2) * (32 * G_SCR_LD);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            __syncwarp();
            *reinterpret_cast<float4*>(scr + lane * G_SCR_LD) = make_float4(o[8 * half], o[8 * half + 1], o[8 * half + 2], o[8 * half + 3]);
            *reinterpret_cast<float4*>(scr + lane * G_SCR_LD + 4) =
                make_float4(o[8 * half + 4], o[8 * half + 5], o[8 * half + 6], o[8 * half + 7]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int r = 16 * i + (lane >> 1), cq = (lane & 1) * 4;
              const float4 v = *reinterpret_cast<const float4*>(scr + r * G_SCR_LD + cq);
              const int grow = mt * GM + q * 32 + r, gcol = col + 8 * half + cq;
              if (grow < L.rows && gcol < L.ldy)  // ldy is a multiple of 4: a quad is inside the row or outside
                *reinterpret_cast<float4*>(L.y_f32 + (size_t)grow * L.ldy + gcol) = v;
            }
          }
        }
      }
      // the last column tile also zero-fills the k padding of the next operand beyond its own columns
      if (L.y_hi != nullptr && nt == L.n_tiles - 1) {
        for (int k8 = (nt * L.ntw + c_end) / 8 + h; k8 < L.kp_next / 8; k8 += G_EPI_GROUPS) {
          const size_t off = (((size_t)mt * (L.kp_next / 8) + k8) * GM + row) * 8;
          *reinterpret_cast<uint4*>(L.y_hi + off) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(L.y_lo + off) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
      if (!released) {  // (a warp without columns in this tile)
        g_fence_before();
        __syncwarp();
        if (lane == 0) g_mbar_arrive(bar_acce(a));
      }
    }
    if (bad && L.flag != nullptr) *L.flag = 1;
  }
  g_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem0), "r"(512) : "memory");
}

// in (f32 | f16, compact rows of `dim`) -> packed split operand [m_tiles][kp / 8][128][8] (hi, lo), optionally (x - min) / range
__global__ void __launch_bounds__(256) g5_stage_in_kernel(const void* __restrict__ in, const int in_dtype, const int64_t rows,
                                                          const int dim, const int kp, const int m_tiles,
                                                          const float* __restrict__ mn, const float* __restrict__ rg,
                                                          __half* __restrict__ hi, __half* __restrict__ lo, int* __restrict__ flag) {
  // One thread per 16-byte unit (8 k values of one row).  A warp covers 4 rows x 8 units: it reads four 256-byte runs of
  // the row-major input and writes eight 64-byte runs of each packed image (with one row per lane every load touched 32
  // rows).  32-bit index arithmetic: a chunk is at most a few million units (checked by the launcher).
  const uint32_t kq = (uint32_t)(kp / 8);  // units per row, a multiple of 8
  const uint32_t total = (uint32_t)m_tiles * kq * GM, G = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += G) {
    const uint32_t w = i >> 5, ln = i & 31u;              // warp-sized group, lane in it
    const uint32_t rg4 = w % (GM / 4), t = w / (GM / 4);  // 4-row group of the tile
    const uint32_t kg = t % (kq / 8), mt = t / (kq / 8);  // 8-unit group of the row, row tile
    const int r = (int)(rg4 * 4 + (ln & 3u));
    const int k8 = (int)(kg * 8 + (ln >> 2));
    const int64_t row = (int64_t)mt * GM + r;
    const uint32_t e = (mt * kq + (uint32_t)k8) * GM + (uint32_t)r;  // unit index in the packed images
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = k8 * 8 + j;
      float x = 0.f;
      if (row < rows && c < dim) {
        x = in_dtype == BB_F16 ? __half2float(reinterpret_cast<const __half*>(in)[row * dim + c]) : reinterpret_cast<const float*>(in)[row * dim + c];
        if (mn != nullptr) x = __fdiv_rn(__fsub_rn(x, __ldg(mn + c)), __ldg(rg + c));
        x *= G_ACT_SCALE;
        if (!(fabsf(x) <= 65504.f) && flag != nullptr) *flag = 1;
      }
      v[j] = x;
    }
    uint4 h, l;
    g_split2(v[0], v[1], h.x, l.x);
    g_split2(v[2], v[3], h.y, l.y);
    g_split2(v[4], v[5], h.z, l.z);
    g_split2(v[6], v[7], h.w, l.w);
    reinterpret_cast<uint4*>(hi)[e] = h;
    reinterpret_cast<uint4*>(lo)[e] = l;
  }
}

inline int g5_round(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

// ---- host side: per-layer geometry, weight packing, launch
struct G5Geom { int K, N, kp, ntw, n_tiles, n_split, n_bufs, n_stages; };
static G5Geom g5_geom(int K, int N) {
  G5Geom g;
  g.K = K; g.N = N;
  g.kp = g5_round(K, GK);
  g.n_tiles = (N + 255) / 256;
  g.ntw = g5_round((N + g.n_tiles - 1) / g.n_tiles, 32);
  g.n_split = g.kp >= 512 ? 2 : 1;  // long contractions: hi * hi and the cross products in separate accumulators
  g.n_bufs = 2 * g.n_split * g.ntw <= 512 ? 2 : 1;
  g.n_stages = std::max(2, std::min(GSTAGES, G_SMEM_MAX / (2 * G_A_BYTES + 2 * g.ntw * GK * 2)));
  return g;
}

int bb_gemm_tc5_prepare(bb_ctx*, Chain* c) {
  const ChainDesc& d = c->desc;
  c->g5_ok = false;
  size_t halves = 0, nb = 0;
  for (int l = 0; l < d.n_layers; ++l) {
    const G5Geom g = g5_geom(d.layer[l].K, d.layer[l].N);
    if (g.ntw > 256 || g.ntw % 32) return BB_ERR_UNSUPPORTED;
    // a model with a non-finite parameter is left to the mma.sync GEMM, whose range guard sees the nan / inf and sends
    // BB_PREC_AUTO callers to the fp32 kernels, where it propagates as it does upstream (the min / max form of the
    // activation in this kernel's epilogue would turn a nan into 0)
    for (size_t i = 0; i < (size_t)d.layer[l].N * d.layer[l].K; ++i)
      if (!std::isfinite(c->w_host[l][i])) return BB_ERR_UNSUPPORTED;
    for (int n = 0; n < d.layer[l].N; ++n)
      if (!std::isfinite(c->b_host[l][n])) return BB_ERR_UNSUPPORTED;
    c->g5_w_off[l] = halves;
    halves += 2 * (size_t)g.n_tiles * g.kp * g.ntw;
    c->g5_b_off[l] = nb;
    nb += (size_t)g.n_tiles * g.ntw;
  }
  std::vector<__half> img(halves, __float2half(0.f));
  std::vector<float> bias(nb, 0.f);
  for (int l = 0; l < d.n_layers; ++l) {
    const int K = d.layer[l].K, N = d.layer[l].N;
    const G5Geom g = g5_geom(K, N);
    double mx = 0.0;
    for (size_t i = 0; i < (size_t)N * K; ++i) mx = std::max(mx, std::fabs(c->w_host[l][i]));
    int sw = 0;  // scale so that max |w| lands in [512, 1024): lo = w - hi of all but the tiniest weights stays a normal fp16
    if (mx > 0.0) sw = (int)std::floor(std::log2(1024.0 / mx) - 1e-9);
    sw = std::max(-14, std::min(24, sw));
    // the accumulator carries 2^sw (weights) x G_ACT_SCALE (input operand); all but the last layer's output is the next
    // operand and leaves scaled by G_ACT_SCALE again: folded into this factor and into the bias (powers of two, and the
    // activations are positively homogeneous, so the values are the ones a separate multiplication would give)
    const double out_scale = l + 1 < d.n_layers ? (double)G_ACT_SCALE : 1.0;
    c->g5_unscale[l] = (float)(std::ldexp(1.0, -sw) / (double)G_ACT_SCALE * out_scale);
    __half* hi = img.data() + c->g5_w_off[l];
    __half* lo = hi + (size_t)g.n_tiles * g.kp * g.ntw;
    for (int n = 0; n < N; ++n) {
      const int nt = n / g.ntw, nn = n - nt * g.ntw;
      for (int k = 0; k < K; ++k) {
        const float w = (float)std::ldexp(c->w_host[l][(size_t)n * K + k], sw);
        uint32_t bits;
        memcpy(&bits, &w, 4);
        bits &= 0xFFFFE000u;
        float wh;
        memcpy(&wh, &bits, 4);
        const size_t idx = (((size_t)nt * (g.kp / 8) + k / 8) * g.ntw + nn) * 8 + (k & 7);
        hi[idx] = __float2half_rn(wh);
        lo[idx] = __float2half_rn(w - wh);
      }
      bias[c->g5_b_off[l] + n] = (float)(c->b_host[l][n] * out_scale);
    }
  }
  if (c->g5_blob_dev) cudaFree(c->g5_blob_dev);
  if (c->g5_bias_dev) cudaFree(c->g5_bias_dev);
  c->g5_blob_dev = nullptr; c->g5_bias_dev = nullptr;
  BB_CUDA(cudaMalloc(&c->g5_blob_dev, img.size() * sizeof(__half)));
  BB_CUDA(cudaMemcpy(c->g5_blob_dev, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  BB_CUDA(cudaMalloc(&c->g5_bias_dev, bias.size() * sizeof(float)));
  BB_CUDA(cudaMemcpy(c->g5_bias_dev, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
  const int smem = G_SMEM_MAX + G_SCR_BYTES;
  BB_CUDA(cudaFuncSetAttribute(gemm_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  c->g5_ok = true;
  return BB_OK;
}

// bytes of one ping-pong operand buffer for `rows` rows: the widest packed operand (hi + lo) or the fp32 output of the last layer
size_t bb_gemm_tc5_buf_bytes(const Chain* c, int64_t rows) {
  const ChainDesc& d = c->desc;
  const int64_t m_tiles = (rows + GM - 1) / GM;
  size_t best = (size_t)m_tiles * GM * g5_round(d.in_dim, GK) * 4;
  for (int l = 0; l < d.n_layers; ++l) {
    const size_t packed = (size_t)m_tiles * GM * g5_round(d.layer[l].N, GK) * 4;
    best = std::max(best, packed);
  }
  return best + 256;
}

// one chunk of rows through all layers; buf0 / buf1: ping-pong buffers of bb_gemm_tc5_buf_bytes.  Returns which buffer holds
// the fp32 output rows (pitch *ld_out floats) through *out_buf.
int bb_gemm_tc5_chunk(bb_ctx* ctx, const Chain* c, const void* in, int in_dtype, int64_t rows, const float* pre_min,
                      const float* pre_range, void* buf0, void* buf1, int* flag_dev, int* out_buf, int* ld_out,
                      cudaStream_t stream) {
  if (!c->g5_ok) return BB_ERR_UNSUPPORTED;
  const ChainDesc& d = c->desc;
  const int m_tiles = (int)((rows + GM - 1) / GM);
  if ((int64_t)m_tiles * GM * (g5_round(d.in_dim, GK) / 8) >= (1ll << 31)) return BB_ERR_INVALID;  // staging kernel: 32-bit indices
  void* buf[2] = {buf0, buf1};
  int cur = 0;
  int kp = g5_round(d.in_dim, GK);
  {
    __half* hi = reinterpret_cast<__half*>(buf[0]);
    __half* lo = hi + (size_t)m_tiles * GM * kp;
    g5_stage_in_kernel<<<ctx->sm_count * 8, 256, 0, stream>>>(in, in_dtype, rows, d.in_dim, kp, m_tiles, pre_min, pre_range, hi, lo,
                                                            flag_dev);
  }
  for (int l = 0; l < d.n_layers; ++l) {
    const G5Geom g = g5_geom(d.layer[l].K, d.layer[l].N);
    const bool last = l + 1 == d.n_layers;
    G5Layer L;
    L.a_hi = reinterpret_cast<const __half*>(buf[cur]);
    L.a_lo = L.a_hi + (size_t)m_tiles * GM * g.kp;
    L.w_hi = reinterpret_cast<const __half*>(c->g5_blob_dev) + c->g5_w_off[l];
    L.w_lo = L.w_hi + (size_t)g.n_tiles * g.kp * g.ntw;
    L.bias = c->g5_bias_dev + c->g5_b_off[l];
    L.kp = g.kp; L.ntw = g.ntw; L.n_tiles = g.n_tiles; L.m_tiles = m_tiles; L.n = g.N;
    L.n_split = g.n_split; L.n_bufs = g.n_bufs; L.n_stages = g.n_stages;
    L.rows = (int)rows; L.act = d.layer[l].act; L.unscale = c->g5_unscale[l]; L.flag = flag_dev;
    if (!last) {
      L.kp_next = g5_round(g.N, GK);
      L.y_hi = reinterpret_cast<__half*>(buf[cur ^ 1]);
      L.y_lo = L.y_hi + (size_t)m_tiles * GM * L.kp_next;
      L.y_f32 = nullptr; L.ldy = 0;
    } else {
      L.kp_next = 0;
      L.y_hi = L.y_lo = nullptr;
      L.y_f32 = reinterpret_cast<float*>(buf[cur ^ 1]);
      L.ldy = (g.N + 3) & ~3;
      *ld_out = L.ldy;
    }
    const int items = m_tiles * g.n_tiles;
    const int grid = items < ctx->sm_count ? items : ctx->sm_count;
    const int smem = g.n_stages * (2 * G_A_BYTES + 2 * g.ntw * GK * 2) + (last ? G_SCR_BYTES : 0);
    gemm_tc5_kernel<<<grid, GTHREADS, smem, stream>>>(L);
    cur ^= 1;
  }
  *out_buf = cur;
  return (int)cudaGetLastError();
}
