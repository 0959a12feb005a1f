// Tensor-core training step of the dense autoencoder (AE / CFD_dense_AE family, 8 Linears).
//
// Replaces the body of the batch loop of training.fit (reference baler/modules/training.py:64-97): forward,
// sum-MSE / n_columns loss (utils.py:195-199), backward, Adam (training.py:266) - the same contract as the fp32
// kernels in bb_train.cu, which stay as the reference-accuracy path for the opt-in L1 chain and AE_Dropout_BN.
//
// Why warp-level MMA (mma.sync m16n8k16, SASS HMMA) and not tcgen05 here: a training step is 512 rows.  tcgen05 needs
// 128 (M) rows per instruction, so a step would occupy 4 SMs whose tensor pipes run the 15 dependent layer passes
// back to back (>= 7 us of MMA issue alone for the 3-product split); 16-row warp tiles put the same step on 32 SMs and
// the weight-gradient products on ~90.  mma.sync runs at 0.5 m16n8k16 / clock / SM here (tools/hmma_bench.cu).  The arithmetic is the same fp16 hi / lo split as the inference kernels:
//   x = hi + lo / 2048, hi = fp16(x), lo = fp16((x - hi) * 2048)      (22 significant bits, lo never subnormal)
//   a * b ~= hi_a hi_b + (hi_a lo_b + lo_a hi_b) / 2048                (three HMMA per k-step, fp32 accumulate)
//
// One persistent cooperative kernel (8 warps per CTA, one CTA per SM) runs a whole epoch (or one step of the step API).
// Per step:
//   phase 1  CTA i < rows / 16 owns 16 batch rows and walks the 8 forward and 7 backward layer passes.  The B operands
//            (weights, pre-packed in mma fragment order as fp16 hi | lo, a forward and a transposed image) stream
//            L2 -> shared memory by bulk copies (TMA 1-D, mbarrier completion) into a 4-stage ring that runs ahead of
//            the math; the warp that is last to finish a stage refills it.  Activations stay in shared memory as hi / lo
//            fp16 [row][feature] (ldmatrix A operands); every layer input X_l and every pre-activation gradient dZ_l is
//            also copied to a global scratch for phase 2.  A warp owns the n-tiles warp, warp + 8, ...; each layer pass
//            is instantiated per tile count (predicated-off HMMAs occupy the tensor pipe like live ones).
//   barrier  (grid-wide, one atomic counter)
//   phase 2  CTA j owns a 32 x 32 block of one layer's weight matrix (bias = one more input column of ones): it bulk-
//            loads the two [rows][32 features] panels, contracts over ALL batch rows (dW = dZ^T X, ldmatrix.trans), so no
//            cross-CTA reduction and no atomics exist anywhere; the Adam update runs in the accumulator registers and the
//            updated weights are written back as fp32 master copy + whole fragments of both packed fp16 images.
//   barrier
// Measured on B200 (tools/train_bench.py, profiles/r02_train_*): what bounds the step is not HMMA throughput but
// per-pass latency - 15 dependent passes, each a barrier, a weight-chunk wait, a k loop of dependent MMA chains at 2
// warps per scheduler, and an epilogue of dependent conversions.
// Data parallel: between the contraction and Adam each rank pushes its gradient tile into every peer's exchange
// buffer over NVLink (peer-mapped memory) as 8-byte {value, step tag} packets and sums the world's tiles in rank order as
// they land - the all-reduce is fused into the weight-gradient kernel, tile by tile, with no fence, flag or barrier on the
// path (SUM, not mean: the loss is a sum, utils.py:195).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include <curand_philox4x32_x.h>

#include "bb_train_tc.cuh"

namespace {

constexpr int NL = 8;
// 8 warps = 256 threads: registers are allocated to CTAs in units of 4 warps, so 9..12 warps would cap every thread at
// 168 registers (measured: spills in the inner loops); with 8 the accumulators, the double-buffered fragments and the
// epilogue live in registers.
constexpr int NWARPS = 8;

constexpr int NTHREADS = NWARPS * 32;
constexpr int TROWS = 16;          // batch rows per phase-1 tile (one m16 MMA tile)
constexpr int RING = 4;            // weight ring stages
constexpr int CHUNK_U4 = 2048;     // uint4 per stage = 64 (k-step, n-tile) fragment blocks of 32 lanes = 32 KB
constexpr int MAX_CHUNKS = 48;
constexpr int P2_BLK = 32;         // edge of a phase-2 weight block
constexpr int P2_ROWS = 512;       // batch rows per panel pass
constexpr int P2_PANEL = P2_ROWS / 16 * 1024;  // bytes of one panel array: 32 tiles x 32 features x 16 rows x 2
constexpr float LO_SCALE = 2048.f, LO_INV = 1.f / 2048.f;
constexpr int MAX_WORLD = 16;
constexpr int BN_PW = 208;         // row width (columns) of the per-tile BatchNorm partial-sum packets
constexpr int BN_TW = 208;         // width of the backward totals scratch in shared memory
constexpr int POLL = 8;            // BatchNorm packets a thread keeps in flight
constexpr int RING_DBN = 3;        // AE_Dropout_BN: one ring stage less, its 32 KB hold the BatchNorm inputs
constexpr float BN_EPS = 1e-5f, BN_MOMENTUM = 0.1f;  // nn.BatchNorm1d defaults (models.py:284-296)

struct TcLayer {
  int K, N, act;
  int KS, NTf;          // forward pass: k-steps over K + 1 (bias column of ones), n-tiles = the consumer's padded K / 8
  int KSb, NTb;         // backward pass (dX = dZ W): k-steps over N, n-tiles = the consumer's padded N_{l-1} / 8
  int w_off, b_off;     // flat fp32 parameter offsets (W (out, in) row-major, then b)
  int imgf_off, imgb_off;  // uint4 offsets of the packed images [k-step][n-tile][lane] = (hi.r0, hi.r1, lo.r0, lo.r1)
  int x_off, ldx;       // shared memory: halves offset and row stride of X_l (layer input, incl. ones + zero pad)
  int xt_off, zt_off;   // global feature-major scratch: first feature row of X_l / dZ_l
};

// One weight chunk = a run of k-steps of one layer pass that fits a ring stage.  Self-contained (no second lookup behind
// it: at 2 warps per scheduler every dependent shared-memory load at a chunk boundary is exposed latency) and 32 bytes,
// so a warp fetches the NEXT chunk's descriptor with two LDS.128 while it works on the current one.
struct TcChunk {
  int src_u4, n_u4;        // producer: source offset / size in the packed images
  short pass, nk;          // layer pass 0..14, k-steps in this chunk
  short NT, flags;         // n-tiles of the pass; flags: 1 first chunk of the pass, 2 last, 4 backward pass
  int a_hi_off, a_lo_off;  // byte offsets in dynamic shared memory of the A operand (hi / lo) at this chunk's first k-step
  int lda_b;               // A row stride, bytes
  int l;                   // layer
};
// What the epilogue of a layer pass needs, byte offsets in dynamic shared memory
struct TcPass {
  int N, act;              // out features / activation of the layer (forward passes)
  int o_hi_off, o_lo_off;  // where the pass writes its output (X_{l+1}, the seed dZ_7, or dZ_{l-1})
  int o_ld_b;
  int pact;                // backward: activation of layer l - 1 ...
  int x_hi_off, x_lo_off;  // ... and X_l, whose sign is the sign of that layer's pre-activation
  int x_ld_b, kind;        // kind: 0 forward (hidden), 1 forward (last layer: loss + seed), 2 backward
  int c0, c1;              // chunks [c0, c1) of the weight stream belong to this pass
  // A operand of the pass (what phase 2 needs feature-major): shared-memory offsets, row stride, padded width, and
  // where it goes in the global scratch (first feature row; 0 = X, 1 = dZ)
  int a_hi_off, a_lo_off, a_ld_b, a_feat, a_gfeat0, a_which, NT;
  // AE_Dropout_BN (models.py:256-313): dropout layer handled in this epilogue (forward pass of encoder layer `drop`; backward
  // pass that produces dZ of encoder layer `drop`) and BatchNorm (forward pass of decoder layer 4 + bn; backward pass that
  // produces dZ of decoder layer 4 + bn); -1 = none
  int drop, bn, pad_[3];
};

struct TcModel {
  TcLayer L[NL];
  int n_chunks, n_chunks_fwd, n_items, n_params, F, RS, dz_ld, smem_x_halves, max_tiles;
  int kind, n_linear;            // 0 AE / 1 AE_Dropout_BN; Linear parameters in the flat vector (BatchNorm gamma / beta follow)
  // AE_Dropout_BN: flat offsets of gamma / beta, feature offset in the concatenated running statistics, width, the fp32
  // shared-memory copy U_i of every BatchNorm input (byte offset, row stride in floats), and the per-feature scratch
  // [mean bn_f_total | 1/sqrt(var + eps) bn_f_total | 4 transient arrays of BN_TW] and the slice partials of a reduction
  // point (byte offsets)
  int bn_g_off[4], bn_b_off[4], bn_f_off[4], bn_n[4], bn_slices[4], u_off[4], u_ld[4], stat_off, red_off, bn_f_total;
  unsigned keep_thr[4];          // dropout: keep when the 32-bit draw is below this (1 - p of models.py:263-275) ...
  float keep_scale[4];           // ... and scale the kept value by 1 / (1 - p)
  alignas(16) TcChunk chunk[MAX_CHUNKS];
  TcPass pass[2 * NL - 1];
};

struct TcItem { int l, n0, k0; };

struct TcPtrs {
  float *params, *m, *v, *grads;
  const uint4* img;
  __half *xt_hi, *xt_lo, *zt_hi, *zt_lo;
  float* loss_part;       // [2][max_tiles]
  const TcItem* items;
  unsigned* bar;
  const float2* stephyper;  // per step of this launch: (lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t))
  int* flag;
  // data parallel (world > 1): peer-mapped exchange buffers and flags, indexed by rank
  int rank, world;
  float* xchg[MAX_WORLD];           // every rank's exchange block: [2][world][n_items][256][4] {value, tag} packets
  unsigned* xflag[MAX_WORLD];       // (reserved)
  unsigned xbase;                   // packet tags are xbase + step-in-launch + 1 (monotonic over launches, never 0)
  int dp_slice;                     // 1: x is the full table, every step takes this rank's share of a global batch
  // AE_Dropout_BN
  int train;                        // 1: dropout on, batch statistics; 0: eval (running statistics)
  float *rm, *rv;                   // running mean / variance, bn_f_total each
  long long* nbt;                   // num_batches_tracked[4]
  uint4* bn_part;                   // [8 reduction points][max_tiles][BN_PW] packets {a, tag, b, tag} of per-tile column sums
  uint4* bn_fin;                    // [8 reduction points][BN_PW] final packets of the columns (statistics / totals of the batch)
  uint4* bnx[MAX_WORLD];            // data parallel: every rank's [2 parities][8 points][world][BN_PW] packets of per-rank sums
  const unsigned char* mask[4];     // injected dropout keep-masks [global batch][width] (parity tests) or nullptr (Philox)
  unsigned long long seed;
  unsigned long long drop_step;     // dropout stream position of the launch's first step
  long long* prof;                  // diagnostics (nullable): clock64 stamps of CTA 0 during step `prof_step` of a launch
  int prof_step;
};

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// mbarrier + bulk copy (TMA 1-D, SASS UBLKCP): one thread moves a whole stage, nobody's LSU slots are spent on it
__device__ __forceinline__ void mbar_init(uint64_t* bar, const uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, const uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, const uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t a = smem_u32(bar);
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, const uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes (other SMs' st.global before the grid barrier, this CTA's shared-memory accesses) -> async-proxy copies
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 32-bit shared-window addresses: generic pointers cost an S2R (shared window base) plus 64-bit arithmetic per access in
// the inner loops (84 instructions per k-step measured, the warps are issue / latency bound at 2 warps per scheduler)
__device__ __forceinline__ void ldsm_x4a(uint32_t (&r)[4], const uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ldsm_x4t(uint32_t (&r)[4], const uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ldsm_x2t(uint32_t& r0, uint32_t& r1, const uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a) : "memory");
}
__device__ __forceinline__ uint4 lds128(const uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(const uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(const uint32_t a, const uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float2 half2_bits_to_float2(const uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); }

// D += A(16x16, row) * B(16x8, col), fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// The split product of one k-step is three HMMAs: hi*hi, hi*lo, lo*hi.  The tensor core adds its 16 products to the
// accumulator with truncation, which over a long contraction biases the hi*hi chain by ~1e-6 (measured: gradient error
// 1.8e-6 -> 3.4e-7 of max); that product therefore starts from zero and is added to the running sum with a rounded fp32
// add.  The two cross products are 2^-11 smaller and accumulate on the tensor core.
__device__ __forceinline__ void split16(const float x, __half& hi, __half& lo) {
  hi = __float2half_rn(x);
  lo = __float2half_rn((x - __half2float(hi)) * LO_SCALE);
}
// (a, b) -> packed fp16 hi pair and packed scaled lo pair
__device__ __forceinline__ void split16x2(const float a, const float b, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((a - hf.x) * LO_SCALE, (b - hf.y) * LO_SCALE);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
// Warp-uniform values the compiler cannot prove uniform (the warp index, anything loaded from memory) make every branch
// on them "potentially divergent": nvcc then guards each mma.sync / ldmatrix behind a WARPSYNC.ALL, ~12 per k-step in the
// inner loops (measured: ~300 cycles per k-step).  A shuffle from lane 0 is uniform by construction.
__device__ __forceinline__ int uni(const int v) { return __shfl_sync(0xffffffffu, v, 0); }

// A spin that never ends must not hang the GPU (a peer rank that died, a protocol bug): trap - the launch fails with an
// error instead - after 2^28 polls (a poll is an L2 round trip: about a minute; a data-parallel peer may legitimately be
// seconds late into an epoch while rank 0 writes files).
__device__ __forceinline__ void spin_guard(unsigned& spins) {
  if (++spins > (1u << 28)) __trap();
}

// every CTA of the (cooperatively launched, co-resident) grid arrives; `target` counts arrivals since the launch
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    unsigned spins = 0;
    while (*reinterpret_cast<volatile unsigned*>(bar) < target) spin_guard(spins);
    __threadfence();
  }
  __syncthreads();
  fence_proxy_async();
}

__device__ __forceinline__ float act_apply(const float v, const int act) {
  if (act == BB_ACT_LEAKY) return v > 0.f ? v : BB_LEAKY * v;
  if (act == BB_ACT_RELU) return fmaxf(v, 0.f);
  return v;
}

// Scratch shared by the two phases.  Per array (hi or lo) of X or dZ:
//   [feature block of 32][16-row tile s][row 16][4 pieces of 8 features = 16 bytes, piece index XOR (row >> 1) & 3]
// so that (a) a phase-1 CTA copies its shared-memory rows out in 16-byte pieces, 1 KB contiguous per (block, tile), (b)
// the panel of a phase-2 item (32 features, all tiles) is ONE contiguous range = one bulk copy, (c) ldmatrix.trans on the
// panel as it lands in shared memory (rows = batch rows = the contraction index of dW = dZ^T X) is bank-conflict free.
__device__ __forceinline__ size_t tm_byte_offset(const int fb, const int tile, const int s_max, const int row, const int piece) {
  return (((size_t)fb * s_max + tile) * 16 + row) * 64 + (size_t)((piece ^ ((row >> 1) & 3)) * 16);
}

// shared [16 rows][ld] hi / lo arrays -> scratch, all threads: one LDS.128 + one STG.128 per (block, row, piece)
__device__ __forceinline__ void write_tile_major(const uint32_t sh_hi, const uint32_t sh_lo, const int ld_b, const int n_feat,
                                                 __half* g_hi, __half* g_lo, const int feat0, const int s_max, const int tile) {
  const int n_fb = (n_feat + 31) >> 5, per = n_fb * 64;
  for (int task = threadIdx.x; task < 2 * per; task += NTHREADS) {
    const int which = task >= per;
    const int rem = task - which * per, fbi = rem >> 6, row = (rem >> 2) & 15, piece = rem & 3;
    const uint4 v = lds128((which ? sh_lo : sh_hi) + row * ld_b + fbi * 64 + piece * 16);
    unsigned char* dst = reinterpret_cast<unsigned char*>(which ? g_lo : g_hi);
    *reinterpret_cast<uint4*>(dst + tm_byte_offset((feat0 >> 5) + fbi, tile, s_max, row, piece)) = v;
  }
}

// Refresh of one parameter's entries in the packed images (flat kernel only; phase 2 writes whole fragment blocks).
// (n, k) of layer l; k == K is the bias.
__device__ __forceinline__ void store_packed(const TcModel& M, const TcPtrs& P, const int l, const int n, const int k, const float p) {
  const TcLayer& L = M.L[l];
  __half hi, lo;
  split16(p, hi, lo);
  __half* img = reinterpret_cast<__half*>(const_cast<uint4*>(P.img));
  {  // forward image: contraction over k, B[k][n] = W[n][k]
    const int ks = k >> 4, nt = n >> 3, lane = ((n & 7) << 2) | ((k & 7) >> 1), reg = (k & 15) >> 3, h = k & 1;
    __half* q = img + ((size_t)L.imgf_off + ((size_t)ks * L.NTf + nt) * 32 + lane) * 8;
    q[reg * 2 + h] = hi;
    q[4 + reg * 2 + h] = lo;
  }
  if (l >= 1 && k < L.K) {  // transposed image: contraction over n, B[n][k] = W[n][k]
    const int ks = n >> 4, nt = k >> 3, lane = ((k & 7) << 2) | ((n & 7) >> 1), reg = (n & 15) >> 3, h = n & 1;
    __half* q = img + ((size_t)L.imgb_off + ((size_t)ks * L.NTb + nt) * 32 + lane) * 8;
    q[reg * 2 + h] = hi;
    q[4 + reg * 2 + h] = lo;
  }
}

// Adam: torch.optim.Adam single-tensor arithmetic (training.py:266), the same expressions as train_adam_kernel in bb_train.cu
__device__ __forceinline__ void adam_math(const float m0, const float v0, const float p0, const float g, const float lr_bc1,
                                          const float inv_sqrt_bc2, const float beta1, const float beta2, const float eps,
                                          float& mm, float& vv, float& p_out) {
  // explicit roundings: the tile epilogue and the flat kernel must produce the same bits (no compiler-chosen contraction)
  mm = __fmaf_rn(__fsub_rn(g, m0), 1.f - beta1, m0);                         // exp_avg.lerp_(g, 1 - beta1)
  vv = __fmaf_rn(beta2, v0, __fmul_rn(__fmul_rn(1.f - beta2, g), g));        // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
  const float denom = __fmaf_rn(__fsqrt_rn(vv), inv_sqrt_bc2, eps);
  p_out = __fmaf_rn(-lr_bc1, __fdiv_rn(mm, denom), p0);
}
__device__ __forceinline__ void adam_one(const TcPtrs& P, const int idx, const float g, const float lr_bc1, const float inv_sqrt_bc2,
                                         const float beta1, const float beta2, const float eps, float& p_out) {
  float mm, vv;
  adam_math(P.m[idx], P.v[idx], P.params[idx], g, lr_bc1, inv_sqrt_bc2, beta1, beta2, eps, mm, vv, p_out);
  P.m[idx] = mm;
  P.v[idx] = vv;
  P.params[idx] = p_out;
}

// shared memory: [phase-1 ring + activations | phase-2 panels + tile] ... [mbarriers]
struct SmemBars {
  uint64_t full[RING];   // weight ring stage filled (bulk copy complete)
  uint64_t grp[4];       // phase-2 panel row groups
  unsigned drained[RING];  // warps done with the stage's current chunk; the last one to arrive refills the stage
};

// running use counts of the mbarriers (parity = count & 1), carried across tiles, items and steps
struct PipeState {
  unsigned chunk;  // weight chunks consumed so far by this CTA
  unsigned pass;   // phase-2 panel passes so far
};

// what a phase-1 tile needs to know about its step
struct StepCtx {
  int rows, n_tiles;         // this rank's rows of the batch / their 16-row tiles
  int row_base;              // position of this rank's first row in the global batch (dropout and injected masks are keyed by it)
  int B;                     // rows of the global batch: the BatchNorm population
  float inv_rows, inv_B, inv_Bm1;  // 1 / rows, 1 / B, 1 / max(B - 1, 1)
  int per[4];                // tiles per slice of a reduction point of BatchNorm i
  unsigned tag;              // packet tag of the step
  unsigned long long dstep;  // dropout stream position of the step
  bool fwd_only;
  long long* prof;           // diagnostics (CTA 0, armed step): clock stamps of the reduction points at [520 + 8 pt + k]
};

// ------------------------------------------------------------------------------------------------ AE_Dropout_BN
// reference baler/modules/models.py:256-313 in train mode:
//   encoder 4 x (Linear -> Dropout(p = .5,.4,.3,.2) -> LeakyReLU)       [activation also on the latent]
//   decoder 3 x (Linear -> LeakyReLU -> BatchNorm1d) + Linear -> BatchNorm1d -> ReLU
// BatchNorm in train mode normalises with the statistics of the WHOLE batch (biased variance, eps 1e-5), which is spread
// over the phase-1 CTAs (16 rows each) and, data parallel, over the ranks.  At each of the 8 reduction points of a step
// (4 forward: column mean / variance; 4 backward: sum dY, sum dY xhat) every CTA publishes its tile's column sums as
// 16-byte {a, tag, b, tag} packets (each 8-byte half carries the step tag, so a reader spins on the data itself: no
// counter, no fence) and reads all tiles' packets back, thread j owning column j, summing in tile order in double.
// Data parallel: CTA 0 of every rank pushes the rank's sums into all peers' buffers over NVLink the same way and every
// CTA combines the ranks' packets in rank order, so all replicas normalise with the same bits - BatchNorm over the
// global batch, exactly what a single GPU computes at batch_size = global batch.
// one thread's epilogue values: rows g, g + 8 x column pairs of its NJ n-tiles
template <int NJ>
using Vals = float[NJ > 0 ? NJ * 2 : 1][2];

// Packet loads are relaxed (strong) loads at the scope of the writer: gpu for the tiles of this GPU, sys for what peers
// push over NVLink.  (Not volatile: ptxas completes a volatile access before it issues the next one.  And not weak
// ld.global.cg / L1::no_allocate either: a weak load carries no inter-thread guarantee, and ptxas did rewrite a
// `while (tag mismatch) reload` loop around one into a single unchecked reload.)  Independent relaxed loads are all in
// flight together; the tag inside each 8-byte half makes a torn read harmless: it is simply retried.
__device__ __forceinline__ uint4 ld_pkt(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_pkt_sys(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_pkt(uint4* p, const float a, const float b, const unsigned tag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag), "r"(__float_as_uint(b)), "r"(tag) : "memory");
}
__device__ __forceinline__ void st_pkt_sys(uint4* p, const float a, const float b, const unsigned tag) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(__float_as_uint(a)), "r"(tag), "r"(__float_as_uint(b)), "r"(tag) : "memory");
}
__device__ __forceinline__ float sum_over_rows(float s) {  // the 8 lanes that hold one column pair differ in lane bits 2..4
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 8);
  s += __shfl_xor_sync(0xffffffffu, s, 16);
  return s;
}
__device__ __forceinline__ int dp_rank_rows(const int B, const int world, const int r) {
  const int base = B / world;
  return base + (r < B - base * world ? 1 : 0);
}

// Reduction point `pt` over the tiles of this rank (and, data parallel, over the ranks) as reduce-scatter + all-gather,
// so that a thread has ONE packet in flight per hop instead of one per tile (reading every tile's packet of every column in
// every CTA measured 1.4 k cycles per 8 packets, 4 rounds for the 200-wide layers: 4 - 12 k cycles per point).
//   hop 1  column j belongs to the CTA of tile j % n_tiles.  Warp w of that CTA takes its q-th column, lane l the packets
//          of tiles l, l + 32, ...; a butterfly over the lanes adds them (fixed pattern: reproducible).  Forward sums are
//          taken about a shift K = the mean of tile 0, the same for all lanes: sum (x - K) and sum (x - K)^2 are additive
//          and M2 = sum (x - K)^2 - (sum (x - K))^2 / n loses nothing to cancellation (|mean - K| is a fraction of sigma).
//          All of it in fp32: FP64 instructions issue at a small fraction of the fp32 rate here.
//          Data parallel: lane r pushes the rank's (mean, M2) or (sum dY, sum dY xhat) of the column to rank r and polls
//          rank r's packet; the ranks are combined in rank order (Chan), so every rank holds the same bits.
//          Lane 0 publishes the result - (mean, 1 / sqrt(var + eps)) or (T1 / B, T2 / B) - as the column's final packet.
//   hop 2  thread j of every CTA reads the final packet of column j and leaves in shared memory what the warps need for
//          their values: mean | inv (kept for the backward pass) and gamma | beta, or T1 / B | T2 / B | gamma inv.
template <bool FWD>
__device__ __forceinline__ void bn_reduce(const TcModel& M, const TcPtrs& P, const StepCtx& sc, const int bi, const int pt,
                                       const int tile, unsigned char* smem) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, N = M.bn_n[bi], fo = M.bn_f_off[bi];
  float* stat = reinterpret_cast<float*>(smem + M.stat_off);
  float* tr = stat + 2 * M.bn_f_total;  // transient: [gamma | beta | T1 / B | T2 / B or gamma inv] x BN_TW
  float gam = 0.f, bet = 0.f;
  if (tid < N) {
    gam = __ldcg(P.params + M.bn_g_off[bi] + tid);
    if (FWD) bet = __ldcg(P.params + M.bn_b_off[bi] + tid);
  }
  const bool pk0 = sc.prof != nullptr && tid == 0;
  if (pk0) sc.prof[520 + 8 * pt + 1] = clock64();
  const int n_tiles = sc.n_tiles;
  const uint4* part = P.bn_part + (size_t)pt * M.max_tiles * BN_PW;
  uint4* fin = P.bn_fin + (size_t)pt * BN_PW;
  const int last = sc.rows - (n_tiles - 1) * TROWS;               // rows of the last tile
  const float n_last = (float)last, inv_last = __frcp_rn(n_last);  // (1 / 16 is exact; 1 / last is used for one tile)
  // ---- hop 1: the columns this CTA owns
  for (int j = tile + warp * n_tiles; j < N; j += NWARPS * n_tiles) {
    uint4 pk[5];
#pragma unroll
    for (int u = 0; u < 5; ++u)
      if (lane + 32 * u < n_tiles) pk[u] = ld_pkt(part + (size_t)(lane + 32 * u) * BN_PW + j);
    bool missing;
    unsigned spins = 0;
    do {
      missing = false;
#pragma unroll
      for (int u = 0; u < 5; ++u)
        if (lane + 32 * u < n_tiles && (pk[u].y != sc.tag || pk[u].w != sc.tag)) {
          pk[u] = ld_pkt(part + (size_t)(lane + 32 * u) * BN_PW + j);
          missing = true;
        }
      spin_guard(spins);
    } while (missing);
    float K = 0.f;
    if (FWD) K = __shfl_sync(0xffffffffu, __uint_as_float(pk[0].x) * (n_tiles == 1 ? inv_last : 1.f / TROWS), 0);
    float S1 = 0.f, S2 = 0.f;
#pragma unroll
    for (int u = 0; u < 5; ++u)
      if (lane + 32 * u < n_tiles) {
        const float a = __uint_as_float(pk[u].x), b = __uint_as_float(pk[u].z);
        if (FWD) {
          const bool is_last = lane + 32 * u == n_tiles - 1;
          const float d = a - (is_last ? n_last : (float)TROWS) * K;
          S1 += d;
          S2 += fmaf(d * d, is_last ? inv_last : 1.f / TROWS, b);
        } else {
          S1 += a;
          S2 += b;
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      S1 += __shfl_xor_sync(0xffffffffu, S1, o);
      S2 += __shfl_xor_sync(0xffffffffu, S2, o);
    }
    float r1, r2;  // FWD: mean, M2 of this rank's rows; else its totals
    if (FWD) {
      const float dm = S1 * sc.inv_rows;
      r1 = K + dm;
      r2 = fmaxf(S2 - S1 * dm, 0.f);
    } else {
      r1 = S1;
      r2 = S2;
    }
    if (P.dp_slice) {
      // lane r talks to rank r; every rank combines the same fp32 numbers in rank order: identical statistics everywhere
      const size_t blk = (((size_t)(sc.tag & 1u) * 8 + pt) * P.world) * BN_PW + j;
      float a = r1, b = r2;
      const int nr = lane < P.world ? dp_rank_rows(sc.B, P.world, lane) : 0;
      if (lane < P.world && lane != P.rank) {
        st_pkt_sys(P.bnx[lane] + blk + (size_t)P.rank * BN_PW, r1, r2, sc.tag);
        if (nr > 0) {
          const uint4* src = P.bnx[P.rank] + blk + (size_t)lane * BN_PW;
          uint4 q = ld_pkt_sys(src);
          unsigned spins_x = 0;
          while (q.y != sc.tag || q.w != sc.tag) { q = ld_pkt_sys(src); spin_guard(spins_x); }
          a = __uint_as_float(q.x);
          b = __uint_as_float(q.z);
        }
      }
      float cn = 0.f, c1 = 0.f, c2 = 0.f;
      for (int r = 0; r < P.world; ++r) {
        const float ar = __shfl_sync(0xffffffffu, a, r), br = __shfl_sync(0xffffffffu, b, r);
        const int nrr = __shfl_sync(0xffffffffu, nr, r);
        if (nrr == 0) continue;
        if (FWD) {  // Chan's pairwise update of (count, mean, M2)
          const float delta = ar - c1, nn = cn + (float)nrr;
          const float w = __fdiv_rn((float)nrr, nn);
          c1 = fmaf(delta, w, c1);
          c2 += fmaf(delta * delta, cn * w, br);
          cn = nn;
        } else {
          c1 += ar;
          c2 += br;
        }
      }
      r1 = c1;
      r2 = c2;
    }
    if (lane == 0) {
      if (FWD) {
        const float inv = __frcp_rn(__fsqrt_rn(fmaf(r2, sc.inv_B, BN_EPS)));  // IEEE roundings: ~1 ulp
        st_pkt(fin + j, r1, inv, sc.tag);
        // running statistics: momentum 0.1, unbiased variance (torch.nn.BatchNorm1d)
        P.rm[fo + j] = (1.f - BN_MOMENTUM) * P.rm[fo + j] + BN_MOMENTUM * r1;
        P.rv[fo + j] = (1.f - BN_MOMENTUM) * P.rv[fo + j] + BN_MOMENTUM * (r2 * sc.inv_Bm1);
        if (j == 0) P.nbt[bi] += 1;
      } else {
        st_pkt(fin + j, r1 * sc.inv_B, r2 * sc.inv_B, sc.tag);
        P.grads[M.bn_b_off[bi] + j] = r1;
        P.grads[M.bn_g_off[bi] + j] = r2;
      }
    }
  }
  if (pk0) sc.prof[520 + 8 * pt + 2] = clock64();
  // ---- hop 2: every CTA gathers the final packets
  if (tid < N) {
    uint4 q = ld_pkt(fin + tid);
    unsigned spins_f = 0;
    while (q.y != sc.tag || q.w != sc.tag) { q = ld_pkt(fin + tid); spin_guard(spins_f); }
    const float v1 = __uint_as_float(q.x), v2 = __uint_as_float(q.z);
    if (FWD) {
      stat[fo + tid] = v1;
      stat[M.bn_f_total + fo + tid] = v2;
      tr[tid] = gam;
      tr[BN_TW + tid] = bet;
    } else {
      tr[2 * BN_TW + tid] = v1;
      tr[3 * BN_TW + tid] = v2;
      tr[tid] = gam * stat[M.bn_f_total + fo + tid];
    }
  }
  if (pk0) sc.prof[520 + 8 * pt + 4] = clock64();
  __syncthreads();
  if (pk0) sc.prof[520 + 8 * pt + 5] = clock64();
}

// eval mode: the running statistics
__device__ __forceinline__ void bn_eval_stats(const TcModel& M, const TcPtrs& P, const int bi, unsigned char* smem) {
  const int j = threadIdx.x, N = M.bn_n[bi], fo = M.bn_f_off[bi];
  float* stat = reinterpret_cast<float*>(smem + M.stat_off);
  if (j < N) {
    stat[fo + j] = __ldcg(P.rm + fo + j);
    stat[M.bn_f_total + fo + j] = 1.f / sqrtf(__ldcg(P.rv + fo + j) + BN_EPS);
    stat[2 * M.bn_f_total + j] = __ldcg(P.params + M.bn_g_off[bi] + j);
    stat[2 * M.bn_f_total + BN_TW + j] = __ldcg(P.params + M.bn_b_off[bi] + j);
  }
  __syncthreads();
}

// Dropout of encoder layer l on this thread's values (rows g, g + 8; column pairs of its n-tiles): Philox4x32-10 keyed by
// (seed, step) with counter (column / 4, global batch row, layer) - the same stream as train_dbn_kernel in bb_train.cu,
// independent of the launch geometry and of how the batch is split over ranks - or the injected keep-masks.
template <int NJ>
__device__ __forceinline__ void dbn_dropout(Vals<NJ>& v, const TcModel& M, const TcPtrs& P, const StepCtx& sc, const int l,
                                            const int N, const int warp, const int g, const int t, const int row0) {
  const float ks = M.keep_scale[l];
  const unsigned thr = M.keep_thr[l];
  const unsigned char* mk = P.mask[l];
  if (mk != nullptr) {
#pragma unroll
    for (int q = 0; q < NJ * 2; ++q) {
      const int col = (warp + NWARPS * (q >> 1)) * 8 + 2 * t, r = row0 + g + 8 * (q & 1);
      bool k0 = false, k1 = false;
      if (r < sc.rows && col < N) {
        const size_t grow = (size_t)(sc.row_base + r);
        k0 = mk[grow * N + col] != 0;
        k1 = col + 1 < N && mk[grow * N + col + 1] != 0;
      }
      v[q][0] = k0 ? v[q][0] * ks : 0.f;
      v[q][1] = k1 ? v[q][1] * ks : 0.f;
    }
    return;
  }
  // One Philox call yields the draws of 4 consecutive columns of a row.  The two lanes of a pair (t, t ^ 1) hold columns
  // 4c..4c+1 and 4c+2..4c+3 of rows g and g + 8: the even lane draws row g, the odd lane row g + 8, and they swap halves.
  const bool odd = (t & 1) != 0;
  const uint2 key = make_uint2((unsigned)P.seed, (unsigned)(P.seed >> 32) ^ (unsigned)(sc.dstep >> 32));
#pragma unroll 2
  for (int j = 0; j < NJ; ++j) {
    const int col = (warp + NWARPS * j) * 8 + 2 * t;
    const unsigned grow = (unsigned)(sc.row_base + row0 + g + (odd ? 8 : 0));
    const uint4 d = curand_Philox4x32_10(make_uint4((unsigned)(col >> 2), grow, (unsigned)l, (unsigned)sc.dstep), key);
    const unsigned o0 = __shfl_xor_sync(0xffffffffu, odd ? d.x : d.z, 1);
    const unsigned o1 = __shfl_xor_sync(0xffffffffu, odd ? d.y : d.w, 1);
    // draws of (row g, col), (row g, col + 1), (row g + 8, col), (row g + 8, col + 1)
    const unsigned a0 = odd ? o0 : d.x, a1 = odd ? o1 : d.y, b0 = odd ? d.z : o0, b1 = odd ? d.w : o1;
    const bool va = row0 + g < sc.rows, vb = row0 + g + 8 < sc.rows;
    v[2 * j][0] = va && col < N && a0 < thr ? v[2 * j][0] * ks : 0.f;
    v[2 * j][1] = va && col + 1 < N && a1 < thr ? v[2 * j][1] * ks : 0.f;
    v[2 * j + 1][0] = vb && col < N && b0 < thr ? v[2 * j + 1][0] * ks : 0.f;
    v[2 * j + 1][1] = vb && col + 1 < N && b1 < thr ? v[2 * j + 1][1] * ks : 0.f;
  }
}

// A BatchNorm epilogue runs in phases around the reduction (phase1_tile): publish -> reduce -> apply.  Nothing stays in
// registers across a reduction: the BatchNorm input a is parked in the fp32 copy U (which the backward pass needs
// anyway) and a gradient dY in the pass's own output slots (the hi / lo words of a column pair hold the two floats), each
// value read back by the thread that wrote it.  The reduction is then one piece of code per direction instead of one per
// epilogue variant, and it has the whole register file for its packets (with this kernel's shared-memory footprint the
// L1 that would back register spills is ~28 KB: a spill is an L2 round trip).
struct Own {  // this thread's position in the epilogue layout
  int warp, g, t, row0;
};
template <int NJ>
__device__ __forceinline__ int own_col(const Own& o, const int q) { return (o.warp + NWARPS * (q >> 1)) * 8 + 2 * o.t; }

// publish (sum, sum of squares about the tile mean) of this tile's columns of BatchNorm input `a` and park a in U
template <int NJ>
__device__ __forceinline__ void bn_publish_fwd(const Vals<NJ>& a, const TcModel& M, const TcPtrs& P, const StepCtx& sc, const int bi,
                                               const int tile, const Own& o, unsigned char* smem) {
  const int N = M.bn_n[bi], uld = M.u_ld[bi];
  float* U = reinterpret_cast<float*>(smem + M.u_off[bi]);
  const bool vr[2] = {o.row0 + o.g < sc.rows, o.row0 + o.g + 8 < sc.rows};
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    const int col = own_col<NJ>(o, q), r = o.g + 8 * (q & 1);
    if (col < N) *reinterpret_cast<float2*>(U + r * uld + col) = make_float2(a[q][0], a[q][1]);
  }
  if (!P.train) return;
  if (sc.prof != nullptr && threadIdx.x == 0) sc.prof[520 + 8 * bi] = clock64();
  const int left = sc.rows - o.row0;
  const float inv_nt = 1.f / (float)(left < TROWS ? left : TROWS);
  uint4* part = P.bn_part + ((size_t)bi * M.max_tiles + tile) * BN_PW;
#if 0
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int col = own_col<NJ>(o, 2 * j);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float a0 = vr[0] ? a[2 * j][e] : 0.f, a1 = vr[1] ? a[2 * j + 1][e] : 0.f;
      const float s = sum_over_rows(a0 + a1);
      const float mean = s * inv_nt;
      const float d0 = vr[0] ? a0 - mean : 0.f, d1 = vr[1] ? a1 - mean : 0.f;
      const float m2 = sum_over_rows(d0 * d0 + d1 * d1);
      if (o.g == 0 && col + e < N) st_pkt(part + col + e, s, m2, sc.tag);
    }
  }
}
#else
  // the column sums over the 16 rows: independent shuffles issued back to back, one butterfly stage at a time (a warp
  // issues in order: a dependent shuffle right behind its source waits out the ~25-cycle latency)
  Vals<NJ> s, m2;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) s[j][e] = (vr[0] ? a[2 * j][e] : 0.f) + (vr[1] ? a[2 * j + 1][e] : 0.f);
#pragma unroll
  for (int st = 4; st <= 16; st <<= 1)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) s[j][e] += __shfl_xor_sync(0xffffffffu, s[j][e], st);
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float mean = s[j][e] * inv_nt;
      const float d0 = vr[0] ? a[2 * j][e] - mean : 0.f, d1 = vr[1] ? a[2 * j + 1][e] - mean : 0.f;
      m2[j][e] = d0 * d0 + d1 * d1;
    }
#pragma unroll
  for (int st = 4; st <= 16; st <<= 1)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) m2[j][e] += __shfl_xor_sync(0xffffffffu, m2[j][e], st);
  if (o.g == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = own_col<NJ>(o, 2 * j);
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (col + e < N) st_pkt(part + col + e, s[j][e], m2[j][e], sc.tag);
    }
  }
}
#endif

// xhat of this thread's values from the parked BatchNorm input (0 outside the valid rows / columns); u = the input itself
template <int NJ>
__device__ __forceinline__ void bn_load_xhat(Vals<NJ>& xh, Vals<NJ>& u, const TcModel& M, const StepCtx& sc, const int bi,
                                             const Own& o, unsigned char* smem) {
  const int N = M.bn_n[bi], uld = M.u_ld[bi], fo = M.bn_f_off[bi];
  const float* U = reinterpret_cast<const float*>(smem + M.u_off[bi]);
  const float* stat = reinterpret_cast<const float*>(smem + M.stat_off);
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    const int col = own_col<NJ>(o, q), r = o.g + 8 * (q & 1);
    const bool vr = o.row0 + r < sc.rows;
    float2 uu = make_float2(0.f, 0.f);
    if (col < N) uu = *reinterpret_cast<const float2*>(U + r * uld + col);
    u[q][0] = uu.x; u[q][1] = uu.y;
#pragma unroll
    for (int e = 0; e < 2; ++e)
      xh[q][e] = vr && col + e < N ? (u[q][e] - stat[fo + col + e]) * stat[M.bn_f_total + fo + col + e] : 0.f;
  }
}

// publish (sum dY, sum dY xhat) of this tile's columns and park dY in the output slots
template <int NJ>
__device__ __forceinline__ void bn_publish_bwd(const Vals<NJ>& dy, const Vals<NJ>& xh, const TcModel& M, const TcPtrs& P,
                                               const StepCtx& sc, const int bi, const int pt, const int tile, const Own& o,
                                               const uint32_t o_hi, const uint32_t o_lo, const uint32_t o_ld) {
  const int N = M.bn_n[bi];
  if (sc.prof != nullptr && threadIdx.x == 0) sc.prof[520 + 8 * pt] = clock64();
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    const uint32_t off = (o.g + 8 * (q & 1)) * o_ld + own_col<NJ>(o, q) * 2;
    sts32(o_hi + off, __float_as_uint(dy[q][0]));
    sts32(o_lo + off, __float_as_uint(dy[q][1]));
  }
  uint4* part = P.bn_part + ((size_t)pt * M.max_tiles + tile) * BN_PW;
  Vals<NJ> s1, s2;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      s1[j][e] = dy[2 * j][e] + dy[2 * j + 1][e];
      s2[j][e] = fmaf(dy[2 * j][e], xh[2 * j][e], dy[2 * j + 1][e] * xh[2 * j + 1][e]);
    }
#pragma unroll
  for (int st = 4; st <= 16; st <<= 1)
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        s1[j][e] += __shfl_xor_sync(0xffffffffu, s1[j][e], st);
        s2[j][e] += __shfl_xor_sync(0xffffffffu, s2[j][e], st);
      }
  if (o.g == 0) {
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = own_col<NJ>(o, 2 * j);
#pragma unroll
      for (int e = 0; e < 2; ++e)
        if (col + e < N) st_pkt(part + col + e, s1[j][e], s2[j][e], sc.tag);
    }
  }
}

// after the backward reduction: dx = gamma inv / B (B dy - sum dy - xhat sum dy xhat), times the LeakyReLU slope of the
// BatchNorm input when the activation sits in front of the BatchNorm; split to hi | lo in place of the parked dY
template <int NJ>
__device__ __forceinline__ void bn_apply_bwd(const TcModel& M, const StepCtx& sc, const int bi, const bool leaky, const Own& o,
                                             const uint32_t o_hi, const uint32_t o_lo, const uint32_t o_ld, unsigned char* smem) {
  const int N = M.bn_n[bi];
  const float* tr = reinterpret_cast<const float*>(smem + M.stat_off) + 2 * M.bn_f_total;
  Vals<NJ> xh, u;
  bn_load_xhat<NJ>(xh, u, M, sc, bi, o, smem);
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    const int col = own_col<NJ>(o, q);
    const bool vr = o.row0 + o.g + 8 * (q & 1) < sc.rows;
    const uint32_t off = (o.g + 8 * (q & 1)) * o_ld + col * 2;
    const float dy[2] = {__uint_as_float(lds32(o_hi + off)), __uint_as_float(lds32(o_lo + off))};
    float dx[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      dx[e] = 0.f;
      if (vr && col + e < N) {
        const int c = col + e;
        dx[e] = tr[c] * (dy[e] - tr[2 * BN_TW + c] - xh[q][e] * tr[3 * BN_TW + c]);
        if (leaky && !(u[q][e] > 0.f)) dx[e] *= BB_LEAKY;
      }
    }
    uint32_t hi, lo;
    split16x2(dx[0], dx[1], hi, lo);
    sts32(o_hi + off, hi);
    sts32(o_lo + off, lo);
  }
}

// after the forward reduction of a hidden decoder layer: X_{l+1} = gamma xhat + beta, the bias column of ones, zero padding
template <int NJ>
__device__ __forceinline__ void bn_apply_fwd(const TcModel& M, const StepCtx& sc, const int bi, const Own& o, const uint32_t o_hi,
                                             const uint32_t o_lo, const uint32_t o_ld, unsigned char* smem) {
  const int N = M.bn_n[bi];
  const float* tr = reinterpret_cast<const float*>(smem + M.stat_off) + 2 * M.bn_f_total;  // gamma | beta
  Vals<NJ> xh, u;
  bn_load_xhat<NJ>(xh, u, M, sc, bi, o, smem);
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    const int col = own_col<NJ>(o, q);
    const bool vr = o.row0 + o.g + 8 * (q & 1) < sc.rows;
    float y[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      y[e] = 0.f;
      if (vr && col + e < N) y[e] = fmaf(xh[q][e], tr[col + e], tr[BN_TW + col + e]);
      if (vr && col + e == N) y[e] = 1.f;
    }
    uint32_t hi, lo;
    split16x2(y[0], y[1], hi, lo);
    const uint32_t off = (o.g + 8 * (q & 1)) * o_ld + col * 2;
    sts32(o_hi + off, hi);
    sts32(o_lo + off, lo);
  }
}

// last layer after the forward reduction: reconstruction = ReLU(gamma xhat + beta), loss = sum (recon - x)^2 / F, the seed
// gradient back through the ReLU; publishes the sums of the BatchNorm backward (train) - the values land in dZ_7's slots
template <int NJ>
__device__ __forceinline__ void bn_loss(const TcModel& M, const TcPtrs& P, const StepCtx& sc, const int tile, const Own& o,
                                        const uint32_t o_hi, const uint32_t o_lo, const uint32_t o_ld, const float* xs, float& loss,
                                        unsigned char* smem) {
  const int F = M.F;
  const float inv_f = 1.f / (float)F;
  const float* tr = reinterpret_cast<const float*>(smem + M.stat_off) + 2 * M.bn_f_total;  // gamma | beta
  Vals<NJ> xh, u;
  bn_load_xhat<NJ>(xh, u, M, sc, 3, o, smem);
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    const int col = own_col<NJ>(o, q), r = o.g + 8 * (q & 1);
    const bool valid = o.row0 + r < sc.rows;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      float gz = 0.f;
      if (valid && col + e < F) {
        const float y = fmaf(xh[q][e], tr[col + e], tr[BN_TW + col + e]);
        const float diff = fmaxf(y, 0.f) - xs[r * F + col + e];
        loss = fmaf(diff * diff, inv_f, loss);
        gz = y > 0.f ? 2.f * diff * inv_f : 0.f;
      }
      u[q][e] = gz;
    }
  }
  if (!sc.fwd_only) bn_publish_bwd<NJ>(u, xh, M, P, sc, 3, 4, tile, o, o_hi, o_lo, o_ld);
}

// ------------------------------------------------------------------------------------------------ phase 1
// A weight chunk goes into ring stage (global chunk number % RING).  The warp that is last to finish a stage's chunk
// (shared counter) immediately refills the stage with the chunk RING positions ahead: no producer warp, nobody waits.
template <int RN>
__device__ __forceinline__ void ring_issue(const TcModel& M, const TcPtrs& P, SmemBars* bars, uint4* ring, const unsigned c_base,
                                           const int c, const int n_chunks) {
  if (c >= n_chunks) return;
  const int2 ch = *reinterpret_cast<const int2*>(&M.chunk[c]);
  const unsigned st = (c_base + c) % RN;
  mbar_expect_tx(&bars->full[st], (uint32_t)ch.y * 16u);
  bulk_g2s(ring + (size_t)st * CHUNK_U4, P.img + ch.x, (uint32_t)ch.y * 16u, &bars->full[st]);
}

// One layer pass for a warp that owns exactly NJ n-tiles (warp, warp + 8, ...): the k loop over the pass's weight
// chunks, then the epilogue.  One instantiation per tile count: predicated-off HMMAs occupy the tensor pipe like live
// ones (measured with ncu: 7.4 k issue slots for 3.2 k useful products when the loop was predicated on 4 tiles).
// Fragments of k-step kk + 1 are fetched before the products of k-step kk are issued, and within a k-step the independent
// products of the warp's tiles are issued back to back (the warp issues in order: a dependent HMMA or an fp32 add right
// behind its HMMA stalls everything after it for the ~21-cycle MMA latency).  The hi * hi products of two consecutive
// k-steps share one zero-started accumulator before the rounded fp32 add.
template <int NJ, bool DBN>
__device__ __forceinline__ void run_pass(const TcModel& M, const TcPtrs& P, SmemBars* bars, uint4* ring, const uint32_t sbase,
                                         const TcPass& ps, const unsigned c_base, const int n_chunks, const int warp, const int lane,
                                         const int row0, const StepCtx& sc, const int tile, unsigned char* smem, const float* xs,
                                         float& loss, long long* prof, const int pass) {
  constexpr int RN = DBN ? RING_DBN : RING;
  const int rows = sc.rows;
  const int g = lane >> 2, t = lane & 3;
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_colb = (lane >> 4) * 16;
  float acc[NJ > 0 ? NJ : 1][2][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][q][i] = 0.f;
  const int NT = ps.NT;
  const uint32_t b_step = NT * 512;
  for (int c = ps.c0; c < ps.c1; ++c) {
    const uint4 d0 = *reinterpret_cast<const uint4*>(&M.chunk[c]), d1 = *(reinterpret_cast<const uint4*>(&M.chunk[c]) + 1);
    const int n_k = d0.z >> 16;
    const unsigned gc = c_base + c, st = gc % RN;
    mbar_wait(&bars->full[st], (gc / RN) & 1u);
    if (NJ > 0) {
      struct Frag { uint32_t ah[4], al[4]; uint4 b[NJ > 0 ? NJ : 1]; };
      const uint32_t pa_hi = sbase + d1.x + a_row * d1.z + a_colb, pa_lo = sbase + d1.y + a_row * d1.z + a_colb;
      const uint32_t pb = sbase + st * (CHUNK_U4 * 16) + (warp * 32 + lane) * 16;
      auto fetch = [&](Frag& f, const int kk) {
        ldsm_x4a(f.ah, pa_hi + kk * 32);
        ldsm_x4a(f.al, pa_lo + kk * 32);
#pragma unroll
        for (int j = 0; j < NJ; ++j) f.b[j] = lds128(pb + kk * b_step + j * (NWARPS * 512));
      };
      float tmp[NJ > 0 ? NJ : 1][4];
      auto products = [&](const Frag& f, const bool fresh, const bool flush) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          if (fresh) tmp[j][0] = tmp[j][1] = tmp[j][2] = tmp[j][3] = 0.f;
          mma16816(tmp[j], f.ah, f.b[j].x, f.b[j].y);
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma16816(acc[j][1], f.ah, f.b[j].z, f.b[j].w);
#pragma unroll
        for (int j = 0; j < NJ; ++j) mma16816(acc[j][1], f.al, f.b[j].x, f.b[j].y);
        if (flush) {
#pragma unroll
          for (int j = 0; j < NJ; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][0][i] += tmp[j][i];
        }
      };
      Frag f0, f1;
      fetch(f0, 0);
      for (int kk = 0; kk < n_k; kk += 2) {
        const bool more = kk + 1 < n_k;
        if (more) fetch(f1, kk + 1);
        products(f0, true, true);
        if (more) {
          if (kk + 2 < n_k) fetch(f0, kk + 2);
          products(f1, true, true);
        }
      }
    }
    __syncwarp();
    if (lane == 0 && atomicAdd(&bars->drained[st], 1u) == NWARPS - 1) {  // last warp out refills the stage
      bars->drained[st] = 0u;
      ring_issue<RN>(M, P, bars, ring, c_base, c + RN, n_chunks);
    }
  }
  if (prof && lane == 0) prof[16 + (pass * NWARPS + warp) * 4 + 1] = clock64();
  // ---- epilogue
  const int N = ps.N, act = ps.act, kind = ps.kind, pact = ps.pact, F = M.F;
  const uint32_t o_hi = sbase + ps.o_hi_off, o_lo = sbase + ps.o_lo_off, o_ld = ps.o_ld_b;
  const uint32_t x_hi = sbase + ps.x_hi_off, x_lo = sbase + ps.x_lo_off, x_ld = ps.x_ld_b;
  const float inv_f = 1.f / (float)F;
  Vals<NJ> v;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      v[j * 2 + h][0] = acc[j][0][2 * h] + acc[j][1][2 * h] * LO_INV;
      v[j * 2 + h][1] = acc[j][0][2 * h + 1] + acc[j][1][2 * h + 1] * LO_INV;
    }
  if (kind == 0) {
    // X_{l+1} = act(X_l W^T + b); column N is the next layer's bias column of ones, beyond it zero padding
    if (DBN && ps.drop >= 0 && P.train) dbn_dropout<NJ>(v, M, P, sc, ps.drop, N, warp, g, t, row0);
    if (DBN && ps.bn >= 0) {
      // decoder layer: LeakyReLU, then BatchNorm over the whole batch
#pragma unroll
      for (int q = 0; q < NJ * 2; ++q) {
        v[q][0] = act_apply(v[q][0], act);
        v[q][1] = act_apply(v[q][1], act);
      }
      bn_publish_fwd<NJ>(v, M, P, sc, ps.bn, tile, Own{warp, g, t, row0}, smem);
      return;  // phase1_tile: reduction, then bn_apply_fwd
    } else {
#pragma unroll
      for (int q = 0; q < NJ * 2; ++q) {
        const int col = (warp + NWARPS * (q >> 1)) * 8 + 2 * t;
        const float one = row0 + g + 8 * (q & 1) < rows ? 1.f : 0.f;
        v[q][0] = col == N ? one : act_apply(v[q][0], act);
        v[q][1] = col + 1 == N ? one : act_apply(v[q][1], act);
      }
    }
  } else if (kind == 1 && DBN) {
    // reconstruction = ReLU(BatchNorm(X_7 W_7^T + b_7)); loss = sum (recon - x)^2 / F; the seed gradient goes back through
    // the ReLU and the BatchNorm to dZ_7
    bn_publish_fwd<NJ>(v, M, P, sc, 3, tile, Own{warp, g, t, row0}, smem);
    return;  // phase1_tile: reduction, bn_loss, reduction, bn_apply_bwd
  } else if (kind == 1) {
    // reconstruction: loss = sum (recon - x)^2 / F, seed gradient dZ_7 = 2 (recon - x) / F
#pragma unroll
    for (int q = 0; q < NJ * 2; ++q) {
      const int col = (warp + NWARPS * (q >> 1)) * 8 + 2 * t, r = g + 8 * (q & 1);
      const bool valid = row0 + r < rows;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float gz = 0.f;
        if (valid && col + e < F) {
          const float diff = act_apply(v[q][e], act) - xs[r * F + col + e];
          loss = fmaf(diff * diff, inv_f, loss);
          gz = 2.f * diff * inv_f;
          if (act == BB_ACT_RELU && v[q][e] <= 0.f) gz = 0.f;
        }
        v[q][e] = gz;
      }
    }
  } else if (DBN && ps.bn >= 0) {
    // v = gradient w.r.t. the output of BatchNorm ps.bn (= dZ_{l} W_{l}); back through the BatchNorm and the LeakyReLU
    // in front of it.  The BatchNorm input U is this thread's own copy from the forward pass.
    const Own o{warp, g, t, row0};
    const int Nb = M.bn_n[ps.bn];
    Vals<NJ> xh, u;
    bn_load_xhat<NJ>(xh, u, M, sc, ps.bn, o, smem);
#pragma unroll
    for (int q = 0; q < NJ * 2; ++q) {
      const int col = own_col<NJ>(o, q);
      const bool vr = row0 + g + 8 * (q & 1) < rows;
      if (!(vr && col < Nb)) v[q][0] = 0.f;
      if (!(vr && col + 1 < Nb)) v[q][1] = 0.f;
    }
    bn_publish_bwd<NJ>(v, xh, M, P, sc, ps.bn, 7 - ps.bn, tile, o, o_hi, o_lo, o_ld);
    return;  // phase1_tile: reduction, then bn_apply_bwd
  } else {
    if (pact != BB_ACT_NONE) {
      // dZ_{l-1} = (dZ_l W_l) * act'_{l-1}; the sign of the pre-activation is the sign of X_l (slope > 0)
      uint32_t sh[NJ > 0 ? NJ * 2 : 1], sl[NJ > 0 ? NJ * 2 : 1];
#pragma unroll
      for (int q = 0; q < NJ * 2; ++q) {
        const uint32_t xo = (g + 8 * (q & 1)) * x_ld + ((warp + NWARPS * (q >> 1)) * 8 + 2 * t) * 2;
        sh[q] = lds32(x_hi + xo);
        sl[q] = lds32(x_lo + xo);
      }
      const float neg = pact == BB_ACT_LEAKY ? BB_LEAKY : 0.f;
#pragma unroll
      for (int q = 0; q < NJ * 2; ++q) {
        const float2 xh = half2_bits_to_float2(sh[q]), xl = half2_bits_to_float2(sl[q]);
        if (!(xh.x > 0.f || (xh.x == 0.f && xl.x > 0.f))) v[q][0] *= neg;
        if (!(xh.y > 0.f || (xh.y == 0.f && xl.y > 0.f))) v[q][1] *= neg;
      }
    }
    // encoder layer of AE_Dropout_BN: the same keep-mask and 1 / (1 - p) as in the forward pass
    if (DBN && ps.drop >= 0 && P.train) dbn_dropout<NJ>(v, M, P, sc, ps.drop, M.L[ps.drop].N, warp, g, t, row0);
  }
#pragma unroll
  for (int q = 0; q < NJ * 2; ++q) {
    uint32_t hi, lo;
    split16x2(v[q][0], v[q][1], hi, lo);
    const uint32_t off = (g + 8 * (q & 1)) * o_ld + ((warp + NWARPS * (q >> 1)) * 8 + 2 * t) * 2;
    sts32(o_hi + off, hi);
    sts32(o_lo + off, lo);
  }
  if (prof && lane == 0) prof[16 + (pass * NWARPS + warp) * 4 + 2] = clock64();
}

// forward + loss + backward of one 16-row tile, all 8 warps alike
template <bool DBN>
__device__ __forceinline__ void phase1_tile(const TcModel& M, const TcPtrs& P, const float* __restrict__ xg, const StepCtx& sc, const int tile,
                            const int loss_slot, unsigned char* smem, SmemBars* bars, PipeState& ps, long long* prof) {
  constexpr int RN = DBN ? RING_DBN : RING;
  const int tid = threadIdx.x, warp = uni(tid >> 5), lane = tid & 31;
  const int rows = sc.rows;
  const bool fwd_only = sc.fwd_only;
  uint4* ring = reinterpret_cast<uint4*>(smem);
  __half* XH = reinterpret_cast<__half*>(smem + (size_t)RN * CHUNK_U4 * 16);
  __half* XL = XH + M.smem_x_halves;
  __half* DZ = XL + M.smem_x_halves;  // [buf 2][hi | lo][16][dz_ld]
  const int row0 = tile * TROWS, F = M.F, s_max = M.RS / TROWS, dz_ld = M.dz_ld;
  float* xs = reinterpret_cast<float*>(DZ + 4 * TROWS * dz_ld);
  float* wl = xs + TROWS * F;
  const int n_pass = fwd_only ? NL : 2 * NL - 1;
  const int n_chunks = fwd_only ? M.n_chunks_fwd : M.n_chunks;
  const unsigned c_base = ps.chunk;
  ps.chunk = c_base + (unsigned)n_chunks;
  const uint32_t sbase = smem_u32(smem);

  if (tid == 0) {
    fence_proxy_async();  // this CTA's earlier generic-proxy accesses of the ring's shared memory precede the bulk writes
    for (int c = 0; c < RN; ++c) ring_issue<RN>(M, P, bars, ring, c_base, c, n_chunks);
  }
  // X_0 = the (already normalised) input rows, hi | lo, plus the bias column of ones and zero padding
  {
    const TcLayer L0 = M.L[0];
    const int width = L0.KS * 16;
    for (int i = tid; i < TROWS * width; i += NTHREADS) {
      const int r = i / width, c = i - r * width;
      const bool valid = row0 + r < rows;
      float v = 0.f;
      if (c < F) {
        v = valid ? __ldg(xg + (size_t)(row0 + r) * F + c) : 0.f;
        xs[r * F + c] = v;
      } else if (c == F) {
        v = valid ? 1.f : 0.f;
      }
      __half hi, lo;
      split16(v, hi, lo);
      XH[L0.x_off + r * L0.ldx + c] = hi;
      XL[L0.x_off + r * L0.ldx + c] = lo;
    }
  }
  float loss = 0.f;
  for (int pass = 0; pass < n_pass; ++pass) {
    __syncthreads();  // the previous pass's epilogue (this pass's A operand) is complete and visible
    const TcPass& pd = M.pass[pass];  // (a reference into shared memory: a copy would keep two dozen registers live across the pass)
    if (prof && lane == 0) prof[16 + (pass * NWARPS + warp) * 4] = clock64();
    // the A operand of this pass goes to phase 2 feature-major (shared by the warps, a few 512-byte runs each)
    write_tile_major(sbase + pd.a_hi_off, sbase + pd.a_lo_off, pd.a_ld_b, pd.a_feat, pd.a_which ? P.zt_hi : P.xt_hi,
                     pd.a_which ? P.zt_lo : P.xt_lo, pd.a_gfeat0, s_max, tile);
    if (prof && lane == 0) prof[16 + (pass * NWARPS + warp) * 4 + 3] = clock64();
    const int nj = uni(pd.NT > warp ? (pd.NT - warp + NWARPS - 1) / NWARPS : 0);
    if (nj >= 4) run_pass<4, DBN>(M, P, bars, ring, sbase, pd, c_base, n_chunks, warp, lane, row0, sc, tile, smem, xs, loss, prof, pass);
    else if (nj == 3) run_pass<3, DBN>(M, P, bars, ring, sbase, pd, c_base, n_chunks, warp, lane, row0, sc, tile, smem, xs, loss, prof, pass);
    else if (nj == 2) run_pass<2, DBN>(M, P, bars, ring, sbase, pd, c_base, n_chunks, warp, lane, row0, sc, tile, smem, xs, loss, prof, pass);
    else if (nj == 1) run_pass<1, DBN>(M, P, bars, ring, sbase, pd, c_base, n_chunks, warp, lane, row0, sc, tile, smem, xs, loss, prof, pass);
    else run_pass<0, DBN>(M, P, bars, ring, sbase, pd, c_base, n_chunks, warp, lane, row0, sc, tile, smem, xs, loss, prof, pass);
    if (DBN && (pd.bn >= 0 || pd.kind == 1)) {
      // BatchNorm epilogue: the pass published this tile's column sums; reduce over the batch, then apply
#define BB_BY_NJ(call)            \
  do {                            \
    if (nj >= 4) { constexpr int NJ = 4; call; }       \
    else if (nj == 3) { constexpr int NJ = 3; call; }  \
    else if (nj == 2) { constexpr int NJ = 2; call; }  \
    else if (nj == 1) { constexpr int NJ = 1; call; }  \
  } while (0)
      const Own o{warp, lane >> 2, lane & 3, row0};
      const uint32_t o_hi = sbase + pd.o_hi_off, o_lo = sbase + pd.o_lo_off, o_ld = pd.o_ld_b;
      int bwd_bi = -1, bwd_pt = 0;
      bool leaky = false;
      if (pd.kind != 2) {
        const int bi = pd.kind == 1 ? 3 : pd.bn;
        if (P.train) bn_reduce<true>(M, P, sc, bi, bi, tile, smem);
        else bn_eval_stats(M, P, bi, smem);
        if (pd.kind == 0) {
          BB_BY_NJ(bn_apply_fwd<NJ>(M, sc, bi, o, o_hi, o_lo, o_ld, smem));
        } else {
          BB_BY_NJ(bn_loss<NJ>(M, P, sc, tile, o, o_hi, o_lo, o_ld, xs, loss, smem));
          if (!fwd_only) { bwd_bi = 3; bwd_pt = 4; }
        }
      } else {
        bwd_bi = pd.bn; bwd_pt = 7 - pd.bn; leaky = true;
      }
      if (bwd_bi >= 0) {
        bn_reduce<false>(M, P, sc, bwd_bi, bwd_pt, tile, smem);
        BB_BY_NJ(bn_apply_bwd<NJ>(M, sc, bwd_bi, leaky, o, o_hi, o_lo, o_ld, smem));
      }
#undef BB_BY_NJ
      if (prof && lane == 0) prof[16 + (pass * NWARPS + warp) * 4 + 2] = clock64();
    }
  }
  __syncthreads();
  if (!fwd_only) {  // dZ_0 (ping-pong buffer (7 - 0) & 1)
    const uint32_t z0 = smem_u32(DZ + ((NL - 1) & 1) * (2 * TROWS * dz_ld));
    write_tile_major(z0, z0 + TROWS * dz_ld * 2, dz_ld * 2, M.L[0].KSb * 16, P.zt_hi, P.zt_lo, M.L[0].zt_off, s_max, tile);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, off);
  if (lane == 0) wl[warp] = loss;
  __syncthreads();
  if (tid == 0) {
    float s = 0.f;
    for (int w = 0; w < NWARPS; ++w) s += wl[w];
    P.loss_part[loss_slot * M.max_tiles + tile] = s;
  }
  __syncthreads();  // shared memory is reused by the next tile / phase
}

// ------------------------------------------------------------------------------------------------ phase 2
// one 32 x 32 block of dW_l = dZ_l^T [X_l | 1] over all rows of the batch, then (exchange,) Adam and re-packing
__device__ __forceinline__ void phase2_item(const TcModel& M, const TcPtrs& P, const int item_idx, const int rows, const int flags,
                            const float lr_bc1, const float inv_sqrt_bc2, const float beta1, const float beta2, const float eps,
                            const unsigned step_tag, const int parity, unsigned char* smem, SmemBars* bars, PipeState& ps,
                            long long* prof) {
  const int tid = threadIdx.x, warp = uni(tid >> 5), lane = tid & 31, g = lane >> 2, t = lane & 3;
  const bool pk = prof && tid == 0;
  if (pk) prof[900] = clock64();
  TcItem it = P.items[item_idx];
  it.l = uni(it.l); it.n0 = uni(it.n0); it.k0 = uni(it.k0);
  TcLayer L = M.L[it.l];
  L.N = uni(L.N); L.K = uni(L.K);
  __syncthreads();  // the previous item's tile staging has been read
  unsigned char* panel = smem;  // [4: dZ hi, dZ lo, X hi, X lo][32 tiles][32 features][32 bytes]
  float* wtile = reinterpret_cast<float*>(smem + 4 * P2_PANEL);  // [32][33] updated weights of the block
  const int mt = warp & 1, nt = (warp >> 1) & 3;
  const int n_base = it.n0 + mt * 16, k_base = it.k0 + nt * 8;
  const bool mma_warp = true;
  const bool active = uni(mma_warp && n_base < L.N && k_base <= L.K) != 0;
  const int ks_total = (rows + TROWS - 1) / TROWS;
  const int s_max = M.RS / TROWS;
  float hh[4] = {0.f, 0.f, 0.f, 0.f}, cross[4] = {0.f, 0.f, 0.f, 0.f}, cross2[4] = {0.f, 0.f, 0.f, 0.f};
  // ldmatrix.trans lane addresses inside one tile's 1 KB [row 16][4 swizzled 16-byte pieces] slab: the stored 8 x 8
  // blocks are [batch row][feature]; transposed they are the (feature x row) A and (row x feature) B fragments
  const int ka = (lane & 7) + ((lane >> 4) & 1) * 8, pa = mt * 2 + ((lane >> 3) & 1);
  const int kb = (lane & 7) + ((lane >> 3) & 1) * 8;
  const uint32_t panel_b = smem_u32(smem);
  const uint32_t a_off = panel_b + ka * 64 + ((pa ^ ((ka >> 1) & 3)) * 16);
  const uint32_t b_off = panel_b + kb * 64 + ((nt ^ ((kb >> 1) & 3)) * 16);
  // Adam state of this thread's four weights: fetched now, used after the contraction (4 dependent L2 round trips
  // otherwise: measured 5.7 k cycles for the update)
  int a_idx[4];
  float a_m[4], a_v[4], a_p[4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = n_base + g + 8 * h, k = k_base + 2 * t + e, q = 2 * h + e;
      a_idx[q] = (active && n < L.N && k <= L.K) ? (k < L.K ? L.w_off + n * L.K + k : L.b_off + n) : -1;
      if (a_idx[q] >= 0 && (flags & TC_ADAM)) { a_m[q] = P.m[a_idx[q]]; a_v[q] = P.v[a_idx[q]]; a_p[q] = P.params[a_idx[q]]; }
    }
  const size_t zblk = (size_t)((L.zt_off + it.n0) >> 5) * s_max, xblk = (size_t)((L.xt_off + it.k0) >> 5) * s_max;

  for (int s0 = 0; s0 < ks_total; s0 += P2_ROWS / TROWS) {
    __syncthreads();  // the previous panel pass (or phase) is done with this shared memory
    if (tid == 0) {
      fence_proxy_async();
#pragma unroll
      for (int grp = 0; grp < 4; ++grp) {
        const int sa = s0 + grp * 8;
        const int nt_g = ks_total - sa < 8 ? ks_total - sa : 8;
        if (nt_g > 0) {
          const uint32_t bytes = (uint32_t)nt_g * 1024u;
          mbar_expect_tx(&bars->grp[grp], 4u * bytes);
          bulk_g2s(panel + 0 * P2_PANEL + grp * 8192, reinterpret_cast<const unsigned char*>(P.zt_hi) + (zblk + sa) * 1024, bytes, &bars->grp[grp]);
          bulk_g2s(panel + 1 * P2_PANEL + grp * 8192, reinterpret_cast<const unsigned char*>(P.zt_lo) + (zblk + sa) * 1024, bytes, &bars->grp[grp]);
          bulk_g2s(panel + 2 * P2_PANEL + grp * 8192, reinterpret_cast<const unsigned char*>(P.xt_hi) + (xblk + sa) * 1024, bytes, &bars->grp[grp]);
          bulk_g2s(panel + 3 * P2_PANEL + grp * 8192, reinterpret_cast<const unsigned char*>(P.xt_lo) + (xblk + sa) * 1024, bytes, &bars->grp[grp]);
        } else {
          mbar_expect_tx(&bars->grp[grp], 0u);  // keeps the phase of every group barrier in step
        }
      }
    }
#pragma unroll
    for (int grp = 0; grp < 4; ++grp) {
      mbar_wait(&bars->grp[grp], ps.pass & 1u);
      if (pk) prof[901 + 2 * grp] = clock64();
      if (active) {
        const int n_ks = uni(ks_total - (s0 + grp * 8) < 8 ? ks_total - (s0 + grp * 8) : 8);
        struct Frag { uint32_t ah[4], al[4], bh0, bh1, bl0, bl1; };
        auto fetch = [&](Frag& f, const int ksl) {
          const uint32_t slab = grp * 8192 + ksl * 1024;
          ldsm_x4t(f.ah, slab + 0 * P2_PANEL + a_off);
          ldsm_x4t(f.al, slab + 1 * P2_PANEL + a_off);
          ldsm_x2t(f.bh0, f.bh1, slab + 2 * P2_PANEL + b_off);
          ldsm_x2t(f.bl0, f.bl1, slab + 3 * P2_PANEL + b_off);
        };
        Frag f0, f1;
        fetch(f0, 0);
        for (int ksl = 0; ksl < n_ks; ksl += 2) {
          const bool more = ksl + 1 < n_ks;
          if (more) fetch(f1, ksl + 1);
          float t0[4] = {0.f, 0.f, 0.f, 0.f}, t1[4] = {0.f, 0.f, 0.f, 0.f};
          mma16816(t0, f0.ah, f0.bh0, f0.bh1);
          mma16816(cross, f0.ah, f0.bl0, f0.bl1);
          if (more) {
            mma16816(t1, f1.ah, f1.bh0, f1.bh1);
            mma16816(cross2, f1.ah, f1.bl0, f1.bl1);
          }
          mma16816(cross, f0.al, f0.bh0, f0.bh1);
          if (more) mma16816(cross2, f1.al, f1.bh0, f1.bh1);
#pragma unroll
          for (int i = 0; i < 4; ++i) hh[i] += t0[i];
          if (more) {
#pragma unroll
            for (int i = 0; i < 4; ++i) hh[i] += t1[i];
            if (ksl + 2 < n_ks) fetch(f0, ksl + 2);
          }
        }
      }
      if (pk) prof[902 + 2 * grp] = clock64();
    }
    ps.pass += 1;
  }
  if (pk) prof[910] = clock64();
  float v[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = hh[i] + (cross[i] + cross2[i]) * LO_INV;
  if (P.world > 1 && (flags & TC_XCHG)) {
    // Fused all-reduce (SUM) of this tile over NVLink peer memory.  The thread's four values travel as two 16-byte
    // {value, tag, value, tag} packets (each 8-byte half carries the step tag: the receiver spins on the data itself, so
    // there is no fence, no flag and no CTA barrier on the path - one NVLink store latency).  Relaxed sys-scope accesses,
    // not volatile ones: the stores to all peers leave back to back and the loads of a group of peers are in flight
    // together (ptxas completes a volatile access before it issues the next one: 28 dependent round trips per thread at
    // 8 GPUs).  Slots are double-buffered by step parity; a slot is rewritten two steps later, which the sender can only
    // reach after the receiver has consumed it (it needs the receiver's packets of the step in between).  The sum runs in
    // rank order on every rank: replicas stay bit-identical.
    // [item][packet 0 | 1][thread]: the 32 lanes of a store instruction write 512 contiguous bytes = whole 128-byte lines
    // on the wire
    const size_t slot = (size_t)item_idx * 2 * NTHREADS + tid;                        // uint4 index inside one rank's block
    const size_t per_rank = (size_t)M.n_items * NTHREADS * 2;
    const size_t mine = ((size_t)parity * P.world + P.rank) * per_rank + slot;
    for (int r = 0; r < P.world; ++r) {
      if (r == P.rank) continue;
      uint4* dst = reinterpret_cast<uint4*>(P.xchg[r]) + mine;
      st_pkt_sys(dst, v[0], v[1], step_tag);
      st_pkt_sys(dst + NTHREADS, v[2], v[3], step_tag);
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    const uint4* inbox = reinterpret_cast<const uint4*>(P.xchg[P.rank]) + (size_t)parity * P.world * per_rank + slot;
    for (int r0 = 0; r0 < P.world; r0 += 4) {
      uint4 q[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (r0 + u < P.world && r0 + u != P.rank) {
          q[u][0] = ld_pkt_sys(inbox + (size_t)(r0 + u) * per_rank);
          q[u][1] = ld_pkt_sys(inbox + (size_t)(r0 + u) * per_rank + NTHREADS);
        }
      bool missing;
      unsigned spins = 0;
      do {
        missing = false;
        spin_guard(spins);
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (r0 + u < P.world && r0 + u != P.rank) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
              if (q[u][h].y != step_tag || q[u][h].w != step_tag) {
                q[u][h] = ld_pkt_sys(inbox + (size_t)(r0 + u) * per_rank + h * NTHREADS);
                missing = true;
              }
          }
      } while (missing);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r0 + u >= P.world) continue;
        if (r0 + u == P.rank) {
#pragma unroll
          for (int i = 0; i < 4; ++i) s[i] += v[i];
        } else {
          s[0] += __uint_as_float(q[u][0].x);
          s[1] += __uint_as_float(q[u][0].z);
          s[2] += __uint_as_float(q[u][1].x);
          s[3] += __uint_as_float(q[u][1].z);
        }
      }
    }
    v[0] = s[0]; v[1] = s[1]; v[2] = s[2]; v[3] = s[3];
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = n_base + g + 8 * h;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = k_base + 2 * t + e;
      float p = 0.f;
      const int q = 2 * h + e, idx = a_idx[q];
      (void)n; (void)k;
      if (idx >= 0) {
        const float gv = v[q];
        if (flags & TC_GRADS) P.grads[idx] = gv;
        if (flags & TC_ADAM) {
          float mm, vv;
          adam_math(a_m[q], a_v[q], a_p[q], gv, lr_bc1, inv_sqrt_bc2, beta1, beta2, eps, mm, vv, p);
          P.m[idx] = mm; P.v[idx] = vv; P.params[idx] = p;
        }
      }
      if (mma_warp) wtile[(mt * 16 + g + 8 * h) * 33 + nt * 8 + 2 * t + e] = p;
    }
  }
  if (pk) prof[911] = clock64();
  if (flags & TC_ADAM) {
    // the block's fragments of both packed images, one whole uint4 (hi.r0, hi.r1, lo.r0, lo.r1) per thread: 512-byte runs
    __syncthreads();
    const int ksl = (tid >> 7) & 1, ntl = (tid >> 5) & 3, gq = lane >> 2, tq = lane & 3;
    uint4* img = const_cast<uint4*>(P.img);
    if (mma_warp) {
      const int ks = (it.k0 >> 4) + ksl, ntile = (it.n0 >> 3) + ntl;
      if (ks < L.KS && ntile < L.NTf) {
        const float* w = wtile + (ntl * 8 + gq) * 33 + ksl * 16 + 2 * tq;
        uint4 o;
        split16x2(w[0], w[1], o.x, o.z);
        split16x2(w[8], w[9], o.y, o.w);
        img[(size_t)L.imgf_off + ((size_t)ks * L.NTf + ntile) * 32 + lane] = o;
      }
    }
    if (mma_warp && it.l >= 1) {
      const int ks = (it.n0 >> 4) + ksl, ntile = (it.k0 >> 3) + ntl;
      if (ks < L.KSb && ntile < L.NTb) {
        const int kk = ntl * 8 + gq;
        const bool w_col = it.k0 + kk < L.K;  // the bias column is not part of the transposed image
        const float* w = wtile + (ksl * 16 + 2 * tq) * 33 + kk;
        uint4 o;
        split16x2(w_col ? w[0] : 0.f, w_col ? w[33] : 0.f, o.x, o.z);
        split16x2(w_col ? w[8 * 33] : 0.f, w_col ? w[9 * 33] : 0.f, o.y, o.w);
        img[(size_t)L.imgb_off + ((size_t)ks * L.NTb + ntile) * 32 + lane] = o;
      }
    }
  }
}

// batch loss = sum of the tile partials in tile order.  Data parallel: this rank's share; the caller sums the epoch
// losses of the ranks (one reduction per epoch instead of one exchange per step)
__device__ void finish_loss(const TcModel& M, const TcPtrs& P, const int n_tiles, const int loss_slot, const int flags,
                            const unsigned step_tag, const int parity, double* loss_accum) {
  float s = 0.f;
  for (int i = 0; i < n_tiles; ++i) s += __ldcg(P.loss_part + loss_slot * M.max_tiles + i);
  if (flags & (TC_GRADS | TC_P1)) P.grads[M.n_params] = s;
  if ((flags & (TC_ADAM | TC_FWD_ONLY)) && loss_accum) *loss_accum += (double)s;
  if (!(s == s) || fabsf(s) > 3.0e38f) *P.flag = 1;
}

template <bool DBN>
__global__ void __launch_bounds__(NTHREADS, 1)
tc_train_kernel(const __grid_constant__ TcModel Mparam, const __grid_constant__ TcPtrs P, const float* __restrict__ x,
                const long long n_rows, const int batch, const int n_steps, const int flags, const float beta1,
                const float beta2, const float eps, double* loss_accum, const int smem_bars_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  // The layer / chunk tables are indexed with run-time values all over the step.  In the kernel-parameter bank that is
  // an indexed LDC per field through the small constant cache (measured: 36 k of a 68 k-cycle phase 1 with the MMAs
  // removed); a shared-memory copy makes them ordinary LDS.
  __shared__ TcModel M;
  for (int i = threadIdx.x; i < (int)(sizeof(TcModel) / 4); i += NTHREADS)
    reinterpret_cast<int*>(&M)[i] = reinterpret_cast<const int*>(&Mparam)[i];
  SmemBars* bars = reinterpret_cast<SmemBars*>(smem + smem_bars_off);
  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(&bars->full[i], 1); bars->drained[i] = 0u; }
    for (int i = 0; i < 4; ++i) mbar_init(&bars->grp[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  PipeState ps = {0u, 0u};
  unsigned target = 0;
  // per-step facts live in shared memory, not in registers: the layer passes run at the register cap, and with this
  // kernel's shared-memory footprint only ~28 KB of L1 would back spills
  __shared__ StepCtx sc;
  for (int step = 0; step < n_steps; ++step) {
    long long r_begin = (long long)step * batch;
    int rows = (int)(n_rows - r_begin < batch ? n_rows - r_begin : batch);
    const int B = rows;
    int row_base = 0;
    if (P.dp_slice) {
      // data parallel: `batch` is the GLOBAL batch (the reference's batch_size) of the full table every rank holds; this
      // rank takes its contiguous share of it, the first (rows % world) ranks one row more (sharded.row_range)
      const int base = rows / P.world, extra = rows - base * P.world;
      row_base = P.rank * base + (P.rank < extra ? P.rank : extra);
      r_begin += row_base;
      rows = base + (P.rank < extra ? 1 : 0);
    }
    const int n_tiles = (rows + TROWS - 1) / TROWS;
    const int slot = step & 1;
    const unsigned tag = P.xbase + (unsigned)step + 1u;
    long long* prof = (P.prof && blockIdx.x == 0 && step == P.prof_step) ? P.prof : nullptr;
    if (threadIdx.x == 0) {
      sc.rows = rows; sc.n_tiles = n_tiles; sc.row_base = row_base; sc.B = B; sc.tag = tag;
      sc.dstep = P.drop_step + (unsigned long long)step;
      sc.fwd_only = (flags & TC_FWD_ONLY) != 0;
      sc.prof = prof;
      if (DBN) {
        sc.inv_rows = 1.f / (float)(rows > 0 ? rows : 1);
        sc.inv_B = 1.f / (float)(B > 0 ? B : 1);
        sc.inv_Bm1 = 1.f / (float)(B > 1 ? B - 1 : 1);
        for (int i = 0; i < 4; ++i) sc.per[i] = (n_tiles + M.bn_slices[i] - 1) / M.bn_slices[i];
      }
    }
    __syncthreads();
    if (prof && threadIdx.x == 0) prof[0] = clock64();
    if (flags & (TC_P1 | TC_FWD_ONLY)) {
      // (AE_Dropout_BN in train mode: the tiles of a batch meet at the BatchNorm reduction points, so every tile needs its
      // own co-resident CTA; the host refuses batches of more than gridDim.x tiles)
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        phase1_tile<DBN>(M, P, x + (size_t)r_begin * M.F, sc, tile, slot, smem, bars, ps, prof);
      if (prof && threadIdx.x == 0) prof[1] = clock64();
      grid_barrier(P.bar, target);
      if (prof && threadIdx.x == 0) prof[2] = clock64();
    }
    if (flags & TC_DW) {
      const float2 sh = P.stephyper[step];
      for (int item = blockIdx.x; item < M.n_items; item += gridDim.x)
        phase2_item(M, P, item, rows, flags, sh.x, sh.y, beta1, beta2, eps, tag, (int)(tag & 1u), smem, bars, ps, prof);
      if (DBN && (flags & TC_ADAM) && blockIdx.x == gridDim.x - 1) {
        // gamma / beta of the four BatchNorms: their gradients are batch totals every CTA held at the reduction points
        // (CTA 0 stored them, already summed over the ranks); one CTA applies Adam
        for (int i = M.n_linear + threadIdx.x; i < M.n_params; i += NTHREADS) {
          float p;
          adam_one(P, i, __ldcg(P.grads + i), sh.x, sh.y, beta1, beta2, eps, p);
        }
      }
      if (prof && threadIdx.x == 0) prof[912] = clock64();
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0 && (flags & (TC_P1 | TC_FWD_ONLY)))
      finish_loss(M, P, n_tiles, slot, flags, tag, (int)(tag & 1u), loss_accum);
    if (prof && threadIdx.x == 0) prof[3] = clock64();
    if ((flags & TC_DW) && step + 1 < n_steps) grid_barrier(P.bar, target);
    if (prof && threadIdx.x == 0) prof[4] = clock64();
  }
}

// Adam (or packing only) over the flat parameter vector: phase 2 of the step API after an external all-reduce, and
// the (re)build of the packed images from the fp32 parameters
__global__ void __launch_bounds__(256)
tc_flat_kernel(const __grid_constant__ TcModel M, const __grid_constant__ TcPtrs P, const int do_adam, const float lr_bc1,
               const float inv_sqrt_bc2, const float beta1, const float beta2, const float eps, double* loss_accum) {
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx < M.n_params) {
    int l = 0, n = 0, k = 0;
#pragma unroll
    for (int q = 0; q < NL; ++q) {
      const TcLayer& L = M.L[q];
      if (idx >= L.w_off && idx < L.b_off) { l = q; n = (idx - L.w_off) / L.K; k = idx - L.w_off - n * L.K; }
      else if (idx >= L.b_off && idx < L.b_off + L.N) { l = q; n = idx - L.b_off; k = L.K; }
    }
    float p = P.params[idx];
    if (do_adam) adam_one(P, idx, P.grads[idx], lr_bc1, inv_sqrt_bc2, beta1, beta2, eps, p);
    if (idx < M.n_linear) store_packed(M, P, l, n, k, p);  // (BatchNorm gamma / beta have no packed image)
  }
  if (do_adam && idx == 0 && loss_accum) *loss_accum += (double)P.grads[M.n_params];
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
struct TcTrainer {
  bb_ctx* ctx = nullptr;
  TcModel M;
  TcPtrs P;
  int max_batch = 0, grid = 0, bars_off = 0;
  size_t smem = 0, img_u4 = 0, xt_bytes = 0, zt_bytes = 0;
  uint4* img = nullptr;
  __half *xt_hi = nullptr, *xt_lo = nullptr, *zt_hi = nullptr, *zt_lo = nullptr;
  float* loss_part = nullptr;
  TcItem* items = nullptr;
  unsigned* bar = nullptr;
  float2* stephyper = nullptr;
  int stephyper_cap = 0;
  int* flag = nullptr;
  long long* prof = nullptr;
  std::vector<float2> host_hyper;
  unsigned xbase = 0;
  // data parallel: one cudaMalloc'd block [flags | exchange buffers] per rank, mapped into every peer by CUDA IPC
  void* dp_mem = nullptr;
  void* dp_peer[MAX_WORLD] = {};
  size_t dp_flag_bytes = 0;
  int xt_feat[NL], zt_feat[NL];
  uint4* bn_part = nullptr;      // AE_Dropout_BN: per-tile BatchNorm packets
  uint4* bn_fin = nullptr;       // ... and the final packet of every column
  size_t dp_bnx_off = 0;         // ... and where the per-rank packets start inside the data-parallel block
  int ring = RING;
};

static int round_up(int a, int b) { return (a + b - 1) / b * b; }
static void dp_release(TcTrainer* t);

int bb_tc_train_create(bb_ctx* ctx, const int* dims, const int* acts, int max_batch, float* params, float* m, float* v,
                       float* grads, int kind, const int* bn_g_off, const int* bn_b_off, int n_params_total, TcTrainer** out) {
  if (!ctx || !dims || !acts || !out || max_batch < 1 || kind < 0 || kind > 1) return BB_ERR_INVALID;
  if (kind == 1 && (!bn_g_off || !bn_b_off)) return BB_ERR_INVALID;
  TcTrainer* t = new (std::nothrow) TcTrainer();
  if (!t) return BB_ERR_NOMEM;
  t->ctx = ctx;
  t->max_batch = max_batch;
  TcModel& M = t->M;
  memset(&M, 0, sizeof(M));
  memset(&t->P, 0, sizeof(t->P));
  M.F = dims[0];
  M.kind = kind;
  t->ring = kind == 1 ? RING_DBN : RING;
  M.RS = round_up(max_batch, P2_ROWS);
  M.max_tiles = (max_batch + TROWS - 1) / TROWS;
  int p = 0, xs = 0, xt = 0, zt = 0, dz_ld = 0;
  size_t img = 0;
  for (int l = 0; l < NL; ++l) {
    TcLayer& L = M.L[l];
    L.K = dims[l]; L.N = dims[l + 1]; L.act = acts[l];
    L.KS = (L.K + 1 + 15) / 16;
    L.KSb = (L.N + 15) / 16;
    L.w_off = p; p += L.K * L.N;
    L.b_off = p; p += L.N;
    L.ldx = round_up(L.KS * 16, 32) + 8;  // whole 32-feature blocks (write_tile_major), +8: ldmatrix rows hit distinct banks
    L.x_off = xs; xs += TROWS * L.ldx;
    t->xt_feat[l] = round_up(L.K + 1, P2_BLK);
    t->zt_feat[l] = round_up(L.N, P2_BLK);
    if (t->xt_feat[l] < L.KS * 16) t->xt_feat[l] = round_up(L.KS * 16, P2_BLK);
    if (t->zt_feat[l] < L.KSb * 16) t->zt_feat[l] = round_up(L.KSb * 16, P2_BLK);
    L.xt_off = xt; xt += t->xt_feat[l];
    L.zt_off = zt; zt += t->zt_feat[l];
    if (round_up(L.KSb * 16, 32) + 8 > dz_ld) dz_ld = round_up(L.KSb * 16, 32) + 8;
  }
  for (int l = 0; l < NL; ++l) {
    TcLayer& L = M.L[l];
    L.NTf = l + 1 < NL ? 2 * M.L[l + 1].KS : 2 * L.KSb;   // the consumer's padded contraction width / 8
    L.NTb = l >= 1 ? 2 * M.L[l - 1].KSb : 0;
    L.imgf_off = (int)img; img += (size_t)L.KS * L.NTf * 32;
    L.imgb_off = (int)img; img += (size_t)L.KSb * L.NTb * 32;
    if (L.NTf > 4 * NWARPS || L.NTb > 4 * NWARPS || L.NTf > CHUNK_U4 / 32 || L.NTb > CHUNK_U4 / 32) { delete t; return BB_ERR_UNSUPPORTED; }
  }
  M.n_linear = p;
  M.n_params = kind == 1 ? n_params_total : p;
  M.smem_x_halves = xs; M.dz_ld = dz_ld;
  if (kind == 1) {
    // dropout keep probabilities 1 - p of models.py:263-275, thresholds exactly as train_dbn_kernel draws them
    const float keep_p[4] = {0.5f, 0.6f, 0.7f, 0.8f};
    for (int i = 0; i < 4; ++i) {
      M.keep_thr[i] = (unsigned)(keep_p[i] * 4294967296.0);
      M.keep_scale[i] = 1.f / keep_p[i];
      M.bn_g_off[i] = bn_g_off[i]; M.bn_b_off[i] = bn_b_off[i];
      M.bn_n[i] = dims[5 + i];
      M.bn_slices[i] = NTHREADS / dims[5 + i] > 0 ? NTHREADS / dims[5 + i] : 1;  // threads beyond the columns take slices of the tile range
      M.bn_f_off[i] = M.bn_f_total;
      M.bn_f_total += dims[5 + i];
      if (dims[5 + i] > BN_PW || dims[5 + i] > NTHREADS) { delete t; return BB_ERR_UNSUPPORTED; }
    }
    if (M.n_params < p + 2 * M.bn_f_total) { delete t; return BB_ERR_INVALID; }
  }
  // weight stream: per pass, groups of k-steps that fit one ring stage
  const int ring_b = t->ring * CHUNK_U4 * 16, xl_base = ring_b + xs * 2, dz_base = xl_base + xs * 2;
  const int dzbuf_b = 2 * TROWS * dz_ld * 2, dzlo_b = TROWS * dz_ld * 2;
  int nc = 0;
  for (int pass = 0; pass < 2 * NL - 1; ++pass) {
    const bool bwd = pass >= NL;
    const int l = bwd ? 2 * NL - 1 - pass : pass;
    const TcLayer& L = M.L[l];
    const int KS = bwd ? L.KSb : L.KS, NT = bwd ? L.NTb : L.NTf, base = bwd ? L.imgb_off : L.imgf_off;
    const int per_max = (CHUNK_U4 / 32) / NT;
    const int n_ch = (KS + per_max - 1) / per_max;
    // A operand of the pass: X_l (forward) or dZ_l in ping-pong buffer (7 - l) & 1 (backward)
    const int a_hi = bwd ? dz_base + ((NL - 1 - l) & 1) * dzbuf_b : ring_b + L.x_off * 2;
    const int a_lo = bwd ? a_hi + dzlo_b : xl_base + L.x_off * 2;
    const int lda_b = (bwd ? dz_ld : L.ldx) * 2;
    int ks = 0;
    for (int c = 0; c < n_ch; ++c) {
      const int len = KS / n_ch + (c < KS % n_ch ? 1 : 0);
      if (nc >= MAX_CHUNKS) { delete t; return BB_ERR_UNSUPPORTED; }
      TcChunk& ch = M.chunk[nc++];
      ch.pass = (short)pass; ch.nk = (short)len; ch.NT = (short)NT;
      ch.flags = (short)((ks == 0 ? 1 : 0) | (ks + len == KS ? 2 : 0) | (bwd ? 4 : 0));
      ch.src_u4 = base + ks * NT * 32; ch.n_u4 = len * NT * 32;
      ch.a_hi_off = a_hi + ks * 32; ch.a_lo_off = a_lo + ks * 32; ch.lda_b = lda_b; ch.l = l;
      ks += len;
    }
    TcPass& ps = M.pass[pass];
    memset(&ps, 0, sizeof(ps));
    ps.N = L.N; ps.act = L.act; ps.NT = NT;
    ps.c0 = nc - n_ch; ps.c1 = nc;
    ps.a_hi_off = a_hi; ps.a_lo_off = a_lo; ps.a_ld_b = lda_b;
    ps.a_feat = (bwd ? L.KSb : L.KS) * 16; ps.a_gfeat0 = bwd ? L.zt_off : L.xt_off; ps.a_which = bwd ? 1 : 0;
    if (!bwd && l < NL - 1) {          // hidden forward layer: writes X_{l+1}
      ps.kind = 0;
      ps.o_hi_off = ring_b + M.L[l + 1].x_off * 2; ps.o_lo_off = xl_base + M.L[l + 1].x_off * 2; ps.o_ld_b = M.L[l + 1].ldx * 2;
    } else if (!bwd) {                 // last layer: loss and the seed gradient dZ_7 -> ping-pong buffer 0
      ps.kind = 1;
      ps.o_hi_off = dz_base; ps.o_lo_off = dz_base + dzlo_b; ps.o_ld_b = dz_ld * 2;
    } else {                           // backward pass of layer l: writes dZ_{l-1} -> buffer (8 - l) & 1
      ps.kind = 2;
      ps.o_hi_off = dz_base + ((NL - l) & 1) * dzbuf_b; ps.o_lo_off = ps.o_hi_off + dzlo_b; ps.o_ld_b = dz_ld * 2;
      ps.pact = M.L[l - 1].act;
      ps.x_hi_off = ring_b + L.x_off * 2; ps.x_lo_off = xl_base + L.x_off * 2; ps.x_ld_b = L.ldx * 2;
    }
    ps.drop = ps.bn = -1;
    if (kind == 1) {
      if (!bwd && l < 4) ps.drop = l;                 // encoder layer: Linear -> Dropout -> LeakyReLU
      if (!bwd && l >= 4 && l < NL - 1) ps.bn = l - 4;  // decoder layer: Linear -> LeakyReLU -> BatchNorm (the last one: kind 1)
      if (bwd && l - 1 < 4) ps.drop = l - 1;          // backward pass of layer l writes dZ_{l-1}
      if (bwd && l - 1 >= 4) ps.bn = l - 1 - 4;
    }
    if (pass == NL - 1) M.n_chunks_fwd = nc;
  }
  M.n_chunks = nc;
  std::vector<TcItem> items;
  for (int l = 0; l < NL; ++l)
    for (int n0 = 0; n0 < M.L[l].N; n0 += P2_BLK)
      for (int k0 = 0; k0 <= M.L[l].K; k0 += P2_BLK) items.push_back({l, n0, k0});
  M.n_items = (int)items.size();
  size_t smem1 = (size_t)t->ring * CHUNK_U4 * 16 + (size_t)2 * xs * 2 + (size_t)4 * TROWS * dz_ld * 2 + (size_t)TROWS * M.F * 4 + NWARPS * 4;
  if (kind == 1) {
    smem1 = (smem1 + 15) / 16 * 16;
    for (int i = 0; i < 4; ++i) {
      M.u_ld[i] = round_up(M.bn_n[i], 8);
      while (M.u_ld[i] % 32 != 8) M.u_ld[i] += 8;  // rows g = 0..3 of a float2 store land in distinct bank groups
      M.u_off[i] = (int)smem1;
      smem1 += (size_t)TROWS * M.u_ld[i] * 4;
    }
    M.stat_off = (int)smem1;
    smem1 += (size_t)(2 * M.bn_f_total + 4 * BN_TW) * 4;
    smem1 = (smem1 + 15) / 16 * 16;
    M.red_off = (int)smem1;
    smem1 += (size_t)NTHREADS * sizeof(float2);
  }
  const size_t smem2 = (size_t)4 * P2_PANEL + 32 * 33 * 4;
  t->bars_off = (int)(((smem1 > smem2 ? smem1 : smem2) + 127) / 128 * 128);
  t->smem = (size_t)t->bars_off + sizeof(SmemBars);
  if (t->smem > ctx->smem_optin) { delete t; return BB_ERR_UNSUPPORTED; }
  t->img_u4 = img;
  t->xt_bytes = (size_t)xt * M.RS * 2;
  t->zt_bytes = (size_t)zt * M.RS * 2;
  int rc = BB_OK;
  auto alloc = [&](void** ptr, size_t bytes) {
    if (rc == BB_OK) rc = (int)cudaMalloc(ptr, bytes);
    if (rc == BB_OK) rc = (int)cudaMemset(*ptr, 0, bytes);
  };
  alloc((void**)&t->img, img * sizeof(uint4));
  alloc((void**)&t->xt_hi, (size_t)xt * M.RS * 2);  // [feature / 32][tile][32 features][16 rows]: same size, see fm_u32_index
  alloc((void**)&t->xt_lo, (size_t)xt * M.RS * 2);
  alloc((void**)&t->zt_hi, (size_t)zt * M.RS * 2);
  alloc((void**)&t->zt_lo, (size_t)zt * M.RS * 2);
  alloc((void**)&t->loss_part, sizeof(float) * 2 * M.max_tiles);
  alloc((void**)&t->items, sizeof(TcItem) * items.size());
  alloc((void**)&t->bar, sizeof(unsigned));
  alloc((void**)&t->flag, sizeof(int));
  alloc((void**)&t->prof, sizeof(long long) * 1024);
  if (kind == 1) alloc((void**)&t->bn_part, sizeof(uint4) * 8 * (size_t)M.max_tiles * BN_PW);
  if (kind == 1) alloc((void**)&t->bn_fin, sizeof(uint4) * 8 * BN_PW);
  if (rc == BB_OK) rc = (int)cudaMemcpy(t->items, items.data(), sizeof(TcItem) * items.size(), cudaMemcpyHostToDevice);
  const void* kern = kind == 1 ? (const void*)tc_train_kernel<true> : (const void*)tc_train_kernel<false>;
  if (rc == BB_OK) rc = (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t->smem);
  int per_sm = 0;
  if (rc == BB_OK) rc = (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NTHREADS, t->smem);
  if (rc == BB_OK && per_sm < 1) rc = BB_ERR_UNSUPPORTED;
  if (rc != BB_OK) { bb_tc_train_destroy(t); return rc; }
  const int want = M.n_items > M.max_tiles ? M.n_items : M.max_tiles;
  t->grid = want < per_sm * ctx->sm_count ? want : per_sm * ctx->sm_count;
  if (const char* g = getenv("BALER_B200_TC_GRID")) {  // diagnostics: profile with fewer CTAs (any grid >= 1 is correct)
    const int v = atoi(g);
    if (v >= 1 && v < t->grid) t->grid = v;
  }
  // AE_Dropout_BN: the tiles of a batch wait for each other at the BatchNorm reduction points: one resident CTA per tile
  if (kind == 1 && M.max_tiles > t->grid) { bb_tc_train_destroy(t); return BB_ERR_UNSUPPORTED; }
  TcPtrs& P = t->P;
  P.params = params; P.m = m; P.v = v; P.grads = grads;
  P.img = t->img; P.xt_hi = t->xt_hi; P.xt_lo = t->xt_lo; P.zt_hi = t->zt_hi; P.zt_lo = t->zt_lo;
  P.loss_part = t->loss_part; P.items = t->items; P.bar = t->bar; P.flag = t->flag;
  P.rank = 0; P.world = 1;
  P.bn_part = t->bn_part;
  P.bn_fin = t->bn_fin;
  P.train = 1;
  *out = t;
  return BB_OK;
}

void bb_tc_train_destroy(TcTrainer* t) {
  if (!t) return;
  dp_release(t);
  void* ptrs[] = {t->img, t->xt_hi, t->xt_lo, t->zt_hi, t->zt_lo, t->loss_part, t->items, t->bar, t->flag, t->stephyper, t->prof,
                  t->bn_part, t->bn_fin};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete t;
}

int bb_tc_train_repack(TcTrainer* t, cudaStream_t s) {
  if (!t) return BB_ERR_INVALID;
  tc_flat_kernel<<<(t->M.n_params + 255) / 256, 256, 0, s>>>(t->M, t->P, 0, 0.f, 0.f, 0.f, 0.f, 0.f, nullptr);
  return (int)cudaGetLastError();
}

static void step_hyper(const TcHyper* h, long long step, float2* out) {
  const double bc1 = 1.0 - std::pow(h->b1d, (double)step);
  const double bc2 = 1.0 - std::pow(h->b2d, (double)step);
  out->x = (float)(h->lr / bc1);
  out->y = (float)(1.0 / std::sqrt(bc2));
}

int bb_tc_train_run(TcTrainer* t, const float* x, int64_t n_rows, int batch, int flags, const TcHyper* h, long long first_step,
                    double* loss_accum, int dp_slice, cudaStream_t s) {
  if (!t || !h || batch < 1 || n_rows < 0) return BB_ERR_INVALID;
  const int local_max = dp_slice && t->P.world > 1 ? (batch + t->P.world - 1) / t->P.world : batch;
  if (local_max > t->max_batch) return BB_ERR_INVALID;
  t->P.dp_slice = dp_slice && t->P.world > 1;
  if (t->P.dp_slice) flags |= TC_XCHG;
  const int n_steps = n_rows == 0 ? 1 : (int)((n_rows + batch - 1) / batch);
  if (n_steps > t->stephyper_cap) {
    if (t->stephyper) cudaFree(t->stephyper);
    t->stephyper = nullptr;
    t->stephyper_cap = 0;
    BB_CUDA(cudaMalloc((void**)&t->stephyper, sizeof(float2) * (size_t)n_steps));
    t->stephyper_cap = n_steps;
  }
  t->host_hyper.resize(n_steps);
  for (int i = 0; i < n_steps; ++i) step_hyper(h, first_step + i, &t->host_hyper[i]);
  BB_CUDA(cudaMemcpyAsync(t->stephyper, t->host_hyper.data(), sizeof(float2) * (size_t)n_steps, cudaMemcpyHostToDevice, s));
  BB_CUDA(cudaMemsetAsync(t->bar, 0, sizeof(unsigned), s));
  t->P.stephyper = t->stephyper;
  t->P.xbase = t->xbase;
  if (flags & TC_DW) t->xbase += (unsigned)n_steps;
  long long nr = n_rows;
  float b1 = h->beta1, b2 = h->beta2, eps = h->eps;
  int ns = n_steps, fl = flags, bt = batch, bo = t->bars_off;
  void* args[] = {(void*)&t->M, (void*)&t->P, (void*)&x, (void*)&nr, (void*)&bt, (void*)&ns, (void*)&fl,
                  (void*)&b1, (void*)&b2, (void*)&eps, (void*)&loss_accum, (void*)&bo};
  const void* kern = t->M.kind == 1 ? (const void*)tc_train_kernel<true> : (const void*)tc_train_kernel<false>;
  return (int)cudaLaunchCooperativeKernel(kern, dim3(t->grid), dim3(NTHREADS), args, t->smem, s);
}

int bb_tc_train_set_bn(TcTrainer* t, float* running_mean, float* running_var, long long* batches_tracked) {
  if (!t || t->M.kind != 1 || !running_mean || !running_var || !batches_tracked) return BB_ERR_INVALID;
  t->P.rm = running_mean; t->P.rv = running_var; t->P.nbt = batches_tracked;
  return BB_OK;
}

int bb_tc_train_set_dropout(TcTrainer* t, unsigned long long seed, const unsigned char* const* masks_dev) {
  if (!t || t->M.kind != 1) return BB_ERR_INVALID;
  t->P.seed = seed;
  for (int i = 0; i < 4; ++i) t->P.mask[i] = masks_dev ? masks_dev[i] : nullptr;
  return BB_OK;
}

void bb_tc_train_set_mode(TcTrainer* t, int train, unsigned long long drop_step) {
  if (!t) return;
  t->P.train = train;
  t->P.drop_step = drop_step;
}

int bb_tc_train_adam_flat(TcTrainer* t, const TcHyper* h, long long step, double* loss_accum, cudaStream_t s) {
  if (!t || !h) return BB_ERR_INVALID;
  float2 sh;
  step_hyper(h, step, &sh);
  tc_flat_kernel<<<(t->M.n_params + 255) / 256, 256, 0, s>>>(t->M, t->P, 1, sh.x, sh.y, h->beta1, h->beta2, h->eps, loss_accum);
  return (int)cudaGetLastError();
}

// host view of the feature-major scratch (see fm_u32_index): value = hi + lo / 2048
static size_t fm_half_index(int f_abs, int row, int s_max) {
  const int tile = row >> 4, rr = row & 15, piece = (f_abs & 31) >> 3;
  return ((((size_t)(f_abs >> 5) * s_max + tile) * 16 + rr) * 64 + (size_t)((piece ^ ((rr >> 1) & 3)) * 16)) / 2 + (f_abs & 7);
}

static int fm_download(TcTrainer* t, int which, std::vector<__half>& hi, std::vector<__half>& lo) {
  const size_t bytes = which == 0 ? t->xt_bytes : t->zt_bytes;
  hi.resize(bytes / 2); lo.resize(bytes / 2);
  BB_CUDA(cudaDeviceSynchronize());
  BB_CUDA(cudaMemcpy(hi.data(), which == 0 ? t->xt_hi : t->zt_hi, bytes, cudaMemcpyDeviceToHost));
  BB_CUDA(cudaMemcpy(lo.data(), which == 0 ? t->xt_lo : t->zt_lo, bytes, cudaMemcpyDeviceToHost));
  return BB_OK;
}

int bb_tc_train_activation_means(TcTrainer* t, int rows, double* out) {
  if (!t || !out || rows < 0 || rows > t->M.RS) return BB_ERR_INVALID;
  const TcModel& M = t->M;
  std::vector<__half> hi, lo;
  const int rc = fm_download(t, 0, hi, lo);
  if (rc != BB_OK) return rc;
  const int layers[6] = {1, 2, 3, 5, 6, 7};  // X_l = LeakyReLU(output of Linear l-1)
  for (int i = 0; i < 6; ++i) {
    const TcLayer& L = M.L[layers[i]];
    for (int j = 0; j < 200; ++j) {
      double s = NAN;
      if (j < L.K && rows > 0) {
        s = 0.0;
        for (int r = 0; r < rows; ++r) {
          const size_t q = fm_half_index(L.xt_off + j, r, M.RS / TROWS);
          s += (double)__half2float(hi[q]) + (double)__half2float(lo[q]) / 2048.0;
        }
        s /= rows;
      }
      out[i * 200 + j] = s;
    }
  }
  return BB_OK;
}

int bb_tc_train_debug_layer(TcTrainer* t, int which, int layer, int rows, float* out, int capacity) {
  if (!t || !out || layer < 0 || layer >= NL || rows < 0 || rows > t->M.RS) return BB_ERR_INVALID;
  const TcModel& M = t->M;
  const TcLayer& L = M.L[layer];
  const int n_feat = which == 0 ? L.K + 1 : L.N;
  if (capacity < n_feat * rows) return BB_ERR_INVALID;
  std::vector<__half> hi, lo;
  const int rc = fm_download(t, which, hi, lo);
  if (rc != BB_OK) return rc;
  const int off = which == 0 ? L.xt_off : L.zt_off;
  for (int f = 0; f < n_feat; ++f)
    for (int r = 0; r < rows; ++r) {
      const size_t q = fm_half_index(off + f, r, M.RS / TROWS);
      out[(size_t)f * rows + r] = __half2float(hi[q]) + __half2float(lo[q]) / 2048.f;
    }
  return n_feat;
}

int bb_tc_train_profile(TcTrainer* t, int step, long long* out_128) {
  if (!t) return BB_ERR_INVALID;
  if (out_128) {
    BB_CUDA(cudaDeviceSynchronize());
    BB_CUDA(cudaMemcpy(out_128, t->prof, sizeof(long long) * 1024, cudaMemcpyDeviceToHost));
  }
  t->P.prof = step >= 0 ? t->prof : nullptr;
  t->P.prof_step = step;
  return t->M.n_chunks;
}

int bb_tc_train_range_flag(TcTrainer* t, int reset, int* out) {
  if (!t || !out) return BB_ERR_INVALID;
  BB_CUDA(cudaDeviceSynchronize());
  BB_CUDA(cudaMemcpy(out, t->flag, sizeof(int), cudaMemcpyDeviceToHost));
  if (reset) BB_CUDA(cudaMemset(t->flag, 0, sizeof(int)));
  return BB_OK;
}

static size_t dp_flag_bytes(const TcTrainer*, int) { return 256; }  // reserved
static size_t dp_xchg_bytes(const TcTrainer* t, int world) {
  // [2 step parities][world source ranks][items][256 threads][4 {value, tag} packets of 8 bytes]
  return (size_t)2 * world * t->M.n_items * NTHREADS * 4 * sizeof(uint2);
}
static size_t dp_total_bytes(const TcTrainer* t, int world) {
  // ... and, AE_Dropout_BN, [2 step parities][8 reduction points][world source ranks][BN_PW] 16-byte packets
  return dp_flag_bytes(t, world) + dp_xchg_bytes(t, world) + (t->M.kind == 1 ? (size_t)2 * 8 * world * BN_PW * sizeof(uint4) : 0);
}

static void dp_release(TcTrainer* t) {
  for (int r = 0; r < MAX_WORLD; ++r) {
    if (t->dp_peer[r] && t->dp_peer[r] != t->dp_mem) cudaIpcCloseMemHandle(t->dp_peer[r]);
    t->dp_peer[r] = nullptr;
  }
  if (t->dp_mem) cudaFree(t->dp_mem);
  t->dp_mem = nullptr;
  t->P.world = 1; t->P.rank = 0;
}

int bb_tc_train_dp_export(TcTrainer* t, int world, unsigned char* handle_out_64) {
  if (!t || world < 2 || world > MAX_WORLD || !handle_out_64) return BB_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  dp_release(t);
  const size_t bytes = dp_total_bytes(t, world);
  BB_CUDA(cudaMalloc(&t->dp_mem, bytes));
  BB_CUDA(cudaMemset(t->dp_mem, 0, bytes));
  BB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  BB_CUDA(cudaIpcGetMemHandle(&h, t->dp_mem));
  memcpy(handle_out_64, &h, 64);
  t->dp_flag_bytes = dp_flag_bytes(t, world);
  return BB_OK;
}

int bb_tc_train_dp_connect(TcTrainer* t, int rank, int world, const unsigned char* handles) {
  if (!t || !t->dp_mem || !handles || world < 2 || world > MAX_WORLD || rank < 0 || rank >= world) return BB_ERR_INVALID;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { t->dp_peer[r] = t->dp_mem; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    BB_CUDA(cudaIpcOpenMemHandle(&t->dp_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  t->P.rank = rank;
  t->P.world = world;
  for (int r = 0; r < world; ++r) {
    t->P.xflag[r] = reinterpret_cast<unsigned*>(t->dp_peer[r]);
    t->P.xchg[r] = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(t->dp_peer[r]) + t->dp_flag_bytes);
    t->P.bnx[r] = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(t->dp_peer[r]) + t->dp_flag_bytes + dp_xchg_bytes(t, world));
  }
  return BB_OK;
}

int bb_tc_train_dp_world(const TcTrainer* t) { return t ? t->P.world : 1; }
