// Fused 4-layer dense chain on tcgen05 / TMEM, statically shaped for the AE family of the reference
// (models.py:116-157: n_features -> 200 -> 100 -> 50 -> z and back), sm_100a.  Same arithmetic and the same
// weight image as bb_chain_tc.cu (fp16 hi/lo split, three products per k-step, fp32 accumulation in TMEM,
// bias as an extra K column, in-place accumulator -> next-operand conversion); what differs is how it is driven:
//
//  * The step program is a compile-time table (template <ENC, KA, NL>), so every TMEM address and every
//    shared-memory matrix descriptor the MMA issuer needs is a constant or a uniform-register add.  The
//    table-driven kernel spent ~130 cycles of issue per MMA on shared-memory table reads and R2UR moves
//    while a 128 x 112 x 16 MMA occupies the tensor pipe for 56 (tools/mma_bench2.cu).
//  * Accumulators are drained in 16-column sub-chunks (= exactly one k-step of the next layer: 8 columns of
//    packed hi | 8 columns of packed lo), software-pipelined (the TMEM load of the next sub-chunk is in
//    flight while the current one is converted), each with its own mbarrier, so the next layer's first MMA
//    starts after one sub-chunk instead of after 32 columns.
//  * No CTA-wide or group-wide barrier in the tile loop.  Input tiles arrive by cp.async.bulk (issued two
//    tiles ahead by the MMA warp, completion on an mbarrier); every epilogue warp stores its own 32 output
//    rows with cp.async.bulk (shared -> global) straight from its slice of the output stage.
//
// Warp roles (576 threads): two independent tile pipelines, each 8 epilogue warps (TMEM lane quarter q =
// warp % 4, column half h) + one MMA-issuer warp.  TMEM per pipeline: X 112 | Y 112 | Z 32 columns.
#include <cmath>
#include <cstdlib>

#include "bb_common.cuh"

namespace {

constexpr int TILE = 128;
constexpr int EPI_GROUPS = 2;           // column groups per pipeline: group h drains sub-chunks h, h + EPI_GROUPS, ...
constexpr int EPI_WARPS = 4 * EPI_GROUPS;  // per pipeline: one warp per TMEM lane quarter and column group
constexpr int NPIPE = 2;
constexpr int NTHREADS = NPIPE * (EPI_WARPS + 1) * 32;
constexpr int REG_X = 0, REG_Y = 112, REG_Z = 224, PIPE_COLS = 256;
constexpr int MAX_SUB = 7;              // 16-column sub-chunks of the widest accumulator region (112)
constexpr int H1 = 208, H2 = 112, H3 = 64;   // padded hidden widths: 200 + bias slot, 100 + 1, 50 + 1

struct Tc4Params {
  const uint8_t* wimg;
  const float* in;
  float* out;
  int64_t n_rows;
  const float *pre_min, *pre_range, *post_min, *post_range;
  float c1[4], c2[4];   // per layer: y = c1 * acc + c2 * |acc|  (s = 2^-sw; leaky: 0.505 s, 0.495 s; relu: 0.5 s, 0.5 s; none: s, 0)
  int epi_col[4], epi_nsub[4];      // steps 0..3: accumulator region (pipeline-relative column) and its 16-column sub-chunks
  float epi_c1[4], epi_c2[4];       // the multipliers of the layer that step drains
  int in_dim, out_dim;
  int* flag;
  uint32_t* trace;      // optional: SM-clock timestamps of CTA 0 / pipeline trace_pipe (64 slots per tile, first 16 tiles)
  int trace_pipe;
  uint32_t smem_off;    // low 18 bits of the shared-window address of the dynamic shared memory (probed once per kernel, see launch_one)
  uint32_t tmem0;       // TMEM base address (0: the CTA owns all 512 columns); a parameter so that addresses are UR adds, not R2UR moves
  uint32_t* probe;      // probe launch: thread 0 writes that address here and the kernel returns
  int pipes;            // tile pipelines in use (2; 1 = experiment: the second pipeline idles)
  int in_f16, out_f16;  // opt-in float16 latent: the decoder's input / the encoder's output rows are fp16 (126 B / row)
};

// ------------------------------------------------------------------------------------------------ static program
struct SMma { int layer, a_col, a_w, ks0, ks_n, n0, n, d_col, acc, dep; };
struct SEpi { int layer, col, w, fin; };
struct SStep { int n_mma; SMma mma[2]; SEpi epi; };

template <bool ENC, int KA, int NL>
struct Prog {
  // padded K / N of the four layers
  static constexpr int Kp(int l) { return ENC ? (l == 0 ? KA : l == 1 ? H1 : l == 2 ? H2 : H3) : (l == 0 ? KA : l == 1 ? H3 : l == 2 ? H2 : H1); }
  static constexpr int Np(int l) { return l < 3 ? Kp(l + 1) : NL; }
  static constexpr uint32_t w_off(int l) {  // byte offset of the hi image of layer l (lo image follows it)
    uint32_t o = 0;
    for (int i = 0; i < l; ++i) o += 4u * Np(i) * Kp(i);
    return o;
  }
  static constexpr uint32_t W_BYTES = w_off(4);
  static constexpr int A1_COL = ENC ? REG_Z : REG_Y;
  static constexpr int N_STEPS = 5;
  static constexpr SStep step(int s) {
    if (ENC) {
      switch (s) {
        case 0: return {1, {{0, REG_Z, KA, 0, KA / 16, 0, 112, REG_X, 0, 0}, {}}, {0, REG_X, 112, 0}};
        case 1: return {2, {{1, REG_X, 112, 0, 7, 0, H2, REG_Y, 0, 1}, {0, REG_Z, KA, 0, KA / 16, 112, 96, REG_X, 0, 0}}, {0, REG_X, 96, 0}};
        case 2: return {1, {{1, REG_X, 96, 7, 6, 0, H2, REG_Y, 1, 1}, {}}, {1, REG_Y, H2, 0}};
        case 3: return {1, {{2, REG_Y, H2, 0, 7, 0, H3, REG_X, 0, 1}, {}}, {2, REG_X, H3, 0}};
        default: return {1, {{3, REG_X, H3, 0, 4, 0, NL, REG_Y, 0, 1}, {}}, {3, REG_Y, NL, 1}};
      }
    } else {
      switch (s) {
        case 0: return {1, {{0, REG_Y, KA, 0, KA / 16, 0, H3, REG_X, 0, 0}, {}}, {0, REG_X, H3, 0}};
        case 1: return {1, {{1, REG_X, H3, 0, 4, 0, H2, REG_Y, 0, 1}, {}}, {1, REG_Y, H2, 0}};
        case 2: return {1, {{2, REG_Y, H2, 0, 7, 0, 112, REG_X, 0, 1}, {}}, {2, REG_X, 112, 0}};
        case 3: return {2, {{3, REG_X, 112, 0, 7, 0, NL, REG_Z, 0, 1}, {2, REG_Y, H2, 0, 7, 112, 96, REG_X, 0, 0}}, {2, REG_X, 96, 0}};
        default: return {1, {{3, REG_X, 96, 7, 6, 0, NL, REG_Z, 1, 1}, {}}, {3, REG_Z, NL, 1}};
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  // a protocol bug must not hang the GPU: trap (-> launch error) after ~1 s of polling
  // back off between polls: hot polling by the ~12 warps that wait at any time took 40 % of the SM's issue slots
  // (measured 32 / 64 / 128 ns: 5.58 / 5.75 / 5.83 G rows/s encode)
  for (uint32_t spins = 0; !mbar_try(bar, parity); ++spins) {
    __nanosleep(128);
    if (spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  // try_wait blocks for a bounded, hardware-chosen time; the first poll usually succeeds on the hot path.  (The
  // suspend-time-hint form compiles to TRYWAIT + NANOSLEEP.SYNCS, whose wake-up cost ~150-300 cycles per wait.)
  if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#define TC4_OUT16(v) "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                     "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
#define TC4_IN16(v) "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), \
                    "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
// each thread receives / writes 16 consecutive 32-bit columns of its own TMEM lane (row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : TC4_OUT16(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
               TC4_IN16(v) : "memory");
}

// packed fp32x2 arithmetic (FMUL2 / FADD2 on sm_100): two elements per instruction
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float lo32(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi32(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// two fp32 values -> packed fp16 hi and lo words: hi = top 11 significant bits (truncated, so that x - hi is
// exact in fp32), lo = fp16_rn(x - hi)
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const float ah = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float bh = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  const uint64_t dl = sub2(pack2(a, b), pack2(ah, bh));
  const __half2 h = __floats2half2_rn(ah, bh);
  const __half2 l = __floats2half2_rn(lo32(dl), hi32(dl));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 16 accumulator columns -> scale + activation -> the next layer's k-step operand: 8 words hi | 8 words lo.
// The activation is y = c * v + d * |v| (leaky: c = 0.505 s, d = 0.495 s; relu: c = d = 0.5 s; none: c = s, d = 0;
// s = 2^-sw undoes the weight scaling): FMUL + FFMA on the FMA pipe instead of two multiplies and an FMNMX, because
// the ALU pipe (FMNMX, LOP3, F2FP: 2 cycles per warp instruction each, tools/alu_bench.cu) is what bounds this loop.
__device__ __forceinline__ void convert16(const uint32_t (&v)[16], const uint64_t c2x, const float d, uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float v0 = __uint_as_float(v[2 * j]), v1 = __uint_as_float(v[2 * j + 1]);
    const uint64_t t = mul2(pack2(v0, v1), c2x);
    split2(fmaf(fabsf(v0), d, lo32(t)), fmaf(fabsf(v1), d, hi32(t)), pk[j], pk[8 + j]);
  }
}

constexpr uint32_t B_DESC_HI = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1 (bit 46)
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  // c_format F32 (1) @4, a/b format F16 (0) @7/@10, a/b K-major, N>>3 @17, M>>4 @24
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}

// one k-step: hi*hi (accumulate flag), then hi*lo and lo*hi (always accumulate).  The warp is converged; branching
// on elect.sync lets ptxas emit bare warp-level UTCHMMA with uniform-register operands (predicating the MMAs inside
// the asm block instead makes it move every operand through R2UR under the elected lane's predicate).
__device__ __forceinline__ void mma_f16(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// FAST (BB_PREC_FAST16, opt-in, OUTSIDE the 1e-5 tolerance): only the hi * hi product, a third of the tensor work.
template <bool FAST>
__device__ __forceinline__ void issue_kstep(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            uint32_t idesc, uint32_t acc) {
  if (elect_one()) {
    const uint64_t bh = (uint64_t)b_hi | ((uint64_t)B_DESC_HI << 32), bl = (uint64_t)b_lo | ((uint64_t)B_DESC_HI << 32);
    mma_f16(d, a_hi, bh, idesc, acc);
    if constexpr (!FAST) {
      mma_f16(d, a_hi, bl, idesc, 1u);
      mma_f16(d, a_lo, bh, idesc, 1u);
    }
  }
  __syncwarp();
}
__device__ __forceinline__ void issue_commit(uint32_t bar) {
  if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  __syncwarp();
}

// barrier slots (8 bytes each): [0] weights; per pipeline g at 1 + g * BARS_PER_PIPE: full_d[2] (even / odd tiles: the
// issuer may run into the next tile while the last accumulator of this one is drained), a1_ready, in_full[2], sub_ready[7]
constexpr int BAR_FULL = 0, BAR_A1 = 2, BAR_IN = 3, BAR_SUB = 5, BARS_PER_PIPE = 5 + MAX_SUB;

// ------------------------------------------------------------------------------------------------ MMA issuer
// All k-steps of MMA M of step S: every address and descriptor is a compile-time constant plus sb4 (a uniform value).
constexpr int kstep_index(const SStep* st, int S, int M) {  // running k-step number of (step S, mma M, k = 0) within a tile
  int n = 0;
  for (int s = 0; s <= S; ++s)
    for (int m = 0; m < st[s].n_mma; ++m) {
      if (s == S && m == M) return n;
      n += st[s].mma[m].ks_n;
    }
  return n;
}
template <class P, int S, int M, bool TRACE, bool FAST>
__device__ __forceinline__ void issue_mma(const uint32_t tcol, const uint32_t bar_p, const uint32_t sb4, uint32_t& par_sub, uint32_t* tr) {
  constexpr SMma mm = P::step(S).mma[M];
  constexpr SStep all[5] = {P::step(0), P::step(1), P::step(2), P::step(3), P::step(4)};
  constexpr int k_index0 = kstep_index(all, S, M);
  constexpr uint32_t np16 = (uint32_t)P::Np(mm.layer) * 16u;  // bytes between the two 8-wide K chunks of a k-step
  constexpr uint32_t idesc = make_idesc(mm.n);
  constexpr uint32_t lbo_f = ((np16 >> 4) & 0x3FFFu) << 16;
  constexpr uint32_t mat = 2u * (uint32_t)P::Np(mm.layer) * (uint32_t)P::Kp(mm.layer);  // bytes of one image (hi or lo)
  constexpr uint32_t b_base = P::w_off(mm.layer) + (uint32_t)mm.ks0 * 2u * np16 + (uint32_t)mm.n0 * 16u;
  bool ready = false;  // sub-chunk k was already seen complete by the probe issued before the previous k-step's MMAs
#pragma unroll
  for (int k = 0; k < mm.ks_n; ++k) {
    if (mm.dep) {  // this k-step's operand is sub-chunk k of the accumulator the previous epilogue rewrites
      if (!ready) mbar_wait(bar_p + 8u * (BAR_SUB + (uint32_t)k), (par_sub >> k) & 1u);
      par_sub ^= 1u << k;
      tc_fence_after();
      // probe the next sub-chunk now: the probe's latency (~80 cycles) hides behind this k-step's MMAs
      if (k + 1 < mm.ks_n) ready = mbar_test(bar_p + 8u * (BAR_SUB + (uint32_t)(k + 1)), (par_sub >> (k + 1)) & 1u);
    }
    if constexpr (TRACE) { if (tr) tr[2 * (k_index0 + k)] = (uint32_t)clock64(); }
    const uint32_t b_off = b_base + (uint32_t)k * 2u * np16;
    // smem_base is 128-byte aligned and the whole window is < 256 KiB, so the 14-bit address field cannot carry
    const uint32_t b_hi = sb4 + ((b_off >> 4) | lbo_f);
    const uint32_t b_lo = sb4 + (((b_off + mat) >> 4) | lbo_f);
    const uint32_t a_hi = tcol + (uint32_t)mm.a_col + 16u * (uint32_t)k;
    issue_kstep<FAST>(tcol + (uint32_t)mm.d_col, a_hi, a_hi + 8u, b_hi, b_lo, idesc, (k > 0 || mm.acc) ? 1u : 0u);
    if constexpr (TRACE) { if (tr) tr[2 * (k_index0 + k) + 1] = (uint32_t)clock64(); }
  }
}
// Experiments that did not pay (profiles/r01_tc4_notes.md): serialising the two pipelines' steps on the tensor pipe, by a
// shared-memory ticket lock or by a token passed through two mbarriers.  Left alone the pipelines fall into lock-step
// (both convert, then both issue MMAs); the token does break that, but a step that waits for sub-chunks then holds the
// pipe idle, and any lane-0 polling loop in the issuer warp makes ptxas treat the warp as divergent and fall back to
// the slow tcgen05.mma issue sequence (4x slower).  Net effect: -1 % .. -75 %.
template <class P, int S, bool TRACE, bool FAST>
__device__ __forceinline__ void issue_step(const uint32_t tcol, const uint32_t bar_p, const uint32_t bar_full, const uint32_t sb4,
                                           uint32_t& par_sub, uint32_t* tr) {
  issue_mma<P, S, 0, TRACE, FAST>(tcol, bar_p, sb4, par_sub, tr);
  if constexpr (P::step(S).n_mma > 1) issue_mma<P, S, 1, TRACE, FAST>(tcol, bar_p, sb4, par_sub, tr);
  issue_commit(bar_full);
}

template <bool ENC, int KA, int NL, bool TRACE, int G, bool FAST>
__device__ __forceinline__ void run_issuer(const Tc4Params& p, const uint32_t smem_base, const uint32_t bars_base,
                                           const uint32_t in_stage_bytes, const uint32_t in0_off) {
  using P = Prog<ENC, KA, NL>;
  // Everything below is warp-uniform.  TMEM: this CTA owns all 512 columns of the SM (one CTA per SM), so the
  // allocation base is column 0 / lane 0 (checked by the caller).  The pipeline index is a template parameter (one copy
  // of the unrolled step program, ~1.3k instructions, per pipeline): with it every TMEM address is parameter + constant.
  // The single-product FAST16 mode stays on the table-driven kernel.
  const uint32_t tcol = p.tmem0 + (uint32_t)G * PIPE_COLS;
  const uint32_t bar_w = bars_base;
  const uint32_t bar_p = bars_base + 8u * (1 + G * BARS_PER_PIPE);
  // descriptors only carry address bits [4, 18): take them from the kernel parameter (a uniform register) rather than
  // from the cvta'd pointer, so that every tcgen05.mma operand is provably warp-uniform
  const uint32_t sb4 = p.smem_off >> 4;
  const bool lane0 = (threadIdx.x & 31) == 0;
  const int64_t n_tiles = G < p.pipes ? (p.n_rows + TILE - 1) / TILE : 0;
  const int64_t tile_stride = (int64_t)gridDim.x * p.pipes;
  const int64_t tile0 = (int64_t)blockIdx.x * p.pipes + G;
  const uint32_t in_s = smem_base + in0_off + (uint32_t)G * 2u * in_stage_bytes;
  const uint32_t row_bytes = (uint32_t)p.in_dim * (p.in_f16 ? 2u : 4u);
  auto trace = [&](int64_t lt, int slot) {
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && G == p.trace_pipe && lane0 && lt < 16) p.trace[lt * 64 + slot] = (uint32_t)clock64();
    }
  };
  // input tile -> stage (lt & 1) by cp.async.bulk; a ragged tail (bytes not a multiple of 16) is finished by hand
  auto load_tile = [&](int64_t tile, int64_t lt) {
    if (lane0 && tile < n_tiles) {
      const int rows = (int)min((int64_t)TILE, p.n_rows - tile * TILE);
      const uint32_t bytes = (uint32_t)rows * row_bytes;
      const uint32_t dst = in_s + (uint32_t)(lt & 1) * in_stage_bytes;
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.in) + (size_t)tile * TILE * row_bytes;
      const uint32_t bar = bar_p + 8u * (BAR_IN + (uint32_t)(lt & 1));
      const uint32_t bulk = bytes & ~15u;
      for (uint32_t o = bulk; o < bytes; o += 2)
        asm volatile("st.shared.b16 [%0], %1;" ::"r"(dst + o), "h"(*reinterpret_cast<const uint16_t*>(src + o)) : "memory");
      if (bulk) {
        mbar_expect_tx(bar, bulk);
        bulk_g2s(dst, src, bulk, bar);
      } else {
        mbar_arrive(bar);
      }
    }
    __syncwarp();
  };
  load_tile(tile0, 0);
  load_tile(tile0 + tile_stride, 1);
  mbar_wait(bar_w, 0);  // weights resident
  uint32_t par_a1 = 0, par_sub = 0;
  int64_t lt = 0;
  for (int64_t tile = tile0; tile < n_tiles; tile += tile_stride, ++lt) {
    trace(lt, 32);
    mbar_wait(bar_p + 8u * BAR_A1, par_a1);
    par_a1 ^= 1u;
    tc_fence_after();
    trace(lt, 33);
    const uint32_t bar_full = bar_p + 8u * (BAR_FULL + (uint32_t)(lt & 1));
    uint32_t* tr = nullptr;  // second half of the trace buffer: per k-step (wait passed, issued) clocks of the issuer
    if constexpr (TRACE) { if (blockIdx.x == 0 && lane0 && lt < 16) tr = p.trace + 1024 * (1 + G) + lt * 64; }  // both pipelines' issuers
    auto after_commit = [&](int s_done) {
      trace(lt, 34 + 5 * s_done + 4);
      if (s_done == 0) load_tile(tile + 2 * tile_stride, lt);  // this tile's stage was consumed before a1_ready; off the s0 critical path
    };
    issue_step<P, 0, TRACE, FAST>(tcol, bar_p, bar_full, sb4, par_sub, tr); after_commit(0);
    issue_step<P, 1, TRACE, FAST>(tcol, bar_p, bar_full, sb4, par_sub, tr); after_commit(1);
    issue_step<P, 2, TRACE, FAST>(tcol, bar_p, bar_full, sb4, par_sub, tr); after_commit(2);
    issue_step<P, 3, TRACE, FAST>(tcol, bar_p, bar_full, sb4, par_sub, tr); after_commit(3);
    issue_step<P, 4, TRACE, FAST>(tcol, bar_p, bar_full, sb4, par_sub, tr); after_commit(4);
  }
}

// ------------------------------------------------------------------------------------------------ loader + epilogue
template <bool ENC, int KA, int NL, bool TRACE>
__device__ __forceinline__ void run_epilogue(const Tc4Params& p, const int g, const int wq, const int h, const uint32_t smem_base,
                                             uint8_t* smem, const uint32_t bars_base, const uint32_t in_stage_bytes,
                                             const uint32_t in0_off, const uint32_t out0_off, const uint32_t out_stage_bytes,
                                             const float2 (*norm_s)[32]) {
  using P = Prog<ENC, KA, NL>;
  const int lane = threadIdx.x & 31;
  const int row = wq * 32 + lane;                                      // tile row == TMEM lane
  const uint32_t tbase = ((uint32_t)(wq * 32) << 16) + (uint32_t)g * PIPE_COLS;  // TMEM address of this warp's lanes, pipeline columns
  const uint32_t bar_p = bars_base + 8u * (1 + g * BARS_PER_PIPE);
  const int64_t n_tiles = g < p.pipes ? (p.n_rows + TILE - 1) / TILE : 0;
  const int64_t tile_stride = (int64_t)gridDim.x * p.pipes;
  const int64_t tile0 = (int64_t)blockIdx.x * p.pipes + g;
  const int in_dim = p.in_dim, out_dim = p.out_dim;
  const uint8_t* in_s = smem + in0_off + (uint32_t)g * 2u * in_stage_bytes;
  const uint32_t out_elt = p.out_f16 ? 2u : 4u;
  uint8_t* out_sb = smem + out0_off + (uint32_t)g * out_stage_bytes + (size_t)(wq * 32 * out_dim) * out_elt;  // this warp's 32 rows
  float* out_s = reinterpret_cast<float*>(out_sb);
  const uint32_t out_s_addr = smem_base + out0_off + (uint32_t)g * out_stage_bytes + (uint32_t)(wq * 32 * out_dim) * out_elt;
  auto trace = [&](int64_t lt, int slot) {
    if constexpr (TRACE) {
      if (blockIdx.x == 0 && g == p.trace_pipe && wq == 0 && lane == 0 && lt < 16 && (h == 0) != (slot == 1 || slot == 2 || (slot >= 29 && slot <= 31))) p.trace[lt * 64 + slot] = (uint32_t)clock64();
    }
  };

  // First operand of a tile (done by the h == 1 warps, one tile ahead, while the h == 0 warps drain the last
  // accumulator): this thread's row -> normalise, 1.0 in the bias slot, split -> TMEM, one k-step (16 features) at a time.
  auto convert_a1 = [&](int64_t tile, int64_t lt) {
    const int rows = (int)min((int64_t)TILE, p.n_rows - tile * TILE);
    mbar_wait(bar_p + 8u * (BAR_IN + (uint32_t)(lt & 1)), (uint32_t)(lt >> 1) & 1u);
    const float* xr = reinterpret_cast<const float*>(in_s + (uint32_t)(lt & 1) * in_stage_bytes) + row * in_dim;
    const __half* xr16 = reinterpret_cast<const __half*>(in_s + (uint32_t)(lt & 1) * in_stage_bytes) + row * in_dim;
    trace(lt - 1, 29);
#pragma unroll
    for (int ks = 0; ks < KA / 16; ++ks) {
      float xv[16];
      const int k0 = ks * 16;
      if (p.in_f16) {  // float16 latent rows
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = (k0 + j < in_dim && row < rows) ? __half2float(xr16[k0 + j]) : 0.f;
      } else if ((in_dim & 3) == 0) {  // 128-bit loads of this row's 16 features
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k0 + 4 * q < in_dim && row < rows) t = reinterpret_cast<const float4*>(xr + k0)[q];
          xv[4 * q] = t.x; xv[4 * q + 1] = t.y; xv[4 * q + 2] = t.z; xv[4 * q + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = (k0 + j < in_dim && row < rows) ? xr[k0 + j] : 0.f;
      }
      // numpy float32 (x - min) / range (data_processing.py:151); here the quotient is x * rcp_rn(range): within 1 ulp of
      // the IEEE quotient, far inside the 1e-5 budget of the latent.  Branch-free: the table also covers "no
      // normalisation" ({0, 1}) and turns the zero loaded beyond the real features into the bias slot's constant one.
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 mr = norm_s[0][k0 + j];
        xv[j] = __fsub_rn(xv[j], mr.x) * mr.y;
      }
      uint32_t pk[16];
#pragma unroll
      for (int j = 0; j < 8; ++j) split2(xv[2 * j], xv[2 * j + 1], pk[j], pk[8 + j]);
      if (ks == KA / 16 - 1) trace(lt - 1, 30);
      tmem_st16(tbase + P::A1_COL + 16u * (uint32_t)ks, pk);
    }
    tc_wait_st();
    trace(lt - 1, 31);
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_p + 8u * BAR_A1);
  };

  int64_t lt = 0;
  bool store_pending = false;
  if (h == 1 && tile0 < n_tiles) convert_a1(tile0, 0);
  for (int64_t tile = tile0; tile < n_tiles; tile += tile_stride, ++lt) {
    const int rows = (int)min((int64_t)TILE, p.n_rows - tile * TILE);
    const uint32_t bar_full = bar_p + 8u * (BAR_FULL + (uint32_t)(lt & 1));
    const uint32_t par0 = (uint32_t)(lt >> 1);  // five phases per tile on each of the two barriers: parity of step s = (par0 + s) & 1
    trace(lt, 0);
    // ---- steps 0..3: drain the accumulator into the next layer's operand, in place.  Sub-chunks h, h + 2, ... of the
    // region are this warp's; the TMEM load of the next one is in flight while the current one is converted.  The loop
    // over steps is rolled (region, width and multipliers come from the parameter table): code size matters here.
#pragma unroll 1
    for (int s = 0; s < P::N_STEPS - 1; ++s) {
      const uint32_t col = tbase + (uint32_t)p.epi_col[s];
      const int n_sub = p.epi_nsub[s];
      const uint64_t c1 = pack2(p.epi_c1[s], p.epi_c1[s]);
      const float c2 = p.epi_c2[s];
      mbar_wait(bar_full, (par0 + (uint32_t)s) & 1u);  // the accumulator of step s is complete
      tc_fence_after();
      trace(lt, 3 + 4 * s);
      uint32_t va[16], vb[16], pk[16];
      auto finish = [&](int i) {
        tmem_st16(col + 16u * (uint32_t)i, pk);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + 8u * (BAR_SUB + (uint32_t)i));  // sub-chunk i is a valid A operand now
      };
      int i = h;
      if (i < n_sub) tmem_ld16(col + 16u * (uint32_t)i, va);
#pragma unroll 1
      for (; i < n_sub; i += 2 * EPI_GROUPS) {
        tc_wait_ld();
        if (i + EPI_GROUPS < n_sub) tmem_ld16(col + 16u * (uint32_t)(i + EPI_GROUPS), vb);
        convert16(va, c1, c2, pk);
        finish(i);
        if (i == h) trace(lt, 4 + 4 * s);
        if (i + EPI_GROUPS < n_sub) {
          tc_wait_ld();
          if (i + 2 * EPI_GROUPS < n_sub) tmem_ld16(col + 16u * (uint32_t)(i + 2 * EPI_GROUPS), va);
          convert16(vb, c1, c2, pk);
          finish(i + EPI_GROUPS);
          if (i == h) trace(lt, 5 + 4 * s);
        }
      }
    }
    if (h >= 2) {
      // nothing to do in the last step for the extra column groups
    } else if (h == 1) {
      // the last accumulator is drained by the h == 0 warps; meanwhile the first operand of the next tile (its TMEM
      // region has been idle since step 1 / step 3, which this warp has seen complete)
      trace(lt, 1);
      if (tile + tile_stride < n_tiles) convert_a1(tile + tile_stride, lt + 1);
      trace(lt, 2);
    } else {
      constexpr SEpi ep = P::step(P::N_STEPS - 1).epi;
      mbar_wait(bar_full, (par0 + (uint32_t)(P::N_STEPS - 1)) & 1u);
      tc_fence_after();
      trace(lt, 3 + 4 * (P::N_STEPS - 1));
      // ---- last accumulator -> scale, activation, range check, un-normalise -> this warp's slice of the out stage
      uint32_t v[NL];
      {
        uint32_t t0[16];
        tmem_ld16(tbase + (uint32_t)ep.col, t0);
        if constexpr (NL == 32) {
          uint32_t t1[16];
          tmem_ld16(tbase + (uint32_t)ep.col + 16u, t1);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = t0[j]; v[16 + j] = t1[j]; }
        } else {
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = t0[j];
        }
      }
      trace(lt, 25);
      if (store_pending) {  // the previous tile's bulk store must have finished READING the stage
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }
      trace(lt, 28);
      // branch-free over the padded width (a runtime `j < out_dim` test per element made this a chain of 16 - 32 dependent
      // branch blocks, ~70 cycles per element): compute everything, then predicated stores
      const float s1 = p.c1[ep.layer], s2 = p.c2[ep.layer];
      float ys[NL];
      uint32_t bad_mask = 0;
#pragma unroll
      for (int j = 0; j < NL; ++j) {
        const float a = __uint_as_float(v[j]);
        const float y = fmaf(fabsf(a), s2, a * s1);
        bad_mask |= (fabsf(y) <= 3.0e38f) ? 0u : (1u << j);  // inf / NaN: an fp16 operand overflowed somewhere upstream
        const float2 rm = norm_s[1][j];
        ys[j] = fmaf(y, rm.x, rm.y);                           // y * range + min (data_processing.py:203); {1, 0} when absent
      }
      const bool bad = (bad_mask & (out_dim >= 32 ? 0xFFFFFFFFu : ((1u << out_dim) - 1u))) != 0u;
      if (p.out_f16) {  // float16 latent rows (the range check above saw the fp32 values)
        __half* o16 = reinterpret_cast<__half*>(out_sb) + lane * out_dim;
#pragma unroll
        for (int j = 0; j < NL; ++j)
          if (j < out_dim) o16[j] = __float2half_rn(ys[j]);
      } else if ((out_dim & 3) == 0) {  // rows are 16-byte multiples: 128-bit stores (2-way instead of 8-way bank conflicts at 24 floats per row)
#pragma unroll
        for (int q = 0; q < NL / 4; ++q)
          if (4 * q < out_dim) reinterpret_cast<float4*>(out_s + lane * out_dim)[q] = make_float4(ys[4 * q], ys[4 * q + 1], ys[4 * q + 2], ys[4 * q + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < NL; ++j)
          if (j < out_dim) out_s[lane * out_dim + j] = ys[j];
      }
      if (bad && row < rows) atomicOr(p.flag, 1);
      trace(lt, 26);
      const int my_rows = min(32, max(0, rows - wq * 32));
      const uint32_t bytes = (uint32_t)(my_rows * out_dim) * out_elt;
      float* gdst = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.out) + ((size_t)tile * TILE + (size_t)wq * 32) * out_dim * out_elt);
      if ((bytes & 15u) == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0 && bytes) { bulk_s2g(gdst, out_s_addr, bytes); }
        store_pending = true;
      } else {  // ragged tail: plain stores
        __syncwarp();
        if (p.out_f16) {
          for (int e = lane; e < my_rows * out_dim; e += 32) reinterpret_cast<__half*>(gdst)[e] = reinterpret_cast<const __half*>(out_sb)[e];
        } else {
          for (int e = lane; e < my_rows * out_dim; e += 32) gdst[e] = out_s[e];
        }
        __syncwarp();
      }
      trace(lt, 27);
    }
    trace(lt, 24);
  }
  if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ kernel
template <bool ENC, int KA, int NL, bool TRACE, bool FAST>
__global__ void __launch_bounds__(NTHREADS, 1) chain_tc4_kernel(const __grid_constant__ Tc4Params p) {
  using P = Prog<ENC, KA, NL>;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[1 + NPIPE * BARS_PER_PIPE];
  __shared__ uint32_t tmem_base_s;
  // [0][k] = {pre_min, 1 / pre_range} of input feature k ({0, 1} without normalisation; the bias slot k == in_dim holds
  // {-1, 1} so that a zero input becomes the constant one; {0, 1} beyond it), [1][j] = {post_range, post_min}
  __shared__ float2 norm_s[2][32];

  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform for the compiler too (role dispatch below)
  const uint32_t smem_base = smem_u32(smem);
  if (p.probe != nullptr) {
    if (tid == 0) *p.probe = smem_base & 0x3FFFFu;
    return;
  }
  if ((smem_base & 0x3FFFFu) != p.smem_off) __trap();
  const uint32_t bars_base = smem_u32(&bars[0]);
  const uint32_t in_stage_bytes = (uint32_t)((TILE * p.in_dim * 4 + 127) & ~127);   // sized for float32 rows either way
  const uint32_t out_stage_bytes = (uint32_t)((TILE * p.out_dim * 4 + 127) & ~127);
  const uint32_t in0_off = (P::W_BYTES + 127u) & ~127u;
  const uint32_t out0_off = in0_off + NPIPE * 2u * in_stage_bytes;

  if (tid == 0) {
    mbar_init(bars_base, 1);
    for (int g = 0; g < NPIPE; ++g) {
      const uint32_t b0 = bars_base + 8u * (1 + g * BARS_PER_PIPE);
      mbar_init(b0 + 8u * BAR_FULL, 1);             // one tcgen05.commit per step (even tiles)
      mbar_init(b0 + 8u * (BAR_FULL + 1), 1);       // (odd tiles)
      mbar_init(b0 + 8u * BAR_A1, 4);               // the four h == 1 warps
      mbar_init(b0 + 8u * BAR_IN, 1);               // the loader's expect_tx
      mbar_init(b0 + 8u * (BAR_IN + 1), 1);
      for (int c = 0; c < MAX_SUB; ++c) mbar_init(b0 + 8u * (BAR_SUB + c), 4);  // the 4 lane-quarter warps of one half
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 64) {  // column (de)normalisation vectors
    const int which = tid >> 5, k = tid & 31;
    float2 v;
    if (which == 0) {
      v = make_float2(k == p.in_dim ? -1.f : 0.f, 1.f);
      if (p.pre_min != nullptr && k < p.in_dim) v = make_float2(p.pre_min[k], __frcp_rn(p.pre_range[k]));  // the loader multiplies by 1 / range
    } else {
      v = make_float2(1.f, 0.f);
      if (p.post_min != nullptr && k < p.out_dim) v = make_float2(p.post_range[k], p.post_min[k]);
    }
    norm_s[which][k] = v;
  }
  if (warp == 0) {  // warp 0 owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // all 512 columns of the SM belong to this CTA, so the base is column 0, lane 0; the issuer relies on that to keep
  // its TMEM addresses compile-time constants
  if (tmem_base_s != 0u || p.tmem0 != 0u) __trap();
  if (tid == 0) {  // resident weight image: global -> smem through the async proxy (what the MMA reads through)
    mbar_expect_tx(bars_base, P::W_BYTES);
    for (uint32_t off = 0; off < P::W_BYTES; off += 32768) {
      const uint32_t n = P::W_BYTES - off < 32768 ? P::W_BYTES - off : 32768;
      bulk_g2s(smem_base + off, p.wimg + off, n, bars_base);
    }
  }

  if (warp == NPIPE * EPI_WARPS) run_issuer<ENC, KA, NL, TRACE, 0, FAST>(p, smem_base, bars_base, in_stage_bytes, in0_off);
  else if (warp == NPIPE * EPI_WARPS + 1) run_issuer<ENC, KA, NL, TRACE, 1, FAST>(p, smem_base, bars_base, in_stage_bytes, in0_off);
  else run_epilogue<ENC, KA, NL, TRACE>(p, warp / EPI_WARPS, warp & 3, (warp >> 2) % EPI_GROUPS, smem_base, smem, bars_base, in_stage_bytes,
                                        in0_off, out0_off, out_stage_bytes, norm_s);

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(0u), "r"(512) : "memory");
}

template <bool ENC, int KA, int NL, bool TRACE = false, bool FAST = false>
int launch_one(const bb_ctx* ctx, Tc4Params p, cudaStream_t stream) {
  using P = Prog<ENC, KA, NL>;
  for (int s = 0; s < P::N_STEPS - 1; ++s) {
    const SEpi ep = P::step(s).epi;
    p.epi_col[s] = ep.col; p.epi_nsub[s] = ep.w / 16;
    p.epi_c1[s] = p.c1[ep.layer]; p.epi_c2[s] = p.c2[ep.layer];
  }
  const size_t in_b = ((size_t)TILE * p.in_dim * 4 + 127) & ~(size_t)127;
  const size_t out_b = ((size_t)TILE * p.out_dim * 4 + 127) & ~(size_t)127;
  const size_t smem_bytes = ((P::W_BYTES + 127u) & ~127u) + NPIPE * (2 * in_b + out_b);
  auto k = chain_tc4_kernel<ENC, KA, NL, TRACE, FAST>;
  // per instantiation and per device (function attributes belong to a device's context)
  constexpr int MAX_DEV = 64;
  static bool attr_set_d[MAX_DEV] = {};
  static uint32_t smem_off_d[MAX_DEV] = {};
  static size_t static_smem_d[MAX_DEV] = {};
  if (ctx->device < 0 || ctx->device >= MAX_DEV) return BB_ERR_UNSUPPORTED;
  bool& attr_set = attr_set_d[ctx->device];
  uint32_t& smem_off = smem_off_d[ctx->device];
  size_t& static_smem = static_smem_d[ctx->device];
  if (!attr_set) {
    cudaFuncAttributes fa;
    BB_CUDA(cudaFuncGetAttributes(&fa, k));
    static_smem = fa.sharedSizeBytes;
    if (smem_bytes + static_smem > ctx->smem_optin) return BB_ERR_UNSUPPORTED;
    BB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctx->smem_optin - static_smem)));
    // where the dynamic shared memory of THIS kernel starts (a link-time constant the host cannot query): one probe launch
    uint32_t* probe_dev = nullptr;
    BB_CUDA(cudaMalloc(&probe_dev, 4));
    Tc4Params q = p;
    q.probe = probe_dev;
    k<<<1, NTHREADS, smem_bytes, stream>>>(q);
    BB_CUDA(cudaMemcpyAsync(&smem_off, probe_dev, 4, cudaMemcpyDeviceToHost, stream));
    BB_CUDA(cudaStreamSynchronize(stream));
    BB_CUDA(cudaFree(probe_dev));
    attr_set = true;
  }
  if (smem_bytes + static_smem > ctx->smem_optin) return BB_ERR_UNSUPPORTED;  // wide rows: the table-driven kernel (one pipeline) takes over
  p.smem_off = smem_off;
  p.tmem0 = 0;
  p.probe = nullptr;
  const int64_t n_tiles = (p.n_rows + TILE - 1) / TILE;
  const int64_t want = (n_tiles + NPIPE - 1) / NPIPE;
  const int grid = (int)(want < ctx->sm_count ? want : ctx->sm_count);
  k<<<grid, NTHREADS, smem_bytes, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace

// Host entry: runs the statically shaped kernel when `c->tc4` says the chain belongs to the family and the buffers
// are float32 and 16-byte aligned (cp.async.bulk); BB_ERR_UNSUPPORTED tells the caller to use the table-driven kernel.
int bb_tc4_launch(bb_ctx* ctx, const Chain* c, const void* in, int in_dtype, int64_t n_rows, const float* pre_min,
                  const float* pre_range, const float* post_min, const float* post_range, void* out, int out_dtype, int fast,
                  int* flag_dev, uint32_t* trace, cudaStream_t stream) {
  const Tc4Plan& t = c->tc4;
  if (!t.ok) return BB_ERR_UNSUPPORTED;
  // float16 rows only where the latent is: the encoder's output, the decoder's input
  if ((in_dtype == BB_F16 && t.enc) || (out_dtype == BB_F16 && !t.enc)) return BB_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(in) & 15u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return BB_ERR_UNSUPPORTED;
  if (n_rows == 0) return BB_OK;
  Tc4Params p;
  p.wimg = reinterpret_cast<const uint8_t*>(c->tc_blob_dev);
  p.in = reinterpret_cast<const float*>(in);
  p.out = reinterpret_cast<float*>(out);
  p.n_rows = n_rows;
  p.pre_min = pre_min; p.pre_range = pre_range; p.post_min = post_min; p.post_range = post_range;
  for (int l = 0; l < 4; ++l) { p.c1[l] = t.c1[l]; p.c2[l] = t.c2[l]; }
  p.in_dim = c->desc.in_dim; p.out_dim = c->desc.out_dim;
  p.flag = flag_dev; p.trace = trace;
  p.in_f16 = in_dtype == BB_F16; p.out_f16 = out_dtype == BB_F16;
  p.pipes = getenv("BALER_B200_TC4_PIPES") ? atoi(getenv("BALER_B200_TC4_PIPES")) : NPIPE;
  p.trace_pipe = getenv("BALER_B200_TRACE_PIPE") ? atoi(getenv("BALER_B200_TRACE_PIPE")) : 0;
  if (trace != nullptr) {  // SM-clock trace build: the two CMS shapes only
    if (t.enc && t.ka == 32 && t.nl == 16) return launch_one<true, 32, 16, true>(ctx, p, stream);
    if (!t.enc && t.ka == 16 && t.nl == 32) return launch_one<false, 16, 32, true>(ctx, p, stream);
    return BB_ERR_UNSUPPORTED;
  }
  if (fast) {  // single-product mode: the CMS shapes (what the bench reports beside the exact mode)
    if (t.enc && t.ka == 32 && t.nl == 16) return launch_one<true, 32, 16, false, true>(ctx, p, stream);
    if (!t.enc && t.ka == 16 && t.nl == 32) return launch_one<false, 16, 32, false, true>(ctx, p, stream);
    return BB_ERR_UNSUPPORTED;
  }
  if (t.enc) {
    if (t.ka == 32 && t.nl == 16) return launch_one<true, 32, 16>(ctx, p, stream);
    if (t.ka == 32 && t.nl == 32) return launch_one<true, 32, 32>(ctx, p, stream);
    if (t.ka == 16 && t.nl == 16) return launch_one<true, 16, 16>(ctx, p, stream);
    if (t.ka == 16 && t.nl == 32) return launch_one<true, 16, 32>(ctx, p, stream);
  } else {
    if (t.ka == 16 && t.nl == 32) return launch_one<false, 16, 32>(ctx, p, stream);
    if (t.ka == 32 && t.nl == 32) return launch_one<false, 32, 32>(ctx, p, stream);
    if (t.ka == 16 && t.nl == 16) return launch_one<false, 16, 16>(ctx, p, stream);
    if (t.ka == 32 && t.nl == 16) return launch_one<false, 32, 16>(ctx, p, stream);
  }
  return BB_ERR_UNSUPPORTED;
}
