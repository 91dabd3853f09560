// tcgen05 (5th-gen tensor core) fused MLP chains — HOLD_MLP_TC.
//
// One persistent CTA per SM walks 128-row tiles through the whole layer chain:
//   * warp 0           : bulk-async (TMA engine) copies of pre-swizzled fp16 weight chunks L2 -> smem ring
//   * warp 1           : tcgen05.mma issuer; D[128 x 256] fp32 accumulators in TMEM, ping-pong per layer
//                        (both walk their loops warp-uniformly and elect the issuing lane, see elect_one())
//   * warps 2..17      : epilogue — tcgen05.ld the accumulator in 32-column hand-offs, bias + activation in fp32,
//                        split into fp16 hi/lo and write the next layer's A operand (SW128 K-major) to smem;
//                        chunk-level mbarriers let layer l+1's MMAs start while layer l's epilogue is running.
// Arithmetic: every fp32 operand x is split x = hi + lo (fp16 each: 11 + 11 mantissa bits) and each product is
// three MMAs (hi*hi + lo*hi + hi*lo, fp32 accumulate) — ~2^-21 relative per product (measured 2.5e-7 on a
// 256-term dot, i.e. fp32-GEMM level), with an absolute floor of ~3e-8 where lo goes subnormal.  This is what the
// 1e-4 parity bar needs: an sdf error eps reaches the density as eps/beta^2 (beta down to 1e-2), so plain bf16
// (2^-9) and even a bf16 split (2^-16, measured 83 % of pixels within 1e-4) do not meet it (SURVEY §7 "hard
// parts").  Range: |x| must stay below 65504 (activations/weights here are O(1e-3..1e2)).  The sdf / rgb heads (1 resp. 3 output
// rows) are fp32 dot products in the epilogue.  Gradients: reverse mode in the same kernel (MODE MLP_SDF_REV): the forward
// epilogues stash softplus'(z_l) as unorm16 (512 KB per CTA: 76 MB for the grid, L2-resident; quantisation step 1.5e-5 moves
// the gradient by <= 1.2e-5 of its scale, tests/test_cpu_stash_quant.py), the backward layers run over transposed weight images.
// Embeddings (tile prologue, skip layer) are inlined branch-free sin/cos (sincos_cw).
// Round-2 A/B on hardware (profiles/r02_variants.md) retired the other variants (base-2-domain epilogue, rebuilt 16-column
// epilogue, CTA-pair kernels, forward-mode gradient): none beat this kernel at equal error.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "embed_phases.h"
#include "mlp_simt.cuh"

namespace hold {

constexpr int kTcRows = 128;
constexpr int kTcStageBytes = 32768;      // one weight stage: [256 n x 32 k] fp16 hi (16 KB) + lo (16 KB)
constexpr int kTcAChunkBytes = 16384;     // one A chunk: [128 rows x 64 k] fp16
constexpr int kTcThreads = 192;
// Power-of-two operand scaling (exact): fp16 operands are fed to the tensor core as A * 2^6 and W * 2^10 so that
// the LOW halves of the hi/lo split stay in fp16's normal range for |a| >= 2e-3, |w| >= 1.2e-4 (unscaled, the low
// half of every |x| < 0.125 is subnormal; measured: the split then gains only 2x over bf16).  The accumulator is
// rescaled by 2^-16 in the epilogue's bias FMA.  Range: |a| < 1023, |w| < 64.
constexpr float kTcScaleA = 64.0f, kTcScaleW = 1024.0f, kTcUnscale = 1.0f / (64.0f * 1024.0f);
// Accumulator-truncation compensation of the Softplus chains.  tcgen05.mma adds into its fp32 accumulator with truncation (round
// toward zero), not round-to-nearest: measured against float64 on hardware (profiles/r02_tc_accumulator_bias.md) every layer of
// 48 accumulations (K = 256 x 3 passes) comes out SHRUNK by ~12 x 2^-24 relative — a bias, identical for the hand and object
// nets and for perturbed weights (zero crossing of the mean sdf error at c = 11.1 / 11.0 / 12.4 / 12.9).  Undoing it in the
// epilogue's unscale multiply (free) cuts the sdf error 5-8x (5.99e-6 -> 9.5e-7 of the sdf scale), which is what the Laplace
// density needs at small beta (error amplification 1 / (2 beta^2), engine/density.py:21-26).
constexpr int kTcAccComp = 12;

constexpr int kTcMaxSteps = 17;
struct TcLayer {
  const uint8_t* wimg;  // pre-swizzled stage images, nst * 32 KB
  const float* bias;    // [256]
  int nst;              // number of 32-wide k stages
  int N;                // valid outputs
};

__global__ void k_scale_vec(const float* __restrict__ src, int n, float c, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = c * src[i];
}

struct TcMlp {
  uint8_t* sdf_img[HOLD_MAX_LAYERS] = {nullptr};
  uint8_t* rgb_img[HOLD_MAX_LAYERS] = {nullptr};
  uint8_t* sdf_imgT[HOLD_MAX_LAYERS] = {nullptr};  // W_l^T images of layers 0..7 for the reverse-mode gradient, 8: feature rows (training)
  uint8_t* rgb_imgT[6] = {nullptr};                // training backward: [0] W_0^T feature part, [1] W_0^T other inputs, [2..4] W_1..3^T
  int sdf_nst[HOLD_MAX_LAYERS], rgb_nst[HOLD_MAX_LAYERS];
  float* bias_s = nullptr;   // [13][256]: biases as the epilogues add them: SDF lin0..7 times kTcScaleA, lin8 (feature rows) plain, colour lin0..3 times kTcScaleA
};

struct TcArgs {
  int P, n_layers;
  TcLayer L[kTcMaxSteps];
  uint16_t* sig;  // reverse mode: per-CTA stash of softplus'(z_l) as unorm16, [grid][8][128][256]
  const float* w_last;
  const float* b_last;
  const float* xc;
  const float* embed_w;
  float* sdf;
  float* grad;
  float* feat;
  const float* normal;
  const float* pose_embed;
  const float* time_code;
  int pts_per_frame, k0;
  float* rgb;
  const SamplerState* st;
  int* err;
  const float* cam;          // background modes: ray origins / directions [R,3] of this frame chunk, its frame code [32]
  const float* dirs;
  const float* frame_code;
  float r_sphere;
  float unscale;             // accumulator -> value: kTcUnscale (times the experimental compensation factor, hold_debug_set key 2)
  // MLP_LINEAR (hold_linear): C[P, nvalid] = A[P, kvalid] . W^T (+ bias): one layer against one packed image
  const float* lin_in;
  float* lin_out;
  const float* in_scale;     // device scalar s (a power of two) or NULL: A is fed as A / s, C comes out times s (keeps tiny gradients
                             // inside the fp16 hi/lo split's range)
  int lda, ldc, kvalid, nvalid;
  int passes;                // MMAs per product: 3 = hi*hi + lo*hi + hi*lo (always, in production); 2 / 1 only through the measurement
                             // hook hold_debug_set(ctx, 3, n) for the sampler rounds (profiles/r02_sampler_precision.md)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as an error code, never as a hung GPU.  On a timeout the waiter
// records its tag in the device error word and raises the CTA's abort flag (shared memory); every other wait polls
// that flag, so the kernel drains in microseconds and hold_ctx_check() reports HOLD_E_STATE.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err, int tag, volatile int* abort_flag) {
  uint32_t done = 0;
  for (unsigned spin = 0; spin < (1u << 22); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return true;
    if ((spin & 63) == 63 && *abort_flag) return false;
  }
  if (err != nullptr) atomicOr(err, 0x100 | (tag << 12));
  *abort_flag = 1;
  return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One lane of a fully converged warp.  The single-thread roles (bulk copies, tcgen05.mma / commit) run their loops
// with ALL lanes and elect the issuing lane per instruction group: with warp-uniform control flow the compiler keeps
// descriptors and barrier addresses in uniform registers; inside an `if (lane == 0)` region it wraps every UTCHMMA /
// UTCBAR / UBLKCP in an ELECT + R2UR.BROADCAST + BRA.U.ANY loop (~25 instructions each), which made the issuer
// thread the bottleneck of the whole kernel (~1 k clk of issue overhead per weight stage).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// UMMA shared-memory descriptor, K-major canonical layouts (cute/arch/mma_sm100_desc.hpp semantics):
// start address >> 4 | LBO (=1, unused for swizzled K-major) | SBO = bytes between 8-row groups | version 1 | layout
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout_type << 61;
  return d;
}
constexpr uint32_t kLayoutSW128 = 2, kLayoutSW64 = 4;
// kind::f16 instruction descriptor: D=f32, A=B=f16 (format 0), both K-major, N=256, M=128
constexpr uint32_t kIdescF16 = (1u << 4) | (0u << 7) | (0u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdescF16N64 = (1u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);   // the same with N = 64

// x = hi + lo: hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    l[i] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of the 16-byte unit holding k = 8j..8j+7 of row r inside one [128 x 64] SW128 K-major A chunk
__device__ __forceinline__ uint32_t a_unit_off(int r, int j) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}

__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// kTcScaleA * Softplus(beta=100)(z) in its overflow-free form max(z,0) + log1p(exp(-|100 z|))/100 (two MUFU ops), from
// zs = kTcScaleA * z (the epilogue's bias FMA produces zs directly: accumulator scale and bias carry the power-of-two factor,
// so every value is bit-identical to scaling afterwards).  nn.Softplus' threshold branch (returns z for 100 z > 20) differs
// from this by < 2.1e-11.  u_out = exp(-|100 z|) for the derivative.
__device__ __forceinline__ float softplus100_scaled(float zs, float& u_out) {
  const float t = zs * (100.0f * 1.4426950408889634f / kTcScaleA);
  const float u = mufu_ex2(-fabsf(t));
  u_out = u;
  const float L = mufu_lg2(1.0f + u);
  return fmaf(L, 0.6931471805599453f * 0.01f * kTcScaleA, fmaxf(zs, 0.f));
}

// unorm16 stash of softplus' in [0, 1]: encode by the magic-number add (round to nearest, no F2I), decode by OR-ing the
// 16 bits into the mantissa of 2^23.  q = round(65535 s); decode returns q as a float (the caller folds 1/65535 into its scale).
__device__ __forceinline__ uint32_t unorm16_pack2(float s0, float s1) {
  const uint32_t b0 = __float_as_uint(fmaf(s0, 65535.0f, 8388608.0f)), b1 = __float_as_uint(fmaf(s1, 65535.0f, 8388608.0f));
  return __byte_perm(b0, b1, 0x5410);
}
__device__ __forceinline__ float unorm16_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410)) - 8388608.0f; }
__device__ __forceinline__ float unorm16_hi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432)) - 8388608.0f; }

// (Measured, profiles/r02_epilogue_experiments.md: making this and the head dot products out-of-line functions behind one
// shared epilogue body -- to shrink the kernel from 49 KB to 28 KB of code -- was 5 % SLOWER than one specialised body per layer
// kind with everything inlined: the branchy shared body costs more per hand-off than the instruction cache misses it avoids.)
struct Embed8 { float v[8]; uint32_t ok; };
template <int D, bool DERIV>
__device__ __forceinline__ Embed8 embed8(int e0, int n_embed, float x0, float x1, float x2, float x3, const float* __restrict__ ew) {
  Embed8 r;
  if (e0 >= n_embed || e0 + 8 <= 0) {   // (warp-uniform) no element of this group exists
#pragma unroll
    for (int i = 0; i < 8; ++i) r.v[i] = 0.f;
    r.ok = 0;
    return r;
  }
  r.ok = embed8_inl<D, DERIV>(e0, n_embed, x0, x1, x2, x3, ew, r.v);
  return r;
}
// the 1- or 3-row output heads: fp32 dot products of 8 activation columns with the head rows
__device__ __forceinline__ float head_dot8(const float* __restrict__ w, float o0, float o1, float o2, float o3, float o4, float o5, float o6, float o7) {
  const float4 w0 = __ldg(reinterpret_cast<const float4*>(w)), w1 = __ldg(reinterpret_cast<const float4*>(w) + 1);
  return (o0 * w0.x + o1 * w0.y + o2 * w0.z + o3 * w0.w) + (o4 * w1.x + o5 * w1.y + o6 * w1.z + o7 * w1.w);
}

constexpr int kTcW = 4;                        // epilogue warps per TMEM lane quarter
constexpr int kTcChunk = 32;                   // accumulator columns per epilogue->MMA hand-off (= one weight stage of k)
constexpr int kTcCW = kTcChunk / kTcW;         // accumulator columns per warp per chunk
constexpr int kTcEpiWarps = 4 * kTcW;
constexpr int kTcEpiThreads = 32 * kTcEpiWarps;
constexpr int kTcThreadsTotal = 64 + kTcEpiThreads;

__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// One arrival per epilogue WARP on the hand-off barrier: every lane orders its own st.shared against the async
// proxy, the warp converges, lane 0 arrives.
__device__ __forceinline__ void handoff_arrive(uint32_t bar, int lane) {
#if !defined(HOLD_TC_EXP) || (HOLD_TC_EXP != 3 && HOLD_TC_EXP != 6)
  fence_proxy_async();
#endif
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory"); }

template <int MODE>
struct TcCfg {
  static constexpr bool kColorLike = (MODE == MLP_COLOR || MODE == MLP_BG_RGB);   // ReLU chain with a 3-row sigmoid head
  static constexpr bool kWide = kColorLike || MODE == MLP_LINEAR;                  // first operand up to 320 wide
  static constexpr int kAChunks = kWide ? 5 : 4;   // 64-wide SW128 A chunks in smem
  static constexpr int kHandoffs = 2 * kAChunks;                 // 32-wide epilogue->MMA hand-offs
  static constexpr int kStages = kWide ? 2 : 3;
  static constexpr int kSmemA = 2 * kAChunks * kTcAChunkBytes;
  static constexpr int kSmemW = kStages * kTcStageBytes;
  static constexpr int kSmemBytes = kSmemA + kSmemW + 256 + 1024;  // + barriers (<= 2*3 + 10 + 8, 8 B each) + 1 KB alignment slack
};

// Per-thread state of an epilogue warp (kept in registers: every user is force-inlined).
struct EpiState {
  uint32_t bAReady, bDFull, t_lane;
  uint8_t* a_row;        // generic pointer to this thread's row inside A chunk 0 (hi part)
  uint32_t u0, lo_off;   // byte offset of the thread's 16-byte unit for even hand-offs (odd: u0 ^ 64); hi -> lo part distance
  int sub, lane, p;
  bool valid;
  float px, py, pz, pw;
  float head0, head1, head2;
  uint32_t d_par;        // bit b = parity to wait for on d_full[b]
  volatile int* abort_flag;
};

// One layer of a forward chain (every mode but the reverse-mode gradient and hold_linear): wait for the accumulator, then 8
// hand-offs of 32 columns; per hand-off this warp turns its 8 columns into the next layer's fp16 hi/lo operand.
//   HEAD : the 1- (sdf) or 3-row (colour) output head is accumulated from this layer's activations (fp32 dot products)
//   WRITE: the activations are the next MMA's operand (false on the chain's last MMA layer)
//   FEAT : the layer's outputs are the 256-d feature vector (no activation, stored to global memory)
//   SKIP : output columns >= a.L[l].N are the point's embedding (the skip connection into the next layer, shape_net.py:116-119)
// One specialised body per layer kind (see the note at embed8).  a.L[l].bias is pre-multiplied by kTcScaleA unless FEAT
// (TcMlp::bias_s).  The hand-off loop is unrolled by two with explicit ping-pong registers for the prefetched bias and
// accumulator columns (a rotating prefetch costs 8 MOVs per hand-off).  The accumulator arrives in four 64-column quarters.
template <int MODE, bool HEAD, bool WRITE, bool FEAT, bool SKIP>
__device__ __forceinline__ bool epi_layer(const TcArgs& a, EpiState& E, const int l) {
  constexpr bool kColorLike = (MODE == MLP_COLOR || MODE == MLP_BG_RGB);
  const int N = a.L[l].N;
  const float* bias = a.L[l].bias + E.sub * 8;
  float4 nb[2][2];
  nb[0][0] = __ldg(reinterpret_cast<const float4*>(bias));      // issued before the wait
  nb[0][1] = __ldg(reinterpret_cast<const float4*>(bias) + 1);
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 7
  const bool rec = blockIdx.x == 0 && (E.p / kTcRows) == 2 * (int)gridDim.x && E.sub == 0 && (threadIdx.x & 127) == 64 && a.sig != nullptr;
  uint32_t* rbuf = reinterpret_cast<uint32_t*>(a.sig) + l * 16;
  if (rec) rbuf[0] = (uint32_t)clock();
#endif
  const uint32_t dbar = E.bDFull + 32 * (l & 1), dpar = (E.d_par >> (l & 1)) & 1;   // four quarter barriers, one parity (each completes once per layer)
  if (!mbar_wait(dbar, dpar, a.err, 4, E.abort_flag)) return false;
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 7
  if (rec) rbuf[1] = (uint32_t)clock();
#endif
  E.d_par ^= (1u << (l & 1));
  tc_fence_after();
  const uint32_t t_col = E.t_lane + (uint32_t)((l & 1) * 256 + E.sub * 8);
  const float us = FEAT ? a.unscale : a.unscale * kTcScaleA;
  uint32_t raw[2][8];
  tc_ld8(t_col, raw[0]);
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 12   // timing experiment: four hand-offs per loop iteration
#pragma unroll 2
#else
#pragma unroll 1
#endif
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int h = 2 * c + hb;
      const int n0 = h * 32 + E.sub * 8;
      if (hb == 0 || c < 3) {   // the next hand-off's bias: an un-prefetched load is a cache round trip on every hand-off's critical path
        nb[hb ^ 1][0] = __ldg(reinterpret_cast<const float4*>(bias + 32 * (h + 1)));
        nb[hb ^ 1][1] = __ldg(reinterpret_cast<const float4*>(bias + 32 * (h + 1)) + 1);
      }
      const float bv[8] = {nb[hb][0].x, nb[hb][0].y, nb[hb][0].z, nb[hb][0].w, nb[hb][1].x, nb[hb][1].y, nb[hb][1].z, nb[hb][1].w};
      tc_wait_ld();
      if (hb == 1 && c < 3) {   // the next hand-off opens the accumulator's next 64-column quarter
        if (!mbar_wait(dbar + 8 * (c + 1), dpar, a.err, 4, E.abort_flag)) return false;
        tc_fence_after();
      }
      if (hb == 0 || c < 3) tc_ld8(t_col + (uint32_t)((h + 1) * 32), raw[hb ^ 1]);   // prefetch the next hand-off's columns
      float out[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        // accumulator -> pre-activation times kTcScaleA (undo the weight scaling, add the scaled bias); out[] is the next
        // layer's operand, i.e. the activation times kTcScaleA (FEAT: the plain output)
        const float zs = fmaf(__uint_as_float(raw[hb][i]), us, bv[i]);
        float e;
        out[i] = FEAT ? zs : (kColorLike ? fmaxf(zs, 0.f) : softplus100_scaled(zs, e));
      }
      if (SKIP && n0 + 8 > N) {  // skip connection: embedding columns of layer 3's output
        const Embed8 ev = (MODE == MLP_BG_SDF) ? embed8<4, false>(n0 - N, kBgEmbed, E.px, E.py, E.pz, E.pw, nullptr)
                                               : embed8<3, false>(n0 - N, kEmbed, E.px, E.py, E.pz, 0.f, a.embed_w);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = (n0 + i >= N) ? (((ev.ok >> i) & 1u) ? kTcScaleA * ev.v[i] : 0.f) : out[i];
      }
      if (HEAD) {
        E.head0 += head_dot8(a.w_last + n0, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
        if (kColorLike) {
          E.head1 += head_dot8(a.w_last + 256 + n0, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
          E.head2 += head_dot8(a.w_last + 512 + n0, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
        }
      }
      if (FEAT) {
        if (E.valid) {
          float4* dst = reinterpret_cast<float4*>(a.feat + (size_t)E.p * kFeat + n0);
          dst[0] = make_float4(out[0], out[1], out[2], out[3]);
          dst[1] = make_float4(out[4], out[5], out[6], out[7]);
        }
      } else if (WRITE) {
        uint4 hi, lo;
        split8(out, hi, lo);
        uint8_t* dst = E.a_row + c * kTcAChunkBytes + (hb ? (E.u0 ^ 64u) : E.u0);
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + E.lo_off) = lo;
        handoff_arrive(E.bAReady + 16 * c + 8 * hb, E.lane);
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 7
        if (rec) rbuf[2 + h] = (uint32_t)clock();
#endif
      }
    }
  }
  return true;
}

// Reverse-mode chain (MODE MLP_SDF_REV), one step = one accumulator: d sdf/d z_7 = w_sdf * s_7;  d sdf/d z_{l-1} = (g_l . W_l) * s_{l-1};
// the embedding columns (skip input of layer 4, input of layer 0) collect d sdf/d embed, chained with d embed/d x_c per column.
//   KIND 0: forward layer l = st (stashes s_l = softplus'(z_l));  SPEC 1: l = 3 (skip columns), SPEC 2: l = 7 (sdf head)
//   KIND 1: the feature layer (st = 8): writes the 256-d feature, seeds g_7 = w_sdf * s_7
//   KIND 2: backward layer l = 16 - st: operand of the next step = acc * s_{l-1};  SPEC 1: l = 4 (its columns >= 217 are d sdf/d embed)
//   KIND 3: st = 16: d sdf / d embed through layer 0's input, no operand
// One specialised body per (KIND, SPEC), same structure as epi_layer (ping-pong registers for the prefetched bias / accumulator
// columns; the stash words are read two hand-offs ahead into the slot that was just consumed).  The monolithic step body this
// replaces issued 223 instructions per 8-column hand-off (ncu) against 107 in the forward-only chain.
// Stash: s_l as unorm16, written and later read by the SAME thread (row, 8 columns per hand-off), L2 only (.cg): 512 KB per CTA.
struct RevState {
  uint16_t* sig;         // this thread's row of the CTA's stash, [8 layers][128 rows][256]
  float gz;              // d sdf / d z (head1 / head2 of the EpiState hold d/dx, d/dy)
};

template <int KIND, int SPEC>
__device__ __forceinline__ bool rev_step(const TcArgs& a, EpiState& E, RevState& R, const int st) {
  constexpr float kInvQ = 1.0f / 65535.0f;
  constexpr int kRowStride = kTcRows * 256;
  const int l = (st <= 8) ? st : 16 - st;
  const float* bias = (KIND <= 1) ? a.L[st].bias + E.sub * 8 : nullptr;
  // stash row read by this step: s_7 for the feature layer's seed, s_{l-1} for backward layer l
  const uint16_t* srow = (KIND == 1) ? R.sig + (size_t)7 * kRowStride + E.sub * 8 : ((KIND == 2) ? R.sig + (size_t)(l - 1) * kRowStride + E.sub * 8 : nullptr);
  float4 nb[2][2];
  uint4 sq[2];
  if (KIND <= 1) {
    nb[0][0] = __ldg(reinterpret_cast<const float4*>(bias));
    nb[0][1] = __ldg(reinterpret_cast<const float4*>(bias) + 1);
  }
  if (KIND == 1 || KIND == 2) {
    sq[0] = __ldcg(reinterpret_cast<const uint4*>(srow));
    sq[1] = __ldcg(reinterpret_cast<const uint4*>(srow + 32));
  }
  const uint32_t dbar = E.bDFull + 32 * (st & 1), dpar = (E.d_par >> (st & 1)) & 1;
  if (!mbar_wait(dbar, dpar, a.err, 4, E.abort_flag)) return false;
  E.d_par ^= (1u << (st & 1));
  tc_fence_after();
  const uint32_t t_col = E.t_lane + (uint32_t)((st & 1) * 256 + E.sub * 8);
  const float us = (KIND == 0) ? a.unscale * kTcScaleA : a.unscale;   // forward layers work on kTcScaleA * z (bias pre-scaled)
  uint32_t raw[2][8];
  tc_ld8(t_col, raw[0]);
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 12   // timing experiment: four hand-offs per loop iteration
#pragma unroll 2
#else
#pragma unroll 1
#endif
  for (int c = 0; c < 4; ++c) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int h = 2 * c + hb;
      const int n0 = h * 32 + E.sub * 8;
      if (KIND <= 1 && (hb == 0 || c < 3)) {
        nb[hb ^ 1][0] = __ldg(reinterpret_cast<const float4*>(bias + 32 * (h + 1)));
        nb[hb ^ 1][1] = __ldg(reinterpret_cast<const float4*>(bias + 32 * (h + 1)) + 1);
      }
      const float bv[8] = {nb[hb][0].x, nb[hb][0].y, nb[hb][0].z, nb[hb][0].w, nb[hb][1].x, nb[hb][1].y, nb[hb][1].z, nb[hb][1].w};
      const uint4 sqv = sq[hb];
      if ((KIND == 1 || KIND == 2) && c < 3) sq[hb] = __ldcg(reinterpret_cast<const uint4*>(srow + 32 * (h + 2)));   // two hand-offs ahead
      tc_wait_ld();
      if (hb == 1 && c < 3) {   // the next hand-off opens the accumulator's next 64-column quarter
        if (!mbar_wait(dbar + 8 * (c + 1), dpar, a.err, 4, E.abort_flag)) return false;
        tc_fence_after();
      }
      if (hb == 0 || c < 3) tc_ld8(t_col + (uint32_t)((h + 1) * 32), raw[hb ^ 1]);
      float out[8];
      if (KIND == 0) {
        float sg[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float e;
          const float zs = fmaf(__uint_as_float(raw[hb][i]), us, bv[i]);
          out[i] = softplus100_scaled(zs, e);
          const float r = mufu_rcp(1.0f + e);
          sg[i] = (zs >= 0.f) ? r : e * r;
        }
        __stcg(reinterpret_cast<uint4*>(R.sig + (size_t)l * kRowStride + n0),
               make_uint4(unorm16_pack2(sg[0], sg[1]), unorm16_pack2(sg[2], sg[3]), unorm16_pack2(sg[4], sg[5]), unorm16_pack2(sg[6], sg[7])));
        if (SPEC == 1 && n0 + 8 > 217) {
          const Embed8 ev = embed8<3, false>(n0 - 217, kEmbed, E.px, E.py, E.pz, 0.f, a.embed_w);
#pragma unroll
          for (int i = 0; i < 8; ++i) out[i] = ((ev.ok >> i) & 1u) ? kTcScaleA * ev.v[i] : out[i];
        }
        if (SPEC == 2) E.head0 += head_dot8(a.w_last + n0, out[0], out[1], out[2], out[3], out[4], out[5], out[6], out[7]);
      } else {
        // softplus' of this hand-off's 8 columns (times 65535; the scale is folded below)
        const float sv[8] = {unorm16_lo(sqv.x), unorm16_hi(sqv.x), unorm16_lo(sqv.y), unorm16_hi(sqv.y),
                             unorm16_lo(sqv.z), unorm16_hi(sqv.z), unorm16_lo(sqv.w), unorm16_hi(sqv.w)};
        if (KIND == 1) {
          if (E.valid) {   // the 256-d feature vector: written once, read once by the colour net -> streaming stores
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaf(__uint_as_float(raw[hb][i]), us, bv[i]);
            float4* dst = reinterpret_cast<float4*>(a.feat + (size_t)E.p * kFeat + n0);
            __stcs(dst, make_float4(f[0], f[1], f[2], f[3]));
            __stcs(dst + 1, make_float4(f[4], f[5], f[6], f[7]));
          }
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w_last + n0));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w_last + n0) + 1);
          const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          // g_7 = w_sdf * s_7, times the operand scale
#pragma unroll
          for (int i = 0; i < 8; ++i) out[i] = (kTcScaleA * kInvQ) * wv[i] * sv[i];
        } else {
          if (KIND == 2) {   // operand of the next backward layer: (acc * unscale) * s_{l-1}, times the operand scale (one constant)
            const float k2 = us * (kTcScaleA * kInvQ);
#pragma unroll
            for (int i = 0; i < 8; ++i) out[i] = (__uint_as_float(raw[hb][i]) * k2) * sv[i];
          }
          if ((KIND == 2 && SPEC == 1 && n0 + 8 > 217) || (KIND == 3 && n0 < 40)) {  // columns that are d sdf / d embed
            const int e0 = (KIND == 2) ? n0 - 217 : n0;
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(raw[hb][i]) * us;
            const Embed8 dv = embed8<3, true>(e0, kEmbed, E.px, E.py, E.pz, 0.f, a.embed_w);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int ec = max(e0 + i, 0), d = ec - 3 * ((ec * 171) >> 9);
              const bool ok = (dv.ok >> i) & 1u;
              const float je = ok ? acc[i] * dv.v[i] : 0.f;
              E.head1 += (d == 0) ? je : 0.f;
              E.head2 += (d == 1) ? je : 0.f;
              R.gz += (d == 2) ? je : 0.f;
              if (KIND == 2) out[i] = ok ? 0.f : out[i];
            }
          }
        }
      }
      if (KIND != 3) {
        uint4 hi, lo;
        split8(out, hi, lo);
        uint8_t* dst = E.a_row + c * kTcAChunkBytes + (hb ? (E.u0 ^ 64u) : E.u0);
        *reinterpret_cast<uint4*>(dst) = hi;
        *reinterpret_cast<uint4*>(dst + E.lo_off) = lo;
        handoff_arrive(E.bAReady + 16 * c + 8 * hb, E.lane);
      }
    }
  }
  return true;
}

template <int MODE>
__global__ void __launch_bounds__(kTcThreadsTotal, 1) k_mlp_tc(TcArgs a) {
  static_assert(MODE == MLP_SDF_ONLY || MODE == MLP_SDF_REV || MODE == MLP_COLOR || MODE == MLP_BG_SDF || MODE == MLP_BG_RGB || MODE == MLP_LINEAR, "chains built on tcgen05");
  if (a.st != nullptr && a.st->done) return;
  using Cfg = TcCfg<MODE>;
  constexpr bool kColorLike = Cfg::kColorLike;
  constexpr int NA = Cfg::kAChunks, NS = Cfg::kStages, NHO = Cfg::kHandoffs;
  constexpr int PPT = kTcRows;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA_hi = base, sA_lo = base + NA * kTcAChunkBytes, sW = base + Cfg::kSmemA;
  const uint32_t sBar = sW + Cfg::kSmemW;
  const uint32_t bWFull = sBar, bWEmpty = sBar + 8 * NS, bAReady = sBar + 16 * NS, bDFull = bAReady + 8 * NHO;
  const uint32_t sTmemPtr = bDFull + 64, sAbort = bDFull + 68;   // d_full: [2 accumulators][4 column quarters]
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to `base`
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gen_base + (sAbort - base));
  const int n_tiles = ceil_div(a.P, PPT);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bWFull + 8 * i, 1); mbar_init(bWEmpty + 8 * i, 1); }
    *abort_flag = 0;
    for (int i = 0; i < NHO; ++i) mbar_init(bAReady + 8 * i, kTcEpiWarps);  // one arrival per epilogue warp
    for (int i = 0; i < 8; ++i) mbar_init(bDFull + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemPtr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + (sTmemPtr - base));

  if (warp == 0) {
    // ============================================================ weight producer (TMA engine, bulk async copies)
    {
      uint32_t stage = 0, phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < a.n_layers; ++l) {
          const uint8_t* src = a.L[l].wimg;
          for (int s = 0; s < a.L[l].nst; ++s) {
            if (!__all_sync(0xffffffffu, mbar_wait(bWEmpty + 8 * stage, phase ^ 1, a.err, 1, abort_flag))) goto tc_done;
            if (elect_one()) {
              mbar_expect_tx(bWFull + 8 * stage, kTcStageBytes);
              bulk_g2s(sW + stage * kTcStageBytes, src + (size_t)s * kTcStageBytes, kTcStageBytes, bWFull + 8 * stage);
            }
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer (whole warp walks the loop, one elected lane issues)
    {
      uint32_t stage = 0, phase = 0;
      uint32_t a_par = 0;  // bit c = parity to wait for on a_ready[c]
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int l = 0; l < a.n_layers; ++l) {
          const uint32_t d_tmem = tmem_u + (uint32_t)((l & 1) * 256);
          const int nst = a.L[l].nst;
          for (int s = 0; s < nst; ++s) {
            const int c = s >> 1;  // 64-wide A chunk holding this 32-k stage
            // hand-off s = columns [32 s, 32 s + 32) of the previous layer's activations
            if (!__all_sync(0xffffffffu, mbar_wait(bAReady + 8 * s, (a_par >> s) & 1, a.err, 2, abort_flag))) goto tc_done;
            a_par ^= (1u << s);
            if (!__all_sync(0xffffffffu, mbar_wait(bWFull + 8 * stage, phase, a.err, 3, abort_flag))) goto tc_done;
            tc_fence_after();
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 7   // cycle accounting (tools/exp_epilogue.py): stage s of layer l issued at
            if (blockIdx.x == 0 && tile == 2 * (int)gridDim.x && lane == 0 && a.sig != nullptr) reinterpret_cast<uint32_t*>(a.sig)[256 + l * 16 + s] = (uint32_t)clock();
#endif
            const uint32_t wb = sW + stage * kTcStageBytes;
            const bool el = elect_one();
            if (s + 1 < nst) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint32_t koff = (uint32_t)(((s & 1) * 2 + j) * 32);  // bytes inside the 128-byte A row
                const uint64_t ahi = umma_desc(sA_hi + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
                const uint64_t alo = umma_desc(sA_lo + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
                const uint64_t whi = umma_desc(wb + j * 32, 512, kLayoutSW64);
                const uint64_t wlo = umma_desc(wb + 16384 + j * 32, 512, kLayoutSW64);
                if (el) {
                  tc_mma(d_tmem, ahi, whi, kIdescF16, (s | j) != 0);
                  if (a.passes >= 2) tc_mma(d_tmem, alo, whi, kIdescF16, 1);
                  if (a.passes >= 3) tc_mma(d_tmem, ahi, wlo, kIdescF16, 1);
                }
              }
            } else {
              // the layer's LAST k stage goes out as four 64-column quarters, each committed on its own barrier: the epilogue
              // starts on columns 0..63 while the other three quarters are still in the tensor pipe (measured: the accumulator-
              // full signal sat ~1.1 k clocks after the last stage's issue, all of it tensor-pipe idle time of the next layer)
#pragma unroll
              for (int qn = 0; qn < 4; ++qn) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint32_t koff = (uint32_t)(((s & 1) * 2 + j) * 32);
                  const uint64_t ahi = umma_desc(sA_hi + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
                  const uint64_t alo = umma_desc(sA_lo + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
                  const uint64_t whi = umma_desc(wb + qn * 4096 + j * 32, 512, kLayoutSW64);          // rows 64 qn.. of the [256 n x 32 k] image
                  const uint64_t wlo = umma_desc(wb + 16384 + qn * 4096 + j * 32, 512, kLayoutSW64);
                  if (el) {
                    tc_mma(d_tmem + (uint32_t)(64 * qn), ahi, whi, kIdescF16N64, (s | j) != 0);
                    if (a.passes >= 2) tc_mma(d_tmem + (uint32_t)(64 * qn), alo, whi, kIdescF16N64, 1);
                    if (a.passes >= 3) tc_mma(d_tmem + (uint32_t)(64 * qn), ahi, wlo, kIdescF16N64, 1);
                  }
                }
                if (el) tc_commit(bDFull + 8 * ((l & 1) * 4 + qn));  // columns [64 qn, 64 qn + 64) of layer l's accumulator complete
              }
            }
            if (el) tc_commit(bWEmpty + 8 * stage);  // frees the weight stage when these MMAs have read it
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ============================================================ epilogue: kTcW warps per TMEM lane quarter; warp
    // `sub` of a quarter owns columns [sub*CW, (sub+1)*CW) of every 64-column chunk, so chunks complete in order
    // and layer l+1's MMAs on chunk c start while chunks c+1.. of layer l are still in the epilogue.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* gA_hi = gen_base;
    uint8_t* gA_lo = gen_base + NA * kTcAChunkBytes;
    float* scratch = reinterpret_cast<float*>(gen_base);  // head partial sums (A region, free at tile end)
    uint32_t d_par = 0;  // bit b = parity to wait for on d_full[b]
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int p = tile * PPT + row;
      const bool valid = p < a.P;
      float px = 0.f, py = 0.f, pz = 0.f, pw = 0.f;
      // ---------------------------------------------------------- prologue: layer-0 A operand (this warp's columns)
      if (MODE == MLP_LINEAR) {
        // rows of a [P, lda] fp32 matrix, columns [0, kvalid) (zero padded to the image's K), optionally divided by *in_scale
        const float isc = (a.in_scale != nullptr) ? kTcScaleA / __ldg(a.in_scale) : kTcScaleA;
        const float* src = a.lin_in + (size_t)p * a.lda;
        const int nst = a.L[0].nst;
        for (int h = 0; h < nst; ++h) {
          const int k0 = h * 32 + sub * 8;
          float x[8];
          if (valid && k0 + 8 <= a.kvalid) {
            const float4 f0 = __ldg(reinterpret_cast<const float4*>(src + k0)), f1 = __ldg(reinterpret_cast<const float4*>(src + k0) + 1);
            x[0] = f0.x, x[1] = f0.y, x[2] = f0.z, x[3] = f0.w, x[4] = f1.x, x[5] = f1.y, x[6] = f1.z, x[7] = f1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = (valid && k0 + i < a.kvalid) ? __ldg(src + k0 + i) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] *= isc;
          uint4 hi, lo;
          split8(x, hi, lo);
          const int c = h >> 1, j = (h & 1) * 4 + sub;
          *reinterpret_cast<uint4*>(gA_hi + c * kTcAChunkBytes + a_unit_off(row, j)) = hi;
          *reinterpret_cast<uint4*>(gA_lo + c * kTcAChunkBytes + a_unit_off(row, j)) = lo;
          handoff_arrive(bAReady + 8 * h, lane);
        }
      } else if (MODE == MLP_BG_SDF) {
        // inverted-sphere point of sample p % 32 of ray p / 32 (background.py:63-68,102-135), PE-10 + frame code: 116 -> 128
        if (valid) {
          const int ray = p / kBgN;
          const float o[3] = {a.cam[3 * ray], a.cam[3 * ray + 1], a.cam[3 * ray + 2]};
          const float d[3] = {a.dirs[3 * ray], a.dirs[3 * ray + 1], a.dirs[3 * ray + 2]};
          float p4[4];
          depth2pts_outside(o, d, bg_depth(p % kBgN, a.r_sphere), a.r_sphere, p4);
          px = p4[0], py = p4[1], pz = p4[2], pw = p4[3];
        }
        const int b = valid ? p / a.pts_per_frame : 0;
        for (int h = 0; h < 4; ++h) {
          float x[8];
          const int e0 = h * 32 + sub * 8;
          const Embed8 ev = embed8<4, false>(e0, kBgEmbed, px, py, pz, pw, nullptr);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int e = e0 + i;
            float v = ((ev.ok >> i) & 1u) ? ev.v[i] : 0.f;
            if (e >= kBgEmbed && e < kBgEmbed + kBgFrame) v = a.frame_code[b * kBgFrame + e - kBgEmbed];
            x[i] = kTcScaleA * v;
          }
          uint4 hi, lo;
          split8(x, hi, lo);
          const int c = h >> 1, j = (h & 1) * 4 + sub;
          *reinterpret_cast<uint4*>(gA_hi + c * kTcAChunkBytes + a_unit_off(row, j)) = hi;
          *reinterpret_cast<uint4*>(gA_lo + c * kTcAChunkBytes + a_unit_off(row, j)) = lo;
          handoff_arrive(bAReady + 8 * h, lane);
        }
      } else if (!kColorLike) {
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 7
        if (blockIdx.x == 0 && tile == 2 * (int)gridDim.x && sub == 0 && (threadIdx.x & 127) == 64 && a.sig != nullptr) reinterpret_cast<uint32_t*>(a.sig)[12] = (uint32_t)clock();
#endif
        if (valid) { px = a.xc[3 * (size_t)p], py = a.xc[3 * (size_t)p + 1], pz = a.xc[3 * (size_t)p + 2]; }
        for (int h = 0; h < 2; ++h) {
          float x[8];
          const Embed8 ev = embed8<3, false>(h * 32 + sub * 8, kEmbed, px, py, pz, 0.f, a.embed_w);
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = ((ev.ok >> i) & 1u) ? kTcScaleA * ev.v[i] : 0.f;
          uint4 hi, lo;
          split8(x, hi, lo);
          const int j = h * 4 + sub;
          *reinterpret_cast<uint4*>(gA_hi + a_unit_off(row, j)) = hi;
          *reinterpret_cast<uint4*>(gA_lo + a_unit_off(row, j)) = lo;
          handoff_arrive(bAReady + 8 * h, lane);
        }
      } else {
        const int b = valid ? p / a.pts_per_frame : 0;
        for (int h = 0; h < 10; ++h) {
          float x[8];
          const int k0 = h * 32 + sub * 8;
          if (h < 8) {
            if (valid) {
              const float4 f0 = *reinterpret_cast<const float4*>(a.feat + (size_t)p * kFeat + k0);
              const float4 f1 = *reinterpret_cast<const float4*>(a.feat + (size_t)p * kFeat + k0 + 4);
              x[0] = f0.x, x[1] = f0.y, x[2] = f0.z, x[3] = f0.w, x[4] = f1.x, x[5] = f1.y, x[6] = f1.z, x[7] = f1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = 0.f;
            }
          } else {
            Embed8 ev;
            ev.ok = 0;
            if (MODE == MLP_BG_RGB) {
              const int ray = valid ? p / kBgN : 0;
              ev = embed8<3, false>(k0 - kFeat, kBgView, a.dirs[3 * ray], a.dirs[3 * ray + 1], a.dirs[3 * ray + 2], 0.f, nullptr);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int e = k0 + i - kFeat;  // [x_c(3), n(3), pose_embed(8), time_code(32)]; background: [view PE-4 (27), frame code (32)]
              float v = 0.f;
              if (MODE == MLP_BG_RGB) {
                if (valid) {
                  if ((ev.ok >> i) & 1u) v = ev.v[i];
                  else if (e >= kBgView && e < kBgView + kBgFrame) v = a.frame_code[b * kBgFrame + e - kBgView];
                }
              } else if (valid) {
                if (e < 3) v = a.xc[3 * (size_t)p + e];
                else if (e < 6) v = a.normal[3 * (size_t)p + e - 3];
                else if (e < 14) v = (a.pose_embed != nullptr) ? a.pose_embed[b * 8 + e - 6] : 0.f;
                else if (e < a.k0 - kFeat) v = a.time_code[b * 32 + e - 14];
              }
              x[i] = v;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] *= kTcScaleA;
          uint4 hi, lo;
          split8(x, hi, lo);
          const int c = h >> 1, j = (h & 1) * 4 + sub;
          *reinterpret_cast<uint4*>(gA_hi + c * kTcAChunkBytes + a_unit_off(row, j)) = hi;
          *reinterpret_cast<uint4*>(gA_lo + c * kTcAChunkBytes + a_unit_off(row, j)) = lo;
          handoff_arrive(bAReady + 8 * h, lane);
        }
      }
      // ---------------------------------------------------------- per-layer epilogues
      float head0 = 0.f, head1 = 0.f, head2 = 0.f, gz_acc = 0.f;
      if (MODE == MLP_LINEAR) {
        // ======== one layer: C = acc * unscale (* in_scale) + bias, fp32 rows of [P, ldc], columns [0, nvalid) ========
        const float* bias = a.L[0].bias;
        const float osc = (a.in_scale != nullptr) ? a.unscale * __ldg(a.in_scale) : a.unscale;
        if (!mbar_wait(bDFull + 24, d_par & 1, a.err, 4, abort_flag)) break;   // the last column quarter: the whole accumulator is complete
        d_par ^= 1u;
        tc_fence_after();
        const uint32_t t_col = t_lane + (uint32_t)(sub * 8);
        uint32_t raw[8];
        tc_ld8(t_col, raw);
        float* dst = a.lin_out + (size_t)p * a.ldc;
#pragma unroll 2
        for (int h = 0; h < 8; ++h) {
          const int n0 = h * 32 + sub * 8;
          tc_wait_ld();
          float out[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) out[i] = __uint_as_float(raw[i]) * osc;
          if (h + 1 < 8) tc_ld8(t_col + (uint32_t)((h + 1) * 32), raw);
          if (bias != nullptr && n0 < a.nvalid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) out[i] += (n0 + i < a.nvalid) ? __ldg(bias + n0 + i) : 0.f;
          }
          if (valid) {
            if (n0 + 8 <= a.nvalid) {
              *reinterpret_cast<float4*>(dst + n0) = make_float4(out[0], out[1], out[2], out[3]);
              *reinterpret_cast<float4*>(dst + n0 + 4) = make_float4(out[4], out[5], out[6], out[7]);
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (n0 + i < a.nvalid) dst[n0 + i] = out[i];
            }
          }
        }
        tc_fence_before();
        epi_bar();   // every warp has read its accumulator columns before the next tile's MMAs overwrite them
        continue;
      }
      if (MODE == MLP_SDF_REV) {
        // ======== reverse-mode gradient: 8 forward layers (stash softplus'), feature layer, 8 backward layers (rev_step) ========
        {
          EpiState E;
          E.bAReady = bAReady, E.bDFull = bDFull, E.t_lane = t_lane;
          E.a_row = gA_hi + (row >> 3) * 1024 + (row & 7) * 128;
          E.u0 = (uint32_t)((sub ^ (row & 7)) << 4), E.lo_off = (uint32_t)(NA * kTcAChunkBytes);
          E.sub = sub, E.lane = lane, E.p = p, E.valid = valid;
          E.px = px, E.py = py, E.pz = pz, E.pw = 0.f;
          E.head0 = 0.f, E.head1 = 0.f, E.head2 = 0.f, E.d_par = d_par, E.abort_flag = abort_flag;
          RevState R;
          R.sig = a.sig + (size_t)blockIdx.x * (8 * kTcRows * 256) + (size_t)row * 256;
          R.gz = 0.f;
          bool ok = true;
          for (int st = 0; st < 17 && ok; ++st) {
            if (st < 8) ok = (st == 3) ? rev_step<0, 1>(a, E, R, st) : ((st == 7) ? rev_step<0, 2>(a, E, R, st) : rev_step<0, 0>(a, E, R, st));
            else if (st == 8) ok = rev_step<1, 0>(a, E, R, st);
            else if (st < 16) ok = (st == 12) ? rev_step<2, 1>(a, E, R, st) : rev_step<2, 0>(a, E, R, st);
            else ok = rev_step<3, 0>(a, E, R, st);
          }
          d_par = E.d_par, head0 = E.head0, head1 = E.head1, head2 = E.head2, gz_acc = R.gz;
        }
        // fixed-order reduction of (sdf head, d/dx, d/dy, d/dz) over the quarter's 4 warps
        tc_fence_before();
        scratch[(sub * 4 + 0) * kTcRows + row] = head0;
        scratch[(sub * 4 + 1) * kTcRows + row] = head1;
        scratch[(sub * 4 + 2) * kTcRows + row] = head2;
        scratch[(sub * 4 + 3) * kTcRows + row] = gz_acc;
        epi_bar();
        if (sub == 0 && valid) {
          float hsum[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float acc = 0.f;
#pragma unroll
            for (int w = 0; w < kTcW; ++w) acc += scratch[(w * 4 + k) * kTcRows + row];
            hsum[k] = acc;
          }
          a.sdf[p] = hsum[0] * (1.0f / kTcScaleA) + a.b_last[0];
          a.grad[3 * (size_t)p] = hsum[1], a.grad[3 * (size_t)p + 1] = hsum[2], a.grad[3 * (size_t)p + 2] = hsum[3];
        }
        epi_bar();
        continue;
      }
      {
        EpiState E;
        E.bAReady = bAReady, E.bDFull = bDFull, E.t_lane = t_lane;
        E.a_row = gA_hi + (row >> 3) * 1024 + (row & 7) * 128;
        E.u0 = (uint32_t)((sub ^ (row & 7)) << 4), E.lo_off = (uint32_t)(NA * kTcAChunkBytes);
        E.sub = sub, E.lane = lane, E.p = p, E.valid = valid;
        E.px = px, E.py = py, E.pz = pz, E.pw = pw;
        E.head0 = 0.f, E.head1 = 0.f, E.head2 = 0.f, E.d_par = d_par, E.abort_flag = abort_flag;
        bool ok = true;
        for (int l = 0; l < a.n_layers && ok; ++l) {
          const bool head_layer = kColorLike ? (l == a.n_layers - 1) : (l == 7);
          const bool last_mma = (l == a.n_layers - 1);
          if (MODE == MLP_BG_SDF && last_mma) ok = epi_layer<MODE, false, false, true, false>(a, E, l);           // the feature rows
          else if (MODE == MLP_BG_SDF && head_layer) ok = epi_layer<MODE, true, true, false, false>(a, E, l);
          else if (head_layer) ok = epi_layer<MODE, true, false, false, false>(a, E, l);
          else if (!kColorLike && l == 3) ok = epi_layer<MODE, false, true, false, true>(a, E, l);
          else ok = epi_layer<MODE, false, true, false, false>(a, E, l);
        }
        d_par = E.d_par, head0 = E.head0, head1 = E.head1, head2 = E.head2;
        if (!ok) break;
      }
      // ---------------------------------------------------------- heads: fixed-order reduction over the quarter's warps
      // (all MMAs of the tile have completed, so the A region is free to hold the partial sums)
      tc_fence_before();
      constexpr int NH = kColorLike ? 3 : 1;
      scratch[(sub * NH + 0) * kTcRows + row] = head0;
      if (kColorLike) {
        scratch[(sub * NH + 1) * kTcRows + row] = head1;
        scratch[(sub * NH + 2) * kTcRows + row] = head2;
      }
      epi_bar();
      if (sub == 0 && valid) {
        float h[NH];
#pragma unroll
        for (int k = 0; k < NH; ++k) {
          float acc = 0.f;
#pragma unroll
          for (int w = 0; w < kTcW; ++w) acc += scratch[(w * NH + k) * kTcRows + row];
          h[k] = acc * (1.0f / kTcScaleA);  // the head saw activations times kTcScaleA
        }
        if (kColorLike) {
#pragma unroll
          for (int k = 0; k < NH; ++k) a.rgb[3 * (size_t)p + k] = 1.0f / (1.0f + __expf(-(h[k] + a.b_last[k])));
        } else {
          a.sdf[p] = h[0] + a.b_last[0];
        }
      }
      epi_bar();  // scratch is overwritten by the next tile's prologue
    }
  }
tc_done:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ packing
// One stage image = [256 n x 32 k] fp16 in the SW64 K-major canonical layout, hi part then lo part.
// W[n][k] = kTcScaleW * scale * fold(v, g)[row_off + n][colmap(k)];  colmap: k -> source column (or -1 => 0).
__global__ void k_tc_pack(const float* __restrict__ v, const float* __restrict__ g, int in_dim, int row_off, int N, int K,
                          int kpad, float scale, int perm_feat_first, uint8_t* __restrict__ img) {
  const int n = blockIdx.x;  // 0..255
  __shared__ float red[32];
  __shared__ float f_sh;
  float f = 0.f;
  const float* vr = v + (size_t)(row_off + (n < N ? n : 0)) * in_dim;
  if (g != nullptr) {
    float ss = 0.f;
    for (int k = threadIdx.x; k < in_dim; k += blockDim.x) ss += vr[k] * vr[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < blockDim.x / 32; ++w) tot += red[w];
      f_sh = g[row_off + (n < N ? n : 0)] / sqrtf(tot);
    }
    __syncthreads();
    f = f_sh;
  } else {
    f = 1.0f;
  }
  for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
    int src = k;
    if (perm_feat_first) {  // colour lin0: A order [feat(256) | x_c, n, pose(14) | time(32)] vs weight order [14 | 256 | 32];
      // background lin0: A order [feat(256) | view, frame (59)] vs weight order [59 | 256].  perm_feat_first = the leading width
      if (k < kFeat) src = perm_feat_first + k;
      else if (k < kFeat + perm_feat_first) src = k - kFeat;
      else src = k;
    }
    float w = (n < N && src < K) ? kTcScaleW * (scale * (vr[src] * f)) : 0.f;
    __half h = __float2half_rn(w);
    __half l = __float2half_rn(w - __half2float(h));
    const int st = k >> 5, kk = k & 31;
    const size_t off = (size_t)st * kTcStageBytes + (size_t)((n >> 3) * 512 + (n & 7) * 64 + ((((kk >> 3) ^ ((n >> 1) & 3))) << 4) + (kk & 7) * 2);
    *reinterpret_cast<__half*>(img + off) = h;
    *reinterpret_cast<__half*>(img + off + 16384) = l;
  }
}

// Transposed image for the reverse-mode gradient: B rows = input index k_in of layer l, K = output index n_out:
// WT[r][n_out] = kTcScaleW * scale * fold(v, g)[row_off + n_out][src(r)]  (zero outside [N_out) x valid sources).
// colmap: 0: src = r (r < K_in);  1: a first layer whose 256 feature inputs follow `lead` other inputs (colour lin0: lead = 14 [x_c, n,
//         pose]; background colour lin0: lead = 59 [view, frame code]), feature part: src = lead + r (r < 256);
//         2: the same layer's other inputs in the A operand's order [leading inputs | inputs after the features]: src = r (r < lead),
//         256 + r (256 + r < K_in)
__global__ void k_tc_pack_T(const float* __restrict__ v, const float* __restrict__ g, int in_dim, int row_off, int N_out, int K_in,
                            int colmap, int lead, float scale, uint8_t* __restrict__ img) {
  const int r = blockIdx.x;   // k_in, 0..255
  const int n = threadIdx.x;  // n_out, 0..255
  int src = -1;
  if (colmap == 0) src = (r < K_in) ? r : -1;
  else if (colmap == 1) src = lead + r;
  else src = (r < lead) ? r : ((256 + r < K_in) ? 256 + r : -1);
  float w = 0.f;
  if (n < N_out && src >= 0) {
    const float* vr = v + (size_t)(row_off + n) * in_dim;
    float f = 1.0f;
    if (g != nullptr) {
      float ss = 0.f;
      for (int k = 0; k < in_dim; ++k) ss += vr[k] * vr[k];
      f = g[row_off + n] / sqrtf(ss);
    }
    w = kTcScaleW * (scale * (vr[src] * f));
  }
  const __half h = __float2half_rn(w);
  const __half l = __float2half_rn(w - __half2float(h));
  const int st = n >> 5, kk = n & 31;
  const size_t off = (size_t)st * kTcStageBytes + (size_t)((r >> 3) * 512 + (r & 7) * 64 + ((((kk >> 3) ^ ((r >> 1) & 3))) << 4) + (kk & 7) * 2);
  *reinterpret_cast<__half*>(img + off) = h;
  *reinterpret_cast<__half*>(img + off + 16384) = l;
}

static int tc_init(hold_ctx*) {
  cudaError_t e;
  e = cudaFuncSetAttribute(k_mlp_tc<MLP_SDF_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<MLP_SDF_ONLY>::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_tc<MLP_COLOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<MLP_COLOR>::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_tc<MLP_SDF_REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<MLP_SDF_REV>::kSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_tc<MLP_LINEAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<MLP_LINEAR>::kSmemBytes);
  if (e != cudaSuccess) { set_error("tcgen05 kernel attribute: %s", cudaGetErrorString(e)); return HOLD_E_CUDA; }
  return HOLD_OK;
}

static void tc_free(NodeState& ns) {
  if (!ns.tc) return;
  for (int l = 0; l < HOLD_MAX_LAYERS; ++l) {
    if (ns.tc->sdf_img[l]) cudaFree(ns.tc->sdf_img[l]);
    if (ns.tc->rgb_img[l]) cudaFree(ns.tc->rgb_img[l]);
    if (ns.tc->sdf_imgT[l]) cudaFree(ns.tc->sdf_imgT[l]);
    if (l < 6 && ns.tc->rgb_imgT[l]) cudaFree(ns.tc->rgb_imgT[l]);
  }
  if (ns.tc->bias_s) cudaFree(ns.tc->bias_s);
  delete ns.tc;
  ns.tc = nullptr;
}

static int tc_pack(hold_ctx* ctx, NodeState& ns, const hold_mlp_weights* sdf, const hold_mlp_weights* rgb, cudaStream_t s) {
  if (!ns.tc) ns.tc = new TcMlp();
  TcMlp& t = *ns.tc;
  for (int l = 0; l < 9; ++l) {
    const int K = (l == 0) ? kEmbed : kHidden, kpad = (l == 0) ? 64 : 256;
    const int N = (l == 3) ? kHidden - kEmbed : kHidden;
    const int row_off = (l == 8) ? 1 : 0;
    t.sdf_nst[l] = kpad / 32;
    if (!t.sdf_img[l]) HOLD_CUDA(cudaMalloc((void**)&t.sdf_img[l], (size_t)t.sdf_nst[l] * kTcStageBytes));
    const float scale = (l == 4) ? (float)(1.0 / sqrt(2.0)) : 1.0f;
    k_tc_pack<<<256, 128, 0, s>>>(sdf->weight_v[l], sdf->weight_g[l], sdf->in_dim[l], row_off, N, K, kpad, scale, 0, t.sdf_img[l]);
    HOLD_LAUNCH_CHECK(ctx);
  }
  for (int l = 0; l < 9; ++l) {  // W_l^T for the reverse-mode gradient (layers 7..0); l = 8: the feature rows (training backward)
    const int K_in = (l == 0) ? kEmbed : kHidden, N_out = (l == 3) ? kHidden - kEmbed : kHidden;
    if (!t.sdf_imgT[l]) HOLD_CUDA(cudaMalloc((void**)&t.sdf_imgT[l], (size_t)8 * kTcStageBytes));
    const float scale = (l == 4) ? (float)(1.0 / sqrt(2.0)) : 1.0f;
    k_tc_pack_T<<<256, 256, 0, s>>>(sdf->weight_v[l], sdf->weight_g[l], sdf->in_dim[l], l == 8 ? 1 : 0, N_out, K_in, 0, 0, scale, t.sdf_imgT[l]);
    HOLD_LAUNCH_CHECK(ctx);
  }
  for (int i = 0; i < 5; ++i) {  // colour net transposes for the training backward: lin0 in two parts (its K = 320 > 256), lin1..3
    const int l = (i < 2) ? 0 : i - 1;
    if (!t.rgb_imgT[i]) HOLD_CUDA(cudaMalloc((void**)&t.rgb_imgT[i], (size_t)8 * kTcStageBytes));
    k_tc_pack_T<<<256, 256, 0, s>>>(rgb->weight_v[l], rgb->weight_g[l], rgb->in_dim[l], 0, 256, rgb->in_dim[l], i == 0 ? 1 : (i == 1 ? 2 : 0), 14, 1.0f,
                                    t.rgb_imgT[i]);
    HOLD_LAUNCH_CHECK(ctx);
  }
  if (!t.bias_s) HOLD_CUDA(cudaMalloc((void**)&t.bias_s, 13 * 256 * sizeof(float)));
  for (int l = 0; l < 13; ++l) {   // (the fp32 packing of this call has already written ns.sdf.bias / ns.rgb.bias on this stream)
    k_scale_vec<<<1, 256, 0, s>>>(l < 9 ? ns.sdf.bias[l] : ns.rgb.bias[l - 9], 256, l == 8 ? 1.0f : kTcScaleA, t.bias_s + 256 * l);
    HOLD_LAUNCH_CHECK(ctx);
  }
  for (int l = 0; l < 4; ++l) {
    const int K = (l == 0) ? rgb->in_dim[0] : 256, kpad = (l == 0) ? 320 : 256;
    t.rgb_nst[l] = kpad / 32;
    if (!t.rgb_img[l]) HOLD_CUDA(cudaMalloc((void**)&t.rgb_img[l], (size_t)t.rgb_nst[l] * kTcStageBytes));
    k_tc_pack<<<256, 128, 0, s>>>(rgb->weight_v[l], rgb->weight_g[l], rgb->in_dim[l], 0, 256, K, kpad, 1.0f, l == 0 ? 14 : 0, t.rgb_img[l]);
    HOLD_LAUNCH_CHECK(ctx);
  }
  return HOLD_OK;
}

// SDF net on P canonical points: sdf only (sampler rounds: 8 layers, sdf head), or sdf + d sdf / d x_c + 256-d feature
// (reverse mode: 8 forward layers, feature layer, 8 backward layers over the transposed images).
static int tc_launch_sdf(hold_ctx* ctx, NodeState& ns, int P, const float* xc, const float* embed_w, float* sdf, float* grad,
                         float* feat, const SamplerState* st, cudaStream_t s) {
  const bool rev = (grad != nullptr) || (feat != nullptr);
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P, a.n_layers = rev ? 17 : 8;
  for (int l = 0; l < 9; ++l) {
    a.L[l].wimg = ns.tc->sdf_img[l], a.L[l].bias = ns.tc->bias_s + 256 * l, a.L[l].nst = ns.tc->sdf_nst[l], a.L[l].N = ns.sdf.N[l];
  }
  a.w_last = ns.sdf.w_last, a.b_last = ns.sdf.b_last;
  a.xc = xc, a.embed_w = embed_w, a.sdf = sdf, a.grad = grad, a.feat = feat, a.st = st, a.err = ctx->dev_err;
  a.passes = (st != nullptr && ctx->sampler_passes >= 1 && ctx->sampler_passes <= 3) ? ctx->sampler_passes : 3;
  a.unscale = kTcUnscale * (1.0f + (float)(ctx->tc_acc_comp >= 0 ? ctx->tc_acc_comp : kTcAccComp) * (1.0f / 16777216.0f));
  const int tiles = ceil_div(P, kTcRows), grid = min(tiles, ctx->sm_count);
#if defined(HOLD_TC_EXP) && HOLD_TC_EXP == 7
  if (!rev) { void* rb = nullptr; if (ws_get(ctx, 12, 4096, &rb)) return HOLD_E_CUDA; a.sig = (uint16_t*)rb; }
#endif
  if (rev) {
    HOLD_REQUIRE(grad != nullptr && feat != nullptr, "sdf eval with gradient needs both grad and feat buffers");
    void* sig = nullptr;
    int rc = ws_get(ctx, 12 /* WS_SIG */, (size_t)grid * 8 * kTcRows * 256 * sizeof(uint16_t), &sig);
    if (rc) return rc;
    a.sig = (uint16_t*)sig;
    for (int i = 0; i < 8; ++i) {
      a.L[9 + i].wimg = ns.tc->sdf_imgT[7 - i], a.L[9 + i].bias = nullptr, a.L[9 + i].nst = 8, a.L[9 + i].N = 256;
    }
    k_mlp_tc<MLP_SDF_REV><<<grid, kTcThreadsTotal, TcCfg<MLP_SDF_REV>::kSmemBytes, s>>>(a);
  } else {
    k_mlp_tc<MLP_SDF_ONLY><<<grid, kTcThreadsTotal, TcCfg<MLP_SDF_ONLY>::kSmemBytes, s>>>(a);
  }
  HOLD_LAUNCH_CHECK(ctx);
  return HOLD_OK;
}

// hold_linear: C[P, nvalid] = A[P, kvalid] . M^T (+ bias) against one packed image (selection of the image: api.cu)
static int tc_launch_linear_img(hold_ctx* ctx, const uint8_t* img, const float* bias, int nst, int P, const float* A, int lda, int kvalid,
                                const float* in_scale, float* Cout, int ldc, int nvalid, cudaStream_t s) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P, a.n_layers = 1;
  HOLD_REQUIRE(lda % 4 == 0 && ldc % 4 == 0 && ((uintptr_t)A & 15) == 0 && ((uintptr_t)Cout & 15) == 0, "hold_linear: rows must be 16-byte aligned");
  a.L[0].wimg = img, a.L[0].bias = bias, a.L[0].nst = nst, a.L[0].N = nvalid;
  a.passes = 3;
  a.lin_in = A, a.lda = lda, a.kvalid = kvalid, a.lin_out = Cout, a.ldc = ldc, a.nvalid = nvalid, a.in_scale = in_scale;
  a.err = ctx->dev_err, a.unscale = kTcUnscale * (1.0f + (float)(ctx->tc_acc_comp >= 0 ? ctx->tc_acc_comp : kTcAccComp) * (1.0f / 16777216.0f));
  const int tiles = ceil_div(P, kTcRows);
  k_mlp_tc<MLP_LINEAR><<<min(tiles, ctx->sm_count), kTcThreadsTotal, TcCfg<MLP_LINEAR>::kSmemBytes, s>>>(a);
  HOLD_LAUNCH_CHECK(ctx);
  return HOLD_OK;
}

static int tc_launch_rgb(hold_ctx* ctx, NodeState& ns, int P, int pts_per_frame, const float* xc, const float* normal,
                         const float* pe, const float* feat, const float* time_code, float* rgb, cudaStream_t s) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P, a.n_layers = 4;
  for (int l = 0; l < 4; ++l) {
    a.L[l].wimg = ns.tc->rgb_img[l], a.L[l].bias = ns.tc->bias_s + 256 * (9 + l), a.L[l].nst = ns.tc->rgb_nst[l], a.L[l].N = 256;
  }
  a.w_last = ns.rgb.w_last, a.b_last = ns.rgb.b_last;
  a.xc = xc, a.normal = normal, a.pose_embed = pe, a.feat = const_cast<float*>(feat), a.time_code = time_code;
  a.pts_per_frame = pts_per_frame, a.k0 = ns.rgb.K[0], a.rgb = rgb, a.err = ctx->dev_err, a.unscale = kTcUnscale, a.passes = 3;
  int tiles = ceil_div(P, kTcRows);
  k_mlp_tc<MLP_COLOR><<<min(tiles, ctx->sm_count), kTcThreadsTotal, TcCfg<MLP_COLOR>::kSmemBytes, s>>>(a);
  HOLD_LAUNCH_CHECK(ctx);
  return HOLD_OK;
}

}  // namespace hold
