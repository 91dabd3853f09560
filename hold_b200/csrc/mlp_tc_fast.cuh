// Sampler-round SDF kernel, FAST variant (HOLD_TC_FAST=1; round-2 A/B against k_mlp_tc<MLP_SDF_ONLY>): same tile / role /
// barrier structure and the same fp16 hi/lo split arithmetic, but the epilogue is rebuilt around its measured cost
// (profiles/r01_ncu_k_mlp_tc0.md: ~135 warp instructions per 8-column round, 47 of them per-round overhead):
//   * base-2-domain softplus of the LEAN scheme (mlp_tc.cuh: t straight from the accumulator by one FMA, S(t) is the next
//     operand, ln2/100 folded into the LEAN weight images) — 6 instead of 8 instructions per element;
//   * each epilogue warp takes 16 columns per round, so a layer is 4 rounds / 4 hand-offs of 64 columns (= one SW128 A chunk =
//     two weight stages) instead of 8 x 32: half the bias loads, TMEM waits, fences, barrier arrivals and loop overhead per
//     element; the MMA tail after the last hand-off grows from 1/8 to 1/4 of a layer (the price);
//   * the round loop is fully unrolled (compile-time column offsets), the head dot product only exists in layer 7's copy.
// No hardware run yet (written after the round-1 GPU budget was spent).
#pragma once
#include "mlp_tc.cuh"

namespace hold {

constexpr int kFastHandoffs = 4;
constexpr int kFastSmemA = 2 * 4 * kTcAChunkBytes;      // hi + lo, 4 chunks of [128 x 64]
constexpr int kFastStages = 3;
constexpr int kFastSmemBytes = kFastSmemA + kFastStages * kTcStageBytes + 256 + 1024;

__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// one layer of the epilogue; HEAD = layer 7 (sdf head, no next operand)
template <bool HEAD>
__device__ __forceinline__ bool fast_layer(const TcArgs& a, int l, uint32_t t_lane, uint32_t bDFull, uint32_t bAReady, uint8_t* gA_hi,
                                           uint8_t* gA_lo, int row, int sub, int lane, float px, float py, float pz, uint32_t& d_par,
                                           volatile int* abort_flag, float& head0) {
  const float* bias = a.L[l].bias;   // LEAN: bias * 100 log2(e)
  if (!mbar_wait(bDFull + 8 * (l & 1), (d_par >> (l & 1)) & 1, a.err, 4, abort_flag)) return false;
  d_par ^= (1u << (l & 1));
  tc_fence_after();
  const uint32_t t_col = t_lane + (uint32_t)((l & 1) * 256 + sub * 16);
  uint32_t raw[16];
  tc_ld16(t_col, raw);
#pragma unroll
  for (int h = 0; h < kFastHandoffs; ++h) {
    const int n0 = h * 64 + sub * 16;
    float4 b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = __ldg(reinterpret_cast<const float4*>(bias + n0) + u);   // in flight during the TMEM wait
    tc_wait_ld();
    float out[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = __uint_as_float(raw[i]);
    if (h + 1 < kFastHandoffs) tc_ld16(t_col + (uint32_t)((h + 1) * 64), raw);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float e;
      out[4 * u + 0] = softplus_t(fmaf(out[4 * u + 0], kLeanAccToT, b[u].x), e);
      out[4 * u + 1] = softplus_t(fmaf(out[4 * u + 1], kLeanAccToT, b[u].y), e);
      out[4 * u + 2] = softplus_t(fmaf(out[4 * u + 2], kLeanAccToT, b[u].z), e);
      out[4 * u + 3] = softplus_t(fmaf(out[4 * u + 3], kLeanAccToT, b[u].w), e);
    }
    if (l == 3 && n0 + 16 > kHidden - kEmbed) {   // skip connection: embedding columns of layer 3's output (A scale 2^6)
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + i >= kHidden - kEmbed) out[i] = kTcScaleA * embed_val(n0 + i - (kHidden - kEmbed), 0, px, py, pz, a.embed_w);
    }
    if (HEAD) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.w_last + n0) + u);   // LEAN: sdf row * ln2/100
        head0 += out[4 * u] * w.x + out[4 * u + 1] * w.y + out[4 * u + 2] * w.z + out[4 * u + 3] * w.w;
      }
    } else {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint4 hi, lo;
        split8(out + 8 * u, hi, lo);
        const uint32_t off = (uint32_t)(h * kTcAChunkBytes) + a_unit_off(row, sub * 2 + u);
        *reinterpret_cast<uint4*>(gA_hi + off) = hi;
        *reinterpret_cast<uint4*>(gA_lo + off) = lo;
      }
      handoff_arrive(bAReady + 8 * h, lane);
    }
  }
  return true;
}

__global__ void __launch_bounds__(kTcThreadsTotal, 1) k_mlp_tc_fast(TcArgs a) {
  if (a.st != nullptr && a.st->done) return;
  constexpr int NS = kFastStages, NHO = kFastHandoffs;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA_hi = base, sA_lo = base + 4 * kTcAChunkBytes, sW = base + kFastSmemA;
  const uint32_t sBar = sW + NS * kTcStageBytes;
  const uint32_t bWFull = sBar, bWEmpty = sBar + 8 * NS, bAReady = sBar + 16 * NS, bDFull = bAReady + 8 * NHO;
  const uint32_t sTmemPtr = bDFull + 16, sAbort = bDFull + 20;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gen_base + (sAbort - base));
  const int n_tiles = ceil_div(a.P, kTcRows);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bWFull + 8 * i, 1); mbar_init(bWEmpty + 8 * i, 1); }
    *abort_flag = 0;
    for (int i = 0; i < NHO; ++i) mbar_init(bAReady + 8 * i, kTcEpiWarps);
    mbar_init(bDFull, 1);
    mbar_init(bDFull + 8, 1);
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
    // ============================================================ weight producer (as k_mlp_tc)
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < 8; ++l) {
        const uint8_t* src = a.L[l].wimg + (a.wcopies > 1 ? (size_t)(blockIdx.x % a.wcopies) * a.L[l].nst * kTcStageBytes : 0);
        for (int s = 0; s < a.L[l].nst; ++s) {
          if (!__all_sync(0xffffffffu, mbar_wait(bWEmpty + 8 * stage, phase ^ 1, a.err, 1, abort_flag))) goto fast_done;
          if (elect_one()) {
            mbar_expect_tx(bWFull + 8 * stage, kTcStageBytes);
            bulk_g2s(sW + stage * kTcStageBytes, src + (size_t)s * kTcStageBytes, kTcStageBytes, bWFull + 8 * stage);
          }
          __syncwarp();
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================================================ MMA issuer: one hand-off = one 64-wide A chunk = two weight stages
    uint32_t stage = 0, phase = 0, a_par = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < 8; ++l) {
        const uint32_t d_tmem = tmem_u + (uint32_t)((l & 1) * 256);
        const int nst = a.L[l].nst;
        for (int s = 0; s < nst; ++s) {
          const int c = s >> 1;
          if ((s & 1) == 0) {
            if (!__all_sync(0xffffffffu, mbar_wait(bAReady + 8 * c, (a_par >> c) & 1, a.err, 2, abort_flag))) goto fast_done;
            a_par ^= (1u << c);
          }
          if (!__all_sync(0xffffffffu, mbar_wait(bWFull + 8 * stage, phase, a.err, 3, abort_flag))) goto fast_done;
          tc_fence_after();
          const uint32_t wb = sW + stage * kTcStageBytes;
          const bool el = elect_one();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t koff = (uint32_t)(((s & 1) * 2 + j) * 32);
            const uint64_t ahi = umma_desc(sA_hi + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
            const uint64_t alo = umma_desc(sA_lo + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
            const uint64_t whi = umma_desc(wb + j * 32, 512, kLayoutSW64);
            const uint64_t wlo = umma_desc(wb + 16384 + j * 32, 512, kLayoutSW64);
            if (el) {
              tc_mma(d_tmem, ahi, whi, kIdescF16, (s | j) != 0);
              tc_mma(d_tmem, alo, whi, kIdescF16, 1);
              tc_mma(d_tmem, ahi, wlo, kIdescF16, 1);
            }
          }
          if (el) tc_commit(bWEmpty + 8 * stage);
          __syncwarp();
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(bDFull + 8 * (l & 1));
        __syncwarp();
      }
    }
  } else {
    // ============================================================ epilogue: 16 warps, 16 columns per warp and round
    const int q = warp & 3, sub = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* gA_hi = gen_base;
    uint8_t* gA_lo = gen_base + 4 * kTcAChunkBytes;
    float* scratch = reinterpret_cast<float*>(gen_base);
    uint32_t d_par = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int p = tile * kTcRows + row;
      const bool valid = p < a.P;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (valid) { px = a.xc[3 * (size_t)p], py = a.xc[3 * (size_t)p + 1], pz = a.xc[3 * (size_t)p + 2]; }
      {  // prologue: the 64-wide layer-0 operand (39 embedding columns, zero padded), this warp's 16 columns
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = kTcScaleA * embed_val(sub * 16 + i, 0, px, py, pz, a.embed_w);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint4 hi, lo;
          split8(x + 8 * u, hi, lo);
          *reinterpret_cast<uint4*>(gA_hi + a_unit_off(row, sub * 2 + u)) = hi;
          *reinterpret_cast<uint4*>(gA_lo + a_unit_off(row, sub * 2 + u)) = lo;
        }
        handoff_arrive(bAReady, lane);
      }
      float head0 = 0.f;
      bool ok = true;
      for (int l = 0; l < 7 && ok; ++l)
        ok = fast_layer<false>(a, l, t_lane, bDFull, bAReady, gA_hi, gA_lo, row, sub, lane, px, py, pz, d_par, abort_flag, head0);
      if (ok) fast_layer<true>(a, 7, t_lane, bDFull, bAReady, gA_hi, gA_lo, row, sub, lane, px, py, pz, d_par, abort_flag, head0);
      // head: fixed-order reduction over the quarter's 4 warps (all MMAs of the tile are complete: the A region is free)
      tc_fence_before();
      scratch[sub * kTcRows + row] = head0;
      epi_bar();
      if (sub == 0 && valid) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < kTcW; ++w) acc += scratch[w * kTcRows + row];
        a.sdf[p] = acc + a.b_last[0];
      }
      epi_bar();
    }
  }
fast_done:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------ reverse mode (shading)
// The 17-step chain of k_mlp_tc<MLP_SDF_REV> (8 forward layers stashing softplus', feature layer, 8 backward layers over the
// transposed images) with the same rebuilt epilogue: 16 columns per warp and round, LEAN forward arithmetic, and the stash holds
// t = 100 z log2(e) — softplus'(z) = sigmoid_t(t) is evaluated where it is consumed (backward rounds, whose MUFU pipe is idle)
// instead of costing the forward rounds a third MUFU per element.
// KIND 0: forward layer l; 1: feature layer + seed of the backward chain; 2: backward through layer l; 3: layer 0 (embedding).
template <int KIND>
__device__ __forceinline__ bool fast_rev_step(const TcArgs& a, int st, int l, uint32_t t_lane, uint32_t bDFull, uint32_t bAReady,
                                              uint8_t* gA_hi, uint8_t* gA_lo, float* sig, int row, int sub, int lane, float px, float py,
                                              float pz, bool valid, int p, uint32_t& d_par, volatile int* abort_flag, float& head0,
                                              float& gx, float& gy, float& gz) {
  constexpr int kSigL = kTcRows * 256;
  const float* side = (KIND <= 1) ? a.L[st].bias : ((KIND == 2) ? sig + (size_t)(l - 1) * kSigL : nullptr);
  if (!mbar_wait(bDFull + 8 * (st & 1), (d_par >> (st & 1)) & 1, a.err, 4, abort_flag)) return false;
  d_par ^= (1u << (st & 1));
  tc_fence_after();
  const uint32_t t_col = t_lane + (uint32_t)((st & 1) * 256 + sub * 16);
  uint32_t raw[16];
  tc_ld16(t_col, raw);
#pragma unroll
  for (int h = 0; h < kFastHandoffs; ++h) {
    const int n0 = h * 64 + sub * 16;
    float4 sv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sv[u] = (KIND == 3) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(side + n0 + 4 * u);
    tc_wait_ld();
    float out[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = __uint_as_float(raw[i]);
    if (h + 1 < kFastHandoffs) tc_ld16(t_col + (uint32_t)((h + 1) * 64), raw);
    if (KIND == 0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float e;
        const float t0 = fmaf(out[4 * u + 0], kLeanAccToT, sv[u].x), t1 = fmaf(out[4 * u + 1], kLeanAccToT, sv[u].y);
        const float t2 = fmaf(out[4 * u + 2], kLeanAccToT, sv[u].z), t3 = fmaf(out[4 * u + 3], kLeanAccToT, sv[u].w);
        *reinterpret_cast<float4*>(sig + (size_t)l * kSigL + n0 + 4 * u) = make_float4(t0, t1, t2, t3);   // the stash holds t
        out[4 * u + 0] = softplus_t(t0, e), out[4 * u + 1] = softplus_t(t1, e);
        out[4 * u + 2] = softplus_t(t2, e), out[4 * u + 3] = softplus_t(t3, e);
      }
      if (l == 3 && n0 + 16 > kHidden - kEmbed) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + i >= kHidden - kEmbed) out[i] = kTcScaleA * embed_val(n0 + i - (kHidden - kEmbed), 0, px, py, pz, a.embed_w);
      }
      if (l == 7) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(a.w_last + n0) + u);
          head0 += out[4 * u] * w.x + out[4 * u + 1] * w.y + out[4 * u + 2] * w.z + out[4 * u + 3] * w.w;
        }
      }
    } else if (KIND == 1) {
      constexpr float ks = kTcScaleA / kLeanAct;   // a.w_last holds w_sdf * ln2/100
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (valid)
          *reinterpret_cast<float4*>(a.feat + (size_t)p * kFeat + n0 + 4 * u) =
              make_float4(fmaf(out[4 * u], kLeanAccToZ, sv[u].x), fmaf(out[4 * u + 1], kLeanAccToZ, sv[u].y),
                          fmaf(out[4 * u + 2], kLeanAccToZ, sv[u].z), fmaf(out[4 * u + 3], kLeanAccToZ, sv[u].w));
        const float4 t7 = *reinterpret_cast<const float4*>(sig + (size_t)7 * kSigL + n0 + 4 * u);
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.w_last + n0) + u);
        out[4 * u + 0] = ks * w.x * sigmoid_t(t7.x), out[4 * u + 1] = ks * w.y * sigmoid_t(t7.y);
        out[4 * u + 2] = ks * w.z * sigmoid_t(t7.z), out[4 * u + 3] = ks * w.w * sigmoid_t(t7.w);
      }
    } else if (KIND == 2) {
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = out[i] * kTcUnscale;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        out[4 * u + 0] = kTcScaleA * acc[4 * u + 0] * sigmoid_t(sv[u].x), out[4 * u + 1] = kTcScaleA * acc[4 * u + 1] * sigmoid_t(sv[u].y);
        out[4 * u + 2] = kTcScaleA * acc[4 * u + 2] * sigmoid_t(sv[u].z), out[4 * u + 3] = kTcScaleA * acc[4 * u + 3] * sigmoid_t(sv[u].w);
      }
      if (l == 4 && n0 + 16 > kHidden - kEmbed) {   // skip input of layer 4: columns 217.. are d sdf / d embed
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          if (n0 + i >= kHidden - kEmbed) {
            const int em = n0 + i - (kHidden - kEmbed), d = em % 3;
            const float je = acc[i] * embed_val(em, d + 1, px, py, pz, a.embed_w);
            gx += (d == 0) ? je : 0.f;
            gy += (d == 1) ? je : 0.f;
            gz += (d == 2) ? je : 0.f;
            out[i] = 0.f;
          }
        }
      }
    } else {   // KIND 3: d sdf / d embed through layer 0's input (columns 0..38)
      if (n0 < 48) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int em = n0 + i;
          if (em < kEmbed) {
            const int d = em % 3;
            const float je = out[i] * kTcUnscale * embed_val(em, d + 1, px, py, pz, a.embed_w);
            gx += (d == 0) ? je : 0.f;
            gy += (d == 1) ? je : 0.f;
            gz += (d == 2) ? je : 0.f;
          }
        }
      }
    }
    if (KIND != 3) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint4 hi, lo;
        split8(out + 8 * u, hi, lo);
        const uint32_t off = (uint32_t)(h * kTcAChunkBytes) + a_unit_off(row, sub * 2 + u);
        *reinterpret_cast<uint4*>(gA_hi + off) = hi;
        *reinterpret_cast<uint4*>(gA_lo + off) = lo;
      }
      handoff_arrive(bAReady + 8 * h, lane);
    }
  }
  return true;
}

__global__ void __launch_bounds__(kTcThreadsTotal, 1) k_mlp_tc_fast_rev(TcArgs a) {
  constexpr int NS = kFastStages, NHO = kFastHandoffs;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA_hi = base, sA_lo = base + 4 * kTcAChunkBytes, sW = base + kFastSmemA;
  const uint32_t sBar = sW + NS * kTcStageBytes;
  const uint32_t bWFull = sBar, bWEmpty = sBar + 8 * NS, bAReady = sBar + 16 * NS, bDFull = bAReady + 8 * NHO;
  const uint32_t sTmemPtr = bDFull + 16, sAbort = bDFull + 20;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gen_base + (sAbort - base));
  const int n_tiles = ceil_div(a.P, kTcRows);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bWFull + 8 * i, 1); mbar_init(bWEmpty + 8 * i, 1); }
    *abort_flag = 0;
    for (int i = 0; i < NHO; ++i) mbar_init(bAReady + 8 * i, kTcEpiWarps);
    mbar_init(bDFull, 1);
    mbar_init(bDFull + 8, 1);
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
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int st = 0; st < 17; ++st) {
        const uint8_t* src = a.L[st].wimg + (a.wcopies > 1 ? (size_t)(blockIdx.x % a.wcopies) * a.L[st].nst * kTcStageBytes : 0);
        for (int s = 0; s < a.L[st].nst; ++s) {
          if (!__all_sync(0xffffffffu, mbar_wait(bWEmpty + 8 * stage, phase ^ 1, a.err, 1, abort_flag))) goto fastrev_done;
          if (elect_one()) {
            mbar_expect_tx(bWFull + 8 * stage, kTcStageBytes);
            bulk_g2s(sW + stage * kTcStageBytes, src + (size_t)s * kTcStageBytes, kTcStageBytes, bWFull + 8 * stage);
          }
          __syncwarp();
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0, a_par = 0;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int st = 0; st < 17; ++st) {
        const uint32_t d_tmem = tmem_u + (uint32_t)((st & 1) * 256);
        const int nst = a.L[st].nst;
        for (int s = 0; s < nst; ++s) {
          const int c = s >> 1;
          if ((s & 1) == 0) {
            if (!__all_sync(0xffffffffu, mbar_wait(bAReady + 8 * c, (a_par >> c) & 1, a.err, 2, abort_flag))) goto fastrev_done;
            a_par ^= (1u << c);
          }
          if (!__all_sync(0xffffffffu, mbar_wait(bWFull + 8 * stage, phase, a.err, 3, abort_flag))) goto fastrev_done;
          tc_fence_after();
          const uint32_t wb = sW + stage * kTcStageBytes;
          const bool el = elect_one();
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint32_t koff = (uint32_t)(((s & 1) * 2 + j) * 32);
            const uint64_t ahi = umma_desc(sA_hi + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
            const uint64_t alo = umma_desc(sA_lo + c * kTcAChunkBytes + koff, 1024, kLayoutSW128);
            const uint64_t whi = umma_desc(wb + j * 32, 512, kLayoutSW64);
            const uint64_t wlo = umma_desc(wb + 16384 + j * 32, 512, kLayoutSW64);
            if (el) {
              tc_mma(d_tmem, ahi, whi, kIdescF16, (s | j) != 0);
              tc_mma(d_tmem, alo, whi, kIdescF16, 1);
              tc_mma(d_tmem, ahi, wlo, kIdescF16, 1);
            }
          }
          if (el) tc_commit(bWEmpty + 8 * stage);
          __syncwarp();
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(bDFull + 8 * (st & 1));
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3, sub = (warp - 2) >> 2, row = q * 32 + lane;
    const uint32_t t_lane = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* gA_hi = gen_base;
    uint8_t* gA_lo = gen_base + 4 * kTcAChunkBytes;
    float* scratch = reinterpret_cast<float*>(gen_base);
    float* sig = a.sig + (size_t)blockIdx.x * (8 * kTcRows * 256) + (size_t)row * 256;
    uint32_t d_par = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int p = tile * kTcRows + row;
      const bool valid = p < a.P;
      float px = 0.f, py = 0.f, pz = 0.f;
      if (valid) { px = a.xc[3 * (size_t)p], py = a.xc[3 * (size_t)p + 1], pz = a.xc[3 * (size_t)p + 2]; }
      {
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = kTcScaleA * embed_val(sub * 16 + i, 0, px, py, pz, a.embed_w);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint4 hi, lo;
          split8(x + 8 * u, hi, lo);
          *reinterpret_cast<uint4*>(gA_hi + a_unit_off(row, sub * 2 + u)) = hi;
          *reinterpret_cast<uint4*>(gA_lo + a_unit_off(row, sub * 2 + u)) = lo;
        }
        handoff_arrive(bAReady, lane);
      }
      float head0 = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
      bool ok = true;
#define FAST_REV(KIND, ST, L) fast_rev_step<KIND>(a, ST, L, t_lane, bDFull, bAReady, gA_hi, gA_lo, sig, row, sub, lane, px, py, pz, valid, p, d_par, abort_flag, head0, gx, gy, gz)
      for (int st = 0; st < 8 && ok; ++st) ok = FAST_REV(0, st, st);
      if (ok) ok = FAST_REV(1, 8, 8);
      for (int st = 9; st < 16 && ok; ++st) ok = FAST_REV(2, st, 16 - st);
      if (ok) ok = FAST_REV(3, 16, 0);
#undef FAST_REV
      tc_fence_before();
      scratch[(sub * 4 + 0) * kTcRows + row] = head0;
      scratch[(sub * 4 + 1) * kTcRows + row] = gx;
      scratch[(sub * 4 + 2) * kTcRows + row] = gy;
      scratch[(sub * 4 + 3) * kTcRows + row] = gz;
      epi_bar();
      if (sub == 0 && valid) {
        float hs[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float acc = 0.f;
#pragma unroll
          for (int w = 0; w < kTcW; ++w) acc += scratch[(w * 4 + k) * kTcRows + row];
          hs[k] = acc;
        }
        a.sdf[p] = hs[0] + a.b_last[0];
        a.grad[3 * (size_t)p] = hs[1], a.grad[3 * (size_t)p + 1] = hs[2], a.grad[3 * (size_t)p + 2] = hs[3];
      }
      epi_bar();
    }
  }
fastrev_done:
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static int tc_fast_init() {
  cudaError_t e = cudaFuncSetAttribute(k_mlp_tc_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_tc_fast_rev, cudaFuncAttributeMaxDynamicSharedMemorySize, kFastSmemBytes);
  if (e != cudaSuccess) { set_error("tcgen05 fast kernel attribute: %s", cudaGetErrorString(e)); return HOLD_E_CUDA; }
  return HOLD_OK;
}

// HOLD_TC_WCOPIES=N: N replicas of every weight image the FAST kernels stream, CTA b reads replica b % N — a test of (and remedy
// for) L2 hot-spotting when 148 CTAs in near lock-step fetch the same 32 KB stage.  Built lazily, rebuilt when weights change.
static int tc_fast_replicas(hold_ctx* ctx, NodeState& ns, cudaStream_t s) {
  const char* e = getenv("HOLD_TC_WCOPIES");
  const int n = e ? atoi(e) : 0;
  if (n <= 1) return 0;
  TcMlp& t = *ns.tc;
  for (int l = 0; l < 9; ++l) {
    const size_t bytes = (size_t)t.sdf_nst[l] * kTcStageBytes;
    if (t.rep_copies != n && t.sdf_imgL_rep[l]) { cudaFree(t.sdf_imgL_rep[l]); t.sdf_imgL_rep[l] = nullptr; }
    if (!t.sdf_imgL_rep[l] && cudaMalloc((void**)&t.sdf_imgL_rep[l], bytes * n) != cudaSuccess) return -1;
    for (int c = 0; c < n; ++c) cudaMemcpyAsync(t.sdf_imgL_rep[l] + bytes * c, t.sdf_imgL[l], bytes, cudaMemcpyDeviceToDevice, s);
    if (t.rep_copies != n && t.sdf_img_rep[l]) { cudaFree(t.sdf_img_rep[l]); t.sdf_img_rep[l] = nullptr; }
    if (!t.sdf_img_rep[l] && cudaMalloc((void**)&t.sdf_img_rep[l], bytes * n) != cudaSuccess) return -1;
    for (int c = 0; c < n; ++c) cudaMemcpyAsync(t.sdf_img_rep[l] + bytes * c, t.sdf_img[l], bytes, cudaMemcpyDeviceToDevice, s);
  }
  for (int l = 0; l < 8; ++l) {
    const size_t bytes = (size_t)8 * kTcStageBytes;
    if (t.rep_copies != n && t.sdf_imgT_rep[l]) { cudaFree(t.sdf_imgT_rep[l]); t.sdf_imgT_rep[l] = nullptr; }
    if (!t.sdf_imgT_rep[l] && cudaMalloc((void**)&t.sdf_imgT_rep[l], bytes * n) != cudaSuccess) return -1;
    for (int c = 0; c < n; ++c) cudaMemcpyAsync(t.sdf_imgT_rep[l] + bytes * c, t.sdf_imgT[l], bytes, cudaMemcpyDeviceToDevice, s);
  }
  t.rep_copies = n;
  (void)ctx;
  return n;
}

static inline bool tc_fast_enabled() {
  const char* e = getenv("HOLD_TC_FAST");
  return e != nullptr && atoi(e) != 0;
}

// sampler-round launch (sdf head only) with the LEAN images / pre-scaled biases / head row
static int tc_fast_launch_sdf(hold_ctx* ctx, NodeState& ns, int P, const float* xc, const float* embed_w, float* sdf,
                              const SamplerState* st, cudaStream_t s) {
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P, a.n_layers = 8;
  for (int l = 0; l < 8; ++l) {
    a.L[l].wimg = ns.tc->sdf_imgL[l], a.L[l].bias = ns.tc->sdf_bias_t[l], a.L[l].nst = ns.tc->sdf_nst[l], a.L[l].N = ns.sdf.N[l];
  }
  a.w_last = ns.tc->w_last_t, a.b_last = ns.sdf.b_last;
  a.xc = xc, a.embed_w = embed_w, a.sdf = sdf, a.st = st, a.err = ctx->dev_err;
  const int nrep = tc_fast_replicas(ctx, ns, s);   // refreshed per launch (a few MB of D2D copies; experiment only)
  HOLD_REQUIRE(nrep >= 0, "out of memory for weight-image replicas");
  if (nrep > 1) {
    a.wcopies = nrep;
    for (int l = 0; l < 8; ++l) a.L[l].wimg = ns.tc->sdf_imgL_rep[l];
  }
  const int tiles = ceil_div(P, kTcRows);
  k_mlp_tc_fast<<<min(tiles, ctx->sm_count), kTcThreadsTotal, kFastSmemBytes, s>>>(a);
  HOLD_LAUNCH_CHECK(ctx);
  return HOLD_OK;
}

// shading launch: sdf + gradient + feature (reverse mode), LEAN forward images and the transposed images for the backward half
static int tc_fast_launch_rev(hold_ctx* ctx, NodeState& ns, int P, const float* xc, const float* embed_w, float* sdf, float* grad,
                              float* feat, cudaStream_t s) {
  HOLD_REQUIRE(grad != nullptr && feat != nullptr, "sdf eval with gradient needs both grad and feat buffers");
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P, a.n_layers = 17;
  for (int l = 0; l < 9; ++l) {
    a.L[l].wimg = ns.tc->sdf_imgL[l], a.L[l].bias = (l < 8) ? ns.tc->sdf_bias_t[l] : ns.sdf.bias[8], a.L[l].nst = ns.tc->sdf_nst[l];
    a.L[l].N = ns.sdf.N[l];
  }
  for (int i = 0; i < 8; ++i) {
    a.L[9 + i].wimg = ns.tc->sdf_imgT[7 - i], a.L[9 + i].bias = nullptr, a.L[9 + i].nst = 8, a.L[9 + i].N = 256;
  }
  a.w_last = ns.tc->w_last_t, a.b_last = ns.sdf.b_last;
  a.xc = xc, a.embed_w = embed_w, a.sdf = sdf, a.grad = grad, a.feat = feat, a.err = ctx->dev_err;
  const int tiles = ceil_div(P, kTcRows), grid = min(tiles, ctx->sm_count);
  void* sig = nullptr;
  int rc = ws_get(ctx, 12 /* WS_SIG */, (size_t)grid * 8 * kTcRows * 256 * sizeof(float), &sig);
  if (rc) return rc;
  a.sig = (float*)sig;
  const int nrep = tc_fast_replicas(ctx, ns, s);
  HOLD_REQUIRE(nrep >= 0, "out of memory for weight-image replicas");
  if (nrep > 1) {
    a.wcopies = nrep;
    for (int l = 0; l < 9; ++l) a.L[l].wimg = ns.tc->sdf_imgL_rep[l];
    for (int i = 0; i < 8; ++i) a.L[9 + i].wimg = ns.tc->sdf_imgT_rep[7 - i];
  }
  k_mlp_tc_fast_rev<<<grid, kTcThreadsTotal, kFastSmemBytes, s>>>(a);
  HOLD_LAUNCH_CHECK(ctx);
  return HOLD_OK;
}

}  // namespace hold
