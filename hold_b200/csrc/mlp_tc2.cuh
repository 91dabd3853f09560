// tcgen05 SDF-net chains on CTA PAIRS with two tiles in flight (HOLD_MLP_TC, sampler rounds + shading gradient).
//
// Why: with one 128-row tile per SM (mlp_tc.cuh) the layer chain is serial — layer l+1's accumulator cannot complete
// before layer l's epilogue has, so the tensor pipe idles through every epilogue tail (profiles/r01_ncu_k_mlp_tc0.md:
// 15.3 k clk per layer against 6.1 k clk of MMA work).  A second independent tile per SM needs its A operand (fp16
// hi/lo, 128 KB per 128 rows) in shared memory too, which does not fit.  A CTA pair driving `cta_group::2` MMAs of
// shape M=128 does: each CTA holds only 64 rows of a tile (A = 64 KB) and half of every weight stage (16 KB), so TWO
// 128-point tiles (X, Y) fit per pair, and the 2x2 TMEM layout of the pair's accumulator (64 rows x 256 columns in
// 128 lanes x 128 columns: lanes 64..127 hold columns 128..255) leaves room for 2 tiles x 2 accumulators.
//
//   epilogue warps : ... epi X(l)   epi Y(l)   epi X(l+1)   epi Y(l+1) ...
//   tensor pipe    : ... MMA Y(l)   MMA X(l+1) MMA Y(l+1)   MMA X(l+2) ...     (X(l+1) starts on X(l)'s first hand-off)
//
// Roles per CTA (576 threads): warp 0 = bulk-copy producer of this CTA's half of each weight stage; warp 1 = MMA
// issuer (leader CTA) / "my half has landed" forwarder (peer CTA); warps 2..17 = epilogue, 4 per TMEM lane quarter.
// Lane quarters 0,1 see rows 0..63 x columns 0..127, quarters 2,3 rows 0..63 x columns 128..255; warp `sub` of a
// quarter owns columns 32 j + 8 sub .. +7 of its half in round j = 0..3, so round j completes the next layer's
// k-chunks j and 4 + j (32 wide = one weight stage); the MMA issuer consumes chunks in the order 0,4,1,5,2,6,3,7.
// Arithmetic is identical to mlp_tc.cuh (fp16 hi/lo split operands, 3 passes, fp32 accumulate).
#pragma once
#include "mlp_tc.cuh"
#include "mlp_tc_fast.cuh"

namespace hold {

constexpr int kT2Rows = 64;                 // rows of a tile held by one CTA of the pair
constexpr int kT2TilePts = 128;             // points per tile (pair)
constexpr int kT2Stages = 5;
constexpr int kT2HalfStage = 16384;         // per CTA and stage: hi 8 KB + lo 8 KB of weight rows n in [128 r, 128 r + 128)
constexpr int kT2APart = 32768;             // hi or lo part of one tile's A operand in one CTA: 4 chunks [64 rows x 64 k]
constexpr int kT2ATile = 2 * kT2APart;
constexpr int kT2SmemA = 2 * kT2ATile;      // tiles X, Y
constexpr int kT2SmemW = kT2Stages * kT2HalfStage;
constexpr int kT2SmemBias = 9 * 256 * 4;
constexpr int kT2SmemBytes = kT2SmemA + kT2SmemW + kT2SmemBias + 512 + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {  // shared::cta address -> shared::cluster address in CTA `rank`
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// the form CUTLASS' ClusterBarrier::arrive(cta_id) uses: default semantics (.release at CTA scope) on the remote barrier; the
// data it publishes (this CTA's shared memory, read by the async proxy) has been ordered by fence.proxy.async before
__device__ __forceinline__ void mbar_arrive_cluster_light(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Bounded wait with cluster-scope acquire (arrivals come from the peer CTA and from multicast commits).  On a
// timeout the waiter records its tag and raises BOTH CTAs' abort flags; every other wait polls its flag.
__device__ __forceinline__ bool mbar_wait2(uint32_t bar, uint32_t parity, int* err, int tag, volatile int* abort_flag) {
  uint32_t done = 0;
  if (*abort_flag) return false;  // after an abort every wait of the pair falls through at once
  for (unsigned spin = 0; spin < (1u << 20); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return true;
    if ((spin & 63) == 63 && *abort_flag) return false;
  }
  if (err != nullptr) atomicOr(err, 0x100 | (tag << 12));
  *abort_flag = 1;
  const uint32_t peer = mapa_rank(smem_u32((const void*)abort_flag), cluster_ctarank() ^ 1u);
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(peer), "r"(1u) : "memory");
  return false;
}
// the same, adding the cycles spent to `acc` when profiling is on (HOLD_TC_PROF=1, cluster 0 only)
__device__ __forceinline__ bool mbar_wait2t(uint32_t bar, uint32_t parity, int* err, int tag, volatile int* abort_flag, bool prof,
                                            long long& acc) {
  if (!prof) return mbar_wait2(bar, parity, err, tag, abort_flag);
  const long long t0 = clock64();
  const bool ok = mbar_wait2(bar, parity, err, tag, abort_flag);
  acc += clock64() - t0;
  return ok;
}
// completion of all previously issued MMAs -> one arrival on the barrier at the same smem offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}

// constants of the pre-scaled softplus: z64 = 64 z; out = 64 softplus_100(z)
constexpr float kT2AccToZ64 = kTcUnscale * kTcScaleA;                              // accumulator -> 64 z
constexpr float kT2Z64ToT = 100.0f * 1.4426950408889634f / kTcScaleA;             // 64 z -> 100 z log2(e)
constexpr float kT2LgToOut = 0.6931471805599453f * 0.01f * kTcScaleA;             // lg2(1 + u) -> 64 log1p(u) / 100

struct T2Epi {            // per-thread constants of an epilogue warp
  uint32_t t_lane;        // TMEM address of this warp's lane quarter, column 0
  uint8_t* gA;            // generic pointer to the A region (tile X hi | X lo | Y hi | Y lo)
  const float* bias;      // smem: [9][256], layers 0..7 pre-multiplied by 64
  uint32_t a_ready;       // shared::cluster address of the leader's a_ready[2][4]
  uint32_t d_full;        // local d_full[2][2]
  volatile int* abort_flag;
  int row, qh, sub, lane; // row 0..63 inside the CTA's half tile; qh = column half; sub = warp of the quarter
  bool prof, coarse, light;
};

// `arrive`: fine-grained mode arrives after every round on a_ready[t][round]; coarse mode (HOLD_TC_DBG & 32) only after a
// warp's last round of the tile-step, on a_ready[t][0] — one fence + one remote arrival per warp and tile-step.
__device__ __forceinline__ void t2_store_a(const T2Epi& e, int t, int n0, int round, bool arrive, const float (&out)[8]) {
  uint4 hi, lo;
  split8(out, hi, lo);
  const int c64 = n0 >> 6, ju = (n0 & 63) >> 3;
  const uint32_t off = (uint32_t)(t * kT2ATile + c64 * 8192 + (e.row >> 3) * 1024 + (e.row & 7) * 128 + ((ju ^ (e.row & 7)) << 4));
  *reinterpret_cast<uint4*>(e.gA + off) = hi;
  *reinterpret_cast<uint4*>(e.gA + off + kT2APart) = lo;
  if (!arrive) return;
  // one arrival per warp on the leader's hand-off barrier of this round (k-chunks j and 4 + j) of tile t
  fence_proxy_async();
  tc_fence_before();
  __syncwarp();
  if (e.lane == 0) {
    if (e.light) mbar_arrive_cluster_light(e.a_ready + 8 * (t * 4 + round));   // HOLD_TC_DBG & 128
    else mbar_arrive_cluster(e.a_ready + 8 * (t * 4 + round));
  }
}

// One (tile, step) of the epilogue.  KIND 0: forward layer l (softplus; REV: stash softplus'); 1: feature layer +
// start of the backward chain; 2: backward through layer l; 3: backward through layer 0 (embedding derivative).
template <int MODE, int KIND>
__device__ __forceinline__ bool t2_epi(const TcArgs& a, const T2Epi& e, int t, int step, int l, bool store_a, float px, float py,
                                       float pz, bool valid, int p, float& h0, float& h1, float& h2, float& h3, uint32_t& d_par, long long& t_wait) {
  const int di = t * 2 + (step & 1);
  if (!mbar_wait2t(e.d_full + 8 * di, (d_par >> di) & 1, a.err, 4, e.abort_flag, e.prof, t_wait)) return false;
  d_par ^= (1u << di);
  tc_fence_after();
  const uint32_t t_col = e.t_lane + (uint32_t)(di * 128 + e.sub * 8);
  uint32_t raw[8];
  tc_ld8(t_col, raw);
  float* sig = nullptr;
  if (MODE == MLP_SDF_REV) sig = a.sig + ((size_t)(blockIdx.x * 2 + t) * 8) * (kT2Rows * 256) + (size_t)e.row * 256;
  constexpr int kSigL = kT2Rows * 256;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n0 = e.qh * 128 + j * 32 + e.sub * 8;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (KIND == 0) {
      s0 = *reinterpret_cast<const float4*>(e.bias + l * 256 + n0);
      s1 = *reinterpret_cast<const float4*>(e.bias + l * 256 + n0 + 4);
    } else if (KIND == 1) {
      s0 = *reinterpret_cast<const float4*>(e.bias + 8 * 256 + n0);
      s1 = *reinterpret_cast<const float4*>(e.bias + 8 * 256 + n0 + 4);
    } else if (KIND == 2) {
      s0 = *reinterpret_cast<const float4*>(sig + (size_t)(l - 1) * kSigL + n0);
      s1 = *reinterpret_cast<const float4*>(sig + (size_t)(l - 1) * kSigL + n0 + 4);
    }
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    tc_wait_ld();
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = __uint_as_float(raw[i]);
    if (j < 3) tc_ld8(t_col + (uint32_t)((j + 1) * 32), raw);
    float out[8];
    if (KIND == 0) {
      float sg[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float z64 = fmaf(acc[i], kT2AccToZ64, sv[i]);
        float u = 0.f;
        if (a.dbg & 2) {
          out[i] = fmaxf(z64, 0.f);
        } else {
          u = mufu_ex2(-fabsf(z64 * kT2Z64ToT));
          const float L = mufu_lg2(1.0f + u);
          out[i] = fmaf(L, kT2LgToOut, fmaxf(z64, 0.f));
        }
        if (MODE == MLP_SDF_REV) {
          const float r = mufu_rcp(1.0f + u);
          sg[i] = (z64 >= 0.f) ? r : u * r;
        }
      }
      if (MODE == MLP_SDF_REV) {
        *reinterpret_cast<float4*>(sig + (size_t)l * kSigL + n0) = make_float4(sg[0], sg[1], sg[2], sg[3]);
        *reinterpret_cast<float4*>(sig + (size_t)l * kSigL + n0 + 4) = make_float4(sg[4], sg[5], sg[6], sg[7]);
      }
      if (l == 3 && n0 + 8 > kHidden - kEmbed) {  // skip connection: embedding columns of layer 3's output
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (n0 + i >= kHidden - kEmbed) out[i] = kTcScaleA * embed_val(n0 + i - (kHidden - kEmbed), 0, px, py, pz, a.embed_w);
      }
      if (l == 7) {
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w_last + n0));
        const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w_last + n0) + 1);
        h0 += out[0] * w0.x + out[1] * w0.y + out[2] * w0.z + out[3] * w0.w;
        h0 += out[4] * w1.x + out[5] * w1.y + out[6] * w1.z + out[7] * w1.w;
      }
    } else if (KIND == 1) {
      if (valid) {
        float4* dst = reinterpret_cast<float4*>(a.feat + (size_t)p * kFeat + n0);
        dst[0] = make_float4(fmaf(acc[0], kTcUnscale, sv[0]), fmaf(acc[1], kTcUnscale, sv[1]), fmaf(acc[2], kTcUnscale, sv[2]),
                             fmaf(acc[3], kTcUnscale, sv[3]));
        dst[1] = make_float4(fmaf(acc[4], kTcUnscale, sv[4]), fmaf(acc[5], kTcUnscale, sv[5]), fmaf(acc[6], kTcUnscale, sv[6]),
                             fmaf(acc[7], kTcUnscale, sv[7]));
      }
      const float4 g0 = *reinterpret_cast<const float4*>(sig + (size_t)7 * kSigL + n0);
      const float4 g1 = *reinterpret_cast<const float4*>(sig + (size_t)7 * kSigL + n0 + 4);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w_last + n0));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w_last + n0) + 1);
      out[0] = kTcScaleA * w0.x * g0.x, out[1] = kTcScaleA * w0.y * g0.y, out[2] = kTcScaleA * w0.z * g0.z, out[3] = kTcScaleA * w0.w * g0.w;
      out[4] = kTcScaleA * w1.x * g1.x, out[5] = kTcScaleA * w1.y * g1.y, out[6] = kTcScaleA * w1.z * g1.z, out[7] = kTcScaleA * w1.w * g1.w;
    } else if (KIND == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = acc[i] * kT2AccToZ64 * sv[i];
      if (l == 4 && n0 + 8 > kHidden - kEmbed) {  // skip input of layer 4: columns 217.. are d sdf / d embed
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (n0 + i >= kHidden - kEmbed) {
            const int em = n0 + i - (kHidden - kEmbed), d = em % 3;
            const float je = acc[i] * kTcUnscale * embed_val(em, d + 1, px, py, pz, a.embed_w);
            h1 += (d == 0) ? je : 0.f;
            h2 += (d == 1) ? je : 0.f;
            h3 += (d == 2) ? je : 0.f;
            out[i] = 0.f;
          }
        }
      }
    } else {  // KIND 3: d sdf / d embed through layer 0's input
      if (n0 < 40) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int em = n0 + i;
          if (em < kEmbed) {
            const int d = em % 3;
            const float je = acc[i] * kTcUnscale * embed_val(em, d + 1, px, py, pz, a.embed_w);
            h1 += (d == 0) ? je : 0.f;
            h2 += (d == 1) ? je : 0.f;
            h3 += (d == 2) ? je : 0.f;
          }
        }
      }
    }
    if (KIND != 3 && store_a) t2_store_a(e, t, n0, e.coarse ? 0 : j, !e.coarse || j == 3, out);
  }
  return true;
}

// SDF-only forward step with the FAST epilogue shape (HOLD_TC_DBG & 256; implies coarse hand-offs): 2 rounds of 16 columns per
// warp and tile-step, one fence + one arrival at the end, no experiment branches in the element loop.
__device__ __forceinline__ bool t2_epi_wide0(const TcArgs& a, const T2Epi& e, int t, int step, bool store_a, float px, float py, float pz,
                                             float& h0, uint32_t& d_par, long long& t_wait) {
  const int l = step;
  const int di = t * 2 + (step & 1);
  if (!mbar_wait2t(e.d_full + 8 * di, (d_par >> di) & 1, a.err, 4, e.abort_flag, e.prof, t_wait)) return false;
  d_par ^= (1u << di);
  tc_fence_after();
  const uint32_t t_col = e.t_lane + (uint32_t)(di * 128 + e.sub * 16);
  uint32_t raw[16];
  tc_ld16(t_col, raw);
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int n0 = e.qh * 128 + j * 64 + e.sub * 16;
    float4 b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) b[u] = *reinterpret_cast<const float4*>(e.bias + l * 256 + n0 + 4 * u);
    tc_wait_ld();
    float out[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = __uint_as_float(raw[i]);
    if (j == 0) tc_ld16(t_col + 64u, raw);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float bb[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float z64 = fmaf(out[4 * u + i], kT2AccToZ64, bb[i]);
        const float uu = mufu_ex2(-fabsf(z64 * kT2Z64ToT));
        out[4 * u + i] = fmaf(mufu_lg2(1.0f + uu), kT2LgToOut, fmaxf(z64, 0.f));
      }
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
        h0 += out[4 * u] * w.x + out[4 * u + 1] * w.y + out[4 * u + 2] * w.z + out[4 * u + 3] * w.w;
      }
    }
    if (store_a) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint4 hi, lo;
        split8(out + 8 * u, hi, lo);
        const int k = n0 + 8 * u, c64 = k >> 6, ju = (k & 63) >> 3;
        const uint32_t off = (uint32_t)(t * kT2ATile + c64 * 8192 + (e.row >> 3) * 1024 + (e.row & 7) * 128 + ((ju ^ (e.row & 7)) << 4));
        *reinterpret_cast<uint4*>(e.gA + off) = hi;
        *reinterpret_cast<uint4*>(e.gA + off + kT2APart) = lo;
      }
    }
  }
  if (store_a) {   // one arrival per warp and tile-step on a_ready[t][0]
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (e.lane == 0) {
      if (e.light) mbar_arrive_cluster_light(e.a_ready + 8 * (t * 4));
      else mbar_arrive_cluster(e.a_ready + 8 * (t * 4));
    }
  }
  return true;
}

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreadsTotal, 1) k_mlp_tc2(TcArgs a) {
  static_assert(MODE == MLP_SDF_ONLY || MODE == MLP_SDF_REV, "pair kernel: SDF chains only");
  if (a.st != nullptr && a.st->done) return;
  constexpr int NS = kT2Stages;
  constexpr int NSTEP = (MODE == MLP_SDF_REV) ? 17 : 8;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sW = base + kT2SmemA, sBias = sW + kT2SmemW, sBar = sBias + kT2SmemBias;
  const uint32_t bWFull = sBar, bWEmpty = sBar + 8 * NS, bAReady = sBar + 16 * NS, bDFull = bAReady + 64;
  const uint32_t sTmemPtr = bDFull + 32, sAbort = sTmemPtr + 4;
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));  // generic pointer to `base`
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  volatile int* abort_flag = reinterpret_cast<volatile int*>(gen_base + (sAbort - base));
  const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int n_super = ceil_div(a.P, 2 * kT2TilePts);
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;
  const bool prof = (a.prof != nullptr) && (cluster_id == 0);
  const uint32_t ns_eff = (a.dbg & 16) ? 3u : (uint32_t)NS;
  long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;  // profiling accumulators of this thread
  const long long t_begin = prof ? clock64() : 0;

  if (threadIdx.x == 0) {
    // leader's w_full: its own expect_tx arrival + the peer's forwarded "my half has landed"
    for (int i = 0; i < NS; ++i) { mbar_init(bWFull + 8 * i, rank == 0 ? 2 : 1); mbar_init(bWEmpty + 8 * i, 1); }
    *abort_flag = 0;
    for (int i = 0; i < 8; ++i) mbar_init(bAReady + 8 * i, 32);  // a_ready[tile][round]: all 16 epilogue warps of both CTAs
    for (int i = 0; i < 4; ++i) mbar_init(bDFull + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {  // biases to shared memory: layers 0..7 as 64 b (operand scale), the feature layer's as is
    float* sb = reinterpret_cast<float*>(gen_base + (sBias - base));
    constexpr int NB = (MODE == MLP_SDF_REV) ? 9 : 8;
    for (int i = threadIdx.x; i < NB * 256; i += blockDim.x) {
      const int l = i >> 8;
      sb[i] = a.L[l].bias[i & 255] * ((l < 8) ? kTcScaleA : 1.0f);
    }
  }
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sTmemPtr), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen_base + (sTmemPtr - base));

  if (warp == 0) {
    // ============================================================ weight producer: this CTA's half of every stage
    // (all lanes walk the loops, one elected lane issues: see elect_one() in mlp_tc.cuh)
    {
      uint32_t stage = 0, phase = 0;
      for (int su = cluster_id; su < n_super; su += n_clusters) {
        for (int step = 0; step < NSTEP; ++step) {
          const uint8_t* src = a.L[step].wimg + (size_t)rank * 8192 +
                               (a.wcopies > 1 ? (size_t)(cluster_id % a.wcopies) * a.L[step].nst * kTcStageBytes : 0);
          const int nst = a.L[step].nst;
          for (int t = 0; t < 2; ++t) {
            for (int si = 0; si < nst; ++si) {
              const int c = (nst == 8) ? ((si >> 1) + 4 * (si & 1)) : si;
              if (!__all_sync(0xffffffffu, mbar_wait2t(bWEmpty + 8 * stage, phase ^ 1, a.err, 1, abort_flag, prof, tp0))) goto tc2_done;
              if (elect_one()) {
                if (a.dbg & 8) {
                  mbar_arrive(bWFull + 8 * stage);
                } else {
                  mbar_expect_tx(bWFull + 8 * stage, kT2HalfStage);
                  bulk_g2s(sW + stage * kT2HalfStage, src + (size_t)c * kTcStageBytes, 8192, bWFull + 8 * stage);
                  bulk_g2s(sW + stage * kT2HalfStage + 8192, src + (size_t)c * kTcStageBytes + 16384, 8192, bWFull + 8 * stage);
                }
              }
              __syncwarp();
              if (++stage == ns_eff) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      if (prof && lane == 0) { a.prof[8 + 2 * rank] = clock64() - t_begin; a.prof[9 + 2 * rank] = tp0; }
    }
  } else if (warp == 1) {
    if (rank == 1) {
      // ========================================================== peer: forward "my half has landed" to the leader's w_full
      uint32_t stage = 0, phase = 0;
      const uint32_t leader_full = mapa_rank(bWFull, 0);
      for (int su = cluster_id; su < n_super; su += n_clusters)
        for (int step = 0; step < NSTEP; ++step)
          for (int k = 0; k < 2 * a.L[step].nst; ++k) {
            if (!__all_sync(0xffffffffu, mbar_wait2t(bWFull + 8 * stage, phase, a.err, 5, abort_flag, prof, tp0))) goto tc2_done;
            if (elect_one()) mbar_arrive_cluster(leader_full + 8 * stage);
            __syncwarp();
            if (++stage == ns_eff) { stage = 0; phase ^= 1; }
          }
      if (prof && lane == 0) a.prof[12] = tp0;
    } else {
      // ========================================================== leader: MMA issuer for the pair
      uint32_t stage = 0, phase = 0;
      uint32_t a_par = 0;  // bit t*4+j = parity to wait for on a_ready[t][j]
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      for (int su = cluster_id; su < n_super; su += n_clusters) {
        for (int step = 0; step < NSTEP; ++step) {
          const int nst = a.L[step].nst;
          for (int t = 0; t < 2; ++t) {
            const uint32_t d_tmem = tmem_u + (uint32_t)((t * 2 + (step & 1)) * 128);
            const uint32_t aT = sA + t * kT2ATile;
            for (int si = 0; si < nst; ++si) {
              const int c = (nst == 8) ? ((si >> 1) + 4 * (si & 1)) : si;
              if ((si & 1) == 0 && (si == 0 || !(a.dbg & (32 | 256)))) {  // one hand-off barrier per epilogue round: chunks j and 4 + j
                // (layer 0: chunks 0, 1; coarse mode: one barrier per tile-step)
                const int bi = t * 4 + (si >> 1);
                if (!__all_sync(0xffffffffu, mbar_wait2t(bAReady + 8 * bi, (a_par >> bi) & 1, a.err, 2, abort_flag, prof, tp1))) goto tc2_done;
                a_par ^= (1u << bi);
              }
              // own half (expect_tx + bytes) and the peer's forwarded arrival complete the same barrier
              if (!__all_sync(0xffffffffu, mbar_wait2t(bWFull + 8 * stage, phase, a.err, 3, abort_flag, prof, tp2))) goto tc2_done;
              tc_fence_after();
              const uint32_t wb = sW + stage * kT2HalfStage;
              const bool el = elect_one();
#pragma unroll
              for (int jj = 0; jj < 2; ++jj) {
                const uint32_t koff = (uint32_t)(((c & 1) * 2 + jj) * 32);  // bytes inside the 128-byte A row
                const uint64_t ahi = umma_desc(aT + (c >> 1) * 8192 + koff, 1024, kLayoutSW128);
                const uint64_t alo = umma_desc(aT + kT2APart + (c >> 1) * 8192 + koff, 1024, kLayoutSW128);
                const uint64_t whi = umma_desc(wb + jj * 32, 512, kLayoutSW64);
                const uint64_t wlo = umma_desc(wb + 8192 + jj * 32, 512, kLayoutSW64);
                if (el) {
                  tc_mma2(d_tmem, ahi, whi, kIdescF16, (si | jj) != 0);
                  if (!(a.dbg & 4)) tc_mma2(d_tmem, alo, whi, kIdescF16, 1);
                  if (!(a.dbg & 1)) tc_mma2(d_tmem, ahi, wlo, kIdescF16, 1);
                }
              }
              if (el) tc_commit2(bWEmpty + 8 * stage);  // frees the stage in both CTAs when these MMAs have read it
              __syncwarp();
              if (++stage == ns_eff) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit2(bDFull + 8 * (t * 2 + (step & 1)));  // accumulator of (tile t, step) complete, both CTAs
            __syncwarp();
          }
        }
      }
      if (prof && lane == 0) { a.prof[0] = clock64() - t_begin; a.prof[1] = tp1; a.prof[2] = tp2; a.prof[3] = 0; }
    }
  } else {
    // ============================================================ epilogue
    T2Epi e;
    const int q = warp & 3;
    e.qh = q >> 1, e.sub = (warp - 2) >> 2, e.lane = lane, e.row = (q & 1) * 32 + lane;
    e.t_lane = tmem + ((uint32_t)(q * 32) << 16);
    e.gA = gen_base;
    e.bias = reinterpret_cast<const float*>(gen_base + (sBias - base));
    e.a_ready = mapa_rank(bAReady, 0);
    e.d_full = bDFull;
    e.abort_flag = abort_flag;
    e.prof = prof;
    e.coarse = (a.dbg & (32 | 256)) != 0;
    e.light = (a.dbg & 128) != 0;
    const int w8 = e.qh * 4 + e.sub;  // index among the 8 warps that share this thread's row
    uint32_t d_par = 0;               // bit t*2+b = parity to wait for on d_full[t][b]
    for (int su = cluster_id; su < n_super; su += n_clusters) {
      const int p0 = su * (2 * kT2TilePts) + (int)rank * kT2Rows + e.row, p1 = p0 + kT2TilePts;
      const bool v0 = p0 < a.P, v1 = p1 < a.P;
      float x0 = 0.f, y0 = 0.f, z0 = 0.f, x1 = 0.f, y1 = 0.f, z1 = 0.f;
      if (v0) { x0 = a.xc[3 * (size_t)p0], y0 = a.xc[3 * (size_t)p0 + 1], z0 = a.xc[3 * (size_t)p0 + 2]; }
      if (v1) { x1 = a.xc[3 * (size_t)p1], y1 = a.xc[3 * (size_t)p1 + 1], z1 = a.xc[3 * (size_t)p1 + 2]; }
      // ---------------------------------------------------------- prologue: layer-0 A operand, k-chunk qh of both tiles
      for (int t = 0; t < 2; ++t) {
        const float px = t ? x1 : x0, py = t ? y1 : y0, pz = t ? z1 : z0;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = kTcScaleA * embed_val(e.qh * 32 + e.sub * 8 + i, 0, px, py, pz, a.embed_w);
        t2_store_a(e, t, e.qh * 32 + e.sub * 8, 0, true, x);
      }
      float ha0 = 0.f, ha1 = 0.f, ha2 = 0.f, ha3 = 0.f, hb0 = 0.f, hb1 = 0.f, hb2 = 0.f, hb3 = 0.f;
      bool ok = true;
#define T2_STEP(KIND, STEP, L, STORE)                                                                                              \
  for (int t = 0; t < 2 && ok; ++t) {                                                                                              \
    float h0 = t ? hb0 : ha0, h1 = t ? hb1 : ha1, h2 = t ? hb2 : ha2, h3 = t ? hb3 : ha3;                                          \
    ok = t2_epi<MODE, KIND>(a, e, t, STEP, L, STORE, t ? x1 : x0, t ? y1 : y0, t ? z1 : z0, t ? v1 : v0, t ? p1 : p0, h0, h1, h2, \
                            h3, d_par, tp0);                                                                                       \
    if (t) { hb0 = h0, hb1 = h1, hb2 = h2, hb3 = h3; } else { ha0 = h0, ha1 = h1, ha2 = h2, ha3 = h3; }                            \
  }
      if (MODE == MLP_SDF_ONLY && (a.dbg & 256)) {
        for (int l = 0; l < 8 && ok; ++l)
          for (int t = 0; t < 2 && ok; ++t) {
            float h0 = t ? hb0 : ha0;
            ok = t2_epi_wide0(a, e, t, l, l < 7, t ? x1 : x0, t ? y1 : y0, t ? z1 : z0, h0, d_par, tp0);
            if (t) hb0 = h0; else ha0 = h0;
          }
      } else {
        for (int l = 0; l < 8 && ok; ++l) { T2_STEP(0, l, l, (MODE == MLP_SDF_REV) || l < 7) }
      }
      if (MODE == MLP_SDF_REV) {
        if (ok) { T2_STEP(1, 8, 8, true) }
        for (int step = 9; step < 16 && ok; ++step) { T2_STEP(2, step, 16 - step, true) }
        if (ok) { T2_STEP(3, 16, 0, false) }
      }
#undef T2_STEP
      // ---------------------------------------------------------- heads: fixed-order reduction over the 8 warps of a row
      // (all MMAs of both tiles have completed, so the A regions are free to hold the partial sums)
      constexpr int NH = (MODE == MLP_SDF_REV) ? 4 : 1;
      tc_fence_before();
      float* scr0 = reinterpret_cast<float*>(gen_base);
      float* scr1 = reinterpret_cast<float*>(gen_base + kT2ATile);
      scr0[(w8 * NH + 0) * kT2Rows + e.row] = ha0;
      scr1[(w8 * NH + 0) * kT2Rows + e.row] = hb0;
      if (MODE == MLP_SDF_REV) {
        scr0[(w8 * NH + 1) * kT2Rows + e.row] = ha1, scr0[(w8 * NH + 2) * kT2Rows + e.row] = ha2, scr0[(w8 * NH + 3) * kT2Rows + e.row] = ha3;
        scr1[(w8 * NH + 1) * kT2Rows + e.row] = hb1, scr1[(w8 * NH + 2) * kT2Rows + e.row] = hb2, scr1[(w8 * NH + 3) * kT2Rows + e.row] = hb3;
      }
      epi_bar();
      if (w8 == 0) {
        for (int t = 0; t < 2; ++t) {
          const float* scr = t ? scr1 : scr0;
          const int p = t ? p1 : p0;
          if (!(t ? v1 : v0)) continue;
          float hs[NH];
#pragma unroll
          for (int k = 0; k < NH; ++k) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) s += scr[(w * NH + k) * kT2Rows + e.row];
            hs[k] = s;
          }
          a.sdf[p] = hs[0] * (1.0f / kTcScaleA) + a.b_last[0];
          if (MODE == MLP_SDF_REV) { a.grad[3 * (size_t)p] = hs[1], a.grad[3 * (size_t)p + 1] = hs[2], a.grad[3 * (size_t)p + 2] = hs[3]; }
        }
      }
      epi_bar();  // the scratch is overwritten by the next super-tile's prologue
    }
    if (prof && lane == 0 && (warp == 2 || warp == 17)) {
      const int o = 16 + 8 * (int)rank + (warp == 17 ? 4 : 0);
      a.prof[o] = clock64() - t_begin;
      a.prof[o + 1] = tp0;
    }
  }
tc2_done:
  tc_fence_before();
  cluster_sync_all();  // the peer's TMEM / smem must outlive the leader's last MMA and the last remote arrivals
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

static int tc2_init() {
  cudaError_t e = cudaFuncSetAttribute(k_mlp_tc2<MLP_SDF_ONLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT2SmemBytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_mlp_tc2<MLP_SDF_REV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kT2SmemBytes);
  if (e != cudaSuccess) { set_error("tcgen05 pair kernel attribute: %s", cudaGetErrorString(e)); return HOLD_E_CUDA; }
  return HOLD_OK;
}

// grid: one CTA pair per two SMs, each pair walks super-tiles of 256 points
static inline int tc2_grid(hold_ctx* ctx, int P) {
  const int n_super = ceil_div(P, 2 * kT2TilePts);
  return 2 * min(n_super, ctx->sm_count / 2);
}

// HOLD_TC_PAIR=1 selects the pair kernels (off by default until validated on hardware).
static inline bool tc2_enabled() {
  const char* e = getenv("HOLD_TC_PAIR");
  return e != nullptr && atoi(e) != 0;
}

// SDF net on P canonical points: sdf only (sampler rounds), or sdf + gradient + feature (reverse mode).
static int tc2_launch_sdf(hold_ctx* ctx, NodeState& ns, int P, const float* xc, const float* embed_w, float* sdf, float* grad,
                          float* feat, const SamplerState* st, cudaStream_t s) {
  const bool rev = (grad != nullptr) || (feat != nullptr);
  static const bool use_jvp = [] { const char* e = getenv("HOLD_TC_GRAD"); return e != nullptr && strcmp(e, "jvp") == 0; }();
  if (tc_fast_enabled() && !(rev && use_jvp)) {   // HOLD_TC_FAST=1: rebuilt epilogue (mlp_tc_fast.cuh)
    if (rev) return tc_fast_launch_rev(ctx, ns, P, xc, embed_w, sdf, grad, feat, s);
    return tc_fast_launch_sdf(ctx, ns, P, xc, embed_w, sdf, st, s);
  }
  if (!tc2_enabled() || (rev && use_jvp)) return tc_launch_sdf(ctx, ns, P, xc, embed_w, sdf, grad, feat, st, s);
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.P = P, a.n_layers = rev ? 17 : 8;
  for (int l = 0; l < 9; ++l) {
    a.L[l].wimg = ns.tc->sdf_img[l], a.L[l].bias = ns.sdf.bias[l], a.L[l].nst = ns.tc->sdf_nst[l], a.L[l].N = ns.sdf.N[l];
  }
  for (int i = 0; i < 8; ++i) {
    a.L[9 + i].wimg = ns.tc->sdf_imgT[7 - i], a.L[9 + i].bias = nullptr, a.L[9 + i].nst = 8, a.L[9 + i].N = 256;
  }
  a.w_last = ns.sdf.w_last, a.b_last = ns.sdf.b_last;
  a.xc = xc, a.embed_w = embed_w, a.sdf = sdf, a.grad = grad, a.feat = feat, a.st = st, a.err = ctx->dev_err;
  { const char* e = getenv("HOLD_TC_DBG"); a.dbg = e ? atoi(e) : 0; }
  if (getenv("HOLD_TC_PROF") != nullptr) {
    void* pr = nullptr;
    int rc = ws_get(ctx, 23 /* WS_PROF (debug) */, 64 * sizeof(long long), &pr);
    if (rc) return rc;
    a.prof = (long long*)pr;
  }
  const int grid = tc2_grid(ctx, P);
  {
    const int nrep = tc_fast_replicas(ctx, ns, s);   // HOLD_TC_WCOPIES=N: cluster c streams replica c % N
    HOLD_REQUIRE(nrep >= 0, "out of memory for weight-image replicas");
    if (nrep > 1) {
      a.wcopies = nrep;
      for (int l = 0; l < 9; ++l) a.L[l].wimg = ns.tc->sdf_img_rep[l];
      for (int i = 0; i < 8; ++i) a.L[9 + i].wimg = ns.tc->sdf_imgT_rep[7 - i];
    }
  }
  if (rev) {
    HOLD_REQUIRE(grad != nullptr && feat != nullptr, "sdf eval with gradient needs both grad and feat buffers");
    void* sig = nullptr;
    int rc = ws_get(ctx, 12 /* WS_SIG */, (size_t)grid * 2 * 8 * kT2Rows * 256 * sizeof(float), &sig);
    if (rc) return rc;
    a.sig = (float*)sig;
    k_mlp_tc2<MLP_SDF_REV><<<grid, kTcThreadsTotal, kT2SmemBytes, s>>>(a);
  } else {
    k_mlp_tc2<MLP_SDF_ONLY><<<grid, kTcThreadsTotal, kT2SmemBytes, s>>>(a);
  }
  HOLD_LAUNCH_CHECK(ctx);
  return HOLD_OK;
}

}  // namespace hold
