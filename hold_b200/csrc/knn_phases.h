// Exact K = 15 nearest MANO vertices with a cheap FILTER scan (HOLD_KNN_FILTER=1, round-2 A/B of k_inverse_warp_hand_rays).
// The production scan evaluates the reference's distance expression ((p - v)^2).sum() — 3 shared loads + 9 arithmetic
// instructions per vertex, 778 vertices per sample.  Here every vertex is first tested with the expanded form
//   d~ = |v|^2 - 2 p.v   (one 16-byte shared load {x, y, z, |v|^2} + 3 FMA)   against   tau - |p|^2 + margin,
// tau = the 15th distance of the previous sample's neighbours re-ranked for this sample (an upper bound of the true 15th
// distance).  `margin` bounds the rounding difference between the two forms (6 roundings at magnitude (|p| + |v|)^2), so the
// filter can only ADD candidates; survivors are re-evaluated with the reference expression and merged in lexicographic
// (distance, index) order — the result is bit-identical to the full exact scan.  Compiles for the host as well
// (tests/host/knn_host.cpp checks it against brute force).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HOLD_HD __host__ __device__ __forceinline__
#else
#ifndef HOLD_HD
#define HOLD_HD inline
#endif
#endif
#if defined(__CUDA_ARCH__)
#define HOLD_FADD(a, b) __fadd_rn((a), (b))
#define HOLD_FMUL(a, b) __fmul_rn((a), (b))
#else
#define HOLD_FADD(a, b) ((a) + (b))   // host build uses -ffp-contract=off
#define HOLD_FMUL(a, b) ((a) * (b))
#endif

namespace hold {
namespace knnf {

constexpr int kK = 15, kNV = 778, kCand = 48;
struct Top {
  float d[kK];
  int i[kK];
};
struct V4 { float x, y, z, q; };   // q = |v|^2

// the reference's expression, no FMA contraction (KNN selection is sensitive to the last bit)
HOLD_HD float exact_dist(const V4& v, float px, float py, float pz) {
  const float dx = px - v.x, dy = py - v.y, dz = pz - v.z;
  return HOLD_FADD(HOLD_FADD(HOLD_FMUL(dx, dx), HOLD_FMUL(dy, dy)), HOLD_FMUL(dz, dz));
}
HOLD_HD void insert(Top& r, float dist, int v) {   // replace the last entry, bubble up; order (distance, index)
  r.d[kK - 1] = dist;
  r.i[kK - 1] = v;
  for (int k = kK - 1; k > 0; --k) {
    const bool sw = (r.d[k] < r.d[k - 1]) || (r.d[k] == r.d[k - 1] && r.i[k] < r.i[k - 1]);
    if (sw) {
      const float td = r.d[k]; r.d[k] = r.d[k - 1]; r.d[k - 1] = td;
      const int ti = r.i[k]; r.i[k] = r.i[k - 1]; r.i[k - 1] = ti;
    }
  }
}
HOLD_HD void full_scan(const V4* sv, float px, float py, float pz, Top& r) {
  for (int k = 0; k < kK; ++k) { r.d[k] = 3.0e38f; r.i[k] = 0x7fffffff; }
  for (int v = 0; v < kNV; ++v) {
    const float dist = exact_dist(sv[v], px, py, pz);
    if (dist < r.d[kK - 1] || (dist == r.d[kK - 1] && v < r.i[kK - 1])) insert(r, dist, v);
  }
}
// r holds the previous sample's neighbours on entry, this sample's on exit.  qmax = max |v|^2 over the vertices.
HOLD_HD void seeded_filter(const V4* sv, float qmax, float px, float py, float pz, Top& r, unsigned short* cand) {
  int seed[kK];
  for (int k = 0; k < kK; ++k) { seed[k] = r.i[k]; r.d[k] = 3.0e38f; r.i[k] = 0x7fffffff; }
  for (int s = 0; s < kK; ++s) insert(r, exact_dist(sv[seed[s]], px, py, pz), seed[s]);
  const float tau = r.d[kK - 1];
  const float pp = px * px + py * py + pz * pz;
  const float margin = 2.0e-6f * (pp + qmax) + 1.0e-30f;
  const float lim = (tau - pp) + margin;
  const float ax = -2.0f * px, ay = -2.0f * py, az = -2.0f * pz;
  int cnt = 0;
  for (int v = 0; v < kNV; ++v) {
    const float dt = fmaf(ax, sv[v].x, fmaf(ay, sv[v].y, fmaf(az, sv[v].z, sv[v].q)));
    if (dt <= lim) {
      if (cnt < kCand) cand[cnt] = (unsigned short)v;
      ++cnt;
    }
  }
  if (cnt > kCand) {   // seeds were poor (large step along the ray): exact full scan
    full_scan(sv, px, py, pz, r);
    return;
  }
  for (int c = 0; c < cnt; ++c) {
    const int v = cand[c];
    bool present = false;
    for (int k = 0; k < kK; ++k) present |= (r.i[k] == v);
    if (present) continue;
    const float dist = exact_dist(sv[v], px, py, pz);
    if (dist < r.d[kK - 1] || (dist == r.d[kK - 1] && v < r.i[kK - 1])) insert(r, dist, v);
  }
}

}  // namespace knnf
}  // namespace hold
