// Reverse mode of the inverse warps (KNNDeformer.forward(inverse=True) + skinning(), model/mano/deformer.py:34-68,145-170;
// ObjectDeformer.forward(inverse=True), model/obj/deformer.py:10-31) with respect to the bone / object transforms — the link
// between a gradient on canonical points (e.g. dL/dsdf * d sdf/d x_c from hold_sdf_eval) and the pose servers' backward
// (hold_mano_lbs_bwd takes g_tfs): the "joint SDF + LBS backward" of BASELINE.json configs[4].  The skinning weights are
// detached in the reference (deformer.py:101), so only the blended transform T = sum_j w_j tfs_j carries gradient:
//   x_c = M^-1 (x - t / s),  M = T[:3,:3], t = T[:3,3], s = T[3,3]
//   q = M^-T g;  dL/dM = -q x_c^T;  dL/dt = -q / s;  dL/ds = (q . t) / s^2;  dL/dx = q
// Written as per-thread phases (host-executable, tests/host/warp_bwd_host.cpp); sums are fixed-order.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HOLD_HD __host__ __device__ __forceinline__
#else
#ifndef HOLD_HD
#define HOLD_HD inline
#endif
#endif

namespace hold {
namespace warpbwd {

constexpr int kK = 15, kJ = 16, kG = 13;   // neighbours, bones, gradient entries per transform: rows 0..2 (12) + [3][3]

HOLD_HD bool inv3x3(const float* A /*row stride 4*/, float* Ai) {
  const float a = A[0], b = A[1], c = A[2], d = A[4], e = A[5], f = A[6], g = A[8], h = A[9], i = A[10];
  const float c0 = e * i - f * h, c1 = f * g - d * i, c2 = d * h - e * g;
  const float id = 1.0f / (a * c0 + b * c1 + c * c2);
  Ai[0] = c0 * id, Ai[1] = (c * h - b * i) * id, Ai[2] = (b * f - c * e) * id;
  Ai[3] = c1 * id, Ai[4] = (a * i - c * g) * id, Ai[5] = (c * d - a * f) * id;
  Ai[6] = c2 * id, Ai[7] = (b * g - a * h) * id, Ai[8] = (a * e - b * d) * id;
  return isfinite(id);
}

// gradient of x_c = M^-1 (x - t/s) w.r.t. the 3x4 top of T and s; G[12] row-major rows 0..2, G[12] = d/ds.  Returns q = dL/dx.
HOLD_HD void point_grad(const float* T /*12: rows 0..2 of the blended transform*/, float s, const float* x, const float* g,
                        float* G /*13*/, float* q /*3*/) {
  float Ai[9];
  inv3x3(T, Ai);
  const float u[3] = {x[0] - T[3] / s, x[1] - T[7] / s, x[2] - T[11] / s};
  float xc[3];
  for (int r = 0; r < 3; ++r) xc[r] = Ai[3 * r] * u[0] + Ai[3 * r + 1] * u[1] + Ai[3 * r + 2] * u[2];
  for (int c = 0; c < 3; ++c) q[c] = Ai[c] * g[0] + Ai[3 + c] * g[1] + Ai[6 + c] * g[2];   // M^-T g
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) G[4 * r + c] = -q[r] * xc[c];
    G[4 * r + 3] = -q[r] / s;
  }
  G[12] = (q[0] * T[3] + q[1] * T[7] + q[2] * T[11]) / (s * s);
}

// hand: skinning weights of one point from its 15 neighbours (query_skinning_weights_multi, deformer.py:84-105)
HOLD_HD void hand_weights(const float* x, const int* idx /*15*/, const float* verts /*[778,3]*/, const float* skin_w /*[778,16]*/,
                          float* w /*16*/) {
  float conf[kK], csum = 0.f;
  for (int k = 0; k < kK; ++k) {
    const float* v = verts + 3 * idx[k];
    const float dx = x[0] - v[0], dy = x[1] - v[1], dz = x[2] - v[2];
    const float d2 = (dx * dx + dy * dy) + dz * dz;
    conf[k] = expf(-fminf(d2, 4.0f));
    csum += conf[k];
  }
  for (int j = 0; j < kJ; ++j) w[j] = 0.f;
  for (int k = 0; k < kK; ++k) {
    const float c = conf[k] / csum;
    for (int j = 0; j < kJ; ++j) w[j] += skin_w[idx[k] * kJ + j] * c;
  }
}

// scratch of one block of nt points: w [nt][16], G [nt][13]
HOLD_HD int scratch_floats(int nt) { return nt * (kJ + kG); }

// P0 (hand): thread = point p (or idle): weights, blended transform, per-point gradient into the scratch; g_x optional
HOLD_HD void hand_p0(int tid, int nt, int p, int P, const float* x, const int* knn_idx, const float* verts, const float* skin_w,
                     const float* tfs /*[16,16]*/, const float* g_xc, float* g_x, float* scr) {
  float* w = scr + tid * kJ;
  float* G = scr + nt * kJ + tid * kG;
  if (p >= P) {
    for (int j = 0; j < kJ; ++j) w[j] = 0.f;
    for (int e = 0; e < kG; ++e) G[e] = 0.f;
    return;
  }
  hand_weights(x + 3 * p, knn_idx + kK * p, verts, skin_w, w);
  float T[12], s = 0.f;
  for (int e = 0; e < 12; ++e) T[e] = 0.f;
  for (int j = 0; j < kJ; ++j) {
    for (int e = 0; e < 12; ++e) T[e] += w[j] * tfs[j * 16 + e];
    s += w[j] * tfs[j * 16 + 15];
  }
  float q[3];
  point_grad(T, s, x + 3 * p, g_xc + 3 * p, G, q);
  if (g_x != nullptr) { g_x[3 * p] = q[0], g_x[3 * p + 1] = q[1], g_x[3 * p + 2] = q[2]; }
}
// P1 (hand): thread o < 16 * 13: partial[j][e] = sum over the block's points of w[p][j] * G[p][e]
HOLD_HD void hand_p1(int tid, int nt, const float* scr, float* partial /*[16*13] of this block*/) {
  for (int o = tid; o < kJ * kG; o += nt) {
    const int j = o / kG, e = o % kG;
    float acc = 0.f;
    for (int p = 0; p < nt; ++p) acc += scr[p * kJ + j] * scr[nt * kJ + p * kG + e];
    partial[o] = acc;
  }
}
// final: g_tfs[j] (4x4) = sum over blocks, fixed order; thread o < 16 * 16
HOLD_HD void hand_final(int o, int n_blocks, const float* partials /*[n_blocks][16*13]*/, float* g_tfs /*[16,16]*/) {
  const int j = o / 16, e = o % 16;
  const int r = e / 4, c = e % 4;
  float acc = 0.f;
  if (r < 3 || c == 3) {
    const int ge = (r < 3) ? e : 12;
    for (int b = 0; b < n_blocks; ++b) acc += partials[(size_t)b * (kJ * kG) + j * kG + ge];
  }
  g_tfs[j * 16 + e] = acc;
}

// object: one rigid transform per frame; scratch G [nt][13]
HOLD_HD void obj_p0(int tid, int nt, int p, int P, const float* x, const float* tf /*[16]*/, const float* g_xc, float* g_x, float* scr) {
  (void)nt;
  float* G = scr + tid * kG;
  if (p >= P) {
    for (int e = 0; e < kG; ++e) G[e] = 0.f;
    return;
  }
  float q[3];
  point_grad(tf, tf[15], x + 3 * p, g_xc + 3 * p, G, q);
  if (g_x != nullptr) { g_x[3 * p] = q[0], g_x[3 * p + 1] = q[1], g_x[3 * p + 2] = q[2]; }
}
HOLD_HD void obj_p1(int tid, int nt, const float* scr, float* partial /*[13]*/) {
  for (int e = tid; e < kG; e += nt) {
    float acc = 0.f;
    for (int p = 0; p < nt; ++p) acc += scr[p * kG + e];
    partial[e] = acc;
  }
}
HOLD_HD void obj_final(int e, int n_blocks, const float* partials /*[n_blocks][13]*/, float* g_tf /*[16]*/) {
  const int r = e / 4, c = e % 4;
  float acc = 0.f;
  if (r < 3 || c == 3) {
    const int ge = (r < 3) ? e : 12;
    for (int b = 0; b < n_blocks; ++b) acc += partials[(size_t)b * kG + ge];
  }
  g_tf[e] = acc;
}

}  // namespace warpbwd
}  // namespace hold
