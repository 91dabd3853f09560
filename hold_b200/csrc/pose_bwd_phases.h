// Reverse mode of the two pose servers (GenericServer.forward, model/mano/server.py:62-99 incl. lbs(),
// utils/external/lbs.py:139-399; ObjectModel.forward, model/obj/object_model.py:29-70) — what torch.autograd does for
// fitting/model.py:117 (optimize_ckpt.py).  One thread block per frame; the work is written as PHASES separated by
// block barriers, each phase a function of (tid, nt) over a per-frame scratch struct.  No shuffles, no atomics: every
// reduction is a fixed-order serial sum, so results are deterministic — and the same phase functions compile for the
// host (tests/host/pose_bwd_host.cpp runs them tid by tid on the CPU against torch.autograd).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HOLD_HD __host__ __device__ __forceinline__
#else
#define HOLD_HD inline
#endif

namespace hold {
namespace posebwd {

constexpr int kV = 778, kJ = 16, kNP = 135;  // MANO vertices, bones, pose-feature length

struct ManoPtrs {  // model (device or host pointers)
  const float *v_template, *shapedirs, *posedirs, *J_regressor, *lbs_weights, *hands_mean;
  int parents[kJ];
  int tips[5];
};

struct ManoFrame {  // one frame's inputs, upstream gradients (nullable) and outputs
  const float *betas, *full_pose, *transl, *scene_scale;  // [10], [48], [3], [1]
  const float* tfs_c_inv;                                  // [16,16] or NULL
  const float *g_verts, *g_jnts, *g_tfs;                   // [778,3], [21,3], [16,16]; each may be NULL
  float *g_betas, *g_pose, *g_transl, *g_scale;            // [10], [48], [3], [1]
};

// scratch (shared memory on the device): sizes in floats
constexpr int kScrBufA = 0;                         // v_shaped, later g_o                  [2334]
constexpr int kScrBufB = kScrBufA + kV * 3;         // v_posed                               [2334]
constexpr int kScrBufC = kScrBufB + kV * 3;         // g_vp, later g_vs                      [2334]
constexpr int kScrJ = kScrBufC + kV * 3;            // joints                                [48]
constexpr int kScrR = kScrJ + kJ * 3;               // rotation matrices                     [144]
constexpr int kScrG = kScrR + kJ * 9;               // world chain                           [256]
constexpr int kScrA = kScrG + kJ * 16;              // relative transforms                   [256]
constexpr int kScrgA = kScrA + kJ * 16;             // d/dA rows 0..2                         [192]
constexpr int kScrgR = kScrgA + kJ * 12;            // d/dR                                   [144]
constexpr int kScrgJ = kScrgR + kJ * 9;             // d/dJ                                   [48]
constexpr int kScrgPf = kScrgJ + kJ * 3;            // d/d pose feature                       [136]
constexpr int kScrPose = kScrgPf + 136;             // pose + mean                            [48]
constexpr int kScrRed = kScrPose + 48;              // per-thread partial sums                [4 * nt]
HOLD_HD int mano_scratch_floats(int nt) { return kScrRed + 4 * nt; }

HOLD_HD void rodrigues_fwd(const float* rv, float* Rm) {  // batch_rodrigues, lbs.py:298-329 (same as geom.cuh)
  const float ax = rv[0] + 1e-8f, ay = rv[1] + 1e-8f, az = rv[2] + 1e-8f;
  const float ang = sqrtf(ax * ax + ay * ay + az * az);
  const float rx = rv[0] / ang, ry = rv[1] / ang, rz = rv[2] / ang;
  const float c = cosf(ang), s = sinf(ang);
  const float Km[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
  float KK[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) KK[3 * i + j] = Km[3 * i] * Km[j] + Km[3 * i + 1] * Km[3 + j] + Km[3 * i + 2] * Km[6 + j];
  for (int e = 0; e < 9; ++e) Rm[e] = ((e % 4 == 0) ? 1.f : 0.f) + s * Km[e] + (1.f - c) * KK[e];
}

// d/d rv of rodrigues_fwd: R = I + sin(n) K(u) + (1 - cos n) K(u)^2, n = |rv + 1e-8|, u = rv / n
HOLD_HD void rodrigues_bwd(const float* rv, const float* gR, float* g_rv) {
  const float a[3] = {rv[0] + 1e-8f, rv[1] + 1e-8f, rv[2] + 1e-8f};
  const float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const float u[3] = {rv[0] / n, rv[1] / n, rv[2] / n};
  const float c = cosf(n), s = sinf(n);
  const float K[9] = {0.f, -u[2], u[1], u[2], 0.f, -u[0], -u[1], u[0], 0.f};
  float KK[9], M[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) KK[3 * i + j] = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
  float g_n = 0.f;
  for (int e = 0; e < 9; ++e) g_n += gR[e] * (c * K[e] + s * KK[e]);
  // M = s gR + (1 - c) (gR K^T + K^T gR)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float t1 = 0.f, t2 = 0.f;
      for (int k = 0; k < 3; ++k) {
        t1 += gR[3 * i + k] * K[3 * j + k];   // gR K^T
        t2 += K[3 * k + i] * gR[3 * k + j];   // K^T gR
      }
      M[3 * i + j] = s * gR[3 * i + j] + (1.f - c) * (t1 + t2);
    }
  const float g_u[3] = {M[7] - M[5], M[2] - M[6], M[3] - M[1]};
  const float dotp = rv[0] * g_u[0] + rv[1] * g_u[1] + rv[2] * g_u[2];
  const float inv = 1.0f / n, inv3 = inv * inv * inv;
  for (int k = 0; k < 3; ++k) g_rv[k] = g_n * a[k] * inv + g_u[k] * inv - a[k] * dotp * inv3;
}

// ---------------------------------------------------------------------------------------------- MANO phases
// P0: v_shaped, pose (+ mean), rotation matrices
HOLD_HD void mano_p0(int tid, int nt, const ManoPtrs& m, const ManoFrame& f, float* scr) {
  float* vs = scr + kScrBufA;
  for (int e = tid; e < kV * 3; e += nt) {
    float acc = 0.f;
    for (int l = 0; l < 10; ++l) acc += f.betas[l] * m.shapedirs[e * 10 + l];
    vs[e] = m.v_template[e] + acc;
  }
  if (tid < kJ) {
    float* pose = scr + kScrPose;
    for (int c = 0; c < 3; ++c) {
      const int q = tid * 3 + c;
      pose[q] = f.full_pose[q] + (q >= 3 ? m.hands_mean[q - 3] : 0.f);
    }
    rodrigues_fwd(pose + 3 * tid, scr + kScrR + 9 * tid);
  }
}
// P1: joints
HOLD_HD void mano_p1(int tid, int nt, const ManoPtrs& m, const ManoFrame&, float* scr) {
  const float* vs = scr + kScrBufA;
  for (int o = tid; o < kJ * 3; o += nt) {
    const int j = o / 3, c = o % 3;
    float acc = 0.f;
    for (int v = 0; v < kV; ++v) acc += m.J_regressor[j * kV + v] * vs[3 * v + c];
    scr[kScrJ + o] = acc;
  }
}
// P2: v_posed; kinematic chain G and relative transforms A (one thread)
HOLD_HD void mano_p2(int tid, int nt, const ManoPtrs& m, const ManoFrame&, float* scr) {
  const float* vs = scr + kScrBufA;
  const float* Rm = scr + kScrR;
  const float* J = scr + kScrJ;
  float* vp = scr + kScrBufB;
  for (int e = tid; e < kV * 3; e += nt) {
    float acc = 0.f;
    for (int q = 0; q < kNP; ++q) {
      const int j = 1 + q / 9, k = q % 9;
      acc += (Rm[j * 9 + k] - ((k % 4 == 0) ? 1.f : 0.f)) * m.posedirs[q * (kV * 3) + e];
    }
    vp[e] = acc + vs[e];
  }
  if (tid == 0) {
    float* G = scr + kScrG;
    float* A = scr + kScrA;
    for (int i = 0; i < kJ; ++i) {
      float Tm[16];
      const int par = m.parents[i];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) Tm[4 * r + c] = Rm[i * 9 + 3 * r + c];
        Tm[4 * r + 3] = J[3 * i + r] - (i > 0 ? J[3 * par + r] : 0.f);
      }
      Tm[12] = Tm[13] = Tm[14] = 0.f, Tm[15] = 1.f;
      if (i == 0) {
        for (int e = 0; e < 16; ++e) G[e] = Tm[e];
      } else {
        for (int r = 0; r < 4; ++r)
          for (int c = 0; c < 4; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 4; ++k) acc += G[par * 16 + 4 * r + k] * Tm[4 * k + c];
            G[i * 16 + 4 * r + c] = acc;
          }
      }
    }
    for (int i = 0; i < kJ; ++i) {
      for (int e = 0; e < 16; ++e) A[i * 16 + e] = G[i * 16 + e];
      for (int r = 0; r < 4; ++r) {
        const float acc = G[i * 16 + 4 * r] * J[3 * i] + G[i * 16 + 4 * r + 1] * J[3 * i + 1] + G[i * 16 + 4 * r + 2] * J[3 * i + 2];
        A[i * 16 + 4 * r + 3] = G[i * 16 + 4 * r + 3] - acc;
      }
    }
  }
}
// P3: per vertex: skinned point o, g_o = s g_v (tips folded in), g_vp = T_v^T g_o; partial sums of g_t, g_s
HOLD_HD void mano_p3(int tid, int nt, const ManoPtrs& m, const ManoFrame& f, float* scr) {
  const float* A = scr + kScrA;
  const float* vp = scr + kScrBufB;
  float* go = scr + kScrBufA;   // v_shaped is dead from here on
  float* gvp = scr + kScrBufC;
  const float s = f.scene_scale[0], t0 = f.transl[0], t1 = f.transl[1], t2 = f.transl[2];
  float pt[3] = {0.f, 0.f, 0.f}, ps = 0.f;
  for (int v = tid; v < kV; v += nt) {
    float gv[3] = {0.f, 0.f, 0.f};
    if (f.g_verts != nullptr) { gv[0] = f.g_verts[3 * v], gv[1] = f.g_verts[3 * v + 1], gv[2] = f.g_verts[3 * v + 2]; }
    if (f.g_jnts != nullptr)
      for (int k = 0; k < 5; ++k)
        if (m.tips[k] == v) { gv[0] += f.g_jnts[3 * (kJ + k)], gv[1] += f.g_jnts[3 * (kJ + k) + 1], gv[2] += f.g_jnts[3 * (kJ + k) + 2]; }
    float T[12];
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int j = 0; j < kJ; ++j) {
      const float w = m.lbs_weights[v * kJ + j];
      for (int e = 0; e < 12; ++e) T[e] += w * A[j * 16 + e];
    }
    const float px = vp[3 * v], py = vp[3 * v + 1], pz = vp[3 * v + 2];
    float o[3];
    for (int r = 0; r < 3; ++r) o[r] = T[4 * r] * px + T[4 * r + 1] * py + T[4 * r + 2] * pz + T[4 * r + 3];
    for (int r = 0; r < 3; ++r) go[3 * v + r] = s * gv[r];
    for (int c = 0; c < 3; ++c) gvp[3 * v + c] = T[c] * (s * gv[0]) + T[4 + c] * (s * gv[1]) + T[8 + c] * (s * gv[2]);
    pt[0] += s * gv[0], pt[1] += s * gv[1], pt[2] += s * gv[2];
    ps += gv[0] * (o[0] + t0) + gv[1] * (o[1] + t1) + gv[2] * (o[2] + t2);
  }
  float* red = scr + kScrRed + 4 * tid;
  red[0] = pt[0], red[1] = pt[1], red[2] = pt[2], red[3] = ps;
}
// P4: d/dA (192 outputs, fixed-order sums over the vertices) + the tfs term; thread 0 of the tail reduces g_t, g_s
HOLD_HD void mano_p4(int tid, int nt, const ManoPtrs& m, const ManoFrame& f, float* scr) {
  const float* go = scr + kScrBufA;
  const float* vp = scr + kScrBufB;
  const float* A = scr + kScrA;
  const float s = f.scene_scale[0];
  for (int o = tid; o < kJ * 12; o += nt) {
    const int j = o / 12, r = (o % 12) / 4, c = o % 4;
    float acc = 0.f;
    for (int v = 0; v < kV; ++v) acc += m.lbs_weights[v * kJ + j] * go[3 * v + r] * (c < 3 ? vp[3 * v + c] : 1.f);
    if (f.g_tfs != nullptr) {
      // tfs_j = As_j Ci_j  =>  d/dAs_j = g_tfs_j Ci_j^T;  As = [s A_rows, + s t]
      float gAs = 0.f;
      if (f.tfs_c_inv != nullptr) {
        for (int k = 0; k < 4; ++k) gAs += f.g_tfs[j * 16 + 4 * r + k] * f.tfs_c_inv[j * 16 + 4 * c + k];
      } else {
        gAs = f.g_tfs[j * 16 + 4 * r + c];
      }
      acc += s * gAs;
    }
    scr[kScrgA + o] = acc;
  }
  if (tid == nt - 1) {
    float gt[3] = {0.f, 0.f, 0.f}, gs = 0.f;
    for (int k = 0; k < nt; ++k) {
      const float* red = scr + kScrRed + 4 * k;
      gt[0] += red[0], gt[1] += red[1], gt[2] += red[2], gs += red[3];
    }
    const float t[3] = {f.transl[0], f.transl[1], f.transl[2]};
    const float* G = scr + kScrG;
    if (f.g_jnts != nullptr) {
      for (int i = 0; i < kJ; ++i)
        for (int r = 0; r < 3; ++r) {
          const float gj = f.g_jnts[3 * i + r];
          gt[r] += s * gj;
          gs += gj * (G[i * 16 + 4 * r + 3] + t[r]);
        }
    }
    if (f.g_tfs != nullptr) {
      for (int j = 0; j < kJ; ++j)
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 4; ++c) {
            float gAs = 0.f;
            if (f.tfs_c_inv != nullptr) {
              for (int k = 0; k < 4; ++k) gAs += f.g_tfs[j * 16 + 4 * r + k] * f.tfs_c_inv[j * 16 + 4 * c + k];
            } else {
              gAs = f.g_tfs[j * 16 + 4 * r + c];
            }
            gs += gAs * A[j * 16 + 4 * r + c];
            if (c == 3) { gs += gAs * t[r]; gt[r] += s * gAs; }
          }
    }
    f.g_transl[0] = gt[0], f.g_transl[1] = gt[1], f.g_transl[2] = gt[2];
    f.g_scale[0] = gs;
  }
}
// P5: thread 0: A -> (G, J), chain backward -> d/dR, d/dJ.  Threads 1..: d/d pose feature (135 dots over g_vp)
HOLD_HD void mano_p5(int tid, int nt, const ManoPtrs& m, const ManoFrame& f, float* scr) {
  if (tid == 0) {
    const float* G = scr + kScrG;
    const float* J = scr + kScrJ;
    const float* Rm = scr + kScrR;
    const float* gA = scr + kScrgA;
    float* gR = scr + kScrgR;
    float* gJ = scr + kScrgJ;
    float gGR[kJ * 9], gGt[kJ * 3];
    const float s = f.scene_scale[0];
    for (int i = 0; i < kJ; ++i) {
      for (int r = 0; r < 3; ++r) {
        const float gat = gA[i * 12 + 4 * r + 3];
        for (int c = 0; c < 3; ++c) gGR[i * 9 + 3 * r + c] = gA[i * 12 + 4 * r + c] - gat * J[3 * i + c];
        gGt[i * 3 + r] = gat + ((f.g_jnts != nullptr) ? s * f.g_jnts[3 * i + r] : 0.f);
      }
      for (int c = 0; c < 3; ++c) {
        float acc = 0.f;
        for (int r = 0; r < 3; ++r) acc += G[i * 16 + 4 * r + c] * gA[i * 12 + 4 * r + 3];
        gJ[3 * i + c] = -acc;
      }
      for (int e = 0; e < 9; ++e) gR[i * 9 + e] = 0.f;
    }
    for (int i = kJ - 1; i >= 0; --i) {
      float gTR[9], gTt[3];
      if (i == 0) {
        for (int e = 0; e < 9; ++e) gTR[e] = gGR[e];
        for (int r = 0; r < 3; ++r) gTt[r] = gGt[r];
      } else {
        const int p = m.parents[i];
        const float* Gp = G + p * 16;
        float Tt[3];
        for (int r = 0; r < 3; ++r) Tt[r] = J[3 * i + r] - J[3 * p + r];
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 3; ++k) acc += Gp[4 * k + r] * gGR[i * 9 + 3 * k + c];   // Gp^R^T gGR_i
            gTR[3 * r + c] = acc;
          }
          float acc = 0.f;
          for (int k = 0; k < 3; ++k) acc += Gp[4 * k + r] * gGt[i * 3 + k];
          gTt[r] = acc;
        }
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) {
            float acc = 0.f;
            for (int k = 0; k < 3; ++k) acc += gGR[i * 9 + 3 * r + k] * Rm[i * 9 + 3 * c + k];  // gGR_i R_i^T
            gGR[p * 9 + 3 * r + c] += acc + gGt[i * 3 + r] * Tt[c];
          }
          gGt[p * 3 + r] += gGt[i * 3 + r];
          gJ[3 * p + r] -= gTt[r];
        }
      }
      for (int e = 0; e < 9; ++e) gR[i * 9 + e] += gTR[e];
      for (int r = 0; r < 3; ++r) gJ[3 * i + r] += gTt[r];
    }
  } else {
    const float* gvp = scr + kScrBufC;
    for (int q = tid - 1; q < kNP; q += nt - 1) {
      float acc = 0.f;
      const float* row = m.posedirs + (size_t)q * (kV * 3);
      for (int e = 0; e < kV * 3; ++e) acc += row[e] * gvp[e];
      scr[kScrgPf + q] = acc;
    }
  }
}
// P6: d/dR += pose-feature term; g_vs = g_vp + J_regressor^T g_J (in place in buffer C)
HOLD_HD void mano_p6(int tid, int nt, const ManoPtrs& m, const ManoFrame&, float* scr) {
  for (int q = tid; q < kNP; q += nt) scr[kScrgR + 9 + q] += scr[kScrgPf + q];
  const float* gJ = scr + kScrgJ;
  float* gvs = scr + kScrBufC;
  for (int e = tid; e < kV * 3; e += nt) {
    const int v = e / 3, c = e % 3;
    float acc = 0.f;
    for (int j = 0; j < kJ; ++j) acc += m.J_regressor[j * kV + v] * gJ[3 * j + c];
    gvs[e] += acc;
  }
}
// P7: outputs
HOLD_HD void mano_p7(int tid, int nt, const ManoPtrs& m, const ManoFrame& f, float* scr) {
  const float* gvs = scr + kScrBufC;
  for (int l = tid; l < 10; l += nt) {
    float acc = 0.f;
    for (int e = 0; e < kV * 3; ++e) acc += m.shapedirs[e * 10 + l] * gvs[e];
    f.g_betas[l] = acc;
  }
  for (int j = tid; j < kJ; j += nt) rodrigues_bwd(scr + kScrPose + 3 * j, scr + kScrgR + 9 * j, f.g_pose + 3 * j);
}

// ---------------------------------------------------------------------------------------------- object phases
struct ObjFrame {
  const float *rot, *trans, *scene_scale;  // [3], [3], [1]
  float obj_scale;
  const float* denorm;                     // [16]
  const float* pts;                        // [Nv,3]
  int Nv;
  const float *g_verts, *g_tfs;            // [Nv,3], [16]; each may be NULL
  float *g_rot, *g_trans, *g_scene_scale, *g_obj_scale;  // [3], [3], [1], [1]
};
constexpr int kObjR = 0, kObjT = 9, kObjRed = 32;  // scratch: R [9], T [16], then [nt][16] partial sums of d/dT
HOLD_HD int obj_scratch_floats(int nt) { return kObjRed + 16 * nt + 16; }

HOLD_HD void axis_angle_quat(const float* a, float& ang, float& k, float* q) {
  ang = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  const float half = ang * 0.5f;
  k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.f) : (sinf(half) / ang);
  q[0] = cosf(half), q[1] = a[0] * k, q[2] = a[1] * k, q[3] = a[2] * k;
}

HOLD_HD void obj_p0(int tid, int, const ObjFrame& f, float* scr) {
  if (tid != 0) return;
  float ang, k, q[4];
  axis_angle_quat(f.rot, ang, k, q);
  const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
  const float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
  const float Rm[9] = {1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                       two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                       two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)};
  const float s = f.scene_scale[0];
  float M[16];
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) M[4 * r + c] = s * Rm[3 * r + c] * f.obj_scale;
    M[4 * r + 3] = s * f.trans[r];
  }
  M[12] = M[13] = M[14] = 0.f, M[15] = 1.f;
  for (int e = 0; e < 9; ++e) scr[kObjR + e] = Rm[e];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float acc = 0.f;
      for (int qq = 0; qq < 4; ++qq) acc += M[4 * r + qq] * f.denorm[4 * qq + c];
      scr[kObjT + 4 * r + c] = acc;
    }
}
HOLD_HD void obj_p1(int tid, int nt, const ObjFrame& f, float* scr) {
  const float* T = scr + kObjT;
  float gT[16];
  for (int e = 0; e < 16; ++e) gT[e] = 0.f;
  if (f.g_verts != nullptr) {
    for (int v = tid; v < f.Nv; v += nt) {
      const float x[4] = {f.pts[3 * v], f.pts[3 * v + 1], f.pts[3 * v + 2], 1.f};
      float o[4];
      for (int r = 0; r < 4; ++r) o[r] = T[4 * r] * x[0] + T[4 * r + 1] * x[1] + T[4 * r + 2] * x[2] + T[4 * r + 3];
      const float g0 = f.g_verts[3 * v], g1 = f.g_verts[3 * v + 1], g2 = f.g_verts[3 * v + 2];
      const float go[4] = {g0 / o[3], g1 / o[3], g2 / o[3], -(g0 * o[0] + g1 * o[1] + g2 * o[2]) / (o[3] * o[3])};
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) gT[4 * r + c] += go[r] * x[c];
    }
  }
  for (int e = 0; e < 16; ++e) scr[kObjRed + 16 * tid + e] = gT[e];
}
HOLD_HD void obj_p2(int tid, int nt, const ObjFrame& f, float* scr) {
  if (tid >= 16) return;
  float acc = (f.g_tfs != nullptr) ? f.g_tfs[tid] : 0.f;
  for (int k = 0; k < nt; ++k) acc += scr[kObjRed + 16 * k + tid];
  scr[kObjRed + 16 * nt + tid] = acc;
}
HOLD_HD void obj_p3(int tid, int nt, const ObjFrame& f, float* scr) {
  if (tid != 0) return;
  const float* gT = scr + kObjRed + 16 * nt;
  const float* Rm = scr + kObjR;
  float gM[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      float acc = 0.f;
      for (int k = 0; k < 4; ++k) acc += gT[4 * r + k] * f.denorm[4 * c + k];   // gT D^T
      gM[4 * r + c] = acc;
    }
  const float s = f.scene_scale[0], os = f.obj_scale;
  float gR[9], rg = 0.f, tg = 0.f;
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) {
      gR[3 * r + c] = s * os * gM[4 * r + c];
      rg += Rm[3 * r + c] * gM[4 * r + c];
    }
    tg += f.trans[r] * gM[4 * r + 3];
    f.g_trans[r] = s * gM[4 * r + 3];
  }
  f.g_scene_scale[0] = os * rg + tg;
  f.g_obj_scale[0] = s * rg;
  // axis-angle -> quaternion -> matrix, reverse
  float ang, k, q[4];
  axis_angle_quat(f.rot, ang, k, q);
  const float qr = q[0], qi = q[1], qj = q[2], qk = q[3];
  const float N = qr * qr + qi * qi + qj * qj + qk * qk, ts = 2.0f / N;
  const float* g = gR;
  const float g_ts = -g[0] * (qj * qj + qk * qk) + g[1] * (qi * qj - qk * qr) + g[2] * (qi * qk + qj * qr) + g[3] * (qi * qj + qk * qr) -
                     g[4] * (qi * qi + qk * qk) + g[5] * (qj * qk - qi * qr) + g[6] * (qi * qk - qj * qr) + g[7] * (qj * qk + qi * qr) -
                     g[8] * (qi * qi + qj * qj);
  float g_qr = ts * (-g[1] * qk + g[2] * qj + g[3] * qk - g[5] * qi - g[6] * qj + g[7] * qi);
  float g_qi = ts * (g[1] * qj + g[2] * qk + g[3] * qj - 2 * g[4] * qi - g[5] * qr + g[6] * qk + g[7] * qr - 2 * g[8] * qi);
  float g_qj = ts * (-2 * g[0] * qj + g[1] * qi + g[2] * qr + g[3] * qi + g[5] * qk - g[6] * qr + g[7] * qk - 2 * g[8] * qj);
  float g_qk = ts * (-2 * g[0] * qk - g[1] * qr + g[2] * qi + g[3] * qr - 2 * g[4] * qk + g[5] * qj + g[6] * qi + g[7] * qj);
  const float cc = -g_ts * ts * ts;
  g_qr += cc * qr, g_qi += cc * qi, g_qj += cc * qj, g_qk += cc * qk;
  const float g_k = g_qi * f.rot[0] + g_qj * f.rot[1] + g_qk * f.rot[2];
  const float half = 0.5f * ang;
  const bool small = fabsf(ang) < 1e-6f;
  const float dk = small ? (-ang / 24.f) : ((0.5f * cosf(half) * ang - sinf(half)) / (ang * ang));
  const float g_ang = -0.5f * sinf(half) * g_qr + dk * g_k;
  const float gq[3] = {g_qi, g_qj, g_qk};
  for (int c = 0; c < 3; ++c) f.g_rot[c] = k * gq[c] + ((ang > 0.f) ? g_ang * f.rot[c] / ang : 0.f);
}

}  // namespace posebwd
}  // namespace hold
