// Geometry kernels: camera rays (a1), sphere exit (a2), MANO LBS server (a16), object transform (a17),
// KNN-weighted inverse / forward skinning (a6), rigid warp (a7), normals from the skinning Jacobian (a10).
#pragma once
#include "common.cuh"
#include "knn_phases.h"

namespace hold {

// ------------------------------------------------------------------------------------------------ a1
// get_camera_params + lift (datasets/utils.py:230-282), pose-matrix branch.
__global__ void k_camera_rays(int B, int P, const float* __restrict__ uv, const float* __restrict__ pose,
                              const float* __restrict__ K, float* __restrict__ dirs, float* __restrict__ cam) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * P) return;
  int b = i / P;
  const float* Kb = K + b * 16;
  const float* Pb = pose + b * 16;
  float fx = Kb[0], fy = Kb[5], cx = Kb[2], cy = Kb[6], sk = Kb[1];
  float x = uv[2 * i], y = uv[2 * i + 1], z = 1.0f;
  float xl = (x - cx + cy * sk / fy - sk * y / fy) / fx * z;
  float yl = (y - cy) / fy * z;
  float w[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) w[r] = Pb[r * 4 + 0] * xl + Pb[r * 4 + 1] * yl + Pb[r * 4 + 2] * z + Pb[r * 4 + 3];
  float c0 = Pb[3], c1 = Pb[7], c2 = Pb[11];
  float dx = w[0] - c0, dy = w[1] - c1, dz = w[2] - c2;
  float n = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);  // F.normalize eps
  dirs[3 * i] = dx / n, dirs[3 * i + 1] = dy / n, dirs[3 * i + 2] = dz / n;
  cam[3 * i] = c0, cam[3 * i + 1] = c1, cam[3 * i + 2] = c2;
}

// ------------------------------------------------------------------------------------------------ a6
// Exact K-nearest (K=15) of 778 vertices held in shared memory; ascending squared distance, ties -> lower index.
// Distance arithmetic is (dx*dx + dy*dy) + dz*dz without FMA contraction so that index selection is
// reproducible against the oracle's torch expression ((p - v)**2).sum(-1).
struct Knn15 {
  float d[kKnn];
  int i[kKnn];
};

__device__ __forceinline__ void knn15(const float* __restrict__ sv /*smem [778*3]*/, float px, float py, float pz,
                                      Knn15& r) {
#pragma unroll
  for (int k = 0; k < kKnn; ++k) { r.d[k] = 3.0e38f; r.i[k] = 0; }
  for (int v = 0; v < kVerts; ++v) {
    float dx = px - sv[3 * v], dy = py - sv[3 * v + 1], dz = pz - sv[3 * v + 2];
    float dist = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    if (dist < r.d[kKnn - 1]) {
      r.d[kKnn - 1] = dist;
      r.i[kKnn - 1] = v;
#pragma unroll
      for (int k = kKnn - 1; k > 0; --k) {
        if (r.d[k] < r.d[k - 1]) {
          float td = r.d[k]; r.d[k] = r.d[k - 1]; r.d[k - 1] = td;
          int ti = r.i[k]; r.i[k] = r.i[k - 1]; r.i[k - 1] = ti;
        }
      }
    }
  }
}

// query_skinning_weights_multi (model/mano/deformer.py:84-105) + blend of the 16 bone transforms
// (`einsum("bpn,bnij->bpij")`, deformer.py:165): returns the top three rows of T = sum_j w_j tfs_j and
// s = sum_j w_j tfs_j[3][3] (== sum of weights).
__device__ __forceinline__ void blend_tf(const Knn15& nn, const float* __restrict__ skin_w /*[778,16] global*/,
                                         const float* __restrict__ stf /*smem [16*16]*/, float T[12], float& s,
                                         float& dmin) {
  float conf[kKnn], csum = 0.f;
#pragma unroll
  for (int k = 0; k < kKnn; ++k) {
    conf[k] = expf(-fminf(nn.d[k], 4.0f));
    csum += conf[k];
  }
  dmin = sqrtf(fminf(nn.d[0], 4.0f));
  float w[kJoints];
#pragma unroll
  for (int j = 0; j < kJoints; ++j) w[j] = 0.f;
#pragma unroll
  for (int k = 0; k < kKnn; ++k) {
    float c = conf[k] / csum;
    const float4* row = reinterpret_cast<const float4*>(skin_w + nn.i[k] * kJoints);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = __ldg(row + q);
      w[4 * q + 0] += v.x * c; w[4 * q + 1] += v.y * c; w[4 * q + 2] += v.z * c; w[4 * q + 3] += v.w * c;
    }
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) T[e] = 0.f;
  s = 0.f;
#pragma unroll
  for (int j = 0; j < kJoints; ++j) {
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] += w[j] * stf[j * 16 + e];
    s += w[j] * stf[j * 16 + 15];
  }
}

__device__ __forceinline__ bool inv3(const float* A /*row-major 3x3 with row stride `rs`*/, int rs, float Ai[9]) {
  float a = A[0], b = A[1], c = A[2], d = A[rs], e = A[rs + 1], f = A[rs + 2], g = A[2 * rs], h = A[2 * rs + 1],
        i = A[2 * rs + 2];
  float c0 = e * i - f * h, c1 = f * g - d * i, c2 = d * h - e * g;
  float det = a * c0 + b * c1 + c * c2;
  float id = 1.0f / det;
  Ai[0] = c0 * id; Ai[1] = (c * h - b * i) * id; Ai[2] = (b * f - c * e) * id;
  Ai[3] = c1 * id; Ai[4] = (a * i - c * g) * id; Ai[5] = (c * d - a * f) * id;
  Ai[6] = c2 * id; Ai[7] = (b * g - a * h) * id; Ai[8] = (a * e - b * d) * id;
  return isfinite(id);
}

// Points along rays -> canonical points.  One block works on one frame (blockIdx.y) so that the frame's
// posed vertices and bone transforms sit in shared memory.
//   MODE_Z: x = cam + z * dir with z from a [R, zstride] buffer, ns samples per ray (sampler rounds, shading)
//   MODE_X: x given directly ([B, P, 3])
// hand:   x_c = (T^-1 [x;1])[:3]  with T the KNN-weighted blend (skinning(inverse=True), deformer.py:162-166)
// object: x_c = (tfs^-1 [x;1])[:3] (obj/deformer.py:21-31)
// FWD (forward_skinning, deformer.py:70-82 / obj/deformer.py:40-46): the input is a CANONICAL point, the KNN runs against the
// canonical vertices (`verts`, one set for all frames: vert_stride 0) and the blend is applied as is: x_d = (T [x_c;1])[:3].
template <bool HAND, bool FROM_Z, bool FWD = false>
__global__ void __launch_bounds__(128)
k_inverse_warp(int pts_per_frame, int ns, int zstride, const float* __restrict__ zbuf,
               const float* __restrict__ cam, const float* __restrict__ dirs, const float* __restrict__ xin,
               const float* __restrict__ tfs, const float* __restrict__ verts, const float* __restrict__ skin_w,
               float* __restrict__ xc, int* __restrict__ knn_idx, uint8_t* __restrict__ outlier,
               const SamplerState* __restrict__ st, int* __restrict__ err) {
  if (st != nullptr && st->done) return;  // sampler already converged: later rounds are no-ops
  __shared__ float sv[HAND ? kVerts * 3 : 4];
  __shared__ float stf[HAND ? kJoints * 16 : 16];
  __shared__ float sinv[12];
  const int b = blockIdx.y;
  if (HAND) {
    for (int t = threadIdx.x; t < kVerts * 3; t += blockDim.x) sv[t] = verts[(FWD ? (size_t)0 : (size_t)b * kVerts * 3) + t];
    for (int t = threadIdx.x; t < kJoints * 16; t += blockDim.x) stf[t] = tfs[(size_t)b * kJoints * 16 + t];
  } else {
    if (threadIdx.x < 16) stf[threadIdx.x] = tfs[b * 16 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0 && !FWD) {
      float Ai[9];
      bool ok = inv3(stf, 4, Ai);
      float s = stf[15];
      for (int r = 0; r < 3; ++r) {
        sinv[4 * r] = Ai[3 * r], sinv[4 * r + 1] = Ai[3 * r + 1], sinv[4 * r + 2] = Ai[3 * r + 2];
        sinv[4 * r + 3] = -(Ai[3 * r] * stf[3] + Ai[3 * r + 1] * stf[7] + Ai[3 * r + 2] * stf[11]) / s;
      }
      if (!ok) atomicOr(err, kErrNonFinite);
    }
  }
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pts_per_frame) return;
  size_t gp = (size_t)b * pts_per_frame + p;
  float x, y, z;
  if (FROM_Z) {
    size_t ray = gp / ns;
    int k = (int)(gp - ray * ns);
    float t = zbuf[ray * zstride + k];
    // mul then add, no FMA contraction: the reference forms points = cam + z * dir as two torch ops, and the
    // KNN selection downstream is sensitive to the last bit of the point
    x = __fadd_rn(cam[3 * ray], __fmul_rn(t, dirs[3 * ray]));
    y = __fadd_rn(cam[3 * ray + 1], __fmul_rn(t, dirs[3 * ray + 1]));
    z = __fadd_rn(cam[3 * ray + 2], __fmul_rn(t, dirs[3 * ray + 2]));
  } else {
    x = xin[3 * gp], y = xin[3 * gp + 1], z = xin[3 * gp + 2];
  }
  float ox, oy, oz;
  if (HAND) {
    Knn15 nn;
    knn15(sv, x, y, z, nn);
    float T[12], s, dmin;
    blend_tf(nn, skin_w, stf, T, s, dmin);
    if (FWD) {
      ox = (T[0] * x + T[1] * y + T[2] * z + T[3]) / s;
      oy = (T[4] * x + T[5] * y + T[6] * z + T[7]) / s;
      oz = (T[8] * x + T[9] * y + T[10] * z + T[11]) / s;
    } else {
      float Ai[9];
      inv3(T, 4, Ai);
      float rx = x - T[3] / s, ry = y - T[7] / s, rz = z - T[11] / s;
      ox = Ai[0] * rx + Ai[1] * ry + Ai[2] * rz;
      oy = Ai[3] * rx + Ai[4] * ry + Ai[5] * rz;
      oz = Ai[6] * rx + Ai[7] * ry + Ai[8] * rz;
    }
    if (knn_idx != nullptr) {
#pragma unroll
      for (int k = 0; k < kKnn; ++k) knn_idx[gp * kKnn + k] = nn.i[k];
    }
    if (outlier != nullptr) outlier[gp] = dmin > 0.1f;
  } else if (FWD) {
    const float w = stf[12] * x + stf[13] * y + stf[14] * z + stf[15];
    ox = (stf[0] * x + stf[1] * y + stf[2] * z + stf[3]) / w;
    oy = (stf[4] * x + stf[5] * y + stf[6] * z + stf[7]) / w;
    oz = (stf[8] * x + stf[9] * y + stf[10] * z + stf[11]) / w;
  } else {
    ox = sinv[0] * x + sinv[1] * y + sinv[2] * z + sinv[3];
    oy = sinv[4] * x + sinv[5] * y + sinv[6] * z + sinv[7];
    oz = sinv[8] * x + sinv[9] * y + sinv[10] * z + sinv[11];
  }
  xc[3 * gp] = ox, xc[3 * gp + 1] = oy, xc[3 * gp + 2] = oz;
}

// Hand-node variant of k_inverse_warp<true, true> for the hot path (sampler rounds, shading): a thread walks `kSeg` consecutive
// samples of one ray and finds each sample's 15 nearest posed vertices with the seeded, cluster-pruned exact search of
// knn_phases.h (same neighbours, same order as the full scan of knn15).  A warp holds 32 CONSECUTIVE rays at the same depth
// segment, so its lanes prune nearly the same vertex groups.  Same arithmetic downstream, same results.
constexpr int kSeg = 32;   // samples per thread at frame-sized launches; small launches (a training batch) use shorter walks
__global__ void __launch_bounds__(128)
k_inverse_warp_hand_rays(int rays_per_frame, int ns, int seg_len, int zstride, const float* __restrict__ zbuf,
                         const float* __restrict__ cam, const float* __restrict__ dirs, const float* __restrict__ tfs,
                         const float* __restrict__ verts, const float* __restrict__ skin_w, const unsigned short* __restrict__ perm,
                         float* __restrict__ xc, const SamplerState* __restrict__ st) {
  if (st != nullptr && st->done) return;
  __shared__ float sv[kVerts * 3];                               // original order: seeds are original indices
  __shared__ knnc::V4 svc[knnc::kNCl * knnc::kClSize];           // cluster order
  __shared__ knnc::Cl scl[knnc::kNCl];
  __shared__ float stf[kJoints * 16];
  __shared__ unsigned short scand[128 * knnc::kCand];
  unsigned short* cand = scand + threadIdx.x * knnc::kCand;
  const int b = blockIdx.y;
  for (int t = threadIdx.x; t < kVerts * 3; t += blockDim.x) sv[t] = verts[(size_t)b * kVerts * 3 + t];
  for (int t = threadIdx.x; t < kJoints * 16; t += blockDim.x) stf[t] = tfs[(size_t)b * kJoints * 16 + t];
  for (int j = threadIdx.x; j < knnc::kNCl * knnc::kClSize; j += blockDim.x) {
    const int v = perm[j];
    knnc::V4 e;
    if (v == 0xFFFF) { e.x = e.y = e.z = 1.0e18f; e.idx = -1; }
    else { const float* pv = verts + ((size_t)b * kVerts + v) * 3; e.x = pv[0], e.y = pv[1], e.z = pv[2]; e.idx = v; }
    svc[j] = e;
  }
  __syncthreads();
  if (threadIdx.x < knnc::kNCl) scl[threadIdx.x] = knnc::make_cluster(svc + threadIdx.x * knnc::kClSize);
  __syncthreads();
  const int segs = (ns + seg_len - 1) / seg_len;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rays_per_frame * segs) return;
  const int seg = t / rays_per_frame, ray_in_frame = t - seg * rays_per_frame;
  const size_t ray = (size_t)b * rays_per_frame + ray_in_frame;
  const float cx = cam[3 * ray], cy = cam[3 * ray + 1], cz = cam[3 * ray + 2];
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
  knnc::Top top;
  const int k0 = seg * seg_len, k1 = min(ns, k0 + seg_len);
  for (int k = k0; k < k1; ++k) {
    const float tz = zbuf[ray * zstride + k];
    const float x = __fadd_rn(cx, __fmul_rn(tz, dx)), y = __fadd_rn(cy, __fmul_rn(tz, dy)), z = __fadd_rn(cz, __fmul_rn(tz, dz));
    if (k == k0) knnc::full_scan(svc, x, y, z, top);
    else knnc::seeded_clustered(sv, svc, scl, x, y, z, top, cand);
    Knn15 nn;
#pragma unroll
    for (int j = 0; j < kKnn; ++j) { nn.d[j] = top.d[j]; nn.i[j] = top.i[j]; }
    float T[12], s, dmin;
    blend_tf(nn, skin_w, stf, T, s, dmin);
    float Ai[9];
    inv3(T, 4, Ai);
    const float rx = x - T[3] / s, ry = y - T[7] / s, rz = z - T[11] / s;
    const size_t gp = ray * ns + k;
    xc[3 * gp] = Ai[0] * rx + Ai[1] * ry + Ai[2] * rz;
    xc[3 * gp + 1] = Ai[3] * rx + Ai[4] * ry + Ai[5] * rz;
    xc[3 * gp + 2] = Ai[6] * rx + Ai[7] * ry + Ai[8] * rz;
  }
}

// extract_features' normal (engine/volsdf_utils.py:66-102): J = d x_d / d x_c of forward skinning with detached
// weights = (sum_j w_j tfs_j)[:3,:3] (hand; KNN against the CANONICAL vertices, deformer.py:70-82) or tfs[:3,:3]
// (object); n = normalize(g . J^-1, eps 1e-6).  Also the Laplace density of the sample (engine/density.py:21-26).
template <bool HAND>
__global__ void __launch_bounds__(128)
k_normals_density(int pts_per_frame, const float* __restrict__ xc, const float* __restrict__ grad,
                  const float* __restrict__ sdf, const float* __restrict__ tfs, const float* __restrict__ cano_verts,
                  const float* __restrict__ skin_w, const float* __restrict__ beta_param, float beta_min,
                  float* __restrict__ normal, float* __restrict__ density) {
  __shared__ float sv[HAND ? kVerts * 3 : 4];
  __shared__ float stf[HAND ? kJoints * 16 : 16];
  __shared__ float sJi[9];
  const int b = blockIdx.y;
  if (HAND) {
    for (int t = threadIdx.x; t < kVerts * 3; t += blockDim.x) sv[t] = cano_verts[t];
    for (int t = threadIdx.x; t < kJoints * 16; t += blockDim.x) stf[t] = tfs[(size_t)b * kJoints * 16 + t];
  } else {
    if (threadIdx.x < 16) stf[threadIdx.x] = tfs[b * 16 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) inv3(stf, 4, sJi);
  }
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pts_per_frame) return;
  size_t gp = (size_t)b * pts_per_frame + p;
  float Ji[9];
  if (HAND) {
    Knn15 nn;
    knn15(sv, xc[3 * gp], xc[3 * gp + 1], xc[3 * gp + 2], nn);
    float T[12], s, dmin;
    blend_tf(nn, skin_w, stf, T, s, dmin);
    inv3(T, 4, Ji);
  } else {
#pragma unroll
    for (int e = 0; e < 9; ++e) Ji[e] = sJi[e];
  }
  float g0 = grad[3 * gp], g1 = grad[3 * gp + 1], g2 = grad[3 * gp + 2];
  float n0 = g0 * Ji[0] + g1 * Ji[3] + g2 * Ji[6];
  float n1 = g0 * Ji[1] + g1 * Ji[4] + g2 * Ji[7];
  float n2 = g0 * Ji[2] + g1 * Ji[5] + g2 * Ji[8];
  float nn_ = fmaxf(sqrtf(n0 * n0 + n1 * n1 + n2 * n2), 1e-6f);
  normal[3 * gp] = n0 / nn_, normal[3 * gp + 1] = n1 / nn_, normal[3 * gp + 2] = n2 / nn_;
  if (density != nullptr) {
    float beta = fabsf(beta_param[0]) + beta_min;
    density[gp] = laplace_density(sdf[gp], beta);
  }
}

// ------------------------------------------------------------------------------------------------ a16
// GenericServer.forward (model/mano/server.py:62-99) around lbs() (utils/external/lbs.py:139-251): one CTA per
// frame, everything on chip, 1 launch instead of ~60.
__device__ __forceinline__ void rodrigues(const float* rv, float* Rm /*3x3*/) {
  // batch_rodrigues, lbs.py:298-329
  float ax = rv[0] + 1e-8f, ay = rv[1] + 1e-8f, az = rv[2] + 1e-8f;
  float ang = sqrtf(ax * ax + ay * ay + az * az);
  float rx = rv[0] / ang, ry = rv[1] / ang, rz = rv[2] / ang;
  float c = cosf(ang), s = sinf(ang);
  float Km[9] = {0.f, -rz, ry, rz, 0.f, -rx, -ry, rx, 0.f};
  float KK[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) KK[3 * i + j] = Km[3 * i] * Km[j] + Km[3 * i + 1] * Km[3 + j] + Km[3 * i + 2] * Km[6 + j];
  for (int e = 0; e < 9; ++e) Rm[e] = ((e % 4 == 0) ? 1.f : 0.f) + s * Km[e] + (1.f - c) * KK[e];
}

struct ManoDev {
  const float *v_template, *shapedirs, *posedirs, *J_regressor, *lbs_weights, *hands_mean;
  int parents[kJoints];
  int tips[5];
};

__global__ void __launch_bounds__(256)
k_mano_lbs(ManoDev m, const float* __restrict__ betas, const float* __restrict__ full_pose,
           const float* __restrict__ transl, const float* __restrict__ scene_scale,
           const float* __restrict__ tfs_c_inv, float* __restrict__ verts, float* __restrict__ jnts,
           float* __restrict__ tfs, float* __restrict__ v_posed_out) {
  extern __shared__ float sm[];
  float* v_shaped = sm;                    // [778*3]
  float* v_posed = v_shaped + kVerts * 3;  // [778*3]
  float* J = v_posed + kVerts * 3;         // [16*3]
  float* Rm = J + kJoints * 3;             // [16*9]
  float* pf = Rm + kJoints * 9;            // [135]
  float* G = pf + 136;                     // [16*16] world chain
  float* A = G + kJoints * 16;             // [16*16] relative transforms
  float* sbeta = A + kJoints * 16;         // [10]
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  if (tid < 10) sbeta[tid] = betas[b * 10 + tid];
  __syncthreads();
  // v_shaped = v_template + shapedirs . betas  (blend_shapes, lbs.py:274-295)
  for (int e = tid; e < kVerts * 3; e += nt) {
    float acc = 0.f;
#pragma unroll
    for (int l = 0; l < 10; ++l) acc += sbeta[l] * m.shapedirs[e * 10 + l];
    v_shaped[e] = m.v_template[e] + acc;
  }
  if (tid < kJoints) {
    float pose3[3];
    for (int c = 0; c < 3; ++c) {
      int q = tid * 3 + c;
      pose3[c] = full_pose[b * 48 + q] + (q >= 3 ? m.hands_mean[q - 3] : 0.f);  // MANO.forward: full_pose += pose_mean
    }
    rodrigues(pose3, Rm + tid * 9);
  }
  __syncthreads();
  // J = J_regressor . v_shaped (vertices2joints, lbs.py:254-271): one warp per (joint, coord) pair
  {
    int warp = tid / 32, lane = tid % 32, nw = nt / 32;
    for (int o = warp; o < kJoints * 3; o += nw) {
      int j = o / 3, c = o % 3;
      float acc = 0.f;
      for (int v = lane; v < kVerts; v += 32) acc += m.J_regressor[j * kVerts + v] * v_shaped[3 * v + c];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (lane == 0) J[o] = acc;
    }
  }
  if (tid < 135) {
    int j = 1 + tid / 9, e = tid % 9;
    pf[tid] = Rm[j * 9 + e] - ((e % 4 == 0) ? 1.f : 0.f);
  }
  __syncthreads();
  // v_posed = v_shaped + pose_feature . posedirs (lbs.py:209-216)
  for (int e = tid; e < kVerts * 3; e += nt) {
    float acc = 0.f;
    for (int q = 0; q < 135; ++q) acc += pf[q] * m.posedirs[q * (kVerts * 3) + e];
    v_posed[e] = acc + v_shaped[e];
    v_posed_out[(size_t)b * kVerts * 3 + e] = v_posed[e];
  }
  // kinematic chain (batch_rigid_transform, lbs.py:345-399): serial over 16 joints, trivially small
  if (tid == 0) {
    for (int i = 0; i < kJoints; ++i) {
      float Tm[16];
      int par = m.parents[i];
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
    for (int i = 0; i < kJoints; ++i) {
      // rel = G - pad(G . [J;0])  -> only the last column changes
      for (int e = 0; e < 16; ++e) A[i * 16 + e] = G[i * 16 + e];
      for (int r = 0; r < 4; ++r) {
        float acc = G[i * 16 + 4 * r] * J[3 * i] + G[i * 16 + 4 * r + 1] * J[3 * i + 1] + G[i * 16 + 4 * r + 2] * J[3 * i + 2];
        A[i * 16 + 4 * r + 3] = G[i * 16 + 4 * r + 3] - acc;
      }
    }
  }
  __syncthreads();
  const float s = scene_scale[b];
  const float t0 = transl[3 * b], t1 = transl[3 * b + 1], t2 = transl[3 * b + 2];
  // skinning: verts = (sum_j W[v,j] A_j) [v_posed;1]  (lbs.py:228-240), then server scaling (server.py:84-88)
  for (int v = tid; v < kVerts; v += nt) {
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    for (int j = 0; j < kJoints; ++j) {
      float w = m.lbs_weights[v * kJoints + j];
#pragma unroll
      for (int e = 0; e < 12; ++e) T[e] += w * A[j * 16 + e];
    }
    float px = v_posed[3 * v], py = v_posed[3 * v + 1], pz = v_posed[3 * v + 2];
    float o[3];
    for (int r = 0; r < 3; ++r) o[r] = T[4 * r] * px + T[4 * r + 1] * py + T[4 * r + 2] * pz + T[4 * r + 3];
    float* vo = verts + ((size_t)b * kVerts + v) * 3;
    vo[0] = o[0] * s + t0 * s, vo[1] = o[1] * s + t1 * s, vo[2] = o[2] * s + t2 * s;
  }
  __syncthreads();
  // joints: 16 chain joints + 5 tip vertices (vertex_joint_selector), scaled
  if (tid < 21) {
    float* jo = jnts + ((size_t)b * 21 + tid) * 3;
    if (tid < kJoints) {
      jo[0] = G[tid * 16 + 3] * s + t0 * s, jo[1] = G[tid * 16 + 7] * s + t1 * s, jo[2] = G[tid * 16 + 11] * s + t2 * s;
    } else {
      const float* vs = verts + ((size_t)b * kVerts + m.tips[tid - kJoints]) * 3;
      jo[0] = vs[0], jo[1] = vs[1], jo[2] = vs[2];
    }
  }
  // tfs = scaled A (. tfs_c_inv)  (server.py:90-96)
  if (tid < kJoints) {
    float As[16];
    for (int e = 0; e < 16; ++e) As[e] = A[tid * 16 + e];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) As[4 * r + c] *= s;
    As[3] += t0 * s, As[7] += t1 * s, As[11] += t2 * s;
    float* o = tfs + ((size_t)b * kJoints + tid) * 16;
    if (tfs_c_inv == nullptr) {
      for (int e = 0; e < 16; ++e) o[e] = As[e];
    } else {
      const float* Ci = tfs_c_inv + tid * 16;
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
          float acc = 0.f;
          for (int k = 0; k < 4; ++k) acc += As[4 * r + k] * Ci[4 * k + c];
          o[4 * r + c] = acc;
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------ a17
// ObjectModel.forward (model/obj/object_model.py:29-70); axis_angle_to_matrix via quaternions (common/rot.py).
__global__ void k_object_tf(int B, int Nv, const float* __restrict__ rot, const float* __restrict__ trans,
                            const float* __restrict__ scene_scale, float obj_scale, const float* __restrict__ denorm,
                            const float* __restrict__ pts, float* __restrict__ tfs, float* __restrict__ verts) {
  __shared__ float T[16];
  int b = blockIdx.y;
  if (threadIdx.x == 0) {
    float ax = rot[3 * b], ay = rot[3 * b + 1], az = rot[3 * b + 2];
    float ang = sqrtf(ax * ax + ay * ay + az * az);
    float half = ang * 0.5f;
    float k = (fabsf(ang) < 1e-6f) ? (0.5f - ang * ang / 48.f) : (sinf(half) / ang);
    float qr = cosf(half), qi = ax * k, qj = ay * k, qk = az * k;
    float two_s = 2.0f / (qr * qr + qi * qi + qj * qj + qk * qk);
    float Rm[9] = {1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                   two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                   two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)};
    float s = scene_scale[b];
    float M[16];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) M[4 * r + c] = s * Rm[3 * r + c] * obj_scale;
      M[4 * r + 3] = s * trans[3 * b + r];
    }
    M[12] = M[13] = M[14] = 0.f, M[15] = 1.f;
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < 4; ++c) {
        float acc = 0.f;
        for (int q = 0; q < 4; ++q) acc += M[4 * r + q] * denorm[4 * q + c];
        T[4 * r + c] = acc;
      }
    for (int e = 0; e < 16; ++e) tfs[b * 16 + e] = T[e];
  }
  __syncthreads();
  if (verts == nullptr) return;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < Nv; v += gridDim.x * blockDim.x) {
    float x = pts[3 * v], y = pts[3 * v + 1], z = pts[3 * v + 2];
    float o[4];
    for (int r = 0; r < 4; ++r) o[r] = T[4 * r] * x + T[4 * r + 1] * y + T[4 * r + 2] * z + T[4 * r + 3];
    float* vo = verts + ((size_t)b * Nv + v) * 3;
    vo[0] = o[0] / o[3], vo[1] = o[1] / o[3], vo[2] = o[2] / o[3];
  }
}

}  // namespace hold
