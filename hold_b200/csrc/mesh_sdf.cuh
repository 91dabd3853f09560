// Kernels of the training-only loss targets; arithmetic in mesh_sdf_phases.h.
#pragma once
#include "common.cuh"
#include "mesh_sdf_phases.h"

namespace hold {

constexpr int kMeshTile = 256;   // faces staged in shared memory per pass (9 KB)

// One thread per point, faces streamed through shared memory in tiles; block = 128 points of one frame (blockIdx.y).
// verts: [Bv, V, 3] with Bv == B or 1 (broadcast); faces [F, 3].
__global__ void __launch_bounds__(128)
k_mesh_sdf(int P, int V, int F, int verts_bstride, const float* __restrict__ points, const float* __restrict__ verts,
           const int* __restrict__ faces, float* __restrict__ sdf, int* __restrict__ face_idx) {
  __shared__ float tri[kMeshTile * 9];
  const int b = blockIdx.y, p = blockIdx.x * blockDim.x + threadIdx.x;
  const float* vb = verts + (size_t)b * verts_bstride;
  float pt[3] = {0.f, 0.f, 0.f};
  if (p < P) { pt[0] = points[((size_t)b * P + p) * 3], pt[1] = points[((size_t)b * P + p) * 3 + 1], pt[2] = points[((size_t)b * P + p) * 3 + 2]; }
  meshsdf::PointAcc acc;
  meshsdf::acc_init(acc);
  for (int f0 = 0; f0 < F; f0 += kMeshTile) {
    const int nf = min(kMeshTile, F - f0);
    __syncthreads();
    for (int e = threadIdx.x; e < nf * 3; e += blockDim.x) {
      const int v = faces[(size_t)(f0 + e / 3) * 3 + e % 3];
      tri[e * 3] = vb[3 * v], tri[e * 3 + 1] = vb[3 * v + 1], tri[e * 3 + 2] = vb[3 * v + 2];
    }
    __syncthreads();
    if (p < P)
      for (int f = 0; f < nf; ++f) meshsdf::acc_face(acc, pt, tri + 9 * f, f0 + f);
  }
  if (p < P) {
    sdf[(size_t)b * P + p] = meshsdf::acc_sdf(acc);
    if (face_idx != nullptr) face_idx[(size_t)b * P + p] = acc.best_f;
  }
}

// check_off_in_surface_points_cano_mesh's reduction (volsdf_utils.py:209-217): per ray, the minimum signed distance over
// its samples -> off-surface (min > threshold) and in-surface (min <= 0) flags.
__global__ void k_off_in_surface(int R, int S, const float* __restrict__ sdf, float threshold, uint8_t* __restrict__ off,
                                 uint8_t* __restrict__ in) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float m = 3.0e38f;
  for (int k = 0; k < S; ++k) m = fminf(m, sdf[(size_t)r * S + k]);
  if (off != nullptr) off[r] = m > threshold;
  if (in != nullptr) in[r] = m <= 0.0f;
}

}  // namespace hold
