// Marching cubes kernels (hold_mc_mark / hold_mc_emit): one thread per grid node, arithmetic in mc_phases.h.
#pragma once
#include "common.cuh"
#include "mc_phases.h"

namespace hold {

__global__ void k_mc_mark(mc::Dims d, const float* __restrict__ vol, float level, int32_t* __restrict__ flags, int32_t* __restrict__ ntri) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)d.n0 * d.n1 * d.n2;
  if (n >= total) return;
  const int k = (int)(n % d.n2), j = (int)((n / d.n2) % d.n1), i = (int)(n / ((int64_t)d.n1 * d.n2));
  mc::node_flags(vol, d, i, j, k, level, flags + n * 3);
  if (i + 1 < d.n0 && j + 1 < d.n1 && k + 1 < d.n2)
    ntri[((int64_t)i * (d.n1 - 1) + j) * (d.n2 - 1) + k] = kMcNTri[mc::cell_case(vol, d, i, j, k, level)];
}

__global__ void k_mc_emit(mc::Dims d, const float* __restrict__ vol, float level, const int32_t* __restrict__ flags,
                          const int64_t* __restrict__ vid, const int64_t* __restrict__ toff, float* __restrict__ verts, int32_t* __restrict__ faces) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)d.n0 * d.n1 * d.n2;
  if (n >= total) return;
  const int k = (int)(n % d.n2), j = (int)((n / d.n2) % d.n1), i = (int)(n / ((int64_t)d.n1 * d.n2));
#pragma unroll
  for (int a = 0; a < 3; ++a)
    if (flags[n * 3 + a]) mc::edge_vertex(vol, d, i, j, k, a, level, verts + 3 * vid[n * 3 + a]);
  if (i + 1 < d.n0 && j + 1 < d.n1 && k + 1 < d.n2) {
    const int64_t c = ((int64_t)i * (d.n1 - 1) + j) * (d.n2 - 1) + k;
    const int cs = mc::cell_case(vol, d, i, j, k, level);
    if (kMcNTri[cs]) mc::cell_faces(cs, d, i, j, k, vid, faces + 3 * toff[c]);
  }
}

}  // namespace hold
