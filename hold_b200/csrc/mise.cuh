// Kernels of the GPU MISE (first half); semantics and per-item code in mise_phases.h.
#pragma once
#include "common.cuh"
#include "mise_phases.h"

namespace hold {

__global__ void k_mise_init(mise::Grid g) {
  const int n = g.res0 + 1, t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n * n) return;
  const int s0 = 1 << g.depth, i = t / (n * n), j = (t / n) % n, k = t % n;
  g.state[mise::pidx(g, i * s0, j * s0, k * s0)] = mise::kUnknown;   // corners of the coarse voxels (mise.pyx:76-88)
}

__global__ void k_mise_collect(mise::Grid g, int* __restrict__ queue, unsigned int* __restrict__ counter, unsigned int cap) {
  const size_t np = (size_t)g.G * g.G * g.G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x) {
    if (g.state[i] == mise::kUnknown) {
      const unsigned int n = atomicAdd(counter, 1u);
      if (n < cap) queue[n] = (int)i;
    }
  }
}

__global__ void k_mise_coords(mise::Grid g, const int* __restrict__ queue, int n, int* __restrict__ coords) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int i = queue[t];
  coords[3 * t] = i / (g.G * g.G), coords[3 * t + 1] = (i / g.G) % g.G, coords[3 * t + 2] = i % g.G;
}

__global__ void k_mise_scatter(mise::Grid g, const int* __restrict__ queue, int n, const float* __restrict__ values) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  g.val[queue[t]] = values[t];
  g.state[queue[t]] = mise::kKnown;
}

__global__ void k_mise_mark(mise::Grid g) {
  const size_t np = (size_t)g.G * g.G * g.G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x) {
    if (g.state[i] != mise::kKnown) continue;
    mise::mark_point(g, (int)(i / ((size_t)g.G * g.G)), (int)((i / g.G) % g.G), (int)(i % g.G));
  }
}

__global__ void k_mise_subdivide(mise::Grid g, int L) {
  const int n = g.res0 << L;
  const size_t nv = (size_t)n * n * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (size_t)gridDim.x * blockDim.x)
    mise::subdivide_voxel(g, L, (int)(i / ((size_t)n * n)), (int)((i / n) % n), (int)(i % n));
}

__global__ void k_mise_dense_init(mise::Grid g, float* __restrict__ out) {
  const size_t np = (size_t)g.G * g.G * g.G;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (g.state[i] == mise::kKnown) ? g.val[i] : __int_as_float(0x7fc00000);
}
// axis 0: thread = (j, k) line along x; 1: (i, k) along y; 2: (i, j) along z
__global__ void k_mise_fill(mise::Grid g, float* __restrict__ out, int axis) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.G * g.G) return;
  const int a = t / g.G, b = t % g.G;
  if (axis == 0) mise::fill_x(g, out, a, b);
  else if (axis == 1) mise::fill_y(g, out, a, b);
  else mise::fill_z(g, out, a, b);
}

}  // namespace hold

struct hold_mise {
  hold_ctx* ctx = nullptr;
  hold::mise::Grid g;
  int* queue = nullptr;
  unsigned int* counter = nullptr;
  size_t np = 0;
  int n_last = 0;   // points handed out by the last query, awaiting update
};
