// Marching cubes on a dense value grid (hold_mc_*; SURVEY §8f rank 4: the last host step of the reference's generate_mesh,
// utils/meshing.py:51).  Host/device functions: the kernels (mc.cuh) run one of these per grid node / cell; tests/test_cpu_mc.py
// compiles them with g++ and compares with the numpy restatement (oracle/marching_cubes.py) bit for bit.
// Conventions (same as the restatement): node INSIDE when value < level; one vertex on every grid edge whose ends differ, at
// lower_end + (level - v0) / (v1 - v0) along the edge, index coordinates, float32; vertex order = exclusive scan over
// [node (C order)][axis], face order = exclusive scan of the per-cell triangle counts over the cells (C order), table order inside a
// cell; normals (right-hand rule) towards increasing values.  Case tables: mc_tables.h (generated, tools/gen_mc_tables.py).
#pragma once
#include <stdint.h>

#include "mc_tables.h"

#if defined(__CUDACC__)
#define HOLD_MHD __device__ __forceinline__   // (the tables are __device__ data under nvcc: device-only there, plain inline under g++)
#else
#define HOLD_MHD inline
#endif

namespace hold {
namespace mc {

struct Dims { int n0, n1, n2; };

HOLD_MHD int64_t node_index(const Dims& d, int i, int j, int k) { return ((int64_t)i * d.n1 + j) * d.n2 + k; }

// flags[3]: is there a vertex on the edge from this node along axis a
HOLD_MHD void node_flags(const float* vol, const Dims& d, int i, int j, int k, float level, int32_t* flags) {
  const int64_t n = node_index(d, i, j, k);
  const bool in0 = vol[n] < level;
  flags[0] = (i + 1 < d.n0) ? (int32_t)(in0 != (vol[n + (int64_t)d.n1 * d.n2] < level)) : 0;
  flags[1] = (j + 1 < d.n1) ? (int32_t)(in0 != (vol[n + d.n2] < level)) : 0;
  flags[2] = (k + 1 < d.n2) ? (int32_t)(in0 != (vol[n + 1] < level)) : 0;
}

HOLD_MHD int cell_case(const float* vol, const Dims& d, int i, int j, int k, float level) {
  int c = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int64_t n = node_index(d, i + (q & 1), j + ((q >> 1) & 1), k + ((q >> 2) & 1));
    c |= (vol[n] < level) ? (1 << q) : 0;
  }
  return c;
}

// position of the vertex on the edge from node (i, j, k) along `axis`
HOLD_MHD void edge_vertex(const float* vol, const Dims& d, int i, int j, int k, int axis, float level, float* xyz) {
  const int64_t n = node_index(d, i, j, k);
  const int64_t step = (axis == 0) ? (int64_t)d.n1 * d.n2 : ((axis == 1) ? d.n2 : 1);
  const float v0 = vol[n], v1 = vol[n + step];
  const float t = (level - v0) / (v1 - v0);
  xyz[0] = (float)i + ((axis == 0) ? t : 0.f);
  xyz[1] = (float)j + ((axis == 1) ? t : 0.f);
  xyz[2] = (float)k + ((axis == 2) ? t : 0.f);
}

// the triangles of cell (i, j, k): vertex ids from the exclusive scan of the node flags; returns the triangle count
HOLD_MHD int cell_faces(int c, const Dims& d, int i, int j, int k, const int64_t* vid, int32_t* faces) {
  const int nt = kMcNTri[c];
  for (int t = 0; t < nt; ++t)
    for (int q = 0; q < 3; ++q) {
      const int e = kMcTri[c][3 * t + q];
      const int64_t n = node_index(d, i + kMcEdge[e][0], j + kMcEdge[e][1], k + kMcEdge[e][2]);
      faces[3 * t + q] = (int32_t)vid[n * 3 + kMcEdge[e][3]];
    }
  return nt;
}

}  // namespace mc
}  // namespace hold
