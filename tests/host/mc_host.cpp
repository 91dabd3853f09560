// Host build of hold_b200/csrc/mc_phases.h for tests/test_cpu_mc.py: the loops the kernels of mc.cuh run one thread per node / cell.
#include <vector>

#include "../../hold_b200/csrc/mc_phases.h"

using namespace hold;

extern "C" {
// pass 1: flags [n0*n1*n2*3], ntri [(n0-1)(n1-1)(n2-1)]
void mc_mark(int n0, int n1, int n2, const float* vol, float level, int32_t* flags, int32_t* ntri) {
  mc::Dims d{n0, n1, n2};
  for (int i = 0; i < n0; ++i)
    for (int j = 0; j < n1; ++j)
      for (int k = 0; k < n2; ++k) {
        mc::node_flags(vol, d, i, j, k, level, flags + mc::node_index(d, i, j, k) * 3);
        if (i + 1 < n0 && j + 1 < n1 && k + 1 < n2)
          ntri[((int64_t)i * (n1 - 1) + j) * (n2 - 1) + k] = kMcNTri[mc::cell_case(vol, d, i, j, k, level)];
      }
}
// pass 2: vid / toff = exclusive scans of flags / ntri
void mc_emit(int n0, int n1, int n2, const float* vol, float level, const int32_t* flags, const int64_t* vid, const int64_t* toff, float* verts,
             int32_t* faces) {
  mc::Dims d{n0, n1, n2};
  for (int i = 0; i < n0; ++i)
    for (int j = 0; j < n1; ++j)
      for (int k = 0; k < n2; ++k) {
        const int64_t n = mc::node_index(d, i, j, k);
        for (int a = 0; a < 3; ++a)
          if (flags[n * 3 + a]) mc::edge_vertex(vol, d, i, j, k, a, level, verts + 3 * vid[n * 3 + a]);
        if (i + 1 < n0 && j + 1 < n1 && k + 1 < n2) {
          const int64_t c = ((int64_t)i * (n1 - 1) + j) * (n2 - 1) + k;
          mc::cell_faces(mc::cell_case(vol, d, i, j, k, level), d, i, j, k, vid, faces + 3 * toff[c]);
        }
      }
}
}
