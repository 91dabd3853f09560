// Host run of the filtered KNN (hold_b200/csrc/knn_phases.h) along rays: seeded from sample to sample exactly as the kernel
// does; returns indices and distances for comparison with brute force.
#include <vector>

#include "../../hold_b200/csrc/knn_phases.h"

using namespace hold::knnf;

extern "C" int knn_filter_host(int n_rays, int ns, const float* verts /*[778,3]*/, const float* cam, const float* dirs,
                               const float* z /*[n_rays, ns]*/, int* idx /*[n_rays, ns, 15]*/, float* dist, int* n_fallback) {
  std::vector<V4> sv(kNV);
  float qmax = 0.f;
  for (int v = 0; v < kNV; ++v) {
    sv[v].x = verts[3 * v], sv[v].y = verts[3 * v + 1], sv[v].z = verts[3 * v + 2];
    sv[v].q = sv[v].x * sv[v].x + sv[v].y * sv[v].y + sv[v].z * sv[v].z;
    qmax = fmaxf(qmax, sv[v].q);
  }
  unsigned short cand[kCand];
  int fb = 0;
  for (int r = 0; r < n_rays; ++r) {
    Top nn;
    for (int k = 0; k < ns; ++k) {
      const float t = z[r * ns + k];
      const float x = cam[3 * r] + t * dirs[3 * r], y = cam[3 * r + 1] + t * dirs[3 * r + 1], zz = cam[3 * r + 2] + t * dirs[3 * r + 2];
      if (k == 0) full_scan(sv.data(), x, y, zz, nn);
      else seeded_filter(sv.data(), qmax, x, y, zz, nn, cand);
      for (int j = 0; j < kK; ++j) { idx[(r * ns + k) * kK + j] = nn.i[j]; dist[(r * ns + k) * kK + j] = nn.d[j]; }
    }
  }
  *n_fallback = fb;
  return 0;
}
