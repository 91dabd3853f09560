// Host run of the cluster-pruned, seeded exact KNN (hold_b200/csrc/knn_phases.h) along rays, exactly as the kernel
// k_inverse_warp_hand_rays walks it; returns indices and distances for comparison with brute force.
#include <vector>

#include "../../hold_b200/csrc/knn_phases.h"

using namespace hold::knnc;

extern "C" int knn_cluster_host(int n_rays, int ns, const float* verts /*[778,3] posed*/, const float* cano /*[778,3]*/,
                                const float* skin_w /*[778,16]*/, const float* cam, const float* dirs, const float* z /*[n_rays, ns]*/,
                                int* idx /*[n_rays, ns, 15]*/, float* dist, int* n_fallback, double* mean_visited) {
  std::vector<unsigned short> perm(kNCl * kClSize);
  cluster_order(cano, skin_w, 16, perm.data());
  std::vector<V4> svc(kNCl * kClSize);
  for (int j = 0; j < kNCl * kClSize; ++j) {
    if (perm[j] == 0xFFFF) { svc[j].x = svc[j].y = svc[j].z = 1.0e18f; svc[j].idx = -1; }
    else { svc[j].x = verts[3 * perm[j]], svc[j].y = verts[3 * perm[j] + 1], svc[j].z = verts[3 * perm[j] + 2]; svc[j].idx = perm[j]; }
  }
  std::vector<Cl> cl(kNCl);
  for (int k = 0; k < kNCl; ++k) cl[k] = make_cluster(&svc[k * kClSize]);
  unsigned short cand[kCand];
  int fb = 0;
  long long vis = 0, nvis = 0;
  for (int r = 0; r < n_rays; ++r) {
    Top nn;
    for (int k = 0; k < ns; ++k) {
      const float t = z[r * ns + k];
      const float x = cam[3 * r] + t * dirs[3 * r], y = cam[3 * r + 1] + t * dirs[3 * r + 1], zz = cam[3 * r + 2] + t * dirs[3 * r + 2];
      if (k == 0) full_scan(svc.data(), x, y, zz, nn);
      else {
        const int v = seeded_clustered(verts, svc.data(), cl.data(), x, y, zz, nn, cand);
        if (v < 0) ++fb; else { vis += v; ++nvis; }
      }
      for (int j = 0; j < kK; ++j) { idx[(r * ns + k) * kK + j] = nn.i[j]; dist[(r * ns + k) * kK + j] = nn.d[j]; }
    }
  }
  *n_fallback = fb;
  *mean_visited = nvis ? (double)vis / (double)nvis : 0.0;
  return 0;
}
