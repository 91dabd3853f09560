// Host run of the per-point arithmetic of k_mesh_sdf (hold_b200/csrc/mesh_sdf_phases.h).
#include "../../hold_b200/csrc/mesh_sdf_phases.h"

using namespace hold::meshsdf;

extern "C" int mesh_sdf_host(int P, int V, int F, const float* points, const float* verts, const int* faces, float* sdf, int* face_idx) {
  (void)V;
  for (int p = 0; p < P; ++p) {
    PointAcc acc;
    acc_init(acc);
    for (int f = 0; f < F; ++f) {
      float tri[9];
      for (int e = 0; e < 3; ++e)
        for (int k = 0; k < 3; ++k) tri[3 * e + k] = verts[3 * faces[3 * f + e] + k];
      acc_face(acc, points + 3 * p, tri, f);
    }
    sdf[p] = acc_sdf(acc);
    if (face_idx) face_idx[p] = acc.best_f;
  }
  return 0;
}
