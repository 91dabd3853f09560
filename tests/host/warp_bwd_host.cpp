// Host run of the inverse-warp backward phases (hold_b200/csrc/warp_bwd_phases.h), block by block, thread by thread.
#include <vector>

#include "../../hold_b200/csrc/warp_bwd_phases.h"

using namespace hold::warpbwd;

extern "C" int warp_bwd_hand_host(int nt, int P, const float* x, const int* knn_idx, const float* verts, const float* skin_w,
                                  const float* tfs, const float* g_xc, float* g_x, float* g_tfs) {
  const int nb = (P + nt - 1) / nt;
  std::vector<float> scr(scratch_floats(nt)), partials((size_t)nb * kJ * kG);
  for (int b = 0; b < nb; ++b) {
    for (int t = 0; t < nt; ++t) hand_p0(t, nt, b * nt + t, P, x, knn_idx, verts, skin_w, tfs, g_xc, g_x, scr.data());
    for (int t = 0; t < nt; ++t) hand_p1(t, nt, scr.data(), partials.data() + (size_t)b * kJ * kG);
  }
  for (int o = 0; o < 256; ++o) hand_final(o, nb, partials.data(), g_tfs);
  return 0;
}

extern "C" int warp_bwd_obj_host(int nt, int P, const float* x, const float* tf, const float* g_xc, float* g_x, float* g_tf) {
  const int nb = (P + nt - 1) / nt;
  std::vector<float> scr(nt * kG), partials((size_t)nb * kG);
  for (int b = 0; b < nb; ++b) {
    for (int t = 0; t < nt; ++t) obj_p0(t, nt, b * nt + t, P, x, tf, g_xc, g_x, scr.data());
    for (int t = 0; t < nt; ++t) obj_p1(t, nt, scr.data(), partials.data() + (size_t)b * kG);
  }
  for (int e = 0; e < 16; ++e) obj_final(e, nb, partials.data(), g_tf);
  return 0;
}
