// Host emulation of the pose-server backward kernels: the SAME phase functions the CUDA kernels run
// (hold_b200/csrc/pose_bwd_phases.h), executed tid by tid with a block barrier = end of the tid loop.
// Built by tests/test_cpu_pose_bwd.py with g++ into tests/_build/ (git-ignored) and called through ctypes.
#include <vector>

#include "../../hold_b200/csrc/pose_bwd_phases.h"

using namespace hold::posebwd;

extern "C" int pose_bwd_mano_host(int nt, int B, const float* v_template, const float* shapedirs, const float* posedirs,
                                  const float* J_regressor, const float* lbs_weights, const float* hands_mean,
                                  const int* parents, const int* tips, const float* betas, const float* full_pose,
                                  const float* transl, const float* scene_scale, const float* tfs_c_inv, const float* g_verts,
                                  const float* g_jnts, const float* g_tfs, float* g_betas, float* g_pose, float* g_transl,
                                  float* g_scale) {
  ManoPtrs m;
  m.v_template = v_template, m.shapedirs = shapedirs, m.posedirs = posedirs, m.J_regressor = J_regressor;
  m.lbs_weights = lbs_weights, m.hands_mean = hands_mean;
  for (int i = 0; i < kJ; ++i) m.parents[i] = (i == 0) ? 0 : parents[i];
  for (int i = 0; i < 5; ++i) m.tips[i] = tips[i];
  std::vector<float> scr(mano_scratch_floats(nt));
  for (int b = 0; b < B; ++b) {
    ManoFrame f;
    f.betas = betas + b * 10, f.full_pose = full_pose + b * 48, f.transl = transl + b * 3, f.scene_scale = scene_scale + b;
    f.tfs_c_inv = tfs_c_inv;
    f.g_verts = g_verts ? g_verts + b * kV * 3 : nullptr;
    f.g_jnts = g_jnts ? g_jnts + b * 21 * 3 : nullptr;
    f.g_tfs = g_tfs ? g_tfs + b * kJ * 16 : nullptr;
    f.g_betas = g_betas + b * 10, f.g_pose = g_pose + b * 48, f.g_transl = g_transl + b * 3, f.g_scale = g_scale + b;
    for (int t = 0; t < nt; ++t) mano_p0(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p1(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p2(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p3(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p4(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p5(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p6(t, nt, m, f, scr.data());
    for (int t = 0; t < nt; ++t) mano_p7(t, nt, m, f, scr.data());
  }
  return 0;
}

extern "C" int pose_bwd_object_host(int nt, int B, const float* rot, const float* trans, const float* scene_scale,
                                    float obj_scale, const float* denorm, const float* pts, int Nv, const float* g_verts,
                                    const float* g_tfs, float* g_rot, float* g_trans, float* g_scene_scale, float* g_obj_scale) {
  std::vector<float> scr(obj_scratch_floats(nt));
  for (int b = 0; b < B; ++b) {
    ObjFrame f;
    f.rot = rot + b * 3, f.trans = trans + b * 3, f.scene_scale = scene_scale + b, f.obj_scale = obj_scale, f.denorm = denorm;
    f.pts = pts, f.Nv = Nv;
    f.g_verts = g_verts ? g_verts + b * Nv * 3 : nullptr;
    f.g_tfs = g_tfs ? g_tfs + b * 16 : nullptr;
    f.g_rot = g_rot + b * 3, f.g_trans = g_trans + b * 3, f.g_scene_scale = g_scene_scale + b, f.g_obj_scale = g_obj_scale + b;
    for (int t = 0; t < nt; ++t) obj_p0(t, nt, f, scr.data());
    for (int t = 0; t < nt; ++t) obj_p1(t, nt, f, scr.data());
    for (int t = 0; t < nt; ++t) obj_p2(t, nt, f, scr.data());
    for (int t = 0; t < nt; ++t) obj_p3(t, nt, f, scr.data());
  }
  return 0;
}
