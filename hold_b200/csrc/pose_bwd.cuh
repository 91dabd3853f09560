// SURVEY §8f rank 4 (second half): reverse mode of the pose servers for optimize_ckpt.py (fitting/model.py:117).
// The arithmetic lives in pose_bwd_phases.h (shared with the host emulation in tests/host/); here are the two
// block-per-frame kernels that run the phases between block barriers.
#pragma once
#include "common.cuh"
#include "pose_bwd_phases.h"

namespace hold {

__global__ void __launch_bounds__(256)
k_mano_lbs_bwd(posebwd::ManoPtrs m, const float* __restrict__ betas, const float* __restrict__ full_pose,
               const float* __restrict__ transl, const float* __restrict__ scene_scale, const float* __restrict__ tfs_c_inv,
               const float* __restrict__ g_verts, const float* __restrict__ g_jnts, const float* __restrict__ g_tfs,
               float* __restrict__ g_betas, float* __restrict__ g_pose, float* __restrict__ g_transl,
               float* __restrict__ g_scale) {
  extern __shared__ float scr[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  posebwd::ManoFrame f;
  f.betas = betas + (size_t)b * 10, f.full_pose = full_pose + (size_t)b * 48, f.transl = transl + (size_t)b * 3;
  f.scene_scale = scene_scale + b, f.tfs_c_inv = tfs_c_inv;
  f.g_verts = g_verts ? g_verts + (size_t)b * kVerts * 3 : nullptr;
  f.g_jnts = g_jnts ? g_jnts + (size_t)b * 21 * 3 : nullptr;
  f.g_tfs = g_tfs ? g_tfs + (size_t)b * kJoints * 16 : nullptr;
  f.g_betas = g_betas + (size_t)b * 10, f.g_pose = g_pose + (size_t)b * 48, f.g_transl = g_transl + (size_t)b * 3;
  f.g_scale = g_scale + b;
  posebwd::mano_p0(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p1(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p2(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p3(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p4(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p5(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p6(tid, nt, m, f, scr);
  __syncthreads();
  posebwd::mano_p7(tid, nt, m, f, scr);
}

__global__ void __launch_bounds__(256)
k_object_tf_bwd(int Nv, const float* __restrict__ rot, const float* __restrict__ trans, const float* __restrict__ scene_scale,
                float obj_scale, const float* __restrict__ denorm, const float* __restrict__ pts,
                const float* __restrict__ g_verts, const float* __restrict__ g_tfs, float* __restrict__ g_rot,
                float* __restrict__ g_trans, float* __restrict__ g_scene_scale, float* __restrict__ g_obj_scale) {
  extern __shared__ float scr[];
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  posebwd::ObjFrame f;
  f.rot = rot + (size_t)b * 3, f.trans = trans + (size_t)b * 3, f.scene_scale = scene_scale + b, f.obj_scale = obj_scale;
  f.denorm = denorm, f.pts = pts, f.Nv = Nv;
  f.g_verts = g_verts ? g_verts + (size_t)b * Nv * 3 : nullptr;
  f.g_tfs = g_tfs ? g_tfs + (size_t)b * 16 : nullptr;
  f.g_rot = g_rot + (size_t)b * 3, f.g_trans = g_trans + (size_t)b * 3, f.g_scene_scale = g_scene_scale + b, f.g_obj_scale = g_obj_scale + b;
  posebwd::obj_p0(tid, nt, f, scr);
  __syncthreads();
  posebwd::obj_p1(tid, nt, f, scr);
  __syncthreads();
  posebwd::obj_p2(tid, nt, f, scr);
  __syncthreads();
  posebwd::obj_p3(tid, nt, f, scr);
}

}  // namespace hold
