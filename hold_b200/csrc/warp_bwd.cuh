// Kernels of the inverse-warp backward (arithmetic in warp_bwd_phases.h): block = 128 points of one frame (blockIdx.y), per-block
// partial sums to a workspace, a second kernel adds them in block order (deterministic).
#pragma once
#include "common.cuh"
#include "warp_bwd_phases.h"

namespace hold {

__global__ void __launch_bounds__(128)
k_inverse_warp_bwd_hand(int P, const float* __restrict__ x, const int* __restrict__ knn_idx, const float* __restrict__ verts,
                        const float* __restrict__ skin_w, const float* __restrict__ tfs, const float* __restrict__ g_xc,
                        float* __restrict__ g_x, float* __restrict__ partials) {
  __shared__ float stf[kJoints * 16];
  __shared__ float scr[128 * (warpbwd::kJ + warpbwd::kG)];
  const int b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  for (int t = tid; t < kJoints * 16; t += nt) stf[t] = tfs[(size_t)b * kJoints * 16 + t];
  __syncthreads();
  const size_t o = (size_t)b * P;
  warpbwd::hand_p0(tid, nt, blockIdx.x * nt + tid, P, x + o * 3, knn_idx + o * warpbwd::kK, verts + (size_t)b * kVerts * 3, skin_w, stf,
                   g_xc + o * 3, g_x ? g_x + o * 3 : nullptr, scr);
  __syncthreads();
  warpbwd::hand_p1(tid, nt, scr, partials + ((size_t)b * gridDim.x + blockIdx.x) * (warpbwd::kJ * warpbwd::kG));
}

__global__ void k_inverse_warp_bwd_hand_final(int n_blocks, const float* __restrict__ partials, float* __restrict__ g_tfs) {
  const int b = blockIdx.x, o = threadIdx.x;   // 256 threads
  warpbwd::hand_final(o, n_blocks, partials + (size_t)b * n_blocks * (warpbwd::kJ * warpbwd::kG), g_tfs + (size_t)b * 256);
}

__global__ void __launch_bounds__(128)
k_inverse_warp_bwd_obj(int P, const float* __restrict__ x, const float* __restrict__ tfs, const float* __restrict__ g_xc,
                       float* __restrict__ g_x, float* __restrict__ partials) {
  __shared__ float stf[16];
  __shared__ float scr[128 * warpbwd::kG];
  const int b = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
  if (tid < 16) stf[tid] = tfs[b * 16 + tid];
  __syncthreads();
  const size_t o = (size_t)b * P;
  warpbwd::obj_p0(tid, nt, blockIdx.x * nt + tid, P, x + o * 3, stf, g_xc + o * 3, g_x ? g_x + o * 3 : nullptr, scr);
  __syncthreads();
  warpbwd::obj_p1(tid, nt, scr, partials + ((size_t)b * gridDim.x + blockIdx.x) * warpbwd::kG);
}

__global__ void k_inverse_warp_bwd_obj_final(int n_blocks, const float* __restrict__ partials, float* __restrict__ g_tfs) {
  const int b = blockIdx.x, e = threadIdx.x;   // 16 threads
  warpbwd::obj_final(e, n_blocks, partials + (size_t)b * n_blocks * warpbwd::kG, g_tfs + (size_t)b * 16);
}

}  // namespace hold
