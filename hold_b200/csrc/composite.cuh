// merge_factors (hold/hold_utils.py:76-121) + density2weight (engine/volsdf_utils.py:220-251) + the five
// integrals of volumetric_render (hold/hold_utils.py:243-271).  The per-node z lists are already sorted, so
// the reference's torch.sort over the concatenation is a stable n-way merge (ties -> lower node first).
#pragma once
#include "common.cuh"

namespace hold {

struct CompositeArgs {
  int n, R, S;
  const float* color[HOLD_MAX_NODES];
  const float* normal[HOLD_MAX_NODES];
  const float* density[HOLD_MAX_NODES];
  const float* z[HOLD_MAX_NODES];
  int class_id[HOLD_MAX_NODES];
  hold_render_out out;
  int drop_head, drop_tail;  // (n-1, n) for the composite; (0, 0) for a single node
  int single_zmax_last;      // per-node render: z_max = z[:, -1] (hold_net.py:79-80)
};

__global__ void __launch_bounds__(128) k_composite(CompositeArgs a) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= a.R) return;
  const int n = a.n, S = a.S, total = n * S;
  int head[HOLD_MAX_NODES];
  const float* zr[HOLD_MAX_NODES];
#pragma unroll
  for (int k = 0; k < HOLD_MAX_NODES; ++k) { head[k] = 0; zr[k] = (k < n) ? a.z[k] + (size_t)r * S : nullptr; }
  const int first = a.drop_head, last = total - a.drop_tail;  // keep merged positions [first, last)
  const int M = last - first;
  // z_max: merged position total - n (composite) or the node's last z
  float acc_rgb[3] = {0.f, 0.f, 0.f}, acc_n[3] = {0.f, 0.f, 0.f}, acc_sem[4] = {0.f, 0.f, 0.f, 0.f};
  float acc_mask = 0.f, acc_depth = 0.f, cum = 0.f;
  // pending sample (needs the next z to close its interval)
  bool have = false;
  float pz = 0.f, psig = 0.f, pc[3] = {0, 0, 0}, pn[3] = {0, 0, 0};
  int pcls = 0, kept = 0;
  float* wout = (a.out.fg_weights != nullptr) ? a.out.fg_weights + (size_t)r * M : nullptr;

  auto close_interval = [&](float znext) {
    float fe = (znext - pz) * psig;
    float alpha = 1.0f - expf(-fe);
    float T = expf(-cum);
    float w = alpha * T;
    cum += fe;
    acc_rgb[0] += pc[0] * w, acc_rgb[1] += pc[1] * w, acc_rgb[2] += pc[2] * w;
    acc_n[0] += pn[0] * w, acc_n[1] += pn[1] * w, acc_n[2] += pn[2] * w;
    acc_mask += w;
    acc_depth += pz * w;
    acc_sem[pcls] += w;
    if (wout != nullptr) wout[kept - 1] = w;
  };

  for (int t = 0; t < total; ++t) {
    int best = -1;
    float bz = 0.f;
#pragma unroll
    for (int k = 0; k < HOLD_MAX_NODES; ++k) {
      if (k < n && head[k] < S) {
        float zk = zr[k][head[k]];
        if (best < 0 || zk < bz) { best = k; bz = zk; }
      }
    }
    int j = head[best]++;
    if (t >= first && t < last) {
      if (have) close_interval(bz);
      size_t off = (size_t)r * S + j;
      pz = bz;
      psig = a.density[best][off];
      const float* c = a.color[best] + 3 * off;
      const float* nn = a.normal[best] + 3 * off;
      pc[0] = c[0], pc[1] = c[1], pc[2] = c[2];
      pn[0] = nn[0], pn[1] = nn[1], pn[2] = nn[2];
      pcls = a.class_id[best];
      have = true;
      ++kept;
    } else if (t >= last) {
      // first dropped tail element: merged[total - n] == z_max for the composite (drop_tail == n)
      if (have && t == total - n && !a.single_zmax_last) { close_interval(bz); have = false; }
    }
  }
  if (have) close_interval(pz);  // per-node render: z_max = own last z -> zero-length last interval
  if (a.out.fg_rgb) { a.out.fg_rgb[3 * r] = acc_rgb[0], a.out.fg_rgb[3 * r + 1] = acc_rgb[1], a.out.fg_rgb[3 * r + 2] = acc_rgb[2]; }
  if (a.out.normal) { a.out.normal[3 * r] = acc_n[0], a.out.normal[3 * r + 1] = acc_n[1], a.out.normal[3 * r + 2] = acc_n[2]; }
  if (a.out.mask_prob) a.out.mask_prob[r] = fminf(fmaxf(acc_mask, 0.f), 1.f);
  if (a.out.depth) a.out.depth[r] = acc_depth;
  if (a.out.fg_semantics) {
#pragma unroll
    for (int c = 0; c < 4; ++c) a.out.fg_semantics[4 * r + c] = acc_sem[c];
  }
  if (a.out.bg_weights) a.out.bg_weights[r] = expf(-cum);
}

}  // namespace hold
