// Host run of the MISE restatement (hold_b200/csrc/mise_phases.h): the same per-item functions the CUDA kernels call,
// driven sequentially.  Interface mirrors the kernels' driver in api.cu: create / query / update / to_dense / destroy.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../hold_b200/csrc/mise_phases.h"

using namespace hold::mise;

struct HostMise {
  Grid g;
  std::vector<float> val;
  std::vector<uint8_t> state;
  std::vector<uint8_t> sub[kMaxDepth];
  std::vector<unsigned int> mark[kMaxDepth];
  std::vector<int> queue;   // lattice indices handed out by the last query
};

extern "C" void* mise_host_create(int res0, int depth, float threshold) {
  HostMise* h = new HostMise();
  Grid& g = h->g;
  g.res0 = res0, g.depth = depth, g.R = res0 << depth, g.G = g.R + 1, g.threshold = threshold;
  const size_t np = (size_t)g.G * g.G * g.G;
  h->val.assign(np, 0.f);
  h->state.assign(np, kNoPoint);
  g.val = h->val.data(), g.state = h->state.data();
  for (int L = 0; L < depth; ++L) {
    const size_t n = (size_t)(res0 << L) * (res0 << L) * (res0 << L);
    h->sub[L].assign(n, 0), h->mark[L].assign(n, 0);
    g.sub[L] = h->sub[L].data(), g.mark[L] = h->mark[L].data();
  }
  const int s0 = 1 << depth;   // initial lattice: the corners of the coarse voxels (mise.pyx:76-88)
  for (int i = 0; i <= res0; ++i)
    for (int j = 0; j <= res0; ++j)
      for (int k = 0; k <= res0; ++k) h->state[pidx(g, i * s0, j * s0, k * s0)] = kUnknown;
  return h;
}
extern "C" void mise_host_destroy(void* p) { delete (HostMise*)p; }

extern "C" int mise_host_query(void* p, int* coords /*[cap,3]*/, int cap) {
  HostMise* h = (HostMise*)p;
  const Grid& g = h->g;
  h->queue.clear();
  for (int x = 0; x < g.G; ++x)
    for (int y = 0; y < g.G; ++y)
      for (int z = 0; z < g.G; ++z)
        if (g.state[pidx(g, x, y, z)] == kUnknown) {
          const int n = (int)h->queue.size();
          if (n < cap) { coords[3 * n] = x, coords[3 * n + 1] = y, coords[3 * n + 2] = z; }
          h->queue.push_back((int)pidx(g, x, y, z));
        }
  return (int)h->queue.size();
}

extern "C" void mise_host_update(void* p, const float* values) {
  HostMise* h = (HostMise*)p;
  Grid& g = h->g;
  for (size_t n = 0; n < h->queue.size(); ++n) { g.val[h->queue[n]] = values[n]; g.state[h->queue[n]] = kKnown; }
  for (int L = 0; L < g.depth; ++L) std::fill(h->mark[L].begin(), h->mark[L].end(), 0u);
  for (int x = 0; x < g.G; ++x)
    for (int y = 0; y < g.G; ++y)
      for (int z = 0; z < g.G; ++z)
        if (g.state[pidx(g, x, y, z)] == kKnown) mark_point(g, x, y, z);
  // all levels in one sweep, coarse to fine: a child created in this sweep has no marks and stays a leaf
  for (int L = 0; L < g.depth; ++L) {
    const int n = g.res0 << L;
    for (int x = 0; x < n; ++x)
      for (int y = 0; y < n; ++y)
        for (int z = 0; z < n; ++z) subdivide_voxel(g, L, x, y, z);
  }
}

extern "C" void mise_host_to_dense(void* p, float* out) {
  HostMise* h = (HostMise*)p;
  const Grid& g = h->g;
  const size_t np = (size_t)g.G * g.G * g.G;
  for (size_t i = 0; i < np; ++i) out[i] = (g.state[i] == kKnown) ? g.val[i] : NAN;
  for (int j = 0; j < g.G; ++j)
    for (int k = 0; k < g.G; ++k) fill_x(g, out, j, k);
  for (int i = 0; i < g.G; ++i)
    for (int k = 0; k < g.G; ++k) fill_y(g, out, i, k);
  for (int i = 0; i < g.G; ++i)
    for (int j = 0; j < g.G; ++j) fill_z(g, out, i, j);
}
