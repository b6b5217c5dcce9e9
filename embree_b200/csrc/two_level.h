// two_level.h -- host side of the two-level assembly (build.cu assemble_scene; kernels/bvh/bvh_builder_twolevel.cpp:35-240): the small
// top-level BVH8 over the kept per-mesh BVHs.  Host code only (std::vector), shared by build.cu and the CPU emulation the tests run
// (tests/emu/emu.cpp), so the top-level construction is checked without a GPU.
#pragma once
#include <algorithm>
#include <vector>

#include "rt_core.cuh"

namespace rtk {
struct TopItem { float lo[3], hi[3]; int sub; };
// emits the top-level node for items[0..n) at out[q]; n >= 2.  Groups of <= 8 items become one node whose children are the items' root
// copies; larger sets are cut into <= 8 groups by repeated median splits of the largest group along the longest axis of its centres.
inline void top_emit(std::vector<Node8>& out, uint32_t q, std::vector<TopItem> items, const std::vector<Node8>& roots) {
  std::vector<std::vector<TopItem>> groups;
  if (items.size() <= 8) for (auto& it : items) groups.push_back({it});
  else {
    groups.push_back(std::move(items));
    while (groups.size() < 8) {
      size_t big = 0;
      for (size_t g = 1; g < groups.size(); ++g) if (groups[g].size() > groups[big].size()) big = g;
      if (groups[big].size() < 2) break;
      std::vector<TopItem> src = std::move(groups[big]);
      float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
      for (auto& it : src) for (int a = 0; a < 3; ++a) { const float c = it.lo[a] + it.hi[a]; clo[a] = fminf(clo[a], c); chi[a] = fmaxf(chi[a], c); }
      int ax = 0;
      for (int a = 1; a < 3; ++a) if (chi[a] - clo[a] > chi[ax] - clo[ax]) ax = a;
      const size_t mid = src.size() / 2;
      std::nth_element(src.begin(), src.begin() + mid, src.end(), [ax](const TopItem& x, const TopItem& y) { return x.lo[ax] + x.hi[ax] < y.lo[ax] + y.hi[ax]; });
      groups[big] = std::vector<TopItem>(src.begin(), src.begin() + mid);
      groups.push_back(std::vector<TopItem>(src.begin() + mid, src.end()));
    }
  }
  const int n = (int)groups.size();
  ChildBox cb[8];
  float plo[3] = {INFINITY, INFINITY, INFINITY}, phi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int c = 0; c < n; ++c) {
    for (int a = 0; a < 3; ++a) { cb[c].lo[a] = INFINITY; cb[c].hi[a] = -INFINITY; }
    for (auto& it : groups[c]) for (int a = 0; a < 3; ++a) { cb[c].lo[a] = fminf(cb[c].lo[a], it.lo[a]); cb[c].hi[a] = fmaxf(cb[c].hi[a], it.hi[a]); }
    for (int a = 0; a < 3; ++a) { plo[a] = fminf(plo[a], cb[c].lo[a]); phi[a] = fmaxf(phi[a], cb[c].hi[a]); }
  }
  uint8_t slot_of[8];
  assign_slots(cb, n, plo, phi, slot_of);
  int child_at[8];
  for (int s = 0; s < 8; ++s) child_at[s] = -1;
  for (int c = 0; c < n; ++c) child_at[slot_of[c]] = c;
  const uint32_t child_base = (uint32_t)out.size();
  out.resize(out.size() + n);                       // every child is an internal node: n consecutive nodes in slot order
  Node8 nd;
  for (int k = 0; k < 24; ++k) nd.w[k] = 0;
  encode_node_boxes(nd, plo, phi, cb, slot_of, n);
  uint32_t imask = 0;
  for (int s = 0; s < 8; ++s) if (child_at[s] >= 0) imask |= 1u << s;
  nd.w[3] = (nd.w[3] & 0x00FFFFFFu) | (imask << 24);
  nd.w[4] = child_base; nd.w[5] = 0;
  out[q] = nd;
  uint32_t rank = 0;
  for (int s = 0; s < 8; ++s) {
    const int c = child_at[s];
    if (c < 0) continue;
    if (groups[c].size() == 1) out[child_base + rank] = roots[groups[c][0].sub];
    else top_emit(out, child_base + rank, std::move(groups[c]), roots);
    ++rank;
  }
}

// the whole top level for `items` (>= 1): node 0 is the root; with a single mesh the root IS that mesh's root node
inline std::vector<Node8> build_top_level(const std::vector<TopItem>& items, const std::vector<Node8>& roots) {
  std::vector<Node8> top(1);
  if (items.size() == 1) top[0] = roots[items[0].sub];
  else top_emit(top, 0, items, roots);
  return top;
}
}  // namespace rtk
