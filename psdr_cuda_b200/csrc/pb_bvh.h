// psdr-b200: host BVH2 builder interface (see pb_bvh.cpp).
#pragma once
#include <vector>

namespace pb {

struct HostNode {
    float llo[3], lhi[3], rlo[3], rhi[3];
    int left, right;   // >= 0 inner node index; < 0 leaf: v = ~code, first = v >> 3, count = (v & 7) + 1
};

// p0e1e2: 9 floats per triangle (p0, e1, e2). `order[i]` = global triangle id stored at leaf slot i.
// nodes[0] is the root and is always an inner node (an absent child is a far-away degenerate box over slot 0).
void build_bvh(const float *p0e1e2, int n, std::vector<HostNode> &nodes, std::vector<int> &order);

}  // namespace pb
