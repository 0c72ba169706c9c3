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

// 4-wide node obtained by collapsing the binary tree (children in SoA form). child: >= 0 inner node index, < 0 and
// != kBvh4Empty a leaf code (same encoding as HostNode), kBvh4Empty = absent (its box is unreachable).
constexpr int kBvh4Empty = (int)0x80000000;
struct HostNode4 {
    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    int child[4];
    int pad[4];
};
void collapse_bvh4(const std::vector<HostNode> &n2, std::vector<HostNode4> &n4);

}  // namespace pb
