// psdr-b200: host BVH2 builder (binned SAH) — replaces optixAccelBuild (include/psdr/scene/optix.h:277-340).
//
// Output is the 64-byte two-child node layout of pb_scene.cuh plus the triangle order of the leaves. Boxes are
// padded so the (rounded) slab test can never cull a triangle the exact ray/triangle test accepts: the traversal
// then returns exactly the closest Möller–Trumbore hit, independent of tree shape.
#include "pb_bvh.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace pb {
namespace {

struct Box {
    float lo[3], hi[3];
    void reset() { for (int k = 0; k < 3; ++k) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -lo[k]; } }
    void grow(const Box &b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
    void grow(const float *p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
    float half_area() const {
        float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
        if (dx < 0.f) return 0.f;
        return dx * dy + dy * dz + dz * dx;
    }
};

constexpr int kBins = 16;
constexpr int kMaxLeaf = 4;

struct Builder {
    const std::vector<Box> &tb;
    std::vector<float> cen;     // 3 per triangle
    std::vector<int> order;
    std::vector<HostNode> nodes;

    explicit Builder(const std::vector<Box> &b) : tb(b) {
        int n = (int)b.size();
        cen.resize(3 * (size_t)n);
        order.resize(n);
        for (int i = 0; i < n; ++i) {
            order[i] = i;
            for (int k = 0; k < 3; ++k) cen[3 * (size_t)i + k] = 0.5f * (b[i].lo[k] + b[i].hi[k]);
        }
    }

    static int leaf_code(int first, int count) { return ~((first << 3) | (count - 1)); }

    // returns child code for the range [first, first+count) and its box
    int build(int first, int count, Box &box) {
        box.reset();
        Box cb;
        cb.reset();
        for (int i = first; i < first + count; ++i) { box.grow(tb[order[i]]); cb.grow(&cen[3 * (size_t)order[i]]); }
        if (count <= kMaxLeaf) {
            // try to keep small leaves only if SAH says so: always split down to <= kMaxLeaf, leaf when <= kMaxLeaf and cheap
            if (count <= 2) return leaf_code(first, count);
        }
        int axis = 0;
        float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
        if (ext[1] > ext[axis]) axis = 1;
        if (ext[2] > ext[axis]) axis = 2;
        int mid = -1;
        if (ext[axis] > 0.f) {
            float best_cost = std::numeric_limits<float>::infinity();
            int best_axis = -1, best_bin = -1;
            for (int ax = 0; ax < 3; ++ax) {
                if (!(ext[ax] > 0.f)) continue;
                Box bins[kBins];
                int cnt[kBins] = {0};
                for (auto &b : bins) b.reset();
                float scale = (float)kBins / ext[ax];
                for (int i = first; i < first + count; ++i) {
                    int t = order[i];
                    int b = std::min(kBins - 1, std::max(0, (int)((cen[3 * (size_t)t + ax] - cb.lo[ax]) * scale)));
                    bins[b].grow(tb[t]);
                    cnt[b]++;
                }
                float right_area[kBins];
                Box acc;
                acc.reset();
                int rc[kBins];
                int c = 0;
                for (int b = kBins - 1; b > 0; --b) { acc.grow(bins[b]); c += cnt[b]; right_area[b] = acc.half_area(); rc[b] = c; }
                acc.reset();
                c = 0;
                for (int b = 0; b < kBins - 1; ++b) {
                    acc.grow(bins[b]);
                    c += cnt[b];
                    if (c == 0 || rc[b + 1] == 0) continue;
                    float cost = acc.half_area() * (float)c + right_area[b + 1] * (float)rc[b + 1];
                    if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
                }
            }
            float leaf_cost = box.half_area() * (float)count;
            if (best_axis >= 0 && !(count <= kMaxLeaf && leaf_cost <= best_cost + box.half_area())) {
                float scale = (float)kBins / ext[best_axis];
                auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int t) {
                    int b = std::min(kBins - 1, std::max(0, (int)((cen[3 * (size_t)t + best_axis] - cb.lo[best_axis]) * scale)));
                    return b <= best_bin;
                });
                mid = (int)(it - order.begin());
                if (mid == first || mid == first + count) mid = -1;
            } else if (count <= kMaxLeaf) {
                return leaf_code(first, count);
            }
        }
        if (mid < 0) {
            if (count <= 8 && !(ext[axis] > 0.f)) return leaf_code(first, count);   // coincident centroids
            mid = first + count / 2;
            std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                             [&](int a, int b) { return cen[3 * (size_t)a + axis] < cen[3 * (size_t)b + axis]; });
        }
        int id = (int)nodes.size();
        nodes.emplace_back();
        Box lb, rb;
        int l = build(first, mid - first, lb);
        int r = build(mid, first + count - mid, rb);
        HostNode &n = nodes[id];
        for (int k = 0; k < 3; ++k) { n.llo[k] = lb.lo[k]; n.lhi[k] = lb.hi[k]; n.rlo[k] = rb.lo[k]; n.rhi[k] = rb.hi[k]; }
        n.left = l; n.right = r;
        return id;
    }
};

// pad = 1e-4 relative to the coordinate (rounded ray/triangle test vs exact boxes) + 2.5e-7 of the scene extent (error of
// the fused slab test for ray origins inside the scene, pb_trace.cuh)
void pad_box(float *lo, float *hi, float extent) {
    for (int k = 0; k < 3; ++k) {
        if (!(hi[k] < std::numeric_limits<float>::max())) continue;
        float pad = 1e-4f * std::max(1.f, std::max(std::fabs(lo[k]), std::fabs(hi[k]))) + 2.5e-7f * extent;
        lo[k] -= pad; hi[k] += pad;
    }
}

}  // namespace

void build_bvh(const float *p0e1e2, int n, std::vector<HostNode> &nodes, std::vector<int> &order) {
    std::vector<Box> tb(n);
    for (int i = 0; i < n; ++i) {
        const float *t = p0e1e2 + 9 * (size_t)i;
        float a[3] = {t[0], t[1], t[2]}, b[3], c[3];
        for (int k = 0; k < 3; ++k) { b[k] = t[k] + t[3 + k]; c[k] = t[k] + t[6 + k]; }
        tb[i].reset();
        tb[i].grow(a); tb[i].grow(b); tb[i].grow(c);
    }
    Builder B(tb);
    B.nodes.reserve(n > 0 ? 2 * (size_t)n : 1);
    const float far_away = std::numeric_limits<float>::max();
    if (n == 0) {   // no geometry: both children are unreachable degenerate boxes over (dummy) leaf slot 0
        HostNode root;
        for (int k = 0; k < 3; ++k) root.llo[k] = root.lhi[k] = root.rlo[k] = root.rhi[k] = far_away;
        root.left = root.right = Builder::leaf_code(0, 1);
        B.nodes.push_back(root);
        B.order.assign(1, 0);
    } else {
        Box box;
        int code = B.build(0, n, box);
        if (code < 0) {   // a single leaf: wrap it in a root whose right child is unreachable
            HostNode root;
            for (int k = 0; k < 3; ++k) { root.llo[k] = box.lo[k]; root.lhi[k] = box.hi[k]; root.rlo[k] = root.rhi[k] = far_away; }
            root.left = code; root.right = Builder::leaf_code(0, 1);
            B.nodes.push_back(root);
        }
    }
    float extent = 0.f;
    for (const Box &b : tb) for (int k = 0; k < 3; ++k) extent = std::max(extent, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
    for (auto &nd : B.nodes) { pad_box(nd.llo, nd.lhi, extent); pad_box(nd.rlo, nd.rhi, extent); }
    {   // breadth-first renumbering: a tree level is a contiguous index range (the device refit walks the levels bottom-up)
        std::vector<int> order_bfs, new_id(B.nodes.size(), -1);
        order_bfs.reserve(B.nodes.size());
        order_bfs.push_back(0);
        for (size_t h = 0; h < order_bfs.size(); ++h) {
            const HostNode &nd = B.nodes[order_bfs[h]];
            if (nd.left >= 0) order_bfs.push_back(nd.left);
            if (nd.right >= 0) order_bfs.push_back(nd.right);
        }
        for (size_t k = 0; k < order_bfs.size(); ++k) new_id[order_bfs[k]] = (int)k;
        std::vector<HostNode> re(order_bfs.size());
        for (size_t k = 0; k < order_bfs.size(); ++k) {
            re[k] = B.nodes[order_bfs[k]];
            if (re[k].left >= 0) re[k].left = new_id[re[k].left];
            if (re[k].right >= 0) re[k].right = new_id[re[k].right];
        }
        B.nodes.swap(re);
    }
    nodes.swap(B.nodes);
    order.swap(B.order);
}

}  // namespace pb
