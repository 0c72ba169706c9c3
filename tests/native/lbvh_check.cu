// Host-compiled check of the device LBVH's per-node arithmetic (psdr_cuda_b200/csrc/pb_lbvh.cuh): for random, clustered and duplicated
// centroids the radix tree of lbvh_inner_node must be a binary tree in which every sorted slot lies in exactly one leaf, every child's range
// partitions its parent's, leaves hold at most leaf_max slots, and every inner node referenced from the root is referenced once.
//   nvcc -x cu -std=c++17 -O1 --expt-relaxed-constexpr -o lbvh_check tests/native/lbvh_check.cu && ./lbvh_check
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>

#include "../../psdr_cuda_b200/csrc/pb_lbvh.cuh"

using namespace pb;

static int g_fail = 0;
#define CHECK(c, ...) do { if (!(c)) { if (g_fail < 20) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } ++g_fail; } } while (0)

static void run_case(int n, int leaf_max, int kind, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<unsigned> codes(n);
    const float3 lo = f3(-1.f, -2.f, 0.5f), inv_ext = f3(1.f / 4.f, 1.f / 3.f, 1.f / 0.25f);
    for (int i = 0; i < n; ++i) {
        float x = U(rng), y = U(rng), z = U(rng);
        if (kind == 1) { x = 0.5f + 0.01f * x; y = 0.25f + 0.01f * y; z = 0.75f + 0.001f * z; }        // one tight cluster: long common prefixes
        if (kind == 2 && (i % 3)) { x = 0.3f; y = 0.6f; z = 0.9f; }                                      // two thirds share one code: index tie-break
        if (kind == 3) { x = (i % 7) / 7.f; y = 0.f; z = 0.f; }                                          // seven distinct codes
        codes[i] = lbvh_morton30(lo.x + 4.f * x, lo.y + 3.f * y, lo.z + 0.25f * z, lo, inv_ext);
        CHECK(codes[i] < (1u << 30), "code out of range");
    }
    std::sort(codes.begin(), codes.end());
    std::vector<LbvhNode> nodes(n - 1);
    for (int i = 0; i < n - 1; ++i) nodes[i] = lbvh_inner_node(codes.data(), n, i, leaf_max);
    CHECK(nodes[0].first == 0 && nodes[0].last == n - 1, "root range [%d, %d]", nodes[0].first, nodes[0].last);
    std::vector<int> covered(n, 0), visits(n - 1, 0);
    std::vector<int> stack = {0};
    int depth_guard = 0;
    while (!stack.empty() && depth_guard++ < 4 * n) {
        const int i = stack.back(); stack.pop_back();
        CHECK(i >= 0 && i < n - 1, "inner index %d", i);
        if (i < 0 || i >= n - 1) continue;
        ++visits[i];
        const LbvhNode &nd = nodes[i];
        int split = -1;   // the left child's last slot
        const int ref[2] = {nd.left, nd.right};
        int lo_s = nd.first;
        for (int s = 0; s < 2; ++s) {
            int first, last;
            if (ref[s] < 0) {
                const int v = ~ref[s];
                first = v >> 3; last = first + (v & 7);
                CHECK(last - first + 1 <= leaf_max, "leaf of %d slots (max %d)", last - first + 1, leaf_max);
                for (int k = first; k <= last && k < n; ++k) ++covered[k];
            } else {
                CHECK(ref[s] < n - 1, "child index %d", ref[s]);
                if (ref[s] >= n - 1) continue;
                first = nodes[ref[s]].first; last = nodes[ref[s]].last;
                CHECK(last - first + 1 > leaf_max, "inner child with %d slots should be a leaf", last - first + 1);
                stack.push_back(ref[s]);
            }
            CHECK(first == lo_s, "child %d of node %d starts at %d, expected %d", s, i, first, lo_s);
            if (s == 0) split = last;
            lo_s = last + 1;
        }
        CHECK(lo_s == nd.last + 1, "children of node %d end at %d, node ends at %d", i, lo_s - 1, nd.last);
        CHECK(split >= nd.first && split < nd.last, "split %d outside [%d, %d)", split, nd.first, nd.last);
    }
    for (int k = 0; k < n; ++k) CHECK(covered[k] == 1, "slot %d covered %d times (n %d leaf_max %d kind %d)", k, covered[k], n, leaf_max, kind);
    for (int i = 0; i < n - 1; ++i) CHECK(visits[i] <= 1, "inner node %d referenced %d times", i, visits[i]);
}

int main() {
    int cases = 0;
    for (int n : {17, 64, 1000, 50000})
        for (int leaf_max : {1, 2, 4, 8})
            for (int kind = 0; kind < 4; ++kind) { run_case(n, leaf_max, kind, 1234u + 17u * (unsigned)cases); ++cases; }
    if (g_fail) { std::printf("lbvh_check: %d failures\n", g_fail); return 1; }
    std::printf("lbvh_check: ok (%d cases)\n", cases);
    return 0;
}
