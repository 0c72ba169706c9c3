// psdr-b200: device-side first build of the scene BVH (replaces optixAccelBuild, include/psdr/scene/optix.h:277-340, for scenes whose
// topology changes often; the default first build is the host's binned SAH, pb_bvh.cpp, whose trees traverse faster).
//
// Linear BVH after Karras 2012: 30-bit Morton codes of the triangle centroids, a radix sort, one thread per inner node finds its
// range and split from the common prefixes of neighbouring codes. Only the topology is built here: ranges of at most g_lbvh_leaf_max
// sorted triangles become leaves (the leaf encoding of pb_bvh.h), the live inner nodes are renumbered breadth-first — children after
// parents, every level a contiguous index range — and the boxes come from the same level-by-level refit kernels that follow a
// vertex edit (pb_configure.cu). The traversal returns the exact closest hit whatever the tree looks like (ties go to the lower
// triangle id), so hits are bit-identical to those of the SAH tree.
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "pb_host.h"
#include "pb_kernels.h"
#include "pb_lbvh.cuh"

namespace pb {

int g_lbvh_leaf_max = 2;   // most triangles per leaf (1..8; debug key lbvh_leaf): 8 / 4 / 2 / 1 -> 5.59 / 6.10 / 6.36 / 6.44 Grays/s on the bench (profiles/r02ar_*)

__global__ void k_lbvh_morton(int n, const TriRec *__restrict__ tri, float3 lo, float3 inv_ext, unsigned *__restrict__ codes, int *__restrict__ ids) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *q = reinterpret_cast<const float4 *>(tri + i);
    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
    const float third = 1.f / 3.f;
    const float cx = q0.x + (q1.x + q2.x) * third, cy = q0.y + (q1.y + q2.y) * third, cz = q0.z + (q1.z + q2.z) * third;
    codes[i] = lbvh_morton30(cx, cy, cz, lo, inv_ext);
    ids[i] = i;
}

// inner node i of the n - 1: its two child references (pb_lbvh.cuh)
__global__ void k_lbvh_karras(int n, int leaf_max, const unsigned *__restrict__ codes, int2 *__restrict__ children) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const LbvhNode nd = lbvh_inner_node(codes, n, i, leaf_max);
    children[i] = make_int2(nd.left, nd.right);
}

// Breadth-first renumbering, one launch per level. lev[L] .. lev[L + 1]: new indices of level L (lev[0] = 0, lev[1] = 1: the root);
// bfs[new index] = Karras index. The kernel of level L appends the inner children of its nodes behind lev[L + 1] and writes the
// nodes of level L (links in the new numbering, boxes zero until the refit); k_lbvh_close_level then fixes lev[L + 2].
__global__ void k_lbvh_level(int L, const int2 *__restrict__ children, int *__restrict__ bfs, int *__restrict__ lev, unsigned *__restrict__ counter,
                             BvhNode *__restrict__ nodes) {
    const int start = lev[L], end = lev[L + 1];
    const int j = start + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= end) return;
    const int2 ch = children[bfs[j]];
    int link[2] = {ch.x, ch.y};
#pragma unroll
    for (int s = 0; s < 2; ++s)
        if (link[s] >= 0) {
            const int pos = end + (int)atomicAdd(counter, 1u);
            bfs[pos] = link[s];
            link[s] = pos;
        }
    float4 *np = reinterpret_cast<float4 *>(nodes + j);
    np[0] = np[1] = np[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    np[3] = make_float4(__int_as_float(link[0]), __int_as_float(link[1]), 0.f, 0.f);
}
__global__ void k_lbvh_close_level(int L, int *__restrict__ lev, unsigned *__restrict__ counter) {
    lev[L + 2] = lev[L + 1] + (int)*counter;
    *counter = 0u;
}

// Builds the topology into `nodes` / `order` (order[leaf slot] = triangle id) and returns the level offsets (host) for the refit.
// scratch grows as needed. Returns false when the tree is deeper than the level table (the caller falls back to the host builder).
bool lbvh_build(cudaStream_t st, int n, const TriRec *tri, const float *scene_lo, const float *scene_hi, DevBuf &scratch, int *order, BvhNode *nodes,
                std::vector<int> &level_off) {
    constexpr int kMaxLevels = 96;
    const int leaf_max = std::min(8, std::max(1, g_lbvh_leaf_max));
    PB_ASSERT_MSG(n > 16, "internal: lbvh_build needs more than 16 triangles");
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const unsigned *)nullptr, (unsigned *)nullptr, (const int *)nullptr, (int *)nullptr, n, 0, 30, st);
    const size_t o_codes = 0, o_sorted = o_codes + al((size_t)n * 4), o_ids = o_sorted + al((size_t)n * 4), o_children = o_ids + al((size_t)n * 4),
                 o_bfs = o_children + al((size_t)n * 8), o_lev = o_bfs + al((size_t)n * 4), o_counter = o_lev + al((kMaxLevels + 2) * 4), o_cub = o_counter + 256;
    scratch.reserve(o_cub + al(cub_bytes));
    char *base = static_cast<char *>(scratch.p);
    unsigned *codes = reinterpret_cast<unsigned *>(base + o_codes), *sorted = reinterpret_cast<unsigned *>(base + o_sorted);
    int *ids = reinterpret_cast<int *>(base + o_ids), *bfs = reinterpret_cast<int *>(base + o_bfs), *lev = reinterpret_cast<int *>(base + o_lev);
    int2 *children = reinterpret_cast<int2 *>(base + o_children);
    unsigned *counter = reinterpret_cast<unsigned *>(base + o_counter);
    const float3 lo = f3(scene_lo[0], scene_lo[1], scene_lo[2]);
    const float3 inv_ext = f3(1.f / fmaxf(scene_hi[0] - scene_lo[0], 1e-20f), 1.f / fmaxf(scene_hi[1] - scene_lo[1], 1e-20f), 1.f / fmaxf(scene_hi[2] - scene_lo[2], 1e-20f));
    const unsigned g = (unsigned)((n + 255) / 256);
    k_lbvh_morton<<<g, 256, 0, st>>>(n, tri, lo, inv_ext, codes, ids);
    PB_CUDA(cub::DeviceRadixSort::SortPairs(base + o_cub, cub_bytes, codes, sorted, ids, order, n, 0, 30, st));
    k_lbvh_karras<<<g, 256, 0, st>>>(n, leaf_max, sorted, children);
    const int init_lev[2] = {0, 1};
    PB_CUDA(cudaMemsetAsync(lev, 0, (kMaxLevels + 2) * sizeof(int), st));
    PB_CUDA(cudaMemcpyAsync(lev, init_lev, sizeof(init_lev), cudaMemcpyHostToDevice, st));
    PB_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), st));
    PB_CUDA(cudaMemsetAsync(bfs, 0, sizeof(int), st));   // the root is Karras node 0
    for (int L = 0; L < kMaxLevels; ++L) {
        k_lbvh_level<<<g, 256, 0, st>>>(L, children, bfs, lev, counter, nodes);
        k_lbvh_close_level<<<1, 1, 0, st>>>(L, lev, counter);
    }
    int h_lev[kMaxLevels + 2];
    PB_CUDA(cudaMemcpyAsync(h_lev, lev, sizeof(h_lev), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    PB_CUDA(cudaGetLastError());
    if (h_lev[kMaxLevels + 1] != h_lev[kMaxLevels]) return false;   // the last level still had inner children
    level_off.clear();
    for (int L = 0; L <= kMaxLevels; ++L) {
        level_off.push_back(h_lev[L]);
        if (L > 0 && h_lev[L + 1] == h_lev[L]) break;   // level L is empty: h_lev[L] is the node count
    }
    return true;
}

}  // namespace pb
