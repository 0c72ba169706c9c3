// psdr-b200: ray regrouping for the traversal kernel — a one-pass counting sort of a wavefront's rays by
// (direction bin, origin cell), with inactive lanes compacted away.
//
// The reference hands OptiX every lane in lane order (src/scene/scene_optix.cpp:101-119). Without RT cores the traversal is
// SIMT code whose cost is set by the longest ray of each warp: in lane order the 32 rays of a warp leave one pixel's hit
// point in 32 unrelated directions and only 5-7 lanes are active per instruction (ncu, profiles/r01_*). After the sort a
// warp holds rays that start in the same region and point the same way, so they walk the same nodes for about as long.
// Only the thread<->ray assignment changes: every hit is written back to its own lane's slot, results are unchanged.
#include <algorithm>

#include "pb_kernels.h"
#include "pb_trace.cuh"

namespace pb {

constexpr int kSortBins = 4096;   // 6 bits direction (octahedral 8x8, Morton) x 6 bits origin cell (4x4x4, Morton)

PB_D int sort_key(float4 a, float4 b, float3 lo, float3 inv_ext) {
    if (!(a.w > 0.f)) return kSortBins;
    const int db = direction_bin(f3(b));
    const int cx = min(3, max(0, (int)((a.x - lo.x) * inv_ext.x * 4.f)));
    const int cy = min(3, max(0, (int)((a.y - lo.y) * inv_ext.y * 4.f)));
    const int cz = min(3, max(0, (int)((a.z - lo.z) * inv_ext.z * 4.f)));
    const int cell = (cx & 1) | ((cy & 1) << 1) | ((cz & 1) << 2) | ((cx & 2) << 2) | ((cy & 2) << 3) | ((cz & 2) << 4);
    return (db << 6) | cell;
}

__global__ void __launch_bounds__(256) k_sort_hist(long long n, const RayRec *__restrict__ rays, float3 lo, float3 inv_ext, unsigned *__restrict__ hist) {
    __shared__ unsigned s_hist[kSortBins + 1];
    for (int t = threadIdx.x; t <= kSortBins; t += blockDim.x) s_hist[t] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
        atomicAdd(&s_hist[sort_key(ldg4(rp), ldg4(rp + 1), lo, inv_ext)], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t <= kSortBins; t += blockDim.x) if (s_hist[t]) atomicAdd(hist + t, s_hist[t]);
}

// exclusive scan of the kSortBins + 1 counters in place; hist[kSortBins + 1] receives the number of active rays
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned *__restrict__ hist) {
    __shared__ unsigned s_part[1024];
    const int t = threadIdx.x;
    constexpr int per = (kSortBins + 1 + 1023) / 1024;   // 5
    unsigned v[per], sum = 0;
    for (int k = 0; k < per; ++k) { const int idx = t * per + k; v[k] = idx <= kSortBins ? hist[idx] : 0u; sum += v[k]; }
    s_part[t] = sum;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned x = t >= o ? s_part[t - o] : 0u;
        __syncthreads();
        s_part[t] += x;
        __syncthreads();
    }
    unsigned base = s_part[t] - sum;
    for (int k = 0; k < per; ++k) {
        const int idx = t * per + k;
        if (idx <= kSortBins) { hist[idx] = base; if (idx == kSortBins) hist[kSortBins + 1] = base; }
        base += v[k];
    }
}

__global__ void __launch_bounds__(256) k_sort_scatter(long long n, const RayRec *__restrict__ rays, float3 lo, float3 inv_ext, unsigned *__restrict__ cursor,
                                                      unsigned *__restrict__ perm, HitRec *__restrict__ hits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
    const int key = sort_key(ldg4(rp), ldg4(rp + 1), lo, inv_ext);
    if (key == kSortBins) {   // inactive lane: a miss, and no thread of the traversal kernel is spent on it
        reinterpret_cast<float4 *>(hits)[i] = make_float4(__int_as_float(-1), __int_as_float(-1), -1.f, -1.f);
        return;
    }
    perm[atomicAdd(cursor + key, 1u)] = (unsigned)i;
}

template <bool FMA_SLAB>
__global__ void __launch_bounds__(128) k_trace_perm(const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf, const unsigned *__restrict__ n_active,
                                                    const unsigned *__restrict__ perm, const RayRec *__restrict__ rays, HitRec *__restrict__ hits) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= __ldg(n_active)) return;
    const unsigned i = __ldg(perm + j);
    const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
    const float4 a = ldg4(rp), b = ldg4(rp + 1);
    const Hit h = trace_closest_spec<FMA_SLAB>(nodes, leaf, f3(a), f3(b), a.w, b.w);
    reinterpret_cast<float4 *>(hits)[i] = make_float4(__int_as_float(h.tri), __int_as_float(h.shape), h.u, h.v);
}

static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

// hist: kSortBins + 2 unsigned (zeroed here); perm: n unsigned
void launch_trace_sorted(cudaStream_t st, const SceneView &S, long long n, const RayRec *rays, HitRec *hits, float3 lo, float3 hi, unsigned *hist, unsigned *perm) {
    if (n <= 0) return;
    const float3 inv_ext = f3(1.f / fmaxf(hi.x - lo.x, 1e-20f), 1.f / fmaxf(hi.y - lo.y, 1e-20f), 1.f / fmaxf(hi.z - lo.z, 1e-20f));
    cudaMemsetAsync(hist, 0, (kSortBins + 2) * sizeof(unsigned), st);
    k_sort_hist<<<(unsigned)std::min<long long>(nblk(n, 256), 148 * 8), 256, 0, st>>>(n, rays, lo, inv_ext, hist);
    k_sort_scan<<<1, 1024, 0, st>>>(hist);
    k_sort_scatter<<<nblk(n, 256), 256, 0, st>>>(n, rays, lo, inv_ext, hist, perm, hits);
    // after the scatter the cursor of bin k has advanced to the start of bin k+1; hist[kSortBins + 1] still holds the active count
    k_trace_perm<true><<<nblk(n, 128), 128, 0, st>>>(S.nodes, S.leaf, hist + kSortBins + 1, perm, rays, hits);
}

}  // namespace pb
