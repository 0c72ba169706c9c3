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
#include "pb_trace2.cuh"
#include "pb_sortkey.cuh"

namespace pb {

int g_sort_mode = SORT_CELL_OCTANT;   // 5: origin cell (8x8x8, Morton) x direction octant; 0: 6-bit direction bin x 4x4x4 cells (A/B: profiles/r02e_*; 16^3 cells without direction: r02m_*)
int g_trace_kernel = 3;     // 3 persistent streaming kernel; 1 one ray per thread over the same nodes (A/B, profiles/r02a_*)
int g_trace_chunk = 128;    // most rays per chunk grab of the streaming kernel (128 vs 256: +1 % on 8 Mi-lane wavefronts, profiles/r02q_*)
int g_trace_blocks = 8;     // debug: persistent blocks per SM launched (<= 8 resident)
int g_trace_node_min = 16;  // streaming kernel: node steps continue while at least this many lanes descend (1 = plain while-while)

// The histogram pass is the only one that reads the rays: it leaves each ray's 13-bit key in `keys` (2 B instead of 32 B for
// the scatter pass to read back). rays == nullptr: the producer of the rays (k_shade) has written the keys already.
__global__ void __launch_bounds__(1024) k_sort_hist(long long n, const RayRec *__restrict__ rays, float3 lo, float3 inv_ext, int mode, unsigned *__restrict__ hist,
                                                    unsigned short *__restrict__ keys) {
    extern __shared__ unsigned s_hist[];
    for (int t = threadIdx.x; t <= kSortBins; t += blockDim.x) s_hist[t] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int key;
        if (rays) {
            const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
            const float4 a = ldg4(rp), b = ldg4(rp + 1);
            key = sort_key(f3(a), a.w, f3(b), lo, inv_ext, mode);
            keys[i] = (unsigned short)key;
        } else {
            key = __ldcs(keys + i);
        }
        atomicAdd(&s_hist[key], 1u);
    }
    __syncthreads();
    for (int t = threadIdx.x; t <= kSortBins; t += blockDim.x) if (s_hist[t]) atomicAdd(hist + t, s_hist[t]);
}

// exclusive scan of the kSortBins + 1 counters in place; hist[kSortBins + 1] receives the number of active rays
__global__ void __launch_bounds__(1024) k_sort_scan(unsigned *__restrict__ hist, unsigned long long *__restrict__ active_total) {
    __shared__ unsigned s_part[1024];
    const int t = threadIdx.x;
    constexpr int per = (kSortBins + 1 + 1023) / 1024;
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
        if (idx <= kSortBins) {
            hist[idx] = base;
            if (idx == kSortBins) { hist[kSortBins + 1] = base; if (active_total) atomicAdd(active_total, (unsigned long long)base); }   // rays actually traced
        }
        base += v[k];
    }
}

// Scatter with warp- and block-level aggregation. Rays arrive in pixel order, so the lanes in flight at any moment share an origin
// cell and fall into a few dozen of the 4096 bins: a warp first groups its lanes by key (match.any: at most 8 octants of one cell),
// one lane per group ranks the group inside the block with a shared-memory atomic, and one global atomic per (block, non-empty
// bin) reserves the block's range. A block handles kScatterItems rays per thread: the zeroing / reservation sweep over the 4096
// shared counters is paid once per 8192+ rays, and a bin's rays of one block land in consecutive slots (whole sectors).
// COPY: instead of the permutation (stream position -> ray slot) the pass writes the rays themselves in stream order and the inverse map
// (ray slot -> stream position, ~0 for an inactive lane): the traversal kernel then reads its rays and writes its hits contiguously, and the
// consumer of the hits follows `inv`.
template <int THREADS, int ITEMS, bool COPY>
__global__ void __launch_bounds__(THREADS) k_sort_scatter(long long n, const unsigned short *__restrict__ keys, unsigned *__restrict__ cursor,
                                                          unsigned *__restrict__ perm, HitRec *__restrict__ hits, const RayRec *__restrict__ rays,
                                                          RayRec *__restrict__ sorted, unsigned *__restrict__ inv) {
    extern __shared__ unsigned s_cnt[];
    for (int t = threadIdx.x; t <= kSortBins; t += THREADS) s_cnt[t] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * (THREADS * ITEMS) + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u, lt_mask = (1u << lane) - 1u;
    unsigned kr[ITEMS];   // key << 16 | rank inside the block (< 65536)
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const long long i = base + (long long)k * THREADS;
        unsigned key = kSortBins, rank = 0;
        if (i < n) {
            key = __ldcs(keys + i);
            if (key == kSortBins) {   // inactive lane: a miss
                if (COPY) inv[i] = 0xffffffffu;
                else reinterpret_cast<float4 *>(hits)[i] = make_float4(__int_as_float(-1), __int_as_float(-1), -1.f, -1.f);
            }
        }
        const unsigned grp = __match_any_sync(0xffffffffu, key);
        if (key != kSortBins) {
            unsigned first = 0;
            const int leader = __ffs(grp) - 1;
            if ((int)lane == leader) first = atomicAdd(&s_cnt[key], (unsigned)__popc(grp));
            rank = __shfl_sync(grp, first, leader) + (unsigned)__popc(grp & lt_mask);
        }
        kr[k] = key << 16 | rank;
    }
    __syncthreads();
    {   // reserve the block's range of every non-empty bin: the atomics of a thread's bins are issued together
        constexpr int PER = kSortBins / THREADS;
        unsigned c[PER], g[PER];
#pragma unroll
        for (int k = 0; k < PER; ++k) c[k] = s_cnt[threadIdx.x + k * THREADS];
#pragma unroll
        for (int k = 0; k < PER; ++k) g[k] = c[k] ? atomicAdd(cursor + threadIdx.x + k * THREADS, c[k]) : 0u;
#pragma unroll
        for (int k = 0; k < PER; ++k) s_cnt[threadIdx.x + k * THREADS] = g[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const unsigned key = kr[k] >> 16;
        if (key != kSortBins) {
            const unsigned pos = s_cnt[key] + (kr[k] & 0xffffu);
            const long long i = base + (long long)k * THREADS;
            if (COPY) {
                const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
                const float4 a = __ldcs(rp), b = __ldcs(rp + 1);
                float4 *sp = reinterpret_cast<float4 *>(sorted + pos);
                sp[0] = a; sp[1] = b;
                inv[i] = pos;
            } else {
                perm[pos] = (unsigned)i;
            }
        }
    }
}

// one ray per thread over the compact nodes (pb_trace2.cuh)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_trace_compact(const BvhNodeC *__restrict__ nodes, const LeafTri *__restrict__ leaf, const unsigned *__restrict__ n_active,
                                                             const unsigned *__restrict__ perm, const RayRec *__restrict__ rays, HitRec *__restrict__ hits) {
    const unsigned j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= __ldg(n_active)) return;
    const unsigned i = __ldcs(perm + j);
    const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
    const float4 a = __ldcs(rp), b = __ldcs(rp + 1);
    const Hit h = trace_closest_c(nodes, leaf, f3(a), f3(b), a.w, b.w);
    __stcs(reinterpret_cast<float4 *>(hits) + i, make_float4(__int_as_float(h.tri), __int_as_float(h.shape), h.u, h.v));
}

// sorted-copy mode, for consumers that index hits by ray slot: hits[i] = sorted_hits[inv[i]] (a miss where inv[i] = ~0)
__global__ void __launch_bounds__(256) k_unpermute_hits(long long n, const unsigned *__restrict__ inv, const HitRec *__restrict__ sorted_hits, HitRec *__restrict__ hits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned p = __ldcs(inv + i);
    float4 h = make_float4(__int_as_float(-1), __int_as_float(-1), -1.f, -1.f);
    if (p != 0xffffffffu) h = __ldcs(reinterpret_cast<const float4 *>(sorted_hits) + p);
    reinterpret_cast<float4 *>(hits)[i] = h;
}

static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }
void launch_unpermute_hits(cudaStream_t st, long long n, const unsigned *inv, const HitRec *sorted_hits, HitRec *hits) {
    if (n > 0) k_unpermute_hits<<<nblk(n, 256), 256, 0, st>>>(n, inv, sorted_hits, hits);
}

// hist: kSortBins + 2 unsigned (zeroed here); perm: n unsigned; keys: n unsigned short
// sorted != nullptr: sorted-copy mode (see k_sort_scatter<COPY>): `hits` receives the hits in stream order, inv[ray slot] their positions
void launch_trace_sorted(cudaStream_t st, const SceneView &S, long long n, const RayRec *rays, HitRec *hits, float3 lo, float3 hi, unsigned *hist, unsigned *perm,
                         unsigned short *keys, unsigned *stream_counter, unsigned long long *active_total, cudaEvent_t ev0, cudaEvent_t ev1, int mode, bool keys_ready,
                         RayRec *sorted, unsigned *inv) {
    if (n <= 0) return;
    const float3 inv_ext = f3(1.f / fmaxf(hi.x - lo.x, 1e-20f), 1.f / fmaxf(hi.y - lo.y, 1e-20f), 1.f / fmaxf(hi.z - lo.z, 1e-20f));
    cudaMemsetAsync(hist, 0, (kSortBins + 2) * sizeof(unsigned), st);   // hist must hold kSortBins + 2 counters
    const int cnt_bytes = (kSortBins + 1) * (int)sizeof(unsigned);
    k_sort_hist<<<(unsigned)std::min<long long>(nblk(n, 1024), 148), 1024, cnt_bytes, st>>>(n, keys_ready ? nullptr : rays, lo, inv_ext, mode, hist, keys);
    k_sort_scan<<<1, 1024, 0, st>>>(hist, active_total);
    if (sorted) k_sort_scatter<1024, 8, true><<<nblk(n, 1024 * 8), 1024, cnt_bytes, st>>>(n, keys, hist, perm, hits, rays, sorted, inv);
    else k_sort_scatter<1024, 8, false><<<nblk(n, 1024 * 8), 1024, cnt_bytes, st>>>(n, keys, hist, perm, hits, rays, sorted, inv);   // 512 x 8, 1024 x 4, 512 x 16, 256 x 16: the same within noise; 1024 x 16 / x 32: -1.5 % (profiles/r02ab_*)
    // after the scatter the cursor of bin k has advanced to the start of bin k+1; hist[kSortBins + 1] still holds the active count
    if (g_trace_kernel == 3) cudaMemsetAsync(stream_counter, 0, sizeof(unsigned), st);
    if (ev0) cudaEventRecord(ev0, st);   // the pair brackets the traversal kernel alone (roofline: 48 B per traced ray / this duration)
    if (g_trace_kernel == 3) {
        StreamArgs A;
        A.nodes = S.nodes_c; A.leaf = S.leaf; A.n_active = hist + kSortBins + 1; A.perm = sorted ? nullptr : perm; A.rays = sorted ? sorted : rays; A.hits = hits; A.counter = stream_counter;
        A.chunk_max = (unsigned)std::max(32, g_trace_chunk & ~31);
        const unsigned grid = (unsigned)std::min<long long>(nblk(n, 128), 148LL * std::max(1, std::min(8, g_trace_blocks)));   // persistent: 8 blocks of 4 warps per SM
        if (g_trace_node_min == 1) k_trace_stream<1, 12, 8, 2><<<grid, 128, 0, st>>>(A);    // node steps until no lane descends (plain while-while; 40 % slower)
        else k_trace_stream<16, 12, 8, 2><<<grid, 128, 0, st>>>(A);
    } else {
        k_trace_compact<8><<<nblk(n, 128), 128, 0, st>>>(S.nodes_c, S.leaf, hist + kSortBins + 1, perm, rays, hits);
    }
    if (ev1) cudaEventRecord(ev1, st);
}

}  // namespace pb
