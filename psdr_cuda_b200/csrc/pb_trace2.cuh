// psdr-b200: second-generation BVH2 traversal for the sorted wavefront (replaces the OptiX launch of
// src/scene/scene_optix.cpp:80-126 / cuda/psdr_cuda.cu:9-45; same contract as pb_trace.cuh: closest Moller-Trumbore hit with
// ties to the lowest triangle id, so the answer is the brute-force one whatever the tree looks like).
//
// ncu on the first-generation kernel (profiles/r01c_k_trace_perm_final_*) shows three limits at once, all near 2/3 busy:
// issue slots, the ALU pipe (FMNMX / compares / selects: 40 of the 80 SASS instructions per node visit, and that pipe takes
// 2 cycles per warp instruction) and the L1 data pipe (64 bytes written back to registers per lane per node). This kernel
// attacks all three per node visit:
//   * boxes as centre + half extent: t_c = c*(1/d) - o/d, t_near/far = t_c -/+ h*|1/d|, issued as nine packed FFMA2 (two fp32 FMAs per
//     instruction, sm_100a) on the FMA pipe, and no per-axis min/max pair; what is left on the ALU pipe is two 3-input min/max,
//     two clamps and the compare per box. 44 SASS instructions per node visit instead of 80.
//   * 48-byte nodes (three 128-bit loads): centres and z half extents in fp32, x / y half extents as bf16 rounded up (left child's
//     in the high halves, used as they are: the low garbage bits only grow the box; right child's in the low halves, one shift).
//   * a sentinel at the bottom of the stack (no empty-stack test per pop).
// and a persistent streaming kernel on top (below) lifts the active lanes per instruction from 17-19 to 24-25 of 32.
#pragma once
#include "pb_trace.cuh"

namespace pb {

// ray constants of the slab test, laid out as the operand pairs of the packed FFMA2 (sm_100a fma.rn.f32x2): (x, y) and (z, z)
struct RaySetup {
    float2 ixy, izz;   // 1/d
    float2 oxy, ozz;   // o/d
};
PB_D RaySetup ray_setup(float3 o, float3 d) {
    RaySetup r;
    const float ix = fminf(fmaxf(clamp_idir(d.x), -1e30f), 1e30f);
    const float iy = fminf(fmaxf(clamp_idir(d.y), -1e30f), 1e30f);
    const float iz = fminf(fmaxf(clamp_idir(d.z), -1e30f), 1e30f);
    r.ixy = make_float2(ix, iy); r.izz = make_float2(iz, iz);
    r.oxy = make_float2(o.x * ix, o.y * iy); r.ozz = make_float2(o.z * iz, o.z * iz);
    return r;
}
// two fp32 FMAs in one instruction (FFMA2); ptxas folds the negations / absolute values into operand modifiers
PB_D float2 fma2(float2 a, float2 b, float2 c) {
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}
PB_D float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
PB_D float2 abs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }

// exact utils.h:67-77 test of the triangles of one leaf (sign-only early outs, accepted hits bit-identical to the oracle's)
PB_D void leaf_test(const LeafTri *__restrict__ leaf, int code, float3 o, float3 d, float tmax, Hit &best) {
    const int v = ~code;
    const int first = v >> 3, cnt = (v & 7) + 1;
    for (int i = 0; i < cnt; ++i) {
        const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
        const F8 t01 = ldg256(tp);
        const float4 ta = t01.lo, tb = t01.hi, tc = ldg4(tp + 2);
        const float3 p0 = f3(ta), e1 = f3(tb), e2 = f3(tc);
        const float3 h = cross(d, e2);
        const float a = dot(e1, h);
        const float3 sv = sub3_rn(o, p0);
        const float U = dot(sv, h);
        const bool guard = fabsf(a) <= 1e18f;
        if (guard && U * a < 0.f && fabsf(U) >= 1e-20f) continue;
        const float3 q = cross(sv, e1);
        const float V = dot(d, q);
        if (guard && V * a < 0.f && fabsf(V) >= 1e-20f) continue;
        const float f = div_rn(1.f, a);
        const float u = mul_rn(f, U), w = mul_rn(f, V), t = mul_rn(f, dot(e2, q));
        const int id = __float_as_int(ta.w);
        if (u >= 0.f && w >= 0.f && add_rn(u, w) <= 1.f && t > kRayEpsilon && t < tmax &&
            (t < best.t || (t == best.t && (best.tri < 0 || id < best.tri)))) {
            best.t = t; best.u = u; best.v = w; best.tri = id; best.shape = __float_as_int(tb.w);
        }
    }
}

// one inner-node visit: returns the next node (near child, or a popped entry)
PB_D int node_step(const BvhNodeC *__restrict__ nodes, int node, const RaySetup &R, float tbest, int *__restrict__ stack, int &sp) {
    const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
    const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2);
    // a = (cL.x, cL.y, cL.z, cR.z)   b = (cR.x, cR.y, left, right)   c = (hx, hy, hL.z, hR.z), hx / hy = bf16(hL) << 16 | bf16(hR)
    const float2 aI = abs2(R.ixy), aIz = abs2(R.izz);
    const float2 tcL = fma2(make_float2(a.x, a.y), R.ixy, neg2(R.oxy));
    const float2 tcZ = fma2(make_float2(a.z, a.w), R.izz, neg2(R.ozz));
    const float2 tcR = fma2(make_float2(b.x, b.y), R.ixy, neg2(R.oxy));
    const float2 hL = make_float2(c.x, c.y), hZ = make_float2(c.z, c.w);
    const float2 hR = make_float2(__uint_as_float(__float_as_uint(c.x) << 16), __uint_as_float(__float_as_uint(c.y) << 16));
    const float2 tnL = fma2(neg2(hL), aI, tcL), tfL = fma2(hL, aI, tcL);
    const float2 tnZ = fma2(neg2(hZ), aIz, tcZ), tfZ = fma2(hZ, aIz, tcZ);
    const float2 tnR = fma2(neg2(hR), aI, tcR), tfR = fma2(hR, aI, tcR);
    const float ln = fmaxf(fmaxf(tnL.x, tnL.y), tnZ.x), lf = fminf(fminf(tfL.x, tfL.y), tfZ.x);
    const float rn = fmaxf(fmaxf(tnR.x, tnR.y), tnZ.y), rf = fminf(fminf(tfR.x, tfR.y), tfZ.y);
    const bool hl = fmaxf(ln, 0.f) <= fminf(lf, tbest), hr = fmaxf(rn, 0.f) <= fminf(rf, tbest);
    const int cl = __float_as_int(b.z), cr = __float_as_int(b.w);
    const bool sw = hr && (!hl || rn < ln);     // go right first
    const int near = sw ? cr : cl, far = sw ? cl : cr;
    if (hl && hr) { stack[sp++] = far; return near; }
    if (hl || hr) return near;
    return stack[--sp];
}

// while-while traversal over the 48-byte nodes; stack[0] holds the sentinel
PB_D Hit trace_closest_c(const BvhNodeC *__restrict__ nodes, const LeafTri *__restrict__ leaf, float3 o, float3 d, float tmax, float t_occ) {
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    const RaySetup R = ray_setup(o, d);
    int stack[64];
    stack[0] = kTraverseDone;
    int sp = 1;
    int node = 0;
    while (node != kTraverseDone) {
        while (node >= 0) node = node_step(nodes, node, R, best.t, stack, sp);
        if (node == kTraverseDone) break;
        leaf_test(leaf, node, o, d, tmax, best);
        if (best.t <= t_occ) break;   // occlusion query: any hit closer than t_occ decides it
        node = stack[--sp];
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}

// ---- persistent streaming traversal -------------------------------------------------------------------------------------
// One-ray-per-thread kernels leave 13 of 32 lanes idle per instruction on this workload (ncu): a warp lasts as long as its longest
// ray, and lanes that reached a leaf wait for lanes still descending. Here a warp is a persistent worker over the sorted ray
// stream: it grabs chunks of up to kStreamChunk consecutive (hence similar) rays with one atomic, stages them 32 at a time in shared memory
// with cp.async (the gather through the sort permutation is off the critical path: no register scoreboard waits on it), and hands a
// staged ray to every lane whose ray has terminated. Lanes are inner-node, triangle or idle lanes; the warp alternates between
// node steps (while at least kNodeMin lanes descend) and single-triangle steps, so both instruction streams run nearly full.
constexpr int kStreamRing = 64;      // staged rays per warp (two blocks of 32)
constexpr unsigned kStreamChunk = 128;    // most rays per atomic grab (StreamArgs::chunk_max)

PB_D void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
PB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
PB_D void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

struct StreamArgs {
    const BvhNodeC *nodes;
    const LeafTri *leaf;
    const unsigned *n_active;   // device counter: rays in the sorted stream
    const unsigned *perm;       // stream position -> ray slot; null: `rays` is the sorted copy and a hit is written at its ray's stream position
    const RayRec *rays;
    HitRec *hits;
    unsigned *counter;          // chunk cursor (zeroed before the launch)
    unsigned chunk_max;         // most rays per grab (kStreamChunk; debug: pb_debug_set trace_chunk)
};

template <int NODE_MIN, int REFILL_MIN, int MINB, int UNROLL>
__global__ void __launch_bounds__(128, MINB) k_trace_stream(StreamArgs A) {
    __shared__ __align__(16) float4 s_ray[4][kStreamRing][2];
    __shared__ unsigned s_src[4][kStreamRing];
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    float4(*ring)[2] = s_ray[wid];
    unsigned *rsrc = s_src[wid];
    const unsigned n = __ldg(A.n_active);

    // warp-uniform stream state
    unsigned chunk_next = 0, chunk_end = 0;   // unstaged part of the current chunk
    unsigned staged = 0, ready = 0, taken = 0;   // counts of rays staged (cp.async issued) / landed / handed out
    bool exhausted = false;
    // block whose permutation entries are in flight: position + count, entry in p_src
    unsigned blk_cnt = 0, p_src = 0;
    auto fetch_block = [&]() {   // A(b): pick the next block of the stream and start loading its permutation entries
        blk_cnt = 0;
        if (exhausted) return;
        if (chunk_next >= chunk_end) {
            // guided self-scheduling: chunks shrink with the work that is left (kStreamChunk rays while there is plenty, 32 at the end),
            // so the last warps finish together; with fixed 256-ray chunks the tail was 10 % of a 12 M-ray launch (one GPU's share of
            // an 8-GPU job). The counter is read without ordering: a stale value only changes the size asked for.
            unsigned base = 0, size = 0;
            if (lane == 0) {
                const unsigned cur = *reinterpret_cast<volatile unsigned *>(A.counter);
                const unsigned rem = cur < n ? n - cur : 0u;
                size = min(A.chunk_max, max(32u, (rem / (gridDim.x * 8u)) & ~31u));
                base = atomicAdd(A.counter, size);
            }
            base = __shfl_sync(full, base, 0);
            size = __shfl_sync(full, size, 0);
            if (base >= n) { exhausted = true; return; }
            chunk_next = base; chunk_end = min(base + size, n);
        }
        blk_cnt = min(32u, chunk_end - chunk_next);
        if ((unsigned)lane < blk_cnt) p_src = A.perm ? __ldcs(A.perm + chunk_next + lane) : chunk_next + lane;
        chunk_next += blk_cnt;
    };
    auto stage_block = [&]() {   // B(b): gather the block's rays into the ring, then A(b+1)
        if (blk_cnt) {
            if ((unsigned)lane < blk_cnt) {
                const unsigned slot = (staged + lane) & (kStreamRing - 1);
                const float4 *rp = reinterpret_cast<const float4 *>(A.rays + p_src);
                cp_async16(&ring[slot][0], rp);
                cp_async16(&ring[slot][1], rp + 1);
                rsrc[slot] = p_src;
            }
            cp_async_commit();
            staged += blk_cnt;
        }
        fetch_block();
    };
    fetch_block();
    stage_block();
    stage_block();

    // per-lane traversal state
    float3 o = f3(0.f), d = f3(0.f);
    float tmax = 0.f, t_occ = 0.f;
    RaySetup R;
    R.ixy = R.izz = R.oxy = R.ozz = make_float2(0.f, 0.f);
    Hit best;
    best.tri = -1; best.shape = -1; best.u = best.v = -1.f; best.t = 0.f;
    int stack[64];
    stack[0] = kTraverseDone;
    int sp = 1;
    int node = kTraverseDone;
    unsigned my_src = 0xffffffffu;

    for (;;) {
        // ---- retire finished rays, hand out staged ones
        const bool idle = (node == kTraverseDone);
        const unsigned m_idle = __ballot_sync(full, idle);
        const int n_idle = __popc(m_idle);
        if (n_idle >= REFILL_MIN || n_idle == 32) {
            if (idle && my_src != 0xffffffffu) {
                __stcs(reinterpret_cast<float4 *>(A.hits) + my_src,
                       make_float4(__int_as_float(best.tri), __int_as_float(best.shape), best.u, best.v));
                my_src = 0xffffffffu;
            }
            if (ready - taken < (unsigned)n_idle && staged > ready) { cp_async_wait_all(); __syncwarp(); ready = staged; }
            const unsigned k = min((unsigned)n_idle, ready - taken);
            const unsigned rank = __popc(m_idle & lt_mask);
            if (idle && rank < k) {
                const unsigned slot = (taken + rank) & (kStreamRing - 1);
                const float4 a = ring[slot][0], b = ring[slot][1];
                my_src = rsrc[slot];
                o = f3(a); d = f3(b); tmax = a.w; t_occ = b.w;
                R = ray_setup(o, d);
                best.tri = -1; best.shape = -1; best.u = best.v = -1.f; best.t = tmax;
                sp = 1;
                node = 0;
            }
            taken += k;
            __syncwarp();
            if (staged - taken <= 32u) stage_block();
            if (k == 0 && n_idle == 32) {
                if (exhausted && taken == staged && blk_cnt == 0) break;
                continue;   // rays are still on their way
            }
        }
        // ---- node steps
        do {
#pragma unroll
            for (int k = 0; k < UNROLL; ++k)
                if (node >= 0) node = node_step(A.nodes, node, R, best.t, stack, sp);
        } while (__popc(__ballot_sync(full, node >= 0)) >= NODE_MIN);
        // ---- one triangle per lane holding a leaf
        if (node < 0 && node != kTraverseDone) {
            const int v = ~node;
            leaf_test(A.leaf, ~(v & ~7), o, d, tmax, best);   // the first triangle of the leaf
            if (best.t <= t_occ) node = kTraverseDone;     // occluded: nothing else matters
            else node = (v & 7) ? ~((((v >> 3) + 1) << 3) | ((v & 7) - 1)) : stack[--sp];
        }
    }
}

}  // namespace pb
