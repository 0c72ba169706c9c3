// psdr-b200: BVH2 closest-hit traversal — replaces the OptiX programs cuda/psdr_cuda.cu:9-45 and the launch in
// src/scene/scene_optix.cpp:80-126.
//
// Contract (same as the reference's closest-hit + miss programs): closest Möller–Trumbore hit with t in
// (RayEpsilon, tmax); outputs global triangle id, shape id and barycentrics (u,v) = weights of vertices 1 and 2;
// -1/-1/-1/-1 on a miss. Ties in t go to the lowest triangle id, so the result is the brute-force answer and does
// not depend on tree shape (boxes are padded on the host, pb_bvh.cpp). Unlike the reference, an inactive lane
// (tmax < 0) is not traced.
#pragma once
#include "pb_scene.cuh"

namespace pb {

struct Hit { int tri, shape; float u, v, t; };

PB_D float clamp_idir(float d) {
    float r = div_rn(1.f, d);
    // +-inf (d == +-0 or denormal) would turn 0*inf into NaN in the slab test; a huge finite value keeps it conservative
    return fminf(fmaxf(r, -3.402823466e38f), 3.402823466e38f);
}

template <int STACK = 64>
PB_D Hit trace_closest(const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf, float3 o, float3 d, float tmax) {
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    const float ix = clamp_idir(d.x), iy = clamp_idir(d.y), iz = clamp_idir(d.z);
    int stack[STACK];
    int sp = 0;
    int node = 0;
    while (true) {
        if (node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2), l = ldg4(np + 3);
            float t0, t1;
            t0 = (a.x - o.x) * ix; t1 = (a.w - o.x) * ix;
            float ln = fminf(t0, t1), lf = fmaxf(t0, t1);
            t0 = (a.y - o.y) * iy; t1 = (b.x - o.y) * iy;
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            t0 = (a.z - o.z) * iz; t1 = (b.y - o.z) * iz;
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            t0 = (b.z - o.x) * ix; t1 = (c.y - o.x) * ix;
            float rn = fminf(t0, t1), rf = fmaxf(t0, t1);
            t0 = (b.w - o.y) * iy; t1 = (c.z - o.y) * iy;
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            t0 = (c.x - o.z) * iz; t1 = (c.w - o.z) * iz;
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            const bool hl = fmaxf(ln, 0.f) <= fminf(lf, best.t), hr = fmaxf(rn, 0.f) <= fminf(rf, best.t);
            int cl = __float_as_int(l.x), cr = __float_as_int(l.y);
            if (hl && hr) {
                if (rn < ln) { int t = cl; cl = cr; cr = t; }
                stack[sp++] = cr;
                node = cl;
                continue;
            }
            if (hl) { node = cl; continue; }
            if (hr) { node = cr; continue; }
        } else {
            const int v = ~node;
            const int first = v >> 3, cnt = (v & 7) + 1;
            for (int i = 0; i < cnt; ++i) {
                const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
                const float4 ta = ldg4(tp), tb = ldg4(tp + 1), tc = ldg4(tp + 2);
                float u, w, t;
                ray_intersect_triangle(f3(ta), f3(tb), f3(tc), o, d, u, w, t);
                const int id = __float_as_int(ta.w);
                if (u >= 0.f && w >= 0.f && add_rn(u, w) <= 1.f && t > kRayEpsilon && t < tmax &&
                    (t < best.t || (t == best.t && (best.tri < 0 || id < best.tri)))) {
                    best.t = t; best.u = u; best.v = w; best.tri = id; best.shape = __float_as_int(tb.w);
                }
            }
        }
        if (sp == 0) break;
        node = stack[--sp];
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}


// while-while variant: every thread first descends inner nodes until it holds a leaf, then intersects leaves; the two
// phases no longer serialise against each other inside a warp (Aila & Laine 2009).
constexpr int kTraverseDone = (int)0x80000000;

template <int STACK = 64>
PB_D Hit trace_closest_ww(const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf, float3 o, float3 d, float tmax) {
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    const float ix = clamp_idir(d.x), iy = clamp_idir(d.y), iz = clamp_idir(d.z);
    int stack[STACK];
    int sp = 0;
    int node = 0;
    while (node != kTraverseDone) {
        while (node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            const float4 a = ldg4(np), b = ldg4(np + 1), c = ldg4(np + 2), l = ldg4(np + 3);
            float t0, t1;
            t0 = (a.x - o.x) * ix; t1 = (a.w - o.x) * ix;
            float ln = fminf(t0, t1), lf = fmaxf(t0, t1);
            t0 = (a.y - o.y) * iy; t1 = (b.x - o.y) * iy;
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            t0 = (a.z - o.z) * iz; t1 = (b.y - o.z) * iz;
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            t0 = (b.z - o.x) * ix; t1 = (c.y - o.x) * ix;
            float rn = fminf(t0, t1), rf = fmaxf(t0, t1);
            t0 = (b.w - o.y) * iy; t1 = (c.z - o.y) * iy;
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            t0 = (c.x - o.z) * iz; t1 = (c.w - o.z) * iz;
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            const bool hl = fmaxf(ln, 0.f) <= fminf(lf, best.t), hr = fmaxf(rn, 0.f) <= fminf(rf, best.t);
            int cl = __float_as_int(l.x), cr = __float_as_int(l.y);
            if (hl && hr) {
                if (rn < ln) { int t = cl; cl = cr; cr = t; }
                stack[sp++] = cr;
                node = cl;
            } else if (hl) node = cl;
            else if (hr) node = cr;
            else node = sp ? stack[--sp] : kTraverseDone;
        }
        while (node < 0 && node != kTraverseDone) {
            const int v = ~node;
            const int first = v >> 3, cnt = (v & 7) + 1;
            for (int i = 0; i < cnt; ++i) {
                const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
                const float4 ta = ldg4(tp), tb = ldg4(tp + 1), tc = ldg4(tp + 2);
                float u, w, t;
                ray_intersect_triangle(f3(ta), f3(tb), f3(tc), o, d, u, w, t);
                const int id = __float_as_int(ta.w);
                if (u >= 0.f && w >= 0.f && add_rn(u, w) <= 1.f && t > kRayEpsilon && t < tmax &&
                    (t < best.t || (t == best.t && (best.tri < 0 || id < best.tri)))) {
                    best.t = t; best.u = u; best.v = w; best.tri = id; best.shape = __float_as_int(tb.w);
                }
            }
            node = sp ? stack[--sp] : kTraverseDone;
        }
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}

// Speculative while-while traversal (Aila & Laine 2009, "postponed leaf"): a lane that reaches a leaf parks it and keeps
// descending inner nodes until every lane of the warp has parked one, then the warp tests triangles together. The ncu
// source view of the plain loop shows why: the triangle test is 30 % of the warp-instructions at 5 of 32 lanes active.
//   FMA_SLAB  box test as fma(lo, 1/d, -o/d) (12 FFMA instead of 12 FADD + 12 FMUL). Its rounding error in world units is
//             2^-24 |o|, covered by the box padding when ray origins lie inside the scene (wavefront rays do; the
//             camera / user rays of pb_trace may not and use the subtract-multiply form).
//   t_occ > 0 occlusion query: stop as soon as any hit closer than t_occ is known (shadow rays only need to know that
//             the emitter sample is blocked; direct.cpp:130-131).
// The triangle test rejects on the sign of the unnormalised barycentrics before paying for the division; survivors run
// the exact utils.h:67-77 arithmetic, so accepted hits are bit-identical to the oracle's.
// 256-bit read-only global load (sm_100a: LDG.E.256): a 64-byte BVH node or leaf triangle is two instructions instead of four / three
struct F8 { float4 lo, hi; };
PB_D F8 ldg256(const void *p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
}

template <bool FMA_SLAB, int STACK = 64, bool LD256 = false>
PB_D Hit trace_closest_spec(const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf, float3 o, float3 d, float tmax, float t_occ) {
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    float ix = clamp_idir(d.x), iy = clamp_idir(d.y), iz = clamp_idir(d.z);
    float ox = o.x, oy = o.y, oz = o.z;
    if (FMA_SLAB) {
        ix = fminf(fmaxf(ix, -1e30f), 1e30f); iy = fminf(fmaxf(iy, -1e30f), 1e30f); iz = fminf(fmaxf(iz, -1e30f), 1e30f);
        ox = o.x * ix; oy = o.y * iy; oz = o.z * iz;
    }
    int stack[STACK];
    int sp = 0;
    int node = 0, parked = 0;   // parked: a leaf code (< 0) waiting to be intersected, 0 = none
    while (node != kTraverseDone) {
        bool searching = true;
        while (node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            float4 a, b, c, l;
            if (LD256) { const F8 n0 = ldg256(np), n1 = ldg256(np + 2); a = n0.lo; b = n0.hi; c = n1.lo; l = n1.hi; }
            else { a = ldg4(np); b = ldg4(np + 1); c = ldg4(np + 2); l = ldg4(np + 3); }
            float t0, t1;
#define PB_SLAB(lo, hi, oc, ic) if (FMA_SLAB) { t0 = fmaf(lo, ic, -oc); t1 = fmaf(hi, ic, -oc); } else { t0 = (lo - oc) * ic; t1 = (hi - oc) * ic; }
            PB_SLAB(a.x, a.w, ox, ix)
            float ln = fminf(t0, t1), lf = fmaxf(t0, t1);
            PB_SLAB(a.y, b.x, oy, iy)
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            PB_SLAB(a.z, b.y, oz, iz)
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            PB_SLAB(b.z, c.y, ox, ix)
            float rn = fminf(t0, t1), rf = fmaxf(t0, t1);
            PB_SLAB(b.w, c.z, oy, iy)
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            PB_SLAB(c.x, c.w, oz, iz)
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
#undef PB_SLAB
            const bool hl = fmaxf(ln, 0.f) <= fminf(lf, best.t), hr = fmaxf(rn, 0.f) <= fminf(rf, best.t);
            int cl = __float_as_int(l.x), cr = __float_as_int(l.y);
            if (hl && hr) {
                if (rn < ln) { int t = cl; cl = cr; cr = t; }
                stack[sp++] = cr;
                node = cl;
            } else if (hl) node = cl;
            else if (hr) node = cr;
            else node = sp ? stack[--sp] : kTraverseDone;
            if (node < 0 && node != kTraverseDone && parked == 0) {   // first leaf: park it, keep descending
                searching = false;
                parked = node;
                node = sp ? stack[--sp] : kTraverseDone;
            }
            if (__ballot_sync(__activemask(), searching) == 0) break;
        }
        while (parked < 0) {
            const int v = ~parked;
            const int first = v >> 3, cnt = (v & 7) + 1;
            for (int i = 0; i < cnt; ++i) {
                const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
                float4 ta, tb, tc;
                if (LD256) { const F8 t01 = ldg256(tp); ta = t01.lo; tb = t01.hi; tc = ldg4(tp + 2); }
                else { ta = ldg4(tp); tb = ldg4(tp + 1); tc = ldg4(tp + 2); }
                const float3 p0 = f3(ta), e1 = f3(tb), e2 = f3(tc);
                // unnormalised barycentrics with the oracle's op order; sign-only early outs (guarded against underflow of u, v)
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
            if (best.t <= t_occ) { node = kTraverseDone; break; }   // occluded: nothing else matters
            parked = 0;
            if (node < 0 && node != kTraverseDone) { parked = node; node = sp ? stack[--sp] : kTraverseDone; }
        }
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}

// Same traversal with the first SK stack levels in shared memory (column tid of an SK x blockDim array: conflict-free whatever
// the lanes' stack depths are, where the local-memory stack costs one L1 wavefront per distinct depth in the warp); deeper
// entries spill to a small local array.
template <bool FMA_SLAB, int SK, bool LD256>
PB_D Hit trace_closest_spec_sstack(int *__restrict__ s_stack, const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf, float3 o, float3 d, float tmax, float t_occ) {
    constexpr int STACK = 64 - SK;
    const int tid = threadIdx.x, nthr = blockDim.x;
#define PB_PUSH(x) do { if (sp < SK) s_stack[sp * nthr + tid] = (x); else stack[sp - SK] = (x); ++sp; } while (0)
#define PB_POP() (--sp, sp < SK ? s_stack[sp * nthr + tid] : stack[sp - SK])
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    float ix = clamp_idir(d.x), iy = clamp_idir(d.y), iz = clamp_idir(d.z);
    float ox = o.x, oy = o.y, oz = o.z;
    if (FMA_SLAB) {
        ix = fminf(fmaxf(ix, -1e30f), 1e30f); iy = fminf(fmaxf(iy, -1e30f), 1e30f); iz = fminf(fmaxf(iz, -1e30f), 1e30f);
        ox = o.x * ix; oy = o.y * iy; oz = o.z * iz;
    }
    int stack[STACK];
    int sp = 0;
    int node = 0, parked = 0;   // parked: a leaf code (< 0) waiting to be intersected, 0 = none
    while (node != kTraverseDone) {
        bool searching = true;
        while (node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            float4 a, b, c, l;
            if (LD256) { const F8 n0 = ldg256(np), n1 = ldg256(np + 2); a = n0.lo; b = n0.hi; c = n1.lo; l = n1.hi; }
            else { a = ldg4(np); b = ldg4(np + 1); c = ldg4(np + 2); l = ldg4(np + 3); }
            float t0, t1;
#define PB_SLAB(lo, hi, oc, ic) if (FMA_SLAB) { t0 = fmaf(lo, ic, -oc); t1 = fmaf(hi, ic, -oc); } else { t0 = (lo - oc) * ic; t1 = (hi - oc) * ic; }
            PB_SLAB(a.x, a.w, ox, ix)
            float ln = fminf(t0, t1), lf = fmaxf(t0, t1);
            PB_SLAB(a.y, b.x, oy, iy)
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            PB_SLAB(a.z, b.y, oz, iz)
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            PB_SLAB(b.z, c.y, ox, ix)
            float rn = fminf(t0, t1), rf = fmaxf(t0, t1);
            PB_SLAB(b.w, c.z, oy, iy)
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            PB_SLAB(c.x, c.w, oz, iz)
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
#undef PB_SLAB
            const bool hl = fmaxf(ln, 0.f) <= fminf(lf, best.t), hr = fmaxf(rn, 0.f) <= fminf(rf, best.t);
            int cl = __float_as_int(l.x), cr = __float_as_int(l.y);
            if (hl && hr) {
                if (rn < ln) { int t = cl; cl = cr; cr = t; }
                PB_PUSH(cr);
                node = cl;
            } else if (hl) node = cl;
            else if (hr) node = cr;
            else node = sp ? PB_POP() : kTraverseDone;
            if (node < 0 && node != kTraverseDone && parked == 0) {   // first leaf: park it, keep descending
                searching = false;
                parked = node;
                node = sp ? PB_POP() : kTraverseDone;
            }
            if (__ballot_sync(__activemask(), searching) == 0) break;
        }
        while (parked < 0) {
            const int v = ~parked;
            const int first = v >> 3, cnt = (v & 7) + 1;
            for (int i = 0; i < cnt; ++i) {
                const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
                float4 ta, tb, tc;
                if (LD256) { const F8 t01 = ldg256(tp); ta = t01.lo; tb = t01.hi; tc = ldg4(tp + 2); }
                else { ta = ldg4(tp); tb = ldg4(tp + 1); tc = ldg4(tp + 2); }
                const float3 p0 = f3(ta), e1 = f3(tb), e2 = f3(tc);
                // unnormalised barycentrics with the oracle's op order; sign-only early outs (guarded against underflow of u, v)
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
            if (best.t <= t_occ) { node = kTraverseDone; break; }   // occluded: nothing else matters
            parked = 0;
            if (node < 0 && node != kTraverseDone) { parked = node; node = sp ? PB_POP() : kTraverseDone; }
        }
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}
#undef PB_PUSH
#undef PB_POP

// ---- shared-memory staged traversal ------------------------------------------------------------------------------------
// ncu shows the per-thread kernels bound by L1 wavefronts: every lane fetches its own 64-byte node with four 16-byte loads,
// each a separate wavefront (about one per cycle per SM), ~70 of them per ray, whatever the tree arity or loop structure.
// The first kTopNodes nodes of the breadth-first ordered tree (the levels every ray walks through) are therefore staged in
// shared memory as four float4 planes, where 32 scattered 16-byte reads cost a handful of cycles instead of 32.
constexpr int kTopNodes = 3072;   // 3072 x 64 B = 192 KiB of the 227 KiB a block may own

template <bool FMA_SLAB, int STACK = 64>
PB_D Hit trace_closest_smem(const float4 *__restrict__ s_nodes, int top, int cap, const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf,
                            float3 o, float3 d, float tmax, float t_occ) {
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    float ix = clamp_idir(d.x), iy = clamp_idir(d.y), iz = clamp_idir(d.z);
    float ox = o.x, oy = o.y, oz = o.z;
    if (FMA_SLAB) {
        ix = fminf(fmaxf(ix, -1e30f), 1e30f); iy = fminf(fmaxf(iy, -1e30f), 1e30f); iz = fminf(fmaxf(iz, -1e30f), 1e30f);
        ox = o.x * ix; oy = o.y * iy; oz = o.z * iz;
    }
    int stack[STACK];
    int sp = 0;
    int node = 0, parked = 0;
    while (node != kTraverseDone) {
        bool searching = true;
        while (node >= 0) {
            float4 a, b, c, l;
            if (node < top) { a = s_nodes[node]; b = s_nodes[cap + node]; c = s_nodes[2 * cap + node]; l = s_nodes[3 * cap + node]; }
            else { const float4 *np = reinterpret_cast<const float4 *>(nodes + node); a = ldg4(np); b = ldg4(np + 1); c = ldg4(np + 2); l = ldg4(np + 3); }
            float t0, t1;
#define PB_SLAB(lo, hi, oc, ic) if (FMA_SLAB) { t0 = fmaf(lo, ic, -oc); t1 = fmaf(hi, ic, -oc); } else { t0 = (lo - oc) * ic; t1 = (hi - oc) * ic; }
            PB_SLAB(a.x, a.w, ox, ix)
            float ln = fminf(t0, t1), lf = fmaxf(t0, t1);
            PB_SLAB(a.y, b.x, oy, iy)
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            PB_SLAB(a.z, b.y, oz, iz)
            ln = fmaxf(ln, fminf(t0, t1)); lf = fminf(lf, fmaxf(t0, t1));
            PB_SLAB(b.z, c.y, ox, ix)
            float rn = fminf(t0, t1), rf = fmaxf(t0, t1);
            PB_SLAB(b.w, c.z, oy, iy)
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
            PB_SLAB(c.x, c.w, oz, iz)
            rn = fmaxf(rn, fminf(t0, t1)); rf = fminf(rf, fmaxf(t0, t1));
#undef PB_SLAB
            const bool hl = fmaxf(ln, 0.f) <= fminf(lf, best.t), hr = fmaxf(rn, 0.f) <= fminf(rf, best.t);
            int cl = __float_as_int(l.x), cr = __float_as_int(l.y);
            if (hl && hr) {
                if (rn < ln) { int t = cl; cl = cr; cr = t; }
                stack[sp++] = cr;
                node = cl;
            } else if (hl) node = cl;
            else if (hr) node = cr;
            else node = sp ? stack[--sp] : kTraverseDone;
            if (node < 0 && node != kTraverseDone && parked == 0) {
                searching = false;
                parked = node;
                node = sp ? stack[--sp] : kTraverseDone;
            }
            if (__ballot_sync(__activemask(), searching) == 0) break;
        }
        while (parked < 0) {
            const int v = ~parked;
            const int first = v >> 3, cnt = (v & 7) + 1;
            for (int i = 0; i < cnt; ++i) {
                const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
                const float4 ta = ldg4(tp), tb = ldg4(tp + 1), tc = ldg4(tp + 2);
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
            if (best.t <= t_occ) { node = kTraverseDone; break; }
            parked = 0;
            if (node < 0 && node != kTraverseDone) { parked = node; node = sp ? stack[--sp] : kTraverseDone; }
        }
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}

// cooperative load of the top of the tree into the four shared-memory planes
PB_D int stage_top_nodes(float4 *s_nodes, const BvhNode *__restrict__ nodes, int num_nodes, int cap) {
    const int top = min(num_nodes, cap);
    for (int k = threadIdx.x; k < top * 4; k += blockDim.x) {
        const int n = k >> 2, q = k & 3;
        s_nodes[q * cap + n] = ldg4(reinterpret_cast<const float4 *>(nodes + n) + q);
    }
    __syncthreads();
    return top;
}

// ---- 4-wide traversal ------------------------------------------------------------------------------------------------
// One node fetch (128 B) tests four boxes; the hit children are ordered with a 5-comparator network on integer keys
// (the entry distance with the child slot in its two low mantissa bits), the nearest is descended into and the others
// are pushed with their entry distance so that a pop can discard them without touching memory once a closer hit exists.
PB_D void sort2(int &a, int &b) { const int lo = min(a, b), hi = max(a, b); a = lo; b = hi; }

template <bool FMA_SLAB, int STACK = 48>
PB_D Hit trace_closest_bvh4(const BvhNode4 *__restrict__ nodes, const LeafTri *__restrict__ leaf, float3 o, float3 d, float tmax, float t_occ) {
    Hit best;
    best.tri = -1; best.shape = -1; best.u = -1.f; best.v = -1.f; best.t = tmax;
    if (!(tmax > 0.f)) { best.t = INFINITY; return best; }
    float ix = clamp_idir(d.x), iy = clamp_idir(d.y), iz = clamp_idir(d.z);
    float ox = o.x, oy = o.y, oz = o.z;
    if (FMA_SLAB) {
        ix = fminf(fmaxf(ix, -1e30f), 1e30f); iy = fminf(fmaxf(iy, -1e30f), 1e30f); iz = fminf(fmaxf(iz, -1e30f), 1e30f);
        ox = o.x * ix; oy = o.y * iy; oz = o.z * iz;
    }
    int stk[STACK];
    float stn[STACK];
    int sp = 0;
    int node = 0;
    const int kInf = 0x7f800000;
    while (true) {
        if (node >= 0) {
            const float4 *np = reinterpret_cast<const float4 *>(nodes + node);
            const float4 lx = ldg4(np), ly = ldg4(np + 1), lz = ldg4(np + 2), hx = ldg4(np + 3), hy = ldg4(np + 4), hz = ldg4(np + 5), ch = ldg4(np + 6);
            const int c0 = __float_as_int(ch.x), c1 = __float_as_int(ch.y), c2 = __float_as_int(ch.z), c3 = __float_as_int(ch.w);
            int key[4];
#define PB_BOX(c, CH, LX, LY, LZ, HX, HY, HZ)                                                                                      \
            {                                                                                                                     \
                float a0, a1, b0, b1, c0, c1;                                                                                     \
                if (FMA_SLAB) { a0 = fmaf(LX, ix, -ox); a1 = fmaf(HX, ix, -ox); b0 = fmaf(LY, iy, -oy); b1 = fmaf(HY, iy, -oy); c0 = fmaf(LZ, iz, -oz); c1 = fmaf(HZ, iz, -oz); } \
                else { a0 = (LX - ox) * ix; a1 = (HX - ox) * ix; b0 = (LY - oy) * iy; b1 = (HY - oy) * iy; c0 = (LZ - oz) * iz; c1 = (HZ - oz) * iz; } \
                const float tn = fmaxf(fmaxf(fminf(a0, a1), fminf(b0, b1)), fmaxf(fminf(c0, c1), 0.f));                            \
                const float tf = fminf(fminf(fmaxf(a0, a1), fmaxf(b0, b1)), fminf(fmaxf(c0, c1), best.t));                          \
                key[c] = (tn <= tf && CH != kTraverseDone) ? ((__float_as_int(tn) & ~3) | c) : kInf;                                                      \
            }
            PB_BOX(0, c0, lx.x, ly.x, lz.x, hx.x, hy.x, hz.x)
            PB_BOX(1, c1, lx.y, ly.y, lz.y, hx.y, hy.y, hz.y)
            PB_BOX(2, c2, lx.z, ly.z, lz.z, hx.z, hy.z, hz.z)
            PB_BOX(3, c3, lx.w, ly.w, lz.w, hx.w, hy.w, hz.w)
#undef PB_BOX
            sort2(key[0], key[1]); sort2(key[2], key[3]); sort2(key[0], key[2]); sort2(key[1], key[3]); sort2(key[1], key[2]);
#define PB_CHILD(k) (((k) & 3) == 0 ? c0 : ((k) & 3) == 1 ? c1 : ((k) & 3) == 2 ? c2 : c3)
            if (key[0] != kInf) {
                if (key[3] != kInf) { stk[sp] = PB_CHILD(key[3]); stn[sp++] = __int_as_float(key[3] & ~3); }
                if (key[2] != kInf) { stk[sp] = PB_CHILD(key[2]); stn[sp++] = __int_as_float(key[2] & ~3); }
                if (key[1] != kInf) { stk[sp] = PB_CHILD(key[1]); stn[sp++] = __int_as_float(key[1] & ~3); }
                node = PB_CHILD(key[0]);
                continue;
            }
#undef PB_CHILD
        } else {
            const int v = ~node;
            const int first = v >> 3, cnt = (v & 7) + 1;
            for (int i = 0; i < cnt; ++i) {
                const float4 *tp = reinterpret_cast<const float4 *>(leaf + first + i);
                const float4 ta = ldg4(tp), tb = ldg4(tp + 1), tc = ldg4(tp + 2);
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
            if (best.t <= t_occ) break;
        }
        // pop, discarding entries that a closer hit has made irrelevant (ties must survive: <=)
        bool found = false;
        while (sp > 0) {
            --sp;
            if (stn[sp] <= best.t) { node = stk[sp]; found = true; break; }
        }
        if (!found) break;
    }
    if (best.tri < 0) best.t = INFINITY;
    return best;
}

// 6-bit direction bin (octahedral map, 8x8, Morton order) used to regroup the rays of a block into coherent warps
PB_D int direction_bin(float3 d) {
    const float inv = 1.f / (fabsf(d.x) + fabsf(d.y) + fabsf(d.z));
    float px = d.x * inv, py = d.y * inv;
    if (d.z < 0.f) {
        const float qx = (1.f - fabsf(py)) * (px >= 0.f ? 1.f : -1.f), qy = (1.f - fabsf(px)) * (py >= 0.f ? 1.f : -1.f);
        px = qx; py = qy;
    }
    int ux = min(7, max(0, (int)((px * .5f + .5f) * 8.f))), uy = min(7, max(0, (int)((py * .5f + .5f) * 8.f)));
    // interleave 3+3 bits
    int m = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) m |= (((ux >> b) & 1) << (2 * b)) | (((uy >> b) & 1) << (2 * b + 1));
    return m;
}

}  // namespace pb
