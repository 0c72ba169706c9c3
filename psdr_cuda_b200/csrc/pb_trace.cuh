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


constexpr int kTraverseDone = (int)0x80000000;

// 256-bit read-only global load (sm_100a: LDG.E.256): a 64-byte leaf triangle is two instructions instead of three
struct F8 { float4 lo, hi; };
PB_D F8 ldg256(const void *p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
    return r;
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
