// psdr-b200: helpers shared by the wavefront kernels (lane mapping, record load/store, vertex reconstruction,
// film accumulation).
#pragma once
#include "pb_adjoint_math.cuh"
#include "pb_kernels.h"
#include "pb_shade.cuh"

#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>

namespace pb {

// shard-local lane index -> (pixel, global lane id). A shard owns samples [s0, s0 + spp_local) of every pixel (sample sharding) or
// every sample of the pixels of its image-row tiles (pixel sharding, BASELINE.json configs[3]); either way the global lane id
// (= RNG stream id, integrator.cpp:76) is pix*spp + s, so results do not depend on the GPU count or on the partition.
PB_D long long global_lane(const RenderParams &P, int i, int &pix) {
    const long long li = P.local0 + i;
    const unsigned sl = (unsigned)P.spp_local;
    int lp, s;
    if ((sl & (sl - 1u)) == 0u) {   // the usual power-of-two sample count: shift / mask instead of the 64-bit division routine (uniform branches)
        const int sh = 31 - __clz((int)sl);
        lp = (int)(li >> sh); s = (int)((unsigned)li & (sl - 1u));
    } else if (li < 0x100000000LL) {
        const unsigned q = (unsigned)li / sl;
        lp = (int)q; s = (int)((unsigned)li - q * sl);
    } else {
        lp = (int)(li / P.spp_local); s = (int)(li - (long long)lp * P.spp_local);
    }
    if (P.tile_rows > 0) {   // local row -> global row of this shard's tiles (only the last tile of the image can be partial)
        const int lr = lp / P.width, x = lp - lr * P.width;
        const int t = lr / P.tile_rows;
        lp = ((t * P.world + P.rank) * P.tile_rows + (lr - t * P.tile_rows)) * P.width + x;
    }
    pix = lp;
    return (long long)lp * P.spp + P.s0 + s;
}

// the lane's generator at stream position `jump`: from the context's seed table when it covers the lane (one 128-bit load), else hashed
PB_D Rng make_rng(const RenderParams &P, long long lane, RngJump jump) {
    if (P.rng_seed && lane < P.rng_seed_count) {
        const ulonglong2 s = __ldg(P.rng_seed + lane);
        RngSeed r;
        r.state = s.x; r.inc = s.y;
        return Rng(r, jump);
    }
    return Rng((uint64_t)lane, jump);
}

PB_D void lane_pixel_sample(const RenderParams &P, int pix, float2 jitter, float &sx, float &sy) {
    const int x = pix % P.width, y = pix / P.width;
    sx = div_rn(add_rn((float)x, jitter.x), (float)P.width);
    sy = div_rn(add_rn((float)y, jitter.y), (float)P.height);
}

struct Vertex { Its its; const BsdfRec *bsdf; bool active; HitRec h; float3 ro, rd; float2 film; };   // film: the camera ray's film sample (first event of renderD)

PB_D HitRec load_hit(const HitRec *p) {
    const float4 h = ldg4(reinterpret_cast<const float4 *>(p));
    HitRec r;
    r.tri = __float_as_int(h.x); r.shape = __float_as_int(h.y); r.u = h.z; r.v = h.w;
    return r;
}
// hit of ray j of lane i of an event (EventBuffers::hits / inv)
PB_D HitRec event_hit(const EventBuffers &E, int j, int n, int i) {
    if (E.inv) {
        const unsigned p = __ldg(E.inv + (size_t)j * n + i);
        if (p == 0xffffffffu) { HitRec r; r.tri = -1; r.shape = -1; r.u = r.v = -1.f; return r; }
        return load_hit(E.hits + p);
    }
    return load_hit(E.hits + (size_t)j * n + i);
}
PB_D void store_ray(RayRec *p, float3 o, float3 d, float tmax, float t_occ = 0.f) {
    // one 256-bit store (sm_100a STG.E.256): the 32-byte record is exactly one sector (step time unchanged against two 128-bit stores,
    // profiles/r02an_*). t_occ > 0: occlusion query, any hit closer than t_occ ends the traversal
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(o.x), "f"(o.y), "f"(o.z), "f"(tmax), "f"(d.x), "f"(d.y), "f"(d.z), "f"(t_occ)
                 : "memory");
}

// Start the dependent gathers of an event's connection rays early: read each ray's hit (so that the 16-byte record is in
// L1 when load_hit asks for it) and prefetch the 128-byte triangle row it points at while the lane works on its own vertex.
// Costs no registers; the kernels that reconstruct its1 are bound by this chain of L2 round trips (profiles/r01b_*).
PB_D void prefetch_line(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
PB_D void prefetch_event_hits(const SceneView &S, const HitRec *__restrict__ hits, int rays, int n, int i) {
    for (int j = 0; j < rays; ++j) {
        const int tri = __ldg(reinterpret_cast<const int *>(hits + (size_t)j * n + i));
        if (tri >= 0) {
            const char *row = reinterpret_cast<const char *>(S.tri + tri);
            prefetch_line(row); prefetch_line(row + 32); prefetch_line(row + 64); prefetch_line(row + 96);
        }
    }
}

// EV (template int of the event kernels): -1 = read the event class from BounceParams at run time; otherwise bit 0 = first event
// (camera vertex), bit 1 = the AD formulation — known at launch, so the instantiation drops the other class' code and registers
// (the camera-ray regeneration of renderD's first event is the largest piece).
template <int EV> PB_D bool ev_depth0(const BounceParams &B) { return EV < 0 ? B.depth == 0 : (EV & 1) != 0; }
template <int EV> PB_D bool ev_ad(const BounceParams &B) { return EV < 0 ? B.ad != 0 : (EV & 2) != 0; }

template <int EV = -1>
PB_D Vertex load_vertex(const RenderParams &P, const BounceParams &B, int i, const EventBuffers &E) {
    const HitRec *hit_cur = E.hit_cur;
    Vertex v;
    if (ev_depth0<EV>(B)) {
        if (ev_ad<EV>(B)) {   // renderD: the camera hit is differentiated in solid-angle form (scene.cpp:355-376)
            int pix;
            const long long lane = global_lane(P, i, pix);
            Rng rng = make_rng(P, lane, P.jump0);
            const float2 j = rng.next_2d();
            float sx, sy;
            lane_pixel_sample(P, pix, j, sx, sy);
            float3 o, d;
            sample_primary_ray(P.cam, sx, sy, o, d);
            v.h = load_hit(hit_cur + i);
            v.ro = o; v.rd = d; v.film = make_float2(sx, sy);
            v.its = reconstruct_its_primary(P.S, v.h, o, d);
        } else {
            v.h = load_hit(hit_cur + i);
            v.ro = transform_pos(P.cam.to_world, f3(0.f)); v.rd = f3(0.f);
            v.its = reconstruct_its(P.S, v.h, v.ro);
        }
    } else {
        HitRec h;
        if (E.inv_cur) {   // sorted-copy traversal: the previous event's hits are in stream order
            const unsigned p = __ldg(E.inv_cur + i);
            if (p == 0xffffffffu) { h.tri = -1; h.shape = -1; h.u = h.v = -1.f; }
            else h = load_hit(hit_cur + p);
        } else {
            h = load_hit(hit_cur + i);
        }
        if (ldg4(E.thr_in + i).w != 0.f) h.tri = -1;   // dead path (zero throughput or no continuation): nothing downstream can contribute
        v.h = h;
        v.ro = f3(ldg4(E.prev_pos + i)); v.rd = f3(0.f);
        v.its = reconstruct_its(P.S, h, v.ro);
    }
    v.active = v.its.valid;
    v.bsdf = its_bsdf(P.S, v.its);
    if (P.S.emitter_env >= 0) v.active = v.active && v.bsdf != nullptr;   // direct.cpp:54-57
    return v;
}

// Vertex record (EventBuffers::pos / vb / vc): what the event kernels need of a path vertex. k_shade reconstructs the vertex from its hit,
// the previous position and the triangle row and leaves the record; k_resolve and k_adjoint_lin read it instead of rebuilding the vertex.
// (Having k_resolve(k) write vertex k+1's record — it reconstructs that point as the end of its continuation ray — was 2 % slower: the kernel
// is the one with no registers to spare, profiles/r02ag_*.)
PB_D void store_vertex_rec(float4 *pa, float4 *pb, float4 *pc, int i, const Its &its, bool alive) {
    if (alive) {
        pa[i] = make_float4(its.p.x, its.p.y, its.p.z, its.wi.x);
        pb[i] = make_float4(its.sh.n.x, its.sh.n.y, its.sh.n.z, its.wi.y);
        pc[i] = make_float4(its.uv.x, its.uv.y, __int_as_float(its.shape), its.wi.z);
    } else {   // no hit / dead path: what reconstruct_its returns for an invalid hit
        pa[i] = make_float4(0.f, 0.f, 0.f, 0.f); pb[i] = make_float4(0.f, 0.f, 0.f, 0.f); pc[i] = make_float4(0.f, 0.f, __int_as_float(-1), 0.f);
    }
}
PB_D Vertex load_vertex_rec(const RenderParams &P, const EventBuffers &E, int i) {
    const float4 a = ldg4(E.pos + i), b = ldg4(E.vb + i), c = ldg4(E.vc + i);
    Vertex v;
    v.its.p = f3(a); v.its.wi = f3(a.w, b.w, c.w);
    v.its.sh = Frame(f3(b));
    v.its.uv = make_float2(c.x, c.y);
    v.its.shape = __float_as_int(c.z); v.its.tri = -1;
    v.its.valid = v.its.shape >= 0;
    v.its.n = f3(0.f); v.its.t = 0.f;   // the vertex' own geometric normal / distance are not used by the event kernels
    v.active = v.its.valid;
    v.bsdf = its_bsdf(P.S, v.its);
    if (P.S.emitter_env >= 0) v.active = v.active && v.bsdf != nullptr;   // direct.cpp:54-57
    return v;
}

// warp-segmented sum over lanes that share a pixel (spp consecutive lanes per pixel, integrator.cpp:76-77), then one
// atomicAdd per segment instead of the reference's per-lane scatter_add (integrator.cpp:88)
PB_D void film_accumulate(float *film, int pix, float3 val) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float x = __shfl_down_sync(full, val.x, o), y = __shfl_down_sync(full, val.y, o), z = __shfl_down_sync(full, val.z, o);
        const int p2 = __shfl_down_sync(full, pix, o);
        if (lane + o < 32 && p2 == pix) { val.x += x; val.y += y; val.z += z; }
    }
    const int prev = __shfl_up_sync(full, pix, 1);
    if (pix >= 0 && (lane == 0 || prev != pix)) {
        atomicAdd(film + 3 * (size_t)pix, val.x);
        atomicAdd(film + 3 * (size_t)pix + 1, val.y);
        atomicAdd(film + 3 * (size_t)pix + 2, val.z);
    }
}

// scatter one triangle's adjoint into the triangle-table gradient (only meshes that require a gradient carry bit3)
PB_D bool geom_mode(const SceneView &S) { return S.tri_grad != nullptr || S.tri_tangent != nullptr; }
PB_D void jvp_add(const SceneView &S, float v) {
    if (isfinite(v)) S.jvp_acc[blockIdx.x * blockDim.x + threadIdx.x] += v;   // every wavefront kernel maps thread -> lane this way
}
PB_D void tri_grad_scatter(const SceneView &S, int tri, const TriGrad &g) {
    const float v[22] = {g.p0.x, g.p0.y, g.p0.z, g.e1.x, g.e1.y, g.e1.z, g.e2.x, g.e2.y, g.e2.z, g.n0.x, g.n0.y, g.n0.z,
                         g.n1.x, g.n1.y, g.n1.z, g.n2.x, g.n2.y, g.n2.z, g.fn.x, g.fn.y, g.fn.z, g.area};
    bool ok = true;   // a degenerate sample (zero-length connection, zero pdf) must not poison the whole gradient
#pragma unroll
    for (int k = 0; k < 22; ++k) ok = ok && isfinite(v[k]);
    if (!ok) return;
    if (S.tri_tangent) {   // forward mode: <local gradient, tangent of the triangle record>
        const float *t = S.tri_tangent + (size_t)tri * kTriGradStride;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 22; ++k) if (v[k] != 0.f) s = fmaf(v[k], __ldg(t + k), s);
        jvp_add(S, s);
        return;
    }
    float *p = S.tri_grad + (size_t)tri * kTriGradStride;
#pragma unroll
    for (int k = 0; k < 22; ++k) if (v[k] != 0.f) atomicAdd(p + k, v[k]);
}
// scalar version of film_accumulate for one channel of the derivative image
PB_D void film_accumulate1(float *img, int pix, int channel, float val) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float x = __shfl_down_sync(full, val, o);
        const int p2 = __shfl_down_sync(full, pix, o);
        if (lane + o < 32 && p2 == pix) val += x;
    }
    const int prev = __shfl_up_sync(full, pix, 1);
    if (pix >= 0 && (lane == 0 || prev != pix) && val != 0.f) atomicAdd(img + 3 * (size_t)pix + channel, val);
}

// ---- sensor pose adjoint ----------------------------------------------------------------------------------------------
// Add `g[k]` to 16 consecutive floats at `dst` from whatever lanes of the warp are here together: coalesced-group reduction,
// then one atomic per entry and group (every lane targets the same 16 addresses, per-lane atomics would serialise).
PB_D void sensor_atomic_add16(float *dst, const float *g) {
    namespace cg = cooperative_groups;
    const cg::coalesced_group grp = cg::coalesced_threads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (!grp.any(g[k] != 0.f)) continue;
        const float v = cg::reduce(grp, isfinite(g[k]) ? g[k] : 0.f, cg::plus<float>());
        if (grp.thread_rank() == 0 && v != 0.f) atomicAdd(dst + k, v);
    }
}
// adjoint (g_o, g_d) of a camera ray through film sample (sx, sy): o = M (0,0,0,1), d = M3 d_cam (perspective.cpp:120-136).
// Reverse mode accumulates dL/dM, forward mode returns <dL/dM, tangent of M>.
PB_D float sensor_ray_adjoint(const SceneView &S, const SensorRec &cam, float sx, float sy, float3 g_o, float3 g_d) {
    if (!S.sensor_grad || !finite3(g_o) || !finite3(g_d)) return 0.f;
    const float3 dc = normalize(transform_pos(cam.sample_to_camera, f3(sx, sy, 0.f)));
    const float g[16] = {g_d.x * dc.x, g_d.x * dc.y, g_d.x * dc.z, g_o.x, g_d.y * dc.x, g_d.y * dc.y, g_d.y * dc.z, g_o.y,
                         g_d.z * dc.x, g_d.z * dc.y, g_d.z * dc.z, g_o.z, 0.f, 0.f, 0.f, 0.f};
    if (S.tri_tangent) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 12; ++k) s = fmaf(g[k], __ldg(S.sensor_grad + k), s);
        return s;
    }
    sensor_atomic_add16(S.sensor_grad, g);
    return 0.f;
}
// adjoint of the homogeneous coordinates (g0, g1, g3 for rows 0, 1, 3) of world_to_sample (x, 1) (perspective.cpp:85-96)
PB_D float sensor_projection_adjoint(const SceneView &S, float3 x, float g0, float g1, float g3) {
    if (!S.sensor_grad || !isfinite(g0) || !isfinite(g1) || !isfinite(g3)) return 0.f;
    const float g[16] = {g0 * x.x, g0 * x.y, g0 * x.z, g0, g1 * x.x, g1 * x.y, g1 * x.z, g1, 0.f, 0.f, 0.f, 0.f, g3 * x.x, g3 * x.y, g3 * x.z, g3};
    if (S.tri_tangent) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) s = fmaf(g[k], __ldg(S.sensor_grad + 16 + k), s);
        return s;
    }
    sensor_atomic_add16(S.sensor_grad + 16, g);
    return 0.f;
}

struct TriFull { float3 p0, e1, e2, n0, n1, n2, fn; float area; int flags; };
PB_D TriFull load_tri_full(const SceneView &S, int tri) {
    const float4 *q = reinterpret_cast<const float4 *>(S.tri + tri);
    const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2), q3 = ldg4(q + 3), q4 = ldg4(q + 4), q5 = ldg4(q + 5), q6 = ldg4(q + 6);
    TriFull t;
    t.p0 = f3(q0); t.area = q0.w; t.e1 = f3(q1); t.e2 = f3(q2); t.flags = __float_as_int(q2.w);
    t.n0 = f3(q3); t.n1 = f3(q4); t.n2 = f3(q5); t.fn = f3(q6);
    return t;
}
// adjoint of a path-space point q = p0 + u e1 + v e2 with face normal / Jacobian J = A/detach(A) (scene.cpp:306,326-342,
// mesh.cpp:317-328): g_q, g_n, g_J -> the triangle's record
PB_D void point_on_triangle_scatter(const SceneView &S, int tri, float u, float v, float3 g_q, float3 g_n, float g_J) {
    const float4 *q = reinterpret_cast<const float4 *>(S.tri + tri);
    const float4 q0 = ldg4(q), q2 = ldg4(q + 2);
    if (!(__float_as_int(q2.w) & 8)) return;
    TriGrad g;
    g.p0 = g_q; g.e1 = g_q * u; g.e2 = g_q * v; g.fn = g_n; g.area = g_J / q0.w;
    tri_grad_scatter(S, tri, g);
}

}  // namespace pb
