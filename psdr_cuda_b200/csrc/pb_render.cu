// psdr-b200: wavefront kernels of the interior integral (primal): ray generation + primary trace, per-event shade
// (sample BSDF / emitter, emit rays), generic trace, per-event resolve (MIS-weighted contributions, throughput update,
// film accumulation).
//
// Restates Integrator::__render (src/integrator/integrator.cpp:64-95) and DirectIntegrator::__Li
// (src/integrator/direct.cpp:47-163) as a fixed pipeline over RayRec/HitRec wavefront buffers. The sampler is
// stateless: every kernel re-derives the lane's PCG32 stream from the global lane id and jumps to the position the
// reference's lock-step wavefront would be at, so no RNG state is stored and a render can be replayed.
#include <algorithm>

#include "pb_kernels.h"
#include "pb_shade.cuh"
#include "pb_trace.cuh"
#include "pb_trace2.cuh"
#include "pb_sortkey.cuh"
#include "pb_wavefront.cuh"

namespace pb {

__global__ void __launch_bounds__(128) k_trace(const BvhNode *__restrict__ nodes, const LeafTri *__restrict__ leaf, long long n,
                                               const RayRec *__restrict__ rays, HitRec *__restrict__ hits, float *__restrict__ t_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *rp = reinterpret_cast<const float4 *>(rays + i);
    const float4 a = ldg4(rp), b = ldg4(rp + 1);
    const Hit h = trace_closest(nodes, leaf, f3(a), f3(b), a.w);
    reinterpret_cast<float4 *>(hits)[i] = make_float4(__int_as_float(h.tri), __int_as_float(h.shape), h.u, h.v);
    if (t_out) t_out[i] = h.t;
}

// integrator.cpp:76-85 + the first ray launch of direct.cpp:48
__global__ void __launch_bounds__(128) k_primary(RenderParams P, HitRec *__restrict__ hit0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int pix;
    const long long lane = global_lane(P, i, pix);
    Rng rng = make_rng(P, lane, P.jump0);
    const float2 j = rng.next_2d();
    float sx, sy;
    lane_pixel_sample(P, pix, j, sx, sy);
    float3 o, d;
    sample_primary_ray(P.cam, sx, sy, o, d);
    // camera rays in lane order are coherent (the samples of a pixel are neighbours): one ray per thread over the compact nodes
    // (their slack covers the camera positions: pb_capi.cu configure)
    const Hit h = trace_closest_c(P.S.nodes_c, P.S.leaf, o, d, INFINITY, 0.f);
    reinterpret_cast<float4 *>(hit0)[i] = make_float4(__int_as_float(h.tri), __int_as_float(h.shape), h.u, h.v);
}

// sample the connections of one scattering event and emit their rays (direct.cpp:69-76, 120-129)
template <bool SIMPLE, int EV>
__global__ void __launch_bounds__(256) k_shade(RenderParams P, BounceParams B, EventBuffers E) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    RayRec *__restrict__ rays_out = E.rays;
    const Vertex v = load_vertex<EV>(P, B, i, E);
    store_vertex_rec(E.pos, E.vb, E.vc, i, v.its, v.its.valid);   // k_resolve / k_adjoint_lin read the record instead of reconstructing the vertex again
    int pix_unused;
    Rng rng = make_rng(P, global_lane(P, i, pix_unused), B.jump);
    for (int j = 0; j < B.nb; ++j) {
        const float3 s3 = rng.next_3d();
        const BsdfSample bs = bsdf_sample<SIMPLE>(v.bsdf, v.its, s3, v.active);
        const bool a1 = v.active && bs.valid;
        const float3 wo_w = v.its.sh.to_world(bs.wo);
        store_ray(rays_out + (size_t)j * P.n + i, v.its.p, wo_w, a1 ? INFINITY : -1.f);
        if (SIMPLE) E.conn[(size_t)j * P.n + i] = make_float2(bs.wo.z, 0.f);
        if (E.keys) E.keys[(size_t)j * P.n + i] = (unsigned short)sort_key(v.its.p, a1 ? 1.f : -1.f, wo_w, B.sort_lo, B.sort_inv_ext, B.sort_mode);
    }
    for (int j = 0; j < B.nl; ++j) {
        const float2 s2 = rng.next_2d();
        const PositionSample ps = sample_emitter_position<SIMPLE>(P.S, v.its.p, s2, v.active);
        const bool a1 = v.active && ps.valid;
        float3 wo = ps.p - v.its.p;
        const float dist_sqr = squared_norm(wo);
        const float dist = safe_sqrt(dist_sqr);
        wo = wo / dist;
        if (SIMPLE) E.conn[(size_t)(B.nb + j) * P.n + i] = make_float2(ps.pdf, dist_sqr);
        // The connection is valid iff the closest hit lies beyond dist - ShadowEpsilon and is an emitter (direct.cpp:130-131):
        // nothing past the sampled point can change that, and any closer hit decides it, so the ray is bounded and
        // flagged as an occlusion query.
        // Every BSDF here is zero unless both cosines are positive (diffuse.cpp:28, roughconductor.cpp:43), so the connection
        // contributes exactly 0 (value and derivative) whatever the ray hits: such lanes are not traced.
        const bool lit = v.its.wi.z > 0.f && dot(wo, v.its.sh.n) > 0.f;
        store_ray(rays_out + (size_t)(B.nb + j) * P.n + i, v.its.p, wo, (a1 && lit) ? fmaf(dist, 1e-5f, dist) + 2e-3f : -1.f, dist - kShadowEpsilon);
        if (E.keys) E.keys[(size_t)(B.nb + j) * P.n + i] = (unsigned short)sort_key(v.its.p, (a1 && lit) ? 1.f : -1.f, wo, B.sort_lo, B.sort_inv_ext, B.sort_mode);
    }
}

// direct.cpp:77-113 (BSDF-sampled connections) and 130-158 (emitter-sampled connections) for one scattering event
template <int MINB, bool PREFETCH, bool SIMPLE, int EV>
__global__ void __launch_bounds__(256, MINB) k_resolve(RenderParams P, BounceParams B, EventBuffers E, float *__restrict__ film) {
    const HitRec *__restrict__ hits = E.hits;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = i < P.n;
    int pix = -1;
    float3 out = f3(0.f);
    if (in_range) {
        if (PREFETCH && !E.inv) prefetch_event_hits(P.S, hits, B.nb + B.nl, P.n, i);
        const long long lane = global_lane(P, i, pix);
        const Vertex v = load_vertex_rec(P, E, i);
        const Its &its = v.its;
        // SIMPLE (diffuse + area lights): k_shade left what this kernel needs of its samples (EventBuffers::conn, directions in `rays`);
        // otherwise the sampler is re-derived and the sampling repeated
        Rng rng = SIMPLE ? Rng() : make_rng(P, lane, B.jump);
        float3 L = f3(0.f), w_cont = f3(0.f);
        bool has_cont = false;
        // Diffuse scenes, AD formulation: L_k and w_k are linear in the vertex' reflectance, L_k = rho * A and w_k = rho * c
        // (f = rho cos_o / pi, diffuse.cpp:25-33; pdf and MIS weight do not depend on rho). A retained render keeps (A, c): the
        // reflectance adjoint of the event is then gL * A + gw * c (k_adjoint_lin) with no connection to reconstruct.
        const bool keep_lin = SIMPLE && ev_ad<EV>(B) && E.lin != nullptr;
        float3 lin_a = f3(0.f);
        float lin_c = 0.f;
        const float inv_nb = B.nb > 0 ? 1.f / (float)B.nb : 0.f, inv_nl = B.nl > 0 ? 1.f / (float)B.nl : 0.f;
        for (int j = 0; j < B.nb; ++j) {
            BsdfSample bs;
            if (SIMPLE) {   // diffuse.cpp:47-55: pdf = cos / pi, valid iff the incident side is the front
                const float cos_o = v.bsdf ? __ldg(&E.conn[(size_t)j * P.n + i].x) : 0.f;
                bs.wo = f3(0.f, 0.f, cos_o); bs.pdf = kInvPi * cos_o; bs.valid = v.bsdf && v.active && its.wi.z > 0.f;
            } else {
                const float3 s3 = rng.next_3d();
                bs = bsdf_sample<SIMPLE>(v.bsdf, its, s3, v.active);
            }
            bool a1 = v.active && bs.valid;
            const HitRec h1 = event_hit(E, j, P.n, i);
            const Its its1 = reconstruct_its(P.S, h1, its.p);
            a1 = a1 && its1.valid;
            const bool cont = a1;
            a1 = a1 && is_emitter(P.S, its1.shape);
            if (a1 || (B.carry && j == 0 && cont)) {
                float3 bsdf_val;
                float pdf0, dk = 0.f;   // dk: d(bsdf_val) / d(rho)
                if (ev_ad<EV>(B)) {   // direct.cpp:83-95
                    float3 wo = its1.p - its.p;
                    wo = wo / its1.t;
                    const float3 wo_l = its.sh.to_local(wo);
                    bsdf_val = bsdf_eval<SIMPLE>(v.bsdf, its, wo_l, true);
                    const float G = fabsf(dot(its1.n, -wo)) / sqr(its1.t);
                    pdf0 = bs.pdf * G;
                    bsdf_val = bsdf_val * (G / pdf0);
                    if (keep_lin && v.bsdf && its.wi.z > 0.f && wo_l.z > 0.f) dk = (kInvPi * wo_l.z) * (G / pdf0);
                } else {      // direct.cpp:96-106
                    const float3 d1 = SIMPLE ? f3(ldg4(reinterpret_cast<const float4 *>(E.rays + (size_t)j * P.n + i) + 1)) : its.sh.to_world(bs.wo);
                    bsdf_val = bsdf_eval<SIMPLE>(v.bsdf, its, bs.wo, true);
                    const float G = fabsf(dot(its1.n, -d1)) / sqr(its1.t);
                    pdf0 = bs.pdf * G;
                    bsdf_val = bsdf_val / bs.pdf;
                }
                if (a1) {
                    float weight = inv_nb;
                    if (B.nl > 0) weight *= mis_weight(pdf0, emitter_position_pdf<SIMPLE>(P.S, its.p, its1, true));
                    const float3 Le = emitter_Le<SIMPLE>(P.S, its1, true);
                    L += Le * bsdf_val * weight;
                    if (keep_lin) lin_a += Le * (dk * weight);
                }
                if (B.carry && j == 0 && cont) { w_cont = bsdf_val; has_cont = true; lin_c = dk; }
            }
        }
        for (int j = 0; j < B.nl; ++j) {
            PositionSample ps;
            float3 wo;
            float dist_sqr, dist;
            if (SIMPLE) {
                const float2 cr = __ldg(E.conn + (size_t)(B.nb + j) * P.n + i);
                ps.pdf = cr.x; ps.valid = v.active;
                dist_sqr = cr.y; dist = safe_sqrt(dist_sqr);
                wo = f3(ldg4(reinterpret_cast<const float4 *>(E.rays + (size_t)(B.nb + j) * P.n + i) + 1));
            } else {
                const float2 s2 = rng.next_2d();
                ps = sample_emitter_position<SIMPLE>(P.S, v.its.p, s2, v.active);
                wo = ps.p - its.p;
                dist_sqr = squared_norm(wo);
                dist = safe_sqrt(dist_sqr);
                wo = wo / dist;
            }
            bool a1 = v.active && ps.valid;
            const HitRec h1 = event_hit(E, B.nb + j, P.n, i);
            const Its its1 = reconstruct_its(P.S, h1, its.p);
            a1 = a1 && its1.valid && (its1.t > dist - kShadowEpsilon) && is_emitter(P.S, its1.shape);
            if (a1) {
                const float G = fabsf(dot(its1.n, -wo)) / dist_sqr;
                const float3 wo_local = its.sh.to_local(wo);
                float3 bsdf_val = bsdf_eval<SIMPLE>(v.bsdf, its, wo_local, true);
                const float pdf1 = bsdf_pdf<SIMPLE>(v.bsdf, its, wo_local, true) * G;
                bsdf_val = bsdf_val * (G / ps.pdf);
                float weight = inv_nl;
                if (B.nb > 0) weight *= mis_weight(ps.pdf, pdf1);
                const float3 Le = emitter_Le<SIMPLE>(P.S, its1, true);
                L += Le * bsdf_val * weight;
                if (keep_lin && v.bsdf && its.wi.z > 0.f && wo_local.z > 0.f) lin_a += Le * ((kInvPi * wo_local.z) * (G / ps.pdf) * weight);
            }
        }
        if (keep_lin) E.lin[i] = make_float4(lin_a.x, lin_a.y, lin_a.z, lin_c);
        float3 thr = f3(1.f), rad;
        if (ev_depth0<EV>(B)) {
            rad = B.hide_emitters ? f3(0.f) : emitter_Le<SIMPLE>(P.S, its, its.valid);   // direct.cpp:51
        } else {
            thr = f3(ldg4(E.thr_in + i)); rad = f3(E.rad[i]);
        }
        rad += thr * L;
        if (B.last) out = zero_nonfinite(rad) * P.inv_spp;   // integrator.cpp:87-91
        if (E.thr_out) {
            const float3 t2 = thr * w_cont;
            // A path without a continuation is dead. A zero-throughput path contributes nothing either, but its deeper
            // events still enter the derivative w.r.t. whatever made the throughput zero (d(rho X)/d rho = X at rho = 0),
            // so it is only dropped by renderC.
            const bool dead = !has_cont || (!ev_ad<EV>(B) && !(t2.x != 0.f || t2.y != 0.f || t2.z != 0.f));
            E.thr_out[i] = make_float4(t2.x, t2.y, t2.z, dead ? 1.f : 0.f);
        }
        if (E.rad) E.rad[i] = make_float4(rad.x, rad.y, rad.z, 0.f);
    }
    if (B.last && film) film_accumulate(film, pix, out);
}

// field.cpp:34-54
__global__ void __launch_bounds__(256) k_field(RenderParams P, int field, const HitRec *__restrict__ hit0, float *__restrict__ film, float4 *__restrict__ rad_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int pix = -1;
    float3 out = f3(0.f);
    if (i < P.n) {
        global_lane(P, i, pix);
        const Its its = reconstruct_its(P.S, load_hit(hit0 + i), transform_pos(P.cam.to_world, f3(0.f)));
        if (its.valid) {
            switch (field) {
                case FIELD_SILHOUETTE: out = f3(1.f); break;
                case FIELD_POSITION: out = its.p; break;
                case FIELD_DEPTH: out = f3(its.t); break;
                case FIELD_GEONORMAL: out = its.n; break;
                case FIELD_SHNORMAL: out = its.sh.n; break;
                default: out = f3(its.uv.x, its.uv.y, 0.f); break;
            }
        }
        if (rad_out) rad_out[i] = make_float4(out.x, out.y, out.z, 0.f);
        out = zero_nonfinite(out) * P.inv_spp;
    }
    if (film) film_accumulate(film, pix, out);
}

static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

int g_shade_tune = 0;
int g_shade_simple = 1;
// rays in lane order through the general kernel (any origin; pb_trace with a distance output, debug)
void launch_trace(cudaStream_t st, const SceneView &S, long long n, const RayRec *rays, HitRec *hits, float *t_out) {
    if (n > 0) k_trace<<<nblk(n, 128), 128, 0, st>>>(S.nodes, S.leaf, n, rays, hits, t_out);
}
void launch_primary(cudaStream_t st, const RenderParams &P, HitRec *hit0) {
    if (P.n > 0) k_primary<<<nblk(P.n, 128), 128, 0, st>>>(P, hit0);
}
void launch_shade(cudaStream_t st, const RenderParams &P, const BounceParams &B, const EventBuffers &E) {
    if (P.n <= 0) return;
    const unsigned g = nblk(P.n, 256);
    if (P.S.simple && g_shade_simple) {
        switch ((B.depth == 0 ? 1 : 0) | (B.ad ? 2 : 0)) {
            case 0: k_shade<true, 0><<<g, 256, 0, st>>>(P, B, E); break;
            case 1: k_shade<true, 1><<<g, 256, 0, st>>>(P, B, E); break;
            case 2: k_shade<true, 2><<<g, 256, 0, st>>>(P, B, E); break;
            default: k_shade<true, 3><<<g, 256, 0, st>>>(P, B, E); break;
        }
    } else k_shade<false, -1><<<g, 256, 0, st>>>(P, B, E);
}
void launch_resolve(cudaStream_t st, const RenderParams &P, const BounceParams &B, const EventBuffers &E, float *film) {
    if (P.n <= 0) return;
    const unsigned g = nblk(P.n, 256);
    if (P.S.simple && g_shade_simple) {   // diffuse BSDFs + area emitters only: the instantiation without rough-conductor / envmap code
        const int ev = (B.depth == 0 ? 1 : 0) | (B.ad ? 2 : 0);
        if (g_shade_tune == 6) {   // run-time event class (what the specialised instantiations are measured against)
            k_resolve<4, false, true, -1><<<g, 256, 0, st>>>(P, B, E, film);
        } else if (g_shade_tune == 2) {
            switch (ev) {
                case 0: k_resolve<3, false, true, 0><<<g, 256, 0, st>>>(P, B, E, film); break;
                case 1: k_resolve<3, false, true, 1><<<g, 256, 0, st>>>(P, B, E, film); break;
                case 2: k_resolve<3, false, true, 2><<<g, 256, 0, st>>>(P, B, E, film); break;
                default: k_resolve<3, false, true, 3><<<g, 256, 0, st>>>(P, B, E, film); break;
            }
        } else {
            switch (ev) {
                case 0: k_resolve<4, false, true, 0><<<g, 256, 0, st>>>(P, B, E, film); break;
                case 1: k_resolve<4, false, true, 1><<<g, 256, 0, st>>>(P, B, E, film); break;
                case 2: k_resolve<4, false, true, 2><<<g, 256, 0, st>>>(P, B, E, film); break;
                default: k_resolve<4, false, true, 3><<<g, 256, 0, st>>>(P, B, E, film); break;
            }
        }
        return;
    }
    switch (g_shade_tune) {   // debug: resident blocks per SM forced through the register cap / early prefetch of the connection hits
        case 1: k_resolve<2, true, false, -1><<<g, 256, 0, st>>>(P, B, E, film); break;    // + L1 prefetch of the connection hits: 27 % slower (L1 thrash)
        case 2: k_resolve<3, false, false, -1><<<g, 256, 0, st>>>(P, B, E, film); break;
        case 4: k_resolve<2, false, false, -1><<<g, 256, 0, st>>>(P, B, E, film); break;   // 127 registers, no spills, 25 % occupancy
        default: k_resolve<4, false, false, -1><<<g, 256, 0, st>>>(P, B, E, film); break;  // 64 registers + 360 B spills, 50 % occupancy: 5 % faster step
    }
}
void launch_field(cudaStream_t st, const RenderParams &P, int field, const HitRec *hit0, float *film, float4 *rad_out) {
    if (P.n > 0) k_field<<<nblk(P.n, 256), 256, 0, st>>>(P, field, hit0, film, rad_out);
}

}  // namespace pb
