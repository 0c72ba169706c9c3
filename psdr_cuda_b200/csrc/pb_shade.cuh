// psdr-b200: per-lane shading primitives of the interior integral (primal flavour).
//
// Restates, as device functions over the tables of pb_scene.cuh:
//   Scene::ray_intersect<false>            src/scene/scene.cpp:289-354   -> reconstruct_its
//   Bitmap::eval                           src/core/bitmap.cpp:43-89     -> tex_eval3 / tex_eval1
//   Diffuse::__eval/__sample/__pdf         src/bsdf/diffuse.cpp:25-82    -> bsdf_eval / bsdf_sample / bsdf_pdf
//   RoughConductor / GGX / fresnel         src/bsdf/{roughconductor,ggx}.cpp, include/psdr/utils.h:149-164
//   AreaLight::eval, Mesh::sample_position src/emitter/area.cpp:20-62, src/shape/mesh.cpp:306-342
//   Scene::sample_emitter_position / pdf   src/scene/scene.cpp:427-453
//   DiscreteDistribution::sample_reuse     src/core/pmf.cpp:30-50
//   PerspectiveCamera::sample_primary_ray  src/sensor/perspective.cpp:120-127
#pragma once
#include "pb_scene.cuh"

namespace pb {

struct Its {
    float3 p, n, wi;
    Frame sh;
    float2 uv;
    float t;
    int tri, shape;
    bool valid;
};

struct TriGeom { float3 p0, e1, e2, fn; float area; int mesh, flags; };

PB_D TriGeom load_tri_geom(const SceneView &S, int tri) {
    const float4 *q = reinterpret_cast<const float4 *>(S.tri + tri);
    const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2), q6 = ldg4(q + 6);
    TriGeom g;
    g.p0 = f3(q0); g.area = q0.w;
    g.e1 = f3(q1); g.mesh = __float_as_int(q1.w);
    g.e2 = f3(q2); g.flags = __float_as_int(q2.w);
    g.fn = f3(q6);
    return g;
}

// Scene::ray_intersect<false> (also the primal of the path-space flavour): scene.cpp:326-354
PB_D Its reconstruct_its(const SceneView &S, const HitRec &h, float3 ray_o) {
    Its its;
    its.valid = h.tri >= 0;
    its.tri = h.tri; its.shape = h.shape;
    if (!its.valid) { its.t = 0.f; its.p = its.n = its.wi = f3(0.f); its.uv = make_float2(0.f, 0.f); return its; }
    const float4 *q = reinterpret_cast<const float4 *>(S.tri + h.tri);
    const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2), q6 = ldg4(q + 6);
    const int flags = __float_as_int(q2.w);
    its.n = f3(q6);
    float3 sh_n = its.n;
    float4 q3, q4, q5;
    if (!(flags & 1) || (flags & 2)) { q3 = ldg4(q + 3); q4 = ldg4(q + 4); q5 = ldg4(q + 5); }
    if (!(flags & 1)) {
        const float3 n0 = f3(q3), n1 = f3(q4), n2 = f3(q5);
        sh_n = normalize(bilinear(n0, n1 - n0, n2 - n0, h.u, h.v));
    }
    its.p = bilinear(f3(q0), f3(q1), f3(q2), h.u, h.v);
    float3 dir = its.p - ray_o;
    its.t = norm(dir);
    dir = f3(div_rn(dir.x, its.t), div_rn(dir.y, its.t), div_rn(dir.z, its.t));
    its.sh = Frame(sh_n);
    its.wi = its.sh.to_local(-dir);
    if (flags & 2) {
        const float4 q7 = ldg4(q + 7);
        const float u0x = q3.w, u0y = q4.w, u1x = q5.w, u1y = q6.w, u2x = q7.x, u2y = q7.y;
        its.uv = make_float2(fma_rn(u1x - u0x, h.u, fma_rn(u2x - u0x, h.v, u0x)), fma_rn(u1y - u0y, h.u, fma_rn(u2y - u0y, h.v, u0y)));
    } else {
        its.uv = make_float2(0.f, 0.f);
    }
    return its;
}

// Scene::ray_intersect<true,false> — the solid-angle flavour renderD uses for the camera ray (scene.cpp:355-376):
// (u,v,t) re-derived from the ray by Möller–Trumbore, p = o + t d, wi = to_local(-d)
PB_D Its reconstruct_its_primary(const SceneView &S, const HitRec &h, float3 ray_o, float3 ray_d) {
    Its its;
    its.valid = h.tri >= 0;
    its.tri = h.tri; its.shape = h.shape;
    if (!its.valid) { its.t = 0.f; its.p = its.n = its.wi = f3(0.f); its.uv = make_float2(0.f, 0.f); return its; }
    const float4 *q = reinterpret_cast<const float4 *>(S.tri + h.tri);
    const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2), q6 = ldg4(q + 6);
    const int flags = __float_as_int(q2.w);
    float u, v, t;
    ray_intersect_triangle(f3(q0), f3(q1), f3(q2), ray_o, ray_d, u, v, t);
    its.n = f3(q6);
    float3 sh_n = its.n;
    float4 q3, q4, q5;
    if (!(flags & 1) || (flags & 2)) { q3 = ldg4(q + 3); q4 = ldg4(q + 4); q5 = ldg4(q + 5); }
    if (!(flags & 1)) {
        const float3 n0 = f3(q3), n1 = f3(q4), n2 = f3(q5);
        sh_n = normalize(bilinear(n0, n1 - n0, n2 - n0, u, v));
    }
    its.p = f3(fma_rn(ray_d.x, t, ray_o.x), fma_rn(ray_d.y, t, ray_o.y), fma_rn(ray_d.z, t, ray_o.z));
    its.t = t;
    its.sh = Frame(sh_n);
    its.wi = its.sh.to_local(-ray_d);
    if (flags & 2) {
        const float4 q7 = ldg4(q + 7);
        const float u0x = q3.w, u0y = q4.w, u1x = q5.w, u1y = q6.w, u2x = q7.x, u2y = q7.y;
        its.uv = make_float2(fma_rn(u1x - u0x, u, fma_rn(u2x - u0x, v, u0x)), fma_rn(u1y - u0y, u, fma_rn(u2y - u0y, v, u0y)));
    } else {
        its.uv = make_float2(0.f, 0.f);
    }
    return its;
}

// ---- textures (bitmap.cpp:43-89) ------------------------------------------------------------------------
struct TexTap { int idx; float w0x, w1x, w0y, w1y; bool constant; };
PB_D TexTap tex_tap(const TexRef &t, float2 uv, bool flip_v = true) {
    TexTap r;
    r.constant = (t.w == 1 && t.h == 1);
    if (r.constant) { r.idx = 0; r.w0x = r.w0y = 1.f; r.w1x = r.w1y = 0.f; return r; }
    if (flip_v) uv.y = -uv.y;
    uv.x -= floorf(uv.x); uv.y -= floorf(uv.y);
    uv.x *= (float)(t.w - 1); uv.y *= (float)(t.h - 1);
    int px = (int)floorf(uv.x), py = (int)floorf(uv.y);
    r.w1x = uv.x - (float)px; r.w1y = uv.y - (float)py; r.w0x = 1.f - r.w1x; r.w0y = 1.f - r.w1y;
    px = min(px, t.w - 2); py = min(py, t.h - 2);
    r.idx = py * t.w + px;
    return r;
}
PB_D float tex_fetch(const TexRef &t, const TexTap &tap, int ch) {
    const float *d = t.data;
    if (tap.constant) return __ldg(d + ch);
    const int c = t.c, i = tap.idx;
    const float v00 = __ldg(d + i * c + ch), v10 = __ldg(d + (i + 1) * c + ch), v01 = __ldg(d + (i + t.w) * c + ch), v11 = __ldg(d + (i + t.w + 1) * c + ch);
    const float a = fmaf(tap.w0x, v00, tap.w1x * v10), b = fmaf(tap.w0x, v01, tap.w1x * v11);
    return fmaf(tap.w0y, a, tap.w1y * b);
}
PB_D float3 tex_eval3(const TexRef &t, float2 uv, bool flip_v = true) {
    if (t.w == 1 && t.h == 1) return f3(__ldg(t.data), __ldg(t.data + 1), __ldg(t.data + 2));
    TexTap tap = tex_tap(t, uv, flip_v);
    return f3(tex_fetch(t, tap, 0), tex_fetch(t, tap, 1), tex_fetch(t, tap, 2));
}
PB_D float tex_eval1(const TexRef &t, float2 uv, bool flip_v = true) {
    if (t.w == 1 && t.h == 1) return __ldg(t.data);
    TexTap tap = tex_tap(t, uv, flip_v);
    return tex_fetch(t, tap, 0);
}

// ---- discrete distribution (pmf.cpp:17-50) ---------------------------------------------------------------
// first i in [0, n-1] with cmf[i] >= x (n-1 if none)
PB_D int cmf_search(const float *__restrict__ cmf, int n, float x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cmf + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}
PB_D int sample_reuse(const float *__restrict__ cmf, const float *__restrict__ pmf, int n, float sum, float &u, float &pdf) {
    if (n == 1) { pdf = 1.f; return 0; }
    u *= sum;
    const int idx = cmf_search(cmf, n, u);
    if (idx > 0) u -= __ldg(cmf + idx - 1);
    const float p = __ldg(pmf + idx);
    if (p > 0.f) u = div_rn(u, p);
    u = fminf(fmaxf(u, 0.f), 1.f);
    pdf = div_rn(p, sum);
    return idx;
}

// ---- BSDFs ---------------------------------------------------------------------------------------------------
namespace ggx {
PB_D float eval(float au, float av, float3 m) {   // ggx.cpp:15-34
    const float result = 1.f / (kPi * au * av * sqr(sqr(m.x / au) + sqr(m.y / av) + sqr(m.z)));
    return result * m.z > 1e-5f ? result : 0.f;
}
PB_D float smith_g1(float au, float av, float3 v, float3 m) {   // ggx.cpp:79-93
    const float xy_alpha_2 = sqr(au * v.x) + sqr(av * v.y);
    const float tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z);
    float result = 2.f / (1.f + sqrtf(1.f + tan_theta_alpha_2));
    if (xy_alpha_2 == 0.f) result = 1.f;
    if (dot(v, m) * v.z <= 0.f) result = 0.f;
    return result;
}
PB_D float2 sample_visible_11(float cos_theta_i, float sx, float sy) {   // ggx.cpp:96-105
    float2 p = square_to_uniform_disk_concentric(sx, sy);
    const float s = .5f * (1.f + cos_theta_i);
    const float a = safe_sqrt(1.f - sqr(p.x));
    p.y = fmaf(p.y, s, fmaf(-a, s, a));
    const float x = p.x, y = p.y, z = safe_sqrt(1.f - fmaf(p.x, p.x, p.y * p.y));
    const float sin_theta_i = safe_sqrt(1.f - sqr(cos_theta_i));
    const float nrm = 1.f / fmaf(sin_theta_i, y, cos_theta_i * z);
    return make_float2(fmaf(cos_theta_i, y, -(sin_theta_i * z)) * nrm, x * nrm);
}
PB_D float3 sample(float au, float av, float3 wi, float sx, float sy) {   // ggx.cpp:37-76
    const float3 wi_p = normalize(f3(au * wi.x, av * wi.y, wi.z));
    const float sin_theta_2 = fmaf(wi_p.x, wi_p.x, sqr(wi_p.y));
    const float inv_sin_theta = 1.f / sqrtf(sin_theta_2);
    const bool degenerate = fabsf(sin_theta_2) <= 4.f * kEpsilon;   // frame.h:103-117
    const float sin_phi = degenerate ? 0.f : fminf(fmaxf(wi_p.y * inv_sin_theta, -1.f), 1.f);
    const float cos_phi = degenerate ? 1.f : fminf(fmaxf(wi_p.x * inv_sin_theta, -1.f), 1.f);
    float2 slope = sample_visible_11(wi_p.z, sx, sy);
    slope = make_float2(fmaf(cos_phi, slope.x, -(sin_phi * slope.y)) * au, fmaf(sin_phi, slope.x, cos_phi * slope.y) * av);
    return normalize(f3(-slope.x, -slope.y, 1.f));
}
}  // namespace ggx

PB_D float3 fresnel_conductor(float3 eta_r, float3 eta_i, float cos_theta_i) {   // utils.h:149-164
    const float c2 = sqr(cos_theta_i), s2 = 1.f - c2, s4 = sqr(s2);
    float out[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float er = getc(eta_r, k), ei = getc(eta_i, k);
        const float temp_1 = sqr(er) - sqr(ei) - s2;
        const float a_2_pb_2 = safe_sqrt(sqr(temp_1) + 4.f * sqr(ei * er));
        const float a = safe_sqrt(.5f * (a_2_pb_2 + temp_1));
        const float term_1 = a_2_pb_2 + c2, term_2 = 2.f * cos_theta_i * a;
        const float r_s = (term_1 - term_2) / (term_1 + term_2);
        const float term_3 = a_2_pb_2 * c2 + s4, term_4 = term_2 * s2;
        const float r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
        out[k] = .5f * (r_s + r_p);
    }
    return f3(out[0], out[1], out[2]);
}

PB_D const BsdfRec *its_bsdf(const SceneView &S, const Its &its) {
    if (!its.valid) return nullptr;
    const int b = S.meshes[its.shape].bsdf;
    return b >= 0 ? S.bsdfs + b : nullptr;
}

// SIMPLE (template flag of the shading functions and of the kernels that call them): the scene has only diffuse BSDFs and area
// emitters — the host knows at configure — so the rough-conductor / environment-map code is compiled out of that instantiation
// (it costs registers even when it never runs).
template <bool SIMPLE = false>
PB_D float3 bsdf_eval(const BsdfRec *b, const Its &its, float3 wo, bool active) {
    if (!active || !b) return f3(0.f);
    const float cos_i = its.wi.z, cos_o = wo.z;
    if (!(cos_i > 0.f && cos_o > 0.f)) return f3(0.f);
    if (SIMPLE || b->type == BSDF_DIFFUSE)   // diffuse.cpp:25-33
        return tex_eval3(b->tex[TEX_REFLECTANCE], its.uv) * kInvPi * cos_o;
    // roughconductor.cpp:40-56
    const float au = tex_eval1(b->tex[TEX_ALPHA_U], its.uv), av = tex_eval1(b->tex[TEX_ALPHA_V], its.uv);
    const float3 H = normalize(wo + its.wi);
    const float D = ggx::eval(au, av, H);
    if (D == 0.f) return f3(0.f);
    const float G = ggx::smith_g1(au, av, its.wi, H) * ggx::smith_g1(au, av, wo, H);
    const float result = D * G / (4.f * its.wi.z);
    const float3 F = fresnel_conductor(tex_eval3(b->tex[TEX_ETA], its.uv), tex_eval3(b->tex[TEX_K], its.uv), dot(its.wi, H));
    return F * result * tex_eval3(b->tex[TEX_SPECULAR], its.uv);
}
template <bool SIMPLE = false>
PB_D float bsdf_pdf(const BsdfRec *b, const Its &its, float3 wo, bool active) {
    if (!b) return 0.f;
    if (SIMPLE || b->type == BSDF_DIFFUSE) {   // diffuse.cpp:69-82
        if (!active || !(its.wi.z > 0.f && wo.z > 0.f)) return 0.f;
        return kInvPi * wo.z;
    }
    // roughconductor.cpp:60-75 — the mask it computes is never applied to the result
    const float3 m = normalize(wo + its.wi);
    const float au = tex_eval1(b->tex[TEX_ALPHA_U], its.uv), av = tex_eval1(b->tex[TEX_ALPHA_V], its.uv);
    return ggx::eval(au, av, m) * ggx::smith_g1(au, av, its.wi, m) / (4.f * its.wi.z);
}
struct BsdfSample { float3 wo; float pdf; bool valid; };
template <bool SIMPLE = false>
PB_D BsdfSample bsdf_sample(const BsdfRec *b, const Its &its, float3 smp, bool active) {
    BsdfSample bs;
    bs.wo = f3(0.f); bs.pdf = 0.f; bs.valid = false;
    if (!b) return bs;
    if (SIMPLE || b->type == BSDF_DIFFUSE) {   // diffuse.cpp:47-55 — consumes tail<2>(sample)
        bs.wo = square_to_cosine_hemisphere(smp.y, smp.z);
        bs.pdf = kInvPi * bs.wo.z;
        bs.valid = active && its.wi.z > 0.f;
        return bs;
    }
    // roughconductor.cpp:79-93 — consumes head<2>(sample)
    const float au = tex_eval1(b->tex[TEX_ALPHA_U], its.uv), av = tex_eval1(b->tex[TEX_ALPHA_V], its.uv);
    const float3 m = ggx::sample(au, av, its.wi, smp.x, smp.y);
    const float two_dot = 2.f * dot(its.wi, m);
    bs.wo = f3(fmaf(m.x, two_dot, -its.wi.x), fmaf(m.y, two_dot, -its.wi.y), fmaf(m.z, two_dot, -its.wi.z));
    bs.pdf = bsdf_pdf<false>(b, its, bs.wo, active);
    bs.valid = active && its.wi.z > 0.f && bs.pdf != 0.f && bs.wo.z > 0.f;
    return bs;
}

// ---- emitters --------------------------------------------------------------------------------------------------
PB_D bool is_emitter(const SceneView &S, int shape) { return shape >= 0 && S.meshes[shape].emitter >= 0; }

PB_D float safe_acos(float x) { return acosf(fminf(fmaxf(x, -1.f), 1.f)); }

// EnvironmentMap::eval_direction (envmap.cpp:42-58): world direction -> lat-long lookup (no v flip) * scale
PB_D float3 env_eval_direction(const EmitterRec &em, float3 wi_world) {
    const float3 v = transform_dir(em.env_from_world, wi_world);
    float2 uv = make_float2(atan2f(v.x, -v.z) * kInvTwoPi, safe_acos(v.y) * kInvPi);
    uv.x -= floorf(uv.x); uv.y -= floorf(uv.y);
    return tex_eval3(em.env_radiance, uv, false) * em.env_scale;
}

template <bool SIMPLE = false>
PB_D float3 emitter_Le(const SceneView &S, const Its &its, bool active) {   // intersection.h:36-38
    if (!active || !its.valid) return f3(0.f);
    const int e = S.meshes[its.shape].emitter;
    if (e < 0) return f3(0.f);
    const EmitterRec &em = S.emitters[e];
    if (SIMPLE || em.type == EMITTER_AREA) return its.wi.z > 0.f ? em.radiance : f3(0.f);   // area.cpp:20-29
    return env_eval_direction(em, -its.sh.to_world(its.wi));                      // envmap.cpp:29-39
}

// utils.h:129-145: exit point of a ray that starts inside the scene box
PB_D void ray_intersect_scene_aabb(float3 o, float3 d, float3 lo, float3 hi, float &t, float3 &n, float &G) {
    const float3 t1 = f3((lo.x - o.x) / d.x, (lo.y - o.y) / d.y, (lo.z - o.z) / d.z), t2 = f3((hi.x - o.x) / d.x, (hi.y - o.y) / d.y, (hi.z - o.z) / d.z);
    const float3 t2p = f3(t1.x > t2.x ? t1.x : t2.x, t1.y > t2.y ? t1.y : t2.y, t1.z > t2.z ? t1.z : t2.z);
    int idx = 0;
    t = t2p.x;
    if (t2p.y < t) { t = t2p.y; idx = 1; }
    if (t2p.z < t) { t = t2p.z; idx = 2; }
    n = f3(0.f);
    const float sg = -copysignf(1.f, getc(d, idx));
    if (idx == 0) n.x = sg; else if (idx == 1) n.y = sg; else n.z = sg;
    G = dot(n, -d) * (1.f / sqr(t));
}

struct PositionSample { float3 p, n; float pdf; int tri; float s, t; bool valid; };
template <bool SIMPLE = false>
PB_D PositionSample sample_emitter_position(const SceneView &S, float3 ref_p, float2 smp, bool active) {   // scene.cpp:427-447
    PositionSample r;
    int ei = 0;
    float emitter_pdf = 1.f;
    if (S.num_emitters > 1) ei = sample_reuse(S.emitter_cmf, S.emitter_pmf, S.num_emitters, S.emitter_sum, smp.y, emitter_pdf);
    const EmitterRec &em = S.emitters[ei];
    if (SIMPLE || em.type == EMITTER_AREA) {   // mesh.cpp:306-330
        float face_pdf;
        const int f = sample_reuse(em.face_cmf, em.face_pmf, em.num_faces, em.face_sum, smp.x, face_pdf);
        const float2 st = square_to_uniform_triangle(smp.x, smp.y);
        r.tri = em.face_offset + f;
        r.s = st.x; r.t = st.y;
        const TriGeom g = load_tri_geom(S, r.tri);
        r.p = bilinear(g.p0, g.e1, g.e2, st.x, st.y);
        r.n = g.fn;
        r.pdf = S.meshes[em.mesh].inv_total_area * emitter_pdf;
    } else {                          // envmap.cpp:72-111
        float pdf;
        const int cell = sample_reuse(em.env_cmf, em.env_pmf, em.env_cells, em.env_sum, smp.y, pdf);
        const int cy = cell % em.env_res_y, cx = cell / em.env_res_y;
        const float sx = (smp.x + (float)cx) * (1.f / (float)em.env_res_x), sy = (smp.y + (float)cy) * (1.f / (float)em.env_res_y);
        pdf *= (float)em.env_cells;
        const float theta = sy * kPi, phi = sx * kTwoPi;
        float st, ct, sp, cp;
        sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
        float3 d = f3(sp * st, ct, -(cp * st));   // sphdir(theta, phi) = (cp st, sp st, ct) permuted to (y, z, -x)
        const float inv_sin_theta = 1.f / safe_sqrt(fmaxf(sqr(d.x) + sqr(d.z), sqr(kEpsilon)));
        if (pdf > kEpsilon) pdf *= inv_sin_theta * (.5f / sqr(kPi));
        d = transform_dir(em.env_to_world, d);
        float t, G;
        ray_intersect_scene_aabb(ref_p, d, em.env_lower, em.env_upper, t, r.n, G);
        r.p = ref_p + d * t;
        r.pdf = pdf * G * emitter_pdf;
        r.tri = -1; r.s = r.t = 0.f;
    }
    r.valid = active;
    return r;
}
template <bool SIMPLE = false>
PB_D float emitter_position_pdf(const SceneView &S, float3 ref_p, const Its &its, bool active) {   // scene.cpp:451-453
    if (!active || !its.valid) return 0.f;
    const MeshRec &m = S.meshes[its.shape];
    if (m.emitter < 0) return 0.f;
    const EmitterRec &em = S.emitters[m.emitter];
    if (SIMPLE || em.type == EMITTER_AREA) return em.sampling_weight * m.inv_total_area;   // area.cpp:58-62
    // envmap.cpp:125-143 (no sampling_weight factor)
    float3 d = its.p - ref_p;
    const float dist2 = squared_norm(d);
    d = d / safe_sqrt(dist2);
    const float G = fabsf(dot(d, its.n)) / dist2;
    d = transform_dir(em.env_from_world, d);
    const float factor = G * (1.f / safe_sqrt(fmaxf(sqr(d.x) + sqr(d.z), sqr(kEpsilon)))) * (.5f / sqr(kPi));
    float u = atan2f(d.x, -d.z) * kInvTwoPi, v = safe_acos(d.y) * kInvPi;
    u -= floorf(u); v -= floorf(v);
    const int ix = (int)floorf(u * (float)em.env_res_x), iy = (int)floorf(v * (float)em.env_res_y);
    if (!(ix >= 0 && ix < em.env_res_x && iy >= 0 && iy < em.env_res_y)) return 0.f;
    return (__ldg(em.env_pmf + ix * em.env_res_y + iy) / em.env_sum) * (float)em.env_cells * factor;
}

// ---- sensor (perspective.cpp:120-127) -------------------------------------------------------------------------------
PB_D void sample_primary_ray(const SensorRec &cam, float sx, float sy, float3 &o, float3 &d) {
    const float3 dc = normalize(transform_pos(cam.sample_to_camera, f3(sx, sy, 0.f)));
    o = transform_pos(cam.to_world, f3(0.f));
    d = transform_dir(cam.to_world, dc);
}

PB_D float mis_weight(float pdf1, float pdf2) { const float w1 = sqr(pdf1), w2 = sqr(pdf2); return w1 / (w1 + w2); }   // direct.cpp:17-21

}  // namespace pb
