// psdr-b200: derivative of the rough-conductor scattering event (RoughConductor / GGXDistribution / fresnel in their
// ad = true flavour: src/bsdf/roughconductor.cpp:40-93, src/bsdf/ggx.cpp:9-105, include/psdr/utils.h:149-164) as it appears in
// DirectIntegrator::__Li<true> (src/integrator/direct.cpp:67-159).
//
// What is attached there (and therefore differentiated here): the BSDF value f(wi, wo; alpha, eta, k, spec), the pdf of the
// BSDF-sampled direction *including the dependence of the sampled direction on alpha and wi* (bs.pdf = pdf(its, bs.wo), with
// bs.wo = sample(its, u)), pdf1 of the emitter-sampled direction, and through them both MIS weights; the geometric term inside
// pdf0 / pdf1 is detached (direct.cpp:93,147). The functions are templates over the scalar T = Dual<N> (pb_dual.cuh).
#pragma once
#include "pb_dual.cuh"
#include "pb_shade.cuh"

namespace pb {
namespace rc {

template <class T> PB_HD T ggx_eval(const T &au, const T &av, const V3<T> &m) {   // ggx.cpp:15-34
    const T result = 1.f / ((au * av) * dsqr(dsqr(m.x / au) + dsqr(m.y / av) + dsqr(m.z)) * kPi);
    return val(result) * val(m.z) > 1e-5f ? result : T(0.f);
}
template <class T> PB_HD T smith_g1(const T &au, const T &av, const V3<T> &v, const V3<T> &m) {   // ggx.cpp:79-93
    const T xy_alpha_2 = dsqr(au * v.x) + dsqr(av * v.y);
    T result = 2.f / (1.f + dsqrt(1.f + xy_alpha_2 / dsqr(v.z)));
    if (val(xy_alpha_2) == 0.f) result = T(1.f);
    if (val(vdot(v, m)) * val(v.z) <= 0.f) result = T(0.f);
    return result;
}
// ggx.cpp:96-105; `p` is the concentric-disk image of the (constant) random numbers
template <class T> PB_HD void sample_visible_11(const T &cos_theta_i, float2 p, T &slope_x, T &slope_y) {
    const T s = (1.f + cos_theta_i) * .5f;
    const float a = safe_sqrt(1.f - sqr(p.x));
    const T py = a + s * (p.y - a);   // lerp(a, p.y, s)
    const T z = dsafe_sqrt(1.f - (dsqr(py) + sqr(p.x)));
    const T sin_theta_i = dsafe_sqrt(1.f - dsqr(cos_theta_i));
    const T nrm = 1.f / (sin_theta_i * py + cos_theta_i * z);
    slope_x = (cos_theta_i * py - sin_theta_i * z) * nrm;
    slope_y = nrm * p.x;
}
template <class T> PB_HD V3<T> ggx_sample(const T &au, const T &av, const V3<T> &wi, float2 disk) {   // ggx.cpp:37-76
    const V3<T> wi_p = vnormalize(V3<T>(au * wi.x, av * wi.y, wi.z));
    const T sin_theta_2 = dsqr(wi_p.x) + dsqr(wi_p.y);
    const bool degenerate = fabsf(val(sin_theta_2)) <= 4.f * kEpsilon;   // frame.h:103-117
    T sin_phi(0.f), cos_phi(1.f);
    if (!degenerate) {
        const T inv_sin_theta = 1.f / dsqrt(sin_theta_2);
        sin_phi = dclamp(wi_p.y * inv_sin_theta, -1.f, 1.f);
        cos_phi = dclamp(wi_p.x * inv_sin_theta, -1.f, 1.f);
    }
    T sx, sy;
    sample_visible_11<T>(wi_p.z, disk, sx, sy);
    const T slx = (cos_phi * sx - sin_phi * sy) * au, sly = (sin_phi * sx + cos_phi * sy) * av;
    return vnormalize(V3<T>(-slx, -sly, T(1.f)));
}
template <class T> PB_HD T fresnel1(const T &eta_r, const T &eta_i, const T &cos_theta_i) {   // utils.h:149-164, one channel
    const T c2 = dsqr(cos_theta_i), s2 = 1.f - c2, s4 = dsqr(s2);
    const T temp_1 = dsqr(eta_r) - dsqr(eta_i) - s2;
    const T a_2_pb_2 = dsafe_sqrt(dsqr(temp_1) + dsqr(eta_i * eta_r) * 4.f);
    const T a = dsafe_sqrt((a_2_pb_2 + temp_1) * .5f);
    const T term_1 = a_2_pb_2 + c2, term_2 = cos_theta_i * a * 2.f;
    const T r_s = (term_1 - term_2) / (term_1 + term_2);
    const T term_3 = a_2_pb_2 * c2 + s4, term_4 = term_2 * s2;
    const T r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
    return (r_s + r_p) * .5f;
}
// D * G / (4 cos_i): the scalar part of RoughConductor::__eval (roughconductor.cpp:40-56), 0 outside its masks
template <class T> PB_HD T eval_scalar(const T &au, const T &av, const V3<T> &wi, const V3<T> &wo, const V3<T> &H) {
    if (!(val(wi.z) > 0.f && val(wo.z) > 0.f)) return T(0.f);
    const T D = ggx_eval<T>(au, av, H);
    if (val(D) == 0.f) return T(0.f);
    return D * smith_g1<T>(au, av, wi, H) * smith_g1<T>(au, av, wo, H) / (wi.z * 4.f);
}
template <class T> PB_HD T pdf(const T &au, const T &av, const V3<T> &wi, const V3<T> &wo) {   // roughconductor.cpp:60-75 (mask unused)
    const V3<T> m = vnormalize(wo + wi);
    return ggx_eval<T>(au, av, m) * smith_g1<T>(au, av, wi, m) / (wi.z * 4.f);
}
// pdf of the BSDF-sampled direction as a function of (alpha, wi): roughconductor.cpp:79-93
template <class T> PB_HD T sampled_pdf(const T &au, const T &av, const V3<T> &wi, float2 disk) {
    const V3<T> m = ggx_sample<T>(au, av, wi, disk);
    const T two_dot = vdot(wi, m) * 2.f;
    const V3<T> wo(m.x * two_dot - wi.x, m.y * two_dot - wi.y, m.z * two_dot - wi.z);
    return pdf<T>(au, av, wi, wo);
}

struct Tex { float au, av; float3 eta, k, spec; };
PB_D Tex load_tex(const BsdfRec *b, float2 uv) {
    Tex t;
    t.au = tex_eval1(b->tex[TEX_ALPHA_U], uv); t.av = tex_eval1(b->tex[TEX_ALPHA_V], uv);
    t.eta = tex_eval3(b->tex[TEX_ETA], uv); t.k = tex_eval3(b->tex[TEX_K], uv); t.spec = tex_eval3(b->tex[TEX_SPECULAR], uv);
    return t;
}
PB_D bool wants_tex_grad(const BsdfRec *b) {
    return b && b->type == BSDF_ROUGHCONDUCTOR &&
           (b->tex[TEX_ALPHA_U].grad || b->tex[TEX_ALPHA_V].grad || b->tex[TEX_ETA].grad || b->tex[TEX_K].grad || b->tex[TEX_SPECULAR].grad);
}

// local gradient of one event's loss-weighted value w.r.t. the 11 texture-evaluated BSDF parameters
struct TexGrad {
    float au, av;
    float3 eta, k, spec;
    PB_HD TexGrad() : au(0.f), av(0.f), eta(f3(0.f)), k(f3(0.f)), spec(f3(0.f)) {}
    PB_HD bool finite() const { return isfinite(au) && isfinite(av) && finite3(eta) && finite3(k) && finite3(spec); }
    PB_HD void add(const TexGrad &o) { au += o.au; av += o.av; eta += o.eta; k += o.k; spec += o.spec; }
};

// Shared tail of both branches: value = sum_ch coef_ch(alpha) * spec_ch * F_ch(eta_ch, k_ch) with
// coef_ch = base(alpha) * (wA(alpha) * gA_ch + gB_ch); base and wA are Dual<2> in (alpha_u, alpha_v).
PB_HD void finish_tex_grad(const Tex &t, float cos_h, const Dual<2> &base, const Dual<2> &wA, float3 gA, float3 gB, TexGrad &g) {
    typedef Dual<2> D2;
    float F[3], dF_eta[3], dF_k[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const D2 f = fresnel1<D2>(D2::seed(getc(t.eta, c), 0), D2::seed(getc(t.k, c), 1), D2(cos_h));
        F[c] = f.v; dF_eta[c] = f.d[0]; dF_k[c] = f.d[1];
    }
    const float ga[3] = {gA.x, gA.y, gA.z}, gb[3] = {gB.x, gB.y, gB.z}, sp[3] = {t.spec.x, t.spec.y, t.spec.z};
    D2 total(0.f);
    float g_spec[3], g_eta[3], g_k[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const D2 coef = base * (wA * ga[c] + gb[c]);
        total += coef * (sp[c] * F[c]);
        g_spec[c] = coef.v * F[c];
        g_eta[c] = coef.v * sp[c] * dF_eta[c];
        g_k[c] = coef.v * sp[c] * dF_k[c];
    }
    TexGrad r;
    r.au = total.d[0]; r.av = total.d[1];
    r.spec = f3(g_spec[0], g_spec[1], g_spec[2]); r.eta = f3(g_eta[0], g_eta[1], g_eta[2]); r.k = f3(g_k[0], g_k[1], g_k[2]);
    if (r.finite()) g.add(r);   // degenerate sample (zero pdf, grazing direction): no contribution rather than a poisoned gradient
}

// BSDF-sampled connection (direct.cpp:67-113 with ad = true): value = f * G J / pdf0 * (Le weight [emitter hit] + S_next [continuation]),
// pdf0 = bs.pdf * detach(G), weight = mis(pdf0, p_em) / nb. gA = dL/d(radiance) * T_k * Le (zero if the hit is no emitter),
// gB = dL/d(radiance) * T_k * S_{k+1} (zero without continuation).
PB_HD void bsdf_branch_tex_grad(const Tex &t, float3 wi, float3 wo_l, float3 s3, float G_geo, float p_em, bool use_mis, float inv_nb,
                               float3 gA, float3 gB, TexGrad &g) {
    typedef Dual<2> D2;
    const D2 au = D2::seed(t.au, 0), av = D2::seed(t.av, 1);
    const V3<D2> wi_d(wi), wo_d(wo_l);
    const float3 Hf = normalize(wo_l + wi);
    const D2 R = eval_scalar<D2>(au, av, wi_d, wo_d, V3<D2>(Hf));
    if (R.v == 0.f) return;
    const float2 disk = square_to_uniform_disk_concentric(s3.x, s3.y);   // roughconductor.cpp:87 uses head<2>(sample)
    const D2 pdf0 = sampled_pdf<D2>(au, av, wi_d, disk) * G_geo;
    D2 wA(inv_nb);
    if (use_mis) { const D2 w1 = dsqr(pdf0); wA = w1 / (w1 + sqr(p_em)) * inv_nb; }
    finish_tex_grad(t, dot(wi, Hf), R * G_geo / pdf0, wA, gA, gB, g);
}
// emitter-sampled connection (direct.cpp:119-159): value = f * G J / ps.pdf * Le * mis(ps.pdf, pdf1 * detach(G)) / nl
PB_HD void light_branch_tex_grad(const Tex &t, float3 wi, float3 wo_l, float G_geo, float ps_pdf, bool use_mis, float inv_nl, float3 gA, TexGrad &g) {
    typedef Dual<2> D2;
    const D2 au = D2::seed(t.au, 0), av = D2::seed(t.av, 1);
    const V3<D2> wi_d(wi), wo_d(wo_l);
    const float3 Hf = normalize(wo_l + wi);
    const D2 R = eval_scalar<D2>(au, av, wi_d, wo_d, V3<D2>(Hf));
    if (R.v == 0.f) return;
    D2 wA(inv_nl);
    if (use_mis) { const D2 pdf1 = pdf<D2>(au, av, wi_d, wo_d) * G_geo; const float w1 = sqr(ps_pdf); wA = w1 / (w1 + dsqr(pdf1)) * inv_nl; }
    finish_tex_grad(t, dot(wi, Hf), R * (G_geo / ps_pdf), wA, gA, f3(0.f), g);
}

// ---- geometry: local gradient of one connection w.r.t. the points and normals it is built from ----------------------------------
// inputs x[15] = (p, sh_n, a, q, n_q): the shaded point and its shading normal, the origin `a` of the ray that found it
// (scene.cpp:343-349: wi = to_local(-(p - a)/|p - a|); at the camera vertex `a` is the ray direction, wi = to_local(-ray.d),
// scene.cpp:368, whose adjoint only matters for the sensor pose), the far end
// q of the connection and the geometric normal there. value = f(wi, wo) * G / pdf * (Le weight + S_next) as in direct.cpp:83-113 /
// 133-158; the Jacobian J = A/detach(A) of q multiplies it (value 1), so its adjoint is the value itself.
template <class T> struct FrameT {   // frame.h:9-52
    V3<T> s, t, n;
    PB_HD explicit FrameT(const V3<T> &v) : n(v) {
        const float sg = copysignf(1.f, val(v.z));
        const bool neg = signbit(val(v.z));
        const T a = -1.f / (sg + v.z);
        const T b = v.x * v.y * a;
        const T sx = dsqr(v.x) * a;
        s = V3<T>((neg ? -sx : sx) + 1.f, neg ? -b : b, neg ? v.x : -v.x);
        t = V3<T>(b, sg + dsqr(v.y) * a, -v.y);
    }
    PB_HD V3<T> to_local(const V3<T> &v) const { return V3<T>(vdot(v, s), vdot(v, t), vdot(v, n)); }
};

template <int N>
PB_HD Dual<N> branch_value(const Tex &t, const float *x, int seed0, bool primary, float3 rd, bool light, float2 disk, float p_other, bool use_mis,
                          float inv_cnt, float3 gA, float3 gB, float g_frozen = -1.f) {   // g_frozen: tests only — the detached G of pdf0 / pdf1 held at this value
    typedef Dual<N> D;
    D in[15];
#pragma unroll
    for (int k = 0; k < 15; ++k) {
        in[k] = D(x[k]);
#pragma unroll
        for (int j = 0; j < N; ++j) if (k == seed0 + j) in[k].d[j] = 1.f;
    }
    const V3<D> p(in[0], in[1], in[2]), shn(in[3], in[4], in[5]), a(in[6], in[7], in[8]), q(in[9], in[10], in[11]), nq(in[12], in[13], in[14]);
    const FrameT<D> fr(shn);
    const V3<D> wi = primary ? fr.to_local(-a) : fr.to_local(-vnormalize(p - a));   // camera vertex: `a` is the ray direction
    const V3<D> dv = q - p;
    const D r2 = vdot(dv, dv);
    const V3<D> wo = dv / dsqrt(r2);
    const V3<D> wo_l = fr.to_local(wo);
    const V3<D> H = vnormalize(wo_l + wi);
    const D au(t.au), av(t.av);
    const D R = eval_scalar<D>(au, av, wi, wo_l, H);
    if (R.v == 0.f) return D(0.f);
    const D G = dabs(vdot(nq, wo)) / r2;
    const float G_det = g_frozen > 0.f ? g_frozen : G.v;   // detach(G_val), direct.cpp:93,147
    D wA(inv_cnt), base;
    if (!light) {   // pdf0 = bs.pdf * detach(G)
        const D pdf0 = sampled_pdf<D>(au, av, wi, disk) * G_det;
        if (use_mis) { const D w1 = dsqr(pdf0); wA = w1 / (w1 + sqr(p_other)) * inv_cnt; }
        base = R * G / pdf0;
    } else {        // pdf1 = bsdf.pdf * detach(G); p_other = ps.pdf
        if (use_mis) { const D pdf1 = pdf<D>(au, av, wi, wo_l) * G_det; const float w1 = sqr(p_other); wA = w1 / (w1 + dsqr(pdf1)) * inv_cnt; }
        base = R * G / p_other;
    }
    const D cos_h = vdot(wi, H);
    const float ga[3] = {gA.x, gA.y, gA.z}, gb[3] = {gB.x, gB.y, gB.z}, sp[3] = {t.spec.x, t.spec.y, t.spec.z};
    D sum(0.f);
#pragma unroll
    for (int c = 0; c < 3; ++c) sum += (wA * ga[c] + gb[c]) * fresnel1<D>(D(getc(t.eta, c)), D(getc(t.k, c)), cos_h) * sp[c];
    return sum * base;
}

struct GeomGrad { float3 p, shn, a, q, nq; float c0; };
// false: degenerate sample (non-finite derivative), nothing to add
PB_HD bool branch_geom_grad(const Tex &t, float3 p, float3 shn, float3 a, float3 q, float3 nq, bool primary, float3 rd, bool light, float2 disk,
                           float p_other, bool use_mis, float inv_cnt, float3 gA, float3 gB, GeomGrad &g) {
    const float x[15] = {p.x, p.y, p.z, shn.x, shn.y, shn.z, a.x, a.y, a.z, q.x, q.y, q.z, nq.x, nq.y, nq.z};
    float d[15];
    bool ok = true;
    float c0 = 0.f;
#pragma unroll 1
    for (int chunk = 0; chunk < 3; ++chunk) {
        const Dual<5> c = branch_value<5>(t, x, 5 * chunk, primary, rd, light, disk, p_other, use_mis, inv_cnt, gA, gB);
        ok = ok && dfinite(c);
        c0 = c.v;
#pragma unroll
        for (int j = 0; j < 5; ++j) d[5 * chunk + j] = c.d[j];
    }
    if (!ok) return false;
    g.p = f3(d[0], d[1], d[2]); g.shn = f3(d[3], d[4], d[5]); g.a = f3(d[6], d[7], d[8]);
    g.q = f3(d[9], d[10], d[11]); g.nq = f3(d[12], d[13], d[14]); g.c0 = c0;
    return true;
}

// ---- scatter of the local gradient into the textures (Bitmap::eval's backward, bitmap.cpp:43-89) -------------------------------
// reverse mode: atomics into the gradient segment; forward mode (S.tri_tangent set): dot with the texture's tangent.
template <int C> PB_D void tex_scatter(const SceneView &S, const TexRef &t, float2 uv, const float *gv) {
    if (!t.grad) return;
    bool any = false;
#pragma unroll
    for (int c = 0; c < C; ++c) any = any || gv[c] != 0.f;
    if (!any) return;
    const TexTap tap = tex_tap(t, uv);
    const int n = tap.constant ? 1 : 4;
    const float w[4] = {tap.w0y * tap.w0x, tap.w0y * tap.w1x, tap.w1y * tap.w0x, tap.w1y * tap.w1x};
    const int idx[4] = {tap.idx, tap.idx + 1, tap.idx + t.w, tap.idx + t.w + 1};
    if (S.tri_tangent) {
        float s = 0.f;
        for (int k = 0; k < n; ++k)
#pragma unroll
            for (int c = 0; c < C; ++c) s = fmaf(w[k] * gv[c], __ldg(t.grad + idx[k] * C + c), s);
        if (isfinite(s)) S.jvp_acc[blockIdx.x * blockDim.x + threadIdx.x] += s;
        return;
    }
    for (int k = 0; k < n; ++k)
#pragma unroll
        for (int c = 0; c < C; ++c) if (gv[c] != 0.f) atomicAdd(t.grad + idx[k] * C + c, gv[c] * w[k]);
}

// Per-thread accumulator for 1x1 textures (the common case: one atomic per warp and parameter instead of one per lane);
// bitmap textures scatter immediately.
PB_D void emit_tex_grad(const SceneView &S, const BsdfRec *b, float2 uv, const TexGrad &g, TexGrad &acc_const) {
    const bool all_const = b->tex[TEX_ALPHA_U].w * b->tex[TEX_ALPHA_U].h == 1 && b->tex[TEX_ALPHA_V].w * b->tex[TEX_ALPHA_V].h == 1 &&
                           b->tex[TEX_ETA].w * b->tex[TEX_ETA].h == 1 && b->tex[TEX_K].w * b->tex[TEX_K].h == 1 &&
                           b->tex[TEX_SPECULAR].w * b->tex[TEX_SPECULAR].h == 1;
    if (all_const && !S.tri_tangent) { acc_const.add(g); return; }
    const float e[3] = {g.eta.x, g.eta.y, g.eta.z}, k[3] = {g.k.x, g.k.y, g.k.z}, s[3] = {g.spec.x, g.spec.y, g.spec.z};
    tex_scatter<1>(S, b->tex[TEX_ALPHA_U], uv, &g.au);
    tex_scatter<1>(S, b->tex[TEX_ALPHA_V], uv, &g.av);
    tex_scatter<3>(S, b->tex[TEX_ETA], uv, e);
    tex_scatter<3>(S, b->tex[TEX_K], uv, k);
    tex_scatter<3>(S, b->tex[TEX_SPECULAR], uv, s);
}
// Block-level flush of the 1x1-texture accumulators (grouped by BSDF id) and of the environment map's scale accumulator:
// warp shuffles, then shared-memory slots, then one global atomic per block and parameter. Every thread of the block calls it.
constexpr int kConstSlots = 8;                        // BSDF ids below this reduce through shared memory, others per warp
constexpr int kFlushFloats = kConstSlots * 11 + 1;    // + the environment map's scale
PB_D void flush_const_tex_grad(const SceneView &S, int bsdf_id, const TexGrad &acc, float env_scale_acc, float *s_buf) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < kFlushFloats; t += blockDim.x) s_buf[t] = 0.f;
    __syncthreads();
    const bool has = bsdf_id >= 0 && (acc.au != 0.f || acc.av != 0.f || acc.eta.x != 0.f || acc.eta.y != 0.f || acc.eta.z != 0.f || acc.k.x != 0.f ||
                                      acc.k.y != 0.f || acc.k.z != 0.f || acc.spec.x != 0.f || acc.spec.y != 0.f || acc.spec.z != 0.f);
    unsigned remaining = __ballot_sync(full, has);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int key = __shfl_sync(full, bsdf_id, leader);
        const bool mine = has && bsdf_id == key;
        const unsigned grp = __ballot_sync(full, mine);
        float v[11] = {acc.au, acc.av, acc.eta.x, acc.eta.y, acc.eta.z, acc.k.x, acc.k.y, acc.k.z, acc.spec.x, acc.spec.y, acc.spec.z};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            float x = mine ? v[k] : 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
            v[k] = x;
        }
        if (lane == leader) {
            if (key < kConstSlots) {
#pragma unroll
                for (int k = 0; k < 11; ++k) if (v[k] != 0.f) atomicAdd(s_buf + key * 11 + k, v[k]);
            } else {
                const BsdfRec &b = S.bsdfs[key];
                if (b.tex[TEX_ALPHA_U].grad && v[0] != 0.f) atomicAdd(b.tex[TEX_ALPHA_U].grad, v[0]);
                if (b.tex[TEX_ALPHA_V].grad && v[1] != 0.f) atomicAdd(b.tex[TEX_ALPHA_V].grad, v[1]);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (b.tex[TEX_ETA].grad && v[2 + c] != 0.f) atomicAdd(b.tex[TEX_ETA].grad + c, v[2 + c]);
                    if (b.tex[TEX_K].grad && v[5 + c] != 0.f) atomicAdd(b.tex[TEX_K].grad + c, v[5 + c]);
                    if (b.tex[TEX_SPECULAR].grad && v[8 + c] != 0.f) atomicAdd(b.tex[TEX_SPECULAR].grad + c, v[8 + c]);
                }
            }
        }
        remaining &= ~grp;
    }
    float es = env_scale_acc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(full, es, o);
    if (lane == 0 && es != 0.f) atomicAdd(s_buf + kConstSlots * 11, es);
    __syncthreads();
    for (int t = threadIdx.x; t < kFlushFloats; t += blockDim.x) {
        const float v = s_buf[t];
        if (v == 0.f) continue;
        if (t == kConstSlots * 11) {
            if (S.emitter_env >= 0 && S.emitters[S.emitter_env].env_scale_grad) atomicAdd(S.emitters[S.emitter_env].env_scale_grad, v);
            continue;
        }
        const int b = t / 11, k = t - 11 * b;
        if (b >= S.num_bsdfs) continue;
        const BsdfRec &br = S.bsdfs[b];
        float *gp = k == 0 ? br.tex[TEX_ALPHA_U].grad : k == 1 ? br.tex[TEX_ALPHA_V].grad : k < 5 ? br.tex[TEX_ETA].grad : k < 8 ? br.tex[TEX_K].grad : br.tex[TEX_SPECULAR].grad;
        const int off = k < 2 ? 0 : (k - 2) % 3;
        if (gp) atomicAdd(gp + off, v);
    }
}

}  // namespace rc
}  // namespace pb
