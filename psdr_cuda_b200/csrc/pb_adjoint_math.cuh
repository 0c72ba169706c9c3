// psdr-b200: reverse-mode building blocks for the geometry terms (hand-written replacements for what Enoki's tape
// records in Scene::ray_intersect<true,*> src/scene/scene.cpp:289-384, process_mesh src/shape/mesh.cpp:19-51 and
// ray_intersect_triangle include/psdr/utils.h:67-77). Each *_vjp takes the forward inputs and the adjoint of the output
// and returns / accumulates the adjoints of the inputs. Host-callable so they can be unit-tested against finite
// differences without a GPU (tests/native/adjoint_check.cu).
#pragma once
#include "pb_math.cuh"

namespace pb {

// gradient record of one triangle-table entry: the 22 AD-tracked floats of TriangleInfo (types.h:136-146)
struct TriGrad {
    float3 p0, e1, e2, n0, n1, n2, fn;
    float area;
    PB_HD TriGrad() : p0(f3(0.f)), e1(f3(0.f)), e2(f3(0.f)), n0(f3(0.f)), n1(f3(0.f)), n2(f3(0.f)), fn(f3(0.f)), area(0.f) {}
};
constexpr int kTriGradStride = 24;   // floats per record in the device buffer (22 used)

PB_HD float pdot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PB_HD float3 pcross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

// y = x / |x|
PB_HD float3 normalize_vjp(float3 x, float3 gy) {
    const float inv = 1.f / sqrtf(pdot(x, x));
    const float3 y = x * inv;
    return (gy - y * pdot(y, gy)) * inv;
}

// (u, v, t) = ray_intersect_triangle(p0, e1, e2, o, d)   (utils.h:67-77)
struct RayTriGrad { float3 p0, e1, e2, o, d; };
PB_HD RayTriGrad ray_intersect_triangle_vjp(float3 p0, float3 e1, float3 e2, float3 o, float3 d, float gu, float gv, float gt) {
    const float3 h = pcross(d, e2);
    const float a = pdot(e1, h), f = 1.f / a;
    const float3 s = o - p0, q = pcross(s, e1);
    const float gf = gu * pdot(s, h) + gv * pdot(d, q) + gt * pdot(e2, q);
    float3 g_s = h * (gu * f), g_h = s * (gu * f);
    float3 g_d = q * (gv * f);
    const float3 g_q = d * (gv * f) + e2 * (gt * f);
    float3 g_e2 = q * (gt * f);
    const float ga = -gf * f * f;
    float3 g_e1 = h * ga;
    g_h += e1 * ga;
    // q = s x e1
    g_s += pcross(e1, g_q);
    g_e1 += pcross(g_q, s);
    // h = d x e2
    g_d += pcross(e2, g_h);
    g_e2 += pcross(g_h, d);
    RayTriGrad r;
    r.p0 = -g_s; r.e1 = g_e1; r.e2 = g_e2; r.o = g_s; r.d = g_d;
    return r;
}

// The geometric factor every connection shares:  c = cos_o * G * J,  wo = (q - p)/|q - p|,  cos_o = wo . sh_n,
// G = |n_q . (-wo)| / |q - p|^2   (direct.cpp:84-95 and 123-152 with f = rho/pi * cos_o factored out).
struct ConnGrad { float3 p, q, sh_n, n_q; float J; };
PB_HD float connection_value(float3 p, float3 q, float3 sh_n, float3 n_q, float J) {
    const float3 dv = q - p;
    const float r2 = pdot(dv, dv), r = sqrtf(r2);
    const float3 wo = dv * (1.f / r);
    return pdot(wo, sh_n) * fabsf(pdot(n_q, wo)) / r2 * J;
}
PB_HD ConnGrad connection_vjp(float3 p, float3 q, float3 sh_n, float3 n_q, float J, float gc) {
    const float3 dv = q - p;
    const float r2 = pdot(dv, dv), r = sqrtf(r2), inv_r = 1.f / r, inv_r2 = 1.f / r2;
    const float3 wo = dv * inv_r;
    const float A = pdot(wo, sh_n);
    const float Bq = -pdot(n_q, wo);
    const float sg = Bq < 0.f ? -1.f : 1.f, aB = fabsf(Bq);
    ConnGrad g;
    g.sh_n = wo * (gc * aB * J * inv_r2);
    g.n_q = wo * (-gc * A * J * inv_r2 * sg);
    g.J = gc * A * aB * inv_r2;
    const float3 dA = (sh_n - wo * A) * inv_r;              // dA/ddv
    const float3 dB = (n_q + wo * Bq) * (-inv_r);           // dBq/ddv
    const float3 dc = (dA * (aB * inv_r2) + dB * (A * sg * inv_r2) + wo * (-2.f * A * aB * inv_r2 * inv_r)) * (J * gc);
    g.q = dc; g.p = -dc;
    return g;
}

// shading normal: sh_n = normalize(n0 + u (n1 - n0) + v (n2 - n0))   (scene.cpp:331-335 / 360-364)
PB_HD void shading_normal_vjp(float3 n0, float3 n1, float3 n2, float u, float v, float3 g_shn, TriGrad &tg, float &gu, float &gv) {
    const float3 m = n0 + (n1 - n0) * u + (n2 - n0) * v;
    const float3 gm = normalize_vjp(m, g_shn);
    tg.n0 += gm * (1.f - u - v); tg.n1 += gm * u; tg.n2 += gm * v;
    gu += pdot(gm, n1 - n0); gv += pdot(gm, n2 - n0);
}

// mesh.cpp:19-51 backward for one face: c = e1 x e2, fn = c/|c|, area = |c|/2, plus the face's share g_cn of the
// vertex-normal sums (each vertex normal is normalize(sum c) — the area weights cancel under the normalisation)
PB_HD void face_vjp(float3 e1, float3 e2, float3 g_fn, float g_area, float3 g_cn, float3 &g_e1, float3 &g_e2) {
    const float3 c = pcross(e1, e2);
    const float len = sqrtf(pdot(c, c));
    const float3 fn = c * (1.f / len);
    const float3 g_c = normalize_vjp(c, g_fn) + fn * (0.5f * g_area) + g_cn;
    g_e1 += pcross(e2, g_c);
    g_e2 += pcross(g_c, e1);
}

}  // namespace pb
