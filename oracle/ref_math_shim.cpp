// TEST INFRASTRUCTURE (oracle/): C entry points around psdr-cuda's OWN math source, compiled unmodified from where it lies under
// /root/reference — include/psdr/core/warp.h, include/psdr/core/frame.h, include/psdr/utils.h, src/bsdf/{ggx,diffuse,roughconductor}.cpp,
// src/core/bitmap.cpp (+ their headers) — against the scalar Enoki stand-in oracle/ref_stub/ (Enoki is an un-vendored dependency of the reference). Built by
// oracle/build_ref.sh into oracle/_ref/libref_math.so; tests/test_ref_math.py pins the oracle's restatement of these formulas to it.
// Every function runs the reference's `C` flavour (ad = false) on one lane.
#include <psdr/psdr.h>
#include <psdr/core/ray.h>
#include <psdr/core/warp.h>
#include <psdr/core/frame.h>
#include "src/bsdf/ggx.cpp"   // resolved against -I<reference root>: the reference's translation units, as they are
#include "src/core/sampler.cpp"
#include "src/core/bitmap_loader.cpp"
#include "src/core/bitmap.cpp"
#include "src/bsdf/diffuse.cpp"
#include "src/bsdf/roughconductor.cpp"

using namespace psdr;

static Vector3fC v3(const float *p) { return Vector3fC(p[0], p[1], p[2]); }
static Vector2fC v2(const float *p) { return Vector2fC(p[0], p[1]); }
static void put(float *o, const Vector3fC &v) { o[0] = v.x().v; o[1] = v.y().v; o[2] = v.z().v; }
static void put(float *o, const Vector2fC &v) { o[0] = v.x().v; o[1] = v.y().v; }

extern "C" {

// include/psdr/core/warp.h
void ref_square_to_uniform_disk_concentric(const float *s, float *out) { put(out, warp::square_to_uniform_disk_concentric<false>(v2(s))); }
void ref_square_to_cosine_hemisphere(const float *s, float *out) { put(out, warp::square_to_cosine_hemisphere<false>(v2(s))); }
float ref_square_to_cosine_hemisphere_pdf(const float *v) { return warp::square_to_cosine_hemisphere_pdf<false>(v3(v)).v; }
void ref_square_to_uniform_triangle(const float *s, float *out) { put(out, warp::square_to_uniform_triangle<false>(v2(s))); }

// include/psdr/core/frame.h
void ref_frame(const float *n, float *s_out, float *t_out) { FrameC f(v3(n)); put(s_out, f.s); put(t_out, f.t); }
void ref_frame_to_local(const float *n, const float *v, float *out) { FrameC f(v3(n)); put(out, f.to_local(v3(v))); }
void ref_frame_to_world(const float *n, const float *v, float *out) { FrameC f(v3(n)); put(out, f.to_world(v3(v))); }

// include/psdr/utils.h
void ref_ray_intersect_triangle(const float *p0, const float *e1, const float *e2, const float *o, const float *d, float *uvt) {
    auto [uv, t] = ray_intersect_triangle<false>(v3(p0), v3(e1), v3(e2), RayC(v3(o), v3(d)));
    uvt[0] = uv.x().v; uvt[1] = uv.y().v; uvt[2] = t.v;
}
void ref_bilinear(const float *p0, const float *e1, const float *e2, const float *st, float *out) { put(out, bilinear<false>(v3(p0), v3(e1), v3(e2), v2(st))); }
void ref_sphdir(float theta, float phi, float *out) { put(out, sphdir<false>(FloatC(theta), FloatC(phi))); }
float ref_rgb2luminance(const float *rgb) { return rgb2luminance<false>(v3(rgb)).v; }
int ref_sign_eps(float x, float eps) { return sign<false>(FloatC(x), eps).v; }
void ref_fresnel(const float *eta, const float *k, float cos_theta_i, float *out) { put(out, fresnel<false>(v3(eta), v3(k), FloatC(cos_theta_i))); }
void ref_ray_intersect_scene_aabb(const float *o, const float *d, const float *lo, const float *hi, float *t_n_G) {
    auto [t, n, G] = ray_intersect_scene_aabb<false>(RayC(v3(o), v3(d)), v3(lo), v3(hi));
    t_n_G[0] = t.v; put(t_n_G + 1, n); t_n_G[4] = G.v;
}

// src/bsdf/ggx.cpp
float ref_ggx_eval(float au, float av, const float *m) { return GGXDistribution(FloatD(au), FloatD(av)).eval<false>(v3(m)).v; }
float ref_ggx_smith_g1(float au, float av, const float *v, const float *m) { return GGXDistribution(FloatD(au), FloatD(av)).smith_g1<false>(v3(v), v3(m)).v; }
float ref_ggx_G(float au, float av, const float *wi, const float *wo, const float *m) { return GGXDistribution(FloatD(au), FloatD(av)).G<false>(v3(wi), v3(wo), v3(m)).v; }
void ref_ggx_sample(float au, float av, const float *wi, const float *s3, float *out) { put(out, GGXDistribution(FloatD(au), FloatD(av)).sample<false>(v3(wi), v3(s3))); }
void ref_ggx_sample_visible_11(float cos_theta_i, const float *s2, float *out) { put(out, GGXDistribution().sample_visible_11<false>(FloatC(cos_theta_i), v2(s2))); }

// src/bsdf/diffuse.cpp and src/bsdf/roughconductor.cpp with constant (1x1) textures (src/core/bitmap.cpp; the one-lane stand-in has no
// multi-texel arrays). its.wi is the only field of the intersection these methods read besides uv.
static IntersectionC its_from_wi(const float *wi) { IntersectionC its; its.wi = v3(wi); its.uv = Vector2fC(0.f, 0.f); return its; }
static ScalarVector3f sv3(const float *p) { return ScalarVector3f(p[0], p[1], p[2]); }
void ref_diffuse_eval(const float *rho, const float *wi, const float *wo, float *out) { put(out, Diffuse(sv3(rho)).eval(its_from_wi(wi), v3(wo), MaskC(true))); }
float ref_diffuse_pdf(const float *rho, const float *wi, const float *wo) { return Diffuse(sv3(rho)).pdf(its_from_wi(wi), v3(wo), MaskC(true)).v; }
int ref_diffuse_sample(const float *rho, const float *wi, const float *s3, float *wo_pdf) {
    BSDFSampleC bs = Diffuse(sv3(rho)).sample(its_from_wi(wi), v3(s3), MaskC(true));
    put(wo_pdf, bs.wo); wo_pdf[3] = bs.pdf.v;
    return bs.is_valid.v ? 1 : 0;
}
static RoughConductor make_rc(const float *p) {   // alpha_u, alpha_v, eta[3], k[3], specular_reflectance[3]
    return RoughConductor(Bitmap1fD(p[0]), Bitmap1fD(p[1]), Bitmap3fD(sv3(p + 2)), Bitmap3fD(sv3(p + 5)), Bitmap3fD(sv3(p + 8)));
}
void ref_rc_eval(const float *prm, const float *wi, const float *wo, float *out) { put(out, make_rc(prm).eval(its_from_wi(wi), v3(wo), MaskC(true))); }
float ref_rc_pdf(const float *prm, const float *wi, const float *wo) { return make_rc(prm).pdf(its_from_wi(wi), v3(wo), MaskC(true)).v; }
int ref_rc_sample(const float *prm, const float *wi, const float *s3, float *wo_pdf) {
    BSDFSampleC bs = make_rc(prm).sample(its_from_wi(wi), v3(s3), MaskC(true));
    put(wo_pdf, bs.wo); wo_pdf[3] = bs.pdf.v;
    return bs.is_valid.v ? 1 : 0;
}

// src/core/sampler.cpp: lane `lane` of a Sampler seeded with arange(count) (scene.cpp:65-79): its first n uniform floats, then one
// next_2d and one next_nd<3> (their component order is the C++ argument-evaluation order of the compiler that builds this file: gcc)
void ref_sampler_lane(uint64_t lane, int n, float *out_1d, float *out_2d, float *out_3d) {
    enoki::stub_lane() = lane;
    Sampler s;
    s.seed(UInt64C(lane));
    for (int i = 0; i < n; ++i) out_1d[i] = s.next_1d<false>().v;
    put(out_2d, s.next_2d<false>());
    put(out_3d, s.next_nd<3, false>());
    enoki::stub_lane() = 0;
}

}  // extern "C"
