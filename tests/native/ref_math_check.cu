// CPU check (host-compiled with nvcc, no GPU): the PRODUCT's per-lane math — csrc/pb_math.cuh (warps, shading frame, ray / triangle, bilinear,
// luminance, the psdr sampler streams) and csrc/pb_rc.cuh (GGX distribution, visible-normal sampling, Smith G1, conductor Fresnel, the
// rough-conductor eval / pdf / sample) — against the REFERENCE'S OWN SOURCE: oracle/_ref/libref_math.so is psdr-cuda's warp.h, frame.h,
// utils.h, ggx.cpp, roughconductor.cpp, sampler.cpp compiled unmodified (oracle/build_ref.sh). argv[1] = path of that library; with
// argv[2] = oracle/_ref/libref_render.so also the DERIVATIVES: pb_rc.cuh's local forward-mode duals of the rough-conductor eval / pdf /
// sampled pdf with respect to (alpha_u, alpha_v, eta, k, specular reflectance) against the tangents the reference's own roughconductor.cpp
// produces in its D flavour (its detach() placement included) under the forward-mode stand-in for Enoki's autodiff.
// The same functions are what the sm_100a kernels inline (the host path of pb_math.cuh uses fmaf / sqrtf where the device uses
// __fmaf_rn / __fsqrt_rn: IEEE-identical), so this pins the kernels' arithmetic to the reference function by function, not image by image.
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "../../psdr_cuda_b200/csrc/pb_rc.cuh"

using namespace pb;

static std::mt19937 rng(11);
static float U(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
static float3 unit3() { std::normal_distribution<float> g; float3 v = f3(g(rng), g(rng), g(rng)); const float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); return f3(v.x / n, v.y / n, v.z / n); }
static int bad = 0, checked = 0;
static bool close(float a, float b, float ulps, float atol = 0.f) { return std::fabs((double)a - b) <= ulps * 1.2e-7 * std::fmax(std::fabs(a), std::fabs(b)) + atol; }
static void expect(bool ok, const char *what, float a, float b) { ++checked; if (!ok) { if (bad < 20) std::printf("MISMATCH %s: reference %.9g product %.9g\n", what, a, b); ++bad; } }

int main(int argc, char **argv) {
    if (argc < 2) { std::printf("usage: ref_math_check <libref_math.so>\n"); return 2; }
    void *h = dlopen(argv[1], RTLD_NOW);
    if (!h) { std::printf("cannot load %s: %s\n", argv[1], dlerror()); return 2; }
#define SYM(name, type) auto name = (type)dlsym(h, #name); if (!name) { std::printf("missing %s\n", #name); return 2; }
    SYM(ref_square_to_uniform_disk_concentric, void (*)(const float *, float *))
    SYM(ref_square_to_cosine_hemisphere, void (*)(const float *, float *))
    SYM(ref_square_to_uniform_triangle, void (*)(const float *, float *))
    SYM(ref_frame, void (*)(const float *, float *, float *))
    SYM(ref_frame_to_local, void (*)(const float *, const float *, float *))
    SYM(ref_frame_to_world, void (*)(const float *, const float *, float *))
    SYM(ref_ray_intersect_triangle, void (*)(const float *, const float *, const float *, const float *, const float *, float *))
    SYM(ref_bilinear, void (*)(const float *, const float *, const float *, const float *, float *))
    SYM(ref_rgb2luminance, float (*)(const float *))
    SYM(ref_fresnel, void (*)(const float *, const float *, float, float *))
    SYM(ref_ggx_eval, float (*)(float, float, const float *))
    SYM(ref_ggx_smith_g1, float (*)(float, float, const float *, const float *))
    SYM(ref_ggx_sample, void (*)(float, float, const float *, const float *, float *))
    SYM(ref_rc_eval, void (*)(const float *, const float *, const float *, float *))
    SYM(ref_rc_pdf, float (*)(const float *, const float *, const float *))
    SYM(ref_rc_sample, int (*)(const float *, const float *, const float *, float *))
    SYM(ref_sampler_lane, void (*)(uint64_t, int, float *, float *, float *))

    for (int it = 0; it < 3000; ++it) {
        // warps (warp.h:14-80)
        const float s[2] = {U(0, 1), U(0, 1)};
        float r2[2], r3[3];
        ref_square_to_uniform_disk_concentric(s, r2);
        const float2 d = square_to_uniform_disk_concentric(s[0], s[1]);
        expect(close(r2[0], d.x, 4, 2e-7f) && close(r2[1], d.y, 4, 2e-7f), "square_to_uniform_disk_concentric", r2[0], d.x);
        ref_square_to_cosine_hemisphere(s, r3);
        const float3 c = square_to_cosine_hemisphere(s[0], s[1]);
        expect(close(r3[0], c.x, 4, 2e-7f) && close(r3[1], c.y, 4, 2e-7f) && close(r3[2], c.z, 4, 3e-7f), "square_to_cosine_hemisphere", r3[2], c.z);
        ref_square_to_uniform_triangle(s, r2);
        const float2 t = square_to_uniform_triangle(s[0], s[1]);
        expect(close(r2[0], t.x, 4, 2e-7f) && close(r2[1], t.y, 4, 2e-7f), "square_to_uniform_triangle", r2[0], t.x);
        // frame (frame.h:9-52)
        const float3 n = unit3(), v = f3(U(-2, 2), U(-2, 2), U(-2, 2));
        const float nn[3] = {n.x, n.y, n.z}, vv[3] = {v.x, v.y, v.z};
        float fs[3], ft[3], fl[3];
        ref_frame(nn, fs, ft);
        const Frame F(n);
        expect(fs[0] == F.s.x && fs[1] == F.s.y && fs[2] == F.s.z && ft[0] == F.t.x && ft[1] == F.t.y && ft[2] == F.t.z, "frame basis", fs[0], F.s.x);
        ref_frame_to_local(nn, vv, fl);
        float3 l = F.to_local(v);
        expect(close(fl[0], l.x, 4, 3e-7f) && close(fl[1], l.y, 4, 3e-7f) && close(fl[2], l.z, 4, 3e-7f), "frame.to_local", fl[0], l.x);
        ref_frame_to_world(nn, vv, fl);
        l = F.to_world(v);
        expect(close(fl[0], l.x, 4, 3e-7f) && close(fl[1], l.y, 4, 3e-7f) && close(fl[2], l.z, 4, 3e-7f), "frame.to_world", fl[0], l.x);
        // ray / triangle (utils.h:67-77), bilinear (utils.h:49-51), luminance (utils.h:61-63)
        const float3 p0 = f3(U(-1, 1), U(-1, 1), U(-1, 1)), e1 = f3(U(-1, 1), U(-1, 1), U(-1, 1)), e2 = f3(U(-1, 1), U(-1, 1), U(-1, 1)), o = f3(U(-3, 3), U(-3, 3), U(-3, 3)), dir = unit3();
        const float a0[3] = {p0.x, p0.y, p0.z}, a1[3] = {e1.x, e1.y, e1.z}, a2[3] = {e2.x, e2.y, e2.z}, ao[3] = {o.x, o.y, o.z}, ad[3] = {dir.x, dir.y, dir.z};
        float uvt[3], pu, pv, pt;
        ref_ray_intersect_triangle(a0, a1, a2, ao, ad, uvt);
        ray_intersect_triangle(p0, e1, e2, o, dir, pu, pv, pt);
        const float mag = std::fmax(1.f, std::fmax(std::fabs(uvt[0]), std::fmax(std::fabs(uvt[1]), std::fabs(uvt[2]))));
        expect(close(uvt[0], pu, 2, 1e-7f * mag) && close(uvt[1], pv, 2, 1e-7f * mag) && close(uvt[2], pt, 2, 1e-7f * mag), "ray_intersect_triangle", uvt[2], pt);
        ref_bilinear(a0, a1, a2, s, r3);
        const float3 b = bilinear(p0, e1, e2, s[0], s[1]);
        expect(r3[0] == b.x && r3[1] == b.y && r3[2] == b.z, "bilinear", r3[0], b.x);
        const float rgb[3] = {U(0, 5), U(0, 5), U(0, 5)};
        expect(close(ref_rgb2luminance(rgb), luminance(f3(rgb[0], rgb[1], rgb[2])), 2), "rgb2luminance", ref_rgb2luminance(rgb), luminance(f3(rgb[0], rgb[1], rgb[2])));
        // Fresnel (utils.h:149-164), GGX (ggx.cpp)
        const float eta[3] = {U(0.1f, 2.f), U(0.1f, 2.f), U(0.1f, 2.f)}, k[3] = {U(0.5f, 5.f), U(0.5f, 5.f), U(0.5f, 5.f)}, ct = U(0.01f, 1.f);
        ref_fresnel(eta, k, ct, r3);
        for (int ch = 0; ch < 3; ++ch) expect(close(r3[ch], rc::fresnel1<float>(eta[ch], k[ch], ct), 8), "fresnel", r3[ch], rc::fresnel1<float>(eta[ch], k[ch], ct));
        const float au = U(0.08f, 0.8f), av = U(0.08f, 0.8f);
        float3 m = unit3(); m.z = std::fabs(m.z);
        float3 wi = unit3(); wi.z = std::fabs(wi.z) + 1e-3f; { const float q = std::sqrt(wi.x * wi.x + wi.y * wi.y + wi.z * wi.z); wi = f3(wi.x / q, wi.y / q, wi.z / q); }
        float3 wo = unit3(); wo.z = std::fabs(wo.z);
        const float am[3] = {m.x, m.y, m.z}, awi[3] = {wi.x, wi.y, wi.z}, awo[3] = {wo.x, wo.y, wo.z};
        const V3<float> M(m), WI(wi), WO(wo);
        expect(close(ref_ggx_eval(au, av, am), rc::ggx_eval<float>(au, av, M), 8), "ggx eval", ref_ggx_eval(au, av, am), rc::ggx_eval<float>(au, av, M));
        expect(close(ref_ggx_smith_g1(au, av, awi, am), rc::smith_g1<float>(au, av, WI, M), 8), "ggx smith_g1", ref_ggx_smith_g1(au, av, awi, am), rc::smith_g1<float>(au, av, WI, M));
        const float s3[3] = {U(0.01f, 0.99f), U(0.01f, 0.99f), U(0.01f, 0.99f)};
        const float2 disk = square_to_uniform_disk_concentric(s3[0], s3[1]);   // roughconductor.cpp:87: GGX::sample gets head<2>(sample)
        // the stretched direction is nearly the normal for small alpha and sin = sqrt(1 - cos^2) amplifies its last bit by 1 / sin^2
        const double wx = au * wi.x, wy = av * wi.y, wn = std::sqrt(wx * wx + wy * wy + (double)wi.z * wi.z);
        const float amp = (float)(1.0 / std::fmax(1e-6, (wx * wx + wy * wy) / (wn * wn)));
        ref_ggx_sample(au, av, awi, s3, r3);
        const V3<float> ms = rc::ggx_sample<float>(au, av, WI, disk);
        const float tol = 2e-6f + 2.4e-7f * amp;
        expect(close(r3[0], ms.x, 32, tol) && close(r3[1], ms.y, 32, tol) && close(r3[2], ms.z, 32, tol), "ggx sample", r3[2], ms.z);
        // RoughConductor eval / pdf / sample (roughconductor.cpp:40-93)
        float prm[11] = {au, av, eta[0], eta[1], eta[2], k[0], k[1], k[2], U(0.3f, 1.f), U(0.3f, 1.f), U(0.3f, 1.f)};
        ref_rc_eval(prm, awi, awo, r3);
        const V3<float> H = vnormalize(WO + WI);
        const float sc = rc::eval_scalar<float>(au, av, WI, WO, H);
        for (int ch = 0; ch < 3; ++ch) {
            const float mine = sc == 0.f ? 0.f : rc::fresnel1<float>(eta[ch], k[ch], vdot(WI, H)) * sc * prm[8 + ch];
            expect(close(r3[ch], mine, 32, 1e-9f), "rough conductor eval", r3[ch], mine);
        }
        const float pr = ref_rc_pdf(prm, awi, awo), pm = wo.z > 0.f ? rc::pdf<float>(au, av, WI, WO) : 0.f;
        expect(close(pr, pm, 32, 1e-9f), "rough conductor pdf", pr, pm);
        float wp[4];
        const int valid = ref_rc_sample(prm, awi, s3, wp);
        const float two = vdot(WI, ms) * 2.f;
        const V3<float> wos(ms.x * two - wi.x, ms.y * two - wi.y, ms.z * two - wi.z);
        const float ps = rc::sampled_pdf<float>(au, av, WI, disk);
        const int mine_valid = (wi.z > 0.f && ps > 0.f && wos.z > 0.f) ? 1 : 0;
        expect(valid == mine_valid, "rough conductor sample validity", (float)valid, (float)mine_valid);
        if (valid && mine_valid) {
            expect(close(wp[0], wos.x, 32, tol) && close(wp[1], wos.y, 32, tol) && close(wp[2], wos.z, 32, tol), "rough conductor sample wo", wp[2], wos.z);
            expect(std::fabs(wp[3] - ps) <= (1e-4f + 3e-6f * amp) * std::fmax(std::fabs(wp[3]), 1e-3f), "rough conductor sample pdf", wp[3], ps);
        }
    }
    // sampler streams (sampler.cpp:8-54, sampler.h:20-31): stream `lane` from its first draw, then a 2D and a 3D sample (component order = the
    // compiler's argument evaluation order in the reference; right to left under gcc)
    const uint64_t lanes[] = {0, 1, 2, 3, 12345, 67108863, (1ull << 31) + 17, (1ull << 33) + 5};
    for (uint64_t lane : lanes) {
        float r1[16], r2[2], r3[3];
        ref_sampler_lane(lane, 16, r1, r2, r3);
        Rng g(lane, make_jump(0));
        bool ok = true;
        for (int i = 0; i < 16; ++i) ok = ok && (g.next_1d() == r1[i]);
        const float2 a = g.next_2d();
        const float3 c = g.next_3d();
        ok = ok && a.x == r2[0] && a.y == r2[1] && c.x == r3[0] && c.y == r3[1] && c.z == r3[2];
        expect(ok, "sampler stream", r1[0], 0.f);
        // and jumped ahead: the stream position a second render starts from
        Rng j(lane, make_jump(5));
        expect(j.next_1d() == r1[5] && j.next_1d() == r1[6], "sampler stream after a jump of 5", r1[5], 0.f);
    }
    if (argc > 2) {
        void *hr = dlopen(argv[2], RTLD_NOW);
        if (!hr) { std::printf("cannot load %s: %s\n", argv[2], dlerror()); return 2; }
        auto ref_rc_d = (int (*)(int, const float *, const float *, const float *, const float *, float *, float *))dlsym(hr, "ref_rc_d");
        if (!ref_rc_d) { std::printf("missing ref_rc_d\n"); return 2; }
        using D1 = Dual<1>;
        auto dual = [](float v, float t) { D1 x(v); x.d[0] = t; return x; };
        auto dclose = [](float a, float b, float scale) { return std::fabs((double)a - b) <= 2e-3 * std::fmax(std::fabs(a), std::fabs(b)) + 1e-4 * scale; };
        for (int it = 0; it < 2000; ++it) {
            float prm[11] = {U(0.1f, 0.7f), U(0.1f, 0.7f), U(0.2f, 1.5f), U(0.2f, 1.5f), U(0.2f, 1.5f), U(1.5f, 4.f), U(1.5f, 4.f), U(1.5f, 4.f), U(0.3f, 1.f), U(0.3f, 1.f), U(0.3f, 1.f)};
            float tn[11];
            for (auto &x : tn) x = U(-1.f, 1.f);
            float3 wi = unit3(), wo = unit3();
            wi.z = std::fabs(wi.z) + 0.05f; wo.z = std::fabs(wo.z) + 0.05f;
            { const float q = std::sqrt(wi.x * wi.x + wi.y * wi.y + wi.z * wi.z); wi = f3(wi.x / q, wi.y / q, wi.z / q); }
            { const float q = std::sqrt(wo.x * wo.x + wo.y * wo.y + wo.z * wo.z); wo = f3(wo.x / q, wo.y / q, wo.z / q); }
            const float awi[3] = {wi.x, wi.y, wi.z}, awo[3] = {wo.x, wo.y, wo.z}, s3[3] = {U(0.05f, 0.95f), U(0.05f, 0.95f), U(0.05f, 0.95f)};
            const D1 au = dual(prm[0], tn[0]), av = dual(prm[1], tn[1]);
            const V3<D1> WI(D1(wi.x), D1(wi.y), D1(wi.z)), WO(D1(wo.x), D1(wo.y), D1(wo.z));
            float rv[3], rt[3];
            // eval = Fresnel * D G / (4 cos_i) * specular reflectance, attached to all eleven parameters
            if (ref_rc_d(0, prm, tn, awi, awo, rv, rt) != 0) { std::printf("ref_rc_d failed\n"); return 2; }
            const V3<D1> H = vnormalize(WO + WI);
            const D1 sc = rc::eval_scalar<D1>(au, av, WI, WO, H);
            for (int ch = 0; ch < 3; ++ch) {
                const D1 mine = rc::fresnel1<D1>(dual(prm[2 + ch], tn[2 + ch]), dual(prm[5 + ch], tn[5 + ch]), vdot(WI, H)) * sc * dual(prm[8 + ch], tn[8 + ch]);
                const float scale = std::fabs(rv[ch]) * 10.f + 1e-3f;
                expect(close(rv[ch], mine.v, 64, 1e-9f), "D eval value", rv[ch], mine.v);
                expect(sc.v == 0.f || dclose(rt[ch], mine.d[0], scale), "D eval tangent", rt[ch], mine.d[0]);
            }
            // pdf(wi, wo) and the pdf of the sampled direction, attached to the roughness
            ref_rc_d(1, prm, tn, awi, awo, rv, rt);
            D1 p = rc::pdf<D1>(au, av, WI, WO);
            expect(close(rv[0], p.v, 64, 1e-9f) && dclose(rt[0], p.d[0], std::fabs(rv[0]) * 10.f + 1e-3f), "D pdf", rt[0], p.d[0]);
            ref_rc_d(2, prm, tn, awi, s3, rv, rt);
            const float2 disk = square_to_uniform_disk_concentric(s3[0], s3[1]);
            p = rc::sampled_pdf<D1>(au, av, WI, disk);
            if (rv[0] > 1e-3f && p.v > 1e-3f)
                expect(std::fabs(rv[0] - p.v) <= 2e-3f * rv[0] && dclose(rt[0], p.d[0], std::fabs(rv[0]) * 20.f + 1e-2f), "D sampled pdf", rt[0], p.d[0]);
        }
    }
    std::printf("ref_math_check: %s (%d comparisons, %d mismatches)\n", bad ? "FAILED" : "ok", checked, bad);
    return bad ? 1 : 0;
}
