// CPU check (host-compiled with nvcc, no GPU): the per-lane shading primitives the sm_100a kernels inline — csrc/pb_shade.cuh: bitmap lookup,
// discrete-distribution sampling with sample reuse, the float GGX / Fresnel path of the primal kernels, BSDF eval / pdf / sample for diffuse
// and rough-conductor records, the scene-box exit of environment-map samples — against the REFERENCE'S OWN SOURCE compiled for the CPU:
// argv[1] = oracle/_ref/libref_math.so (ggx.cpp, diffuse.cpp, roughconductor.cpp, utils.h), argv[2] = oracle/_ref/libref_render.so
// (bitmap.cpp, pmf.cpp; and, when argv[3] = the tests/ directory is given, perspective.cpp's camera rays and envmap.cpp's direction lookup on
// the fixture scenes loaded and configured by the reference's own Scene). pb_shade.cuh is device code; for this check its functions are compiled __host__ __device__ (PB_D predefined) and
// its three device intrinsics (__ldg, __float_as_int) read memory / bits directly.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#define PB_D __host__ __device__ __forceinline__
template <class T> __host__ __device__ inline T host_ldg(const T *p) { return *p; }
#define __ldg(p) host_ldg(p)
__host__ __device__ inline int host_float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
#define __float_as_int(x) host_float_as_int(x)
#include "../../psdr_cuda_b200/csrc/pb_shade.cuh"

using namespace pb;

static std::mt19937 rng(23);
static float U(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
static float3 unit3() { std::normal_distribution<float> g; float3 v = f3(g(rng), g(rng), g(rng)); const float n = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z); return f3(v.x / n, v.y / n, v.z / n); }
static int bad = 0, checked = 0;
static bool close(float a, float b, float ulps, float atol = 0.f) { return std::fabs((double)a - b) <= ulps * 1.2e-7 * std::fmax(std::fabs(a), std::fabs(b)) + atol; }
static void expect(bool ok, const char *what, float a, float b) { ++checked; if (!ok) { if (bad < 20) std::printf("MISMATCH %s: reference %.9g product %.9g\n", what, a, b); ++bad; } }
static TexRef tex(const float *d, int w, int h, int c) { TexRef t; t.data = d; t.grad = nullptr; t.w = w; t.h = h; t.c = c; t.pad = 0; return t; }

int main(int argc, char **argv) {
    if (argc < 3) { std::printf("usage: ref_shade_check <libref_math.so> <libref_render.so> [tests dir]\n"); return 2; }
    void *hm = dlopen(argv[1], RTLD_NOW), *hr = dlopen(argv[2], RTLD_NOW);
    if (!hm || !hr) { std::printf("cannot load the reference libraries: %s\n", dlerror()); return 2; }
#define SYM(lib, name, type) auto name = (type)dlsym(lib, #name); if (!name) { std::printf("missing %s\n", #name); return 2; }
    SYM(hm, ref_fresnel, void (*)(const float *, const float *, float, float *))
    SYM(hm, ref_ggx_eval, float (*)(float, float, const float *))
    SYM(hm, ref_ggx_smith_g1, float (*)(float, float, const float *, const float *))
    SYM(hm, ref_ggx_sample, void (*)(float, float, const float *, const float *, float *))
    SYM(hm, ref_ggx_sample_visible_11, void (*)(float, const float *, float *))
    SYM(hm, ref_diffuse_eval, void (*)(const float *, const float *, const float *, float *))
    SYM(hm, ref_diffuse_pdf, float (*)(const float *, const float *, const float *))
    SYM(hm, ref_diffuse_sample, int (*)(const float *, const float *, const float *, float *))
    SYM(hm, ref_rc_eval, void (*)(const float *, const float *, const float *, float *))
    SYM(hm, ref_rc_pdf, float (*)(const float *, const float *, const float *))
    SYM(hm, ref_rc_sample, int (*)(const float *, const float *, const float *, float *))
    SYM(hm, ref_ray_intersect_scene_aabb, void (*)(const float *, const float *, const float *, const float *, float *))
    SYM(hr, ref_bitmap_eval, int (*)(int, int, int, const float *, const float *, int, int, float *))
    SYM(hr, ref_discrete_sample, int (*)(const float *, int, const float *, int, int, int *, float *, float *))

    for (int it = 0; it < 3000; ++it) {
        const float au = U(0.08f, 0.8f), av = U(0.08f, 0.8f);
        float3 m = unit3(); m.z = std::fabs(m.z);
        float3 wi = unit3(), wo = unit3();
        if (it % 5) { wi.z = std::fabs(wi.z) + 1e-3f; wo.z = std::fabs(wo.z); const float q = std::sqrt(wi.x * wi.x + wi.y * wi.y + wi.z * wi.z); wi = f3(wi.x / q, wi.y / q, wi.z / q); }
        const float am[3] = {m.x, m.y, m.z}, awi[3] = {wi.x, wi.y, wi.z}, awo[3] = {wo.x, wo.y, wo.z};
        const float s3[3] = {U(0.01f, 0.99f), U(0.01f, 0.99f), U(0.01f, 0.99f)};
        float r2[2], r3[3], r4[4];
        // the float GGX / Fresnel path of the primal kernels (pb_shade.cuh:160-215)
        expect(close(ref_ggx_eval(au, av, am), ggx::eval(au, av, m), 8), "ggx::eval", ref_ggx_eval(au, av, am), ggx::eval(au, av, m));
        if (wi.z > 0.f) {
            expect(close(ref_ggx_smith_g1(au, av, awi, am), ggx::smith_g1(au, av, wi, m), 8), "ggx::smith_g1", ref_ggx_smith_g1(au, av, awi, am), ggx::smith_g1(au, av, wi, m));
            ref_ggx_sample_visible_11(wi.z, s3, r2);
            const float2 sv = ggx::sample_visible_11(wi.z, s3[0], s3[1]);
            expect(close(r2[0], sv.x, 16, 1e-6f) && close(r2[1], sv.y, 16, 1e-6f), "ggx::sample_visible_11", r2[0], sv.x);
        }
        const float eta[3] = {U(0.1f, 2.f), U(0.1f, 2.f), U(0.1f, 2.f)}, k[3] = {U(0.5f, 5.f), U(0.5f, 5.f), U(0.5f, 5.f)}, ct = U(0.01f, 1.f);
        ref_fresnel(eta, k, ct, r3);
        const float3 fc = fresnel_conductor(f3(eta[0], eta[1], eta[2]), f3(k[0], k[1], k[2]), ct);
        expect(close(r3[0], fc.x, 8) && close(r3[1], fc.y, 8) && close(r3[2], fc.z, 8), "fresnel_conductor", r3[0], fc.x);
        // BSDF records with constant textures: diffuse.cpp / roughconductor.cpp through bsdf_eval / bsdf_pdf / bsdf_sample
        const float rho[3] = {U(0, 1), U(0, 1), U(0, 1)};
        const float prm[11] = {au, av, eta[0], eta[1], eta[2], k[0], k[1], k[2], U(0.3f, 1.f), U(0.3f, 1.f), U(0.3f, 1.f)};
        BsdfRec diff, rcb;
        std::memset(&diff, 0, sizeof diff); std::memset(&rcb, 0, sizeof rcb);
        diff.type = BSDF_DIFFUSE; diff.tex[TEX_REFLECTANCE] = tex(rho, 1, 1, 3);
        rcb.type = BSDF_ROUGHCONDUCTOR;
        rcb.tex[TEX_ALPHA_U] = tex(prm, 1, 1, 1); rcb.tex[TEX_ALPHA_V] = tex(prm + 1, 1, 1, 1); rcb.tex[TEX_ETA] = tex(prm + 2, 1, 1, 3); rcb.tex[TEX_K] = tex(prm + 5, 1, 1, 3);
        rcb.tex[TEX_SPECULAR] = tex(prm + 8, 1, 1, 3);
        Its its;
        its.wi = wi; its.uv = make_float2(0.f, 0.f); its.valid = true; its.shape = 0; its.tri = 0;
        const float3 smp = f3(s3[0], s3[1], s3[2]);
        ref_diffuse_eval(rho, awi, awo, r3);
        float3 e = bsdf_eval<false>(&diff, its, wo, true);
        expect(close(r3[0], e.x, 2) && close(r3[1], e.y, 2) && close(r3[2], e.z, 2), "diffuse eval", r3[0], e.x);
        e = bsdf_eval<true>(&diff, its, wo, true);   // the instantiation of diffuse-only scenes
        expect(close(r3[0], e.x, 2) && close(r3[1], e.y, 2) && close(r3[2], e.z, 2), "diffuse eval (SIMPLE)", r3[0], e.x);
        expect(close(ref_diffuse_pdf(rho, awi, awo), bsdf_pdf<false>(&diff, its, wo, true), 2), "diffuse pdf", ref_diffuse_pdf(rho, awi, awo), bsdf_pdf<false>(&diff, its, wo, true));
        int v = ref_diffuse_sample(rho, awi, s3, r4);
        BsdfSample bs = bsdf_sample<false>(&diff, its, smp, true);
        expect(v == (bs.valid ? 1 : 0) && close(r4[0], bs.wo.x, 4, 3e-7f) && close(r4[1], bs.wo.y, 4, 3e-7f) && close(r4[2], bs.wo.z, 4, 3e-7f) && close(r4[3], bs.pdf, 4, 3e-7f), "diffuse sample", r4[3], bs.pdf);
        ref_rc_eval(prm, awi, awo, r3);
        e = bsdf_eval<false>(&rcb, its, wo, true);
        expect(close(r3[0], e.x, 32, 1e-9f) && close(r3[1], e.y, 32, 1e-9f) && close(r3[2], e.z, 32, 1e-9f), "rough conductor eval", r3[0], e.x);
        if (wi.z > 0.f) {
            expect(close(ref_rc_pdf(prm, awi, awo), bsdf_pdf<false>(&rcb, its, wo, true), 32, 1e-9f), "rough conductor pdf", ref_rc_pdf(prm, awi, awo), bsdf_pdf<false>(&rcb, its, wo, true));
            const double wx = au * wi.x, wy = av * wi.y, wn = std::sqrt(wx * wx + wy * wy + (double)wi.z * wi.z);
            const float amp = (float)(1.0 / std::fmax(1e-6, (wx * wx + wy * wy) / (wn * wn))), tol = 2e-6f + 2.4e-7f * amp;
            ref_ggx_sample(au, av, awi, s3, r3);
            const float3 ms = ggx::sample(au, av, wi, s3[0], s3[1]);
            expect(close(r3[0], ms.x, 32, tol) && close(r3[1], ms.y, 32, tol) && close(r3[2], ms.z, 32, tol), "ggx::sample", r3[2], ms.z);
            v = ref_rc_sample(prm, awi, s3, r4);
            bs = bsdf_sample<false>(&rcb, its, smp, true);
            expect(v == (bs.valid ? 1 : 0), "rough conductor sample validity", (float)v, bs.valid ? 1.f : 0.f);
            if (v && bs.valid) {
                expect(close(r4[0], bs.wo.x, 32, tol) && close(r4[1], bs.wo.y, 32, tol) && close(r4[2], bs.wo.z, 32, tol), "rough conductor sample wo", r4[2], bs.wo.z);
                expect(std::fabs(r4[3] - bs.pdf) <= (1e-4f + 3e-6f * amp) * std::fmax(std::fabs(r4[3]), 1e-3f), "rough conductor sample pdf", r4[3], bs.pdf);
            }
        }
        // scene-box exit of an environment-map sample (utils.h:129-145)
        const float lo[3] = {-2, -3, -1}, hi[3] = {4, 2, 5};
        const float org[3] = {lo[0] + (hi[0] - lo[0]) * U(0.05f, 0.95f), lo[1] + (hi[1] - lo[1]) * U(0.05f, 0.95f), lo[2] + (hi[2] - lo[2]) * U(0.05f, 0.95f)};
        const float3 d = unit3();
        const float ad[3] = {d.x, d.y, d.z};
        float tng[5], t, G;
        float3 n;
        ref_ray_intersect_scene_aabb(org, ad, lo, hi, tng);
        ray_intersect_scene_aabb(f3(org[0], org[1], org[2]), d, f3(lo[0], lo[1], lo[2]), f3(hi[0], hi[1], hi[2]), t, n, G);
        expect(close(tng[0], t, 4) && tng[1] == n.x && tng[2] == n.y && tng[3] == n.z && close(tng[4], G, 4), "ray_intersect_scene_aabb", tng[0], t);
    }
    // bitmap lookups (bitmap.cpp:56-96): 1 and 3 channels, flipped and unflipped v, wrap-around, the clamped last texel
    for (int rep = 0; rep < 6; ++rep) {
        const int w = rep % 3 == 0 ? 2 : (rep % 3 == 1 ? 7 : 16), hgt = rep % 3 == 0 ? 2 : (rep % 3 == 1 ? 5 : 9), ch = rep < 3 ? 3 : 1, n = 600;
        std::vector<float> data((size_t)w * hgt * ch), uv(2 * n), out((size_t)n * ch);
        for (auto &x : data) x = U(0, 1);
        for (int i = 0; i < n; ++i) { uv[2 * i] = U(-1.5f, 2.5f); uv[2 * i + 1] = U(-1.5f, 2.5f); }
        uv[0] = 0.f; uv[1] = 0.f; uv[2] = 1.f; uv[3] = 1.f; uv[4] = 0.999999f; uv[5] = 0.5f; uv[6] = 0.5f; uv[7] = -1.f;
        const TexRef tr = tex(data.data(), w, hgt, ch);
        for (int flip = 0; flip < 2; ++flip) {
            if (ref_bitmap_eval(ch, w, hgt, data.data(), uv.data(), n, flip, out.data()) != 0) { std::printf("ref_bitmap_eval failed\n"); return 2; }
            for (int i = 0; i < n; ++i) {
                const float2 q = make_float2(uv[2 * i], uv[2 * i + 1]);
                if (ch == 3) { const float3 e = tex_eval3(tr, q, flip != 0); expect(close(out[3 * i], e.x, 4, 2e-7f) && close(out[3 * i + 1], e.y, 4, 2e-7f) && close(out[3 * i + 2], e.z, 4, 2e-7f), "tex_eval3", out[3 * i], e.x); }
                else { const float e = tex_eval1(tr, q, flip != 0); expect(close(out[i], e, 4, 2e-7f), "tex_eval1", out[i], e); }
            }
        }
    }
    // DiscreteDistribution::sample_reuse (pmf.cpp:30-50) on inclusive fp32 prefix sums
    for (int n : {1, 2, 7, 1000}) {
        std::vector<float> pmf(n), cmf(n);
        for (auto &x : pmf) x = U(0, 2);
        if (n > 3) pmf[2] = 0.f;
        float acc = 0.f;
        for (int i = 0; i < n; ++i) { acc += pmf[i]; cmf[i] = acc; }
        const int m = 500;
        std::vector<float> u(m), uo(m), pdf(m);
        std::vector<int> idx(m);
        for (auto &x : u) x = U(0, 1);
        u[0] = 0.f; u[1] = 0.5f; u[2] = 0.999999f;
        if (ref_discrete_sample(pmf.data(), n, u.data(), m, 1, idx.data(), pdf.data(), uo.data()) != 0) { std::printf("ref_discrete_sample failed\n"); return 2; }
        for (int i = 0; i < m; ++i) {
            float x = u[i], p;
            const int k = sample_reuse(cmf.data(), pmf.data(), n, acc, x, p);
            expect(k == idx[i] && close(pdf[i], p, 2) && close(uo[i], x, 4, 2e-6f), "sample_reuse", (float)idx[i], (float)k);
        }
    }
    if (argc > 3) {
        SYM(hr, ref_scene_load, void *(*)(const char *, const char *, int, int, int, int, int))
        SYM(hr, ref_scene_configure, int (*)(void *))
        SYM(hr, ref_scene_free, void (*)(void *))
        SYM(hr, ref_get_sensor, int (*)(void *, int, float *))
        SYM(hr, ref_sample_primary_ray, int (*)(void *, int, const float *, int, float *, float *))
        SYM(hr, ref_get_envmap, int (*)(void *, float *, float *))
        SYM(hr, ref_env_eval_direction, int (*)(void *, const float *, int, float *))
        const std::string dir = argv[3];
        for (const char *name : {"cbox_bunny", "tree", "bunny_env", "bunny_env_2"}) {
            void *sc = ref_scene_load((dir + "/data/scenes/" + name + ".xml").c_str(), dir.c_str(), 40, 24, 1, 0, 0);
            if (!sc || ref_scene_configure(sc) != 0) { std::printf("cannot load %s through the reference\n", name); return 2; }
            // camera rays (perspective.cpp:120-136) from the reference's own matrices
            float cam[55];
            ref_get_sensor(sc, 0, cam);
            SensorRec C;
            std::memset(&C, 0, sizeof C);
            for (int i = 0; i < 16; ++i) { C.sample_to_camera.m[i] = cam[i]; C.world_to_sample.m[i] = cam[16 + i]; C.to_world.m[i] = cam[32 + i]; }
            const int n = 2000;
            std::vector<float> smp(2 * n), ro(3 * n), rd(3 * n);
            for (auto &x : smp) x = U(0, 1);
            ref_sample_primary_ray(sc, 0, smp.data(), n, ro.data(), rd.data());
            for (int i = 0; i < n; ++i) {
                float3 o, d;
                sample_primary_ray(C, smp[2 * i], smp[2 * i + 1], o, d);
                expect(ro[3 * i] == o.x && ro[3 * i + 1] == o.y && ro[3 * i + 2] == o.z, "camera ray origin", ro[3 * i], o.x);
                expect(close(rd[3 * i], d.x, 4, 2e-7f) && close(rd[3 * i + 1], d.y, 4, 2e-7f) && close(rd[3 * i + 2], d.z, 4, 2e-7f), "camera ray direction", rd[3 * i + 2], d.z);
            }
            // environment map lookup by direction (envmap.cpp:42-58; bitmap.cpp:56-96 without the v flip)
            float env[19];
            if (std::strncmp(name, "bunny_env", 9) == 0 && ref_get_envmap(sc, env, nullptr) == 0) {
                const int w = (int)env[1], h = (int)env[2];
                std::vector<float> texels((size_t)w * h * 3);
                ref_get_envmap(sc, env, texels.data());
                EmitterRec em;
                std::memset(&em, 0, sizeof em);
                em.type = EMITTER_ENVMAP; em.env_radiance = tex(texels.data(), w, h, 3); em.env_scale = env[0];
                for (int i = 0; i < 16; ++i) em.env_from_world.m[i] = env[3 + i];
                std::vector<float> dirs(3 * n), out(3 * n);
                for (int i = 0; i < n; ++i) { const float3 d = unit3(); dirs[3 * i] = d.x; dirs[3 * i + 1] = d.y; dirs[3 * i + 2] = d.z; }
                dirs[0] = 0.f; dirs[1] = 1.f; dirs[2] = 0.f; dirs[3] = 0.f; dirs[4] = -1.f; dirs[5] = 0.f;   // the poles
                ref_env_eval_direction(sc, dirs.data(), n, out.data());
                for (int i = 0; i < n; ++i) {
                    const float3 e = env_eval_direction(em, f3(dirs[3 * i], dirs[3 * i + 1], dirs[3 * i + 2]));
                    // a last-bit difference in atan2 / acos moves the lookup by 1e-7 of the map; neighbouring texels of an HDR map differ
                    const float tol = 2e-4f * std::fmax(1.f, std::fmax(out[3 * i], std::fmax(out[3 * i + 1], out[3 * i + 2])));
                    expect(close(out[3 * i], e.x, 64, tol) && close(out[3 * i + 1], e.y, 64, tol) && close(out[3 * i + 2], e.z, 64, tol), "env_eval_direction", out[3 * i], e.x);
                }
            }
            ref_scene_free(sc);
        }
    }
    std::printf("ref_shade_check: %s (%d comparisons, %d mismatches)\n", bad ? "FAILED" : "ok", checked, bad);
    return bad ? 1 : 0;
}
