// ORACLE — TEST INFRASTRUCTURE ONLY (parity pinned against the reference's own source run on the CPU; see orc_math.hpp header and DESIGN.md §2).
//
// C entry points over the CPU restatement so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg can drive it through ctypes. The product never links or loads this library.
#include <chrono>
#include <cstdio>
#include <omp.h>

#include "orc_render.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;
struct Handle {
    Scene scene;
};
M4f mat16(const float *m) {
    M4f r;
    if (m) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = m[4 * i + j];
    return r;
}
Bitmap bitmap(const float *d, int w, int h, int c) {
    Bitmap b;
    b.w = w; b.h = h; b.c = c;
    b.data.assign(d, d + (size_t)w * h * c);
    return b;
}
template <class F> int guard(F &&f) {
    try { f(); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return 1; }
}
Bitmap *bsdf_slot(Bsdf &b, int which) {
    switch (which) {
        case 0: return &b.reflectance;
        case 1: return &b.alpha_u;
        case 2: return &b.alpha_v;
        case 3: return &b.eta;
        case 4: return &b.k;
        default: return &b.specular_reflectance;
    }
}
}  // namespace

extern "C" {

const char *orc_last_error() { return g_err.c_str(); }
int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

void *orc_scene_new() { return new Handle(); }
void orc_scene_free(void *h) { delete (Handle *)h; }

int orc_set_options(void *h, int w, int hh, int spp, int sppe, int sppse) {
    Scene &s = ((Handle *)h)->scene;
    s.opts.width = w; s.opts.height = hh; s.opts.spp = spp; s.opts.sppe = sppe; s.opts.sppse = sppse;
    return 0;
}
// which: 0 reflectance(3) 1 alpha_u(1) 2 alpha_v(1) 3 eta(3) 4 k(3) 5 specular_reflectance(3)
int orc_add_bsdf(void *h, int type) {
    Scene &s = ((Handle *)h)->scene;
    Bsdf b;
    b.type = type;
    s.bsdfs.push_back(b);
    return (int)s.bsdfs.size() - 1;
}
int orc_set_bsdf_texture(void *h, int bsdf, int which, const float *data, int w, int hh) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        Bitmap *slot = bsdf_slot(s.bsdfs.at(bsdf), which);
        *slot = bitmap(data, w, hh, slot->c);
    });
}
int orc_set_bsdf_tangent(void *h, int bsdf, int which, const float *tang) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        Bitmap *slot = bsdf_slot(s.bsdfs.at(bsdf), which);
        if (tang) slot->tang.assign(tang, tang + slot->data.size()); else slot->tang.clear();
    });
}
int orc_add_mesh(void *h, int nv, int nf, const float *verts, const int *faces, int nuv, const float *uvs, const int *uv_faces,
                 int face_normals, int enable_edges, int bsdf, const float *to_world) {
    Scene &s = ((Handle *)h)->scene;
    Mesh m;
    m.nv = nv; m.nf = nf;
    m.vraw.assign(verts, verts + 3 * (size_t)nv);
    m.faces.assign(faces, faces + 3 * (size_t)nf);
    m.has_uv = nuv > 0;
    if (m.has_uv) { m.uvs.assign(uvs, uvs + 2 * (size_t)nuv); m.uv_faces.assign(uv_faces, uv_faces + 3 * (size_t)nf); }
    m.face_normals = face_normals != 0;
    m.enable_edges = enable_edges != 0;
    m.bsdf = bsdf;
    m.to_world_raw = mat16(to_world);
    int rc = guard([&] { build_edge_list(m); });
    if (rc) return -1;
    s.meshes.push_back(std::move(m));
    return (int)s.meshes.size() - 1;
}
int orc_set_mesh_vertices(void *h, int mesh, const float *verts) {
    return guard([&] { Mesh &m = ((Handle *)h)->scene.meshes.at(mesh); m.vraw.assign(verts, verts + 3 * (size_t)m.nv); });
}
int orc_set_mesh_transform(void *h, int mesh, const float *mat, int left) {
    return guard([&] { Mesh &m = ((Handle *)h)->scene.meshes.at(mesh); (left ? m.left : m.right) = mat16(mat); });
}
int orc_set_mesh_vertex_tangent(void *h, int mesh, const float *tang) {
    return guard([&] {
        Mesh &m = ((Handle *)h)->scene.meshes.at(mesh);
        if (tang) m.vraw_t.assign(tang, tang + 3 * (size_t)m.nv); else m.vraw_t.clear();
    });
}
int orc_set_mesh_transform_tangent(void *h, int mesh, const float *tang, int left) {
    return guard([&] {
        Mesh &m = ((Handle *)h)->scene.meshes.at(mesh);
        M4f t;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) t.m[i][j] = tang ? tang[4 * i + j] : 0.f;
        if (left) { m.left_t = t; m.has_left_t = tang != nullptr; } else { m.right_t = t; m.has_right_t = tang != nullptr; }
    });
}
int orc_add_area_emitter(void *h, int mesh, const float *radiance) {
    Scene &s = ((Handle *)h)->scene;
    Emitter e;
    e.type = EMITTER_AREA; e.mesh = mesh; e.radiance = V3f(radiance[0], radiance[1], radiance[2]);
    s.emitters.push_back(e);
    s.meshes.at(mesh).emitter = (int)s.emitters.size() - 1;
    return (int)s.emitters.size() - 1;
}
int orc_add_envmap(void *h, int w, int hh, const float *rgb, float scale, const float *to_world) {
    Scene &s = ((Handle *)h)->scene;
    Emitter e;
    e.type = EMITTER_ENVMAP;
    e.env_radiance = bitmap(rgb, w, hh, 3);
    e.env_scale = scale;
    e.env_to_world_raw = mat16(to_world);
    s.emitters.push_back(e);
    s.emitter_env = (int)s.emitters.size() - 1;
    return s.emitter_env;
}
/* forward-mode tangents of the environment map's radiance texels (w*h*3, may be null) and of its scale */
int orc_set_envmap_tangent(void *h, const float *radiance_t, float scale_t) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        Emitter &e = s.emitters.at(s.emitter_env);
        if (radiance_t) e.env_radiance.tang.assign(radiance_t, radiance_t + e.env_radiance.data.size()); else e.env_radiance.tang.clear();
        e.env_scale_t = scale_t;
    });
}
/* forward-mode tangent of the matrix EnvironmentMap.set_transform sets (to_world = left * raw, envmap.cpp:23); null clears it */
int orc_set_envmap_transform_tangent(void *h, const float *tang) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        Emitter &e = s.emitters.at(s.emitter_env);
        e.env_has_t = tang != nullptr;
        if (tang) e.env_left_t = mat16(tang);
    });
}
int orc_set_envmap_transform(void *h, const float *left) {
    return guard([&] { Scene &s = ((Handle *)h)->scene; s.emitters.at(s.emitter_env).env_left = mat16(left); });
}
/* forward-mode tangent of a mesh's texture coordinates (same layout as the uv array given to orc_add_mesh); null clears it */
int orc_set_mesh_uv_tangent(void *h, int mesh, const float *tang) {
    return guard([&] {
        Mesh &m = ((Handle *)h)->scene.meshes.at(mesh);
        if (tang) m.uvs_t.assign(tang, tang + m.uvs.size()); else m.uvs_t.clear();
    });
}
/* forward-mode tangent of a sensor's to_world (Sensor.to_world is an AD leaf, src/psdr.cpp:220-224); null clears it */
int orc_set_sensor_transform_tangent(void *h, int sensor, const float *tang) {
    return guard([&] {
        Sensor &c = ((Handle *)h)->scene.sensors.at(sensor);
        c.has_t = tang != nullptr;
        if (tang) c.to_world_t = mat16(tang);
    });
}
int orc_add_sensor(void *h, float fov_x, float near_clip, float far_clip, const float *to_world) {
    Scene &s = ((Handle *)h)->scene;
    Sensor c;
    c.fov_x = fov_x; c.near_clip = near_clip; c.far_clip = far_clip; c.to_world = mat16(to_world);
    s.sensors.push_back(c);
    return (int)s.sensors.size() - 1;
}
int orc_configure(void *h) { return guard([&] { ((Handle *)h)->scene.configure(); }); }

// ---- introspection of configured tables (for parity tests of the product's configure kernels) ------------
int orc_num_triangles(void *h) { return (int)((Handle *)h)->scene.tri.size(); }
int orc_num_meshes(void *h) { return (int)((Handle *)h)->scene.meshes.size(); }
// 22 floats per triangle: p0 e1 e2 n0 n1 n2 face_normal face_area (types.h:136-146)
int orc_get_triangle_info(void *h, float *out) {
    Scene &s = ((Handle *)h)->scene;
    for (size_t i = 0; i < s.tri.size(); ++i) {
        const auto &t = s.tri[i];
        const V3<Dual> *v[7] = {&t.p0, &t.e1, &t.e2, &t.n0, &t.n1, &t.n2, &t.face_normal};
        for (int k = 0; k < 7; ++k) for (int c = 0; c < 3; ++c) out[22 * i + 3 * k + c] = (*v[k])[c].v;
        out[22 * i + 21] = t.face_area.v;
    }
    return 0;
}
int orc_mesh_num_edges(void *h, int mesh) { return (int)((Handle *)h)->scene.meshes.at(mesh).edges.size() / 5; }
int orc_mesh_get_edges(void *h, int mesh, int *out) {
    const auto &e = ((Handle *)h)->scene.meshes.at(mesh).edges;
    std::copy(e.begin(), e.end(), out);
    return 0;
}
int orc_num_sec_edges(void *h) { return (int)((Handle *)h)->scene.sec_edges.size(); }
// 16 floats per secondary edge: p0 e1 n0 n1 p2 is_boundary (edge.h:50-65)
int orc_get_sec_edges(void *h, float *out) {
    Scene &s = ((Handle *)h)->scene;
    for (size_t i = 0; i < s.sec_edges.size(); ++i) {
        const auto &e = s.sec_edges[i];
        const V3<Dual> *v[5] = {&e.p0, &e.e1, &e.n0, &e.n1, &e.p2};
        for (int k = 0; k < 5; ++k) for (int c = 0; c < 3; ++c) out[16 * i + 3 * k + c] = (*v[k])[c].v;
        out[16 * i + 15] = e.is_boundary ? 1.f : 0.f;
    }
    return 0;
}
int orc_num_primary_edges(void *h, int sensor) { return (int)((Handle *)h)->scene.sensors.at(sensor).edges.size(); }
// 7 floats per primary edge: p0.xy p1.xy edge_normal.xy edge_length (edge.h:28-40)
int orc_get_primary_edges(void *h, int sensor, float *out) {
    const auto &ed = ((Handle *)h)->scene.sensors.at(sensor).edges;
    for (size_t i = 0; i < ed.size(); ++i) {
        out[7 * i + 0] = ed[i].p0.x.v; out[7 * i + 1] = ed[i].p0.y.v; out[7 * i + 2] = ed[i].p1.x.v; out[7 * i + 3] = ed[i].p1.y.v;
        out[7 * i + 4] = ed[i].edge_normal.x; out[7 * i + 5] = ed[i].edge_normal.y; out[7 * i + 6] = ed[i].edge_length;
    }
    return 0;
}
// camera: sample_to_camera(16) world_to_sample(16) to_world(16) camera_pos(3) camera_dir(3) inv_area(1)
int orc_get_sensor(void *h, int sensor, float *out) {
    const Sensor &c = ((Handle *)h)->scene.sensors.at(sensor);
    const M4<Dual> *m[3] = {&c.sample_to_camera, &c.world_to_sample, &c.to_world_d};
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[16 * k + 4 * i + j] = m[k]->m[i][j].v;
    for (int k = 0; k < 3; ++k) { out[48 + k] = c.camera_pos[k].v; out[51 + k] = c.camera_dir[k].v; }
    out[54] = c.inv_area.v;
    return 0;
}

// ---- RNG known-answer helpers ---------------------------------------------------------------------------
void orc_pcg32_kat(uint64_t initstate, uint64_t initseq, int n, uint32_t *out) {
    PCG32 r;
    r.seed(initstate, initseq);
    for (int i = 0; i < n; ++i) out[i] = r.next_uint32();
}
void orc_sampler_kat(uint64_t lane, int n, uint32_t *out_u32, float *out_f32) {
    SamplerLane a = SamplerLane::make(lane), b = SamplerLane::make(lane);
    for (int i = 0; i < n; ++i) { out_u32[i] = a.rng.next_uint32(); out_f32[i] = b.next_1d(); }
}
int orc_discrete_sample_reuse(const float *pmf, int n, const float *u_in, int m, int *idx, float *pdf, float *u_out) {
    DiscreteDistribution d;
    d.init(std::vector<float>(pmf, pmf + n));
    for (int i = 0; i < m; ++i) { float u = u_in[i]; auto pr = d.sample_reuse(u); idx[i] = pr.first; pdf[i] = pr.second; u_out[i] = u; }
    return 0;
}

// ---- closest hit (cuda/psdr_cuda.cu:9-45 contract: tri id, shape id, u, v; -1 on miss) -------------------------
int orc_trace(void *h, int64_t n, const float *o, const float *d, const float *tmax, int *tri, int *shape, float *u, float *v, float *t, int brute) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        if (!s.ready) throw std::runtime_error("Input scene must be configured!");
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < n; ++i) {
            Ray<float> r(V3f(o[3 * i], o[3 * i + 1], o[3 * i + 2]), V3f(d[3 * i], d[3 * i + 1], d[3 * i + 2]));
            r.tmax = tmax ? tmax[i] : kInf;
            Hit hit = brute ? s.accel.closest_brute(r) : s.accel.closest(r);
            tri[i] = hit.tri; shape[i] = hit.tri >= 0 ? s.tri_mesh[hit.tri] : -1;
            u[i] = hit.u; v[i] = hit.v;
            if (t) t[i] = hit.t;
        }
    });
}

// ---- integrators -------------------------------------------------------------------------------------------
void *orc_integrator_new(int kind, int bsdf_samples, int light_samples, int hide_emitters, int field, int max_depth) {
    Integrator *I = new Integrator();
    I->kind = kind; I->bsdf_samples = bsdf_samples; I->light_samples = light_samples;
    I->hide_emitters = hide_emitters != 0; I->field = field; I->max_depth = max_depth;
    return I;
}
void orc_integrator_free(void *I) { delete (Integrator *)I; }
int orc_render_c(void *h, void *I, int sensor, float *out) {
    return guard([&] { renderC(*(Integrator *)I, ((Handle *)h)->scene, sensor, out); });
}
int orc_render_d(void *h, void *I, int sensor, float *out, float *out_t) {
    return guard([&] { renderD(*(Integrator *)I, ((Handle *)h)->scene, sensor, out, out_t); });
}
// debugging aid: radiance of one interior lane of a freshly seeded sampler (stream position 0), C or D formulation
int orc_debug_lane(void *h, void *Iv, int sensor, int64_t lane, int ad, float *out3) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        const Integrator &I = *(Integrator *)Iv;
        const RenderOption &o = s.opts;
        SamplerLane smp = SamplerLane::make((uint64_t)lane);
        int pix = (int)(lane / o.spp);
        V2f j = smp.next_2d();
        float bx = (float)(pix % o.width), by = (float)(pix / o.width);
        if (ad) {
            V2<Dual> smpl(Dual((bx + j.x) / (float)o.width), Dual((by + j.y) / (float)o.height));
            Ray<Dual> ray = s.sensors[sensor].sample_primary_ray<Dual>(smpl);
            V3<Dual> v = Li<Dual>(I, s, smp, ray);
            for (int k = 0; k < 3; ++k) out3[k] = v[k].v;
        } else {
            V2f smpl((bx + j.x) / (float)o.width, (by + j.y) / (float)o.height);
            Ray<float> ray = s.sensors[sensor].sample_primary_ray<float>(smpl);
            V3f v = Li<float>(I, s, smp, ray);
            for (int k = 0; k < 3; ++k) out3[k] = v[k];
        }
    });
}
// same lane in the D formulation, value and forward-mode tangent (for the tangents set on the scene)
int orc_debug_lane_d(void *h, void *Iv, int sensor, int64_t lane, float *out3, float *out3_t) {
    return guard([&] {
        Scene &s = ((Handle *)h)->scene;
        const Integrator &I = *(Integrator *)Iv;
        const RenderOption &o = s.opts;
        SamplerLane smp = SamplerLane::make((uint64_t)lane);
        int pix = (int)(lane / o.spp);
        V2f j = smp.next_2d();
        float bx = (float)(pix % o.width), by = (float)(pix / o.width);
        V2<Dual> smpl(Dual((bx + j.x) / (float)o.width), Dual((by + j.y) / (float)o.height));
        Ray<Dual> ray = s.sensors[sensor].sample_primary_ray<Dual>(smpl);
        V3<Dual> v = Li<Dual>(I, s, smp, ray);
        for (int k = 0; k < 3; ++k) { out3[k] = v[k].v; out3_t[k] = v[k].d; }
    });
}
int orc_preprocess_secondary_edges(void *h, void *I, int sensor, const int *reso4, int nrounds) {
    return guard([&] { preprocess_secondary_edges(*(Integrator *)I, ((Handle *)h)->scene, sensor, reso4, nrounds); });
}
// ---- the oracle's math, one call per reference function, for tests/test_ref_math.py (pinned against oracle/_ref/libref_math.so) ----
static V3f mv3(const float *p) { return V3f(p[0], p[1], p[2]); }
static void mput(float *o, const V3f &v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
void orc_math_square_to_uniform_disk_concentric(const float *s, float *out) { V2f r = square_to_uniform_disk_concentric(V2f(s[0], s[1])); out[0] = r.x; out[1] = r.y; }
void orc_math_square_to_cosine_hemisphere(const float *s, float *out) { mput(out, square_to_cosine_hemisphere(V2f(s[0], s[1]))); }
void orc_math_square_to_uniform_triangle(const float *s, float *out) { V2f r = square_to_uniform_triangle(V2f(s[0], s[1])); out[0] = r.x; out[1] = r.y; }
void orc_math_frame(const float *n, float *s_out, float *t_out) { Frame<float> f(mv3(n)); mput(s_out, f.s); mput(t_out, f.t); }
void orc_math_frame_to_local(const float *n, const float *v, float *out) { Frame<float> f(mv3(n)); mput(out, f.to_local(mv3(v))); }
void orc_math_frame_to_world(const float *n, const float *v, float *out) { Frame<float> f(mv3(n)); mput(out, f.to_world(mv3(v))); }
void orc_math_ray_intersect_triangle(const float *p0, const float *e1, const float *e2, const float *o, const float *d, float *uvt) {
    ray_intersect_triangle<float>(mv3(p0), mv3(e1), mv3(e2), Ray<float>(mv3(o), mv3(d)), uvt[0], uvt[1], uvt[2]);
}
void orc_math_bilinear(const float *p0, const float *e1, const float *e2, const float *st, float *out) { mput(out, bilinear<float>(mv3(p0), mv3(e1), mv3(e2), V2f(st[0], st[1]))); }
float orc_math_rgb2luminance(const float *rgb) { return rgb2luminance(mv3(rgb)); }
int orc_math_sign_eps(float x, float eps) { return sign_eps(x, eps); }
void orc_math_fresnel(const float *eta, const float *k, float cos_theta_i, float *out) { mput(out, fresnel<float>(mv3(eta), mv3(k), cos_theta_i)); }
void orc_math_ray_intersect_scene_aabb(const float *o, const float *d, const float *lo, const float *hi, float *t_n_G) {
    V3f n; ray_intersect_scene_aabb(mv3(o), mv3(d), mv3(lo), mv3(hi), t_n_G[0], n, t_n_G[4]); mput(t_n_G + 1, n);
}
float orc_math_ggx_eval(float au, float av, const float *m) { return ggx::eval<float>(au, av, mv3(m)); }
float orc_math_ggx_smith_g1(float au, float av, const float *v, const float *m) { return ggx::smith_g1<float>(au, av, mv3(v), mv3(m)); }
void orc_math_ggx_sample(float au, float av, const float *wi, const float *s3, float *out) { mput(out, ggx::sample<float>(au, av, mv3(wi), mv3(s3))); }
void orc_math_ggx_sample_visible_11(float cos_theta_i, const float *s2, float *out) { V2f r = ggx::sample_visible_11<float>(cos_theta_i, V2f(s2[0], s2[1])); out[0] = r.x; out[1] = r.y; }

void orc_math_sampler_lane(uint64_t lane, int n, float *out_1d, float *out_2d, float *out_3d) {
    SamplerLane s = SamplerLane::make(lane);
    for (int i = 0; i < n; ++i) out_1d[i] = s.next_1d();
    V2f a = s.next_2d(); out_2d[0] = a.x; out_2d[1] = a.y;
    mput(out_3d, s.next_3d());
}
// BSDFs with constant textures through the oracle's Scene::bsdf_* (one mesh, one BSDF)
static Scene &bsdf_scene(int type, const float *prm) {
    static thread_local Scene s;
    if (s.meshes.empty()) { s.meshes.emplace_back(); s.meshes[0].bsdf = 0; s.bsdfs.emplace_back(); }
    Bsdf &b = s.bsdfs[0];
    b.type = type;
    if (type == BSDF_DIFFUSE) b.reflectance = Bitmap::constant3(prm[0], prm[1], prm[2]);
    else {
        b.alpha_u = Bitmap::constant1(prm[0]); b.alpha_v = Bitmap::constant1(prm[1]);
        b.eta = Bitmap::constant3(prm[2], prm[3], prm[4]); b.k = Bitmap::constant3(prm[5], prm[6], prm[7]);
        b.specular_reflectance = Bitmap::constant3(prm[8], prm[9], prm[10]);
    }
    return s;
}
static Intersection<float> its_wi(const float *wi) { Intersection<float> its; its.shape = 0; its.tri = 0; its.wi = mv3(wi); its.uv = V2f(0.f, 0.f); return its; }
void orc_math_bsdf_eval(int type, const float *prm, const float *wi, const float *wo, float *out) { mput(out, bsdf_scene(type, prm).bsdf_eval<float>(its_wi(wi), mv3(wo), true)); }
float orc_math_bsdf_pdf(int type, const float *prm, const float *wi, const float *wo) { return bsdf_scene(type, prm).bsdf_pdf<float>(its_wi(wi), mv3(wo), true); }
int orc_math_bsdf_sample(int type, const float *prm, const float *wi, const float *s3, float *wo_pdf) {
    BSDFSample<float> bs = bsdf_scene(type, prm).bsdf_sample<float>(its_wi(wi), mv3(s3), true);
    mput(wo_pdf, bs.wo); wo_pdf[3] = bs.pdf;
    return bs.valid ? 1 : 0;
}

int orc_reseed(void *h) {   // drop sampler state so that the next configure() starts the streams afresh
    Scene &s = ((Handle *)h)->scene;
    for (auto &v : s.samplers) v.clear();
    return 0;
}

}  // extern "C"
