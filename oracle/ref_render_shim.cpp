// TEST INFRASTRUCTURE (oracle/): C entry points around psdr-cuda's OWN renderer, compiled unmodified from where it lies under
// /root/reference (every src/**/*.cpp except the pybind module, the OptiX glue and the embedded PTX) against the CPU stand-in for
// Enoki in oracle/ref_dyn/. Built by oracle/build_ref.sh into oracle/_ref/libref_render.so; tests/test_ref_render.py runs the
// reference's Scene::load_file / configure / Integrator::renderC / renderD through it and pins the oracle to the result.
//
// Two things are NOT the reference's here, because they are external to it:
//  * Enoki  -> oracle/ref_dyn/enoki_dyn.h (host arrays; forward-mode tangents instead of the reverse-mode tape),
//  * OptiX  -> Scene_OptiX below (scene_optix.cpp replaced): the closest hit over all triangles with t in (RayEpsilon, tmax),
//              computed with the arithmetic of the reference's ray_intersect_triangle (utils.h:67-77), ties to the lowest triangle id
//              (a small BVH only prunes; it returns what testing every triangle returns);
//              it fills the same outputs as cuda/psdr_cuda.cu:27-45 (global triangle id, shape id, barycentrics; -1 on a miss).
#include <unistd.h>

#include <cmath>
#include <cstring>
#include <string>

#define protected public   // the entry points below read Scene's tables (m_triangle_info, m_sec_edge_info) for the table parity tests
#include <misc/Exception.h>
#include <psdr/psdr.h>
#include <psdr/core/ray.h>
#include <psdr/core/intersection.h>
#include <psdr/core/sampler.h>
#include <psdr/core/bitmap.h>
#include <psdr/core/transform.h>
#include <psdr/bsdf/diffuse.h>
#include <psdr/bsdf/roughconductor.h>
#include <psdr/emitter/area.h>
#include <psdr/emitter/envmap.h>
#include <psdr/sensor/perspective.h>
#include <psdr/shape/mesh.h>
#include <psdr/scene/scene_optix.h>
#include <psdr/scene/scene.h>
#include <psdr/integrator/direct.h>
#include <psdr/integrator/field.h>
#undef protected

// ---- stand-in for the OptiX acceleration structure -----------------------------------------------------------------------------------------
struct PathTracerState {
    struct Tri { float p0[3], e1[3], e2[3]; int shape, id; };
    std::vector<Tri> tris;                 // in BVH leaf order; `id` is the global triangle id
    struct Node { float lo[3], hi[3]; int left, right, first, count; };   // leaf: count > 0
    std::vector<Node> nodes;               // a median-split BVH: an accelerator only, boxes padded so that no hit is lost
    int build(int first, int count) {
        Node n;
        for (int c = 0; c < 3; ++c) { n.lo[c] = std::numeric_limits<float>::max(); n.hi[c] = -std::numeric_limits<float>::max(); }
        float clo[3] = {n.lo[0], n.lo[1], n.lo[2]}, chi[3] = {n.hi[0], n.hi[1], n.hi[2]};
        for (int i = first; i < first + count; ++i) for (int c = 0; c < 3; ++c) {
            const Tri &t = tris[i];
            const float a = t.p0[c], b = t.p0[c] + t.e1[c], d = t.p0[c] + t.e2[c], ce = (a + b + d) / 3.f;
            n.lo[c] = std::min(n.lo[c], std::min(a, std::min(b, d))); n.hi[c] = std::max(n.hi[c], std::max(a, std::max(b, d)));
            clo[c] = std::min(clo[c], ce); chi[c] = std::max(chi[c], ce);
        }
        for (int c = 0; c < 3; ++c) { const float pad = 1e-4f * (1.f + std::abs(n.lo[c]) + std::abs(n.hi[c])); n.lo[c] -= pad; n.hi[c] += pad; }
        n.left = n.right = -1; n.first = first; n.count = count;
        const int me = (int)nodes.size();
        nodes.push_back(n);
        if (count > 4) {
            int axis = 0;
            for (int c = 1; c < 3; ++c) if (chi[c] - clo[c] > chi[axis] - clo[axis]) axis = c;
            const int mid = first + count / 2;
            std::nth_element(tris.begin() + first, tris.begin() + mid, tris.begin() + first + count, [axis](const Tri &x, const Tri &y) {
                return x.p0[axis] * 3.f + x.e1[axis] + x.e2[axis] < y.p0[axis] * 3.f + y.e1[axis] + y.e2[axis]; });
            const int l = build(first, mid - first), r = build(mid, first + count - mid);
            nodes[me].left = l; nodes[me].right = r; nodes[me].count = 0;
        }
        return me;
    }
};

namespace psdr {

// the per-launch output buffers of scene_optix.h:10-17, (re)allocated when the wavefront size changes
void Intersection_OptiX::reserve(int64_t size) {
    PSDR_ASSERT(size > 0);
    if (m_size == size) return;
    triangle_id = shape_id = zero<IntC>(size);
    uv = zero<Vector2fC>(size);
    m_size = size;
}

Scene_OptiX::Scene_OptiX() { m_accel = nullptr; }
Scene_OptiX::~Scene_OptiX() { delete m_accel; }

void Scene_OptiX::configure(const std::vector<Mesh *> &meshes) {
    PSDR_ASSERT(!meshes.empty());
    if (m_accel == nullptr) m_accel = new PathTracerState();
    m_accel->tris.clear();
    m_accel->nodes.clear();
    int offset = 0;
    for (size_t s = 0; s < meshes.size(); ++s) {
        const Mesh *mesh = meshes[s];
        PSDR_ASSERT(static_cast<int>(slices(mesh->m_vertex_buffer)) == mesh->m_num_vertices * 3);
        PSDR_ASSERT(static_cast<int>(slices(mesh->m_face_buffer)) == mesh->m_num_faces * 3);
        const float *vb = mesh->m_vertex_buffer.data();   // what scene_optix.cpp:48-63 hands to OptiX
        const int *fb = mesh->m_face_buffer.data();
        for (int f = 0; f < mesh->m_num_faces; ++f) {
            PathTracerState::Tri t;
            const float *a = vb + 3 * fb[3 * f], *b = vb + 3 * fb[3 * f + 1], *c = vb + 3 * fb[3 * f + 2];
            for (int k = 0; k < 3; ++k) { t.p0[k] = a[k]; t.e1[k] = b[k] - a[k]; t.e2[k] = c[k] - a[k]; }
            t.shape = (int)s; t.id = offset + f;
            m_accel->tris.push_back(t);
        }
        offset += mesh->m_num_faces;
    }
    m_accel->build(0, (int)m_accel->tris.size());
}

bool Scene_OptiX::is_ready() const { return m_accel != nullptr; }

static inline float dot3(const float *a, const float *b) { return std::fma(a[0], b[0], std::fma(a[1], b[1], a[2] * b[2])); }
static inline void cross3(const float *a, const float *b, float *r) {
    r[0] = std::fma(a[1], b[2], -(a[2] * b[1])); r[1] = std::fma(a[2], b[0], -(a[0] * b[2])); r[2] = std::fma(a[0], b[1], -(a[1] * b[0]));
}
static bool hits_box(const PathTracerState::Node &b, const float *o, const float *d, float tmax) {
    float t0 = 0.f, t1 = tmax;
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.f) { if (o[k] < b.lo[k] || o[k] > b.hi[k]) return false; continue; }
        float a = (b.lo[k] - o[k]) / d[k], c = (b.hi[k] - o[k]) / d[k];
        if (a > c) std::swap(a, c);
        t0 = std::max(t0, a * 0.999f - 1e-3f); t1 = std::min(t1, c * 1.001f + 1e-3f);
        if (!(t0 <= t1)) return false;
    }
    return true;
}

template <bool ad>
Vector2i<ad> Scene_OptiX::ray_intersect(const Ray<ad> &ray, Mask<ad> &active) const {
    const int m = static_cast<int>(slices(ray.o));
    m_its.reserve(m);
    const RayC r = [&] { if constexpr (ad) return detach(ray); else return ray; }();
    int *tri_out = m_its.triangle_id.data(), *shape_out = m_its.shape_id.data();
    float *u_out = m_its.uv.x().data(), *v_out = m_its.uv.y().data();
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < m; ++i) {
        tri_out[i] = shape_out[i] = -1; u_out[i] = v_out[i] = -1.f;   // __miss__psdr_ms
        if (!bool(lval(active, i))) continue;                       // the caller masks these lanes out anyway (active &= hit)
        const float o[3] = {lval(r.o.x(), i), lval(r.o.y(), i), lval(r.o.z(), i)}, d[3] = {lval(r.d.x(), i), lval(r.d.y(), i), lval(r.d.z(), i)};
        float best = lval(r.tmax, i);
        int stack[64], sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const auto &node = m_accel->nodes[stack[--sp]];
            if (!hits_box(node, o, d, best)) continue;
            if (node.count == 0) { stack[sp++] = node.left; stack[sp++] = node.right; continue; }
            for (int k = node.first; k < node.first + node.count; ++k) {
                const auto &t = m_accel->tris[k];
                float h[3], s[3], q[3];
                cross3(d, t.e2, h);                                  // utils.h:68-76
                const float a = dot3(t.e1, h), f = 1.f / a;
                for (int c = 0; c < 3; ++c) s[c] = o[c] - t.p0[c];
                const float u = f * dot3(s, h);
                cross3(s, t.e1, q);
                const float v = f * dot3(d, q), tt = f * dot3(t.e2, q);
                // the closest hit; equal distances go to the lowest triangle id, whatever order the leaves are visited in
                if (u >= 0.f && v >= 0.f && u + v <= 1.f && tt > RayEpsilon && (tt < best || (tt == best && tri_out[i] >= 0 && t.id < tri_out[i]))) {
                    best = tt; tri_out[i] = t.id; shape_out[i] = t.shape; u_out[i] = u; v_out[i] = v;
                }
            }
        }
    }
    active &= (m_its.shape_id >= 0) && (m_its.triangle_id >= 0);
    return Vector2i<ad>(m_its.shape_id, m_its.triangle_id);
}
template Vector2iC Scene_OptiX::ray_intersect(const RayC &ray, MaskC &active) const;
template Vector2iD Scene_OptiX::ray_intersect(const RayD &ray, MaskD &active) const;

}  // namespace psdr

// ---- C entry points -------------------------------------------------------------------------------------------------------------------------
using namespace psdr;
static thread_local std::string g_err;
template <class F> static int guard(F &&f) {
    try { f(); return 0; } catch (const std::exception &e) { g_err = e.what(); return 1; }
}
static void set_tangent(FloatD &x, const float *t, size_t n, size_t stride, size_t off) {
    if (!t) { x.g = FloatC(); return; }
    if (x.size() != n) throw Exception("tangent size mismatch: array has " + std::to_string(x.size()) + " lanes, tangent " + std::to_string(n));
    x.g.d.resize(n);
    for (size_t i = 0; i < n; ++i) x.g.d[i] = t[i * stride + off];
}
template <size_t k> static void set_tangent(Array<FloatD, k> &x, const float *t, size_t n) { for (size_t c = 0; c < k; ++c) set_tangent(x[c], t, n, k, c); }
static void set_tangent(Matrix4fD &m, const float *t) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) set_tangent(m(i, j), t ? t + 4 * i + j : nullptr, 1, 1, 0); }

// what SceneLoader made of an XML file (before configure): one line per object, numbers printed with %.9g, for comparing loaders
static void put_mat(std::ostringstream &os, const Matrix4fD &m) { char b[32]; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { snprintf(b, sizeof b, " %.9g", (double)m(i, j)[0]); os << b; } }
template <class V> static void put_tex(std::ostringstream &os, const char *name, const V &data, const ScalarVector2i &reso, int ch) {
    char b[32];
    os << " " << name << "[" << reso.x() << "x" << reso.y() << "]";
    if (reso.x() == 1 && reso.y() == 1) for (int c = 0; c < ch; ++c) { float v; if constexpr (is_arr_v<V>) v = data[c][0]; else v = data[0]; snprintf(b, sizeof b, " %.9g", (double)v); os << b; }
}

extern "C" {

const char *ref_last_error() { return g_err.c_str(); }
void ref_set_matvec_plain(int on) { enoki::matvec_plain() = on != 0; }   // see oracle/ref_dyn/enoki_dyn.h
void ref_set_inverse_rounded(int on) { enoki::inverse_rounded() = on != 0; }
void ref_set_dot_from_first(int on) { enoki::dot_from_first() = on != 0; }

// Scene::load_file(xml, false) with the working directory the scene's relative paths expect, RenderOption overrides, then configure()
void *ref_scene_load(const char *xml_path, const char *cwd, int width, int height, int spp, int sppe, int sppse) {
    Scene *scene = nullptr;
    if (guard([&] {
            char old[4096];
            if (!getcwd(old, sizeof old)) throw Exception("getcwd failed");
            if (cwd && chdir(cwd) != 0) throw Exception(std::string("cannot chdir to ") + cwd);
            try {
                scene = new Scene();
                scene->load_file(xml_path, false);
            } catch (...) { (void)!chdir(old); delete scene; scene = nullptr; throw; }
            (void)!chdir(old);
            if (width > 0) { scene->m_opts.width = width; scene->m_opts.height = height; }
            if (spp >= 0) { scene->m_opts.spp = spp; scene->m_opts.sppe = sppe; scene->m_opts.sppse = sppse; }
            scene->m_opts.log_level = 0;
        })) return nullptr;
    return scene;
}
void ref_scene_free(void *s) { delete (Scene *)s; }
int ref_scene_configure(void *s) { return guard([&] { ((Scene *)s)->configure(); }); }
void ref_scene_options(void *s, int *out5) { const RenderOption &o = ((Scene *)s)->m_opts; out5[0] = o.width; out5[1] = o.height; out5[2] = o.spp; out5[3] = o.sppe; out5[4] = o.sppse; }
int ref_scene_num_meshes(void *s) { return ((Scene *)s)->m_num_meshes; }

// kind 0: DirectIntegrator(bsdf_samples, light_samples) with m_hide_emitters; kind 1: FieldExtractionIntegrator(field)
void *ref_integrator_new(int kind, int bsdf_samples, int light_samples, int hide_emitters, const char *field) {
    Integrator *I = nullptr;
    if (guard([&] {
            if (kind == 0) { auto *d = new DirectIntegrator(bsdf_samples, light_samples); d->m_hide_emitters = hide_emitters != 0; I = d; }
            else I = new FieldExtractionIntegrator(field);
        })) return nullptr;
    return I;
}
void ref_integrator_free(void *I) { delete (Integrator *)I; }

static void put_image(const float *src, size_t n_src, float *out, int npix, int c) { for (int i = 0; i < npix; ++i) out[3 * i + c] = n_src == 0 ? 0.f : src[n_src == 1 ? 0 : i]; }
// Integrator::renderC -> out[npix][3]
int ref_render_c(void *s, void *I, int sensor, float *out) {
    return guard([&] {
        Scene &scene = *(Scene *)s;
        SpectrumC img = ((Integrator *)I)->renderC(scene, sensor);
        for (int c = 0; c < 3; ++c) put_image(img[c].data(), img[c].size(), out, scene.m_opts.width * scene.m_opts.height, c);
    });
}
// Integrator::renderD -> out[npix][3] and the forward-mode tangent image out_t[npix][3] for the tangents currently set on the scene
int ref_render_d(void *s, void *I, int sensor, float *out, float *out_t) {
    return guard([&] {
        Scene &scene = *(Scene *)s;
        SpectrumD img = ((Integrator *)I)->renderD(scene, sensor);
        const int npix = scene.m_opts.width * scene.m_opts.height;
        for (int c = 0; c < 3; ++c) { put_image(img[c].v.data(), img[c].v.size(), out, npix, c); put_image(img[c].g.data(), img[c].g.size(), out_t, npix, c); }
    });
}
// debugging aid: the per-lane radiance Integrator::__render scatters (integrator.cpp:72-87 up to, not including, the scatter_add), for lanes
// of a freshly configured scene; consumes sampler 0 exactly like one render
int ref_debug_li(void *s, void *Iv, int sensor, int ad, int skip_dims, float *out) {   // skip_dims: sampler dimensions dropped between the pixel jitter and Li
    return guard([&] {
        Scene &scene = *(Scene *)s;
        const RenderOption &opts = scene.m_opts;
        const Integrator &I = *(Integrator *)Iv;
        const int64_t n = (int64_t)opts.width * opts.height * opts.spp;
        auto run = [&](auto tag) {
            constexpr bool AD = decltype(tag)::value;
            Int<AD> idx = arange<Int<AD>>(n);
            if (opts.spp > 1) idx /= opts.spp;
            Vector2f<AD> base = gather<Vector2f<AD>>(meshgrid(arange<Float<AD>>(opts.width), arange<Float<AD>>(opts.height)), idx);
            Vector2f<AD> samples = (base + scene.m_samplers[0].next_2d<AD>()) / ScalarVector2f(opts.width, opts.height);
            Ray<AD> ray = scene.m_sensors[sensor]->sample_primary_ray(samples);
            for (int k = 0; k < skip_dims; ++k) scene.m_samplers[0].next_1d<false>();
            Spectrum<AD> value = I.Li(scene, scene.m_samplers[0], ray);
            for (int c = 0; c < 3; ++c) for (int64_t i = 0; i < n; ++i) out[3 * i + c] = lval(detach(value[c]), i);
        };
        if (ad) run(std::true_type()); else run(std::false_type());
    });
}
int ref_preprocess_secondary_edges(void *s, void *I, int sensor, const int *reso4, int nrounds) {
    return guard([&] { ((Integrator *)I)->preprocess_secondary_edges(*(Scene *)s, sensor, ScalarVector4i(reso4[0], reso4[1], reso4[2], reso4[3]), nrounds); });
}

// ---- parameters: tangents (what ek.set_requires_gradient + ek.forward seed in the reference's Python flow) and values --------------------------
// which: 0 reflectance | 1 alpha_u | 2 alpha_v | 3 eta | 4 k | 5 specular_reflectance; t has one entry (x3) per texel, or null to clear
int ref_set_bsdf_tangent(void *s, int bsdf, int which, const float *t) {
    return guard([&] {
        BSDF *b = ((Scene *)s)->m_bsdfs.at(bsdf);
        if (auto *d = dynamic_cast<Diffuse *>(b)) { PSDR_ASSERT(which == 0); set_tangent(d->m_reflectance.m_data, t, slices(d->m_reflectance.m_data)); return; }
        auto *r = dynamic_cast<RoughConductor *>(b);
        PSDR_ASSERT(r != nullptr);
        switch (which) {
            case 1: set_tangent(r->m_alpha_u.m_data, t, slices(r->m_alpha_u.m_data), 1, 0); break;
            case 2: set_tangent(r->m_alpha_v.m_data, t, slices(r->m_alpha_v.m_data), 1, 0); break;
            case 3: set_tangent(r->m_eta.m_data, t, slices(r->m_eta.m_data)); break;
            case 4: set_tangent(r->m_k.m_data, t, slices(r->m_k.m_data)); break;
            case 5: set_tangent(r->m_specular_reflectance.m_data, t, slices(r->m_specular_reflectance.m_data)); break;
            default: PSDR_ASSERT(false);
        }
    });
}
int ref_set_bsdf_texture(void *s, int bsdf, int which, const float *data, int width, int height) {
    return guard([&] {
        BSDF *b = ((Scene *)s)->m_bsdfs.at(bsdf);
        const size_t n = (size_t)width * height;
        auto fill3 = [&](Bitmap3fD &bm) { std::vector<float> ch(n); Vector3fD v; for (int c = 0; c < 3; ++c) { for (size_t i = 0; i < n; ++i) ch[i] = data[3 * i + c]; v[c] = FloatD::copy(ch.data(), n); } bm.m_data = v; bm.m_resolution = ScalarVector2i(width, height); };
        auto fill1 = [&](Bitmap1fD &bm) { bm.m_data = FloatD::copy(data, n); bm.m_resolution = ScalarVector2i(width, height); };
        if (auto *d = dynamic_cast<Diffuse *>(b)) { PSDR_ASSERT(which == 0); fill3(d->m_reflectance); return; }
        auto *r = dynamic_cast<RoughConductor *>(b);
        PSDR_ASSERT(r != nullptr);
        switch (which) {
            case 1: fill1(r->m_alpha_u); break;
            case 2: fill1(r->m_alpha_v); break;
            case 3: fill3(r->m_eta); break;
            case 4: fill3(r->m_k); break;
            case 5: fill3(r->m_specular_reflectance); break;
            default: PSDR_ASSERT(false);
        }
    });
}
int ref_mesh_num_vertices(void *s, int mesh) { return ((Scene *)s)->m_meshes.at(mesh)->m_num_vertices; }
int ref_set_mesh_vertex_tangent(void *s, int mesh, const float *t) {   // d(m_vertex_positions_raw)[nv][3]; needs configure() afterwards
    return guard([&] { Mesh *m = ((Scene *)s)->m_meshes.at(mesh); set_tangent(m->m_vertex_positions_raw, t, m->m_num_vertices); m->m_ready = false; });
}
int ref_set_mesh_uv_tangent(void *s, int mesh, const float *t) {
    return guard([&] { Mesh *m = ((Scene *)s)->m_meshes.at(mesh); set_tangent(m->m_vertex_uv, t, slices(m->m_vertex_uv)); m->m_ready = false; });
}
int ref_set_mesh_transform(void *s, int mesh, const float *mat16, int left) {   // Mesh::set_transform
    return guard([&] { Matrix4fD M; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M(i, j) = FloatD(mat16[4 * i + j]); ((Scene *)s)->m_meshes.at(mesh)->set_transform(M, left != 0); });
}
int ref_set_mesh_transform_tangent(void *s, int mesh, const float *t16, int left) {
    return guard([&] { Mesh *m = ((Scene *)s)->m_meshes.at(mesh); set_tangent(left ? m->m_to_world_left : m->m_to_world_right, t16); m->m_ready = false; });
}
int ref_set_sensor_transform_tangent(void *s, int sensor, const float *t16) {
    return guard([&] { set_tangent(((Scene *)s)->m_sensors.at(sensor)->m_to_world, t16); });
}
int ref_set_envmap_tangent(void *s, const float *radiance_t, float scale_t) {
    return guard([&] {
        EnvironmentMap *e = ((Scene *)s)->m_emitter_env;
        PSDR_ASSERT(e != nullptr);
        set_tangent(e->m_radiance.m_data, radiance_t, slices(e->m_radiance.m_data));
        if (scale_t != 0.f) set_tangent(e->m_scale, &scale_t, 1, 1, 0); else set_tangent(e->m_scale, nullptr, 1, 1, 0);
    });
}
int ref_set_envmap_transform_tangent(void *s, const float *t16) {   // d(m_to_world_left), EnvironmentMap::set_transform's argument
    return guard([&] { EnvironmentMap *e = ((Scene *)s)->m_emitter_env; PSDR_ASSERT(e != nullptr); set_tangent(e->m_to_world_left, t16); e->m_ready = false; });
}

// ---- tables, in the oracle's introspection layouts (oracle/orc_capi.cpp) ----------------------------------------------------------------------
int ref_num_triangles(void *s) { return (int)slices(((Scene *)s)->m_triangle_info); }
int ref_get_triangle_info(void *s, float *out) {   // [n][22]: p0 e1 e2 n0 n1 n2 face_normal face_area
    const TriangleInfoD &t = ((Scene *)s)->m_triangle_info;
    const Vector3fD *v[7] = {&t.p0, &t.e1, &t.e2, &t.n0, &t.n1, &t.n2, &t.face_normal};
    const size_t n = slices(t);
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 7; ++k) for (int c = 0; c < 3; ++c) out[22 * i + 3 * k + c] = (*v[k])[c][i];
        out[22 * i + 21] = t.face_area[i];
    }
    return 0;
}
int ref_num_sec_edges(void *s) { return ((Scene *)s)->m_opts.sppse > 0 ? (int)slices(((Scene *)s)->m_sec_edge_info) : 0; }
int ref_get_sec_edges(void *s, float *out) {   // [n][16]: p0 e1 n0 n1 p2 is_boundary
    const SecondaryEdgeInfo &e = ((Scene *)s)->m_sec_edge_info;
    const Vector3fD *v[5] = {&e.p0, &e.e1, &e.n0, &e.n1, &e.p2};
    const size_t n = slices(e);
    for (size_t i = 0; i < n; ++i) {
        for (int k = 0; k < 5; ++k) for (int c = 0; c < 3; ++c) out[16 * i + 3 * k + c] = (*v[k])[c][i];
        out[16 * i + 15] = e.is_boundary[i] ? 1.f : 0.f;
    }
    return 0;
}
int ref_num_primary_edges(void *s, int sensor) { const Sensor *c = ((Scene *)s)->m_sensors.at(sensor); return c->m_enable_edges ? (int)slices(c->m_edge_info) : 0; }
int ref_get_primary_edges(void *s, int sensor, float *out) {   // [n][7]: p0.xy p1.xy edge_normal.xy edge_length
    const PrimaryEdgeInfo &e = ((Scene *)s)->m_sensors.at(sensor)->m_edge_info;
    const size_t n = slices(e);
    for (size_t i = 0; i < n; ++i) {
        out[7 * i + 0] = e.p0.x()[i]; out[7 * i + 1] = e.p0.y()[i]; out[7 * i + 2] = e.p1.x()[i]; out[7 * i + 3] = e.p1.y()[i];
        out[7 * i + 4] = e.edge_normal.x()[i]; out[7 * i + 5] = e.edge_normal.y()[i]; out[7 * i + 6] = e.edge_length[i];
    }
    return 0;
}
int ref_get_sensor(void *s, int sensor, float *out) {   // sample_to_camera(16) world_to_sample(16) to_world(16) camera_pos(3) camera_dir(3) inv_area(1)
    return guard([&] {
        auto *c = dynamic_cast<PerspectiveCamera *>(((Scene *)s)->m_sensors.at(sensor));
        PSDR_ASSERT(c != nullptr);
        const Matrix4fD *m[3] = {&c->m_sample_to_camera, &c->m_world_to_sample, &c->m_to_world};
        for (int k = 0; k < 3; ++k) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[16 * k + 4 * i + j] = (*m[k])(i, j)[0];
        for (int k = 0; k < 3; ++k) { out[48 + k] = c->m_camera_pos[k][0]; out[51 + k] = c->m_camera_dir[k][0]; }
        out[54] = c->m_inv_area[0];
    });
}
int ref_mesh_num_edges(void *s, int mesh) { return (int)slices(((Scene *)s)->m_meshes.at(mesh)->m_edge_indices); }
int ref_mesh_get_edges(void *s, int mesh, int *out) {   // [n][5]: Mesh::m_edge_indices (mesh.cpp:143-203)
    const auto &e = ((Scene *)s)->m_meshes.at(mesh)->m_edge_indices;
    const size_t n = slices(e);
    for (size_t i = 0; i < n; ++i) for (int k = 0; k < 5; ++k) out[5 * i + k] = e[k][i];
    return 0;
}
// Scene::sample_boundary_segment_direct (scene.cpp:456-492) on n samples -> [n][17]: p0 edge edge2 p2 n pdf is_valid
int ref_sample_boundary_segment_direct(void *s, int64_t n, const float *sample3, float *out) {
    return guard([&] {
        Scene &scene = *(Scene *)s;
        std::vector<float> buf(n);
        Vector3fC S3;
        for (int c = 0; c < 3; ++c) { for (int64_t i = 0; i < n; ++i) buf[i] = sample3[3 * i + c]; S3[c] = FloatC::copy(buf.data(), n); }
        BoundarySegSampleDirect b = scene.sample_boundary_segment_direct(S3, MaskC(true));
        const Vector3fC p0 = detach(b.p0);
        const Vector3fC *v[5] = {&p0, &b.edge, &b.edge2, &b.p2, &b.n};
        for (int64_t i = 0; i < n; ++i) {
            for (int k = 0; k < 5; ++k) for (int c = 0; c < 3; ++c) { const FloatC &a = (*v[k])[c]; out[17 * i + 3 * k + c] = a[slices(a) == 1 ? 0 : i]; }
            out[17 * i + 15] = b.pdf[slices(b.pdf) == 1 ? 0 : i];
            out[17 * i + 16] = b.is_valid[slices(b.is_valid) == 1 ? 0 : i] ? 1.f : 0.f;
        }
    });
}
// Scene::ray_intersect<false> on n rays -> global triangle id, shape index into m_meshes, OptiX-style barycentrics and its.t
int ref_trace(void *s, int64_t n, const float *o, const float *d, int *tri, int *shape, float *u, float *v, float *t) {
    return guard([&] {
        Scene &scene = *(Scene *)s;
        std::vector<float> buf(n);
        Vector3fC O, D;
        for (int c = 0; c < 3; ++c) { for (int64_t i = 0; i < n; ++i) buf[i] = o[3 * i + c]; O[c] = FloatC::copy(buf.data(), n); for (int64_t i = 0; i < n; ++i) buf[i] = d[3 * i + c]; D[c] = FloatC::copy(buf.data(), n); }
        IntersectionC its = scene.ray_intersect<false>(RayC(O, D), MaskC(true));
        const Intersection_OptiX &h = scene.m_optix->m_its;
        for (int64_t i = 0; i < n; ++i) {
            tri[i] = h.triangle_id[i]; u[i] = h.uv.x()[i]; v[i] = h.uv.y()[i]; t[i] = its.t[i];
            shape[i] = -1;
            for (int k = 0; k < scene.m_num_meshes; ++k) if (its.shape[i] == scene.m_meshes[k]) shape[i] = k;
        }
    });
}

// ---- the reference's utility classes that its Python module exposes (src/psdr.cpp:140-164, 100-118, 242-265), for pinning the host-side
//      mirrors in psdr_cuda_b200/compat/psdr_cuda/_surface.py ---------------------------------------------------------------------------------
// DiscreteDistribution::init + sample (pmf.cpp:7-27) and sample_reuse (pmf.cpp:30-50)
int ref_discrete_sample(const float *pmf, int n, const float *u, int m, int reuse, int *idx, float *pdf, float *u_out) {
    return guard([&] {
        DiscreteDistribution d;
        d.init(FloatC::copy(pmf, n));
        FloatC samples = FloatC::copy(u, m);
        auto r = reuse ? d.sample_reuse<false>(samples) : d.sample(samples);
        for (int i = 0; i < m; ++i) { idx[i] = lval(r.first, i); pdf[i] = lval(r.second, i); u_out[i] = lval(samples, i); }
    });
}
// HyperCubeDistribution<ndim>: set_resolution + set_mass + sample_reuse + pdf (cube_distrb.cpp:8-62); samples[m][ndim] in, warped samples out
int ref_hypercube(int ndim, const int *reso, const float *mass, const float *samples, int m, float *warped, float *pdf_sample, float *pdf_eval, int *cells) {
    return guard([&] {
        auto run = [&](auto &hc, auto reso_v, auto smp) {
            hc.set_resolution(reso_v);
            hc.set_mass(FloatC::copy(mass, hc.m_num_cells));
            std::vector<float> buf(m);
            for (int c = 0; c < ndim; ++c) { for (int i = 0; i < m; ++i) buf[i] = samples[i * ndim + c]; smp[c] = FloatC::copy(buf.data(), m); }
            FloatC pe = hc.pdf(smp);
            FloatC ps = hc.sample_reuse(smp);
            for (int i = 0; i < m; ++i) { pdf_sample[i] = lval(ps, i); pdf_eval[i] = lval(pe, i); for (int c = 0; c < ndim; ++c) warped[i * ndim + c] = lval(smp[c], i); }
            for (int k = 0; k < hc.m_num_cells; ++k) for (int c = 0; c < ndim; ++c) cells[k * ndim + c] = lval(hc.m_cells[c], k);
        };
        if (ndim == 2) { HyperCubeDistribution2f hc; run(hc, ScalarVector2i(reso[0], reso[1]), Vector2fC()); }
        else { HyperCubeDistribution3f hc; run(hc, ScalarVector3i(reso[0], reso[1], reso[2]), Vector3fC()); }
    });
}
// Bitmap<channels>::eval<false>(uv, flip_v) (bitmap.cpp:56-96) on a width x height texture given as [h][w][channels]
int ref_bitmap_eval(int channels, int width, int height, const float *data, const float *uv, int m, int flip_v, float *out) {
    return guard([&] {
        const size_t n = (size_t)width * height;
        std::vector<float> buf(std::max<size_t>(n, m));
        Vector2fC q;
        for (int c = 0; c < 2; ++c) { for (int i = 0; i < m; ++i) buf[i] = uv[2 * i + c]; q[c] = FloatC::copy(buf.data(), m); }
        if (channels == 1) {
            Bitmap1fD bm(width, height, FloatD::copy(data, n));
            FloatC r = bm.eval<false>(q, flip_v != 0);
            for (int i = 0; i < m; ++i) out[i] = lval(r, i);
        } else {
            Vector3fD v;
            for (int c = 0; c < 3; ++c) { for (size_t i = 0; i < n; ++i) buf[i] = data[3 * i + c]; v[c] = FloatD::copy(buf.data(), n); }
            Bitmap3fD bm(width, height, v);
            Vector3fC r = bm.eval<false>(q, flip_v != 0);
            for (int i = 0; i < m; ++i) for (int c = 0; c < 3; ++c) out[3 * i + c] = lval(r[c], i);
        }
    });
}
// Mesh::sample_position(sample2) of a configured scene's mesh (mesh.cpp:277-303; needs an emitter mesh: the others drop their triangle
// table after Scene::configure, scene.cpp:245-262) -> p[m][3], n[m][3], pdf[m]
int ref_mesh_sample_position(void *s, int mesh, const float *sample2, int m, float *p, float *n, float *pdf) {
    return guard([&] {
        const Mesh *M = ((Scene *)s)->m_meshes.at(mesh);
        std::vector<float> buf(m);
        Vector2fC q;
        for (int c = 0; c < 2; ++c) { for (int i = 0; i < m; ++i) buf[i] = sample2[2 * i + c]; q[c] = FloatC::copy(buf.data(), m); }
        PositionSampleC ps = M->sample_position(q, MaskC(true));
        for (int i = 0; i < m; ++i) { for (int c = 0; c < 3; ++c) { p[3 * i + c] = lval(ps.p[c], i); n[3 * i + c] = lval(ps.n[c], i); } pdf[i] = lval(ps.pdf, i); }
    });
}
// Mesh::m_vertex_normals_raw after configure (mesh.cpp:19-51 on the object-space positions)
int ref_mesh_vertex_normals(void *s, int mesh, float *out) {
    return guard([&] {
        const Mesh *M = ((Scene *)s)->m_meshes.at(mesh);
        for (int i = 0; i < M->m_num_vertices; ++i) for (int c = 0; c < 3; ++c) out[3 * i + c] = M->m_vertex_normals_raw[c][i];
    });
}
// keys of Scene::m_param_map (scene_loader.cpp:184-240, 343-352), '\n'-separated
const char *ref_param_map_keys(void *s) {
    static thread_local std::string keys;
    keys.clear();
    std::vector<std::string> v;
    for (const auto &kv : ((Scene *)s)->m_param_map) v.push_back(kv.first + "=" + kv.second.type_name() + "\t" + kv.second.to_string());
    std::sort(v.begin(), v.end());
    for (const auto &k : v) keys += k + "\n";
    return keys.c_str();
}

const char *ref_scene_describe(void *sp) {
    static thread_local std::string out;
    const Scene &s = *(Scene *)sp;
    std::ostringstream os;
    char b[64];
    os << "opts " << s.m_opts.width << " " << s.m_opts.height << " " << s.m_opts.spp << " " << s.m_opts.sppe << " " << s.m_opts.sppse << "\n";
    for (const Sensor *c : s.m_sensors) {
        auto *p = dynamic_cast<const PerspectiveCamera *>(c);
        snprintf(b, sizeof b, "sensor %.9g %.9g %.9g", (double)p->m_fov_x, (double)p->m_near_clip, (double)p->m_far_clip); os << b; put_mat(os, c->m_to_world); os << "\n";
    }
    for (const BSDF *bs : s.m_bsdfs) {
        os << "bsdf " << bs->type_name() << " id=" << bs->m_id;
        if (auto *d = dynamic_cast<const Diffuse *>(bs)) put_tex(os, "reflectance", d->m_reflectance.m_data, d->m_reflectance.m_resolution, 3);
        if (auto *r = dynamic_cast<const RoughConductor *>(bs)) {
            put_tex(os, "alpha_u", r->m_alpha_u.m_data, r->m_alpha_u.m_resolution, 1); put_tex(os, "alpha_v", r->m_alpha_v.m_data, r->m_alpha_v.m_resolution, 1);
            put_tex(os, "eta", r->m_eta.m_data, r->m_eta.m_resolution, 3); put_tex(os, "k", r->m_k.m_data, r->m_k.m_resolution, 3);
            put_tex(os, "specular_reflectance", r->m_specular_reflectance.m_data, r->m_specular_reflectance.m_resolution, 3);
        }
        os << "\n";
    }
    if (s.m_emitter_env) {
        snprintf(b, sizeof b, "envmap %.9g %dx%d", (double)s.m_emitter_env->m_scale[0], s.m_emitter_env->m_radiance.m_resolution.x(), s.m_emitter_env->m_radiance.m_resolution.y());
        os << b; put_mat(os, s.m_emitter_env->m_to_world_raw); os << "\n";
    }
    for (const Mesh *m : s.m_meshes) {
        os << "mesh id=" << m->m_id << " nv=" << m->m_num_vertices << " nf=" << m->m_num_faces << " uv=" << (m->m_has_uv ? (int)slices(m->m_vertex_uv) : 0)
           << " face_normals=" << m->m_use_face_normals << " edges=" << m->m_enable_edges << " bsdf=" << (m->m_bsdf ? m->m_bsdf->m_id : std::string("-"));
        if (auto *a = dynamic_cast<const AreaLight *>(m->m_emitter)) { snprintf(b, sizeof b, " radiance %.9g %.9g %.9g", (double)a->m_radiance[0][0], (double)a->m_radiance[1][0], (double)a->m_radiance[2][0]); os << b; }
        os << " to_world"; put_mat(os, m->m_to_world_raw); os << "\n";
    }
    out = os.str();
    return out.c_str();
}
// SceneLoader::load_from_string on an XML text (relative file names resolve against cwd)
void *ref_scene_load_string(const char *xml, const char *cwd) {
    Scene *scene = nullptr;
    if (guard([&] {
            char old[4096];
            if (!getcwd(old, sizeof old)) throw Exception("getcwd failed");
            if (cwd && chdir(cwd) != 0) throw Exception(std::string("cannot chdir to ") + cwd);
            try { scene = new Scene(); scene->load_string(xml, false); } catch (...) { (void)!chdir(old); delete scene; scene = nullptr; throw; }
            (void)!chdir(old);
            scene->m_opts.log_level = 0;
        })) return nullptr;
    return scene;
}

// Mesh::dump (mesh.cpp:318-392) of a loaded scene's mesh
int ref_mesh_dump(void *s, int mesh, const char *path) { return guard([&] { ((Scene *)s)->m_meshes.at(mesh)->dump(path); }); }

// PerspectiveCamera::sample_primary_ray(samples) (perspective.cpp:120-136, the C flavour) -> o[n][3], d[n][3]
int ref_sample_primary_ray(void *s, int sensor, const float *samples2, int n, float *o, float *d) {
    return guard([&] {
        std::vector<float> buf(n);
        Vector2fC q;
        for (int c = 0; c < 2; ++c) { for (int i = 0; i < n; ++i) buf[i] = samples2[2 * i + c]; q[c] = FloatC::copy(buf.data(), n); }
        RayC r = ((Scene *)s)->m_sensors.at(sensor)->sample_primary_ray(q);
        for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) { o[3 * i + c] = lval(r.o[c], i); d[3 * i + c] = lval(r.d[c], i); }
    });
}
// the configured environment map: scale, resolution, m_from_world, and its radiance texels [h][w][3]
int ref_get_envmap(void *s, float *scale_w_h_from_world19, float *radiance) {
    return guard([&] {
        const EnvironmentMap *e = ((Scene *)s)->m_emitter_env;
        PSDR_ASSERT(e != nullptr && e->m_ready);
        scale_w_h_from_world19[0] = e->m_scale[0]; scale_w_h_from_world19[1] = (float)e->m_radiance.m_resolution.x(); scale_w_h_from_world19[2] = (float)e->m_radiance.m_resolution.y();
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) scale_w_h_from_world19[3 + 4 * i + j] = e->m_from_world(i, j)[0];
        if (radiance) { const size_t n = slices(e->m_radiance.m_data); for (size_t i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) radiance[3 * i + c] = e->m_radiance.m_data[c][i]; }
    });
}
// EnvironmentMap::eval_direction<false>(wi) (envmap.cpp:42-58)
int ref_env_eval_direction(void *s, const float *dirs, int n, float *out) {
    return guard([&] {
        const EnvironmentMap *e = ((Scene *)s)->m_emitter_env;
        PSDR_ASSERT(e != nullptr);
        std::vector<float> buf(n);
        Vector3fC w;
        for (int c = 0; c < 3; ++c) { for (int i = 0; i < n; ++i) buf[i] = dirs[3 * i + c]; w[c] = FloatC::copy(buf.data(), n); }
        SpectrumC v = e->eval_direction<false>(w, MaskC(true));
        for (int i = 0; i < n; ++i) for (int c = 0; c < 3; ++c) out[3 * i + c] = lval(v[c], i);
    });
}

// RoughConductor eval / pdf / sample in the D flavour with a forward-mode tangent on its eleven (constant-texture) parameters
// [alpha_u, alpha_v, eta(3), k(3), specular_reflectance(3)] (roughconductor.cpp:40-93): values and tangents. what: 0 eval (3 + 3 floats),
// 1 pdf (1 + 1), 2 the pdf of sample(wi, sample3) (1 + 1; 0 when the sample is invalid)
int ref_rc_d(int what, const float *prm, const float *prm_t, const float *wi, const float *wo_or_sample, float *out, float *out_t) {
    return guard([&] {
        auto f1 = [&](int k) { FloatD x(prm[k]); x.g = FloatC(prm_t[k]); return Bitmap1fD(1, 1, x); };
        auto f3 = [&](int k) { Vector3fD x; for (int c = 0; c < 3; ++c) { x[c] = FloatD(prm[k + c]); x[c].g = FloatC(prm_t[k + c]); } return Bitmap3fD(1, 1, x); };
        RoughConductor rc(f1(0), f1(1), f3(2), f3(5), f3(8));
        IntersectionD its;
        its.wi = Vector3fD(wi[0], wi[1], wi[2]);
        its.uv = Vector2fD(0.f, 0.f);
        const Vector3fD w(wo_or_sample[0], wo_or_sample[1], wo_or_sample[2]);
        auto put = [&](const FloatD &x, int k) { out[k] = lval(x, 0); out_t[k] = ltan(x, 0); };
        if (what == 0) { SpectrumD v = rc.eval(its, w, MaskD(true)); for (int c = 0; c < 3; ++c) put(v[c], c); }
        else if (what == 1) put(rc.pdf(its, w, MaskD(true)), 0);
        else { BSDFSampleD bs = rc.sample(its, w, MaskD(true)); put(bs.pdf & bs.is_valid, 0); }
    });
}

}  // extern "C"
