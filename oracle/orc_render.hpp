// ORACLE — TEST INFRASTRUCTURE ONLY (parity pinned against the reference's own source run on the CPU; see orc_math.hpp header and DESIGN.md §2).
//
// CPU restatement of psdr-cuda's sensor, scene, BSDF, emitter and integrator layers. R = float is the
// reference's "C" flavour, R = Dual (one forward-mode tangent) its "D" flavour; detach() drops the tangent.
// Every function cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include "orc_scene.hpp"

namespace orc {

struct RenderOption { int width = 0, height = 0, spp = 0, sppe = 0, sppse = 0; };   // types.h:171-182

// ---- tiny helpers ----------------------------------------------------------------------------------
template <class R> inline V3<R> lift(const V3f &a) { return V3<R>(R(a.x), R(a.y), R(a.z)); }
template <class R> inline V3<R> cast3(const V3<Dual> &a) {
    if constexpr (std::is_same_v<R, Dual>) return a; else return detach(a);
}
template <class R> inline V2<R> cast2(const V2<Dual> &a) {
    if constexpr (std::is_same_v<R, Dual>) return a; else return detach(a);
}
template <class R> inline R cast1(const Dual &a) {
    if constexpr (std::is_same_v<R, Dual>) return a; else return a.v;
}
template <class R> inline M4<R> castM(const M4<Dual> &a) {
    if constexpr (std::is_same_v<R, Dual>) return a; else return detach(a);
}
template <class R> inline void zero_nonfinite(V3<R> &v) {   // integrator.cpp:87 (isfinite looks at the primal)
    for (int k = 0; k < 3; ++k) if (!std::isfinite(val(v[k]))) v[k] = R(0.f);
}
inline int sign_eps(float x, float eps) { return x > eps ? 1 : (x < -eps ? -1 : 0); }   // utils.h:32-38

// ---- transforms (transform.h:14-79; Enoki translate/scale/rotate) -----------------
inline M4f m_translate(float x, float y, float z) { M4f m; m.m[0][3] = x; m.m[1][3] = y; m.m[2][3] = z; return m; }
inline M4f m_scale(float x, float y, float z) { M4f m; m.m[0][0] = x; m.m[1][1] = y; m.m[2][2] = z; return m; }
inline M4f m_rotate(V3f a, float angle_rad) {   // right-handed rotation about an (assumed unit) axis
    float s = std::sin(angle_rad), c = std::cos(angle_rad), cm = 1.f - c;
    M4f m;
    m.m[0][0] = a.x * a.x * cm + c;       m.m[0][1] = a.x * a.y * cm - a.z * s; m.m[0][2] = a.x * a.z * cm + a.y * s;
    m.m[1][0] = a.y * a.x * cm + a.z * s; m.m[1][1] = a.y * a.y * cm + c;       m.m[1][2] = a.y * a.z * cm - a.x * s;
    m.m[2][0] = a.z * a.x * cm - a.y * s; m.m[2][1] = a.z * a.y * cm + a.x * s; m.m[2][2] = a.z * a.z * cm + c;
    return m;
}
inline M4f m_perspective(float fov_deg, float near_, float far_) {   // transform.h:45-59
    float recip = 1.f / (far_ - near_);
    float t = std::tan((fov_deg * .5f) * (kPi / 180.f)), cot = 1.f / t;   // enoki::tan(deg_to_rad(fov * .5f)), deg_to_rad(a) = a * (Pi / 180)
    M4f m;
    m.m[0][0] = cot; m.m[1][1] = cot; m.m[2][2] = far_ * recip; m.m[3][3] = 0.f;
    m.m[2][3] = -near_ * far_ * recip; m.m[3][2] = 1.f;
    return m;
}
inline M4f m_look_at(V3f origin, V3f target, V3f up) {   // transform.h:67-79
    V3f dir = normalize(target - origin), left = normalize(cross(up, dir)), new_up = cross(dir, left);
    M4f m;
    for (int i = 0; i < 3; ++i) { m.m[i][0] = left[i]; m.m[i][1] = new_up[i]; m.m[i][2] = dir[i]; m.m[i][3] = origin[i]; }
    return m;
}

// =================================================================================================
// Emitters: src/emitter/area.cpp, src/emitter/envmap.cpp
// =================================================================================================
struct Emitter {
    int type = EMITTER_AREA;
    V3f radiance;           // area.h:26
    int mesh = -1;          // area.h:27
    float sampling_weight = 0.f;
    // envmap (envmap.h)
    Bitmap env_radiance;    // .tang: forward-mode tangent of the radiance texels (EnvironmentMap.radiance.data is an AD leaf, psdr.cpp:236)
    float env_scale = 1.f, env_scale_t = 0.f;   // scale and its tangent (psdr.cpp:237)
    M4f env_to_world_raw, env_left, env_left_t;   // env_left_t: forward-mode tangent of the matrix set by EnvironmentMap.set_transform (psdr.cpp:238)
    bool env_has_t = false;
    M4f env_to_world, env_from_world;
    M4<Dual> env_from_world_d;
    V3f lower, upper;
    HyperCube<2> env_distrb;
    std::vector<float> env_cell_lum;
};

// =================================================================================================
// Sensor: src/sensor/sensor.cpp, src/sensor/perspective.cpp
// =================================================================================================
template <class R> struct PrimaryEdgeSample { R x_dot_n = R(0.f); int idx = -1; Ray<float> ray_n, ray_p; float pdf = 0.f; };
struct SensorDirectSample { V2f q; int pixel_idx = -1; float sensor_val = 0.f; bool valid = false; };

struct Sensor {
    float fov_x = 45.f, near_clip = 0.1f, far_clip = 1e4f;
    M4f to_world, to_world_t;
    bool has_t = false;
    // configured
    int W = 0, H = 0;
    float aspect = 1.f;
    M4<Dual> to_world_d, sample_to_camera, camera_to_sample, world_to_sample, sample_to_world;
    V3<Dual> camera_pos, camera_dir;
    Dual inv_area;
    std::vector<PrimEdge<Dual>> edges;
    DiscreteDistribution edge_distrb;
    bool enable_edges = false;

    // perspective.cpp:11-32
    void configure_matrices(int w, int h) {
        W = w; H = h;
        aspect = (float)w / (float)h;
        float det = to_world.m[0][0] * (to_world.m[1][1] * to_world.m[2][2] - to_world.m[1][2] * to_world.m[2][1]) -
                    to_world.m[0][1] * (to_world.m[1][0] * to_world.m[2][2] - to_world.m[1][2] * to_world.m[2][0]) +
                    to_world.m[0][2] * (to_world.m[1][0] * to_world.m[2][1] - to_world.m[1][1] * to_world.m[2][0]);
        if (!(std::fabs(det - 1.f) < kEpsilon)) throw std::runtime_error("Sensor transformation should not involve scaling!");
        M4f c2s = m_scale(-0.5f, -0.5f * aspect, 1.f) * m_translate(-1.f, -1.f / aspect, 0.f) * m_perspective(fov_x, near_clip, far_clip);
        camera_to_sample = M4<Dual>(c2s);
        sample_to_camera = M4<Dual>(inverse(c2s));
        to_world_d = M4<Dual>(to_world);
        if (has_t) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) to_world_d.m[i][j].d = to_world_t.m[i][j];
        world_to_sample = camera_to_sample * inverse(to_world_d);
        sample_to_world = to_world_d * sample_to_camera;
        camera_pos = transform_pos(to_world_d, V3<Dual>());
        camera_dir = transform_dir(to_world_d, V3<Dual>(Dual(0.f), Dual(0.f), Dual(1.f)));
        using D3 = V3<Dual>;
        D3 v00 = transform_pos(sample_to_camera, D3(Dual(0.f), Dual(0.f), Dual(0.f))),
           v10 = transform_pos(sample_to_camera, D3(Dual(1.f), Dual(0.f), Dual(0.f))),
           v11 = transform_pos(sample_to_camera, D3(Dual(1.f), Dual(1.f), Dual(0.f))),
           vc = transform_pos(sample_to_camera, D3(Dual(.5f), Dual(.5f), Dual(0.f)));
        inv_area = Dual(1.f) / (norm(v00 - v10) * norm(v11 - v10)) * squared_norm(vc);
    }

    // perspective.cpp:120-136
    template <class R> Ray<R> sample_primary_ray(const V2<R> &s) const {
        V3<R> d = normalize(transform_pos(castM<R>(sample_to_camera), V3<R>(s.x, s.y, R(0.f))));
        M4<R> tw = castM<R>(to_world_d);
        return Ray<R>(transform_pos(tw, V3<R>()), transform_dir(tw, d));
    }
    // perspective.cpp:139-155
    SensorDirectSample sample_direct(const V3f &p) const {
        SensorDirectSample r;
        V3f q3 = transform_pos(detach(world_to_sample), p);
        r.q = V2f(q3.x, q3.y);
        int ix = (int)std::floor(r.q.x * (float)W), iy = (int)std::floor(r.q.y * (float)H);
        r.valid = ix >= 0 && ix < W && iy >= 0 && iy < H;
        r.pixel_idx = r.valid ? iy * W + ix : -1;
        V3f dir = p - detach(camera_pos);
        float dist2 = squared_norm(dir);
        dir = dir / safe_sqrt(dist2);
        float cosTheta = dot(detach(camera_dir), dir);
        float rc = 1.f / cosTheta;
        r.sensor_val = (1.f / dist2) * (rc * rc * rc) * inv_area.v;
        return r;
    }
    // perspective.cpp:158-200
    PrimaryEdgeSample<Dual> sample_primary_edge(float sample1) const {
        PrimaryEdgeSample<Dual> r;
        auto [edge_idx, pdf] = edge_distrb.sample_reuse(sample1);
        const PrimEdge<Dual> &info = edges[edge_idx];
        r.pdf = pdf / info.edge_length;
        V2f en = info.edge_normal;
        V2<Dual> p_(fma_(info.p0.x, Dual(1.f - sample1), info.p1.x * sample1), fma_(info.p0.y, Dual(1.f - sample1), info.p1.y * sample1));
        V2f p = detach(p_);
        r.x_dot_n = dot(p_, V2<Dual>(Dual(en.x), Dual(en.y)));
        int ix = (int)std::floor(p.x * (float)W), iy = (int)std::floor(p.y * (float)H);
        bool valid = ix >= 0 && ix < W && iy >= 0 && iy < H;
        r.idx = valid ? iy * W + ix : -1;
        r.ray_p = sample_primary_ray<float>(V2f(p.x + kEdgeEpsilon * en.x, p.y + kEdgeEpsilon * en.y));
        r.ray_n = sample_primary_ray<float>(V2f(p.x - kEdgeEpsilon * en.x, p.y - kEdgeEpsilon * en.y));
        return r;
    }
};

struct BoundarySegSampleDirect {   // records.h:35-45
    V3<Dual> p0;
    V3f edge, edge2, p2, n;
    float pdf = 0.f;
    bool valid = false;
};

// =================================================================================================
// Scene: src/scene/scene.cpp
// =================================================================================================
struct Scene {
    RenderOption opts;
    std::vector<Mesh> meshes;
    std::vector<Bsdf> bsdfs;
    std::vector<Emitter> emitters;
    std::vector<Sensor> sensors;
    int emitter_env = -1;
    bool has_bound_mesh = false;
    // samplers (scene.cpp:65-79): per-lane PCG32 state, re-seeded only when the lane count changes
    std::vector<SamplerLane> samplers[3];
    // configured global tables (scene.cpp:205-244)
    std::vector<TriangleInfo<Dual>> tri;
    std::vector<std::array<V2f, 3>> tri_uv, tri_uv_t;   // per-triangle texture coordinates and their tangents
    std::vector<int> tri_mesh;          // shape id per global triangle
    std::vector<SecEdge<Dual>> sec_edges;
    DiscreteDistribution sec_edge_distrb, emitters_distrb;
    TriAccel accel;
    V3f lower, upper;
    bool ready = false;

    static void seed_sampler(std::vector<SamplerLane> &s, int64_t count) {
        if ((int64_t)s.size() == count) return;
        s.resize(count);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < count; ++i) s[i] = SamplerLane::make((uint64_t)i);
    }

    void configure();
    template <class R, bool PS> Intersection<R> ray_intersect(const Ray<R> &ray, bool active = true, TriangleInfo<Dual> *out_info = nullptr) const;

    // intersection.h:32-38 / area.cpp:20-29
    template <class R> V3<R> Le(const Intersection<R> &its, bool active) const;
    bool is_emitter(int shape) const { return shape >= 0 && meshes[shape].emitter >= 0; }
    template <class R> PositionSample<R> mesh_sample_position(const Mesh &m, V2<R> sample2) const;
    template <class R> PositionSample<R> sample_emitter_position(const V3<R> &ref_p, V2<R> sample2, bool active) const;
    template <class R> float emitter_position_pdf(const V3<R> &ref_p, const Intersection<R> &its, bool active) const;
    BoundarySegSampleDirect sample_boundary_segment_direct(V3f sample3, bool active = true) const;

    // BSDF vcalls (bsdf.h:27-34), dispatched on the hit mesh's BSDF
    template <class R> V3<R> bsdf_eval(const Intersection<R> &its, const V3<R> &wo, bool active) const;
    template <class R> BSDFSample<R> bsdf_sample(const Intersection<R> &its, const V3<R> &sample, bool active) const;
    template <class R> R bsdf_pdf(const Intersection<R> &its, const V3<R> &wo, bool active) const;
};

inline void configure_envmap(Emitter &e);

// perspective.cpp:39-111
inline void configure_primary_edges(Sensor &s, const Scene &scene) {
    s.edges.clear();
    s.enable_edges = false;
    if (scene.opts.sppe <= 0) return;
    using D3 = V3<Dual>;
    for (const Mesh &mesh : scene.meshes) {
        if (!mesh.enable_edges) continue;
        int ne = (int)mesh.edges.size() / 5, kept = 0;
        for (int e = 0; e < ne; ++e) {
            const int *ed = &mesh.edges[5 * e];
            bool valid = ed[3] >= 0;
            D3 e0 = normalize(s.camera_pos - mesh.tri[ed[2]].p0), n0 = mesh.tri[ed[2]].face_normal;
            D3 e1, n1;   // masked gathers return 0 on boundary edges; normalize(cam - 0) is still evaluated
            if (valid) { e1 = normalize(s.camera_pos - mesh.tri[ed[3]].p0); n1 = mesh.tri[ed[3]].face_normal; }
            else e1 = normalize(s.camera_pos);
            bool keep;
            if (mesh.face_normals) {
                bool skip = valid && ((val(dot(e0, n0)) < kEpsilon && val(dot(e1, n1)) < kEpsilon) || (val(dot(n0, n1)) > 1.f - kEpsilon));
                keep = !skip;
            } else {
                keep = !valid || ((val(dot(e0, n0)) > kEpsilon) != (val(dot(e1, n1)) > kEpsilon));
            }
            if (!keep) continue;
            ++kept;
            D3 q0 = transform_pos(s.world_to_sample, mesh.vworld[ed[0]]), q1 = transform_pos(s.world_to_sample, mesh.vworld[ed[1]]);
            PrimEdge<Dual> pe;
            pe.p0 = V2<Dual>(q0.x, q0.y); pe.p1 = V2<Dual>(q1.x, q1.y);
            V2f ev(q1.x.v - q0.x.v, q1.y.v - q0.y.v);
            float len = norm(ev);
            ev = V2f(ev.x / len, ev.y / len);
            pe.edge_normal = V2f(-ev.y, ev.x);
            pe.edge_length = len;
            s.edges.push_back(pe);
        }
        if (kept == 0) throw std::runtime_error("PSDR_ASSERT(slices(info) > 0)");   // perspective.cpp:67
    }
    if (!s.edges.empty()) {
        std::vector<float> len(s.edges.size());
        for (size_t i = 0; i < len.size(); ++i) len[i] = s.edges[i].edge_length;
        s.edge_distrb.init(len);
        s.enable_edges = true;
    }
}

inline void Scene::configure() {   // scene.cpp:56-278
    const int64_t npix = (int64_t)opts.width * opts.height;
    if (opts.spp > 0) seed_sampler(samplers[0], npix * opts.spp);
    if (opts.sppe > 0) seed_sampler(samplers[1], npix * opts.sppe);
    if (opts.sppse > 0) seed_sampler(samplers[2], npix * opts.sppse);
    if (meshes.empty()) throw std::runtime_error("Missing meshes!");
    if (has_bound_mesh) { meshes.pop_back(); has_bound_mesh = false; }   // the oracle rebuilds the bounding mesh every time
    for (int k = 0; k < 3; ++k) { lower[k] = std::numeric_limits<float>::max(); upper[k] = std::numeric_limits<float>::min(); }
    for (Mesh &m : meshes) {
        configure_mesh(m);
        for (int v = 0; v < m.nv; ++v)
            for (int k = 0; k < 3; ++k) { lower[k] = min_(lower[k], m.vworld[v][k].v); upper[k] = max_(upper[k], m.vworld[v][k].v); }
    }
    if (sensors.empty()) throw std::runtime_error("Missing sensor!");
    for (Sensor &s : sensors) {
        s.configure_matrices(opts.width, opts.height);
        configure_primary_edges(s, *this);
        for (int k = 0; k < 3; ++k) { lower[k] = min_(lower[k], s.camera_pos[k].v); upper[k] = max_(upper[k], s.camera_pos[k].v); }
    }
    if (emitter_env >= 0) {   // scene.cpp:135-180
        float margin = min_(min_((upper.x - lower.x) * 0.05f, (upper.y - lower.y) * 0.05f), (upper.z - lower.z) * 0.05f);
        for (int k = 0; k < 3; ++k) { lower[k] -= margin; upper[k] += margin; }
        Emitter &env = emitters[emitter_env];
        env.lower = lower; env.upper = upper;
        static const int face_data[3][12] = {{0, 0, 1, 1, 2, 2, 0, 0, 0, 0, 4, 4}, {1, 3, 5, 7, 3, 7, 5, 4, 2, 6, 7, 6}, {3, 2, 7, 3, 7, 6, 1, 5, 6, 4, 5, 7}};
        Mesh b;
        b.nv = 8; b.nf = 12; b.face_normals = true; b.enable_edges = false; b.bsdf = -1; b.emitter = emitter_env;
        b.vraw.resize(24); b.faces.resize(36);
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 3; ++j) b.vraw[3 * i + j] = (i & (1 << j)) ? upper[j] : lower[j];
        for (int f = 0; f < 12; ++f) for (int j = 0; j < 3; ++j) b.faces[3 * f + j] = face_data[j][f];
        configure_mesh(b);
        meshes.push_back(b);
        has_bound_mesh = true;
    }
    if (!emitters.empty()) {   // scene.cpp:183-196
        std::vector<float> w;
        for (Emitter &e : emitters) {
            if (e.type == EMITTER_AREA) e.sampling_weight = meshes[e.mesh].total_area * rgb2luminance(e.radiance);   // area.cpp:10-17
            else { configure_envmap(e); e.sampling_weight = 1.f; }
            w.push_back(e.sampling_weight);
        }
        emitters_distrb.init(w);
        float inv = 1.f / emitters_distrb.sum;
        for (Emitter &e : emitters) e.sampling_weight *= inv;
    }
    tri.clear(); tri_uv.clear(); tri_uv_t.clear(); tri_mesh.clear(); sec_edges.clear();
    for (size_t i = 0; i < meshes.size(); ++i) {
        Mesh &m = meshes[i];
        m.face_offset = (int)tri.size();
        for (int f = 0; f < m.nf; ++f) {
            tri.push_back(m.tri[f]);
            std::array<V2f, 3> uv;
            if (m.has_uv) for (int k = 0; k < 3; ++k) { int j = m.uv_faces[3 * f + k]; uv[k] = V2f(m.uvs[2 * j], m.uvs[2 * j + 1]); }
            tri_uv.push_back(uv);
            std::array<V2f, 3> uvt;
            if (m.has_uv && !m.uvs_t.empty()) for (int k = 0; k < 3; ++k) { int j = m.uv_faces[3 * f + k]; uvt[k] = V2f(m.uvs_t[2 * j], m.uvs_t[2 * j + 1]); }
            tri_uv_t.push_back(uvt);
            tri_mesh.push_back((int)i);
        }
        if (opts.sppse > 0 && m.enable_edges) for (auto &s : m.sec_edges) sec_edges.push_back(s);
    }
    if (opts.sppse > 0) {   // scene.cpp:219-235
        std::vector<float> len(sec_edges.size());
        for (size_t i = 0; i < len.size(); ++i) len[i] = norm(detach(sec_edges[i].e1));
        sec_edge_distrb.init(len);
    }
    std::vector<V3f> P0(tri.size()), E1(tri.size()), E2(tri.size());
    for (size_t i = 0; i < tri.size(); ++i) { P0[i] = detach(tri[i].p0); E1[i] = detach(tri[i].e1); E2[i] = detach(tri[i].e2); }
    accel.build(P0, E1, E2);
    ready = true;
}

// scene.cpp:289-384
template <class R, bool PS>
inline Intersection<R> Scene::ray_intersect(const Ray<R> &ray, bool active, TriangleInfo<Dual> *out_info) const {
    constexpr bool ad = std::is_same_v<R, Dual>;
    static_assert(ad || !PS);
    Intersection<R> its;
    Ray<float> rc(detach(ray.o), detach(ray.d));
    rc.tmax = ray.tmax;
    Hit hit = accel.closest(rc);   // OptiX traces every lane; `active` only masks the gathers
    if (!active || hit.tri < 0) {
        if (out_info) *out_info = TriangleInfo<Dual>();
        its.t = R(0.f);   // masked gathers give zero-filled records
        return its;
    }
    if (out_info) *out_info = tri[hit.tri];
    TriangleInfo<R> ti = cast_tri<R>(tri[hit.tri]);
    const auto &tuv = tri_uv[hit.tri];
    const int shape = tri_mesh[hit.tri];
    const bool fn_mask = meshes[shape].face_normals;
    its.tri = hit.tri;
    its.shape = shape;
    if constexpr (ad && PS) its.J = ti.face_area / detach(ti.face_area); else its.J = R(1.f);
    its.n = ti.face_normal;
    V2<R> tuv0(R(tuv[0].x), R(tuv[0].y)), tuve1(R(tuv[1].x - tuv[0].x), R(tuv[1].y - tuv[0].y)), tuve2(R(tuv[2].x - tuv[0].x), R(tuv[2].y - tuv[0].y));
    if constexpr (ad) {   // Mesh.vertex_uv tangents (m_triangle_uv is gathered from the AD vertex uvs, mesh.cpp:232-236)
        const auto &tt = tri_uv_t[hit.tri];
        tuv0.x.d = tt[0].x; tuv0.y.d = tt[0].y;
        tuve1.x.d = tt[1].x - tt[0].x; tuve1.y.d = tt[1].y - tt[0].y;
        tuve2.x.d = tt[2].x - tt[0].x; tuve2.y.d = tt[2].y - tt[0].y;
    }
    if constexpr (!ad || PS) {
        V2<R> uv(R(hit.u), R(hit.v));
        V3<R> sh_n = normalize(bilinear(ti.n0, ti.n1 - ti.n0, ti.n2 - ti.n0, uv));
        if (fn_mask) sh_n = its.n;
        its.p = bilinear(ti.p0, ti.e1, ti.e2, uv);
        V3<R> dir = its.p - ray.o;
        its.t = norm(dir);
        dir = dir / its.t;
        its.sh = Frame<R>(sh_n);
        its.wi = its.sh.to_local(-dir);
        its.uv = bilinear2(tuv0, tuve1, tuve2, uv);
    } else {
        R u, v, t;
        ray_intersect_triangle<R>(ti.p0, ti.e1, ti.e2, ray, u, v, t);
        V2<R> uv(u, v);
        V3<R> sh_n = normalize(bilinear(ti.n0, ti.n1 - ti.n0, ti.n2 - ti.n0, uv));
        if (fn_mask) sh_n = its.n;
        its.p = ray(t);
        its.t = t;
        its.sh = Frame<R>(sh_n);
        its.wi = its.sh.to_local(-ray.d);
        its.uv = bilinear2(tuv0, tuve1, tuve2, uv);
    }
    return its;
}

// =================================================================================================
// Environment map: src/emitter/envmap.cpp (all detached in sampling; eval differentiable through to_world only
// for transforms the oracle does not parameterise, so it is evaluated on floats lifted to R)
// =================================================================================================
inline void ray_intersect_scene_aabb(const V3f &o, const V3f &d, const V3f &lo, const V3f &hi, float &t, V3f &n, float &G) {   // utils.h:129-145
    V3f t1 = (lo - o) / d, t2 = (hi - o) / d;
    V3f t2p(max_(t1.x, t2.x), max_(t1.y, t2.y), max_(t1.z, t2.z));
    int idx = 0;
    t = t2p.x;
    if (t2p.y < t) { t = t2p.y; idx = 1; }
    if (t2p.z < t) { t = t2p.z; idx = 2; }
    n = V3f();
    n[idx] = -sign1(d[idx]);
    G = dot(n, -d) * (1.f / sqr(t));
}
inline V3f env_eval_direction(const Emitter &e, const V3f &wi_world) {   // envmap.cpp:42-58
    V3f wi = transform_dir(e.env_from_world, wi_world);
    V2f uv(atan2_(wi.x, -wi.z) * kInvTwoPi, safe_acos(wi.y) * kInvPi);
    uv.x = uv.x - std::floor(uv.x); uv.y = uv.y - std::floor(uv.y);
    V3f r = e.env_radiance.eval3<float>(uv, false);
    return r * e.env_scale;
}
// envmap.cpp:42-58 in its ad = true flavour: attached to the direction (through atan2 / acos and the bilinear weights), to the
// radiance texels and to the scale. m_from_world is a constant here (the oracle does not parameterise the envmap transform).
inline V3<Dual> env_eval_direction_d(const Emitter &e, const V3<Dual> &wi_world) {
    V3<Dual> wi = transform_dir(e.env_from_world_d, wi_world);   // m_from_world is attached (envmap.cpp:46)
    V2<Dual> uv(atan2_(wi.x, -wi.z) * kInvTwoPi, safe_acos(wi.y) * kInvPi);
    uv.x = uv.x - floor_(uv.x); uv.y = uv.y - floor_(uv.y);
    V3<Dual> r = e.env_radiance.eval3<Dual>(uv, false);
    return r * Dual(e.env_scale, e.env_scale_t);
}
inline void configure_envmap(Emitter &e) {   // envmap.cpp:10-26
    int w = e.env_radiance.w, h = e.env_radiance.h;
    if (!(w > 1 && h > 1)) throw std::runtime_error("envmap resolution");
    int res[2] = {(w - 1) << 1, (h - 1) << 1};   // cells (x, y); the last dimension (y) is the fastest
    e.env_distrb.set_resolution(res);
    std::vector<float> mass((size_t)res[0] * res[1]);
    for (int i = 0; i < res[0]; ++i)
        for (int j = 0; j < res[1]; ++j) {
            V2f uv(((float)i + .5f) * e.env_distrb.unit[0], ((float)j + .5f) * e.env_distrb.unit[1]);
            V3f c = e.env_radiance.eval3<float>(uv, false);
            float theta = ((float)j + .5f) * (kPi / (float)res[1]);
            mass[(size_t)i * res[1] + j] = rgb2luminance(c) * std::sin(theta);
        }
    e.env_distrb.set_mass(mass);
    e.env_to_world = e.env_left * e.env_to_world_raw;
    e.env_from_world = inverse(e.env_to_world);
    {
        M4<Dual> left_d(e.env_left);
        if (e.env_has_t) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) left_d.m[i][j].d = e.env_left_t.m[i][j];
        e.env_from_world_d = inverse(left_d * M4<Dual>(e.env_to_world_raw));
    }
    // m_sampling_weight keeps its default 1.f (emitter.h:27) until Scene::configure normalises it
}

// =================================================================================================
// BSDFs: src/bsdf/diffuse.cpp, src/bsdf/ggx.cpp, src/bsdf/roughconductor.cpp
// =================================================================================================
namespace ggx {
template <class R> inline R eval(R au, R av, const V3<R> &m) {   // ggx.cpp:15-34
    R alpha_uv = au * av, cos_theta = m.z;
    R result = R(1.f) / (kPi * alpha_uv * sqr(sqr(m.x / au) + sqr(m.y / av) + sqr(m.z)));
    return val(result * cos_theta) > 1e-5f ? result : R(0.f);
}
template <class R> inline R smith_g1(R au, R av, const V3<R> &v, const V3<R> &m) {   // ggx.cpp:79-93
    R xy_alpha_2 = sqr(au * v.x) + sqr(av * v.y);
    R tan_theta_alpha_2 = xy_alpha_2 / sqr(v.z);
    R result = R(2.f) / (R(1.f) + sqrt_(R(1.f) + tan_theta_alpha_2));
    if (val(xy_alpha_2) == 0.f) result = R(1.f);
    if (val(dot(v, m) * v.z) <= 0.f) result = R(0.f);
    return result;
}
template <class R> inline V2<R> sample_visible_11(R cos_theta_i, const V2<R> &sample) {   // ggx.cpp:96-105
    V2<R> p = square_to_uniform_disk_concentric(sample);
    R s = .5f * (R(1.f) + cos_theta_i);
    R a = safe_sqrt(R(1.f) - sqr(p.x));
    p.y = fma_(p.y, s, fma_(-a, s, a));   // lerp(a, p.y, s) = fma(b,t, fnma(a,t,a))
    R x = p.x, y = p.y, z = safe_sqrt(R(1.f) - dot(p, p));
    R sin_theta_i = safe_sqrt(R(1.f) - sqr(cos_theta_i));
    R nrm = R(1.f) / fma_(sin_theta_i, y, cos_theta_i * z);
    return V2<R>(fma_(cos_theta_i, y, -(sin_theta_i * z)) * nrm, x * nrm);
}
template <class R> inline V3<R> sample(R au, R av, const V3<R> &wi, const V3<R> &smp) {   // ggx.cpp:37-76
    V3<R> wi_p = normalize(V3<R>(au * wi.x, av * wi.y, wi.z));
    // frame.h sin_phi / cos_phi
    R sin_theta_2 = fma_(wi_p.x, wi_p.x, sqr(wi_p.y));
    R inv_sin_theta = R(1.f) / sqrt_(sin_theta_2);
    bool degenerate = std::fabs(val(sin_theta_2)) <= 4.f * kEpsilon;   // frame.h:103-117 (psdr Epsilon)
    R sin_phi = degenerate ? R(0.f) : clamp_(wi_p.y * inv_sin_theta, -1.f, 1.f);
    R cos_phi = degenerate ? R(1.f) : clamp_(wi_p.x * inv_sin_theta, -1.f, 1.f);
    V2<R> slope = sample_visible_11<R>(wi_p.z, V2<R>(smp.x, smp.y));
    slope = V2<R>(fma_(cos_phi, slope.x, -(sin_phi * slope.y)) * au, fma_(sin_phi, slope.x, cos_phi * slope.y) * av);
    return normalize(V3<R>(-slope.x, -slope.y, R(1.f)));
}
}  // namespace ggx

template <class R> inline V3<R> fresnel(const V3<R> &eta_r, const V3<R> &eta_i, R cos_theta_i) {   // utils.h:149-164
    R c2 = sqr(cos_theta_i), s2 = R(1.f) - c2, s4 = sqr(s2);
    V3<R> out;
    for (int k = 0; k < 3; ++k) {
        R temp_1 = sqr(eta_r[k]) - sqr(eta_i[k]) - s2;
        R a_2_pb_2 = safe_sqrt(sqr(temp_1) + 4.f * sqr(eta_i[k] * eta_r[k]));
        R a = safe_sqrt(.5f * (a_2_pb_2 + temp_1));
        R term_1 = a_2_pb_2 + c2, term_2 = 2.f * cos_theta_i * a;
        R r_s = (term_1 - term_2) / (term_1 + term_2);
        R term_3 = a_2_pb_2 * c2 + s4, term_4 = term_2 * s2;
        R r_p = r_s * (term_3 - term_4) / (term_3 + term_4);
        out[k] = .5f * (r_s + r_p);
    }
    return out;
}

template <class R> inline V3<R> Scene::bsdf_eval(const Intersection<R> &its, const V3<R> &wo, bool active) const {
    if (!active || its.shape < 0 || meshes[its.shape].bsdf < 0) return V3<R>();
    const Bsdf &b = bsdfs[meshes[its.shape].bsdf];
    R cos_i = its.wi.z, cos_o = wo.z;
    if (b.type == BSDF_DIFFUSE) {   // diffuse.cpp:25-33
        if (!(val(cos_i) > 0.f && val(cos_o) > 0.f)) return V3<R>();
        return b.reflectance.eval3<R>(its.uv) * R(kInvPi) * cos_o;
    }
    // roughconductor.cpp:40-56
    if (!(val(cos_i) > 0.f && val(cos_o) > 0.f)) return V3<R>();
    R au = b.alpha_u.eval1<R>(its.uv), av = b.alpha_v.eval1<R>(its.uv);
    V3<R> H = normalize(wo + its.wi);
    R D = ggx::eval<R>(au, av, H);
    if (val(D) == 0.f) return V3<R>();
    R G = ggx::smith_g1<R>(au, av, its.wi, H) * ggx::smith_g1<R>(au, av, wo, H);
    R result = D * G / (4.f * its.wi.z);
    V3<R> F = fresnel<R>(b.eta.eval3<R>(its.uv), b.k.eval3<R>(its.uv), dot(its.wi, H));
    V3<R> spec = b.specular_reflectance.eval3<R>(its.uv);
    return F * result * spec;
}
template <class R> inline R Scene::bsdf_pdf(const Intersection<R> &its, const V3<R> &wo, bool active) const {
    if (its.shape < 0 || meshes[its.shape].bsdf < 0) return R(0.f);
    const Bsdf &b = bsdfs[meshes[its.shape].bsdf];
    if (b.type == BSDF_DIFFUSE) {   // diffuse.cpp:69-82 (detached cosines)
        if (!active) return R(0.f);
        float cos_i = val(its.wi.z), cos_o = val(wo.z);
        if (!(cos_i > 0.f && cos_o > 0.f)) return R(0.f);
        return R(kInvPi * cos_o);
    }
    // roughconductor.cpp:60-75: `active` is computed but never applied to the result
    R cos_i = its.wi.z;
    V3<R> m = normalize(wo + its.wi);
    R au = b.alpha_u.eval1<R>(its.uv), av = b.alpha_v.eval1<R>(its.uv);
    return ggx::eval<R>(au, av, m) * ggx::smith_g1<R>(au, av, its.wi, m) / (4.f * cos_i);
}
template <class R> inline BSDFSample<R> Scene::bsdf_sample(const Intersection<R> &its, const V3<R> &sample, bool active) const {
    BSDFSample<R> bs;
    if (its.shape < 0 || meshes[its.shape].bsdf < 0) return bs;
    const Bsdf &b = bsdfs[meshes[its.shape].bsdf];
    R cos_i = its.wi.z;
    if (b.type == BSDF_DIFFUSE) {   // diffuse.cpp:47-55: uses tail<2>(sample)
        bs.wo = square_to_cosine_hemisphere(V2<R>(sample.y, sample.z));
        bs.pdf = R(kInvPi) * bs.wo.z;
        bs.valid = active && val(cos_i) > 0.f;
        return bs;
    }
    // roughconductor.cpp:79-93
    R au = b.alpha_u.eval1<R>(its.uv), av = b.alpha_v.eval1<R>(its.uv);
    V3<R> m = ggx::sample<R>(au, av, its.wi, sample);
    R two_dot = 2.f * dot(its.wi, m);
    bs.wo = V3<R>(fma_(m.x, two_dot, -its.wi.x), fma_(m.y, two_dot, -its.wi.y), fma_(m.z, two_dot, -its.wi.z));
    bs.pdf = bsdf_pdf<R>(its, bs.wo, active);
    bs.valid = active && val(cos_i) > 0.f && val(bs.pdf) != 0.f && val(bs.wo.z) > 0.f;
    return bs;
}

// =================================================================================================
// Emitter queries
// =================================================================================================
template <class R> inline V3<R> Scene::Le(const Intersection<R> &its, bool active) const {
    if (!active || its.shape < 0) return V3<R>();
    int e = meshes[its.shape].emitter;
    if (e < 0) return V3<R>();
    const Emitter &em = emitters[e];
    if (em.type == EMITTER_AREA) return val(its.wi.z) > 0.f ? lift<R>(em.radiance) : V3<R>();   // area.cpp:20-29
    // envmap.cpp:29-39: radiance along -wi (world)
    if constexpr (std::is_same_v<R, Dual>) {
        return env_eval_direction_d(em, -its.sh.to_world(its.wi));   // envmap.cpp:35-38: attached to the direction, the texels and the scale
    } else {
        V3f wi_world = its.sh.to_world(its.wi);
        return env_eval_direction(em, -wi_world);
    }
}
// mesh.cpp:306-330
template <class R> inline PositionSample<R> Scene::mesh_sample_position(const Mesh &m, V2<R> sample2) const {
    PositionSample<R> r;
    float sx = val(sample2.x);
    auto [idx, pdf_unused] = m.face_distrb.sample_reuse(sx);
    (void)pdf_unused;
    sample2.x = R(sx);
    sample2 = square_to_uniform_triangle(sample2);
    TriangleInfo<R> ti = cast_tri<R>(m.tri[idx]);
    if constexpr (std::is_same_v<R, Dual>) r.J = ti.face_area / detach(ti.face_area); else r.J = 1.f;
    r.p = bilinear(ti.p0, ti.e1, ti.e2, sample2);
    r.n = ti.face_normal;
    r.pdf = m.inv_total_area;
    r.valid = true;
    return r;
}
// scene.cpp:427-447, area.cpp:46-50, envmap.cpp:72-95
template <class R> inline PositionSample<R> Scene::sample_emitter_position(const V3<R> &ref_p, V2<R> sample2, bool active) const {
    if (emitters.empty()) throw std::runtime_error("No Emitter!");
    int ei = 0;
    float emitter_pdf = 1.f;
    if (emitters.size() > 1) {
        float sy = val(sample2.y);
        auto pr = emitters_distrb.sample_reuse(sy);
        ei = pr.first; emitter_pdf = pr.second;
        sample2.y = R(sy);
    }
    const Emitter &em = emitters[ei];
    PositionSample<R> r;
    if (em.type == EMITTER_AREA) {
        r = mesh_sample_position<R>(meshes[em.mesh], sample2);
    } else {
        float s[2] = {val(sample2.x), val(sample2.y)};
        float pdf_dir = em.env_distrb.sample_reuse(s);   // envmap.cpp:98-111
        float theta = s[1] * kPi, phi = s[0] * kTwoPi;
        V3f sd(std::cos(phi) * std::sin(theta), std::sin(phi) * std::sin(theta), std::cos(theta));   // sphdir (utils.h:41-45)
        V3f d(sd.y, sd.z, -sd.x);
        float inv_sin_theta = 1.f / safe_sqrt(max_(sqr(d.x) + sqr(d.z), sqr(kEpsilon)));
        if (pdf_dir > kEpsilon) pdf_dir *= inv_sin_theta * (.5f / sqr(kPi));
        d = transform_dir(em.env_to_world, d);
        float t, G; V3f n;
        ray_intersect_scene_aabb(detach(ref_p), d, em.lower, em.upper, t, n, G);
        V3f p = detach(ref_p) + d * t;
        r.p = lift<R>(p); r.n = lift<R>(n); r.J = R(1.f);
        r.pdf = pdf_dir * G;
        r.valid = true;
    }
    if (!active) r.valid = false;
    r.pdf *= emitter_pdf;
    return r;
}
// scene.cpp:451-453, area.cpp:58-62, mesh.cpp:333-342, envmap.cpp:125-143
template <class R> inline float Scene::emitter_position_pdf(const V3<R> &ref_p, const Intersection<R> &its, bool active) const {
    if (!active || its.shape < 0) return 0.f;
    int e = meshes[its.shape].emitter;
    if (e < 0) return 0.f;
    const Emitter &em = emitters[e];
    if (em.type == EMITTER_AREA) return em.sampling_weight * meshes[its.shape].inv_total_area;
    V3f d = detach(its.p) - detach(ref_p);
    float dist2 = squared_norm(d);
    d = d / safe_sqrt(dist2);
    float G = std::fabs(dot(d, detach(its.n))) / dist2;
    d = transform_dir(em.env_from_world, d);
    float factor = G * (1.f / safe_sqrt(max_(sqr(d.x) + sqr(d.z), sqr(kEpsilon)))) * (.5f / sqr(kPi));
    float uvq[2] = {atan2_(d.x, -d.z) * kInvTwoPi, safe_acos(d.y) * kInvPi};
    uvq[0] -= std::floor(uvq[0]); uvq[1] -= std::floor(uvq[1]);
    return em.env_distrb.pdf(uvq) * factor;   // NB no sampling_weight factor (envmap.cpp:125-143)
}

// scene.cpp:456-492
inline BoundarySegSampleDirect Scene::sample_boundary_segment_direct(V3f sample3, bool active) const {
    BoundarySegSampleDirect r;
    float sample1 = sample3.x;
    auto [edge_idx, pdf0] = sec_edge_distrb.sample_reuse(sample1);
    const SecEdge<Dual> &info = sec_edges[edge_idx];
    r.p0 = V3<Dual>(fma_(info.e1.x, Dual(sample1), info.p0.x), fma_(info.e1.y, Dual(sample1), info.p0.y), fma_(info.e1.z, Dual(sample1), info.p0.z));
    V3f e1 = detach(info.e1);
    r.edge = normalize(e1);
    r.edge2 = detach(info.p2) - detach(info.p0);
    V3f p0 = detach(r.p0);
    pdf0 /= norm(e1);
    PositionSample<float> ps2 = sample_emitter_position<float>(p0, V2f(sample3.y, sample3.z), active);
    r.p2 = ps2.p; r.n = ps2.n;
    V3f e = r.p2 - p0;
    float distSqr = squared_norm(e);
    e = e / safe_sqrt(distSqr);
    float cosTheta = dot(r.n, -e);
    int sgn0 = sign_eps(dot(detach(info.n0), e), kEdgeEpsilon), sgn1 = sign_eps(dot(detach(info.n1), e), kEdgeEpsilon);
    r.valid = active && cosTheta > kEpsilon && ((info.is_boundary && sgn0 != 0) || (!info.is_boundary && sgn0 * sgn1 < 0));
    r.pdf = r.valid ? pdf0 * ps2.pdf * (distSqr / cosTheta) : 0.f;
    return r;
}

// =================================================================================================
// Integrators: src/integrator/{integrator,direct,field}.cpp + the multi-bounce PathIntegrator this repo defines
// (depth-1 PathIntegrator == DirectIntegrator(1,1), checked in tests)
// =================================================================================================
enum IntegratorKind { INTEG_DIRECT = 0, INTEG_FIELD = 1, INTEG_PATH = 2 };
enum FieldKind { FIELD_SILHOUETTE = 0, FIELD_POSITION, FIELD_DEPTH, FIELD_GEONORMAL, FIELD_SHNORMAL, FIELD_UV };

struct Integrator {
    int kind = INTEG_DIRECT;
    int bsdf_samples = 1, light_samples = 1;   // direct.cpp:30-32
    bool hide_emitters = false;
    int field = FIELD_SILHOUETTE;
    int max_depth = 1;                         // path: number of scattering events
    std::vector<std::unique_ptr<HyperCube<3>>> warpper;   // direct.cpp:166-204 guiding grids per sensor
};

template <class R> inline R mis_weight(R pdf1, R pdf2) { R w1 = sqr(pdf1), w2 = sqr(pdf2); return w1 / (w1 + w2); }   // direct.cpp:17-21

// One BSDF-sampling connection from `its` (direct.cpp:67-113). On return *next (if given) holds the sampled hit.
template <class R>
inline V3<R> bsdf_branch(const Scene &scene, const Intersection<R> &its, bool active, const V3f &smp, bool use_mis, float inv_count,
                         Intersection<R> *next = nullptr, V3<R> *thru = nullptr, bool *next_active = nullptr) {
    constexpr bool ad = std::is_same_v<R, Dual>;
    BSDFSample<R> bs = scene.bsdf_sample<R>(its, lift<R>(smp), active);
    bool active1 = active && bs.valid;
    Ray<R> ray1(its.p, its.sh.to_world(bs.wo));
    Intersection<R> its1 = scene.ray_intersect<R, ad>(ray1, active1);
    active1 = active1 && its1.valid();
    bool cont = active1;
    active1 = active1 && scene.is_emitter(its1.shape);
    V3<R> bsdf_val;
    R pdf0;
    // In path mode the continuation needs the weight even when hit1 is not an emitter
    bool evalw = next ? cont : active1;
    if constexpr (ad) {
        V3<R> wo = its1.p - its.p;
        wo = wo / its1.t;
        bsdf_val = scene.bsdf_eval<R>(its, its.sh.to_local(wo), evalw);
        R cos_val = dot(its1.n, -wo);
        R G_val = abs_(cos_val) / sqr(its1.t);
        pdf0 = bs.pdf * detach(G_val);
        bsdf_val = bsdf_val * (G_val * its1.J / pdf0);
    } else {
        bsdf_val = scene.bsdf_eval<R>(its, bs.wo, evalw);
        R cos_val = dot(its1.n, -ray1.d);
        R G_val = abs_(cos_val) / sqr(its1.t);
        pdf0 = bs.pdf * G_val;
        bsdf_val = bsdf_val / bs.pdf;
    }
    if (next) { *next = its1; *thru = bsdf_val; *next_active = cont; }
    if (!active1) return V3<R>();
    R weight = R(inv_count);
    if (use_mis) weight = weight * mis_weight<R>(pdf0, R(scene.emitter_position_pdf<R>(its.p, its1, active1)));
    return scene.Le<R>(its1, active1) * bsdf_val * weight;
}
// One emitter-sampling connection from `its` (direct.cpp:119-159)
template <class R>
inline V3<R> light_branch(const Scene &scene, const Intersection<R> &its, bool active, const V2f &smp, bool use_mis, float inv_count) {
    constexpr bool ad = std::is_same_v<R, Dual>;
    PositionSample<R> ps = scene.sample_emitter_position<R>(its.p, V2<R>(R(smp.x), R(smp.y)), active);
    bool active1 = active && ps.valid;
    V3<R> wo = ps.p - its.p;
    R dist_sqr = squared_norm(wo);
    R dist = safe_sqrt(dist_sqr);
    wo = wo / dist;
    Ray<R> ray1(its.p, wo);
    Intersection<R> its1 = scene.ray_intersect<R, ad>(ray1, active1);
    active1 = active1 && its1.valid();
    active1 = active1 && (val(its1.t) > val(dist) - kShadowEpsilon) && scene.is_emitter(its1.shape);
    if (!active1) return V3<R>();
    R cos_val = dot(its1.n, -wo);
    R G_val = abs_(cos_val) / dist_sqr;
    V3<R> wo_local = its.sh.to_local(wo);
    V3<R> bsdf_val = scene.bsdf_eval<R>(its, wo_local, active1);
    R pdf1 = scene.bsdf_pdf<R>(its, wo_local, active1);
    bsdf_val = bsdf_val * (G_val * ps.J / ps.pdf);
    if constexpr (ad) pdf1 = pdf1 * detach(G_val); else pdf1 = pdf1 * G_val;
    R weight = R(inv_count);
    if (use_mis) weight = weight * mis_weight<R>(R(ps.pdf), pdf1);
    return scene.Le<R>(its1, active1) * bsdf_val * weight;
}

// direct.cpp:47-163 / field.cpp:34-54 / PathIntegrator. `smp` advances one lane of the sampler in lock-step
// (every lane draws whether or not it is active, as the wavefront does).
template <class R>
inline V3<R> Li(const Integrator &I, const Scene &scene, SamplerLane &smp, const Ray<R> &ray, bool active = true) {
    Intersection<R> its = scene.ray_intersect<R, false>(ray, I.kind == INTEG_FIELD ? true : active);
    if (I.kind == INTEG_FIELD) {
        if (!(active && its.valid())) return V3<R>();
        switch (I.field) {
            case FIELD_SILHOUETTE: return V3<R>(R(1.f));
            case FIELD_POSITION: return its.p;
            case FIELD_DEPTH: return V3<R>(its.t);
            case FIELD_GEONORMAL: return its.n;
            case FIELD_SHNORMAL: return its.sh.n;
            default: return V3<R>(its.uv.x, its.uv.y, R(0.f));
        }
    }
    active = active && its.valid();
    V3<R> result = I.hide_emitters ? V3<R>() : scene.Le<R>(its, active);
    if (scene.emitter_env >= 0) active = active && its.shape >= 0 && scene.meshes[its.shape].bsdf >= 0;
    if (I.kind == INTEG_DIRECT) {
        for (int i = 0; i < I.bsdf_samples; ++i) {
            V3f s3 = smp.next_3d();
            result += bsdf_branch<R>(scene, its, active, s3, I.light_samples > 0, 1.f / (float)I.bsdf_samples);
        }
        for (int i = 0; i < I.light_samples; ++i) {
            V2f s2 = smp.next_2d();
            result += light_branch<R>(scene, its, active, s2, I.bsdf_samples > 0, 1.f / (float)I.light_samples);
        }
        return result;
    }
    // PathIntegrator: per scattering event one BSDF-sampled ray (emitter hit with MIS + continuation) and one NEE ray
    V3<R> throughput(R(1.f));
    for (int depth = 0; depth < I.max_depth; ++depth) {
        V3f s3 = smp.next_3d();
        V2f s2 = smp.next_2d();
        if (!active) continue;   // dead lanes keep drawing in lock-step
        Intersection<R> nxt;
        V3<R> w;
        bool nact = false;
        V3<R> c_b = bsdf_branch<R>(scene, its, active, s3, true, 1.f, &nxt, &w, &nact);
        V3<R> c_l = light_branch<R>(scene, its, active, s2, true, 1.f);
        result += throughput * (c_b + c_l);
        throughput = throughput * w;
        its = nxt;
        active = nact;
        if (scene.emitter_env >= 0) active = active && its.shape >= 0 && scene.meshes[its.shape].bsdf >= 0;
    }
    return result;
}

// integrator.cpp:64-95. out/out_t: W*H*3 interleaved image and (R = Dual) its tangent.
// scene.cpp:404: emitter sampling on a scene without emitters is an error. Raised here, before the OpenMP loops (an exception must not leave
// a parallel region): Li samples emitters when the integrator has light samples (every Path event does), the secondary-edge sampler always.
inline void require_emitters(const Integrator &I, const Scene &scene, bool edge_sampler = false) {
    const bool needs = edge_sampler || (I.kind == INTEG_DIRECT && I.light_samples > 0) || I.kind == INTEG_PATH;
    if (needs && scene.emitters.empty()) throw std::runtime_error("No Emitter!");
}
template <class R>
inline void render_interior(const Integrator &I, Scene &scene, int sensor_id, float *out, float *out_t) {
    if (!scene.ready) throw std::runtime_error("Input scene must be configured!");
    if (sensor_id < 0 || sensor_id >= (int)scene.sensors.size()) throw std::runtime_error("Invalid sensor id!");
    require_emitters(I, scene);
    const RenderOption &o = scene.opts;
    const int npix = o.width * o.height;
    if (o.spp <= 0) return;
    const Sensor &sensor = scene.sensors[sensor_id];
    auto &lanes = scene.samplers[0];
#pragma omp parallel for schedule(dynamic, 64)
    for (int pix = 0; pix < npix; ++pix) {
        float acc[3] = {0, 0, 0}, acc_t[3] = {0, 0, 0};
        for (int s = 0; s < o.spp; ++s) {
            int64_t lane = (int64_t)pix * o.spp + s;
            SamplerLane &smp = lanes[lane];
            V2f j = smp.next_2d();
            float bx = (float)(pix % o.width), by = (float)(pix / o.width);
            V2<R> samples(R((bx + j.x) / (float)o.width), R((by + j.y) / (float)o.height));
            Ray<R> ray = sensor.sample_primary_ray<R>(samples);
            V3<R> value = Li<R>(I, scene, smp, ray);
            zero_nonfinite(value);
            for (int k = 0; k < 3; ++k) { acc[k] += val(value[k]); acc_t[k] += tan_(value[k]); }
        }
        for (int k = 0; k < 3; ++k) {
            if (o.spp > 1) { acc[k] /= (float)o.spp; acc_t[k] /= (float)o.spp; }
            out[3 * pix + k] += acc[k];
            if (out_t) out_t[3 * pix + k] += acc_t[k];
        }
    }
}

// integrator.cpp:98-119. Adds only a tangent (value -= detach(value)).
inline void render_primary_edges(const Integrator &I, Scene &scene, int sensor_id, float *out_t) {
    const RenderOption &o = scene.opts;
    const Sensor &sensor = scene.sensors[sensor_id];
    if (!sensor.enable_edges) return;
    require_emitters(I, scene);
    const int64_t n = (int64_t)o.width * o.height * o.sppe;
    auto &lanes = scene.samplers[1];
    std::vector<float> contrib((size_t)n * 3, 0.f);
    std::vector<int> cidx((size_t)n, -1);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t lane = 0; lane < n; ++lane) {
        SamplerLane &smp = lanes[lane];
        PrimaryEdgeSample<Dual> es = sensor.sample_primary_edge(smp.next_1d());
        bool valid = es.idx >= 0;
        // operator-(Li(ray_n), Li(ray_p)): GCC evaluates the right operand first
        V3f Lp = Li<float>(I, scene, smp, es.ray_p, valid);
        V3f Ln = Li<float>(I, scene, smp, es.ray_n, valid);
        V3f delta = Ln - Lp;
        V3<Dual> value;
        for (int k = 0; k < 3; ++k) value[k] = es.x_dot_n * Dual(delta[k] / es.pdf);
        zero_nonfinite(value);
        if (o.sppe > 1) for (int k = 0; k < 3; ++k) value[k] = value[k] / (float)o.sppe;
        if (valid) { cidx[lane] = es.idx; for (int k = 0; k < 3; ++k) contrib[3 * lane + k] = value[k].d; }
    }
    for (int64_t lane = 0; lane < n; ++lane)
        if (cidx[lane] >= 0) for (int k = 0; k < 3; ++k) out_t[3 * cidx[lane] + k] += contrib[3 * lane + k];
}

// direct.cpp:225-316. Returns pixel index (-1 if none); value0 (guiding) when !ad, tangent-only value when ad.
template <bool ad>
inline int eval_secondary_edge(const Scene &scene, const Sensor &sensor, const V3f &sample3, V3f &out_val) {
    out_val = V3f();
    BoundarySegSampleDirect bss = scene.sample_boundary_segment_direct(sample3);
    bool valid = bss.valid;
    V3f _p0 = detach(bss.p0), _p2 = bss.p2, _dir = normalize(_p2 - _p0);
    TriangleInfo<Dual> tri_info;
    Intersection<float> _its2 = scene.ray_intersect<float, false>(Ray<float>(_p0, _dir), valid, ad ? &tri_info : nullptr);
    valid = valid && _its2.valid() && norm(_its2.p - _p2) < kShadowEpsilon;
    Intersection<float> _its1 = scene.ray_intersect<float, false>(Ray<float>(_p0, -_dir), valid);
    valid = valid && _its1.valid();
    V3f _p1 = _its1.p;
    SensorDirectSample sds = sensor.sample_direct(_p1);
    valid = valid && sds.valid;
    V3f cam_d;
    Intersection<Dual> its1;
    if constexpr (ad) {
        Ray<Dual> camera_ray = sensor.sample_primary_ray<Dual>(V2<Dual>(Dual(sds.q.x), Dual(sds.q.y)));
        its1 = scene.ray_intersect<Dual, false>(camera_ray, valid);
        valid = valid && its1.valid() && norm(detach(its1.p) - _p1) < kShadowEpsilon;
        cam_d = detach(camera_ray.d);
    } else {
        Ray<float> camera_ray = sensor.sample_primary_ray<float>(sds.q);
        Intersection<float> it = scene.ray_intersect<float, false>(camera_ray, valid);
        valid = valid && it.valid() && norm(it.p - _p1) < kShadowEpsilon;
        cam_d = camera_ray.d;
    }
    float dist = norm(_p2 - _p1), cos2 = std::fabs(dot(bss.n, -_dir));
    V3f e = cross(bss.edge, _dir);
    float sinphi = norm(e);
    V3f proj = normalize(cross(e, bss.n));
    float sinphi2 = norm(cross(_dir, proj));
    float base_v = (_its1.t / dist) * (sinphi / sinphi2) * cos2;
    valid = valid && (sinphi > kEpsilon) && (sinphi2 > kEpsilon);
    V3f d0 = -cam_d;
    V3f d0_local = _its1.sh.to_local(d0);
    // direct.cpp:278-284: with one BSDF (or one mesh) in the scene the reference evaluates meshes[0]'s BSDF WITHOUT looking at the mesh that
    // was hit — so a boundary segment whose camera-side end point lies on the environment map's bounding mesh (no BSDF of its own, and, unlike
    // in Li (direct.cpp:58-61), not masked out here) is shaded with that BSDF. Seen by running the reference's own source
    // (tests/test_ref_render.py); reproduced here and in the CUDA kernel (pb_edges.cu) for parity.
    Intersection<float> _its1_b = _its1;
    if (_its1.valid() && (scene.bsdfs.size() == 1 || scene.meshes.size() == 1)) _its1_b.shape = 0;
    V3f bsdf_val = scene.bsdf_eval<float>(_its1_b, d0_local, valid);
    float correction = std::fabs((_its1.wi.z * dot(d0, _its1.n)) / (d0_local.z * dot(_dir, _its1.n)));
    if (valid) bsdf_val = bsdf_val * correction;
    V3f value0;
    if (valid) value0 = bsdf_val * scene.Le<float>(_its2, valid) * (base_v * sds.sensor_val / bss.pdf);
    if constexpr (ad) {
        V3f n = normalize(cross(bss.n, proj));
        value0 = value0 * (sign1(dot(e, bss.edge2)) * sign1(dot(e, n)));
        Ray<Dual> shadow_ray(its1.p, normalize(bss.p0 - its1.p));
        Dual u, v, t;
        ray_intersect_triangle<Dual>(tri_info.p0, tri_info.e1, tri_info.e2, shadow_ray, u, v, t);
        V3<Dual> u2 = bilinear(lift<Dual>(detach(tri_info.p0)), lift<Dual>(detach(tri_info.e1)), lift<Dual>(detach(tri_info.e2)), V2<Dual>(u, v));
        Dual nv = dot(lift<Dual>(n), u2);
        if (!valid) return -1;
        for (int k = 0; k < 3; ++k) out_val[k] = (Dual(value0[k]) * nv).d;   // result - detach(result)
        // the primal decides finiteness (direct.cpp:215)
        for (int k = 0; k < 3; ++k) if (!std::isfinite(value0[k] * nv.v)) out_val[k] = 0.f;
        return sds.pixel_idx;
    } else {
        out_val = value0;
        return -1;
    }
}

// direct.cpp:166-204
inline void preprocess_secondary_edges(Integrator &I, Scene &scene, int sensor_id, const int reso[4], int nrounds) {
    if (nrounds <= 0) throw std::runtime_error("nrounds > 0");
    if (!scene.ready) throw std::runtime_error("Scene needs to be configured!");
    require_emitters(I, scene, true);
    if (I.warpper.size() != scene.sensors.size()) I.warpper.resize(scene.sensors.size());
    if (!I.warpper[sensor_id]) I.warpper[sensor_id] = std::make_unique<HyperCube<3>>();
    HyperCube<3> &w = *I.warpper[sensor_id];
    w.set_resolution(reso);
    const int ncell = w.num_cells;
    const int64_t ns = (int64_t)ncell * reso[3];
    std::vector<float> result(ncell, 0.f);
    const Sensor &sensor = scene.sensors[sensor_id];
#pragma omp parallel for schedule(dynamic, 256)
    for (int c = 0; c < ncell; ++c) {
        int cc[3];
        w.cell(c, cc);
        float acc = 0.f;
        for (int r = 0; r < reso[3]; ++r) {
            int64_t lane = (int64_t)c * reso[3] + r;
            SamplerLane smp = SamplerLane::make((uint64_t)lane);
            for (int j = 0; j < nrounds; ++j) {
                V3f s = smp.next_3d();
                V3f s3(((float)cc[0] + s.x) * w.unit[0], ((float)cc[1] + s.y) * w.unit[1], ((float)cc[2] + s.z) * w.unit[2]);
                V3f v;
                eval_secondary_edge<false>(scene, sensor, s3, v);
                for (int k = 0; k < 3; ++k) if (!std::isfinite(v[k])) v[k] = 0.f;
                if (reso[3] > 1) v = v / (float)reso[3];
                acc += max_(max_(v.x, v.y), v.z);
            }
        }
        result[c] = acc;
    }
    if (nrounds > 1) for (auto &r : result) r /= (float)nrounds;
    w.set_mass(result);
}

// direct.cpp:207-221
inline void render_secondary_edges(const Integrator &I, Scene &scene, int sensor_id, float *out_t) {
    const RenderOption &o = scene.opts;
    const Sensor &sensor = scene.sensors[sensor_id];
    const int64_t n = (int64_t)o.width * o.height * o.sppse;
    require_emitters(I, scene, true);
    auto &lanes = scene.samplers[2];
    const HyperCube<3> *w = (I.warpper.empty() || !I.warpper[sensor_id]) ? nullptr : I.warpper[sensor_id].get();
    std::vector<float> contrib((size_t)n * 3, 0.f);
    std::vector<int> cidx((size_t)n, -1);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t lane = 0; lane < n; ++lane) {
        V3f s3 = lanes[lane].next_3d();
        float pdf0 = 1.f;
        if (w) { float s[3] = {s3.x, s3.y, s3.z}; pdf0 = w->sample_reuse(s); s3 = V3f(s[0], s[1], s[2]); }
        V3f value;
        int idx = eval_secondary_edge<true>(scene, sensor, s3, value);
        for (int k = 0; k < 3; ++k) if (!std::isfinite(value[k])) value[k] = 0.f;
        if (pdf0 > kEpsilon) value = value / pdf0;
        if (o.sppse > 1) value = value / (float)o.sppse;
        if (idx >= 0) { cidx[lane] = idx; for (int k = 0; k < 3; ++k) contrib[3 * lane + k] = value[k]; }
    }
    for (int64_t lane = 0; lane < n; ++lane)
        if (cidx[lane] >= 0) for (int k = 0; k < 3; ++k) out_t[3 * cidx[lane] + k] += contrib[3 * lane + k];
}

// integrator.cpp:13-60
inline void renderC(const Integrator &I, Scene &scene, int sensor_id, float *out) {
    std::fill(out, out + (size_t)scene.opts.width * scene.opts.height * 3, 0.f);
    render_interior<float>(I, scene, sensor_id, out, nullptr);
}
inline void renderD(const Integrator &I, Scene &scene, int sensor_id, float *out, float *out_t) {
    size_t n = (size_t)scene.opts.width * scene.opts.height * 3;
    std::fill(out, out + n, 0.f);
    std::fill(out_t, out_t + n, 0.f);
    render_interior<Dual>(I, scene, sensor_id, out, out_t);
    if (scene.opts.sppe > 0) render_primary_edges(I, scene, sensor_id, out_t);
    if (scene.opts.sppse > 0 && I.kind != INTEG_FIELD) render_secondary_edges(I, scene, sensor_id, out_t);
}

}  // namespace orc
