// ORACLE — TEST INFRASTRUCTURE ONLY (parity pinned against the reference's own source run on the CPU; see orc_math.hpp header and DESIGN.md §2).
//
// CPU restatement of psdr-cuda's scene layer: RNG, discrete / hyper-cube distributions, bitmap lookup,
// mesh preprocessing + edge lists, perspective sensor, exact closest-hit ray casting, Scene::configure.
// Every function cites the reference file:line it restates (paths relative to /root/reference).
#pragma once
#include <algorithm>
#include <array>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "orc_math.hpp"

namespace orc {

// =================================================================================================
// RNG: src/core/sampler.cpp:8-54, include/psdr/core/sampler.h:18-31, Enoki PCG32
// =================================================================================================
constexpr uint64_t kPCG32DefaultState = 0x853c49e6748fea9bULL;
constexpr uint64_t kPCG32Mult = 0x5851f42d4c957f2dULL;

// sampler.cpp:8-18 instantiated on 64-bit lanes: v0/v1 are 64-bit, `sum` is a 32-bit lane (wraps).
inline uint64_t sample_tea_64(uint64_t v0, uint64_t v1, int rounds = 4) {
    uint32_t sum = 0;
    for (int i = 0; i < rounds; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cULL) ^ (v1 + (uint64_t)sum) ^ ((v1 >> 5) + 0xc8013ea4ULL);
        v1 += ((v0 << 4) + 0xad90777dULL) ^ (v0 + (uint64_t)sum) ^ ((v0 >> 5) + 0x7e95761eULL);
    }
    return v0 + (v1 << 32);
}

struct PCG32 {
    uint64_t state = 0, inc = 0;
    void seed(uint64_t initstate, uint64_t initseq) {
        state = 0;
        inc = (initseq << 1) | 1u;
        next_uint32();
        state += initstate;
        next_uint32();
    }
    uint32_t next_uint32() {
        uint64_t old = state;
        state = old * kPCG32Mult + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    float next_float32() {
        uint32_t u = (next_uint32() >> 9) | 0x3f800000u;
        float f;
        std::memcpy(&f, &u, 4);
        return f - 1.f;
    }
};

// One lane of a psdr Sampler: stream `i` of a sampler seeded with arange(count) (sampler.cpp:29-40).
// The lane remembers how many draws it has consumed so that samplers persist across render calls (F8).
struct SamplerLane {
    PCG32 rng;
    static SamplerLane make(uint64_t lane, uint64_t seed_value_base = 0) {
        SamplerLane s;
        uint64_t seed_value = lane + seed_value_base + kPCG32DefaultState;  // seed_value += m_base_seed
        uint64_t idx = lane;
        s.rng.seed(sample_tea_64(seed_value, idx), sample_tea_64(idx, seed_value));
        return s;
    }
    float next_1d() { return rng.next_float32(); }
    // RNG-dimension order: GCC evaluates call arguments right-to-left: y first, then x.
    V2f next_2d() {
        float y = next_1d();
        float x = next_1d();
        return {x, y};
    }
    // concat(next_nd<1>(), next_nd<2>()) with right-to-left evaluation: s[2], s[1], s[0].
    V3f next_3d() {
        float z = next_1d();
        float y = next_1d();
        float x = next_1d();
        return {x, y, z};
    }
};

// =================================================================================================
// DiscreteDistribution: src/core/pmf.cpp:7-50
// =================================================================================================
struct DiscreteDistribution {
    int size = 0;
    float sum = 0.f;
    std::vector<float> pmf, cmf;
    void init(const std::vector<float> &p) {
        size = (int)p.size();
        pmf = p;
        cmf.resize(size);
        float acc = 0.f;
        for (int i = 0; i < size; ++i) { acc += p[i]; cmf[i] = acc; }   // inclusive psum, sequential fp32
        sum = acc;
    }
    // enoki binary_search(0, size-1, pred): first i in [0,size-1] with cmf[i] >= x (size-1 if none)
    int search(float x) const {
        int lo = 0, hi = size - 1;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if (cmf[mid] < x) lo = mid + 1; else hi = mid;
        }
        return lo;
    }
    // pmf.cpp:17-27
    std::pair<int, float> sample(float u) const {
        if (size == 1) return {0, 1.f};
        int idx = search(u * sum);
        return {idx, pmf[idx] / sum};
    }
    // pmf.cpp:30-50 (sample is rescaled in place for reuse)
    std::pair<int, float> sample_reuse(float &u) const {
        if (size == 1) return {0, 1.f};
        u *= sum;
        int idx = search(u);
        if (idx > 0) u -= cmf[idx - 1];
        float p = pmf[idx];
        if (p > 0.f) u /= p;
        u = std::min(std::max(u, 0.f), 1.f);
        return {idx, p / sum};
    }
    float pmf_normalized(int i) const { return pmf[i] / sum; }
};

// =================================================================================================
// HyperCubeDistribution<n>: src/core/cube_distrb.cpp:8-62 (last dimension fastest)
// =================================================================================================
template <int N> struct HyperCube {
    int reso[N] = {0};
    int num_cells = 0;
    float unit[N];
    DiscreteDistribution distrb;
    bool ready = false;
    void set_resolution(const int *r) {
        int64_t prod = 1;
        for (int i = 0; i < N; ++i) { reso[i] = r[i]; unit[i] = 1.f / (float)r[i]; prod *= r[i]; }
        num_cells = (int)prod;
        ready = false;
    }
    void cell(int idx, int *c) const {
        for (int i = N - 1; i >= 0; --i) { c[i] = idx % reso[i]; idx /= reso[i]; }
    }
    void set_mass(const std::vector<float> &pmf) { distrb.init(pmf); ready = true; }
    // cube_distrb.cpp:41-47
    float sample_reuse(float *s) const {
        auto [idx, pdf] = distrb.sample_reuse(s[N - 1]);
        int c[N];
        cell(idx, c);
        for (int i = 0; i < N; ++i) s[i] = (s[i] + (float)c[i]) * unit[i];
        return pdf * (float)num_cells;
    }
    // cube_distrb.cpp:51-62
    float pdf(const float *p) const {
        int idx = 0;
        bool valid = true;
        for (int i = 0; i < N; ++i) {
            int ip = (int)std::floor(p[i] * (float)reso[i]);
            valid = valid && ip >= 0 && ip < reso[i];
            idx = idx * (i == 0 ? 0 : reso[i]) + ip;
        }
        if (!valid) return 0.f;
        return distrb.pmf_normalized(idx) * (float)num_cells;
    }
};

// =================================================================================================
// Bitmap<C>: src/core/bitmap.cpp:43-89. Texels stored interleaved [pixel*C + ch]; `tang` = tangent (AD leaf).
// =================================================================================================
struct Bitmap {
    int w = 1, h = 1, c = 3;
    std::vector<float> data, tang;
    template <class R> R texel(int idx, int ch) const {
        if constexpr (std::is_same_v<R, Dual>) return Dual(data[idx * c + ch], tang.empty() ? 0.f : tang[idx * c + ch]);
        else return data[idx * c + ch];
    }
    template <class R> void eval(V2<R> uv, bool flip_v, R *out) const {
        if (w == 1 && h == 1) { for (int k = 0; k < c; ++k) out[k] = texel<R>(0, k); return; }
        if (flip_v) uv.y = -uv.y;
        uv.x = uv.x - floor_(uv.x); uv.y = uv.y - floor_(uv.y);
        uv.x = uv.x * (float)(w - 1); uv.y = uv.y * (float)(h - 1);
        int px = (int)std::floor(val(uv.x)), py = (int)std::floor(val(uv.y));
        R w1x = uv.x - (float)px, w1y = uv.y - (float)py, w0x = 1.f - w1x, w0y = 1.f - w1y;
        px = std::min(px, w - 2); py = std::min(py, h - 2);
        int idx = py * w + px;
        for (int k = 0; k < c; ++k) {
            R v00 = texel<R>(idx, k), v10 = texel<R>(idx + 1, k), v01 = texel<R>(idx + w, k), v11 = texel<R>(idx + w + 1, k);
            R v0 = fma_(w0x, v00, w1x * v10), v1 = fma_(w0x, v01, w1x * v11);
            out[k] = fma_(w0y, v0, w1y * v1);
        }
    }
    template <class R> V3<R> eval3(const V2<R> &uv, bool flip_v = true) const { R o[3]; eval<R>(uv, flip_v, o); return {o[0], o[1], o[2]}; }
    template <class R> R eval1(const V2<R> &uv, bool flip_v = true) const { R o[1]; eval<R>(uv, flip_v, o); return o[0]; }
    static Bitmap constant3(float r, float g, float b) { Bitmap t; t.c = 3; t.data = {r, g, b}; return t; }
    static Bitmap constant1(float v) { Bitmap t; t.c = 1; t.data = {v}; return t; }
};

// =================================================================================================
// Records (types.h:136-146, intersection.h:25-54, records.h:11-45, edge.h:28-65)
// =================================================================================================
template <class R> struct TriangleInfo {
    V3<R> p0, e1, e2, n0, n1, n2, face_normal;
    R face_area;
};
template <class R> struct SecEdge { V3<R> p0, e1, n0, n1, p2; bool is_boundary; };
template <class R> struct PrimEdge { V2<R> p0, p1; V2f edge_normal; float edge_length; };

template <class R> struct Intersection {
    V3<R> wi, p;
    R t = R(kInf);
    int shape = -1, tri = -1;   // mesh index (nullptr <=> -1) and global triangle id
    V3<R> n;
    Frame<R> sh;
    V2<R> uv;
    R J = R(1.f);
    bool valid() const { return shape >= 0; }
};
template <class R> struct PositionSample { V3<R> p, n; R J = R(1.f); float pdf = 0.f; bool valid = false; };
template <class R> struct BSDFSample { V3<R> wo; R pdf = R(0.f); bool valid = false; };

enum BsdfType { BSDF_DIFFUSE = 0, BSDF_ROUGHCONDUCTOR = 1 };
struct Bsdf {
    int type = BSDF_DIFFUSE;
    Bitmap reflectance = Bitmap::constant3(.5f, .5f, .5f);                    // diffuse.h:11-14
    Bitmap alpha_u = Bitmap::constant1(.1f), alpha_v = Bitmap::constant1(.1f); // roughconductor.h:11-12
    Bitmap eta = Bitmap::constant3(0.f, 0.f, 0.f), k = Bitmap::constant3(1.f, 1.f, 1.f);
    Bitmap specular_reflectance = Bitmap::constant3(1.f, 1.f, 1.f);
};

enum EmitterType { EMITTER_AREA = 0, EMITTER_ENVMAP = 1 };

struct Mesh {
    int nv = 0, nf = 0;
    std::vector<float> vraw, vraw_t;      // object-space positions (AoS xyz) + tangent (empty = 0)
    std::vector<int> faces;               // nf*3
    bool has_uv = false;
    std::vector<float> uvs;               // n*2
    std::vector<float> uvs_t;             // forward-mode tangent of the texture coordinates (Mesh.vertex_uv is an AD leaf, psdr.cpp:254), or empty
    std::vector<int> uv_faces;            // nf*3
    bool face_normals = false, enable_edges = true;
    int bsdf = -1, emitter = -1;
    M4f to_world_raw, left, right, left_t, right_t;   // *_t = tangents of the AD leaves (mesh.h:19-35)
    bool has_left_t = false, has_right_t = false;
    std::vector<int> edges;               // 5 per edge: v0, v1, f0, f1|-1, opposite vertex of f0 (mesh.cpp:143-203)
    // configured (stored as duals; float views detach):
    std::vector<V3<Dual>> vworld;
    std::vector<TriangleInfo<Dual>> tri;
    std::vector<SecEdge<Dual>> sec_edges;
    float total_area = 0.f, inv_total_area = 0.f;
    DiscreteDistribution face_distrb;
    int face_offset = 0;
    bool ready = false;
};

template <class R> inline TriangleInfo<R> cast_tri(const TriangleInfo<Dual> &t) {
    if constexpr (std::is_same_v<R, Dual>) return t;
    else {
        TriangleInfo<float> o;
        o.p0 = detach(t.p0); o.e1 = detach(t.e1); o.e2 = detach(t.e2);
        o.n0 = detach(t.n0); o.n1 = detach(t.n1); o.n2 = detach(t.n2);
        o.face_normal = detach(t.face_normal); o.face_area = t.face_area.v;
        return o;
    }
}

// mesh.cpp:143-203 — unique undirected edges in (min,max) order; [opp vertex of first face, f0, f1]
inline void build_edge_list(Mesh &m) {
    std::map<std::pair<int, int>, std::vector<int>> edge_map;
    for (int f = 0; f < m.nf; ++f)
        for (int i = 0; i < 3; ++i) {
            int i1 = m.faces[3 * f + i], i2 = m.faces[3 * f + (i + 1) % 3], i3 = m.faces[3 * f + (i + 2) % 3];
            auto key = i1 < i2 ? std::make_pair(i1, i2) : std::make_pair(i2, i1);
            auto it = edge_map.find(key);
            if (it == edge_map.end()) it = edge_map.insert({key, std::vector<int>{i3}}).first;
            it->second.push_back(f);
        }
    m.edges.clear();
    for (auto &kv : edge_map) {
        const auto &v = kv.second;
        if (v.size() > 3) throw std::runtime_error("Edge shared by more than 2 faces");
        if (v.size() == 3 && v[1] == v[2]) throw std::runtime_error("Duplicated faces");
        m.edges.push_back(kv.first.first);
        m.edges.push_back(kv.first.second);
        m.edges.push_back(v[1]);
        m.edges.push_back(v.size() == 3 ? v[2] : -1);
        m.edges.push_back(v[0]);
    }
}

// mesh.cpp:19-51 (process_mesh) applied to world-space positions; sequential fp32 scatter-adds in face order.
inline void process_mesh(const std::vector<V3<Dual>> &vp, const std::vector<int> &faces, int nf,
                         std::vector<TriangleInfo<Dual>> &tri, std::vector<V3<Dual>> *vnormals_out = nullptr) {
    using R = Dual;
    int nv = (int)vp.size();
    tri.resize(nf);
    std::vector<V3<R>> vn(nv);
    std::vector<R> vw(nv);
    std::vector<V3<R>> fn(nf);
    std::vector<R> fa(nf);
    for (int f = 0; f < nf; ++f) {
        auto &t = tri[f];
        t.p0 = vp[faces[3 * f]];
        t.e1 = vp[faces[3 * f + 1]] - t.p0;
        t.e2 = vp[faces[3 * f + 2]] - t.p0;
        fn[f] = cross(t.e1, t.e2);
        fa[f] = norm(fn[f]);
    }
    for (int i = 0; i < 3; ++i)
        for (int f = 0; f < nf; ++f) {
            int v = faces[3 * f + i];
            vn[v] += fn[f];
            vw[v] = vw[v] + fa[f];
        }
    for (int v = 0; v < nv; ++v) vn[v] = normalize(vn[v] / vw[v]);
    for (int f = 0; f < nf; ++f) {
        auto &t = tri[f];
        t.n0 = vn[faces[3 * f]]; t.n1 = vn[faces[3 * f + 1]]; t.n2 = vn[faces[3 * f + 2]];
        t.face_normal = fn[f] / fa[f];
        t.face_area = fa[f] * 0.5f;
    }
    if (vnormals_out) *vnormals_out = vn;
}

// mesh.cpp:215-274 (Mesh::configure)
inline void configure_mesh(Mesh &m) {
    using R = Dual;
    M4<R> L(m.left), Rm(m.right), W(m.to_world_raw);
    if (m.has_left_t) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) L.m[i][j].d = m.left_t.m[i][j];
    if (m.has_right_t) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Rm.m[i][j].d = m.right_t.m[i][j];
    M4<R> to_world = L * W * Rm;
    m.vworld.resize(m.nv);
    for (int v = 0; v < m.nv; ++v) {
        V3<R> p(R(m.vraw[3 * v]), R(m.vraw[3 * v + 1]), R(m.vraw[3 * v + 2]));
        if (!m.vraw_t.empty()) { p.x.d = m.vraw_t[3 * v]; p.y.d = m.vraw_t[3 * v + 1]; p.z.d = m.vraw_t[3 * v + 2]; }
        m.vworld[v] = transform_pos(to_world, p);
    }
    process_mesh(m.vworld, m.faces, m.nf, m.tri);
    float total = 0.f;
    std::vector<float> areas(m.nf);
    for (int f = 0; f < m.nf; ++f) { areas[f] = m.tri[f].face_area.v; total += areas[f]; }
    m.total_area = total;
    m.inv_total_area = 1.f / total;
    m.face_distrb.init(areas);
    m.sec_edges.clear();
    if (m.enable_edges) {
        if (m.edges.empty()) build_edge_list(m);
        int ne = (int)m.edges.size() / 5;
        for (int e = 0; e < ne; ++e) {
            const int *ed = &m.edges[5 * e];
            SecEdge<R> s;
            s.is_boundary = ed[3] < 0;
            s.p0 = m.vworld[ed[0]];
            s.e1 = m.vworld[ed[1]] - s.p0;
            s.n0 = m.tri[ed[2]].face_normal;
            s.n1 = s.is_boundary ? V3<R>() : m.tri[ed[3]].face_normal;
            s.p2 = m.vworld[ed[4]];
            if (val(dot(s.n0, s.n1)) < 1.f - kEdgeEpsilon) m.sec_edges.push_back(s);   // mesh.cpp:262-263
        }
    }
    m.ready = true;
}

// =================================================================================================
// Exact closest hit (stands in for OptiX: cuda/psdr_cuda.cu:9-45, scene_optix.cpp:80-126).
// Closest Möller–Trumbore hit (utils.h:67-77 arithmetic) with t in (1e-3, tmax); ties -> lowest triangle id.
// The BVH is only an accelerator: boxes are padded so it returns exactly what brute force returns.
// =================================================================================================
struct Hit { int tri = -1; float u = -1.f, v = -1.f, t = kInf; };

struct TriAccel {
    struct Node { float lo[3], hi[3]; int left, right, first, count; };
    std::vector<Node> nodes;
    std::vector<int> order;
    std::vector<V3f> p0, e1, e2;

    static bool test(const V3f &P0, const V3f &E1, const V3f &E2, const Ray<float> &ray, float tmax, float &u, float &v, float &t) {
        ray_intersect_triangle<float>(P0, E1, E2, ray, u, v, t);
        return u >= 0.f && v >= 0.f && u + v <= 1.f && t > kRayEpsilon && t < tmax;   // NaN (a == 0) fails every test
    }
    void build(const std::vector<V3f> &P0, const std::vector<V3f> &E1, const std::vector<V3f> &E2) {
        p0 = P0; e1 = E1; e2 = E2;
        int n = (int)p0.size();
        order.resize(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        std::vector<std::array<float, 3>> cen(n), blo(n), bhi(n);
        for (int i = 0; i < n; ++i) {
            V3f a = p0[i], b = p0[i] + e1[i], c = p0[i] + e2[i];
            for (int k = 0; k < 3; ++k) {
                blo[i][k] = std::min(a[k], std::min(b[k], c[k]));
                bhi[i][k] = std::max(a[k], std::max(b[k], c[k]));
                cen[i][k] = 0.5f * (blo[i][k] + bhi[i][k]);
            }
        }
        nodes.clear();
        nodes.reserve(2 * n);
        build_rec(0, n, cen, blo, bhi);
    }
    int build_rec(int first, int count, const std::vector<std::array<float, 3>> &cen,
                  const std::vector<std::array<float, 3>> &blo, const std::vector<std::array<float, 3>> &bhi) {
        int id = (int)nodes.size();
        nodes.push_back(Node());
        float lo[3] = {kInf, kInf, kInf}, hi[3] = {-kInf, -kInf, -kInf}, clo[3] = {kInf, kInf, kInf}, chi[3] = {-kInf, -kInf, -kInf};
        for (int i = first; i < first + count; ++i) {
            int t = order[i];
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], blo[t][k]); hi[k] = std::max(hi[k], bhi[t][k]);
                clo[k] = std::min(clo[k], cen[t][k]); chi[k] = std::max(chi[k], cen[t][k]);
            }
        }
        for (int k = 0; k < 3; ++k) {   // pad: the box test must never cull a triangle the exact test accepts
            float pad = 1e-4f * std::max(1.f, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
            nodes[id].lo[k] = lo[k] - pad; nodes[id].hi[k] = hi[k] + pad;
        }
        int axis = 0;
        for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
        if (count <= 4 || chi[axis] <= clo[axis]) {
            nodes[id].left = nodes[id].right = -1; nodes[id].first = first; nodes[id].count = count;
            return id;
        }
        int mid = first + count / 2;
        std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                         [&](int a, int b) { return cen[a][axis] < cen[b][axis]; });
        nodes[id].first = nodes[id].count = 0;
        int l = build_rec(first, mid - first, cen, blo, bhi);
        int r = build_rec(mid, first + count - mid, cen, blo, bhi);
        nodes[id].left = l; nodes[id].right = r;
        return id;
    }
    Hit closest(const Ray<float> &ray) const {
        Hit best;
        best.t = ray.tmax;
        if (nodes.empty()) { best.t = kInf; return best; }
        float inv[3] = {1.f / ray.d.x, 1.f / ray.d.y, 1.f / ray.d.z};
        int stack[128], sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node &nd = nodes[stack[--sp]];
            float t0 = 0.f, t1 = best.t;
            bool miss = false;
            for (int k = 0; k < 3; ++k) {
                float a = (nd.lo[k] - ray.o[k]) * inv[k], b = (nd.hi[k] - ray.o[k]) * inv[k];
                if (a > b) std::swap(a, b);
                if (a != a || b != b) {   // 0 * inf: direction component is 0 and the origin is on a slab plane
                    if (ray.o[k] < nd.lo[k] || ray.o[k] > nd.hi[k]) { miss = true; break; }
                    continue;
                }
                t0 = std::max(t0, a); t1 = std::min(t1, b);
                if (t0 > t1) { miss = true; break; }
            }
            if (miss) continue;
            if (nd.left < 0) {
                for (int i = nd.first; i < nd.first + nd.count; ++i) {
                    int t = order[i];
                    float u, v, tt;
                    if (test(p0[t], e1[t], e2[t], ray, kInf, u, v, tt) && tt < ray.tmax &&
                        (tt < best.t || (tt == best.t && (best.tri < 0 || t < best.tri)))) {
                        best.t = tt; best.u = u; best.v = v; best.tri = t;
                    }
                }
            } else { stack[sp++] = nd.left; stack[sp++] = nd.right; }
        }
        if (best.tri < 0) best.t = kInf;
        return best;
    }
    Hit closest_brute(const Ray<float> &ray) const {
        Hit best;
        best.t = ray.tmax;
        for (int t = 0; t < (int)p0.size(); ++t) {
            float u, v, tt;
            if (test(p0[t], e1[t], e2[t], ray, kInf, u, v, tt) && tt < ray.tmax && (tt < best.t || (tt == best.t && (best.tri < 0 || t < best.tri)))) {
                best.t = tt; best.u = u; best.v = v; best.tri = t;
            }
        }
        if (best.tri < 0) best.t = kInf;
        return best;
    }
};

}  // namespace orc
