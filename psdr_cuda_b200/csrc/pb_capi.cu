// psdr-b200: implementation of the C ABI declared in include/psdr_b200.h.
//
// Host side of Scene::configure (src/scene/scene.cpp:56-278) and of Integrator::renderC/renderD
// (src/integrator/integrator.cpp:13-95): owns the scene description, prepares the device tables, and drives the
// wavefront kernels batch by batch on the context's stream.
#include "../../include/psdr_b200.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>

#include "pb_adjoint_math.cuh"
#include "pb_host.h"

namespace pb {

// ---- small host matrix helpers (fp32, op order as written: s = s + a*b from 0) ---------------------------------
Mat4h matmul(const Mat4h &a, const Mat4h &b) {
    Mat4h c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s = s + a.m[4 * i + k] * b.m[4 * k + j];
            c.m[4 * i + j] = s;
        }
    return c;
}
// Gauss-Jordan with partial pivoting in double, rounded once to fp32
Mat4h inverse(const Mat4h &A) {
    double a[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { a[i][j] = A.m[4 * i + j]; a[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 8; ++j) std::swap(a[p][j], a[c][j]);
        const double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= inv;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = a[r][c];
            for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j];
        }
    }
    Mat4h R;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.m[4 * i + j] = (float)a[i][4 + j];
    return R;
}
static Mat4 to_dev(const Mat4h &m) { Mat4 r; std::memcpy(r.m, m.m, sizeof(r.m)); return r; }
static Mat4h from_ptr(const float *p) { Mat4h r = Mat4h::identity(); if (p) std::memcpy(r.m, p, sizeof(r.m)); return r; }

// unique undirected edges in (min,max) order: (v0, v1, f0, f1|-1, opposite vertex of f0) — src/shape/mesh.cpp:143-203
static void build_edges(HostMesh &m) {
    struct Half { int lo, hi, face, opp; };
    std::vector<Half> hs;
    hs.reserve(3 * (size_t)m.nf);
    for (int f = 0; f < m.nf; ++f)
        for (int i = 0; i < 3; ++i) {
            const int a = m.faces[3 * f + i], b = m.faces[3 * f + (i + 1) % 3], c = m.faces[3 * f + (i + 2) % 3];
            hs.push_back({std::min(a, b), std::max(a, b), f, c});
        }
    std::stable_sort(hs.begin(), hs.end(), [](const Half &x, const Half &y) { return x.lo != y.lo ? x.lo < y.lo : x.hi < y.hi; });
    m.edges.clear();
    for (size_t i = 0; i < hs.size();) {
        size_t j = i;
        while (j < hs.size() && hs[j].lo == hs[i].lo && hs[j].hi == hs[i].hi) ++j;
        const size_t cnt = j - i;
        PB_ASSERT_MSG(cnt <= 2, "Edge shared by more than 2 faces");
        PB_ASSERT_MSG(!(cnt == 2 && hs[i].face == hs[i + 1].face), "Duplicated faces");
        const int e[5] = {hs[i].lo, hs[i].hi, hs[i].face, cnt == 2 ? hs[i + 1].face : -1, hs[i].opp};
        m.edges.insert(m.edges.end(), e, e + 5);
        i = j;
    }
}

// vertex -> incident faces in the order the reference's three scatter_add passes visit them (mesh.cpp:34-37)
static void build_csr(HostMesh &m) {
    m.csr_off.assign(m.nv + 1, 0);
    for (int i = 0; i < 3 * m.nf; ++i) m.csr_off[m.faces[i] + 1]++;
    for (int v = 0; v < m.nv; ++v) m.csr_off[v + 1] += m.csr_off[v];
    m.csr_face.resize(3 * (size_t)m.nf);
    std::vector<int> cur(m.csr_off.begin(), m.csr_off.end() - 1);
    m.csr_slot.resize(3 * (size_t)m.nf);
    for (int i = 0; i < 3; ++i)
        for (int f = 0; f < m.nf; ++f) {
            const int k = cur[m.faces[3 * f + i]]++;
            m.csr_face[k] = f;
            m.csr_slot[k] = 3 * f + i;
        }
}

// EnvironmentMap::configure (envmap.cpp:10-26): luminance * sin(theta) mass of every cell of the 2(w-1) x 2(h-1) grid, evaluated and
// summed on the device (pb_tables.cu); the host contributes the ry values of sin(theta) (libm, as the reference evaluates them)
static void configure_envmap_distribution(pb_ctx *c, HostEmitter &e) {
    if (!e.env_dirty) return;
    const int w = e.env_radiance.w, h = e.env_radiance.h;
    PB_ASSERT_MSG(w > 1 && h > 1, "Environment map must be larger than 1x1");
    const int rx = (w - 1) << 1, ry = (h - 1) << 1;
    e.env_res[0] = rx; e.env_res[1] = ry;
    std::vector<float> sin_theta(ry);
    for (int j = 0; j < ry; ++j) sin_theta[j] = std::sin(((float)j + .5f) * (kPi / (float)ry));
    const size_t cells = (size_t)rx * ry;
    e.env_radiance.d.upload(e.env_radiance.data, c->stream);
    c->d_env_sin.upload(sin_theta, c->stream);
    e.d_env_pmf.reserve(cells * sizeof(float)); e.d_env_cmf.reserve(cells * sizeof(float));
    c->d_edge_out.reserve(16 * sizeof(int));
    launch_envmap_pmf(c->stream, rx, ry, w, h, e.env_radiance.d.as<float>(), c->d_env_sin.as<float>(), e.d_env_pmf.as<float>());
    launch_seq_cmf(c->stream, (long long)cells, e.d_env_pmf.as<float>(), e.d_env_cmf.as<float>(), c->d_edge_out.as<float>(), nullptr);
    c->launches += 2;
    PB_CUDA(cudaMemcpyAsync(&e.env_sum, c->d_edge_out.p, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(cudaStreamSynchronize(c->stream));   // also: sin_theta is a local
    e.env_dirty = false;
}

// perspective.cpp:11-32
static void configure_sensor(HostSensor &s, int W, int H) {
    const float aspect = (float)W / (float)H;
    const float *t = s.to_world.m;
    const float det = t[0] * (t[5] * t[10] - t[6] * t[9]) - t[1] * (t[4] * t[10] - t[6] * t[8]) + t[2] * (t[4] * t[9] - t[5] * t[8]);
    PB_ASSERT_MSG(std::fabs(det - 1.f) < kEpsilon, "Sensor transformation should not involve scaling!");
    Mat4h sc = Mat4h::identity(), tr = Mat4h::identity(), pe = Mat4h::identity();
    sc.m[0] = -0.5f; sc.m[5] = -0.5f * aspect;
    tr.m[3] = -1.f; tr.m[7] = -1.f / aspect;
    const float recip = 1.f / (s.far_clip - s.near_clip);
    const float tn = std::tan((s.fov_x * .5f) * (kPi / 180.f)), cot = 1.f / tn;   // transform.h:50: tan(deg_to_rad(fov * .5f)), deg_to_rad(a) = a * (Pi / 180)
    pe.m[0] = cot; pe.m[5] = cot; pe.m[10] = s.far_clip * recip; pe.m[15] = 0.f;
    pe.m[11] = -s.near_clip * s.far_clip * recip; pe.m[14] = 1.f;
    const Mat4h c2s = matmul(matmul(sc, tr), pe);
    const Mat4h s2c = inverse(c2s);
    const Mat4h w2s = matmul(c2s, inverse(s.to_world));
    s.c2s = c2s;
    SensorRec &r = s.rec;
    r.sample_to_camera = to_dev(s2c);
    r.to_world = to_dev(s.to_world);
    r.world_to_sample = to_dev(w2s);
    r.camera_pos = transform_pos(r.to_world, f3(0.f));
    r.camera_dir = transform_dir(r.to_world, f3(0.f, 0.f, 1.f));
    const float3 v00 = transform_pos(r.sample_to_camera, f3(0.f, 0.f, 0.f)), v10 = transform_pos(r.sample_to_camera, f3(1.f, 0.f, 0.f)),
                 v11 = transform_pos(r.sample_to_camera, f3(1.f, 1.f, 0.f)), vc = transform_pos(r.sample_to_camera, f3(.5f, .5f, 0.f));
    r.inv_area = 1.f / (norm(v00 - v10) * norm(v11 - v10)) * squared_norm(vc);
    r.width = W; r.height = H;
}

// all wavefront ray launches go through here: counting sort by (origin cell, direction octant) with compaction of inactive lanes,
// then the streaming traversal kernel over the sorted stream (pb_sort.cu, pb_trace2.cuh)
// mode: SORT_CELL_OCTANT for rays that start on surfaces, SORT_DIRECTION for rays that share an origin (camera rays of the edge terms);
// keys_ready: k_shade has left the keys in d_sort_keys
// sorted_inv: sorted-copy mode — `hits` receives the hits in stream order and *sorted_inv the map ray slot -> stream position
static void trace_wavefront(pb_ctx *c, int64_t n, const RayRec *rays, HitRec *hits, cudaEvent_t ev0 = nullptr, cudaEvent_t ev1 = nullptr, int mode = -1,
                            bool keys_ready = false, int lane = 0, const unsigned **sorted_inv = nullptr) {
    if (!c->d_active_total.p) { c->d_active_total.reserve(sizeof(unsigned long long)); cudaMemsetAsync(c->d_active_total.p, 0, sizeof(unsigned long long), c->stream); }
    DevBuf &hist = lane ? c->d_sort_hist1 : c->d_sort_hist, &perm = lane ? c->d_sort_perm1 : c->d_sort_perm, &keys = lane ? c->d_sort_keys1 : c->d_sort_keys,
           &counter = lane ? c->d_stream_counter1 : c->d_stream_counter;
    hist.reserve(40000 * sizeof(unsigned));
    perm.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(unsigned));
    keys.reserve((size_t)std::max<int64_t>(n, 1) * sizeof(unsigned short));
    counter.reserve(sizeof(unsigned));
    RayRec *sorted = nullptr;
    unsigned *inv = nullptr;
    if (sorted_inv) {
        c->d_sorted_rays[lane].reserve((size_t)std::max<int64_t>(n, 1) * sizeof(RayRec));
        c->d_sort_inv[lane].reserve((size_t)std::max<int64_t>(n, 1) * sizeof(unsigned));
        sorted = c->d_sorted_rays[lane].as<RayRec>(); inv = c->d_sort_inv[lane].as<unsigned>();
        *sorted_inv = inv;
    }
    launch_trace_sorted(lane ? c->stream2 : c->stream, c->view, n, rays, hits, f3(c->scene_lo[0], c->scene_lo[1], c->scene_lo[2]),
                        f3(c->scene_hi[0], c->scene_hi[1], c->scene_hi[2]), hist.as<unsigned>(), perm.as<unsigned>(),
                        keys.as<unsigned short>(), counter.as<unsigned>(), c->d_active_total.as<unsigned long long>(), ev0, ev1,
                        mode < 0 ? g_sort_mode : mode, keys_ready, sorted, inv);
    c->launches += 3;
}

// Keep triangle table + BVH + leaf triangles (22 MB for 70 k triangles) resident in L2 while rays, hits and path state stream
// through it: ncu shows the traversal kernel re-reading 3.5 GB of scene data from DRAM per launch otherwise (15 % of its L2
// requests miss). One access-policy window over the scene arena, persisting up to what the device allows.
static void set_l2_window(pb_ctx *c) {
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    if (c->l2_persist && c->d_scene_arena.p) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, c->device) != cudaSuccess || prop.persistingL2CacheMaxSize <= 0) return;
        const size_t want = std::min<size_t>(c->arena_used, (size_t)prop.persistingL2CacheMaxSize);
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        if (cur < want) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
        attr.accessPolicyWindow.base_ptr = c->d_scene_arena.p;
        attr.accessPolicyWindow.num_bytes = std::min<size_t>(c->arena_used, (size_t)prop.accessPolicyMaxWindowSize);
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)std::max<size_t>(1, attr.accessPolicyWindow.num_bytes));
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
    if (c->stream2) cudaStreamSetAttribute(c->stream2, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
}

static cudaEvent_t get_event(pb_ctx *c, size_t idx) {
    while (c->ev_pool.size() <= idx) {
        cudaEvent_t e;
        PB_CUDA(cudaEventCreate(&e));
        c->ev_pool.push_back(e);
    }
    return c->ev_pool[idx];
}

// Primary-edge list of every sensor (perspective.cpp:39-111) and the global secondary-edge table (mesh.cpp:251-264,
// scene.cpp:219-235), built on the device from the resident triangle table and world-space vertices (pb_tables.cu): nothing but
// the two counts, the two sums and the per-mesh kept counts (perspective.cpp:67's assertion) come back to the host.
struct EdgeSrcHost { int v0, v1, f0, f1, v2, mesh; };
static void configure_edges(pb_ctx *c, bool topo_changed) {
    cudaStream_t st = c->stream;
    const bool need_prim = c->sppe > 0, need_sec = c->sppse > 0;
    std::vector<const float *> vw_ptrs(c->meshes.size());
    for (size_t i = 0; i < c->meshes.size(); ++i) vw_ptrs[i] = c->meshes[i].d_vworld.as<float>();
    c->d_mesh_vworld.upload(vw_ptrs, st);
    for (auto &s : c->sensors) { s.num_prim = 0; s.prim_sum = 0.f; }
    c->num_sec = 0; c->sec_sum = 0.f;
    if (!need_prim && !need_sec) return;
    // edge topology of every mesh with enabled edges, in mesh order (uploaded when the topology or the flags change)
    std::vector<int> edge_sig;
    for (const HostMesh &m : c->meshes) { edge_sig.push_back((m.flags & 4) ? (int)m.edges.size() : -1); edge_sig.push_back(m.nf); }
    if (topo_changed || edge_sig != c->edge_sig || !c->d_edge_src.p) {
        std::vector<EdgeSrcHost> src;
        for (size_t mi = 0; mi < c->meshes.size(); ++mi) {
            const HostMesh &m = c->meshes[mi];
            if (!(m.flags & 4)) continue;
            for (size_t e = 0; e < m.edges.size(); e += 5) src.push_back({m.edges[e], m.edges[e + 1], m.edges[e + 2], m.edges[e + 3], m.edges[e + 4], (int)mi});
        }
        c->num_edge_src = (int)src.size();
        c->d_edge_src.upload(src, st);
        PB_CUDA(cudaStreamSynchronize(st));   // src is a local
        c->edge_sig = edge_sig;
    }
    const int E = c->num_edge_src;
    if (E == 0) return;
    const int tiles = (E + 4095) / 4096, nm = (int)c->meshes.size();
    c->d_edge_flags.reserve((size_t)E);
    c->d_edge_local.reserve((size_t)E * sizeof(int));
    c->d_edge_tiles.reserve((size_t)tiles * sizeof(int));
    c->d_edge_out.reserve((size_t)(nm + 4) * sizeof(int));   // [0] count, [1] sum (float), [2..] kept edges per mesh
    int *out = c->d_edge_out.as<int>();
    std::vector<int> h_out(nm + 4);
    const float *const *vw = c->d_mesh_vworld.as<const float *>();
    if (need_prim) {
        for (auto &s : c->sensors) {
            s.d_prim.reserve((size_t)E * sizeof(PrimEdgeRec)); s.d_prim_pmf.reserve((size_t)E * sizeof(float)); s.d_prim_cmf.reserve((size_t)E * sizeof(float));
            launch_primary_edge_table(st, E, c->d_edge_src.p, c->view, vw, s.rec.camera_pos, s.rec.world_to_sample, c->d_edge_flags.as<unsigned char>(),
                                      c->d_edge_local.as<int>(), c->d_edge_tiles.as<int>(), out + 2, nm, s.d_prim.as<PrimEdgeRec>(), s.d_prim_pmf.as<float>(),
                                      s.d_prim_cmf.as<float>(), out, reinterpret_cast<float *>(out + 1));
            c->launches += 6;
            PB_CUDA(cudaMemcpyAsync(h_out.data(), out, (size_t)(nm + 2) * sizeof(int), cudaMemcpyDeviceToHost, st));
            PB_CUDA(cudaStreamSynchronize(st));
            for (int mi = 0; mi < nm; ++mi)
                PB_ASSERT_MSG(!(c->meshes[mi].flags & 4) || c->meshes[mi].edges.empty() || h_out[2 + mi] > 0, "A mesh with enabled edges contributes no primary edge (perspective.cpp:67)");
            s.num_prim = h_out[0];
            std::memcpy(&s.prim_sum, &h_out[1], 4);
        }
    }
    if (need_sec) {
        c->d_sec.reserve((size_t)E * sizeof(SecEdgeRec)); c->d_sec_pmf.reserve((size_t)E * sizeof(float)); c->d_sec_cmf.reserve((size_t)E * sizeof(float));
        launch_secondary_edge_table(st, E, c->d_edge_src.p, c->view, vw, c->d_edge_flags.as<unsigned char>(), c->d_edge_local.as<int>(), c->d_edge_tiles.as<int>(),
                                    c->d_sec.as<SecEdgeRec>(), c->d_sec_pmf.as<float>(), c->d_sec_cmf.as<float>(), out, reinterpret_cast<float *>(out + 1), c->edge_importance);
        c->launches += 5;
        PB_CUDA(cudaMemcpyAsync(h_out.data(), out, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        c->num_sec = h_out[0];
        std::memcpy(&c->sec_sum, &h_out[1], 4);
    }
}

// host copy of the triangle table, fetched when something on the host needs it (BVH build, pb_scene_get_triangle_info)
static void download_triangle_table(pb_ctx *c) {
    if (c->h_tri_valid) return;
    c->h_tri.resize((size_t)c->num_tri * 32);
    if (c->num_tri) PB_CUDA(cudaMemcpyAsync(c->h_tri.data(), c->arena_tri(), (size_t)c->num_tri * sizeof(TriRec), cudaMemcpyDeviceToHost, c->stream));
    PB_CUDA(cudaStreamSynchronize(c->stream));
    c->h_tri_valid = true;
}

static void configure(pb_ctx *c) {
    PB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    PB_ASSERT_MSG(!c->meshes.empty(), "Missing meshes!");
    PB_ASSERT_MSG(!c->sensors.empty(), "Missing sensor!");
    PB_ASSERT_MSG(c->width > 0 && c->height > 0, "Invalid film size!");
    // scene.cpp:65-79 — streams are re-seeded only when the lane count changes
    const int spps[3] = {c->spp, c->sppe, c->sppse};
    for (int k = 0; k < 3; ++k) {
        if (spps[k] <= 0) continue;
        const int64_t count = (int64_t)c->width * c->height * spps[k];
        PB_ASSERT_MSG(count <= std::numeric_limits<int>::max(), "Too many samples (integrator.cpp:73-74)");
        if (c->sampler_count[k] != count) { c->sampler_count[k] = count; c->sampler_offset[k] = 0; }
    }
    // sensors first: the camera positions are part of the scene box the envmap radiates from (scene.cpp:104-119)
    for (auto &s : c->sensors) configure_sensor(s, c->width, c->height);
    // environment lighting: (re)create the bounding mesh as the last mesh (scene.cpp:135-180)
    if (c->has_bound_mesh) { c->meshes.pop_back(); c->has_bound_mesh = false; }
    if (c->emitter_env >= 0) {
        c->meshes.emplace_back();
        HostMesh &b = c->meshes.back();
        static const int face_data[3][12] = {{0, 0, 1, 1, 2, 2, 0, 0, 0, 0, 4, 4}, {1, 3, 5, 7, 3, 7, 5, 4, 2, 6, 7, 6}, {3, 2, 7, 3, 7, 6, 1, 5, 6, 4, 5, 7}};
        b.nv = 8; b.nf = 12; b.flags = 1; b.bsdf = -1; b.emitter = c->emitter_env;
        b.verts.assign(24, 0.f);
        b.faces.resize(36);
        for (int f = 0; f < 12; ++f) for (int j = 0; j < 3; ++j) b.faces[3 * f + j] = face_data[j][f];
        build_edges(b);
        build_csr(b);
        c->emitters[c->emitter_env].mesh = (int)c->meshes.size() - 1;
        c->has_bound_mesh = true;
    }
    // meshes -> triangle table
    int total = 0;
    for (auto &m : c->meshes) { m.face_offset = total; total += m.nf; }
    c->num_tri = total;
    {
        const size_t nt = std::max<size_t>(1, total);
        const size_t tri_b = nt * sizeof(TriRec), node_b = (2 * nt + 1) * sizeof(BvhNode), leaf_b = nt * sizeof(LeafTri), nodec_b = (2 * nt + 1) * sizeof(BvhNodeC);
        if (tri_b != c->arena_tri_bytes || node_b != c->arena_node_bytes || !c->d_scene_arena.p) {   // layout changes: the old tree is gone
            c->bvh_valid = false;
            c->d_scene_arena.reserve(tri_b + node_b + leaf_b + nodec_b);
            c->arena_tri_bytes = tri_b; c->arena_node_bytes = node_b; c->arena_leaf_bytes = leaf_b; c->arena_used = tri_b + node_b + leaf_b + nodec_b;
        }
    }
    bool any_topo = false;
    auto preprocess = [&](size_t i) {
        HostMesh &m = c->meshes[i];
        if (m.topo_dirty) {
            if (!(c->has_bound_mesh && i + 1 == c->meshes.size())) any_topo = true;   // the envmap's box is re-created every time with the same faces
            m.d_faces.upload(m.faces, st);
            m.d_csr_off.upload(m.csr_off, st);
            m.d_csr_face.upload(m.csr_face, st);
            m.d_csr_slot.upload(m.csr_slot, st);
            if (m.flags & 2) { m.d_uvs.upload(m.uvs, st); m.d_uv_faces.upload(m.uv_faces, st); }
            m.topo_dirty = false;
        }
        if (m.verts_dirty) { m.d_vraw.upload(m.verts, st); m.verts_dirty = false; }
        if (m.uv_dirty && (m.flags & 2)) { m.d_uvs.upload(m.uvs, st); m.uv_dirty = false; }   // texture coordinates only: the BVH is untouched
        m.d_vworld.reserve(3 * (size_t)m.nv * sizeof(float));
        m.d_vnormal.reserve(3 * (size_t)m.nv * sizeof(float));
        m.d_fcross.reserve((size_t)m.nf * sizeof(float4));
        m.d_face_area.reserve((size_t)m.nf * sizeof(float));
        m.to_world = matmul(matmul(m.left, m.raw), m.right);   // mesh.cpp:223
        launch_mesh_preprocess(st, m.nv, m.nf, m.face_offset, (int)i, (m.flags & 3) | (m.requires_grad ? 8 : 0), m.d_vraw.as<float>(), to_dev(m.to_world), m.d_faces.as<int>(),
                               m.d_csr_off.as<int>(), m.d_csr_face.as<int>(), m.d_uvs.as<float>(), m.d_uv_faces.as<int>(), m.d_vworld.as<float>(),
                               m.d_fcross.as<float4>(), m.d_vnormal.as<float>(), c->arena_tri(), m.d_face_area.as<float>());
        c->launches += 4;
    };
    const size_t num_regular = c->meshes.size() - (c->has_bound_mesh ? 1 : 0);
    for (size_t i = 0; i < num_regular; ++i) preprocess(i);
    c->h_tri_valid = false;
    const int regular_tris = total - (c->has_bound_mesh ? 12 : 0);
    // bounds of the regular geometry, reduced on the device (6 floats come back)
    c->d_bounds.reserve(6 * sizeof(float));
    auto device_bounds = [&](int ntri, float *lo, float *hi) {
        float lohi[6];
        launch_tri_bounds(st, ntri, c->arena_tri(), c->d_bounds.as<float>());
        c->launches++;
        PB_CUDA(cudaMemcpyAsync(lohi, c->d_bounds.p, sizeof(lohi), cudaMemcpyDeviceToHost, st));
        PB_CUDA(cudaStreamSynchronize(st));
        for (int k = 0; k < 3; ++k) { lo[k] = lohi[k]; hi[k] = lohi[3 + k]; }
    };
    float reg_lo[3], reg_hi[3];
    device_bounds(regular_tris, reg_lo, reg_hi);
    if (c->has_bound_mesh) {
        // scene box: all vertices + camera positions, upper initialised to the smallest positive float (scene.cpp:88-89 quirk),
        // grown by 5 % of its smallest extent (scene.cpp:136-137)
        float lo[3], hi[3];
        for (int k = 0; k < 3; ++k) { lo[k] = std::min(reg_lo[k], std::numeric_limits<float>::max()); hi[k] = std::max(reg_hi[k], std::numeric_limits<float>::min()); }
        for (auto &s : c->sensors) {
            const float p[3] = {s.rec.camera_pos.x, s.rec.camera_pos.y, s.rec.camera_pos.z};
            for (int a = 0; a < 3; ++a) { lo[a] = p[a] < lo[a] ? p[a] : lo[a]; hi[a] = p[a] > hi[a] ? p[a] : hi[a]; }
        }
        float margin = (hi[0] - lo[0]) * 0.05f;
        for (int a = 1; a < 3; ++a) margin = std::min(margin, (hi[a] - lo[a]) * 0.05f);
        for (int a = 0; a < 3; ++a) { lo[a] -= margin; hi[a] += margin; c->env_lower[a] = lo[a]; c->env_upper[a] = hi[a]; }
        HostMesh &b = c->meshes.back();
        for (int i = 0; i < 8; ++i) for (int j = 0; j < 3; ++j) b.verts[3 * i + j] = (i & (1 << j)) ? hi[j] : lo[j];
        b.verts_dirty = true;
        preprocess(c->meshes.size() - 1);
    }
    // face-area pmf / cmf per emitter mesh (mesh.cpp:238-249): sequential fp32 sums as the oracle's, on the device
    {
        int ne = 0;
        for (auto &m : c->meshes) if (m.emitter >= 0) ++ne;
        c->d_area_sums.reserve((size_t)std::max(1, ne) * sizeof(float));
        int k = 0;
        for (auto &m : c->meshes) {
            if (m.emitter < 0) continue;
            m.d_face_cmf.reserve((size_t)std::max(1, m.nf) * sizeof(float));
            launch_seq_cmf(st, m.nf, m.d_face_area.as<float>(), m.d_face_cmf.as<float>(), c->d_area_sums.as<float>() + k++, nullptr);
            c->launches++;
        }
        std::vector<float> sums(std::max(1, ne));
        if (ne) { PB_CUDA(cudaMemcpyAsync(sums.data(), c->d_area_sums.p, (size_t)ne * sizeof(float), cudaMemcpyDeviceToHost, st)); PB_CUDA(cudaStreamSynchronize(st)); }
        k = 0;
        for (auto &m : c->meshes) {
            if (m.emitter < 0) continue;
            m.total_area = sums[k]; m.face_sum = sums[k]; m.inv_total_area = 1.f / sums[k]; ++k;
        }
    }
    // BVH over all triangles (replaces optixAccelBuild, optix.h:277-340)
    {
        if (c->has_bound_mesh) device_bounds(total, c->scene_lo, c->scene_hi);
        else for (int k = 0; k < 3; ++k) { c->scene_lo[k] = reg_lo[k]; c->scene_hi[k] = reg_hi[k]; }
        if (total == 0) for (int k = 0; k < 3; ++k) { c->scene_lo[k] = 0.f; c->scene_hi[k] = 1.f; }
        std::vector<int> sig;
        for (const HostMesh &m : c->meshes) sig.push_back(m.nf);
        const bool can_refit = c->bvh_valid && !any_topo && sig == c->bvh_sig && c->bvh_refits < c->bvh_max_refits && total > 0;
        if (can_refit) {
            float extent = 0.f;
            for (int k = 0; k < 3; ++k) extent = std::max(extent, std::max(std::fabs(c->scene_lo[k]), std::fabs(c->scene_hi[k])));
            launch_build_leaf_tris(st, total, c->d_order.as<int>(), c->arena_tri(), c->arena_leaf());
            c->d_node_boxes.reserve((size_t)c->view.num_nodes * 12 * sizeof(float));
            launch_bvh_refit(st, c->arena_nodes(), c->d_node_boxes.as<float>(), c->arena_leaf(), c->bvh_level_off.data(),
                             (int)c->bvh_level_off.size() - 1, extent);
            c->launches += 1 + (int64_t)c->bvh_level_off.size() - 1;
            c->bvh_refits++; c->bvh_refit_count++;
        } else {
        bool built = false;
        if (c->bvh_builder == 1 && total > 64) {   // device LBVH (pb_lbvh.cu): topology on the device, boxes by the refit kernels
            std::vector<int> lev;
            c->d_order.reserve((size_t)total * sizeof(int));
            if (lbvh_build(st, total, c->arena_tri(), c->scene_lo, c->scene_hi, c->d_lbvh_scratch, c->d_order.as<int>(), c->arena_nodes(), lev)) {
                float extent = 0.f;
                for (int k = 0; k < 3; ++k) extent = std::max(extent, std::max(std::fabs(c->scene_lo[k]), std::fabs(c->scene_hi[k])));
                c->bvh_level_off = lev;
                c->view.num_nodes = lev.back();
                c->bvh_valid = true; c->bvh_sig = sig; c->bvh_refits = 0; c->bvh_builds++;
                launch_build_leaf_tris(st, total, c->d_order.as<int>(), c->arena_tri(), c->arena_leaf());
                c->d_node_boxes.reserve((size_t)c->view.num_nodes * 12 * sizeof(float));
                launch_bvh_refit(st, c->arena_nodes(), c->d_node_boxes.as<float>(), c->arena_leaf(), c->bvh_level_off.data(), (int)c->bvh_level_off.size() - 1, extent);
                c->launches += 4 + 2 * 96 + (int64_t)c->bvh_level_off.size();
                built = true;
            }
        }
        if (!built) {
        download_triangle_table(c);   // the binned-SAH build runs on the host (first build / topology change only)
        std::vector<float> geo(9 * (size_t)total);
        for (int t = 0; t < total; ++t)
            for (int k = 0; k < 3; ++k) for (int a = 0; a < 3; ++a) geo[9 * (size_t)t + 3 * k + a] = c->h_tri[(size_t)t * 32 + 4 * k + a];
        std::vector<HostNode> nodes;
        std::vector<int> order;
        build_bvh(geo.data(), total, nodes, order);
        {   // level ranges of the breadth-first numbering (for the refit); a numbering that is not level-contiguous disables it
            std::vector<int> level(nodes.size(), 0);
            bool ok = true;
            for (size_t i = 0; i < nodes.size(); ++i)
                for (int ch : {nodes[i].left, nodes[i].right})
                    if (ch >= 0) { if (ch <= (int)i) ok = false; else level[ch] = level[i] + 1; }
            for (size_t i = 1; i < nodes.size(); ++i) if (level[i] < level[i - 1]) ok = false;
            c->bvh_level_off.clear();
            if (ok) {
                for (size_t i = 0; i < nodes.size(); ++i) if (i == 0 || level[i] != level[i - 1]) c->bvh_level_off.push_back((int)i);
                c->bvh_level_off.push_back((int)nodes.size());
            }
            c->bvh_valid = ok; c->bvh_sig = sig; c->bvh_refits = 0; c->bvh_builds++;
        }
        std::vector<BvhNode> dn(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) {
            const HostNode &n = nodes[i];
            dn[i].a = make_float4(n.llo[0], n.llo[1], n.llo[2], n.lhi[0]);
            dn[i].b = make_float4(n.lhi[1], n.lhi[2], n.rlo[0], n.rlo[1]);
            dn[i].c = make_float4(n.rlo[2], n.rhi[0], n.rhi[1], n.rhi[2]);
            float l, r;
            std::memcpy(&l, &n.left, 4); std::memcpy(&r, &n.right, 4);
            dn[i].d = make_float4(l, r, 0.f, 0.f);
        }
        PB_ASSERT_MSG(dn.size() * sizeof(BvhNode) <= c->arena_node_bytes, "internal: BVH larger than its arena");
        PB_CUDA(cudaMemcpyAsync(c->arena_nodes(), dn.data(), dn.size() * sizeof(BvhNode), cudaMemcpyHostToDevice, st));
        PB_CUDA(cudaStreamSynchronize(st));   // dn is a local
        c->d_order.upload(order, st);
        if (total == 0) PB_CUDA(cudaMemsetAsync(c->arena_leaf(), 0, sizeof(LeafTri), st));
        else launch_build_leaf_tris(st, total, c->d_order.as<int>(), c->arena_tri(), c->arena_leaf());
        c->launches += 1;
        c->view.num_nodes = (int)dn.size();
        }
        }
    }
    {   // the compact copy of the tree the wavefront traversal reads (pb_trace2.cuh)
        float extent = 0.f;
        for (int k = 0; k < 3; ++k) extent = std::max(extent, std::max(std::fabs(c->scene_lo[k]), std::fabs(c->scene_hi[k])));
        for (const HostSensor &s : c->sensors)   // camera rays go through the compact nodes too: their slab-test slack scales with |o|
            extent = std::max(extent, std::max(std::fabs(s.rec.camera_pos.x), std::max(std::fabs(s.rec.camera_pos.y), std::fabs(s.rec.camera_pos.z))));
        launch_nodes_to_compact(st, c->view.num_nodes, c->arena_nodes(), c->arena_nodes_c(), extent);
        c->launches++;
    }
    // emitters (scene.cpp:183-196, area.cpp:10-17; the envmap keeps its default weight 1, emitter.h:27)
    std::vector<EmitterRec> er(c->emitters.size());
    if (!c->emitters.empty()) {
        std::vector<float> w, cmf;
        float acc = 0.f;
        for (auto &e : c->emitters) {
            const HostMesh &m = c->meshes[e.mesh];
            if (e.type == EMITTER_AREA) e.sampling_weight = m.total_area * (e.radiance[0] * .2126f + e.radiance[1] * .7152f + e.radiance[2] * .0722f);
            else { configure_envmap_distribution(c, e); e.sampling_weight = 1.f; }
            w.push_back(e.sampling_weight);
            acc += e.sampling_weight;
            cmf.push_back(acc);
        }
        c->emitter_sum = acc;
        const float inv = 1.f / acc;
        for (auto &e : c->emitters) e.sampling_weight *= inv;
        c->d_emitter_pmf.upload(w, st);
        c->d_emitter_cmf.upload(cmf, st);
        for (size_t i = 0; i < er.size(); ++i) {
            const HostEmitter &e = c->emitters[i];
            const HostMesh &m = c->meshes[e.mesh];
            EmitterRec &r = er[i];
            std::memset(&r, 0, sizeof(r));
            r.type = e.type; r.mesh = e.mesh; r.sampling_weight = e.sampling_weight;
            r.radiance = f3(e.radiance[0], e.radiance[1], e.radiance[2]);
            r.face_cmf = m.d_face_cmf.as<float>(); r.face_pmf = m.d_face_area.as<float>();
            r.face_sum = m.face_sum; r.num_faces = m.nf; r.face_offset = m.face_offset;
            if (e.type == EMITTER_ENVMAP) {
                r.env_radiance.data = e.env_radiance.d.as<float>(); r.env_radiance.grad = nullptr;
                r.env_radiance.w = e.env_radiance.w; r.env_radiance.h = e.env_radiance.h; r.env_radiance.c = 3;
                r.env_scale = e.env_scale; r.env_res_x = e.env_res[0]; r.env_res_y = e.env_res[1]; r.env_cells = e.env_res[0] * e.env_res[1];
                const Mat4h tw = matmul(e.env_left, e.env_raw);   // envmap.cpp:23
                r.env_to_world = to_dev(tw); r.env_from_world = to_dev(inverse(tw));
                r.env_lower = f3(c->env_lower[0], c->env_lower[1], c->env_lower[2]); r.env_upper = f3(c->env_upper[0], c->env_upper[1], c->env_upper[2]);
                r.env_sum = e.env_sum; r.env_cmf = e.d_env_cmf.as<float>(); r.env_pmf = e.d_env_pmf.as<float>();
            }
        }
    }
    c->d_emitters.upload(er, st);
    c->h_emitters = er;
    // mesh + bsdf tables
    std::vector<MeshRec> mr(c->meshes.size());
    for (size_t i = 0; i < mr.size(); ++i) {
        const HostMesh &m = c->meshes[i];
        std::memset(&mr[i], 0, sizeof(MeshRec));
        mr[i].bsdf = m.bsdf; mr[i].emitter = m.emitter; mr[i].inv_total_area = m.inv_total_area;
        mr[i].face_offset = m.face_offset; mr[i].num_faces = m.nf; mr[i].flags = (m.flags & 3) | (m.requires_grad ? 4 : 0);
        mr[i].uv_faces = (m.flags & 2) ? m.d_uv_faces.as<int>() : nullptr; mr[i].uv_grad = nullptr;
    }
    c->d_meshes.upload(mr, st);
    c->h_meshrecs = mr;
    std::vector<BsdfRec> br(c->bsdfs.size());
    for (size_t i = 0; i < br.size(); ++i) {
        HostBsdf &b = c->bsdfs[i];
        std::memset(&br[i], 0, sizeof(BsdfRec));
        br[i].type = b.type;
        for (int k = 0; k < TEX_COUNT; ++k) {
            HostTexture &t = b.tex[k];
            if (t.dirty) { t.d.upload(t.data, st); t.dirty = false; }
            br[i].tex[k].data = t.d.as<float>(); br[i].tex[k].grad = nullptr;
            br[i].tex[k].w = t.w; br[i].tex[k].h = t.h; br[i].tex[k].c = t.c;
        }
    }
    c->d_bsdfs.upload(br, st);
    c->h_bsdfs = br;
    PB_CUDA(cudaStreamSynchronize(st));
    SceneView &V = c->view;
    V.tri = c->arena_tri(); V.leaf = c->arena_leaf(); V.nodes = c->arena_nodes(); V.nodes_c = c->arena_nodes_c();
    V.meshes = c->d_meshes.as<MeshRec>(); V.bsdfs = c->d_bsdfs.as<BsdfRec>(); V.emitters = c->d_emitters.as<EmitterRec>();
    V.emitter_cmf = c->d_emitter_cmf.as<float>(); V.emitter_pmf = c->d_emitter_pmf.as<float>(); V.emitter_sum = c->emitter_sum;
    V.num_tri = total; V.num_meshes = (int)mr.size(); V.num_bsdfs = (int)br.size(); V.num_emitters = (int)er.size();
    V.emitter_env = c->emitter_env;
    V.simple = (c->emitter_env < 0) ? 1 : 0;
    for (const HostBsdf &hb : c->bsdfs) if (hb.type != PB_BSDF_DIFFUSE) V.simple = 0;
    V.tri_grad = nullptr;
    V.tri_tangent = nullptr; V.jvp_acc = nullptr; V.jvp_image = nullptr; V.jvp_channel = 0; V.sensor_grad = nullptr;
    configure_edges(c, any_topo);
    // gradient layout
    c->grad_segments.clear();
    int64_t off = 0;
    for (size_t i = 0; i < c->bsdfs.size(); ++i)
        for (int k = 0; k < TEX_COUNT; ++k)
            if (c->bsdfs[i].tex[k].requires_grad) {
                const int64_t n = (int64_t)c->bsdfs[i].tex[k].data.size();
                c->grad_segments.push_back({PB_PARAM_BSDF_TEXTURE, (int)i, k, off, n});
                off += n;
            }
    for (size_t i = 0; i < c->meshes.size(); ++i)
        if (c->meshes[i].requires_grad) {
            const int64_t n = 3 * (int64_t)c->meshes[i].nv;
            c->grad_segments.push_back({PB_PARAM_MESH_VERTICES, (int)i, 0, off, n});
            off += n;
        }
    for (size_t i = 0; i < c->meshes.size(); ++i)
        if (c->meshes[i].uv_requires_grad && (c->meshes[i].flags & 2)) {
            const int64_t n = (int64_t)c->meshes[i].uvs.size();
            c->grad_segments.push_back({PB_PARAM_MESH_UV, (int)i, 0, off, n});
            off += n;
        }
    for (size_t i = 0; i < c->sensors.size(); ++i)
        if (c->sensors[i].requires_grad) { c->grad_segments.push_back({PB_PARAM_SENSOR_TRANSFORM, (int)i, 0, off, 16}); off += 16; }
    if (c->emitter_env >= 0) {
        const HostEmitter &e = c->emitters[c->emitter_env];
        if (e.env_radiance.requires_grad) {
            const int64_t n = (int64_t)e.env_radiance.data.size();
            c->grad_segments.push_back({PB_PARAM_ENVMAP_RADIANCE, c->emitter_env, 0, off, n});
            off += n;
        }
        if (e.env_scale_requires_grad) { c->grad_segments.push_back({PB_PARAM_ENVMAP_SCALE, c->emitter_env, 0, off, 1}); off += 1; }
        if (e.env_xf_requires_grad) { c->grad_segments.push_back({PB_PARAM_ENVMAP_TRANSFORM, c->emitter_env, 0, off, 16}); off += 16; }
    }
    set_l2_window(c);
    c->ready = true;
    c->d_generation_at_configure = c->d_generation;
    c->have_last_d = false;
    c->retained_valid = false;
}

// seed table of the stateless sampler: (state, inc) of every lane id any of the three samplers can use, built once per lane count
static void attach_rng_seeds(pb_ctx *c, RenderParams &P) {
    const int64_t npix = (int64_t)c->width * c->height;
    const int64_t need = npix * std::max(c->spp, std::max(c->sppe, c->sppse));
    P.rng_seed = nullptr; P.rng_seed_count = 0;
    if (need <= 0 || !c->rng_seed_table) return;
    if (c->rng_seed_count < need) {
        try { c->d_rng_seed.reserve((size_t)need * sizeof(ulonglong2)); }
        catch (const Error &) { cudaGetLastError(); c->rng_seed_count = 0; return; }   // no room: the kernels hash on the fly
        launch_rng_seed(c->stream, need, c->d_rng_seed.as<ulonglong2>());
        c->launches++;
        c->rng_seed_count = need;
    }
    P.rng_seed = c->d_rng_seed.as<ulonglong2>(); P.rng_seed_count = c->rng_seed_count;
}

// the ray sort's box and mode for k_shade, which leaves every emitted ray's key next to it (the sort then skips reading the rays)
static void set_sort_params(const pb_ctx *c, BounceParams &B) {
    B.sort_lo = f3(c->scene_lo[0], c->scene_lo[1], c->scene_lo[2]);
    B.sort_inv_ext = f3(1.f / fmaxf(c->scene_hi[0] - c->scene_lo[0], 1e-20f), 1.f / fmaxf(c->scene_hi[1] - c->scene_lo[1], 1e-20f),
                        1.f / fmaxf(c->scene_hi[2] - c->scene_lo[2], 1e-20f));
    B.sort_mode = g_sort_mode;
}

struct Plan { int nbounce, nb, nl, draws; };
static Plan make_plan(const pb_integrator &I) {
    Plan p;
    if (I.kind == PB_INTEG_DIRECT) {
        PB_ASSERT_MSG(I.bsdf_samples >= 0 && I.light_samples >= 0 && I.bsdf_samples + I.light_samples > 0, "Invalid DirectIntegrator sample counts");
        p.nbounce = 1; p.nb = I.bsdf_samples; p.nl = I.light_samples;
    } else if (I.kind == PB_INTEG_PATH) {
        PB_ASSERT_MSG(I.max_depth >= 1, "PathIntegrator needs max_depth >= 1");
        p.nbounce = I.max_depth; p.nb = 1; p.nl = 1;
    } else {
        PB_ASSERT_MSG(I.kind == PB_INTEG_FIELD, "Unknown integrator kind");
        p.nbounce = 0; p.nb = p.nl = 0;
    }
    p.draws = 2 + p.nbounce * (3 * p.nb + 2 * p.nl);
    return p;
}

enum Mode { MODE_C = 0, MODE_D = 1, MODE_VJP = 2, MODE_JVP = 3 };

struct Plan;
static void run_edge_terms(pb_ctx *c, const pb_integrator &I, int sensor, const Plan &plan, const RenderParams &Pbase, const float *d_dLdI, size_t *nev);

static EdgeParams make_edge_params(pb_ctx *c, int sensor, const GuideGrid *guide) {
    EdgeParams Q;
    std::memset(&Q, 0, sizeof(Q));
    const HostSensor &s = c->sensors[sensor];
    Q.prim = s.d_prim.as<PrimEdgeRec>(); Q.prim_cmf = s.d_prim_cmf.as<float>(); Q.prim_pmf = s.d_prim_pmf.as<float>();
    Q.prim_sum = s.prim_sum; Q.num_prim = s.num_prim;
    Q.sec = c->d_sec.as<SecEdgeRec>(); Q.sec_cmf = c->d_sec_cmf.as<float>(); Q.sec_pmf = c->d_sec_pmf.as<float>();
    Q.sec_sum = c->sec_sum; Q.num_sec = c->num_sec;
    if (guide && guide->ready) {
        Q.guide_cmf = guide->d_cmf.as<float>(); Q.guide_pmf = guide->d_pmf.as<float>(); Q.guide_sum = guide->sum; Q.guide_cells = guide->cells;
    }
    if (guide) for (int k = 0; k < 3; ++k) Q.guide_res[k] = guide->res[k];
    Q.mesh_gworld = c->d_mesh_gworld.as<float *>();
    Q.mesh_vworld = c->d_mesh_vworld.as<const float *>();
    return Q;
}

static void size_store(EventStore &S, int64_t lanes, int nslots, int R, int64_t ray_lanes, bool keep_lin = false) {
    if ((int)S.pos.size() < nslots) { S.pos.resize(nslots); S.vb.resize(nslots); S.vc.resize(nslots); S.hits.resize(nslots); }
    if (keep_lin) {
        if ((int)S.lin.size() < nslots) S.lin.resize(nslots);
        for (int k = 0; k < nslots; ++k) S.lin[k].reserve((size_t)lanes * sizeof(float4));
    }
    if ((int)S.thr.size() < nslots + 1) S.thr.resize(nslots + 1);
    S.hit0.reserve((size_t)lanes * sizeof(HitRec));
    S.rad.reserve((size_t)lanes * sizeof(float4));
    S.rays.reserve((size_t)ray_lanes * R * sizeof(RayRec));
    for (int k = 0; k < nslots; ++k) {
        S.pos[k].reserve((size_t)lanes * sizeof(float4)); S.vb[k].reserve((size_t)lanes * sizeof(float4)); S.vc[k].reserve((size_t)lanes * sizeof(float4));
        S.hits[k].reserve((size_t)lanes * R * sizeof(HitRec));
    }
    for (int k = 0; k < nslots + 1; ++k) S.thr[k].reserve((size_t)lanes * sizeof(float4));
}

static bool any_geom_jvp(const pb_ctx *c) {   // anything that differentiates geometry (vertices or the sensor pose)
    for (const GradSegment &g : c->grad_segments) if (g.kind == PB_PARAM_MESH_VERTICES || g.kind == PB_PARAM_SENSOR_TRANSFORM) return true;
    return false;
}
// adjoint of to_world from the adjoint of world_to_sample = camera_to_sample * inverse(to_world):  g_M = -Minv^T (C^T g_W) Minv^T
static void fold_w2s_adjoint(const HostSensor &s, const float *g_w2s, float *g_m_accum) {
    const Mat4h Minv = inverse(s.to_world);
    double CtG[16], tmp[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += (double)s.c2s.m[4 * k + i] * g_w2s[4 * k + j]; CtG[4 * i + j] = a; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += (double)Minv.m[4 * k + i] * CtG[4 * k + j]; tmp[4 * i + j] = a; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += tmp[4 * i + k] * (double)Minv.m[4 * j + k]; g_m_accum[4 * i + j] -= (float)a; }
}
// tangent of world_to_sample from the tangent of to_world:  W_t = -C Minv M_t Minv
static void w2s_tangent(const HostSensor &s, const float *m_t, float *w_t) {
    const Mat4h Minv = inverse(s.to_world);
    double a1[16], a2[16];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += (double)Minv.m[4 * i + k] * m_t[4 * k + j]; a1[4 * i + j] = a; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += a1[4 * i + k] * (double)Minv.m[4 * k + j]; a2[4 * i + j] = a; }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += (double)s.c2s.m[4 * i + k] * a2[4 * k + j]; w_t[4 * i + j] = -(float)a; }
}

// the three ray launches + evaluation of the secondary-edge estimator for one batch (direct.cpp:225-316)
static void secondary_edge_batch(pb_ctx *c, const RenderParams &P, const EdgeParams &Q, const float *d_dLdI, float inv_sppse, float *guide_out, int guide_spc) {
    cudaStream_t st = c->stream;
    EventStore &S = c->scratch;
    RayRec *rays = S.rays.as<RayRec>();
    HitRec *hits = S.hits[0].as<HitRec>();
    RayRec *cam_rays = c->d_edge_rays.as<RayRec>();
    HitRec *cam_hits = S.hit0.as<HitRec>();
    launch_edge_secondary_rays(st, P, Q, rays, guide_spc);
    trace_wavefront(c, 2 * (int64_t)P.n, rays, hits);
    launch_edge_secondary_camera(st, P, Q, rays, hits, cam_rays, guide_spc);
    trace_wavefront(c, (int64_t)P.n, cam_rays, cam_hits, nullptr, nullptr, 7 /* SORT_DIRECTION: every ray starts at the camera */);
    if (P.S.tri_tangent) {
        RenderParams Pc = P;
        for (int ch = 0; ch < 3; ++ch) {
            Pc.S.jvp_channel = ch;
            launch_edge_secondary_eval(st, Pc, Q, rays, hits, cam_rays, cam_hits, d_dLdI, inv_sppse, guide_out, guide_spc);
            c->launches++;
        }
    } else {
        launch_edge_secondary_eval(st, P, Q, rays, hits, cam_rays, cam_hits, d_dLdI, inv_sppse, guide_out, guide_spc);
        c->launches++;
    }
    c->launches += 4;
}

// boundary terms of renderD in reverse mode (integrator.cpp:98-119, direct.cpp:207-221); lanes are sharded by index range
static void run_edge_terms(pb_ctx *c, const pb_integrator &I, int sensor, const Plan &plan, const RenderParams &Pbase, const float *d_dLdI, size_t *) {
    cudaStream_t st = c->stream;
    const int64_t npix = (int64_t)c->width * c->height;
    const bool field = (I.kind == PB_INTEG_FIELD);
    // per-mesh world-space vertex adjoint buffers
    std::vector<float *> gw(c->meshes.size(), nullptr);
    for (const GradSegment &g : c->grad_segments)
        if (g.kind == PB_PARAM_MESH_VERTICES) gw[g.id] = c->meshes[g.id].d_gworld.as<float>();
    c->d_mesh_gworld.upload(gw, st);
    const GuideGrid *guide = (I.use_guiding && sensor < (int)c->guides.size()) ? &c->guides[sensor] : nullptr;
    const EdgeParams Q = make_edge_params(c, sensor, guide);
    const int R = std::max(2, plan.nb + plan.nl);
    const int64_t max_lanes = npix * std::max(c->sppe, c->sppse);
    const int64_t B = std::max<int64_t>(1024, std::min<int64_t>(c->batch, ((max_lanes + 1023) / 1024) * 1024));
    size_store(c->scratch, B, 2, R, B);
    c->d_edge_rays.reserve((size_t)B * sizeof(RayRec));
    c->d_edge_rad.reserve((size_t)B * sizeof(float4));
    RenderParams P = Pbase;
    if (P.S.tri_tangent) { c->d_jvp_acc.reserve((size_t)B * sizeof(float)); P.S.jvp_acc = c->d_jvp_acc.as<float>(); }
    P.spp = 1; P.spp_local = 1; P.s0 = 0; P.inv_spp = 1.f; P.tile_rows = 0;   // edge lanes are sharded by index range
    EventStore &S = c->scratch;
    // ---- primary edges
    if (c->sppe > 0 && Q.num_prim > 0) {
        const int64_t N = npix * c->sppe;
        const int64_t l0 = N * c->rank / c->world, l1 = N * (c->rank + 1) / c->world;
        const uint64_t base = c->last_d_offset_e;
        const uint64_t per_event = 3 * plan.nb + 2 * plan.nl, per_li = (uint64_t)plan.nbounce * per_event;
        P.jump0 = make_jump(base);
        for (int64_t start = l0; start < l1; start += B) {
            P.local0 = start; P.n = (int)std::min<int64_t>(B, l1 - start);
            for (int side = 0; side < 2; ++side) {   // side 0 = ray_p is evaluated first (operand order)
                HitRec *hit0 = S.hit0.as<HitRec>();
                launch_edge_primary_rays(st, P, Q, side, S.rays.as<RayRec>());
                trace_wavefront(c, (int64_t)P.n, S.rays.as<RayRec>(), hit0, nullptr, nullptr, 7 /* SORT_DIRECTION */);
                c->launches++;
                if (field) {
                    launch_field(st, P, I.field, hit0, nullptr, S.rad.as<float4>());
                    c->launches++;
                } else {
                    for (int k = 0; k < plan.nbounce; ++k) {
                        BounceParams Bp;
                        Bp.nb = plan.nb; Bp.nl = plan.nl; Bp.depth = k; Bp.last = (k == plan.nbounce - 1); Bp.carry = (I.kind == PB_INTEG_PATH);
                        Bp.hide_emitters = I.hide_emitters; Bp.ad = 0; Bp.rc_grad = 0;
                        Bp.jump = make_jump(base + 1 + (uint64_t)side * per_li + (uint64_t)k * per_event);
                        set_sort_params(c, Bp);
                        EventBuffers E;
                        E.hit_cur = (k == 0) ? hit0 : S.hits[(k - 1) & 1].as<HitRec>();
                        E.hit_prev = nullptr;
                        E.prev_pos = (k == 0) ? nullptr : S.pos[(k - 1) & 1].as<float4>();
                        E.pos = S.pos[k & 1].as<float4>(); E.vb = S.vb[k & 1].as<float4>(); E.vc = S.vc[k & 1].as<float4>();
                        E.rays = S.rays.as<RayRec>();
                        E.hits = S.hits[k & 1].as<HitRec>();
                        E.thr_in = (k == 0) ? nullptr : S.thr[k & 1].as<float4>();
                        E.thr_out = Bp.last ? nullptr : S.thr[(k + 1) & 1].as<float4>();
                        E.rad = S.rad.as<float4>();
                        c->d_sort_keys.reserve((size_t)P.n * (plan.nb + plan.nl) * sizeof(unsigned short));
                        E.keys = c->d_sort_keys.as<unsigned short>();
                        E.lin = nullptr; E.inv = nullptr; E.inv_cur = nullptr; E.lane_list = nullptr; E.lane_count = nullptr;
                        c->d_conn[0].reserve((size_t)P.n * (plan.nb + plan.nl) * sizeof(float2));
                        E.conn = c->d_conn[0].as<float2>();
                        launch_shade(st, P, Bp, E);
                        trace_wavefront(c, (int64_t)P.n * (plan.nb + plan.nl), E.rays, E.hits, nullptr, nullptr, -1, true);
                        launch_resolve(st, P, Bp, E, nullptr);
                        c->launches += 3;
                    }
                }
                if (side == 0) PB_CUDA(cudaMemcpyAsync(c->d_edge_rad.p, S.rad.p, (size_t)P.n * sizeof(float4), cudaMemcpyDeviceToDevice, st));
            }
            if (P.S.tri_tangent) {
                RenderParams Pc = P;
                for (int ch = 0; ch < 3; ++ch) {
                    Pc.S.jvp_channel = ch;
                    launch_edge_primary_grad(st, Pc, Q, c->d_edge_rad.as<float4>(), S.rad.as<float4>(), d_dLdI, 1.f / (float)c->sppe);
                    c->launches++;
                }
            } else {
                launch_edge_primary_grad(st, P, Q, c->d_edge_rad.as<float4>(), S.rad.as<float4>(), d_dLdI, 1.f / (float)c->sppe);
                c->launches++;
            }
        }
    }
    // ---- secondary edges (the base Integrator / FieldExtractionIntegrator has none: integrator.h:24)
    if (c->sppse > 0 && Q.num_sec > 0 && !field) {
        const int64_t N = npix * c->sppse;
        const int64_t l0 = N * c->rank / c->world, l1 = N * (c->rank + 1) / c->world;
        P.jump0 = make_jump(c->last_d_offset_s);
        for (int64_t start = l0; start < l1; start += B) {
            P.local0 = start; P.n = (int)std::min<int64_t>(B, l1 - start);
            secondary_edge_batch(c, P, Q, d_dLdI, 1.f / (float)c->sppse, nullptr, 0);
        }
    }
}

// DirectIntegrator::preprocess_secondary_edges (direct.cpp:166-204): guiding grid over the 3-D secondary-edge sample space
static void preprocess_secondary_edges(pb_ctx *c, int sensor, const int *reso, int nrounds) {
    PB_ASSERT_MSG(nrounds > 0, "nrounds must be positive");
    PB_ASSERT_MSG(c->ready, "Scene needs to be configured!");
    PB_ASSERT_MSG(sensor >= 0 && sensor < (int)c->sensors.size(), "Invalid sensor id!");
    PB_ASSERT_MSG(c->sppse > 0 && c->num_sec > 0, "preprocess_secondary_edges needs sppse > 0 and at least one secondary edge");
    PB_ASSERT_MSG(reso[0] > 0 && reso[1] > 0 && reso[2] > 0 && reso[3] > 0, "Invalid guiding resolution");
    PB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if ((int)c->guides.size() < (int)c->sensors.size()) c->guides.resize(c->sensors.size());
    GuideGrid &G = c->guides[sensor];
    const int64_t cells = (int64_t)reso[0] * reso[1] * reso[2];
    const int64_t N = cells * reso[3];
    PB_ASSERT_MSG(N <= std::numeric_limits<int>::max(), "Too many guiding samples");
    for (int k = 0; k < 3; ++k) G.res[k] = reso[k];
    G.cells = (int)cells; G.ready = false;
    G.d_pmf.reserve((size_t)cells * sizeof(float));
    PB_CUDA(cudaMemsetAsync(G.d_pmf.p, 0, (size_t)cells * sizeof(float), st));
    std::vector<float *> gw(c->meshes.size(), nullptr);
    c->d_mesh_gworld.upload(gw, st);
    const EdgeParams Q = make_edge_params(c, sensor, &G);
    const int64_t B = c->batch;
    size_store(c->scratch, B, 2, 2, B);
    c->d_edge_rays.reserve((size_t)B * sizeof(RayRec));
    RenderParams P;
    std::memset(&P, 0, sizeof(P));
    P.S = c->view; P.cam = c->sensors[sensor].rec;
    P.width = c->width; P.height = c->height; P.spp = 1; P.spp_local = 1; P.s0 = 0; P.inv_spp = 1.f;
    attach_rng_seeds(c, P);
    for (int j = 0; j < nrounds; ++j) {
        P.jump0 = make_jump(3 * (uint64_t)j);   // a private sampler seeded with arange(N) (direct.cpp:185-186)
        for (int64_t start = 0; start < N; start += B) {
            P.local0 = start; P.n = (int)std::min<int64_t>(B, N - start);
            secondary_edge_batch(c, P, Q, nullptr, 1.f, G.d_pmf.as<float>(), reso[3]);
        }
    }
    std::vector<float> pmf((size_t)cells), cmf((size_t)cells);
    PB_CUDA(cudaMemcpyAsync(pmf.data(), G.d_pmf.p, (size_t)cells * sizeof(float), cudaMemcpyDeviceToHost, st));
    PB_CUDA(cudaStreamSynchronize(st));
    float acc = 0.f;
    for (int64_t i = 0; i < cells; ++i) { if (nrounds > 1) pmf[i] /= (float)nrounds; acc += pmf[i]; cmf[i] = acc; }
    G.sum = acc;
    G.d_pmf.upload(pmf, st); G.d_cmf.upload(cmf, st);
    PB_CUDA(cudaStreamSynchronize(st));
    G.ready = acc > 0.f;
}

// interior term: integrator.cpp:64-95 over this shard's samples, in batches.
//   MODE_C / MODE_D   forward render (renderC's / renderD's formulation of the primal). MODE_D additionally retains every
//                     event's hit records, vertex positions and throughputs for the whole shard when they fit
//                     pb_ctx_set_retain_limit (592 B per lane for a depth-5 path), so that the VJP needs no re-tracing.
//   MODE_VJP          adjoint kernels in reverse event order over the retained records; if nothing was retained the
//                     forward pass is replayed batch by batch first (same stream positions as the last renderD).
static void render_interior(pb_ctx *c, const pb_integrator &I, int sensor, float *d_image, Mode mode, const float *d_dLdI = nullptr,
                            float *d_grad = nullptr) {
    // MODE_JVP: d_grad is the (read-only) flat tangent vector and d_image the W*H*3 derivative image; everything else follows MODE_VJP
    const bool jvp = (mode == MODE_JVP);
    float *d_dimage = jvp ? d_image : nullptr;
    if (jvp) mode = MODE_VJP;
    PB_ASSERT_MSG(c->ready, "Input scene must be configured!");
    PB_ASSERT_MSG(sensor >= 0 && sensor < (int)c->sensors.size(), "Invalid sensor id!");
    PB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int64_t npix = (int64_t)c->width * c->height;
    if (d_image) PB_CUDA(cudaMemsetAsync(d_image, 0, (size_t)npix * 3 * sizeof(float), st));
    if (jvp) d_image = nullptr;
    c->last_trace_ms = 0.f; c->last_rays = 0; c->last_primary_ms = 0.f; c->last_trace_launches = 0; c->last_active_rays = 0;
    if (!c->d_active_total.p) c->d_active_total.reserve(sizeof(unsigned long long));
    PB_CUDA(cudaMemsetAsync(c->d_active_total.p, 0, sizeof(unsigned long long), c->stream));
    const bool by_pixels = (c->shard_mode == 1 && c->world > 1);
    const int s0 = by_pixels ? 0 : (int)((int64_t)c->spp * c->rank / c->world), s1 = by_pixels ? c->spp : (int)((int64_t)c->spp * (c->rank + 1) / c->world);
    const int spp_local = s1 - s0;
    // pixel sharding: rows of the tiles t * world + rank
    const int tile_rows = by_pixels ? std::max(1, c->tile_rows > 0 ? c->tile_rows : (c->height + c->world - 1) / c->world) : 0;
    int64_t local_rows = c->height;
    if (by_pixels) {
        local_rows = 0;
        for (int64_t r0 = (int64_t)c->rank * tile_rows; r0 < c->height; r0 += (int64_t)tile_rows * c->world) local_rows += std::min<int64_t>(tile_rows, c->height - r0);
    }
    const Plan plan = make_plan(I);
    const uint64_t base = (mode == MODE_VJP) ? c->last_d_offset : c->sampler_offset[0];
    if (mode == MODE_D) c->retained_valid = false;
    const bool field = (I.kind == PB_INTEG_FIELD);
    if ((c->spp <= 0 || spp_local <= 0) && mode != MODE_VJP) {
        PB_CUDA(cudaStreamSynchronize(st));
        if (c->spp > 0) c->sampler_offset[0] = base + plan.draws;
        if (mode == MODE_D) {   // the edge samplers still advance (integrator.cpp:41-47)
            const uint64_t per_li = (uint64_t)plan.nbounce * (3 * plan.nb + 2 * plan.nl);
            c->last_d_offset_e = c->sampler_offset[1]; c->last_d_offset_s = c->sampler_offset[2];
            if (c->sppe > 0 && c->sensors[sensor].num_prim > 0) c->sampler_offset[1] += 1 + 2 * per_li;
            if (c->sppse > 0 && !field) c->sampler_offset[2] += 3;
        }
        return;
    }
    PB_ASSERT_MSG(field || !c->emitters.empty(), "No Emitter!");
    const int64_t total = (c->spp > 0 && spp_local > 0) ? local_rows * c->width * spp_local : 0;
    const int R = std::max(1, plan.nb + plan.nl);
    // two batches in flight on two streams when the render is short (see pb_host.h); forward mode keeps one (its per-lane accumulator)
    const bool dual = !field && !jvp && total > 2048 && (c->pipeline == 2 || (c->pipeline == 1 && total <= c->pipeline_max_lanes));
    int64_t B = std::max<int64_t>(1024, std::min<int64_t>(c->batch, ((total + 1023) / 1024) * 1024));
    if (dual && total <= B) B = std::max<int64_t>(1024, (((total + 1) / 2 + 1023) / 1024) * 1024);   // a single batch: split it in two
    if (mode == MODE_VJP && c->retained_valid && c->retained_B > 0) B = c->retained_B;   // the retained records are laid out batch by batch ([ray][lane] inside a batch)
    const int D = std::max(1, plan.nbounce);
    const bool lin = c->view.simple != 0 && g_shade_simple != 0;   // diffuse BSDFs + area emitters only: the events' reflectance linearisation is kept for k_adjoint_lin
    const int64_t retain_bytes = total * (16 + 16 + (int64_t)D * (48 + 16 * R + 16 + (lin ? 16 : 0)));
    // every leaf is a reflectance texture of a diffuse scene, reverse mode: k_adjoint_lin runs and no kernel reads an event's hit records after its k_resolve
    bool lin_only = lin && !jvp && g_adjoint_lin != 0;
    for (const GradSegment &g : c->grad_segments) if (g.kind != PB_PARAM_BSDF_TEXTURE) lin_only = false;
    const bool sorted_copy = c->sorted_copy != 0 && !field;
    // which store, and whether the forward pass has to run
    bool use_retained = false, run_forward = true;
    if (mode == MODE_D && !field && !c->grad_segments.empty() && retain_bytes <= c->retain_limit) use_retained = true;
    if (mode == MODE_VJP) {
        if (c->retained_valid && c->retained_kind == I.kind && c->retained_nb == plan.nb && c->retained_nl == plan.nl &&
            c->retained_nbounce == plan.nbounce && c->retained_sensor == sensor && c->retained_hide == I.hide_emitters &&
            (c->retained_hits_by_slot || lin_only)) {
            use_retained = true; run_forward = false;
        }
    }
    if (use_retained && run_forward) {   // the store may not fit next to what else lives on the device: fall back to re-tracing in the VJP
        try { size_store(c->retained, total, D, R, B, lin); }
        catch (const Error &) { cudaGetLastError(); c->retained.release(); use_retained = false; }
    }
    const bool keep = use_retained || mode == MODE_VJP;   // every event has its own slot
    const bool hits_by_slot = keep && !lin_only;          // sorted-copy traversal: hits are un-permuted for the kernels that index them by ray slot
    EventStore &S = use_retained ? c->retained : c->scratch;
    size_store(S, use_retained ? total : B, keep ? D : 2, R, B, keep && lin);
    if (mode == MODE_VJP) c->d_suffix.reserve((size_t)B * sizeof(float4));
    c->d_sort_keys.reserve((size_t)B * R * sizeof(unsigned short));   // k_shade writes the sort keys of the rays it emits
    c->d_conn[0].reserve((size_t)B * R * sizeof(float2));
    if (dual) {   // lane 1: its own stream and per-batch buffers
        if (!c->stream2) {
            PB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
            PB_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
            PB_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
            set_l2_window(c);
        }
        if (!use_retained) size_store(c->scratch1, B, keep ? D : 2, R, B, keep && lin);
        c->d_rays1.reserve((size_t)B * R * sizeof(RayRec));
        if (mode == MODE_VJP) c->d_suffix1.reserve((size_t)B * sizeof(float4));
        c->d_sort_keys1.reserve((size_t)B * R * sizeof(unsigned short));
        c->d_conn[1].reserve((size_t)B * R * sizeof(float2));
    }
    RenderParams P;
    P.S = c->view; P.cam = c->sensors[sensor].rec;
    P.width = c->width; P.height = c->height; P.spp = c->spp; P.inv_spp = 1.f / (float)c->spp;
    P.spp_local = spp_local; P.s0 = s0;
    P.tile_rows = tile_rows; P.rank = c->rank; P.world = c->world;
    attach_rng_seeds(c, P);
    P.jump0 = make_jump(base);
    if (mode == MODE_VJP) {   // BSDF table whose textures point at their gradient segments
        // the tables of configure with their gradient pointers filled in: patched on the host copies, uploaded asynchronously (no read-back)
        std::vector<BsdfRec> &br = c->h_bsdfs_grad;
        br = c->h_bsdfs;
        for (const GradSegment &g : c->grad_segments)
            if (g.kind == PB_PARAM_BSDF_TEXTURE) br[g.id].tex[g.slot].grad = d_grad + g.offset;   // forward mode: the tangent, read only
        c->d_bsdfs_grad.upload(br, st);
        P.S.bsdfs = c->d_bsdfs_grad.as<BsdfRec>();
        {   // mesh table whose uv-gradient pointers point at their segments (Mesh.vertex_uv leaves)
            bool any_uv = false;
            for (const GradSegment &g : c->grad_segments) any_uv = any_uv || g.kind == PB_PARAM_MESH_UV;
            if (any_uv) {
                std::vector<MeshRec> &mr = c->h_meshrecs_grad;
                mr = c->h_meshrecs;
                for (const GradSegment &g : c->grad_segments) if (g.kind == PB_PARAM_MESH_UV) mr[g.id].uv_grad = d_grad + g.offset;
                c->d_meshes_grad.upload(mr, st);
                P.S.meshes = c->d_meshes_grad.as<MeshRec>();
            }
        }
        bool env_grad = false;
        for (const GradSegment &g : c->grad_segments)
            env_grad = env_grad || g.kind == PB_PARAM_ENVMAP_RADIANCE || g.kind == PB_PARAM_ENVMAP_SCALE || g.kind == PB_PARAM_ENVMAP_TRANSFORM;
        if (env_grad) {   // emitter table whose environment map points at its gradient segments
            std::vector<EmitterRec> &er = c->h_emitters_grad;
            er = c->h_emitters;
            for (const GradSegment &g : c->grad_segments) {
                if (g.kind == PB_PARAM_ENVMAP_RADIANCE) er[g.id].env_radiance.grad = d_grad + g.offset;
                if (g.kind == PB_PARAM_ENVMAP_SCALE) er[g.id].env_scale_grad = d_grad + g.offset;
                if (g.kind == PB_PARAM_ENVMAP_TRANSFORM) {   // the kernels see from_world = inverse(left * raw): its adjoint (or tangent) lives in a scratch buffer
                    const HostEmitter &he = c->emitters[g.id];
                    c->d_env_xf_acc.reserve(16 * sizeof(float));
                    if (jvp) {   // F_t = -F (L_t R) F
                        float lt[16], ft[16];
                        PB_CUDA(cudaMemcpyAsync(lt, d_grad + g.offset, sizeof(lt), cudaMemcpyDeviceToHost, st));
                        PB_CUDA(cudaStreamSynchronize(st));
                        Mat4h Lt; std::memcpy(Lt.m, lt, sizeof(lt));
                        const Mat4h F = inverse(matmul(he.env_left, he.env_raw));
                        const Mat4h T = matmul(matmul(F, matmul(Lt, he.env_raw)), F);
                        for (int k = 0; k < 16; ++k) ft[k] = -T.m[k];
                        PB_CUDA(cudaMemcpyAsync(c->d_env_xf_acc.p, ft, sizeof(ft), cudaMemcpyHostToDevice, st));
                        PB_CUDA(cudaStreamSynchronize(st));
                    } else {
                        PB_CUDA(cudaMemsetAsync(c->d_env_xf_acc.p, 0, 16 * sizeof(float), st));
                    }
                    er[g.id].env_xf_grad = c->d_env_xf_acc.as<float>();
                }
            }
            c->d_emitters_grad.upload(er, st);
            P.S.emitters = c->d_emitters_grad.as<EmitterRec>();
        }
        bool any_geom = false;
        for (const GradSegment &g : c->grad_segments) any_geom = any_geom || g.kind == PB_PARAM_MESH_VERTICES || g.kind == PB_PARAM_SENSOR_TRANSFORM;
        for (const GradSegment &g : c->grad_segments)
            if (g.kind == PB_PARAM_SENSOR_TRANSFORM && g.id == sensor) {   // pose of the sensor being rendered (other sensors get a zero gradient)
                c->d_sensor_acc.reserve(32 * sizeof(float));
                if (jvp) {   // tangents of to_world and of world_to_sample
                    float h[32];
                    PB_CUDA(cudaMemcpyAsync(h, d_grad + g.offset, 16 * sizeof(float), cudaMemcpyDeviceToHost, st));
                    PB_CUDA(cudaStreamSynchronize(st));
                    w2s_tangent(c->sensors[sensor], h, h + 16);
                    PB_CUDA(cudaMemcpyAsync(c->d_sensor_acc.p, h, sizeof(h), cudaMemcpyHostToDevice, st));
                    PB_CUDA(cudaStreamSynchronize(st));
                } else {
                    PB_CUDA(cudaMemsetAsync(c->d_sensor_acc.p, 0, 32 * sizeof(float), st));
                }
                P.S.sensor_grad = c->d_sensor_acc.as<float>();
            }
        if (any_geom) {
            c->d_tri_grad.reserve((size_t)std::max(1, c->num_tri) * kTriGradStride * sizeof(float));
            PB_CUDA(cudaMemsetAsync(c->d_tri_grad.p, 0, (size_t)std::max(1, c->num_tri) * kTriGradStride * sizeof(float), st));
            P.S.tri_grad = c->d_tri_grad.as<float>();
            for (const GradSegment &g : c->grad_segments)
                if (g.kind == PB_PARAM_MESH_VERTICES) {
                    HostMesh &m = c->meshes[g.id];
                    m.d_gworld.reserve(3 * (size_t)m.nv * sizeof(float));
                    PB_CUDA(cudaMemsetAsync(m.d_gworld.p, 0, 3 * (size_t)m.nv * sizeof(float), st));
                }
        }
        if (jvp) {   // forward mode: tangent of the triangle table from the vertex tangents (mesh.cpp:19-51,215-231 in forward mode)
            const size_t tt_bytes = (size_t)std::max(1, c->num_tri) * kTriGradStride * sizeof(float);
            c->d_tri_tangent.reserve(tt_bytes);
            PB_CUDA(cudaMemsetAsync(c->d_tri_tangent.p, 0, tt_bytes, st));
            for (const GradSegment &g : c->grad_segments)
                if (g.kind == PB_PARAM_MESH_VERTICES) {
                    HostMesh &m = c->meshes[g.id];
                    m.d_fcross_t.reserve((size_t)m.nf * sizeof(float4));
                    m.d_vnormal_t.reserve(3 * (size_t)m.nv * sizeof(float));
                    launch_mesh_tangent(st, m.nv, m.nf, m.face_offset, m.d_vraw.as<float>(), d_grad + g.offset, to_dev(m.to_world), m.d_vworld.as<float>(),
                                        m.d_faces.as<int>(), m.d_csr_off.as<int>(), m.d_csr_face.as<int>(), m.d_fcross.as<float4>(), m.d_gworld.as<float>(),
                                        m.d_fcross_t.as<float4>(), m.d_vnormal_t.as<float>(), c->d_tri_tangent.as<float>());
                    c->launches += 4;
                }
            c->d_jvp_acc.reserve((size_t)B * sizeof(float));
            P.S.tri_grad = nullptr;
            P.S.tri_tangent = c->d_tri_tangent.as<float>();
            P.S.jvp_acc = c->d_jvp_acc.as<float>();
            P.S.jvp_image = d_dimage;
            P.S.jvp_channel = 0;
        }
    }
    std::vector<BounceParams> bps(plan.nbounce);
    for (int k = 0; k < plan.nbounce; ++k) {
        BounceParams &Bp = bps[k];
        Bp.nb = plan.nb; Bp.nl = plan.nl; Bp.depth = k; Bp.last = (k == plan.nbounce - 1); Bp.carry = (I.kind == PB_INTEG_PATH);
        Bp.hide_emitters = I.hide_emitters; Bp.ad = (mode == MODE_C) ? 0 : 1;
        Bp.rc_grad = 0;
        if (mode == MODE_VJP)
            for (const GradSegment &g : c->grad_segments)
                if (g.kind == PB_PARAM_BSDF_TEXTURE && c->bsdfs[g.id].type == PB_BSDF_ROUGHCONDUCTOR) Bp.rc_grad = 1;
        if (mode == MODE_VJP && any_geom_jvp(c))   // geometry adjoints flow through every rough-conductor vertex on a path
            for (const HostBsdf &hb : c->bsdfs) if (hb.type == PB_BSDF_ROUGHCONDUCTOR) Bp.rc_grad = 1;
        if (mode == MODE_VJP)
            for (const GradSegment &g : c->grad_segments) if (g.kind == PB_PARAM_SENSOR_TRANSFORM) Bp.rc_grad = 1;   // the pose adjoint lives in the extended kernel
        if (mode == MODE_VJP)
            for (const GradSegment &g : c->grad_segments) if (g.kind == PB_PARAM_MESH_UV) Bp.rc_grad = 1;
        if (mode == MODE_VJP && any_geom_jvp(c))   // bitmap textures: the camera vertex' uv moves with the geometry
            for (const HostBsdf &hb : c->bsdfs) for (int k = 0; k < TEX_COUNT; ++k) if (hb.tex[k].w * hb.tex[k].h > 1) Bp.rc_grad = 1;
        if (mode == MODE_VJP && c->emitter_env >= 0) {   // environment map: radiance / scale gradients, and its direction term in the geometry adjoints
            const HostEmitter &he = c->emitters[c->emitter_env];
            if (he.env_radiance.requires_grad || he.env_scale_requires_grad || he.env_xf_requires_grad || any_geom_jvp(c)) Bp.rc_grad = 1;
        }
        Bp.jump = make_jump(base + 2 + (uint64_t)k * (3 * plan.nb + 2 * plan.nl));
        set_sort_params(c, Bp);
    }
    size_t nev = 0, nev_edge = 0;
    cudaStream_t const st0 = st;
    if (dual) {   // everything enqueued so far (film clear, table uploads) precedes both lanes
        PB_CUDA(cudaEventRecord(c->ev_fork, st0));
        PB_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
    }
    int batch_index = 0;
    for (int64_t start = 0; start < total && !(field && mode == MODE_VJP); start += B, ++batch_index) {
        const int lane = dual ? (batch_index & 1) : 0;
        st = lane ? c->stream2 : st0;
        EventStore &S = use_retained ? c->retained : (lane ? c->scratch1 : c->scratch);
        RayRec *const lane_rays = lane ? c->d_rays1.as<RayRec>() : (use_retained ? c->retained.rays.as<RayRec>() : c->scratch.rays.as<RayRec>());
        unsigned short *const lane_keys = lane ? c->d_sort_keys1.as<unsigned short>() : c->d_sort_keys.as<unsigned short>();
        float4 *const lane_suffix = lane ? c->d_suffix1.as<float4>() : c->d_suffix.as<float4>();
        P.local0 = start; P.n = (int)std::min<int64_t>(B, total - start);
        const size_t off = use_retained ? (size_t)start : 0;
        HitRec *hit0 = S.hit0.as<HitRec>() + off;
        auto event = [&](int k) {
            const int sl = keep ? k : (k & 1), sp = keep ? k - 1 : ((k - 1) & 1);
            EventBuffers E;
            E.hit_cur = (k == 0) ? hit0 : S.hits[sp].as<HitRec>() + off * R;
            E.hit_prev = (k == 0 || !keep) ? nullptr : (k == 1 ? hit0 : S.hits[k - 2].as<HitRec>() + off * R);
            E.prev_pos = (k == 0) ? nullptr : S.pos[sp].as<float4>() + off;
            E.pos = S.pos[sl].as<float4>() + off; E.vb = S.vb[sl].as<float4>() + off; E.vc = S.vc[sl].as<float4>() + off;
            E.rays = lane_rays;
            E.hits = S.hits[sl].as<HitRec>() + off * R;
            E.thr_in = (k == 0) ? nullptr : S.thr[keep ? k : (k & 1)].as<float4>() + off;
            E.thr_out = bps[k].last ? nullptr : S.thr[keep ? k + 1 : ((k + 1) & 1)].as<float4>() + off;
            E.rad = S.rad.as<float4>() + off;
            E.keys = lane_keys;
            E.lin = (keep && lin) ? S.lin[sl].as<float4>() + off : nullptr;
            E.inv = nullptr; E.inv_cur = nullptr; E.lane_list = nullptr; E.lane_count = nullptr;
            E.conn = c->d_conn[lane].as<float2>();
            return E;
        };
        if (run_forward) {
            cudaEvent_t e0 = get_event(c, nev++), e1 = get_event(c, nev++);
            PB_CUDA(cudaEventRecord(e0, st));
            launch_primary(st, P, hit0);
            PB_CUDA(cudaEventRecord(e1, st));
            c->launches++;
            if (field) {
                if (mode != MODE_VJP) { launch_field(st, P, I.field, hit0, d_image, nullptr); c->launches++; }
                continue;
            }
            const unsigned *prev_inv = nullptr;   // sorted-copy traversal: the previous event's hits are in stream order, this is their inverse map
            for (int k = 0; k < plan.nbounce; ++k) {
                EventBuffers E = event(k);
                E.inv_cur = prev_inv;
                launch_shade(st, P, bps[k], E);
                cudaEvent_t t0 = get_event(c, nev++), t1 = get_event(c, nev++);
                const int64_t nrays = (int64_t)P.n * (plan.nb + plan.nl);
                if (!sorted_copy) {
                    trace_wavefront(c, nrays, E.rays, E.hits, t0, t1, -1, true, lane);
                } else if (hits_by_slot) {   // someone indexes this event's hits by ray slot later (k_adjoint, the next event's load_vertex): un-permute them
                    c->d_sorted_hits[lane].reserve((size_t)nrays * sizeof(HitRec));
                    const unsigned *inv = nullptr;
                    trace_wavefront(c, nrays, E.rays, c->d_sorted_hits[lane].as<HitRec>(), t0, t1, -1, true, lane, &inv);
                    launch_unpermute_hits(st, nrays, inv, c->d_sorted_hits[lane].as<HitRec>(), E.hits);
                    c->launches++;
                } else {                     // k_resolve is the only reader: it follows the inverse map
                    trace_wavefront(c, nrays, E.rays, E.hits, t0, t1, -1, true, lane, &E.inv);
                }
                prev_inv = E.inv;
                launch_resolve(st, P, bps[k], E, mode == MODE_VJP ? nullptr : d_image);
                c->launches += 3; c->last_rays += nrays; c->last_trace_launches++;
            }
        }
        if (mode == MODE_VJP) {
            for (int ch = 0; ch < (jvp ? 3 : 1); ++ch) {   // forward mode: one pass per colour channel with a unit seed
                if (jvp) { P.S.jvp_channel = ch; PB_CUDA(cudaMemsetAsync(c->d_jvp_acc.p, 0, (size_t)P.n * sizeof(float), st)); }
                for (int k = plan.nbounce - 1; k >= 0; --k) {
                    c->d_adj_list[lane].reserve((size_t)B * sizeof(int)); c->d_adj_count[lane].reserve(sizeof(unsigned));
                    launch_adjoint(st, P, bps[k], event(k), lane_suffix, d_dLdI, c->d_adj_list[lane].as<int>(), c->d_adj_count[lane].as<unsigned>());
                    c->launches++;
                }
            }
        }
    }
    st = st0;
    if (dual) {   // join: whatever follows on the context's stream (edge terms, mesh backward, the caller) sees both lanes' results
        PB_CUDA(cudaEventRecord(c->ev_join, c->stream2));
        PB_CUDA(cudaStreamWaitEvent(st0, c->ev_join, 0));
    }
    if (mode == MODE_VJP && (P.S.tri_grad || (jvp && any_geom_jvp(c)))) run_edge_terms(c, I, sensor, plan, P, d_dLdI, &nev_edge);
    if (mode == MODE_VJP && !jvp)   // adjoint of the envmap's from_world = inverse(left * raw) -> adjoint of `left`: g_L = -(F^T g_F F^T) R^T
        for (const GradSegment &g : c->grad_segments)
            if (g.kind == PB_PARAM_ENVMAP_TRANSFORM) {
                const HostEmitter &he = c->emitters[g.id];
                float gf[16], cur[16];
                PB_CUDA(cudaMemcpyAsync(gf, c->d_env_xf_acc.p, sizeof(gf), cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaMemcpyAsync(cur, d_grad + g.offset, sizeof(cur), cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaStreamSynchronize(st));
                const Mat4h F = inverse(matmul(he.env_left, he.env_raw));
                double a1[16], a2[16];
                for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += (double)F.m[4 * k + i] * gf[4 * k + j]; a1[4 * i + j] = a; }          // F^T g_F
                for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += a1[4 * i + k] * (double)F.m[4 * j + k]; a2[4 * i + j] = -a; }        // -(..) F^T
                for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += a2[4 * i + k] * (double)he.env_raw.m[4 * j + k]; cur[4 * i + j] += (float)a; }   // (..) R^T
                PB_CUDA(cudaMemcpyAsync(d_grad + g.offset, cur, sizeof(cur), cudaMemcpyHostToDevice, st));
                PB_CUDA(cudaStreamSynchronize(st));
            }
    if (mode == MODE_VJP && !jvp && P.S.sensor_grad) {   // fold the world_to_sample adjoint into the to_world adjoint and add it to the gradient vector
        float h[32], cur[16];
        PB_CUDA(cudaMemcpyAsync(h, c->d_sensor_acc.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        for (const GradSegment &g : c->grad_segments)
            if (g.kind == PB_PARAM_SENSOR_TRANSFORM && g.id == sensor) {
                PB_CUDA(cudaMemcpyAsync(cur, d_grad + g.offset, sizeof(cur), cudaMemcpyDeviceToHost, st));
                PB_CUDA(cudaStreamSynchronize(st));
                fold_w2s_adjoint(c->sensors[sensor], h + 16, h);
                for (int k = 0; k < 16; ++k) cur[k] += h[k];
                PB_CUDA(cudaMemcpyAsync(d_grad + g.offset, cur, sizeof(cur), cudaMemcpyHostToDevice, st));
                PB_CUDA(cudaStreamSynchronize(st));
            }
    }
    if (mode == MODE_VJP && P.S.tri_grad) {   // triangle-table adjoint -> object-space vertex gradients (mesh.cpp:19-51,215-231 backward)
        for (const GradSegment &g : c->grad_segments)
            if (g.kind == PB_PARAM_MESH_VERTICES) {
                HostMesh &m = c->meshes[g.id];
                m.d_gnsum.reserve(3 * (size_t)m.nv * sizeof(float));
                m.d_gcorner.reserve(9 * (size_t)m.nf * sizeof(float));
                launch_mesh_backward(st, m.nv, m.nf, m.face_offset, m.d_csr_off.as<int>(), m.d_csr_slot.as<int>(), m.d_fcross.as<float4>(),
                                     m.d_vworld.as<float>(), m.d_faces.as<int>(), m.d_vraw.as<float>(), to_dev(m.to_world), c->d_tri_grad.as<float>(),
                                     m.d_gworld.as<float>(), m.d_gnsum.as<float>(), m.d_gcorner.as<float>(), d_grad + g.offset);
                c->launches += 3;
            }
    }
    PB_CUDA(cudaGetLastError());
    PB_CUDA(cudaStreamSynchronize(st));
    {   // rays the traversal kernels actually traced (inactive lanes are compacted away by the sort)
        unsigned long long act = 0;
        PB_CUDA(cudaMemcpy(&act, c->d_active_total.p, sizeof(act), cudaMemcpyDeviceToHost));
        c->last_active_rays = (int64_t)act;
    }
    // event pairs: per batch one for the primary kernel, then one per k_trace launch (the traversal kernel alone, without the sort)
    {
        const size_t per_batch = 2 + 2 * (size_t)(field ? 0 : plan.nbounce);
        for (size_t i = 0; i + 1 < nev; i += 2) {
            float ms = 0.f;
            PB_CUDA(cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]));
            if (i % per_batch == 0) c->last_primary_ms += ms; else c->last_trace_ms += ms;
        }
    }
    if (mode == MODE_D && use_retained) {
        c->retained_B = B;
        c->retained_hits_by_slot = !sorted_copy || hits_by_slot;
        c->retained_valid = true; c->retained_kind = I.kind; c->retained_nb = plan.nb; c->retained_nl = plan.nl;
        c->retained_nbounce = plan.nbounce; c->retained_sensor = sensor; c->retained_hide = I.hide_emitters;
    }
    if (mode != MODE_VJP) c->sampler_offset[0] = base + plan.draws;
    if (mode == MODE_D) {   // integrator.cpp:41-47: renderD also consumes the two edge samplers
        const uint64_t per_li = (uint64_t)plan.nbounce * (3 * plan.nb + 2 * plan.nl);
        c->last_d_offset_e = c->sampler_offset[1]; c->last_d_offset_s = c->sampler_offset[2];
        if (c->sppe > 0 && c->sensors[sensor].num_prim > 0) c->sampler_offset[1] += 1 + 2 * per_li;
        if (c->sppse > 0 && !field) c->sampler_offset[2] += 3;
    }
}

}  // namespace pb

using namespace pb;

static thread_local std::string g_create_error;

template <class F> static int guard(pb_ctx *c, F &&f) {
    try { f(); return 0; }
    catch (const std::exception &e) { if (c) c->error = e.what(); else g_create_error = e.what(); return 1; }
}
template <class F> static int guard_id(pb_ctx *c, F &&f) {   // returns id >= 0 or -1
    try { return f(); }
    catch (const std::exception &e) { if (c) c->error = e.what(); return -1; }
}

extern "C" {

int pb_version(void) { return 100; }

int pb_ctx_create(int device, pb_ctx **out) {
    *out = nullptr;
    return guard(nullptr, [&] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0) throw Error(std::string("psdr_b200 needs a CUDA device (no CPU fallback): ") + cudaGetErrorString(e));
        PB_ASSERT_MSG(device >= 0 && device < count, "Invalid CUDA device index");
        PB_CUDA(cudaSetDevice(device));
        pb_ctx *c = new pb_ctx();
        c->device = device;
        PB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        *out = c;
    });
}
int pb_ctx_destroy(pb_ctx *c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    pb_dist_finalize(c);
    if (!c->own_stream && c->l2_persist) { c->l2_persist = 0; set_l2_window(c); }   // leave the caller's stream as it was
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join); }
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}
const char *pb_last_error(pb_ctx *c) { return c ? c->error.c_str() : g_create_error.c_str(); }
int pb_ctx_set_batch(pb_ctx *c, int64_t lanes) {
    return guard(c, [&] {
        if (lanes <= 0) lanes = 1 << 25;
        PB_ASSERT_MSG(lanes % 1024 == 0, "batch must be a multiple of 1024 lanes");
        c->batch = lanes;
    });
}
int pb_ctx_set_shard(pb_ctx *c, int rank, int world) {
    return guard(c, [&] { PB_ASSERT_MSG(world >= 1 && rank >= 0 && rank < world, "Invalid shard"); c->rank = rank; c->world = world; c->retained_valid = false; });
}
int pb_ctx_set_shard_mode(pb_ctx *c, int mode, int tile_rows) {
    return guard(c, [&] {
        PB_ASSERT_MSG((mode == PB_SHARD_SAMPLES || mode == PB_SHARD_PIXELS) && tile_rows >= 0, "Invalid shard mode");
        c->shard_mode = mode; c->tile_rows = tile_rows; c->retained_valid = false;
    });
}

int pb_scene_set_options(pb_ctx *c, int w, int h, int spp, int sppe, int sppse) {
    return guard(c, [&] {
        PB_ASSERT_MSG(w >= 0 && h >= 0 && spp >= 0 && sppe >= 0 && sppse >= 0, "Invalid render options");
        c->width = w; c->height = h; c->spp = spp; c->sppe = sppe; c->sppse = sppse;
        c->ready = false;
    });
}
int pb_scene_add_sensor(pb_ctx *c, float fov_x, float near_clip, float far_clip, const float *to_world) {
    return guard_id(c, [&] {
        HostSensor s;
        s.fov_x = fov_x; s.near_clip = near_clip; s.far_clip = far_clip; s.to_world = from_ptr(to_world);
        c->sensors.push_back(std::move(s));
        c->ready = false;
        return (int)c->sensors.size() - 1;
    });
}
int pb_scene_set_sensor_transform(pb_ctx *c, int sensor, const float *to_world) {
    return guard(c, [&] {
        PB_ASSERT_MSG(sensor >= 0 && sensor < (int)c->sensors.size(), "Invalid sensor id!");
        c->sensors[sensor].to_world = from_ptr(to_world);
        c->ready = false;
    });
}
int pb_scene_add_bsdf(pb_ctx *c, int type) {
    return guard_id(c, [&] {
        PB_ASSERT_MSG(type == PB_BSDF_DIFFUSE || type == PB_BSDF_ROUGHCONDUCTOR, "Unsupported BSDF");
        c->bsdfs.emplace_back();
        HostBsdf &b = c->bsdfs.back();
        b.type = type;
        // defaults: diffuse.h:11-14, roughconductor.h:11-12
        const float def3[TEX_COUNT] = {.5f, 0.f, 0.f, 0.f, 1.f, 1.f};
        for (int k = 0; k < TEX_COUNT; ++k) {
            const bool one = (k == TEX_ALPHA_U || k == TEX_ALPHA_V);
            b.tex[k].c = one ? 1 : 3;
            b.tex[k].data.assign(one ? 1 : 3, one ? .1f : def3[k]);
        }
        c->ready = false;
        return (int)c->bsdfs.size() - 1;
    });
}
int pb_scene_set_bsdf_texture(pb_ctx *c, int bsdf, int slot, const float *data, int w, int h) {
    return guard(c, [&] {
        PB_ASSERT_MSG(bsdf >= 0 && bsdf < (int)c->bsdfs.size(), "Invalid BSDF id");
        PB_ASSERT_MSG(slot >= 0 && slot < TEX_COUNT && w >= 1 && h >= 1 && data, "Invalid texture");
        HostTexture &t = c->bsdfs[bsdf].tex[slot];
        t.w = w; t.h = h;
        t.data.assign(data, data + (size_t)w * h * t.c);
        t.dirty = true;
        c->ready = false;
    });
}
int pb_scene_add_mesh(pb_ctx *c, int nv, int nf, const float *verts, const int *faces, int nuv, const float *uvs, const int *uv_faces, int flags,
                      int bsdf, const float *to_world) {
    return guard_id(c, [&] {
        PB_ASSERT_MSG(nv >= 0 && nf >= 0 && (nv == 0 || verts) && (nf == 0 || faces), "Invalid mesh buffers");
        PB_ASSERT_MSG(bsdf >= -1 && bsdf < (int)c->bsdfs.size(), "Unknown BSDF id");
        if (c->has_bound_mesh) {   // the environment map's bounding box is always the last mesh (scene.cpp:135-180): take it off, configure re-creates it
            c->meshes.pop_back(); c->has_bound_mesh = false;
            if (c->emitter_env >= 0) c->emitters[c->emitter_env].mesh = -1;
        }
        c->meshes.emplace_back();
        HostMesh &m = c->meshes.back();
        try {
            m.nv = nv; m.nf = nf;
            m.verts.assign(verts, verts + 3 * (size_t)nv);
            m.faces.assign(faces, faces + 3 * (size_t)nf);
            for (int i = 0; i < 3 * nf; ++i) PB_ASSERT_MSG(m.faces[i] >= 0 && m.faces[i] < nv, "Face index out of range");
            const bool has_uv = nuv > 0 && uvs && uv_faces;
            if (has_uv) {
                m.uvs.assign(uvs, uvs + 2 * (size_t)nuv);
                m.uv_faces.assign(uv_faces, uv_faces + 3 * (size_t)nf);
                for (int i = 0; i < 3 * nf; ++i) PB_ASSERT_MSG(m.uv_faces[i] >= 0 && m.uv_faces[i] < nuv, "UV index out of range");
            }
            // device flags: bit0 face normals, bit1 has uv; bit2 keeps enable_edges on the host side
            m.flags = ((flags & PB_MESH_FACE_NORMALS) ? 1 : 0) | (has_uv ? 2 : 0) | ((flags & PB_MESH_ENABLE_EDGES) ? 4 : 0);
            m.bsdf = bsdf;
            m.raw = from_ptr(to_world);
            build_edges(m);
            build_csr(m);
        } catch (...) { c->meshes.pop_back(); throw; }
        c->ready = false;
        return (int)c->meshes.size() - 1;
    });
}
int pb_scene_set_mesh_vertices(pb_ctx *c, int mesh, const float *verts) {
    return guard(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size() && verts, "Invalid mesh id");
        HostMesh &m = c->meshes[mesh];
        m.verts.assign(verts, verts + 3 * (size_t)m.nv);
        m.verts_dirty = true;
        m.verts_host_stale = false;
        c->ready = false;
    });
}
/* the same two setters from DEVICE memory: an optimiser that keeps its parameters on the GPU (torch.optim on CUDA leaves) updates the scene
 * without a host round trip; the copies are enqueued on the context's stream, configure() then refits / rebuilds on the device */
int pb_scene_set_mesh_vertices_device(pb_ctx *c, int mesh, const float *d_verts) {
    return guard(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size() && d_verts, "Invalid mesh id");
        HostMesh &m = c->meshes[mesh];
        PB_CUDA(cudaSetDevice(c->device));
        m.d_vraw.reserve(3 * (size_t)std::max(1, m.nv) * sizeof(float));
        PB_CUDA(cudaMemcpyAsync(m.d_vraw.p, d_verts, 3 * (size_t)m.nv * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
        m.verts_dirty = false;      // the device copy is the current one (the host copy is refreshed on demand: pb_scene_get_mesh_vertices)
        m.verts_host_stale = true;
        c->ready = false;
    });
}
int pb_scene_set_bsdf_texture_device(pb_ctx *c, int bsdf, int slot, const float *d_data) {
    return guard(c, [&] {
        PB_ASSERT_MSG(bsdf >= 0 && bsdf < (int)c->bsdfs.size(), "Invalid BSDF id");
        PB_ASSERT_MSG(slot >= 0 && slot < TEX_COUNT && d_data, "Invalid texture");
        HostTexture &t = c->bsdfs[bsdf].tex[slot];
        PB_CUDA(cudaSetDevice(c->device));
        const size_t bytes = (size_t)t.w * t.h * t.c * sizeof(float);
        t.d.reserve(bytes);
        PB_CUDA(cudaMemcpyAsync(t.d.p, d_data, bytes, cudaMemcpyDeviceToDevice, c->stream));   // same resolution as set at creation / by the host setter
        t.dirty = false;
        c->ready = false;
    });
}
int pb_scene_get_mesh_vertices(pb_ctx *c, int mesh, float *h_out) {
    return guard(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size() && h_out, "Invalid mesh id");
        HostMesh &m = c->meshes[mesh];
        if (m.verts_host_stale) {
            PB_CUDA(cudaSetDevice(c->device));
            PB_CUDA(cudaMemcpyAsync(m.verts.data(), m.d_vraw.p, 3 * (size_t)m.nv * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
            PB_CUDA(cudaStreamSynchronize(c->stream));
            m.verts_host_stale = false;
        }
        std::copy(m.verts.begin(), m.verts.end(), h_out);
    });
}
/* importance of a secondary edge in the edge distribution: 0 = its length (the reference, scene.cpp:236), 1 = length x exterior dihedral
 * angle (pi for boundary edges) — the alternative the reference keeps under `#if 0` (scene.cpp:230-233): flat creases are rarely silhouettes */
int pb_scene_set_edge_importance(pb_ctx *c, int mode) {
    return guard(c, [&] { PB_ASSERT_MSG(mode == 0 || mode == 1, "Invalid edge importance"); c->edge_importance = mode; c->ready = false; });
}
int pb_scene_set_mesh_uvs(pb_ctx *c, int mesh, const float *uvs) {
    return guard(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size() && uvs && (c->meshes[mesh].flags & 2), "Invalid mesh id, or the mesh has no texture coordinates");
        HostMesh &m = c->meshes[mesh];
        m.uvs.assign(uvs, uvs + m.uvs.size());   // count and indices stay those of pb_scene_add_mesh
        m.uv_dirty = true;
        c->ready = false;
    });
}
int pb_scene_set_mesh_transform(pb_ctx *c, int mesh, const float *mat, int left) {
    return guard(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size(), "Invalid mesh id");
        (left ? c->meshes[mesh].left : c->meshes[mesh].right) = from_ptr(mat);
        c->ready = false;
    });
}
int pb_scene_add_area_emitter(pb_ctx *c, int mesh, const float *radiance) {
    return guard_id(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size() && radiance, "Invalid mesh id");
        HostEmitter e;
        e.type = EMITTER_AREA; e.mesh = mesh;
        for (int k = 0; k < 3; ++k) e.radiance[k] = radiance[k];
        c->emitters.push_back(std::move(e));
        c->meshes[mesh].emitter = (int)c->emitters.size() - 1;
        c->ready = false;
        return (int)c->emitters.size() - 1;
    });
}
int pb_scene_add_envmap(pb_ctx *c, int w, int h, const float *rgb, float scale, const float *to_world) {
    return guard_id(c, [&] {
        PB_ASSERT_MSG(c->emitter_env < 0, "A scene is only allowed to have one envmap!");
        PB_ASSERT_MSG(rgb && w > 1 && h > 1, "Invalid environment map");
        PB_ASSERT_MSG(!c->has_bound_mesh, "internal: bounding mesh present");
        c->emitters.emplace_back();
        HostEmitter &e = c->emitters.back();
        e.type = EMITTER_ENVMAP;
        e.env_radiance.w = w; e.env_radiance.h = h; e.env_radiance.c = 3;
        e.env_radiance.data.assign(rgb, rgb + (size_t)w * h * 3);
        e.env_scale = scale;
        e.env_raw = from_ptr(to_world);
        c->emitter_env = (int)c->emitters.size() - 1;
        c->ready = false;
        return c->emitter_env;
    });
}
int pb_scene_set_envmap_radiance(pb_ctx *c, const float *rgb, float scale) {
    return guard(c, [&] {
        PB_ASSERT_MSG(c->emitter_env >= 0, "No environment map");
        HostEmitter &e = c->emitters[c->emitter_env];
        if (rgb) { e.env_radiance.data.assign(rgb, rgb + e.env_radiance.data.size()); e.env_dirty = true; }   // resolution is fixed at creation
        e.env_scale = scale;
        c->ready = false;
    });
}
int pb_scene_set_envmap_transform(pb_ctx *c, const float *left) {
    return guard(c, [&] {
        PB_ASSERT_MSG(c->emitter_env >= 0, "No environment map");
        c->emitters[c->emitter_env].env_left = from_ptr(left);
        c->ready = false;
    });
}
int pb_scene_num_meshes(pb_ctx *c) { return (int)c->meshes.size(); }
int pb_scene_configure(pb_ctx *c) { return guard(c, [&] { configure(c); }); }
int pb_scene_reseed(pb_ctx *c) {
    for (int k = 0; k < 3; ++k) { c->sampler_count[k] = 0; c->sampler_offset[k] = 0; }
    c->ready = false;
    return 0;
}
int pb_scene_num_triangles(pb_ctx *c) { return c->num_tri; }
int pb_scene_get_triangle_info(pb_ctx *c, float *out) {
    return guard(c, [&] {
        PB_ASSERT_MSG(c->ready, "Input scene must be configured!");
        PB_CUDA(cudaSetDevice(c->device));
        download_triangle_table(c);
        for (int t = 0; t < c->num_tri; ++t) {
            const float *q = &c->h_tri[(size_t)t * 32];
            float *o = out + 22 * (size_t)t;
            for (int a = 0; a < 3; ++a) {
                o[a] = q[a]; o[3 + a] = q[4 + a]; o[6 + a] = q[8 + a];              // p0 e1 e2
                o[9 + a] = q[12 + a]; o[12 + a] = q[16 + a]; o[15 + a] = q[20 + a];  // n0 n1 n2
                o[18 + a] = q[24 + a];                                              // face normal
            }
            o[21] = q[3];
        }
    });
}
/* edge tables as configure built them (device -> host copies for inspection / parity tests) */
int pb_scene_num_primary_edges(pb_ctx *c, int sensor) {
    if (!c || !c->ready || sensor < 0 || sensor >= (int)c->sensors.size()) return -1;
    return c->sensors[sensor].num_prim;
}
int pb_scene_get_primary_edges(pb_ctx *c, int sensor, float *out, float *cmf_out) {   // [n][7]: p0.xy p1.xy edge_normal.xy edge_length; cmf [n]
    return guard(c, [&] {
        PB_ASSERT_MSG(c->ready, "Input scene must be configured!");
        PB_ASSERT_MSG(sensor >= 0 && sensor < (int)c->sensors.size(), "Invalid sensor id!");
        const HostSensor &s = c->sensors[sensor];
        PB_CUDA(cudaSetDevice(c->device));
        std::vector<PrimEdgeRec> recs(s.num_prim);
        if (s.num_prim) PB_CUDA(cudaMemcpyAsync(recs.data(), s.d_prim.p, recs.size() * sizeof(PrimEdgeRec), cudaMemcpyDeviceToHost, c->stream));
        if (s.num_prim && cmf_out) PB_CUDA(cudaMemcpyAsync(cmf_out, s.d_prim_cmf.p, recs.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        PB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < recs.size(); ++i) {
            const PrimEdgeRec &r = recs[i];
            const float v[7] = {r.p0x, r.p0y, r.p1x, r.p1y, r.nx, r.ny, r.len};
            std::copy(v, v + 7, out + 7 * i);
        }
    });
}
int pb_scene_num_secondary_edges(pb_ctx *c) { return (c && c->ready) ? c->num_sec : -1; }
int pb_scene_get_secondary_edges(pb_ctx *c, float *out, float *cmf_out) {   // [n][16]: p0 e1 n0 n1 p2 is_boundary; cmf [n]
    return guard(c, [&] {
        PB_ASSERT_MSG(c->ready, "Input scene must be configured!");
        PB_CUDA(cudaSetDevice(c->device));
        std::vector<SecEdgeRec> recs(c->num_sec);
        if (c->num_sec) PB_CUDA(cudaMemcpyAsync(recs.data(), c->d_sec.p, recs.size() * sizeof(SecEdgeRec), cudaMemcpyDeviceToHost, c->stream));
        if (c->num_sec && cmf_out) PB_CUDA(cudaMemcpyAsync(cmf_out, c->d_sec_cmf.p, recs.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        PB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t i = 0; i < recs.size(); ++i) {
            const SecEdgeRec &r = recs[i];
            const float v[16] = {r.a.x, r.a.y, r.a.z, r.b.x, r.b.y, r.b.z, r.c.x, r.c.y, r.c.z, r.d.x, r.d.y, r.d.z, r.e.x, r.e.y, r.e.z, r.a.w};
            std::copy(v, v + 16, out + 16 * i);
        }
    });
}
int pb_scene_mesh_num_edges(pb_ctx *c, int mesh) {
    if (mesh < 0 || mesh >= (int)c->meshes.size()) return -1;
    return (int)c->meshes[mesh].edges.size() / 5;
}
int pb_scene_mesh_get_edges(pb_ctx *c, int mesh, int *out) {
    return guard(c, [&] {
        PB_ASSERT_MSG(mesh >= 0 && mesh < (int)c->meshes.size(), "Invalid mesh id");
        std::copy(c->meshes[mesh].edges.begin(), c->meshes[mesh].edges.end(), out);
    });
}

int pb_trace(pb_ctx *c, int64_t n, const float *d_rays, void *d_hits, float *d_t) {
    return guard(c, [&] {
        PB_ASSERT_MSG(c->ready, "Input scene must be configured!");
        PB_CUDA(cudaSetDevice(c->device));
        cudaEvent_t e0 = get_event(c, 0), e1 = get_event(c, 1);
        PB_CUDA(cudaEventRecord(e0, c->stream));
        if (d_t) launch_trace(c->stream, c->view, n, reinterpret_cast<const RayRec *>(d_rays), reinterpret_cast<HitRec *>(d_hits), d_t);
        else trace_wavefront(c, n, reinterpret_cast<const RayRec *>(d_rays), reinterpret_cast<HitRec *>(d_hits));   // what the render calls use
        PB_CUDA(cudaEventRecord(e1, c->stream));
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaStreamSynchronize(c->stream));
        PB_CUDA(cudaEventElapsedTime(&c->last_trace_ms, e0, e1));
        c->launches++; c->last_rays = n; c->last_trace_launches = 1; c->last_primary_ms = 0.f;
    });
}
int pb_render_c(pb_ctx *c, const pb_integrator *I, int sensor, float *d_image) {
    return guard(c, [&] { PB_ASSERT_MSG(I && d_image, "Null argument"); render_interior(c, *I, sensor, d_image, MODE_C); });
}
int pb_render_c_host(pb_ctx *c, const pb_integrator *I, int sensor, float *h_image) {
    return guard(c, [&] {
        PB_ASSERT_MSG(I && h_image, "Null argument");
        const size_t bytes = (size_t)c->width * c->height * 3 * sizeof(float);
        DevBuf img;
        img.reserve(bytes);
        render_interior(c, *I, sensor, img.as<float>(), MODE_C);
        PB_CUDA(cudaMemcpy(h_image, img.p, bytes, cudaMemcpyDeviceToHost));
    });
}
int pb_render_d(pb_ctx *c, const pb_integrator *I, int sensor, float *d_image) {
    return guard(c, [&] {
        PB_ASSERT_MSG(I && d_image, "Null argument");
        const uint64_t off = c->sampler_offset[0];
        render_interior(c, *I, sensor, d_image, MODE_D);
        c->last_d_offset = off; c->have_last_d = true;
        c->d_generation++;
    });
}
int pb_render_d_get_state(pb_ctx *c, uint64_t *state) {
    return guard(c, [&] {
        PB_ASSERT_MSG(state, "Null argument");
        PB_ASSERT_MSG(c->have_last_d, "pb_render_d_get_state needs a preceding pb_render_d on the configured scene");
        state[0] = c->last_d_offset; state[1] = c->last_d_offset_e; state[2] = c->last_d_offset_s; state[3] = c->d_generation;
    });
}
int pb_render_d_set_state(pb_ctx *c, const uint64_t *state) {
    return guard(c, [&] {
        PB_ASSERT_MSG(state, "Null argument");
        PB_ASSERT_MSG(c->ready, "Input scene must be configured!");
        PB_ASSERT_MSG(state[3] > c->d_generation_at_configure && state[3] <= c->d_generation, "pb_render_d_set_state: this state is not from a pb_render_d on the scene as configured now");
        if (state[3] != c->d_generation) c->retained_valid = false;   // the retained records belong to a later render: the VJP re-traces
        c->last_d_offset = state[0]; c->last_d_offset_e = state[1]; c->last_d_offset_s = state[2];
        c->have_last_d = true;
    });
}

int pb_preprocess_secondary_edges(pb_ctx *c, int sensor, const int *resolution, int nrounds) {
    return guard(c, [&] { PB_ASSERT_MSG(resolution, "Null argument"); preprocess_secondary_edges(c, sensor, resolution, nrounds); });
}

int pb_sample_boundary_segment_direct(pb_ctx *c, int64_t n, const float *d_sample3, float *d_out) {
    return guard(c, [&] {
        PB_ASSERT_MSG(c->ready, "Scene needs to be configured!");
        PB_ASSERT_MSG(n >= 0 && n <= std::numeric_limits<int>::max() && (n == 0 || (d_sample3 && d_out)), "Invalid arguments");
        PB_ASSERT_MSG(c->sppse > 0 && c->num_sec > 0, "sample_boundary_segment_direct needs sppse > 0 (the secondary-edge table is built by configure only then, scene.cpp:205) and at least one edge");
        PB_ASSERT_MSG(!c->emitters.empty(), "No Emitter!");
        PB_CUDA(cudaSetDevice(c->device));
        const EdgeParams Q = make_edge_params(c, 0, nullptr);
        launch_sample_boundary_segment(c->stream, (int)n, c->view, Q, d_sample3, d_out);
        c->launches++;
        PB_CUDA(cudaGetLastError());
        PB_CUDA(cudaStreamSynchronize(c->stream));
    });
}

int pb_grad_require(pb_ctx *c, int kind, int id, int slot, int enable) {
    return guard(c, [&] {
        if (kind == PB_PARAM_BSDF_TEXTURE) {
            PB_ASSERT_MSG(id >= 0 && id < (int)c->bsdfs.size() && slot >= 0 && slot < TEX_COUNT, "Invalid BSDF texture");
            c->bsdfs[id].tex[slot].requires_grad = enable != 0;
        } else if (kind == PB_PARAM_MESH_VERTICES) {
            PB_ASSERT_MSG(id >= 0 && id < (int)c->meshes.size(), "Invalid mesh id");
            c->meshes[id].requires_grad = enable != 0;
        } else if (kind == PB_PARAM_MESH_UV) {
            PB_ASSERT_MSG(id >= 0 && id < (int)c->meshes.size() && (c->meshes[id].flags & 2), "Invalid mesh id, or the mesh has no texture coordinates");
            c->meshes[id].uv_requires_grad = enable != 0;
        } else if (kind == PB_PARAM_SENSOR_TRANSFORM) {
            PB_ASSERT_MSG(id >= 0 && id < (int)c->sensors.size(), "Invalid sensor id");
            c->sensors[id].requires_grad = enable != 0;
        } else if (kind == PB_PARAM_ENVMAP_RADIANCE || kind == PB_PARAM_ENVMAP_SCALE || kind == PB_PARAM_ENVMAP_TRANSFORM) {
            PB_ASSERT_MSG(c->emitter_env >= 0, "The scene has no environment map");
            if (kind == PB_PARAM_ENVMAP_RADIANCE) c->emitters[c->emitter_env].env_radiance.requires_grad = enable != 0;
            else if (kind == PB_PARAM_ENVMAP_SCALE) c->emitters[c->emitter_env].env_scale_requires_grad = enable != 0;
            else c->emitters[c->emitter_env].env_xf_requires_grad = enable != 0;
        } else throw Error("Unknown parameter kind");
        c->ready = false;
    });
}
int pb_grad_num_segments(pb_ctx *c) { return (int)c->grad_segments.size(); }
int pb_grad_segment(pb_ctx *c, int index, int *kind, int *id, int *slot, int64_t *offset, int64_t *count) {
    return guard(c, [&] {
        PB_ASSERT_MSG(index >= 0 && index < (int)c->grad_segments.size(), "Invalid segment");
        const GradSegment &g = c->grad_segments[index];
        *kind = g.kind; *id = g.id; *slot = g.slot; *offset = g.offset; *count = g.count;
    });
}
static bool bsdf_slot_is_differentiable(int type, int slot) {
    return type == PB_BSDF_DIFFUSE ? slot == PB_TEX_REFLECTANCE : (slot >= PB_TEX_ALPHA_U && slot <= PB_TEX_SPECULAR_REFLECTANCE);
}
int64_t pb_grad_size(pb_ctx *c) { return c->grad_segments.empty() ? 0 : c->grad_segments.back().offset + c->grad_segments.back().count; }
int pb_render_d_vjp(pb_ctx *c, const pb_integrator *I, int sensor, const float *d_dLdI, float *d_grad) {
    return guard(c, [&] {
        PB_ASSERT_MSG(I && d_dLdI && d_grad, "Null argument");
        PB_ASSERT_MSG(c->have_last_d, "pb_render_d_vjp needs a preceding pb_render_d on the configured scene");
        PB_ASSERT_MSG(I->kind != PB_INTEG_FIELD || I->field == PB_FIELD_SILHOUETTE,
                      "FieldExtractionIntegrator: only the silhouette field (zero interior derivative) has gradients so far");
        for (const GradSegment &g : c->grad_segments)
            PB_ASSERT_MSG(g.kind != PB_PARAM_BSDF_TEXTURE || bsdf_slot_is_differentiable(c->bsdfs[g.id].type, g.slot),
                          "pb_render_d_vjp: this texture is not a parameter of its BSDF type (diffuse: reflectance; roughconductor: alpha_u, alpha_v, eta, k, specular_reflectance)");
        render_interior(c, *I, sensor, nullptr, MODE_VJP, d_dLdI, d_grad);
    });
}

int pb_debug_set(pb_ctx *c, const char *key, int64_t value) {
    return guard(c, [&] {
        if (std::strcmp(key, "sort_mode") == 0) pb::g_sort_mode = (int)value;
        else if (std::strcmp(key, "shade_tune") == 0) pb::g_shade_tune = (int)value;
        else if (std::strcmp(key, "shade_simple") == 0) pb::g_shade_simple = (int)value;
        else if (std::strcmp(key, "adjoint_lin") == 0) pb::g_adjoint_lin = (int)value;
        else if (std::strcmp(key, "l2_persist") == 0) { c->l2_persist = (int)value; set_l2_window(c); }
        else if (std::strcmp(key, "rng_seed_table") == 0) { c->rng_seed_table = (int)value; c->rng_seed_count = 0; }
        else if (std::strcmp(key, "trace_kernel") == 0) pb::g_trace_kernel = (int)value;
        else if (std::strcmp(key, "trace_node_min") == 0) pb::g_trace_node_min = (int)value;
        else if (std::strcmp(key, "pipeline") == 0) c->pipeline = (int)value;
        else if (std::strcmp(key, "pipeline_max_lanes") == 0) c->pipeline_max_lanes = value;
        else if (std::strcmp(key, "trace_chunk") == 0) pb::g_trace_chunk = (int)value;
        else if (std::strcmp(key, "trace_blocks") == 0) pb::g_trace_blocks = (int)value;
        else if (std::strcmp(key, "bvh_builder") == 0) { c->bvh_builder = (int)value; c->bvh_valid = false; }   // = pb_ctx_set_bvh_builder (bench.py --debug)
        else if (std::strcmp(key, "lbvh_leaf") == 0) { pb::g_lbvh_leaf_max = (int)value; c->bvh_valid = false; }
        else if (std::strcmp(key, "sorted_copy") == 0) { c->sorted_copy = (int)value; c->retained_valid = false; }
        else throw Error(std::string("Unknown debug key: ") + key);
    });
}
/* scratch ray buffer (the rays of the last event traced by the last batch): (nb+nl)*n RayRecs */
int pb_debug_ray_buffer(pb_ctx *c, int event, void **d_rays, int64_t *bytes) {
    return guard(c, [&] {
        (void)event;
        EventStore &S = c->retained_valid ? c->retained : c->scratch;
        *d_rays = S.rays.p; *bytes = (int64_t)S.rays.bytes;
    });
}
/* retained per-lane radiance of the last pb_render_d (float4 per lane), for debugging */
int pb_debug_retained_rad(pb_ctx *c, void **d_rad, int64_t *bytes) {
    return guard(c, [&] {
        PB_ASSERT_MSG(c->retained_valid, "nothing retained");
        *d_rad = c->retained.rad.p; *bytes = (int64_t)c->retained.rad.bytes;
    });
}
int pb_ctx_set_bvh_refit(pb_ctx *c, int max_consecutive_refits) { c->bvh_max_refits = max_consecutive_refits; return 0; }
int pb_ctx_set_bvh_builder(pb_ctx *c, int builder) {
    return guard(c, [&] {
        PB_ASSERT_MSG(builder == PB_BVH_HOST_SAH || builder == PB_BVH_DEVICE_LBVH, "Unknown BVH builder");
        c->bvh_builder = builder; c->bvh_valid = false;
    });
}
int pb_stats_bvh(pb_ctx *c, int *builds, int *refits) { if (builds) *builds = c->bvh_builds; if (refits) *refits = c->bvh_refit_count; return 0; }
int pb_ctx_set_retain_limit(pb_ctx *c, int64_t bytes) { c->retain_limit = bytes; c->retained_valid = false; return 0; }
int pb_render_d_jvp(pb_ctx *c, const pb_integrator *I, int sensor, const float *d_tangent, float *d_dimage) {
    return guard(c, [&] {
        PB_ASSERT_MSG(I && d_tangent && d_dimage, "Null argument");
        PB_ASSERT_MSG(c->have_last_d, "pb_render_d_jvp needs a preceding pb_render_d on the configured scene");
        PB_ASSERT_MSG(I->kind != PB_INTEG_FIELD || I->field == PB_FIELD_SILHOUETTE,
                      "FieldExtractionIntegrator: only the silhouette field (zero interior derivative) has derivatives so far");
        for (const GradSegment &g : c->grad_segments)
            PB_ASSERT_MSG(g.kind != PB_PARAM_BSDF_TEXTURE || bsdf_slot_is_differentiable(c->bsdfs[g.id].type, g.slot),
                          "pb_render_d_jvp: this texture is not a parameter of its BSDF type (diffuse: reflectance; roughconductor: alpha_u, alpha_v, eta, k, specular_reflectance)");
        render_interior(c, *I, sensor, d_dimage, MODE_JVP, nullptr, const_cast<float *>(d_tangent));
    });
}
int64_t pb_stats_launches(pb_ctx *c) { return c->launches; }
float pb_stats_last_trace_ms(pb_ctx *c) { return c->last_trace_ms; }
int64_t pb_stats_last_rays(pb_ctx *c) { return c->last_rays; }
int64_t pb_stats_last_active_rays(pb_ctx *c) { return c->last_active_rays; }
float pb_stats_last_primary_ms(pb_ctx *c) { return c->last_primary_ms; }
int pb_stats_last_trace_launches(pb_ctx *c) { return c->last_trace_launches; }
int pb_ctx_set_stream(pb_ctx *c, void *stream) {
    return guard(c, [&] {
        PB_CUDA(cudaSetDevice(c->device));
        PB_CUDA(cudaStreamSynchronize(c->stream));
        if (c->own_stream) { cudaStreamDestroy(c->stream); c->own_stream = false; }
        c->stream = reinterpret_cast<cudaStream_t>(stream);
        if (c->ready) set_l2_window(c);
    });
}

}  // extern "C"
