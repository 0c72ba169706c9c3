// psdr-b200: host-side context behind the C ABI (scene description, device tables, wavefront buffers).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "pb_bvh.h"
#include "pb_kernels.h"

namespace pb {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

#define PB_CUDA(call)                                                                                      \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) throw pb::Error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call); \
    } while (0)
#define PB_ASSERT_MSG(cond, msg) do { if (!(cond)) throw pb::Error(msg); } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    ~DevBuf() { if (p) cudaFree(p); }
    void reserve(size_t n) {   // grow-only
        if (n <= bytes) return;
        if (p) { cudaFree(p); p = nullptr; bytes = 0; }
        PB_CUDA(cudaMalloc(&p, n ? n : 16));
        bytes = n;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
    template <class T> void upload(const std::vector<T> &v, cudaStream_t st) {
        reserve(v.size() * sizeof(T));
        if (!v.empty()) PB_CUDA(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
    }
};

struct Mat4h {   // row-major
    float m[16];
    static Mat4h identity() { Mat4h r; for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.f : 0.f; return r; }
};
Mat4h matmul(const Mat4h &a, const Mat4h &b);
Mat4h inverse(const Mat4h &a);

struct HostTexture {
    int w = 1, h = 1, c = 3;
    std::vector<float> data;
    DevBuf d;
    bool dirty = true;
    bool requires_grad = false;
};
struct HostBsdf {
    int type = 0;
    HostTexture tex[TEX_COUNT];
};
struct HostMesh {
    int nv = 0, nf = 0, flags = 0, bsdf = -1, emitter = -1;
    std::vector<float> verts, uvs;
    std::vector<int> faces, uv_faces, edges, csr_off, csr_face, csr_slot;
    Mat4h raw = Mat4h::identity(), left = Mat4h::identity(), right = Mat4h::identity();
    bool verts_host_stale = false;   // the vertices were last set from device memory (pb_scene_set_mesh_vertices_device)
    bool verts_dirty = true, topo_dirty = true, requires_grad = false, uv_requires_grad = false, uv_dirty = false;
    DevBuf d_vraw, d_faces, d_uvs, d_uv_faces, d_csr_off, d_csr_face, d_csr_slot, d_vworld, d_fcross, d_vnormal, d_face_area, d_face_cmf;
    DevBuf d_gworld, d_gnsum, d_gcorner, d_fcross_t, d_vnormal_t;   // VJP scratch: direct world-space vertex adjoint, normal-sum adjoint, per-corner adjoint
    int face_offset = 0;
    float total_area = 0.f, inv_total_area = 0.f, face_sum = 0.f;
    Mat4h to_world = Mat4h::identity();
};
struct HostEmitter {
    int type = 0, mesh = -1;
    float radiance[3] = {0, 0, 0};
    float sampling_weight = 0.f;
    // environment map (envmap.h)
    HostTexture env_radiance;
    float env_scale = 1.f;
    bool env_scale_requires_grad = false;   // env_radiance.requires_grad covers the texels
    bool env_xf_requires_grad = false;      // the matrix EnvironmentMap.set_transform sets
    Mat4h env_raw = Mat4h::identity(), env_left = Mat4h::identity();
    bool env_dirty = true;
    int env_res[2] = {0, 0};
    float env_sum = 0.f;
    DevBuf d_env_pmf, d_env_cmf;
};
struct HostSensor {
    float fov_x = 45.f, near_clip = 0.1f, far_clip = 1e4f;
    Mat4h to_world = Mat4h::identity();
    bool requires_grad = false;   // Sensor.to_world is a differentiable leaf
    Mat4h c2s = Mat4h::identity();   // camera_to_sample (perspective.cpp:14-17), kept for the pose adjoint
    SensorRec rec;
    // primary-edge list (perspective.cpp:39-111)
    int num_prim = 0;
    float prim_sum = 0.f;
    DevBuf d_prim, d_prim_pmf, d_prim_cmf;
};
struct GuideGrid {   // HyperCubeDistribution3f of DirectIntegrator::m_warpper (direct.cpp:166-204)
    int res[3] = {0, 0, 0};
    int cells = 0;
    float sum = 0.f;
    bool ready = false;
    DevBuf d_pmf, d_cmf;
};
struct GradSegment { int kind, id, slot; int64_t offset, count; };
// per-event wavefront records, either one batch worth (scratch) or the whole shard (retained for the VJP)
struct EventStore {
    DevBuf hit0, rad, rays;
    std::vector<DevBuf> pos, vb, vc, hits, thr, lin;   // per event slot: vertex record (a = pos, b, c), hits of its rays, throughput, linearisation
    void release() { hit0.release(); rad.release(); rays.release(); pos.clear(); vb.clear(); vc.clear(); hits.clear(); thr.clear(); lin.clear(); }
};

}  // namespace pb

struct pb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    // description
    int width = 0, height = 0, spp = 0, sppe = 0, sppse = 0;
    std::vector<pb::HostSensor> sensors;
    std::vector<pb::HostBsdf> bsdfs;
    std::vector<pb::HostMesh> meshes;
    std::vector<pb::HostEmitter> emitters;
    int emitter_env = -1;          // index of the environment emitter (scene.h m_emitter_env)
    bool has_bound_mesh = false;   // the last mesh is the envmap's bounding box (scene.cpp:135-180)
    // sharding / tiling
    int rank = 0, world = 1;
    int edge_importance = 0;             // secondary-edge pmf: 0 length (reference), 1 length x dihedral angle (scene.cpp:230-233, `#if 0` there)
    int shard_mode = 0, tile_rows = 0;   // 0: every shard renders spp / world samples of every pixel; 1: image-row tiles of tile_rows rows, dealt round-robin
    void *nccl_comm = nullptr;           // pb_dist.cpp: communicator of pb_dist_init / pb_dist_adopt_comm
    bool nccl_owned = false;
    int64_t batch = 1 << 25;   // 32 Mi lanes: large wavefronts sort into more coherent bins
    // samplers (scene.cpp:65-79): lane count the streams were seeded for and draws consumed so far
    pb::DevBuf d_rng_seed;          // seeded PCG32 (state, inc) per lane id (16 B each), shared by the three samplers
    int64_t rng_seed_count = 0;
    int rng_seed_table = 1;          // debug: 0 = every kernel hashes its lane's seed on the fly (round 1)
    int64_t sampler_count[3] = {0, 0, 0};
    uint64_t sampler_offset[3] = {0, 0, 0};
    // configured device tables
    bool ready = false;
    int num_tri = 0;
    // triangle table | BVH nodes | leaf triangles live in ONE allocation so that a single L2 access-policy window can keep the
    // randomly gathered scene data resident while the wavefront streams through (pb_capi.cu: set_l2_window)
    pb::DevBuf d_scene_arena;
    size_t arena_tri_bytes = 0, arena_node_bytes = 0, arena_leaf_bytes = 0, arena_used = 0;
    pb::TriRec *arena_tri() const { return d_scene_arena.as<pb::TriRec>(); }
    pb::BvhNode *arena_nodes() const { return reinterpret_cast<pb::BvhNode *>(static_cast<char *>(d_scene_arena.p) + arena_tri_bytes); }
    pb::LeafTri *arena_leaf() const { return reinterpret_cast<pb::LeafTri *>(static_cast<char *>(d_scene_arena.p) + arena_tri_bytes + arena_node_bytes); }
    pb::BvhNodeC *arena_nodes_c() const { return reinterpret_cast<pb::BvhNodeC *>(static_cast<char *>(d_scene_arena.p) + arena_tri_bytes + arena_node_bytes + arena_leaf_bytes); }
    int l2_persist = 1;
    pb::DevBuf d_order, d_meshes, d_bsdfs, d_emitters, d_emitter_cmf, d_emitter_pmf;
    // host copies of the small device tables (configure) and their gradient-pointer variants (VJP): persistent, so that the
    // asynchronous uploads never read freed memory and the VJP never reads the tables back
    std::vector<pb::BsdfRec> h_bsdfs, h_bsdfs_grad;
    std::vector<pb::MeshRec> h_meshrecs, h_meshrecs_grad;
    std::vector<pb::EmitterRec> h_emitters, h_emitters_grad;
    std::vector<float> h_tri;   // host copy of the triangle table, fetched on demand (host BVH build, inspection)
    bool h_tri_valid = false;
    // device-side configure (pb_tables.cu): edge topology of all meshes with enabled edges + scratch
    pb::DevBuf d_edge_src, d_edge_flags, d_edge_local, d_edge_tiles, d_edge_out, d_bounds, d_area_sums, d_env_sin;
    int num_edge_src = 0;
    std::vector<int> edge_sig;
    float emitter_sum = 0.f;
    pb::SceneView view;
    // wavefront buffers
    pb::DevBuf d_suffix, d_bsdfs_grad, d_tri_grad, d_tri_tangent, d_jvp_acc, d_sort_hist, d_sort_perm, d_sort_keys, d_stream_counter, d_active_total, d_emitters_grad, d_sensor_acc, d_env_xf_acc, d_meshes_grad;
    float scene_lo[3] = {0, 0, 0}, scene_hi[3] = {1, 1, 1};
    float env_lower[3] = {0, 0, 0}, env_upper[3] = {1, 1, 1};
    // boundary terms
    int num_sec = 0;
    float sec_sum = 0.f;
    pb::DevBuf d_sec, d_sec_pmf, d_sec_cmf, d_mesh_vworld, d_mesh_gworld, d_edge_rays, d_edge_rad;
    std::vector<pb::GuideGrid> guides;
    uint64_t last_d_offset_e = 0, last_d_offset_s = 0;
    pb::EventStore scratch, retained;
    // Second batch in flight (render_interior): short wavefronts (one GPU's share of a multi-GPU job) end in a tail as long as their
    // longest ray, 10 % of a 12 M-ray launch; two half-batches on two streams fill each other's tails. Lane 0 is the context's own
    // stream and buffers, lane 1 owns copies of everything a batch scribbles on.
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    pb::EventStore scratch1;
    pb::DevBuf d_rays1, d_suffix1, d_sort_hist1, d_sort_perm1, d_sort_keys1, d_stream_counter1;
    // sorted-copy traversal (pb_sort.cu k_sort_scatter<COPY>), per lane: the rays in stream order, ray slot -> stream position, hits in stream order
    pb::DevBuf d_sorted_rays[2], d_sort_inv[2], d_sorted_hits[2];
    pb::DevBuf d_adj_list[2], d_adj_count[2];   // lane list of the split geometry adjoint (k_adjoint_split), per lane
    pb::DevBuf d_conn[2];                // EventBuffers::conn of the batch in flight on each lane
    int bvh_builder = 0;                 // first build of the scene BVH: 0 = binned SAH on the host (pb_bvh.cpp), 1 = LBVH on the device (pb_lbvh.cu)
    pb::DevBuf d_lbvh_scratch;
    int sorted_copy = 0;                 // 1: the interior events trace a sorted copy of their rays (debug key sorted_copy; A/B in profiles/)
    bool retained_hits_by_slot = true;   // the retained hit records are indexed by ray slot (false: stream order — only k_adjoint_lin can use the store)
    int pipeline = 1;                       // 0 never, 1 when a render has at most pipeline_max_lanes lanes, 2 always (debug)
    int64_t pipeline_max_lanes = 20 << 20;
    int64_t retain_limit = (int64_t)64 << 30;
    int64_t retained_B = 0;   // batch size of the render that filled the retained store (its layout depends on it)
    bool retained_valid = false;
    int retained_kind = 0, retained_nb = 0, retained_nl = 0, retained_nbounce = 0, retained_sensor = 0, retained_hide = 0;
    // replay info of the last renderD
    uint64_t last_d_offset = 0;
    bool have_last_d = false;
    uint64_t d_generation = 0, d_generation_at_configure = 0;   // pb_render_d calls so far / at the last configure (pb_render_d_{get,set}_state)
    std::vector<pb::GradSegment> grad_segments;
    // stats
    int64_t launches = 0, last_rays = 0, last_active_rays = 0, collectives = 0;
    // BVH refit (vertex-only updates keep the tree and recompute its boxes on the device)
    bool bvh_valid = false;
    int bvh_max_refits = 16, bvh_refits = 0, bvh_builds = 0, bvh_refit_count = 0;
    std::vector<int> bvh_sig, bvh_level_off;
    pb::DevBuf d_node_boxes;
    float last_trace_ms = 0.f, last_primary_ms = 0.f;
    int last_trace_launches = 0;
    bool own_stream = true;
    std::vector<cudaEvent_t> ev_pool;
};
