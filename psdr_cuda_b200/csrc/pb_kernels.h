// psdr-b200: host-callable launchers of the CUDA kernels (internal to the shared library; the public surface is
// include/psdr_b200.h).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "pb_scene.cuh"

namespace pb {

// One scattering event of the interior integral. Direct(b,l) is a single event with (nb,nl) = (b,l);
// the path integrator runs `max_depth` events with (1,1) and continues along BSDF ray 0.
struct BounceParams {
    int nb, nl;             // BSDF-sampled / emitter-sampled connections at this event (direct.cpp:67,119)
    int depth;              // event index (0 = first hit of the camera ray)
    int last;               // 1: this event writes the film
    int carry;              // 1: path mode, keep throughput/radiance state between events
    int hide_emitters;      // direct.h m_hide_emitters
    int ad;                 // 1: the reference's AD formulation of the primal (direct.cpp:83-95), 0: renderC's
    int rc_grad;            // adjoint only: some rough-conductor texture requires a gradient
    RngJump jump;           // stream position of this event's first draw
    float3 sort_lo, sort_inv_ext;   // scene box of the ray sort (k_shade leaves each emitted ray's sort key in EventBuffers::keys)
    int sort_mode;
};

struct RenderParams {
    SceneView S;
    SensorRec cam;
    int width, height, spp;
    float inv_spp;
    long long local0;       // shard-local index of the first lane of this batch
    int spp_local, s0;      // this shard owns samples [s0, s0 + spp_local) of every pixel (sample sharding; all of them under pixel sharding)
    int tile_rows, rank, world;   // pixel sharding: tile_rows > 0 and this shard owns the image-row tiles t * world + rank (tile_rows rows each); 0 = every pixel
    int n;                  // lanes in this batch
    RngJump jump0;          // stream position of the pixel jitter
    const ulonglong2 *rng_seed;   // seeded (state, inc) per global lane id, lanes [0, rng_seed_count) (pb_ctx; nullptr / short: hash on the fly)
    long long rng_seed_count;
};

// device buffers of one scattering event for one batch of n lanes (R = nb + nl rays per lane)
struct EventBuffers {
    const HitRec *hit_cur;   // [n]   hit that created this event's vertex
    const HitRec *hit_prev;  // [n]   adjoint only: hit that created the previous vertex (null at depth 0)
    const float4 *prev_pos;  // [n]   position of the previous vertex (origin of the ray that found hit_cur); unused at depth 0
    float4 *pos;             // [n]   this vertex' record, part a = (p, wi.x). The record (48 B: position, shading normal, local incident
    float4 *vb, *vc;         // [n]   direction, uv, mesh id; b = (n_sh, wi.y), c = (uv, mesh id or -1, wi.z)) is written by k_shade, which reconstructs
                             //       the vertex from its hit, and read by k_resolve / k_adjoint_lin instead of reconstructing it again
    RayRec *rays;            // [R*n] rays of this event (scratch, ray j of lane i at j*n + i)
    HitRec *hits;            // [R*n] their hits
    const float4 *thr_in;    // [n]   throughput T_k (.w != 0: the path is dead); unused at depth 0
    float4 *thr_out;         // [n]   T_{k+1}; may be null on the last event
    float4 *rad;             // [n]   radiance accumulated so far (in/out)
    float2 *conn;            // [R*n] diffuse scenes: what k_resolve needs of the samples k_shade drew, so that it neither re-derives the sampler nor
                             //       repeats the sampling — BSDF ray: (cos theta_o of the sampled direction, -), emitter ray: (position pdf, squared distance);
                             //       the directions themselves are read back from `rays`
    unsigned short *keys;    // [R*n] sort keys of this event's rays, written by k_shade (or null: the sort computes them from the rays)
    const unsigned *inv_cur; // [n]   sorted-copy traversal: hit_cur is in stream order and inv_cur[i] is lane i's position (null: hit_cur is indexed by lane)
    const unsigned *inv;     // [R*n] sorted-copy traversal: `hits` is in stream order and inv[j*n + i] is the position of ray j of lane i (~0: inactive
                             //       lane, a miss); null: hits are indexed by ray slot
    const int *lane_list;    // adjoint kernels: the lanes this launch works on (k_adjoint behind k_adjoint_split), *lane_count of them; null: all n lanes
    const unsigned *lane_count;
    float4 *lin;             // [n]   retained renders of diffuse scenes: the event's linearisation in its vertex' reflectance, (A, c) with
                             //       L_k = rho * A and w_k = rho * c — written by k_resolve, read by k_adjoint_lin (null: not kept)
};

void launch_mesh_preprocess(cudaStream_t st, int nv, int nf, int face_offset, int mesh_id, int flags, const float *vraw, const Mat4 &to_world,
                            const int *faces, const int *csr_off, const int *csr_face, const float *uvs, const int *uv_faces, float *vworld,
                            float4 *fcross, float *vnormal, TriRec *tri, float *face_area);
void launch_mesh_backward(cudaStream_t st, int nv, int nf, int face_offset, const int *csr_off, const int *csr_slot, const float4 *fcross,
                          const float *vworld, const int *faces, const float *vraw, const Mat4 &to_world, const float *tri_grad,
                          const float *g_world_direct, float *g_nsum, float *g_corner, float *grad_out);
void launch_mesh_tangent(cudaStream_t st, int nv, int nf, int face_offset, const float *vraw, const float *vraw_t, const Mat4 &to_world, const float *vworld,
                         const int *faces, const int *csr_off, const int *csr_face, const float4 *fcross, float *vworld_t, float4 *fcross_t, float *vnormal_t,
                         float *tri_tangent);
void launch_bvh_refit(cudaStream_t st, BvhNode *nodes, float *boxes, const LeafTri *leaf, const int *level_off, int num_levels, float extent);
// device-side tables of Scene::configure (pb_tables.cu)
void launch_seq_cmf(cudaStream_t st, long long n, const float *pmf, float *cmf, float *sum_out, const int *n_dev);
void launch_primary_edge_table(cudaStream_t st, int n, const void *edge_src, const SceneView &S, const float *const *vworld, float3 cam, const Mat4 &w2s,
                               unsigned char *flags, int *local, int *tile_sum, int *mesh_kept, int num_meshes, PrimEdgeRec *recs, float *pmf, float *cmf,
                               int *count_out, float *sum_out);
void launch_secondary_edge_table(cudaStream_t st, int n, const void *edge_src, const SceneView &S, const float *const *vworld, unsigned char *flags, int *local,
                                 int *tile_sum, SecEdgeRec *recs, float *pmf, float *cmf, int *count_out, float *sum_out, int importance);
void launch_envmap_pmf(cudaStream_t st, int rx, int ry, int w, int h, const float *texel, const float *sin_theta, float *pmf);
void launch_tri_bounds(cudaStream_t st, int n, const TriRec *tri, float *lohi);
void launch_nodes_to_compact(cudaStream_t st, int num_nodes, const BvhNode *nodes, BvhNodeC *out, float extent);
void launch_rng_seed(cudaStream_t st, long long n, ulonglong2 *out);
void launch_build_leaf_tris(cudaStream_t st, int n, const int *order, const TriRec *tri, LeafTri *leaf);

extern int g_sort_mode;       // debug: sort key (pb_sort.cu)
extern int g_lbvh_leaf_max;
extern int g_adjoint_lin;     // debug: 0 = reflectance adjoints through k_adjoint (connection by connection) even when the linearisation was kept
extern int g_shade_simple;    // debug: 0 = never use the diffuse + area-light instantiations
extern int g_shade_tune;      // debug: k_resolve / k_adjoint variant (0 default)
extern int g_trace_kernel;    // sorted-wavefront traversal kernel: 3 persistent streaming kernel (default), 1 one ray per thread
extern int g_trace_chunk, g_trace_blocks;   // debug knobs of the streaming kernel
extern int g_trace_node_min;  // streaming kernel: node steps continue while at least this many lanes descend
void launch_trace(cudaStream_t st, const SceneView &S, long long n, const RayRec *rays, HitRec *hits, float *t_out);
void launch_trace_sorted(cudaStream_t st, const SceneView &S, long long n, const RayRec *rays, HitRec *hits, float3 lo, float3 hi, unsigned *hist, unsigned *perm,
                         unsigned short *keys, unsigned *stream_counter, unsigned long long *active_total, cudaEvent_t ev0, cudaEvent_t ev1, int mode, bool keys_ready,
                         RayRec *sorted = nullptr, unsigned *inv = nullptr);
struct DevBuf;
bool lbvh_build(cudaStream_t st, int n, const TriRec *tri, const float *scene_lo, const float *scene_hi, DevBuf &scratch, int *order, BvhNode *nodes,
                std::vector<int> &level_off);
void launch_unpermute_hits(cudaStream_t st, long long n, const unsigned *inv, const HitRec *sorted_hits, HitRec *hits);
void launch_primary(cudaStream_t st, const RenderParams &P, HitRec *hit0);
void launch_shade(cudaStream_t st, const RenderParams &P, const BounceParams &B, const EventBuffers &E);
void launch_resolve(cudaStream_t st, const RenderParams &P, const BounceParams &B, const EventBuffers &E, float *film);
// split_list / split_count: scratch (n ints, one counter) for the lane list of the split geometry adjoint, or null
void launch_adjoint(cudaStream_t st, const RenderParams &P, const BounceParams &B, const EventBuffers &E, float4 *suffix, const float *dLdI,
                    int *split_list = nullptr, unsigned *split_count = nullptr);
void launch_edge_primary_rays(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, int side, RayRec *rays);
void launch_edge_primary_grad(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, const float4 *rad_p, const float4 *rad_n, const float *dLdI, float inv_sppe);
void launch_edge_secondary_rays(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, RayRec *rays, int guide_spc);
void launch_edge_secondary_camera(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, const RayRec *rays, const HitRec *hits, RayRec *cam_rays, int guide_spc);
void launch_edge_secondary_eval(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, const RayRec *rays, const HitRec *hits, const RayRec *cam_rays,
                                const HitRec *cam_hits, const float *dLdI, float inv_sppse, float *guide_out, int guide_spc);
void launch_sample_boundary_segment(cudaStream_t st, int n, const SceneView &S, const EdgeParams &Q, const float *sample3, float *out);
void launch_field(cudaStream_t st, const RenderParams &P, int field, const HitRec *hit0, float *film, float4 *rad_out);

}  // namespace pb
