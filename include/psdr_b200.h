/* psdr_b200.h — C ABI of the B200-native path-space differentiable rendering hot path.
 *
 * This is the drop-in boundary for psdr-cuda's `Integrator::renderC / renderD` path. The reference has no C ABI or
 * FFI of its own: its boundary is the pybind11 module `psdr_cuda` (src/psdr.cpp:41-295). Every entry point below
 * cites the reference interface it stands behind; the pybind11 host layer of this repo (psdr_cuda_b200/host) binds
 * exactly these, and INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions: every call returns 0 on success and a non-zero code on failure (never throws across the ABI);
 * pb_last_error(ctx) returns the message (the reference raises psdr::Exception -> Python RuntimeError,
 * include/misc/Exception.h:13-107). Pointers named h_* are host memory, d_* are device memory on the context's GPU.
 * The context owns scene tables, BVH and wavefront buffers; the caller owns every image / gradient buffer it passes.
 * One context per host thread and GPU; no global state. All work is enqueued on the context's stream and the call
 * returns after the stream has been synchronised unless stated otherwise.
 */
#ifndef PSDR_B200_H
#define PSDR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb_ctx pb_ctx;

/* BSDF kinds / texture slots: src/scene/scene_loader.cpp:318-351 (diffuse, roughconductor) */
enum { PB_BSDF_DIFFUSE = 0, PB_BSDF_ROUGHCONDUCTOR = 1 };
enum { PB_TEX_REFLECTANCE = 0, PB_TEX_ALPHA_U = 1, PB_TEX_ALPHA_V = 2, PB_TEX_ETA = 3, PB_TEX_K = 4, PB_TEX_SPECULAR_REFLECTANCE = 5 };
/* integrators: src/psdr.cpp:282-294 (+ the multi-bounce PathIntegrator BASELINE.json names; depth 1 == Direct(1,1)) */
enum { PB_INTEG_DIRECT = 0, PB_INTEG_FIELD = 1, PB_INTEG_PATH = 2 };
/* FieldExtractionIntegrator fields: src/integrator/field.cpp:38-52 */
enum { PB_FIELD_SILHOUETTE = 0, PB_FIELD_POSITION, PB_FIELD_DEPTH, PB_FIELD_GEONORMAL, PB_FIELD_SHNORMAL, PB_FIELD_UV };
/* mesh flags: Mesh::m_use_face_normals / m_enable_edges (include/psdr/shape/mesh.h) */
enum { PB_MESH_FACE_NORMALS = 1, PB_MESH_ENABLE_EDGES = 2 };
/* differentiable leaves: what a gradient segment refers to */
enum { PB_PARAM_BSDF_TEXTURE = 0, PB_PARAM_MESH_VERTICES = 1, PB_PARAM_ENVMAP_RADIANCE = 2, PB_PARAM_ENVMAP_SCALE = 3, PB_PARAM_SENSOR_TRANSFORM = 4 /* Sensor.to_world, 16 floats row-major, id = sensor; src/psdr.cpp:220-224 */,
       PB_PARAM_MESH_UV = 6 /* Mesh.vertex_uv, 2 floats per uv vertex in the order given to pb_scene_add_mesh; src/psdr.cpp:254 */,
       PB_PARAM_ENVMAP_TRANSFORM = 5 /* the matrix EnvironmentMap.set_transform sets (to_world = left * raw), 16 floats; src/psdr.cpp:238 */ };   /* EnvironmentMap.radiance.data / .scale, src/psdr.cpp:236-237 (id, slot ignored) */

typedef struct pb_integrator {
    int kind;           /* PB_INTEG_* */
    int bsdf_samples;   /* DirectIntegrator(bsdf_samples, light_samples): src/integrator/direct.cpp:30-32 */
    int light_samples;
    int hide_emitters;  /* DirectIntegrator.hide_emitters: src/psdr.cpp:294 */
    int field;          /* PB_FIELD_* for PB_INTEG_FIELD */
    int max_depth;      /* PB_INTEG_PATH: number of scattering events */
    int use_guiding;    /* use the grid built by pb_preprocess_secondary_edges for this sensor (DirectIntegrator::m_warpper) */
} pb_integrator;

/* ---- context ------------------------------------------------------------------------------------------------ */
int pb_ctx_create(int device, pb_ctx **out);
int pb_ctx_destroy(pb_ctx *ctx);
const char *pb_last_error(pb_ctx *ctx);            /* ctx may be NULL: message of the last failed pb_ctx_create */
int pb_version(void);
/* lanes per wavefront batch (multiple of 1024); 0 restores the default (2^25). Tiling replaces the reference's
 * "all W*H*spp lanes at once" (src/integrator/integrator.cpp:69-76). */
int pb_ctx_set_batch(pb_ctx *ctx, int64_t lanes);
/* multi-GPU: this context renders shard `rank` of `world`: samples [spp*rank/world, spp*(rank+1)/world) of every pixel
 * (lane ranges for the edge terms). RNG streams stay indexed by the global lane id, so the sum of the shards' images
 * equals the single-GPU image up to fp32 summation order. Default (0, 1). */
int pb_ctx_set_shard(pb_ctx *ctx, int rank, int world);
/* How the interior term is split over ranks. PB_SHARD_SAMPLES (default): every rank renders spp / world samples of every pixel (perfect
 * load balance; the films are partial sums). PB_SHARD_PIXELS: rank r renders every sample of the image-row tiles t * world + r of
 * `tile_rows` rows (0 = height / world: one contiguous block per rank; 1 = interleaved rows); the films are disjoint and their union is
 * bit-identical to the single-GPU image. The boundary terms are always split by lane range (integrator.cpp:98-119). */
enum { PB_SHARD_SAMPLES = 0, PB_SHARD_PIXELS = 1 };
int pb_ctx_set_shard_mode(pb_ctx *ctx, int mode, int tile_rows);
/* pb_render_d keeps every event's hit records / vertex records / throughputs (and, for diffuse scenes, the event's linearisation in
 * the vertex' reflectance) of the whole shard (32 + 112 B per event: 592 B per lane for a depth-5 path) when they fit in `bytes` (default 64 GiB of the 180 GB HBM3e), so that pb_render_d_vjp runs only the
 * adjoint kernels; otherwise the VJP re-traces the forward pass batch by batch. 0 disables retention. */
int pb_ctx_set_retain_limit(pb_ctx *ctx, int64_t bytes);
/* Run on the caller's CUDA stream (a cudaStream_t; NULL = the legacy default stream) instead of the context's own.
 * STREAM CONTRACT: every kernel and copy of a call is enqueued on the context's stream and the call returns after synchronising it. A fresh
 * context owns a private cudaStreamNonBlocking stream, which is NOT ordered against work the caller has in flight elsewhere: buffers passed
 * in (d_image, d_dLdI, d_grad, device-pointer setters) must be complete before the call, or — the normal case with torch / CuPy tensors —
 * the caller passes the stream that produces and consumes them here, once or whenever it changes. */
int pb_ctx_set_stream(pb_ctx *ctx, void *cuda_stream);

/* ---- scene description: what SceneLoader::load_scene builds (src/scene/scene_loader.cpp:208-242) ------------- */
/* RenderOption: include/psdr/types.h:171-182, src/psdr.cpp:53-72 */
int pb_scene_set_options(pb_ctx *ctx, int width, int height, int spp, int sppe, int sppse);
/* PerspectiveCamera(fov_x, near, far) + to_world (row-major 4x4): src/sensor/perspective.cpp, src/psdr.cpp:222-224. Returns id >= 0. */
int pb_scene_add_sensor(pb_ctx *ctx, float fov_x, float near_clip, float far_clip, const float h_to_world[16]);
int pb_scene_set_sensor_transform(pb_ctx *ctx, int sensor, const float h_to_world[16]);
/* Diffuse / RoughConductor with their default textures (src/bsdf/*.h). Returns id >= 0. */
int pb_scene_add_bsdf(pb_ctx *ctx, int type);
/* Bitmap data of one slot, interleaved [ (y*w + x)*C + c ], C = 1 (alpha_u/v) or 3: src/core/bitmap.cpp, src/psdr.cpp:102-120 */
int pb_scene_set_bsdf_texture(pb_ctx *ctx, int bsdf, int slot, const float *h_data, int w, int h);
/* Mesh::load result (src/shape/mesh.cpp:62-141): object-space vertices (xyz AoS), triangle indices, optional uvs. Returns id >= 0. */
int pb_scene_add_mesh(pb_ctx *ctx, int num_vertices, int num_faces, const float *h_vertices, const int *h_faces, int num_uvs,
                      const float *h_uvs, const int *h_uv_faces, int flags, int bsdf, const float h_to_world[16]);
int pb_scene_set_mesh_vertices(pb_ctx *ctx, int mesh, const float *h_vertices);              /* Mesh.vertex_positions setter */
int pb_scene_set_mesh_uvs(pb_ctx *ctx, int mesh, const float *h_uvs);                        /* Mesh.vertex_uv setter (same count as at pb_scene_add_mesh), src/psdr.cpp:254 */
int pb_scene_set_mesh_transform(pb_ctx *ctx, int mesh, const float h_mat[16], int left);     /* Mesh::set_transform, mesh.h:19-26 */
/* AreaLight(radiance, mesh): src/emitter/area.cpp, src/psdr.cpp:230-231. Returns id >= 0. */
int pb_scene_add_area_emitter(pb_ctx *ctx, int mesh, const float h_radiance[3]);
/* EnvironmentMap(file) + scale + to_world (src/emitter/envmap.cpp, src/psdr.cpp:233-238, scene_loader.cpp:291-315): lat-long RGB
 * radiance, interleaved. Must be added before the area emitters to keep the reference's emitter order. At configure the
 * context appends the 12-triangle bounding mesh the envmap radiates from (src/scene/scene.cpp:135-180). Returns id >= 0. */
int pb_scene_add_envmap(pb_ctx *ctx, int w, int h, const float *h_rgb, float scale, const float h_to_world[16]);
int pb_scene_set_envmap_radiance(pb_ctx *ctx, const float *h_rgb /* w*h*3 of the creation size, or NULL to keep */, float scale);   /* EnvironmentMap.radiance.data / .scale edits, src/psdr.cpp:236-237 */
int pb_scene_set_envmap_transform(pb_ctx *ctx, const float h_left[16]);   /* EnvironmentMap::set_transform, envmap.h:18-21 */
int pb_scene_num_meshes(pb_ctx *ctx);                                      /* Scene.num_meshes (includes the bounding mesh) */
/* Scene::configure: src/scene/scene.cpp:56-278 (sampler seeding rule, mesh preprocessing, emitter pmf, triangle table, BVH) */
int pb_scene_configure(pb_ctx *ctx);
/* forget sampler positions so that the next configure() restarts every stream (a fresh Scene in the reference) */
int pb_scene_reseed(pb_ctx *ctx);
int pb_scene_num_triangles(pb_ctx *ctx);
/* configured tables, for inspection: 22 floats per triangle (p0 e1 e2 n0 n1 n2 face_normal area; types.h:136-146) */
int pb_scene_get_triangle_info(pb_ctx *ctx, float *h_out);
/* the two setters above from DEVICE pointers: an optimiser that keeps its parameters on the GPU updates the scene without a host round trip
 * (the reference's Enoki arrays live on the device too: `mesh.vertex_positions = u`, examples/utils/adam.py:63); the copies are enqueued on the
 * context's stream. Texture resolution stays what it was. pb_scene_get_mesh_vertices reads the current positions back (host buffer, 3 nv floats). */
int pb_scene_set_mesh_vertices_device(pb_ctx *ctx, int mesh, const float *d_verts);
int pb_scene_set_bsdf_texture_device(pb_ctx *ctx, int bsdf, int slot, const float *d_data);
int pb_scene_get_mesh_vertices(pb_ctx *ctx, int mesh, float *h_out);
/* importance of a secondary edge in the edge distribution: 0 = its length (the reference, src/scene/scene.cpp:236), 1 = length x exterior
 * dihedral angle, pi for boundary edges — the alternative the reference keeps under `#if 0` (scene.cpp:230-233). Unbiased either way. */
int pb_scene_set_edge_importance(pb_ctx *ctx, int mode);
/* the edge tables Scene::configure builds (on the device: csrc/pb_tables.cu), copied to the host for inspection: primary edges of a sensor
 * (PrimaryEdgeInfo, src/sensor/perspective.cpp:39-111) as 7 floats each (p0.xy, p1.xy, edge_normal.xy, edge_length), secondary edges
 * (SecondaryEdgeInfo, src/shape/mesh.cpp:251-264, src/scene/scene.cpp:219-235) as 16 floats each (p0, e1, n0, n1, p2, is_boundary);
 * cmf (optional) receives the running sums of the edge lengths the samplers search. Built only when sppe > 0 / sppse > 0. */
int pb_scene_num_primary_edges(pb_ctx *ctx, int sensor);
int pb_scene_get_primary_edges(pb_ctx *ctx, int sensor, float *h_out, float *h_cmf);
int pb_scene_num_secondary_edges(pb_ctx *ctx);
int pb_scene_get_secondary_edges(pb_ctx *ctx, float *h_out, float *h_cmf);
/* unique edges of a mesh, 5 ints each (v0 v1 f0 f1|-1 opposite vertex): src/shape/mesh.cpp:143-203, Mesh.edge_indices() */
int pb_scene_mesh_num_edges(pb_ctx *ctx, int mesh);
int pb_scene_mesh_get_edges(pb_ctx *ctx, int mesh, int *h_out);

/* ---- hot path ------------------------------------------------------------------------------------------------- */
/* Scene_OptiX::ray_intersect (src/scene/scene_optix.cpp:80-126, cuda/psdr_cuda.cu:9-45): n rays as 8 floats each
 * (o.xyz, tmax, d.xyz, t_occ; t_occ = 0 for a closest-hit query, > 0 to stop at any hit closer than t_occ), hits as 4 x 32 bit each (tri id, shape id, u, v).
 * d_t != NULL: rays in lane order through the general kernel (any origin), d_t receives the distance. d_t == NULL: the launch the render
 * calls use (counting sort by direction / origin cell, compaction of inactive rays, persistent streaming kernel); origins inside the scene box. */
int pb_trace(pb_ctx *ctx, int64_t n, const float *d_rays, void *d_hits, float *d_t);
/* Integrator::renderC (src/integrator/integrator.cpp:13-29): d_image receives W*H*3 floats, pixel = y*W + x */
int pb_render_c(pb_ctx *ctx, const pb_integrator *integ, int sensor, float *d_image);
/* same call with host buffers: image copied back inside the call */
int pb_render_c_host(pb_ctx *ctx, const pb_integrator *integ, int sensor, float *h_image);
/* Integrator::renderD primal (src/integrator/integrator.cpp:32-60): the AD formulation's image; remembers the sampler
 * positions so that pb_render_d_vjp replays the same paths */
int pb_render_d(pb_ctx *ctx, const pb_integrator *integ, int sensor, float *d_image);
/* The reference's tape lets a loss depend on several renderD images (one per sensor, src/integrator/integrator.cpp:32-60 called in a
 * loop) before one ek.backward. pb_render_d_vjp / _jvp replay "the last pb_render_d"; these two calls make any earlier one the last
 * again: get_state after a pb_render_d returns {interior, primary-edge, secondary-edge sampler positions, serial number}; set_state
 * before the VJP of that image restores them (records retained for a later render are dropped: that VJP re-traces its paths).
 * A state is valid until the next pb_scene_configure. */
int pb_render_d_get_state(pb_ctx *ctx, uint64_t state[4]);
int pb_render_d_set_state(pb_ctx *ctx, const uint64_t state[4]);

/* DirectIntegrator::preprocess_secondary_edges (src/integrator/direct.cpp:166-204, src/psdr.cpp:285): builds the guiding grid
 * resolution[0..2] cells x resolution[3] samples per cell x nrounds over the secondary-edge sample space of `sensor` */
int pb_preprocess_secondary_edges(pb_ctx *ctx, int sensor, const int resolution[4], int nrounds);

/* Scene::sample_boundary_segment_direct (src/scene/scene.cpp:456-492, bound at src/psdr.cpp:274): n samples of 3 floats -> 17 floats each:
 * p0 (point on a face edge), edge (unit), edge2 (to the opposite vertex), p2 and n (point on an emitter), pdf, is_valid (1 / 0).
 * Needs sppse > 0: like the reference, configure builds the secondary-edge table only then. */
int pb_sample_boundary_segment_direct(pb_ctx *ctx, int64_t n, const float *d_sample3, float *d_out);

/* ---- gradients (reverse mode: ek.backward + ek.gradient in the reference, docs/inverse_diff_render.rst:71-79) -- */
/* mark a leaf as requiring a gradient (ek.set_requires_gradient); slot only for PB_PARAM_BSDF_TEXTURE */
int pb_grad_require(pb_ctx *ctx, int param_kind, int id, int slot, int enable);
/* layout of the flat fp32 gradient vector: number of segments, and per segment (kind, id, slot, offset, count) */
int pb_grad_num_segments(pb_ctx *ctx);
int pb_grad_segment(pb_ctx *ctx, int index, int *kind, int *id, int *slot, int64_t *offset, int64_t *count);
int64_t pb_grad_size(pb_ctx *ctx);
/* VJP of renderD: d_dLdI is W*H*3; accumulates (+=) into d_grad (pb_grad_size floats; the caller zeroes it).
 * Includes the primary-edge term when sppe > 0 and the secondary-edge term when sppse > 0 (src/integrator/integrator.cpp:41-47);
 * both exist only in the derivative, so they are evaluated here and not in pb_render_d.
 * On several GPUs each rank holds a private d_grad and the caller all-reduces it once. */
int pb_render_d_vjp(pb_ctx *ctx, const pb_integrator *integ, int sensor, const float *d_dLdI, float *d_grad);

/* ---- several GPUs (BASELINE.json north_star; the reference is single-GPU): one process and one context per GPU, NCCL resolved at run time.
 * pb_dist_unique_id on rank 0, the 128 bytes handed to every rank (any transport), pb_dist_init everywhere (also sets the shard);
 * or pb_dist_adopt_comm with an existing ncclComm_t. pb_allreduce_grads enqueues ONE ncclAllReduce(sum) of the flat gradient vector on
 * the context's stream, i.e. right behind the last adjoint kernel of pb_render_d_vjp — the single exchange step of renderD
 * (replaces nothing in the reference). pb_allreduce_image completes a film after pb_render_c / pb_render_d: an in-place
 * all-gather for equal contiguous pixel tiles, a sum otherwise. Both are no-ops on one GPU. */
int pb_dist_available(void);
int pb_dist_unique_id(pb_ctx *ctx, char id128[128]);
int pb_dist_init(pb_ctx *ctx, const char id128[128], int rank, int world);
int pb_dist_adopt_comm(pb_ctx *ctx, void *nccl_comm, int rank, int world);
int pb_dist_finalize(pb_ctx *ctx);
int pb_allreduce_grads(pb_ctx *ctx, float *d_grad, int64_t count);
int pb_allreduce_image(pb_ctx *ctx, float *d_image);
int64_t pb_stats_collectives(pb_ctx *ctx);

/* forward mode (ek.forward + ek.gradient(image) in the reference, examples/run_test.py:126-129): d_tangent is a flat vector with the
 * layout of the gradient vector (the direction in parameter space), d_dimage receives the W*H*3 derivative image. Needs a
 * preceding pb_render_d like the VJP. Implemented by running the adjoint kernels once per colour channel with a unit seed and
 * contracting each local gradient with the tangent of what it refers to. */
int pb_render_d_jvp(pb_ctx *ctx, const pb_integrator *integ, int sensor, const float *d_tangent, float *d_dimage);

/* ---- instrumentation ---------------------------------------------------------------------------------------- */
/* kernels launched by this context since creation; for the last render / trace call: milliseconds spent in the k_trace
 * launches (CUDA events on the context's stream), how many launches and rays that was, and the milliseconds of the
 * fused raygen + primary-trace kernel — bench.py's gpu_launches and roofline inputs */
int64_t pb_stats_launches(pb_ctx *ctx);
float pb_stats_last_trace_ms(pb_ctx *ctx);
/* Scene::configure after a vertex-only edit keeps the BVH topology and recomputes its boxes on the device (replaces the per-configure
 * optixAccelBuild of include/psdr/scene/optix.h:277-340); after `max_consecutive_refits` refits (default 16; 0 = always rebuild) the
 * binned-SAH build runs again. Results do not depend on it: the traversal returns the exact closest hit for any valid tree. */
int pb_ctx_set_bvh_refit(pb_ctx *ctx, int max_consecutive_refits);
/* First build of the tree (and every rebuild after a topology change): PB_BVH_HOST_SAH (default) = binned SAH on the host, 40-60 ms for
 * 144 k triangles; PB_BVH_DEVICE_LBVH = Morton codes + radix sort + Karras topology on the device, boxes by the refit kernels: the whole
 * configure 4-7 ms, but its trees traverse ~27 % slower (6.4 vs 8.7 Grays/s on the bench) — for scenes whose topology changes every iteration.
 * Same hits either way. Replaces optixAccelBuild, include/psdr/scene/optix.h:277-340. */
enum { PB_BVH_HOST_SAH = 0, PB_BVH_DEVICE_LBVH = 1 };
int pb_ctx_set_bvh_builder(pb_ctx *ctx, int builder);
int pb_stats_bvh(pb_ctx *ctx, int *full_builds, int *refits);
int64_t pb_stats_last_rays(pb_ctx *ctx);
int64_t pb_stats_last_active_rays(pb_ctx *ctx);                                         /* rays the traversal kernels traced (inactive lanes are compacted away) */
int pb_stats_last_trace_launches(pb_ctx *ctx);
float pb_stats_last_primary_ms(pb_ctx *ctx);

/* ---- debugging / tuning hooks (not part of the reference surface) ---------------------------------------------- */
/* A/B switches behind the measurements in profiles/ (defaults in brackets). Process-wide unless noted:
 *   "sort_mode" [5] 5 = origin cell x direction octant, 0 = round 1's direction bin x cell, 7 = direction only
 *   "trace_kernel" [3] 3 = persistent streaming kernel, 1 = one ray per thread; "trace_node_min" [16]; "trace_chunk" [128]; "trace_blocks" [8]
 *   "sorted_copy" [0] (per context) 1 = the ray cast works on a sorted copy of the rays and writes hits in stream order
 *   "adjoint_lin" [1] 0 = reflectance adjoints connection by connection even when the forward pass kept the linearisation
 *   "shade_simple" [1] 0 = never use the diffuse + area-light instantiations; "shade_tune" [0] register caps of the event kernels
 *   "pipeline" [1] (per context) 0 = one batch at a time, 1 = two batches in flight for renders of at most "pipeline_max_lanes" lanes, 2 = always
 *   "l2_persist" [1], "rng_seed_table" [1] (per context)
 *   "bvh_builder" [0] (per context) = pb_ctx_set_bvh_builder; "lbvh_leaf" [2] most triangles per leaf of the device LBVH (1..8)
 * Unknown keys are an error. */
int pb_debug_set(pb_ctx *ctx, const char *key, int64_t value);
int pb_debug_ray_buffer(pb_ctx *ctx, int event, void **d_rays, int64_t *bytes);       /* rays of the last event traced by the last batch */
int pb_debug_retained_rad(pb_ctx *ctx, void **d_rad, int64_t *bytes);                  /* per-lane radiance kept by pb_render_d */

#ifdef __cplusplus
}
#endif
#endif /* PSDR_B200_H */
