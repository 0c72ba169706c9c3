// psdr-b200: device-side scene tables and wavefront records (the HBM layout).
//
// Replaces the reference's Enoki SoA arrays (include/psdr/types.h:136-146 TriangleInfo_, scene.h m_triangle_info /
// m_triangle_uv / m_triangle_face_normals / m_meshes_cuda, cuda/psdr_cuda.h:3-14 Params) with 16-byte-aligned records
// so that every gather and every wavefront read/write is a 128-bit access.
#pragma once
#include "pb_math.cuh"

namespace pb {

// ---- wavefront records ---------------------------------------------------------------------------
struct __align__(16) RayRec {      // 32 B
    float ox, oy, oz, tmax;        // tmax < 0  <=>  lane inactive (the ray is not traced)
    float dx, dy, dz, pad;
};
struct __align__(16) HitRec {      // 16 B: what cuda/psdr_cuda.cu:36-45 writes (tri id, shape id, u, v); -1 on miss
    int tri, shape;
    float u, v;
};

// ---- triangle table: one 128-byte line per triangle, indexed by global triangle id -------------------------------
// q[0] = p0.xyz, face_area     q[1] = e1.xyz, mesh id (int bits)   q[2] = e2.xyz, flags (bit0 face normals, bit1 has uv, bit3 mesh requires grad)
// q[3] = n0.xyz, uv0.x         q[4] = n1.xyz, uv0.y                q[5] = n2.xyz, uv1.x
// q[6] = face_normal.xyz, uv1.y                                    q[7] = uv2.x, uv2.y, -, -
struct __align__(16) TriRec { float4 q[8]; };

// triangles in BVH leaf order for the traversal kernel: p0.xyz,tri id | e1.xyz,mesh id | e2.xyz,- | padding. 64 B so that a
// record is two 256-bit loads (LDG.E.256 on sm_100a) instead of three 128-bit ones.
struct __align__(32) LeafTri { float4 a, b, c, pad; };

// BVH2 node, 64 B: both children's boxes + child links.
// a = (l.lo.x, l.lo.y, l.lo.z, l.hi.x)  b = (l.hi.y, l.hi.z, r.lo.x, r.lo.y)  c = (r.lo.z, r.hi.x, r.hi.y, r.hi.z)
// d = (left, right, -, -) as int bits; child >= 0: inner node index; child < 0: leaf, v = ~child, first = v >> 3, count = (v & 7) + 1
struct __align__(64) BvhNode { float4 a, b, c, d; };

// Compact BVH2 node for the wavefront traversal (pb_trace2.cuh), 48 B = three 128-bit loads: child boxes as centre + half extent,
// centres in fp32, x / y half extents as bf16 rounded up (left child's in the high halves, right child's in the low halves).
// a = (cL.x, cL.y, cL.z, cR.z)   b = (cR.x, cR.y, left, right)   c = (hx, hy, hL.z, hR.z) with hx / hy = bf16(hL) << 16 | bf16(hR) and the z half
// extents in fp32: every operand pair of the traversal's packed FFMA2 is an aligned register pair of a 128-bit load. Links as in BvhNode.
struct __align__(16) BvhNodeC { float4 a, b, c; };

struct MeshRec {           // 48 B, per mesh
    int bsdf, emitter;     // -1 = none
    float inv_total_area;
    int face_offset, num_faces;
    int flags;             // bit0 face normals, bit1 has uv, bit2 the vertex positions are a leaf (requires_grad)
    int pad0, pad1;
    const int *uv_faces;   // 3 uv indices per face, or nullptr
    float *uv_grad;        // VJP: gradient of Mesh.vertex_uv (2 floats per uv vertex; forward mode: its tangent), or nullptr
};
enum { BSDF_DIFFUSE = 0, BSDF_ROUGHCONDUCTOR = 1 };
enum { TEX_REFLECTANCE = 0, TEX_ALPHA_U, TEX_ALPHA_V, TEX_ETA, TEX_K, TEX_SPECULAR, TEX_COUNT };
struct TexRef {            // texel data interleaved [pixel*C + c]
    const float *data;
    float *grad;           // gradient accumulation target (same layout) or nullptr
    int w, h, c, pad;
};
struct BsdfRec {
    int type, pad[3];
    TexRef tex[TEX_COUNT];
};
enum { EMITTER_AREA = 0, EMITTER_ENVMAP = 1 };
struct EmitterRec {
    int type, mesh;
    float sampling_weight, pad;
    float3 radiance;
    float pad2;
    const float *face_cmf, *face_pmf;   // area pmf of the emitter's mesh (mesh.cpp:249)
    float face_sum;
    int num_faces, face_offset, pad3;
    // environment map (envmap.h): lat-long radiance bitmap, scale, rotation, the scene box it radiates from, and the
    // luminance*sin(theta) cell distribution over 2(w-1) x 2(h-1) cells (last dimension fastest)
    TexRef env_radiance;
    float env_scale;
    int env_res_x, env_res_y, env_cells;
    Mat4 env_to_world, env_from_world;
    float3 env_lower; float pad4;
    float3 env_upper; float env_sum;
    const float *env_cmf, *env_pmf;
    float *env_scale_grad;   // VJP: gradient of the scale (forward mode: its tangent), or nullptr
    float *env_xf_grad;      // VJP: 16 floats accumulating the adjoint of env_from_world (forward mode: its tangent), or nullptr
};
struct SensorRec {
    Mat4 sample_to_camera, to_world, world_to_sample;
    float3 camera_pos, camera_dir;
    float inv_area;
    int width, height;
};

// primary (camera-silhouette) edge, 48 B: projected endpoints, unit normal, length (edge.h:28-40) + who owns it
struct __align__(16) PrimEdgeRec { float p0x, p0y, p1x, p1y; float nx, ny, len, pad; int mesh, v0, v1, pad2; };
// secondary edge, 80 B (edge.h:50-65): p0|boundary flag, e1|mesh, n0|v0, n1|v1, p2|-
struct __align__(16) SecEdgeRec { float4 a, b, c, d, e; };
struct EdgeParams {
    const PrimEdgeRec *prim; const float *prim_cmf, *prim_pmf; float prim_sum; int num_prim;
    const SecEdgeRec *sec; const float *sec_cmf, *sec_pmf; float sec_sum; int num_sec;
    const float *guide_cmf, *guide_pmf; float guide_sum; int guide_cells; int guide_res[3];   // direct.cpp:166-204 (null: no guiding)
    float *const *mesh_gworld;          // per mesh: world-space vertex adjoint buffer (nullptr if the mesh needs no gradient)
    const float *const *mesh_vworld;    // per mesh: world-space vertex positions
};

struct SceneView {
    const TriRec *tri;
    const LeafTri *leaf;
    const BvhNode *nodes;
    const BvhNodeC *nodes_c;   // the same tree as `nodes` in the compact layout
    const MeshRec *meshes;
    const BsdfRec *bsdfs;
    const EmitterRec *emitters;
    const float *emitter_cmf, *emitter_pmf;   // scene.cpp:183-196
    float emitter_sum;
    int num_tri, num_nodes, num_meshes, num_bsdfs, num_emitters;
    int emitter_env;
    int simple;             // 1: only diffuse BSDFs and area emitters (kernels use their instantiation without rough-conductor / envmap code)
    float *tri_grad;        // VJP only: kTriGradStride floats per triangle (adjoint of the triangle table), or nullptr
    // forward mode (JVP): the same adjoint kernels run once per colour channel with a unit seed; instead of scattering a
    // local gradient they dot it with the tangent of what it refers to and add the result to the lane's accumulator
    const float *tri_tangent;   // kTriGradStride floats per triangle: tangent of the triangle table, or nullptr (reverse mode)
    float *jvp_acc;             // one float per lane of the batch
    int jvp_channel;            // colour channel of this pass
    float *jvp_image;           // W*H*3 derivative image
    // Sensor.to_world as a differentiable leaf (src/psdr.cpp:220-224): 32 floats or nullptr. Reverse mode: [0,16) accumulates the
    // adjoint of to_world through the camera rays (o = M (0,0,0,1), d = M3 d_cam; perspective.cpp:120-136), [16,32) the adjoint of
    // world_to_sample through the projected primary-edge end points (perspective.cpp:85-96); the host folds the second into the first.
    // Forward mode: the tangents of the two matrices (read only).
    float *sensor_grad;
};

enum { INTEG_DIRECT = 0, INTEG_FIELD = 1, INTEG_PATH = 2 };
enum { FIELD_SILHOUETTE = 0, FIELD_POSITION, FIELD_DEPTH, FIELD_GEONORMAL, FIELD_SHNORMAL, FIELD_UV };

PB_D float4 ldg4(const float4 *p) { return __ldg(p); }

}  // namespace pb
