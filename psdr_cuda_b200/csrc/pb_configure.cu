// psdr-b200: mesh preprocessing kernels — the device half of Scene::configure.
//
// Restates Mesh::configure / process_mesh (src/shape/mesh.cpp:19-51, 215-274) and the global triangle table build
// (src/scene/scene.cpp:205-217). The reference scatter-adds face normals into vertices with atomics; here every vertex
// gathers its incident faces through a CSR list laid out in the same (corner, face) order the oracle sums in, so the
// result is deterministic and bit-identical to the CPU restatement, and the backward pass is a gather as well.
#include "pb_adjoint_math.cuh"
#include "pb_kernels.h"

namespace pb {

__global__ void k_transform_vertices(int nv, const float *__restrict__ vraw, Mat4 to_world, float *__restrict__ vworld) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const float3 p = transform_pos(to_world, f3(vraw[3 * v], vraw[3 * v + 1], vraw[3 * v + 2]));
    vworld[3 * v] = p.x; vworld[3 * v + 1] = p.y; vworld[3 * v + 2] = p.z;
}

// per face: unnormalised normal (cross) and its length -> fcross[4*f] = (cx, cy, cz, |c|)
__global__ void k_face_cross(int nf, const float *__restrict__ vworld, const int *__restrict__ faces, float4 *__restrict__ fcross) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float3 p0 = f3(vworld[3 * i0], vworld[3 * i0 + 1], vworld[3 * i0 + 2]);
    const float3 e1 = sub3_rn(f3(vworld[3 * i1], vworld[3 * i1 + 1], vworld[3 * i1 + 2]), p0);
    const float3 e2 = sub3_rn(f3(vworld[3 * i2], vworld[3 * i2 + 1], vworld[3 * i2 + 2]), p0);
    const float3 c = cross(e1, e2);
    fcross[f] = make_float4(c.x, c.y, c.z, norm(c));
}

// per vertex: normalize(sum(cross) / sum(|cross|)) over incident faces in CSR order (mesh.cpp:32-38)
__global__ void k_vertex_normals(int nv, const int *__restrict__ csr_off, const int *__restrict__ csr_face, const float4 *__restrict__ fcross,
                                 float *__restrict__ vnormal) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    float3 acc = f3(0.f);
    float w = 0.f;
    for (int k = csr_off[v]; k < csr_off[v + 1]; ++k) {
        const float4 c = fcross[csr_face[k]];
        acc = f3(add_rn(acc.x, c.x), add_rn(acc.y, c.y), add_rn(acc.z, c.z));
        w = add_rn(w, c.w);
    }
    const float3 n = normalize(f3(div_rn(acc.x, w), div_rn(acc.y, w), div_rn(acc.z, w)));
    vnormal[3 * v] = n.x; vnormal[3 * v + 1] = n.y; vnormal[3 * v + 2] = n.z;
}

// assemble the 128-byte triangle records at their global offset + the face-area pmf of the mesh
__global__ void k_assemble_triangles(int nf, int face_offset, int mesh_id, int flags, const float *__restrict__ vworld, const int *__restrict__ faces,
                                     const float4 *__restrict__ fcross, const float *__restrict__ vnormal, const float *__restrict__ uvs,
                                     const int *__restrict__ uv_faces, TriRec *__restrict__ tri, float *__restrict__ face_area) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float3 p0 = f3(vworld[3 * i0], vworld[3 * i0 + 1], vworld[3 * i0 + 2]);
    const float3 e1 = sub3_rn(f3(vworld[3 * i1], vworld[3 * i1 + 1], vworld[3 * i1 + 2]), p0);
    const float3 e2 = sub3_rn(f3(vworld[3 * i2], vworld[3 * i2 + 1], vworld[3 * i2 + 2]), p0);
    const float4 c = fcross[f];
    const float3 fn = f3(div_rn(c.x, c.w), div_rn(c.y, c.w), div_rn(c.z, c.w));
    const float area = mul_rn(c.w, 0.5f);
    float2 uv0 = make_float2(0.f, 0.f), uv1 = uv0, uv2 = uv0;
    if (flags & 2) {
        const int j0 = uv_faces[3 * f], j1 = uv_faces[3 * f + 1], j2 = uv_faces[3 * f + 2];
        uv0 = make_float2(uvs[2 * j0], uvs[2 * j0 + 1]); uv1 = make_float2(uvs[2 * j1], uvs[2 * j1 + 1]); uv2 = make_float2(uvs[2 * j2], uvs[2 * j2 + 1]);
    }
    float4 *q = reinterpret_cast<float4 *>(tri + face_offset + f);
    q[0] = make_float4(p0.x, p0.y, p0.z, area);
    q[1] = make_float4(e1.x, e1.y, e1.z, __int_as_float(mesh_id));
    q[2] = make_float4(e2.x, e2.y, e2.z, __int_as_float(flags));
    q[3] = make_float4(vnormal[3 * i0], vnormal[3 * i0 + 1], vnormal[3 * i0 + 2], uv0.x);
    q[4] = make_float4(vnormal[3 * i1], vnormal[3 * i1 + 1], vnormal[3 * i1 + 2], uv0.y);
    q[5] = make_float4(vnormal[3 * i2], vnormal[3 * i2 + 1], vnormal[3 * i2 + 2], uv1.x);
    q[6] = make_float4(fn.x, fn.y, fn.z, uv1.y);
    q[7] = make_float4(uv2.x, uv2.y, 0.f, 0.f);
    face_area[f] = area;
}

// leaf-ordered triangles for traversal: gather p0/e1/e2 by the BVH's triangle order
__global__ void k_build_leaf_tris(int n, const int *__restrict__ order, const TriRec *__restrict__ tri, LeafTri *__restrict__ leaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = order[i];
    const float4 *q = reinterpret_cast<const float4 *>(tri + t);
    const float4 q0 = q[0], q1 = q[1], q2 = q[2];
    LeafTri l;
    l.a = make_float4(q0.x, q0.y, q0.z, __int_as_float(t));
    l.b = make_float4(q1.x, q1.y, q1.z, q1.w);   // w = mesh id
    l.c = make_float4(q2.x, q2.y, q2.z, 0.f);
    l.pad = make_float4(0.f, 0.f, 0.f, 0.f);
    leaf[i] = l;
}

// ---- BVH refit: same tree, new boxes ---------------------------------------------------------------------------------------
// An optimisation loop moves vertices a little every iteration and calls Scene::configure() each time (examples/run_test.py:117),
// where the reference rebuilds its OptiX GAS (optix.h:277-340). Here the binned-SAH build runs on the host (~50-90 ms for 70 k
// triangles); when only vertex positions changed the tree topology is kept and the child boxes are recomputed on the device,
// one launch per tree level from the deepest up (nodes are numbered breadth-first, so a level is a contiguous index range).
// Boxes only have to be conservative: the traversal returns the exact closest hit whatever the tree looks like.
// `boxes`: 12 floats per node, the unpadded child boxes (scratch, rewritten every refit); nodes get the padded ones
// (same padding rule as pb_bvh.cpp pad_box).
__global__ void k_bvh_refit_level(BvhNode *__restrict__ nodes, float *__restrict__ boxes, const LeafTri *__restrict__ leaf, int start, int end, float extent) {
    const int i = start + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= end) return;
    float4 *np = reinterpret_cast<float4 *>(nodes + i);
    const float4 a = np[0], b = np[1], c = np[2], d = np[3];
    const int link[2] = {__float_as_int(d.x), __float_as_int(d.y)};
    float old_lo[2][3] = {{a.x, a.y, a.z}, {b.z, b.w, c.x}}, old_hi[2][3] = {{a.w, b.x, b.y}, {c.y, c.z, c.w}};
    float lo[2][3], hi[2][3], plo[2][3], phi[2][3];
    const float kFar = 3.402823466e+38f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (old_lo[s][0] >= kFar) {   // unreachable child (pb_bvh.cpp): keep it
            for (int k = 0; k < 3; ++k) { lo[s][k] = plo[s][k] = old_lo[s][k]; hi[s][k] = phi[s][k] = old_hi[s][k]; }
            continue;
        }
        for (int k = 0; k < 3; ++k) { lo[s][k] = INFINITY; hi[s][k] = -INFINITY; }
        if (link[s] < 0) {
            const int v = ~link[s], first = v >> 3, cnt = (v & 7) + 1;
            for (int t = 0; t < cnt; ++t) {
                const LeafTri lt = leaf[first + t];
                const float p0[3] = {lt.a.x, lt.a.y, lt.a.z}, e1[3] = {lt.b.x, lt.b.y, lt.b.z}, e2[3] = {lt.c.x, lt.c.y, lt.c.z};
                for (int k = 0; k < 3; ++k) {
                    const float v1 = __fadd_rn(p0[k], e1[k]), v2 = __fadd_rn(p0[k], e2[k]);
                    lo[s][k] = fminf(lo[s][k], fminf(p0[k], fminf(v1, v2)));
                    hi[s][k] = fmaxf(hi[s][k], fmaxf(p0[k], fmaxf(v1, v2)));
                }
            }
        } else {
            const float *cb = boxes + 12 * (size_t)link[s];   // the child's own two (unpadded) child boxes, written by the previous launch
            for (int k = 0; k < 3; ++k) {
                lo[s][k] = fminf(cb[k] >= kFar ? INFINITY : cb[k], cb[6 + k] >= kFar ? INFINITY : cb[6 + k]);
                hi[s][k] = fmaxf(cb[k] >= kFar ? -INFINITY : cb[3 + k], cb[6 + k] >= kFar ? -INFINITY : cb[9 + k]);
            }
        }
        for (int k = 0; k < 3; ++k) {
            const float pad = __fadd_rn(__fmul_rn(1e-4f, fmaxf(1.f, fmaxf(fabsf(lo[s][k]), fabsf(hi[s][k])))), __fmul_rn(2.5e-7f, extent));
            plo[s][k] = lo[s][k] - pad; phi[s][k] = hi[s][k] + pad;
        }
    }
    float *ob = boxes + 12 * (size_t)i;
    for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) { ob[6 * s + k] = lo[s][k]; ob[6 * s + 3 + k] = hi[s][k]; }
    np[0] = make_float4(plo[0][0], plo[0][1], plo[0][2], phi[0][0]);
    np[1] = make_float4(phi[0][1], phi[0][2], plo[1][0], plo[1][1]);
    np[2] = make_float4(plo[1][2], phi[1][0], phi[1][1], phi[1][2]);
}

void launch_bvh_refit(cudaStream_t st, BvhNode *nodes, float *boxes, const LeafTri *leaf, const int *level_off, int num_levels, float extent) {
    for (int l = num_levels - 1; l >= 0; --l) {
        const int start = level_off[l], end = level_off[l + 1];
        if (end > start) k_bvh_refit_level<<<(end - start + 127) / 128, 128, 0, st>>>(nodes, boxes, leaf, start, end, extent);
    }
}

// ---- backward of the mesh preprocessing (mesh.cpp:19-51, 215-231): triangle-table adjoint -> vertex adjoints -------------
// csr_slot[k] = 3*face + corner for the k-th (vertex, incident face) pair, same order as the forward gather.

// per vertex: adjoint of the (unnormalised) normal sum A_v = sum_f c_f from the adjoints of n0/n1/n2 of its faces
__global__ void k_mesh_bwd_vertex_normal(int nv, int face_offset, const int *__restrict__ csr_off, const int *__restrict__ csr_slot,
                                         const float4 *__restrict__ fcross, const float *__restrict__ tri_grad, float *__restrict__ g_nsum) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    float3 acc = f3(0.f), gn = f3(0.f);
    float w = 0.f;
    for (int k = csr_off[v]; k < csr_off[v + 1]; ++k) {
        const int slot = csr_slot[k], f = slot / 3, corner = slot - 3 * f;
        const float4 c = fcross[f];
        acc += f3(c); w += c.w;
        const float *g = tri_grad + (size_t)(face_offset + f) * kTriGradStride + 9 + 3 * corner;
        gn += f3(g[0], g[1], g[2]);
    }
    float3 gA = f3(0.f);
    if (w > 0.f && (gn.x != 0.f || gn.y != 0.f || gn.z != 0.f)) gA = normalize_vjp(acc * (1.f / w), gn) * (1.f / w);
    g_nsum[3 * v] = gA.x; g_nsum[3 * v + 1] = gA.y; g_nsum[3 * v + 2] = gA.z;
}

// per face: adjoints of its three corners from the adjoints of p0/e1/e2/face normal/area and the vertex-normal sums
__global__ void k_mesh_bwd_face(int nf, int face_offset, const float *__restrict__ vworld, const int *__restrict__ faces,
                                const float *__restrict__ tri_grad, const float *__restrict__ g_nsum, float *__restrict__ g_corner) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float3 p0 = f3(vworld[3 * i0], vworld[3 * i0 + 1], vworld[3 * i0 + 2]);
    const float3 e1 = f3(vworld[3 * i1], vworld[3 * i1 + 1], vworld[3 * i1 + 2]) - p0;
    const float3 e2 = f3(vworld[3 * i2], vworld[3 * i2 + 1], vworld[3 * i2 + 2]) - p0;
    const float *g = tri_grad + (size_t)(face_offset + f) * kTriGradStride;
    const float3 g_p0 = f3(g[0], g[1], g[2]);
    float3 g_e1 = f3(g[3], g[4], g[5]), g_e2 = f3(g[6], g[7], g[8]);
    const float3 g_fn = f3(g[18], g[19], g[20]);
    const float g_area = g[21];
    const float3 g_cn = f3(g_nsum[3 * i0] + g_nsum[3 * i1] + g_nsum[3 * i2], g_nsum[3 * i0 + 1] + g_nsum[3 * i1 + 1] + g_nsum[3 * i2 + 1],
                           g_nsum[3 * i0 + 2] + g_nsum[3 * i1 + 2] + g_nsum[3 * i2 + 2]);
    face_vjp(e1, e2, g_fn, g_area, g_cn, g_e1, g_e2);
    const float3 c0 = g_p0 - g_e1 - g_e2;
    float *o = g_corner + 9 * (size_t)f;
    o[0] = c0.x; o[1] = c0.y; o[2] = c0.z; o[3] = g_e1.x; o[4] = g_e1.y; o[5] = g_e1.z; o[6] = g_e2.x; o[7] = g_e2.y; o[8] = g_e2.z;
}

// per vertex: gather the corner adjoints (+ the direct world-space adjoint from the boundary terms), pull back through
// transform_pos (mesh.cpp:223-231) and accumulate into the gradient segment of the object-space vertex positions
__global__ void k_mesh_bwd_vertex(int nv, const int *__restrict__ csr_off, const int *__restrict__ csr_slot, const float *__restrict__ g_corner,
                                  const float *__restrict__ g_world_direct, const float *__restrict__ vraw, Mat4 M, float *__restrict__ grad_out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    float3 gw = f3(g_world_direct[3 * v], g_world_direct[3 * v + 1], g_world_direct[3 * v + 2]);
    for (int k = csr_off[v]; k < csr_off[v + 1]; ++k) {
        const float *c = g_corner + 3 * (size_t)csr_slot[k];
        gw += f3(c[0], c[1], c[2]);
    }
    // p = t.xyz / t.w, t = M (x, 1)
    const float3 x = f3(vraw[3 * v], vraw[3 * v + 1], vraw[3 * v + 2]);
    float t[4];
    for (int i = 0; i < 4; ++i) t[i] = M.m[4 * i] * x.x + M.m[4 * i + 1] * x.y + M.m[4 * i + 2] * x.z + M.m[4 * i + 3];
    const float3 p = f3(t[0] / t[3], t[1] / t[3], t[2] / t[3]);
    const float gt[4] = {gw.x / t[3], gw.y / t[3], gw.z / t[3], -pdot(gw, p) / t[3]};
    for (int j = 0; j < 3; ++j) {
        float acc = 0.f;
        for (int i = 0; i < 4; ++i) acc += M.m[4 * i + j] * gt[i];
        grad_out[3 * v + j] += acc;
    }
}

// ---- forward mode of the mesh preprocessing: vertex tangents -> tangent of the triangle table ---------------------------------
__global__ void k_mesh_tan_vertices(int nv, const float *__restrict__ vraw, const float *__restrict__ vraw_t, Mat4 M, float *__restrict__ vworld_t) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const float3 x = f3(vraw[3 * v], vraw[3 * v + 1], vraw[3 * v + 2]), u = f3(vraw_t[3 * v], vraw_t[3 * v + 1], vraw_t[3 * v + 2]);
    float t[4], dt[4];
    for (int i = 0; i < 4; ++i) {
        t[i] = M.m[4 * i] * x.x + M.m[4 * i + 1] * x.y + M.m[4 * i + 2] * x.z + M.m[4 * i + 3];
        dt[i] = M.m[4 * i] * u.x + M.m[4 * i + 1] * u.y + M.m[4 * i + 2] * u.z;
    }
    for (int k = 0; k < 3; ++k) vworld_t[3 * v + k] = (dt[k] - (t[k] / t[3]) * dt[3]) / t[3];
}
// per face: tangent of the unnormalised normal c = e1 x e2 and of its length -> fcross_t = (dc, d|c|)
__global__ void k_mesh_tan_faces(int nf, const float *__restrict__ vworld, const float *__restrict__ vworld_t, const int *__restrict__ faces,
                                 float4 *__restrict__ fcross_t) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    auto ld = [](const float *a, int i) { return f3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); };
    const float3 p0 = ld(vworld, i0), e1 = ld(vworld, i1) - p0, e2 = ld(vworld, i2) - p0;
    const float3 dp0 = ld(vworld_t, i0), de1 = ld(vworld_t, i1) - dp0, de2 = ld(vworld_t, i2) - dp0;
    const float3 c = pcross(e1, e2), dc = pcross(de1, e2) + pcross(e1, de2);
    fcross_t[f] = make_float4(dc.x, dc.y, dc.z, pdot(c, dc) / sqrtf(pdot(c, c)));
}
__global__ void k_mesh_tan_vnormals(int nv, const int *__restrict__ csr_off, const int *__restrict__ csr_face, const float4 *__restrict__ fcross,
                                    const float4 *__restrict__ fcross_t, float *__restrict__ vnormal_t) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    float3 A = f3(0.f), dA = f3(0.f);
    float w = 0.f, dw = 0.f;
    for (int k = csr_off[v]; k < csr_off[v + 1]; ++k) {
        const float4 c = fcross[csr_face[k]], dc = fcross_t[csr_face[k]];
        A += f3(c); w += c.w; dA += f3(dc); dw += dc.w;
    }
    float3 dn = f3(0.f);
    if (w > 0.f) {
        const float3 B = A * (1.f / w), dB = (dA - B * dw) * (1.f / w);
        const float len = sqrtf(pdot(B, B));
        const float3 n = B * (1.f / len);
        dn = (dB - n * pdot(n, dB)) * (1.f / len);
    }
    vnormal_t[3 * v] = dn.x; vnormal_t[3 * v + 1] = dn.y; vnormal_t[3 * v + 2] = dn.z;
}
__global__ void k_mesh_tan_assemble(int nf, int face_offset, const float *__restrict__ vworld_t, const int *__restrict__ faces, const float4 *__restrict__ fcross,
                                    const float4 *__restrict__ fcross_t, const float *__restrict__ vnormal_t, float *__restrict__ tri_tangent) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nf) return;
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    auto ld = [](const float *a, int i) { return f3(a[3 * i], a[3 * i + 1], a[3 * i + 2]); };
    const float3 dp0 = ld(vworld_t, i0), de1 = ld(vworld_t, i1) - dp0, de2 = ld(vworld_t, i2) - dp0;
    const float4 c = fcross[f], dc = fcross_t[f];
    const float3 fn = f3(c) * (1.f / c.w);
    const float3 dfn = (f3(dc) - fn * pdot(fn, f3(dc))) * (1.f / c.w);
    float *o = tri_tangent + (size_t)(face_offset + f) * kTriGradStride;
    const float3 rows[7] = {dp0, de1, de2, ld(vnormal_t, i0), ld(vnormal_t, i1), ld(vnormal_t, i2), dfn};
    for (int k = 0; k < 7; ++k) { o[3 * k] = rows[k].x; o[3 * k + 1] = rows[k].y; o[3 * k + 2] = rows[k].z; }
    o[21] = 0.5f * dc.w; o[22] = 0.f; o[23] = 0.f;
}

static inline int nblk(int n, int b) { return (n + b - 1) / b; }

void launch_mesh_tangent(cudaStream_t st, int nv, int nf, int face_offset, const float *vraw, const float *vraw_t, const Mat4 &to_world, const float *vworld,
                         const int *faces, const int *csr_off, const int *csr_face, const float4 *fcross, float *vworld_t, float4 *fcross_t, float *vnormal_t,
                         float *tri_tangent) {
    if (nv <= 0 || nf <= 0) return;
    k_mesh_tan_vertices<<<nblk(nv, 256), 256, 0, st>>>(nv, vraw, vraw_t, to_world, vworld_t);
    k_mesh_tan_faces<<<nblk(nf, 256), 256, 0, st>>>(nf, vworld, vworld_t, faces, fcross_t);
    k_mesh_tan_vnormals<<<nblk(nv, 256), 256, 0, st>>>(nv, csr_off, csr_face, fcross, fcross_t, vnormal_t);
    k_mesh_tan_assemble<<<nblk(nf, 256), 256, 0, st>>>(nf, face_offset, vworld_t, faces, fcross, fcross_t, vnormal_t, tri_tangent);
}

void launch_mesh_backward(cudaStream_t st, int nv, int nf, int face_offset, const int *csr_off, const int *csr_slot, const float4 *fcross,
                          const float *vworld, const int *faces, const float *vraw, const Mat4 &to_world, const float *tri_grad,
                          const float *g_world_direct, float *g_nsum, float *g_corner, float *grad_out) {
    if (nv <= 0 || nf <= 0) return;
    k_mesh_bwd_vertex_normal<<<nblk(nv, 256), 256, 0, st>>>(nv, face_offset, csr_off, csr_slot, fcross, tri_grad, g_nsum);
    k_mesh_bwd_face<<<nblk(nf, 256), 256, 0, st>>>(nf, face_offset, vworld, faces, tri_grad, g_nsum, g_corner);
    k_mesh_bwd_vertex<<<nblk(nv, 256), 256, 0, st>>>(nv, csr_off, csr_slot, g_corner, g_world_direct, vraw, to_world, grad_out);
}

void launch_mesh_preprocess(cudaStream_t st, int nv, int nf, int face_offset, int mesh_id, int flags, const float *vraw, const Mat4 &to_world,
                            const int *faces, const int *csr_off, const int *csr_face, const float *uvs, const int *uv_faces, float *vworld,
                            float4 *fcross, float *vnormal, TriRec *tri, float *face_area) {
    if (nv > 0) k_transform_vertices<<<nblk(nv, 256), 256, 0, st>>>(nv, vraw, to_world, vworld);
    if (nf > 0) k_face_cross<<<nblk(nf, 256), 256, 0, st>>>(nf, vworld, faces, fcross);
    if (nv > 0) k_vertex_normals<<<nblk(nv, 256), 256, 0, st>>>(nv, csr_off, csr_face, fcross, vnormal);
    if (nf > 0) k_assemble_triangles<<<nblk(nf, 256), 256, 0, st>>>(nf, face_offset, mesh_id, flags, vworld, faces, fcross, vnormal, uvs, uv_faces, tri, face_area);
}

// BvhNode (padded lo / hi boxes) -> BvhNodeC (centre + half extent). Conservative: c = rn((lo + hi) / 2), h = ru(max(hi - c, c - lo))
// + the rounding slack of the three-FFMA slab test (2^-24 (3 |o| + 2 |c|) in world units <= 4e-7 extent), then rounded up to bf16.
__device__ __forceinline__ unsigned bf16_up(float h) {
    const unsigned u = __float_as_uint(h);
    return (u + 0xffffu) >> 16;   // h >= 0 and finite: next bf16 at or above
}
__global__ void k_nodes_to_compact(int n, const BvhNode *__restrict__ nodes, BvhNodeC *__restrict__ out, float extent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *np = reinterpret_cast<const float4 *>(nodes + i);
    const float4 a = np[0], b = np[1], c = np[2], d = np[3];
    const float lo[2][3] = {{a.x, a.y, a.z}, {b.z, b.w, c.x}}, hi[2][3] = {{a.w, b.x, b.y}, {c.y, c.z, c.w}};
    float cen[2][3], hf[2][3];
    const float slack = __fmul_ru(4e-7f, extent);
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            if (lo[s][k] >= 3.402823466e+38f) { cen[s][k] = lo[s][k]; hf[s][k] = 0.f; continue; }   // unreachable child (pb_bvh.cpp)
            const float m = __fadd_rn(__fmul_rn(0.5f, lo[s][k]), __fmul_rn(0.5f, hi[s][k]));
            cen[s][k] = m;
            hf[s][k] = fmaxf(__fadd_ru(fmaxf(__fsub_ru(hi[s][k], m), __fsub_ru(m, lo[s][k])), slack), 0.f);
        }
    BvhNodeC o;
    o.a = make_float4(cen[0][0], cen[0][1], cen[0][2], cen[1][2]);
    o.b = make_float4(cen[1][0], cen[1][1], d.x, d.y);
    o.c = make_float4(__uint_as_float((bf16_up(hf[0][0]) << 16) | bf16_up(hf[1][0])), __uint_as_float((bf16_up(hf[0][1]) << 16) | bf16_up(hf[1][1])), hf[0][2], hf[1][2]);
    out[i] = o;
}
void launch_nodes_to_compact(cudaStream_t st, int num_nodes, const BvhNode *nodes, BvhNodeC *out, float extent) {
    if (num_nodes > 0) k_nodes_to_compact<<<nblk(num_nodes, 256), 256, 0, st>>>(num_nodes, nodes, out, extent);
}

// seeded PCG32 state of every lane (sampler.cpp:29-40): computed once per context, read by every kernel that draws samples
__global__ void __launch_bounds__(256) k_rng_seed(long long n, ulonglong2 *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RngSeed s = rng_seed((uint64_t)i);
    out[i] = make_ulonglong2(s.state, s.inc);
}
void launch_rng_seed(cudaStream_t st, long long n, ulonglong2 *out) {
    if (n > 0) k_rng_seed<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, out);
}

void launch_build_leaf_tris(cudaStream_t st, int n, const int *order, const TriRec *tri, LeafTri *leaf) {
    if (n > 0) k_build_leaf_tris<<<nblk(n, 256), 256, 0, st>>>(n, order, tri, leaf);
}

}  // namespace pb
