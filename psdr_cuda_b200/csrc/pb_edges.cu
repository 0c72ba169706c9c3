// psdr-b200: boundary-integral kernels — primary (silhouette seen by the camera) and secondary (direct-illumination
// shadow boundary) edge samplers with their gradient estimators, reverse mode.
//
// Restates Integrator::render_primary_edges (src/integrator/integrator.cpp:98-119), PerspectiveCamera::sample_primary_edge
// (src/sensor/perspective.cpp:158-200), Scene::sample_boundary_segment_direct (src/scene/scene.cpp:456-492),
// DirectIntegrator::eval_secondary_edge / render_secondary_edges (src/integrator/direct.cpp:207-316) and
// PerspectiveCamera::sample_direct (src/sensor/perspective.cpp:139-155). Both terms have zero primal (`value -=
// detach(value)`): they exist only in the derivative, so they run inside pb_render_d_vjp and scatter straight into the
// vertex adjoints (world-space vertex buffer of the owning mesh / the triangle-table gradient).
#include "pb_trace.cuh"
#include "pb_wavefront.cuh"

namespace pb {

PB_D PrimEdgeRec load_prim_edge(const PrimEdgeRec *p) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
    const float4 a = ldg4(q), b = ldg4(q + 1), c = ldg4(q + 2);
    PrimEdgeRec r;
    r.p0x = a.x; r.p0y = a.y; r.p1x = a.z; r.p1y = a.w;
    r.nx = b.x; r.ny = b.y; r.len = b.z; r.pad = 0.f;
    r.mesh = __float_as_int(c.x); r.v0 = __float_as_int(c.y); r.v1 = __float_as_int(c.z); r.pad2 = 0;
    return r;
}

struct PrimEdgeSample { float s, pdf, px, py, x_dot_n; int idx, edge; PrimEdgeRec rec; };

// perspective.cpp:158-200
PB_D PrimEdgeSample sample_primary_edge(const EdgeParams &Q, const SensorRec &cam, float sample1) {
    PrimEdgeSample r;
    float pdf;
    r.edge = sample_reuse(Q.prim_cmf, Q.prim_pmf, Q.num_prim, Q.prim_sum, sample1, pdf);
    r.rec = load_prim_edge(Q.prim + r.edge);
    r.s = sample1;
    r.pdf = div_rn(pdf, r.rec.len);
    r.px = fma_rn(r.rec.p0x, 1.f - sample1, mul_rn(r.rec.p1x, sample1));
    r.py = fma_rn(r.rec.p0y, 1.f - sample1, mul_rn(r.rec.p1y, sample1));
    r.x_dot_n = fma_rn(r.px, r.rec.nx, mul_rn(r.py, r.rec.ny));
    const int ix = (int)floorf(r.px * (float)cam.width), iy = (int)floorf(r.py * (float)cam.height);
    const bool valid = ix >= 0 && ix < cam.width && iy >= 0 && iy < cam.height;
    r.idx = valid ? iy * cam.width + ix : -1;
    return r;
}

// camera rays through p +- EdgeEpsilon * n, traced; lanes whose sample falls off the film are inactive (integrator.cpp:104-110)
// (emitted as a wavefront: edge samples of neighbouring lanes are unrelated, so the rays go through the direction sort + streaming kernel)
__global__ void __launch_bounds__(256) k_edge_primary_rays(RenderParams P, EdgeParams Q, int side, RayRec *__restrict__ rays) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int pix;
    const long long lane = global_lane(P, i, pix);
    Rng rng = make_rng(P, lane, P.jump0);
    const PrimEdgeSample es = sample_primary_edge(Q, P.cam, rng.next_1d());
    const float sg = side == 0 ? 1.f : -1.f;   // side 0: ray_p (+n), side 1: ray_n (-n)
    float3 o, d;
    sample_primary_ray(P.cam, es.px + sg * kEdgeEpsilon * es.rec.nx, es.py + sg * kEdgeEpsilon * es.rec.ny, o, d);
    store_ray(rays + i, o, d, es.idx >= 0 ? INFINITY : -1.f);
}

// adjoint of a world-space vertex: reverse mode adds it to the mesh's vertex-adjoint buffer; forward mode (the buffer then
// holds the vertex tangents) returns its dot product with the tangent
PB_D float vertex_adjoint(const SceneView &S, float *buf, int v, float3 g) {
    if (!finite3(g)) return 0.f;   // degenerate samples must not poison the gradient
    if (S.tri_tangent) return g.x * buf[3 * v] + g.y * buf[3 * v + 1] + g.z * buf[3 * v + 2];
    atomicAdd(buf + 3 * v, g.x); atomicAdd(buf + 3 * v + 1, g.y); atomicAdd(buf + 3 * v + 2, g.z);
    return 0.f;
}

// integrator.cpp:111-117: value = x_dot_n * (L_n - L_p) / pdf / sppe; only x_dot_n carries a derivative. Its adjoint goes
// through the two projected endpoints (perspective.cpp:85-96) back to the world-space vertices of the edge.
__global__ void __launch_bounds__(256) k_edge_primary_grad(RenderParams P, EdgeParams Q, const float4 *__restrict__ rad_p, const float4 *__restrict__ rad_n,
                                                           const float *__restrict__ dLdI, float inv_sppe) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int pix;
    const long long lane = global_lane(P, i, pix);
    Rng rng = make_rng(P, lane, P.jump0);
    const PrimEdgeSample es = sample_primary_edge(Q, P.cam, rng.next_1d());
    if (es.idx < 0) return;
    float *gworld = Q.mesh_gworld[es.rec.mesh];
    if (!gworld && !P.S.sensor_grad) return;
    const float3 delta = f3(rad_n[i]) - f3(rad_p[i]);
    const bool jvp = P.S.tri_tangent != nullptr;
    float w = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float v = es.x_dot_n * (getc(delta, c) / es.pdf);
        const float seed = jvp ? (c == P.S.jvp_channel ? 1.f : 0.f) : __ldg(dLdI + 3 * (size_t)es.idx + c);
        if (isfinite(v)) w += seed * (getc(delta, c) / es.pdf) * inv_sppe;
    }
    if (w == 0.f) return;
    float jsum = 0.f;
    // x_dot_n = p_ . n,  p_ = (1-s) q0 + s q1,  q = (M (v,1)).xy / (M (v,1)).w
    const float gq[2][2] = {{w * es.rec.nx * (1.f - es.s), w * es.rec.ny * (1.f - es.s)}, {w * es.rec.nx * es.s, w * es.rec.ny * es.s}};
    const float qv[2][2] = {{es.rec.p0x, es.rec.p0y}, {es.rec.p1x, es.rec.p1y}};
    const int vid[2] = {es.rec.v0, es.rec.v1};
    const float *M = P.cam.world_to_sample.m;
    const float *vw = Q.mesh_vworld[es.rec.mesh];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float3 x = f3(vw[3 * vid[e]], vw[3 * vid[e] + 1], vw[3 * vid[e] + 2]);
        const float tw = M[12] * x.x + M[13] * x.y + M[14] * x.z + M[15];
        const float gt0 = gq[e][0] / tw, gt1 = gq[e][1] / tw, gt3 = -(gq[e][0] * qv[e][0] + gq[e][1] * qv[e][1]) / tw;
        const float3 gx = f3(M[0] * gt0 + M[4] * gt1 + M[12] * gt3, M[1] * gt0 + M[5] * gt1 + M[13] * gt3, M[2] * gt0 + M[6] * gt1 + M[14] * gt3);
        if (gworld) jsum += vertex_adjoint(P.S, gworld, vid[e], gx);
        jsum += sensor_projection_adjoint(P.S, x, gt0, gt1, gt3);
    }
    if (jvp && jsum != 0.f && isfinite(jsum)) atomicAdd(P.S.jvp_image + 3 * (size_t)es.idx + P.S.jvp_channel, jsum);
}

// ---- secondary edges ---------------------------------------------------------------------------------------------------
struct SecEdge { float3 p0, e1, n0, n1, p2; bool boundary; int mesh, v0, v1; };
PB_D SecEdge load_sec_edge(const SecEdgeRec *p) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
    const float4 a = ldg4(q), b = ldg4(q + 1), c = ldg4(q + 2), d = ldg4(q + 3), e = ldg4(q + 4);
    SecEdge r;
    r.p0 = f3(a); r.boundary = a.w != 0.f;
    r.e1 = f3(b); r.mesh = __float_as_int(b.w);
    r.n0 = f3(c); r.v0 = __float_as_int(c.w);
    r.n1 = f3(d); r.v1 = __float_as_int(d.w);
    r.p2 = f3(e);
    return r;
}
PB_D int sign_eps(float x, float eps) { return x > eps ? 1 : (x < -eps ? -1 : 0); }

struct BoundarySample { float3 p0, edge, edge2, p2, n; float pdf, s, guide_pdf; bool valid; SecEdge info; int light_tri; float ls, lt; };

// the lane's three sample dimensions, optionally warped by the guiding grid (direct.cpp:210-212, cube_distrb.cpp:41-47)
PB_D float3 secondary_sample3(const EdgeParams &Q, Rng &rng, float &guide_pdf) {
    float3 s3 = rng.next_3d();
    guide_pdf = 1.f;
    if (Q.guide_cmf) {
        float pdf;
        const int cell = sample_reuse(Q.guide_cmf, Q.guide_pmf, Q.guide_cells, Q.guide_sum, s3.z, pdf);
        int c = cell;
        const int cz = c % Q.guide_res[2]; c /= Q.guide_res[2];
        const int cy = c % Q.guide_res[1]; c /= Q.guide_res[1];
        const int cx = c;
        s3 = f3((s3.x + (float)cx) * (1.f / (float)Q.guide_res[0]), (s3.y + (float)cy) * (1.f / (float)Q.guide_res[1]), (s3.z + (float)cz) * (1.f / (float)Q.guide_res[2]));
        guide_pdf = pdf * (float)Q.guide_cells;
    }
    return s3;
}

// the sample a lane evaluates: the render's (optionally guided) sample, or — while building the guiding grid
// (direct.cpp:182-192, guide_spc > 0) — a stratified sample inside the lane's cell
PB_D float3 lane_sample3(const RenderParams &P, const EdgeParams &Q, int i, Rng &rng, int guide_spc, float &guide_pdf) {
    if (guide_spc <= 0) return secondary_sample3(Q, rng, guide_pdf);
    int c = (int)((P.local0 + i) / guide_spc);
    const int cz = c % Q.guide_res[2]; c /= Q.guide_res[2];
    const int cy = c % Q.guide_res[1]; c /= Q.guide_res[1];
    const float3 u = rng.next_3d();
    guide_pdf = 1.f;
    return f3(((float)c + u.x) * (1.f / (float)Q.guide_res[0]), ((float)cy + u.y) * (1.f / (float)Q.guide_res[1]), ((float)cz + u.z) * (1.f / (float)Q.guide_res[2]));
}

// scene.cpp:456-492
PB_D BoundarySample sample_boundary_segment_direct(const SceneView &S, const EdgeParams &Q, float3 sample3) {
    BoundarySample r;
    float sample1 = sample3.x, pdf0;
    const int edge = sample_reuse(Q.sec_cmf, Q.sec_pmf, Q.num_sec, Q.sec_sum, sample1, pdf0);
    r.info = load_sec_edge(Q.sec + edge);
    r.s = sample1;
    r.p0 = f3(fma_rn(r.info.e1.x, sample1, r.info.p0.x), fma_rn(r.info.e1.y, sample1, r.info.p0.y), fma_rn(r.info.e1.z, sample1, r.info.p0.z));
    r.edge = normalize(r.info.e1);
    r.edge2 = r.info.p2 - r.info.p0;
    pdf0 = div_rn(pdf0, norm(r.info.e1));
    const PositionSample ps2 = sample_emitter_position(S, r.p0, make_float2(sample3.y, sample3.z), true);
    r.p2 = ps2.p; r.n = ps2.n; r.light_tri = ps2.tri; r.ls = ps2.s; r.lt = ps2.t;
    float3 e = r.p2 - r.p0;
    const float dist_sqr = squared_norm(e);
    e = e / safe_sqrt(dist_sqr);
    const float cos_theta = dot(r.n, -e);
    const int sgn0 = sign_eps(dot(r.info.n0, e), kEdgeEpsilon), sgn1 = sign_eps(dot(r.info.n1, e), kEdgeEpsilon);
    r.valid = cos_theta > kEpsilon && ((r.info.boundary && sgn0 != 0) || (!r.info.boundary && sgn0 * sgn1 < 0));
    r.pdf = r.valid ? pdf0 * ps2.pdf * (dist_sqr / cos_theta) : 0.f;
    return r;
}

// Scene::sample_boundary_segment_direct as a call of its own (src/psdr.cpp:274): n samples (3 floats each) -> 17 floats each
// (p0, edge, edge2, p2, n, pdf, is_valid)
__global__ void __launch_bounds__(256) k_sample_boundary_segment(int n, SceneView S, EdgeParams Q, const float *__restrict__ sample3, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const BoundarySample b = sample_boundary_segment_direct(S, Q, f3(sample3[3 * i], sample3[3 * i + 1], sample3[3 * i + 2]));
    float *o = out + 17 * (size_t)i;
    const float3 v[5] = {b.p0, b.edge, b.edge2, b.p2, b.n};
#pragma unroll
    for (int k = 0; k < 5; ++k) { o[3 * k] = v[k].x; o[3 * k + 1] = v[k].y; o[3 * k + 2] = v[k].z; }
    o[15] = b.pdf; o[16] = b.valid ? 1.f : 0.f;
}
void launch_sample_boundary_segment(cudaStream_t st, int n, const SceneView &S, const EdgeParams &Q, const float *sample3, float *out) {
    if (n > 0) k_sample_boundary_segment<<<(n + 255) / 256, 256, 0, st>>>(n, S, Q, sample3, out);
}

// stage A: sample the boundary segment; emit ray 0 (edge point -> emitter point) and ray 1 (opposite direction)
__global__ void __launch_bounds__(256) k_edge_secondary_rays(RenderParams P, EdgeParams Q, RayRec *__restrict__ rays, int guide_spc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    int pix;
    const long long lane = global_lane(P, i, pix);
    Rng rng = make_rng(P, lane, P.jump0);
    float gp;
    const BoundarySample bss = sample_boundary_segment_direct(P.S, Q, lane_sample3(P, Q, i, rng, guide_spc, gp));
    const float3 dir = normalize(bss.p2 - bss.p0);
    store_ray(rays + i, bss.p0, dir, bss.valid ? INFINITY : -1.f);
    store_ray(rays + (size_t)P.n + i, bss.p0, -dir, bss.valid ? INFINITY : -1.f);
}

// perspective.cpp:139-155
struct SensorDirect { float qx, qy, sensor_val; int pixel; bool valid; };
PB_D SensorDirect sensor_sample_direct(const SensorRec &cam, float3 p) {
    SensorDirect r;
    const float3 q = transform_pos(cam.world_to_sample, p);
    r.qx = q.x; r.qy = q.y;
    const int ix = (int)floorf(q.x * (float)cam.width), iy = (int)floorf(q.y * (float)cam.height);
    r.valid = ix >= 0 && ix < cam.width && iy >= 0 && iy < cam.height;
    r.pixel = r.valid ? iy * cam.width + ix : -1;
    float3 dir = p - cam.camera_pos;
    const float dist2 = squared_norm(dir);
    dir = dir / safe_sqrt(dist2);
    const float rc = 1.f / dot(cam.camera_dir, dir);
    r.sensor_val = (1.f / dist2) * (rc * rc * rc) * cam.inv_area;
    return r;
}

// stage B: the two hits decide validity (direct.cpp:234-247); project p1 onto the film and emit the camera ray through it
__global__ void __launch_bounds__(256) k_edge_secondary_camera(RenderParams P, EdgeParams Q, const RayRec *__restrict__ rays, const HitRec *__restrict__ hits,
                                                               RayRec *__restrict__ cam_rays, int guide_spc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const float4 ra = ldg4(reinterpret_cast<const float4 *>(rays + i));
    bool valid = ra.w > 0.f;
    const float3 p0 = f3(ra);
    float3 o = f3(0.f), d = f3(0.f, 0.f, 1.f);
    if (valid) {
        int pix;
        const long long lane = global_lane(P, i, pix);
        Rng rng = make_rng(P, lane, P.jump0);
        float gp;
        const BoundarySample bss = sample_boundary_segment_direct(P.S, Q, lane_sample3(P, Q, i, rng, guide_spc, gp));
        const Its its2 = reconstruct_its(P.S, load_hit(hits + i), p0);
        valid = its2.valid && norm(its2.p - bss.p2) < kShadowEpsilon;
        const Its its1 = reconstruct_its(P.S, load_hit(hits + (size_t)P.n + i), p0);
        valid = valid && its1.valid;
        if (valid) {
            const SensorDirect sds = sensor_sample_direct(P.cam, its1.p);
            valid = sds.valid;
            if (valid) sample_primary_ray(P.cam, sds.qx, sds.qy, o, d);
        }
    }
    store_ray(cam_rays + i, o, d, valid ? INFINITY : -1.f);
}

// stage C: direct.cpp:249-311 — value0 and the normal velocity; reverse mode scatters into the emitter triangle, the
// triangle seen by the camera and the two vertices of the sampled edge. With `guide_out` set it instead accumulates
// hmax(value0) per guiding cell (direct.cpp:166-204; no sign factors, no derivative).
__global__ void __launch_bounds__(256) k_edge_secondary_eval(RenderParams P, EdgeParams Q, const RayRec *__restrict__ rays, const HitRec *__restrict__ hits,
                                                             const RayRec *__restrict__ cam_rays, const HitRec *__restrict__ cam_hits,
                                                             const float *__restrict__ dLdI, float inv_sppse, float *__restrict__ guide_out, int guide_spc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const float4 ca = ldg4(reinterpret_cast<const float4 *>(cam_rays + i)), cb = ldg4(reinterpret_cast<const float4 *>(cam_rays + i) + 1);
    if (!(ca.w > 0.f)) return;
    const float3 cam_o = f3(ca), cam_d = f3(cb);
    const HitRec hc = load_hit(cam_hits + i);
    if (hc.tri < 0) return;
    int pix;
    const long long lane = global_lane(P, i, pix);
    Rng rng = make_rng(P, lane, P.jump0);
    float guide_pdf;
    const float3 s3 = lane_sample3(P, Q, i, rng, guide_out ? guide_spc : 0, guide_pdf);
    const BoundarySample bss = sample_boundary_segment_direct(P.S, Q, s3);
    const float3 p0 = bss.p0, dir = normalize(bss.p2 - p0);
    const HitRec h2 = load_hit(hits + i), h1 = load_hit(hits + (size_t)P.n + i);
    const Its its2 = reconstruct_its(P.S, h2, p0), its1 = reconstruct_its(P.S, h1, p0);
    const float3 p1 = its1.p;
    // the camera ray must see p1 (direct.cpp:253-264): solid-angle flavour in AD mode
    const Its itc = guide_out ? reconstruct_its(P.S, hc, cam_o) : reconstruct_its_primary(P.S, hc, cam_o, cam_d);
    if (!(norm(itc.p - p1) < kShadowEpsilon)) return;
    const SensorDirect sds = sensor_sample_direct(P.cam, p1);
    const float dist = norm(bss.p2 - p1), cos2 = fabsf(dot(bss.n, -dir));
    const float3 e = cross(bss.edge, dir);
    const float sinphi = norm(e);
    const float3 proj = normalize(cross(e, bss.n));
    const float sinphi2 = norm(cross(dir, proj));
    if (!(sinphi > kEpsilon && sinphi2 > kEpsilon)) return;
    const float base_v = (its1.t / dist) * (sinphi / sinphi2) * cos2;
    const float3 d0 = -cam_d, d0_local = its1.sh.to_local(d0);
    // direct.cpp:278-284: with one BSDF (or one mesh) in the scene the reference evaluates meshes[0]'s BSDF without looking at the mesh
    // that was hit, so an end point on the environment map's bounding mesh (no BSDF of its own; not masked out here as it is in Li,
    // direct.cpp:58-61) is shaded with it. Reproduced for parity (seen by running the reference's source: tests/test_ref_render.py).
    const BsdfRec *b1 = its_bsdf(P.S, its1);
    if (its1.valid && (P.S.num_bsdfs == 1 || P.S.num_meshes == 1)) b1 = P.S.meshes[0].bsdf >= 0 ? P.S.bsdfs + P.S.meshes[0].bsdf : nullptr;
    float3 bsdf_val = bsdf_eval(b1, its1, d0_local, true);
    const float correction = fabsf((its1.wi.z * dot(d0, its1.n)) / (d0_local.z * dot(dir, its1.n)));
    bsdf_val = bsdf_val * correction;
    float3 value0 = bsdf_val * emitter_Le(P.S, its2, true) * (base_v * sds.sensor_val / bss.pdf);
    if (guide_out) {
        value0 = zero_nonfinite(value0);
        const float v = hmax(value0) / (float)guide_spc;
        if (v != 0.f) atomicAdd(guide_out + (int)((P.local0 + i) / guide_spc), v);
        return;
    }
    const float3 n = normalize(cross(bss.n, proj));
    value0 = value0 * (copysignf(1.f, dot(e, bss.edge2)) * copysignf(1.f, dot(e, n)));
    // normal velocity: u2 = detach(v0) + u detach(e1) + v detach(e2), (u,v) from the differentiable shadow ray x1 -> edge point
    const TriFull te = [&] {
        const float4 *q = reinterpret_cast<const float4 *>(P.S.tri + h2.tri);
        const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2);
        TriFull t; t.p0 = f3(q0); t.e1 = f3(q1); t.e2 = f3(q2); t.flags = __float_as_int(q2.w); t.area = q0.w;
        return t;
    }();
    const float3 x1 = itc.p;
    const float3 sd_raw = p0 - x1, sd = normalize(sd_raw);
    float u, v, t;
    ray_intersect_triangle(te.p0, te.e1, te.e2, x1, sd, u, v, t);
    const float3 u2 = te.p0 + te.e1 * u + te.e2 * v;
    const float nv = pdot(n, u2);
    const bool jvp = P.S.tri_tangent != nullptr;
    float W = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float prim = getc(value0, c) * nv;
        if (isfinite(prim)) {
            float wv = getc(value0, c) * inv_sppse;
            if (guide_pdf > kEpsilon) wv /= guide_pdf;
            W += (jvp ? (c == P.S.jvp_channel ? 1.f : 0.f) : __ldg(dLdI + 3 * (size_t)sds.pixel + c)) * wv;
        }
    }
    if (W == 0.f || !geom_mode(P.S)) return;
    if (jvp) P.S.jvp_acc[i] = 0.f;
    // d(n . u2): only (u, v) carry derivatives
    const float3 g_u2 = n * W;
    const RayTriGrad rt = ray_intersect_triangle_vjp(te.p0, te.e1, te.e2, x1, sd, pdot(g_u2, te.e1), pdot(g_u2, te.e2), 0.f);
    if (te.flags & 8) { TriGrad g; g.p0 = rt.p0; g.e1 = rt.e1; g.e2 = rt.e2; tri_grad_scatter(P.S, h2.tri, g); }
    const float3 g_x = normalize_vjp(sd_raw, rt.d);           // sd = normalize(p0 - x1)
    const float3 g_x1 = rt.o - g_x, g_p0 = g_x;
    // x1 = cam_o + t_c cam_d with t_c from the camera ray's triangle (scene.cpp:357,366)
    {
        const float4 *q = reinterpret_cast<const float4 *>(P.S.tri + hc.tri);
        const float4 q0 = ldg4(q), q1 = ldg4(q + 1), q2 = ldg4(q + 2);
        float jcam = 0.f;
        if ((__float_as_int(q2.w) & 8) || P.S.sensor_grad) {
            const RayTriGrad rc = ray_intersect_triangle_vjp(f3(q0), f3(q1), f3(q2), cam_o, cam_d, 0.f, 0.f, pdot(g_x1, cam_d));
            if (__float_as_int(q2.w) & 8) {
                TriGrad g; g.p0 = rc.p0; g.e1 = rc.e1; g.e2 = rc.e2;
                tri_grad_scatter(P.S, hc.tri, g);
            }
            jcam = sensor_ray_adjoint(P.S, P.cam, sds.qx, sds.qy, rc.o + g_x1, rc.d + g_x1 * itc.t);   // the camera ray moves with the sensor pose
        }
        if (jvp && jcam != 0.f) P.S.jvp_acc[i] += jcam;
    }
    // edge point p0 = (1-s) v0 + s v1
    float *gworld = Q.mesh_gworld[bss.info.mesh];
    float jsum = 0.f;
    if (gworld) {
        jsum += vertex_adjoint(P.S, gworld, bss.info.v0, g_p0 * (1.f - bss.s));
        jsum += vertex_adjoint(P.S, gworld, bss.info.v1, g_p0 * bss.s);
    }
    if (jvp) {
        jsum += P.S.jvp_acc[i];
        if (jsum != 0.f && isfinite(jsum)) atomicAdd(P.S.jvp_image + 3 * (size_t)sds.pixel + P.S.jvp_channel, jsum);
    }
}

static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

void launch_edge_primary_rays(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, int side, RayRec *rays) {
    if (P.n > 0) k_edge_primary_rays<<<nblk(P.n, 256), 256, 0, st>>>(P, Q, side, rays);
}
void launch_edge_primary_grad(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, const float4 *rad_p, const float4 *rad_n, const float *dLdI, float inv_sppe) {
    if (P.n > 0) k_edge_primary_grad<<<nblk(P.n, 256), 256, 0, st>>>(P, Q, rad_p, rad_n, dLdI, inv_sppe);
}
void launch_edge_secondary_rays(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, RayRec *rays, int guide_spc) {
    if (P.n > 0) k_edge_secondary_rays<<<nblk(P.n, 256), 256, 0, st>>>(P, Q, rays, guide_spc);
}
void launch_edge_secondary_camera(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, const RayRec *rays, const HitRec *hits, RayRec *cam_rays, int guide_spc) {
    if (P.n > 0) k_edge_secondary_camera<<<nblk(P.n, 256), 256, 0, st>>>(P, Q, rays, hits, cam_rays, guide_spc);
}
void launch_edge_secondary_eval(cudaStream_t st, const RenderParams &P, const EdgeParams &Q, const RayRec *rays, const HitRec *hits, const RayRec *cam_rays,
                                const HitRec *cam_hits, const float *dLdI, float inv_sppse, float *guide_out, int guide_spc) {
    if (P.n > 0) k_edge_secondary_eval<<<nblk(P.n, 256), 256, 0, st>>>(P, Q, rays, hits, cam_rays, cam_hits, dLdI, inv_sppse, guide_out, guide_spc);
}

}  // namespace pb
