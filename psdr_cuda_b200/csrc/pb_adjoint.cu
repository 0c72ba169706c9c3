// psdr-b200: reverse-mode (VJP) kernels of the interior integral.
//
// The reference differentiates DirectIntegrator::__Li<true> (src/integrator/direct.cpp:47-163) with Enoki's tape and
// ek.backward; here each scattering event has a hand-written adjoint that replays the event from the stored hit
// records (same RNG positions as the forward renderD) and scatters into the flat gradient vector.
//
// Per lane the forward radiance is  L = Le(x0) + sum_k T_k * L_k,  T_{k+1} = T_k * w_k  (L_k: the event's MIS-weighted
// connections, w_k: its continuation weight). With the suffix  S_k = L_k + w_k * S_{k+1}  the sensitivity of L to the
// parameters touched by event k is  T_k * (dL_k + dw_k * S_{k+1}); the adjoint kernels run k = D-1 .. 0 carrying S.
#include "pb_env_adjoint.cuh"
#include "pb_rc.cuh"
#include "pb_trace.cuh"
#include "pb_wavefront.cuh"

namespace pb {

constexpr int kMaxConstBsdf = 64;

// adjoint of bsdf_eval w.r.t. its textures; `g` is dLoss/d(value). Constant (1x1) textures accumulate into the
// thread-private `acc` (flushed per block), bitmap textures scatter with atomics (Bitmap::eval's backward,
// src/core/bitmap.cpp:43-89: scatter_add into the four texels).
PB_D void bsdf_eval_grad_tex(const SceneView &S, const BsdfRec *b, const Its &its, float3 wo, float3 g, float3 &acc) {
    if (!b) return;
    const float cos_i = its.wi.z, cos_o = wo.z;
    if (!(cos_i > 0.f && cos_o > 0.f)) return;
    if (b->type == BSDF_DIFFUSE) {
        const TexRef &t = b->tex[TEX_REFLECTANCE];
        if (!t.grad) return;
        const float3 gr = g * (kInvPi * cos_o);
        if (!finite3(gr)) return;   // degenerate sample (e.g. zero pdf): no contribution rather than a poisoned gradient
        if (t.w == 1 && t.h == 1) { acc += gr; return; }
        const TexTap tap = tex_tap(t, its.uv);
        const float w[4] = {tap.w0y * tap.w0x, tap.w0y * tap.w1x, tap.w1y * tap.w0x, tap.w1y * tap.w1x};
        const int idx[4] = {tap.idx, tap.idx + 1, tap.idx + t.w, tap.idx + t.w + 1};
        if (S.tri_tangent) {   // forward mode: t.grad holds the texture's tangent
            float sacc = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                sacc += w[k] * (gr.x * __ldg(t.grad + idx[k] * 3) + gr.y * __ldg(t.grad + idx[k] * 3 + 1) + gr.z * __ldg(t.grad + idx[k] * 3 + 2));
            jvp_add(S, sacc);
            return;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            atomicAdd(t.grad + idx[k] * 3 + 0, gr.x * w[k]);
            atomicAdd(t.grad + idx[k] * 3 + 1, gr.y * w[k]);
            atomicAdd(t.grad + idx[k] * 3 + 2, gr.z * w[k]);
        }
    }
    // TODO(roughconductor): alpha_u/alpha_v/eta/k/specular_reflectance adjoints
}

// adjoint of the texture coordinate through one bitmap lookup (Bitmap::eval<ad> is attached to uv, bitmap.cpp:43-89): gv[c] is
// dLoss/d(value_c). At the camera vertex the barycentrics — hence uv — move with the geometry (scene.cpp:355-376).
template <int C> PB_D float2 tex_uv_adjoint(const TexRef &t, float2 uv, const float *gv, bool flip_v = true) {
    if (t.w == 1 && t.h == 1) return make_float2(0.f, 0.f);
    const TexTap tap = tex_tap(t, uv, flip_v);
    const float *d = t.data;
    const int i = tap.idx;
    float gx = 0.f, gy = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        if (gv[c] == 0.f) continue;
        const float v00 = __ldg(d + i * C + c), v10 = __ldg(d + (i + 1) * C + c), v01 = __ldg(d + (i + t.w) * C + c), v11 = __ldg(d + (i + t.w + 1) * C + c);
        gx += gv[c] * (tap.w0y * (v10 - v00) + tap.w1y * (v11 - v01));
        gy += gv[c] * (tap.w0x * (v01 - v00) + tap.w1x * (v11 - v10));
    }
    return make_float2(gx * (float)(t.w - 1), gy * (float)(t.h - 1) * (flip_v ? -1.f : 1.f));
}
PB_D bool bsdf_has_bitmap(const BsdfRec *b) {
    if (!b) return false;
    bool any = false;
#pragma unroll
    for (int k = 0; k < TEX_COUNT; ++k) any = any || (b->tex[k].data != nullptr && b->tex[k].w * b->tex[k].h > 1);
    return any;
}

PB_D float2 rc_uv_adjoint(const BsdfRec *b, float2 uv, const rc::TexGrad &tg) {
    if (!tg.finite()) return make_float2(0.f, 0.f);
    const float e[3] = {tg.eta.x, tg.eta.y, tg.eta.z}, k[3] = {tg.k.x, tg.k.y, tg.k.z}, sp[3] = {tg.spec.x, tg.spec.y, tg.spec.z};
    const float2 a = tex_uv_adjoint<1>(b->tex[TEX_ALPHA_U], uv, &tg.au), c = tex_uv_adjoint<1>(b->tex[TEX_ALPHA_V], uv, &tg.av),
                 d = tex_uv_adjoint<3>(b->tex[TEX_ETA], uv, e), f = tex_uv_adjoint<3>(b->tex[TEX_K], uv, k), g = tex_uv_adjoint<3>(b->tex[TEX_SPECULAR], uv, sp);
    return make_float2(a.x + c.x + d.x + f.x + g.x, a.y + c.y + d.y + f.y + g.y);
}

// block-level reduction of the per-thread constant-texture accumulators, keyed by BSDF id
PB_D void flush_const_tex_grad(const SceneView &S, int bsdf_id, float3 acc, float *s_acc) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int t = threadIdx.x; t < kMaxConstBsdf * 3; t += blockDim.x) s_acc[t] = 0.f;
    __syncthreads();
    const bool has = bsdf_id >= 0 && (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f);
    unsigned remaining = __ballot_sync(full, has);
    while (remaining) {
        const int leader = __ffs(remaining) - 1;
        const int key = __shfl_sync(full, bsdf_id, leader);
        const unsigned grp = __ballot_sync(full, has && bsdf_id == key);
        float3 v = (has && bsdf_id == key) ? acc : f3(0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v.x += __shfl_xor_sync(full, v.x, o); v.y += __shfl_xor_sync(full, v.y, o); v.z += __shfl_xor_sync(full, v.z, o);
        }
        if (lane == leader) {
            if (key < kMaxConstBsdf) {
                atomicAdd(s_acc + key * 3, v.x); atomicAdd(s_acc + key * 3 + 1, v.y); atomicAdd(s_acc + key * 3 + 2, v.z);
            } else {
                float *gp = S.bsdfs[key].tex[TEX_REFLECTANCE].grad;
                atomicAdd(gp, v.x); atomicAdd(gp + 1, v.y); atomicAdd(gp + 2, v.z);
            }
        }
        remaining &= ~grp;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kMaxConstBsdf * 3; t += blockDim.x) {
        const float v = s_acc[t];
        if (v != 0.f) {
            const int b = t / 3;
            if (b < S.num_bsdfs) { float *gp = S.bsdfs[b].tex[TEX_REFLECTANCE].grad; if (gp) atomicAdd(gp + (t - 3 * b), v); }
        }
    }
}

// Reflectance adjoint of one lane's event from the linearisation the forward pass kept (EventBuffers::lin): with L_k = rho * A and
// w_k = rho * c the event's share of dLoss/d(rho) is gL * A + gw * c and the suffix is S_k = rho * (A + c S_{k+1}) — the same sums
// k_adjoint forms connection by connection (bsdf_eval_grad_tex), without reconstructing the connections.
template <int EV>
PB_D void adjoint_lin_lane(const RenderParams &P, const BounceParams &B, const EventBuffers &E, int i, float4 *__restrict__ suffix, const float *__restrict__ dLdI,
                           float3 &acc, int &bsdf_id) {
    const float4 lin = ldg4(E.lin + i);
    const float3 A = f3(lin);
    const float3 S_next = B.last ? f3(0.f) : f3(suffix[i]);
    float3 Sk = f3(0.f);
    if (A.x != 0.f || A.y != 0.f || A.z != 0.f || lin.w != 0.f) {   // (NaN compares unequal: a poisoned event is kept and poisons S_k like k_adjoint's)
        const float4 vc = ldg4(E.vc + i);   // the vertex record's (uv, mesh id)
        const int shape = __float_as_int(vc.z);
        const int bid = shape >= 0 ? P.S.meshes[shape].bsdf : -1;
        if (bid >= 0) {
            const BsdfRec *bsdf = P.S.bsdfs + bid;
            const float2 uv = make_float2(vc.x, vc.y);
            const TexRef &t = bsdf->tex[TEX_REFLECTANCE];
            Sk = tex_eval3(t, uv) * (A + S_next * lin.w);
            if (t.grad) {
                int pix;
                global_lane(P, i, pix);
                const float3 rad_final = f3(ldg4(E.rad + i));
                float3 g = f3(__ldg(dLdI + 3 * (size_t)pix), __ldg(dLdI + 3 * (size_t)pix + 1), __ldg(dLdI + 3 * (size_t)pix + 2)) * P.inv_spp;
                if (!isfinite(rad_final.x)) g.x = 0.f;
                if (!isfinite(rad_final.y)) g.y = 0.f;
                if (!isfinite(rad_final.z)) g.z = 0.f;
                float3 T = f3(1.f);
                if (!ev_depth0<EV>(B)) T = f3(ldg4(E.thr_in + i));
                const float3 gL = g * T;
                const float3 gr = gL * A + gL * S_next * lin.w;
                if (finite3(gr)) {
                    if (t.w == 1 && t.h == 1) {
                        acc = gr;
                        bsdf_id = bid;
                    } else {
                        const TexTap tap = tex_tap(t, uv);
                        const float w[4] = {tap.w0y * tap.w0x, tap.w0y * tap.w1x, tap.w1y * tap.w0x, tap.w1y * tap.w1x};
                        const int idx[4] = {tap.idx, tap.idx + 1, tap.idx + t.w, tap.idx + t.w + 1};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            atomicAdd(t.grad + idx[k] * 3 + 0, gr.x * w[k]);
                            atomicAdd(t.grad + idx[k] * 3 + 1, gr.y * w[k]);
                            atomicAdd(t.grad + idx[k] * 3 + 2, gr.z * w[k]);
                        }
                    }
                }
            }
        }
    }
    if (!ev_depth0<EV>(B)) suffix[i] = make_float4(Sk.x, Sk.y, Sk.z, 0.f);
}
// Geometry adjoints of a diffuse scene (reverse mode): an event touches the triangle table only through its own vertex and the end points of its
// connections. If none of them lies on a mesh whose vertices are a leaf (MeshRec::flags bit 2), the event's adjoint is the reflectance one above.
PB_D bool event_touches_leaf_mesh(const RenderParams &P, const BounceParams &B, const EventBuffers &E, int i) {
    auto leaf_mesh = [&](int shape) { return shape >= 0 && (P.S.meshes[shape].flags & 4) != 0; };
    bool need = leaf_mesh(__float_as_int(__ldg(&E.vc[i].z)));
    for (int j = 0; j < B.nb + B.nl; ++j) {
        const int shape = __ldg(reinterpret_cast<const int *>(E.hits + (size_t)j * P.n + i) + 1);
        // the end of a BSDF-sampled ray is the next vertex; an emitter-sampled connection only counts when it reaches an emitter
        need = need || (leaf_mesh(shape) && (j < B.nb || P.S.meshes[shape].emitter >= 0));
    }
    return need;
}

// adjoint of one scattering event (texture parameters). E.thr_in: T_k; suffix: S_{k+1} in, S_k out; E.rad: the lane's
// final forward radiance (decides which channels integrator.cpp:87 zeroed).
// RC ("extended" events): rough-conductor texture / geometry adjoints and the environment map's radiance / scale / direction
// adjoints (separate instantiation so that the diffuse + area-light kernel keeps its registers)
template <int MINB, bool PREFETCH, bool RC, bool SIMPLE, int EV>
__global__ void __launch_bounds__(256, MINB) k_adjoint(RenderParams P, BounceParams B, EventBuffers E, float4 *__restrict__ suffix, const float *__restrict__ dLdI) {
    const HitRec *__restrict__ hits = E.hits;
    __shared__ float s_acc[kMaxConstBsdf * 3];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (E.lane_list) {   // behind k_adjoint_split: the lanes whose event touches a leaf mesh (the whole block leaves together past the end of the list)
        const unsigned cnt = __ldg(E.lane_count);
        if (blockIdx.x * blockDim.x >= cnt) return;
        i = (unsigned)i < cnt ? __ldg(E.lane_list + i) : P.n;
    }
    float3 acc = f3(0.f);
    int bsdf_id = -1;
    rc::TexGrad rc_acc;
    bool rc_tex = false;
    float env_scale_acc = 0.f;
    __shared__ float s_rc[RC ? rc::kFlushFloats : 1];
    if (i < P.n) {
        if (PREFETCH) prefetch_event_hits(P.S, hits, B.nb + B.nl, P.n, i);
        int pix;
        const long long lane = global_lane(P, i, pix);
        const Vertex v = load_vertex<EV>(P, B, i, E);
        const Its &its = v.its;
        bsdf_id = v.bsdf ? (int)(v.bsdf - P.S.bsdfs) : -1;
        // loss adjoint of this lane's radiance
        const float3 rad_final = f3(ldg4(E.rad + i));
        float3 g;
        if (P.S.tri_tangent) g = f3(P.S.jvp_channel == 0 ? 1.f : 0.f, P.S.jvp_channel == 1 ? 1.f : 0.f, P.S.jvp_channel == 2 ? 1.f : 0.f) * P.inv_spp;
        else g = f3(__ldg(dLdI + 3 * (size_t)pix), __ldg(dLdI + 3 * (size_t)pix + 1), __ldg(dLdI + 3 * (size_t)pix + 2)) * P.inv_spp;
        if (!isfinite(rad_final.x)) g.x = 0.f;
        if (!isfinite(rad_final.y)) g.y = 0.f;
        if (!isfinite(rad_final.z)) g.z = 0.f;
        float3 T = f3(1.f);
        if (!ev_depth0<EV>(B)) T = f3(ldg4(E.thr_in + i));
        const float3 S_next = B.last ? f3(0.f) : f3(suffix[i]);
        const float3 gL = g * T, gw = gL * S_next;
        Rng rng = make_rng(P, lane, B.jump);
        float3 L = f3(0.f), w_cont = f3(0.f);
        const float inv_nb = B.nb > 0 ? 1.f / (float)B.nb : 0.f, inv_nl = B.nl > 0 ? 1.f / (float)B.nl : 0.f;
        const bool geom = geom_mode(P.S) && v.active && v.bsdf && v.bsdf->type == BSDF_DIFFUSE;
        float3 g_p = f3(0.f), g_shn = f3(0.f);   // adjoints of this vertex' position and shading normal
        const bool env_on = RC && P.S.emitter_env >= 0 && (env_wants_grad(P.S) || geom_mode(P.S));
        float3 g_rd_env = f3(0.f);   // Le(x0), direct.cpp:51: its direction is the camera ray's, which only the sensor pose moves
        if (RC && ev_depth0<EV>(B) && !B.hide_emitters && env_on)
            env_le_vjp(P.S, its, v.ro, g, P.S.sensor_grad != nullptr, env_scale_acc, &g_rd_env);
        float2 g_uv = make_float2(0.f, 0.f);   // adjoint of the camera vertex' texture coordinate (bitmap textures only)
        float *uv_leaf = (RC && v.active) ? P.S.meshes[its.shape].uv_grad : nullptr;   // Mesh.vertex_uv of this vertex' mesh is a leaf
        const bool uv_geom = RC && v.active && bsdf_has_bitmap(v.bsdf) && ((ev_depth0<EV>(B) && geom_mode(P.S)) || uv_leaf != nullptr);
        rc::Tex rtex;
        bool geom_rc = false;       // geometry adjoints of a rough-conductor vertex (local duals, pb_rc.cuh)
        float3 g_a = f3(0.f);       // adjoint of the previous vertex' position (enters through wi)
        if (RC) {
            rc_tex = v.active && rc::wants_tex_grad(v.bsdf);
            geom_rc = geom_mode(P.S) && v.active && v.bsdf && v.bsdf->type == BSDF_ROUGHCONDUCTOR;
            if (rc_tex || geom_rc || (uv_geom && v.bsdf->type == BSDF_ROUGHCONDUCTOR)) rtex = rc::load_tex(v.bsdf, its.uv);
        }
        for (int j = 0; j < B.nb; ++j) {
            const float3 s3 = rng.next_3d();
            const BsdfSample bs = bsdf_sample<SIMPLE>(v.bsdf, its, s3, v.active);
            bool a1 = v.active && bs.valid;
            const HitRec h1 = load_hit(hits + (size_t)j * P.n + i);
            const Its its1 = reconstruct_its(P.S, h1, its.p);
            a1 = a1 && its1.valid;
            const bool cont = a1 && B.carry && j == 0;
            a1 = a1 && is_emitter(P.S, its1.shape);
            if (a1 || cont) {
                float3 wo = its1.p - its.p;
                wo = wo / its1.t;
                const float3 wo_l = its.sh.to_local(wo);
                const float3 f = bsdf_eval<SIMPLE>(v.bsdf, its, wo_l, true);
                const float G = fabsf(dot(its1.n, -wo)) / sqr(its1.t);
                const float pdf0 = bs.pdf * G;
                const float scale = G / pdf0;
                float3 gval = f3(0.f);
                if (a1) {
                    float weight = inv_nb;
                    if (B.nl > 0) weight *= mis_weight(pdf0, emitter_position_pdf<SIMPLE>(P.S, its.p, its1, true));
                    const float3 Le = emitter_Le<SIMPLE>(P.S, its1, true);
                    L += Le * f * (scale * weight);
                    gval += gL * Le * (scale * weight);
                }
                if (cont) { w_cont = f * scale; gval += gw * scale; }
                bsdf_eval_grad_tex(P.S, v.bsdf, its, wo_l, gval, acc);
                if (RC && uv_geom && v.bsdf->type == BSDF_DIFFUSE && wo_l.z > 0.f && its.wi.z > 0.f) {
                    const float3 gr = gval * (kInvPi * wo_l.z);
                    if (finite3(gr)) { const float gv3[3] = {gr.x, gr.y, gr.z}; const float2 a = tex_uv_adjoint<3>(v.bsdf->tex[TEX_REFLECTANCE], its.uv, gv3); g_uv.x += a.x; g_uv.y += a.y; }
                }
                if (RC && a1 && env_on) {   // dLoss/dLe of this connection
                    float weight = inv_nb;
                    if (B.nl > 0) weight *= mis_weight(pdf0, emitter_position_pdf<SIMPLE>(P.S, its.p, its1, true));
                    g_p += env_le_vjp(P.S, its1, its.p, gL * f * (scale * weight), geom_mode(P.S), env_scale_acc);
                }
                if (RC && (rc_tex || (uv_geom && v.bsdf->type == BSDF_ROUGHCONDUCTOR))) {
                    const float p_em = (a1 && B.nl > 0) ? emitter_position_pdf<SIMPLE>(P.S, its.p, its1, true) : 0.f;
                    rc::TexGrad tg;
                    rc::bsdf_branch_tex_grad(rtex, its.wi, wo_l, s3, G, p_em, a1 && B.nl > 0, inv_nb, a1 ? gL * emitter_Le<SIMPLE>(P.S, its1, true) : f3(0.f),
                                             cont ? gw : f3(0.f), tg);
                    if (rc_tex) rc::emit_tex_grad(P.S, v.bsdf, its.uv, tg, rc_acc);
                    if (uv_geom) { const float2 a = rc_uv_adjoint(v.bsdf, its.uv, tg); g_uv.x += a.x; g_uv.y += a.y; }
                }
                if (RC && geom_rc) {
                    const float p_em = (a1 && B.nl > 0) ? emitter_position_pdf<SIMPLE>(P.S, its.p, its1, true) : 0.f;
                    rc::GeomGrad gg;
                    if (rc::branch_geom_grad(rtex, its.p, its.sh.n, ev_depth0<EV>(B) ? v.rd : v.ro, its1.p, its1.n, ev_depth0<EV>(B), v.rd, false, square_to_uniform_disk_concentric(s3.x, s3.y),
                                             p_em, a1 && B.nl > 0, inv_nb, a1 ? gL * emitter_Le<SIMPLE>(P.S, its1, true) : f3(0.f), cont ? gw : f3(0.f), gg)) {
                        g_p += gg.p; g_shn += gg.shn; g_a += gg.a;
                        point_on_triangle_scatter(P.S, its1.tri, h1.u, h1.v, gg.q, gg.nq, gg.c0);
                    }
                }
                if (geom && wo_l.z > 0.f && its.wi.z > 0.f) {
                    // value = K * (cos_o G J) with K = rho/pi * (Le weight gL + gw) / pdf0   (diffuse: pdf0 and the MIS weight are detached)
                    const float3 rho = tex_eval3(v.bsdf->tex[TEX_REFLECTANCE], its.uv);
                    float3 kk = f3(0.f);
                    if (a1) {
                        float weight = inv_nb;
                        if (B.nl > 0) weight *= mis_weight(pdf0, emitter_position_pdf<SIMPLE>(P.S, its.p, its1, true));
                        kk += gL * emitter_Le<SIMPLE>(P.S, its1, true) * weight;
                    }
                    if (cont) kk += gw;
                    const float gc = pdot(kk, rho) * kInvPi / pdf0;
                    if (gc != 0.f && isfinite(gc)) {
                        const ConnGrad cg = connection_vjp(its.p, its1.p, its.sh.n, its1.n, 1.f, gc);
                        if (finite3(cg.p) && finite3(cg.sh_n)) { g_p += cg.p; g_shn += cg.sh_n; }
                        point_on_triangle_scatter(P.S, its1.tri, h1.u, h1.v, cg.q, cg.n_q, cg.J);
                    }
                }
            }
        }
        for (int j = 0; j < B.nl; ++j) {
            const float2 s2 = rng.next_2d();
            const PositionSample ps = sample_emitter_position<SIMPLE>(P.S, v.its.p, s2, v.active);
            bool a1 = v.active && ps.valid;
            float3 wo = ps.p - its.p;
            const float dist_sqr = squared_norm(wo);
            const float dist = safe_sqrt(dist_sqr);
            wo = wo / dist;
            const HitRec h1 = load_hit(hits + (size_t)(B.nb + j) * P.n + i);
            const Its its1 = reconstruct_its(P.S, h1, its.p);
            a1 = a1 && its1.valid && (its1.t > dist - kShadowEpsilon) && is_emitter(P.S, its1.shape);
            if (a1) {
                const float G = fabsf(dot(its1.n, -wo)) / dist_sqr;
                const float3 wo_l = its.sh.to_local(wo);
                const float3 f = bsdf_eval<SIMPLE>(v.bsdf, its, wo_l, true);
                const float pdf1 = bsdf_pdf<SIMPLE>(v.bsdf, its, wo_l, true) * G;
                float weight = inv_nl;
                if (B.nb > 0) weight *= mis_weight(ps.pdf, pdf1);
                const float3 Le = emitter_Le<SIMPLE>(P.S, its1, true);
                const float scale = G / ps.pdf * weight;
                L += Le * f * scale;
                bsdf_eval_grad_tex(P.S, v.bsdf, its, wo_l, gL * Le * scale, acc);
                if (RC && uv_geom && v.bsdf->type == BSDF_DIFFUSE && wo_l.z > 0.f && its.wi.z > 0.f) {
                    const float3 gr = gL * Le * (scale * kInvPi * wo_l.z);
                    if (finite3(gr)) { const float gv3[3] = {gr.x, gr.y, gr.z}; const float2 a = tex_uv_adjoint<3>(v.bsdf->tex[TEX_REFLECTANCE], its.uv, gv3); g_uv.x += a.x; g_uv.y += a.y; }
                }
                if (RC && env_on) g_p += env_le_vjp(P.S, its1, its.p, gL * f * scale, geom_mode(P.S), env_scale_acc);
                if (RC && (rc_tex || (uv_geom && v.bsdf->type == BSDF_ROUGHCONDUCTOR))) {
                    rc::TexGrad tg;
                    rc::light_branch_tex_grad(rtex, its.wi, wo_l, G, ps.pdf, B.nb > 0, inv_nl, gL * Le, tg);
                    if (rc_tex) rc::emit_tex_grad(P.S, v.bsdf, its.uv, tg, rc_acc);
                    if (uv_geom) { const float2 a = rc_uv_adjoint(v.bsdf, its.uv, tg); g_uv.x += a.x; g_uv.y += a.y; }
                }
                if (RC && geom_rc) {
                    rc::GeomGrad gg;
                    if (rc::branch_geom_grad(rtex, its.p, its.sh.n, ev_depth0<EV>(B) ? v.rd : v.ro, ps.p, its1.n, ev_depth0<EV>(B), v.rd, true, make_float2(0.f, 0.f), ps.pdf, B.nb > 0, inv_nl,
                                             gL * Le, f3(0.f), gg)) {
                        g_p += gg.p; g_shn += gg.shn; g_a += gg.a;
                        if (ps.tri >= 0) point_on_triangle_scatter(P.S, ps.tri, ps.s, ps.t, gg.q, f3(0.f), gg.c0);   // area-light sample + its Jacobian; envmap samples are detached
                        point_on_triangle_scatter(P.S, its1.tri, h1.u, h1.v, f3(0.f), gg.nq, 0.f);
                    }
                }
                if (geom && wo_l.z > 0.f && its.wi.z > 0.f) {
                    const float3 rho = tex_eval3(v.bsdf->tex[TEX_REFLECTANCE], its.uv);
                    const float gc = pdot(gL * Le, rho) * kInvPi * weight / ps.pdf;
                    if (gc != 0.f && isfinite(gc)) {
                        const ConnGrad cg = connection_vjp(its.p, ps.p, its.sh.n, its1.n, 1.f, gc);
                        if (finite3(cg.p) && finite3(cg.sh_n)) { g_p += cg.p; g_shn += cg.sh_n; }
                        if (ps.tri >= 0) point_on_triangle_scatter(P.S, ps.tri, ps.s, ps.t, cg.q, f3(0.f), cg.J);     // sampled point + its Jacobian (mesh.cpp:317-328); envmap samples are detached (envmap.cpp:72-95)
                        point_on_triangle_scatter(P.S, its1.tri, h1.u, h1.v, f3(0.f), cg.n_q, 0.f);   // normal of the triangle the shadow ray hit
                    }
                }
            }
        }
        if (!ev_depth0<EV>(B)) {
            const float3 Sk = L + w_cont * S_next;
            suffix[i] = make_float4(Sk.x, Sk.y, Sk.z, 0.f);
        }
        if (RC && geom_rc && !ev_depth0<EV>(B) && E.hit_prev && (g_a.x != 0.f || g_a.y != 0.f || g_a.z != 0.f) && finite3(g_a)) {
            // the previous vertex moves wi: chain into its triangle (path-space point, or the camera hit in solid-angle form)
            const HitRec hp = load_hit(E.hit_prev + i);
            if (hp.tri >= 0) {
                if (B.depth == 1) {
                    const TriFull t = load_tri_full(P.S, hp.tri);
                    if ((t.flags & 8) || P.S.sensor_grad) {
                        int pix0;
                        Rng rng0 = make_rng(P, global_lane(P, i, pix0), P.jump0);
                        const float2 jit = rng0.next_2d();
                        float sx, sy;
                        lane_pixel_sample(P, pix0, jit, sx, sy);
                        float3 o, d;
                        sample_primary_ray(P.cam, sx, sy, o, d);
                        const RayTriGrad r = ray_intersect_triangle_vjp(t.p0, t.e1, t.e2, o, d, 0.f, 0.f, pdot(g_a, d));
                        if (t.flags & 8) {
                            TriGrad tg;
                            tg.p0 = r.p0; tg.e1 = r.e1; tg.e2 = r.e2;
                            tri_grad_scatter(P.S, hp.tri, tg);
                        }
                        if (P.S.sensor_grad) {   // the camera hit is o + t d
                            const float jv = sensor_ray_adjoint(P.S, P.cam, sx, sy, r.o + g_a, r.d + g_a * norm(v.ro - o));
                            if (P.S.tri_tangent) jvp_add(P.S, jv);
                        }
                    }
                } else {
                    point_on_triangle_scatter(P.S, hp.tri, hp.u, hp.v, g_a, f3(0.f), 0.f);
                }
            }
        }
        if (RC && uv_leaf && (g_uv.x != 0.f || g_uv.y != 0.f) && isfinite(g_uv.x) && isfinite(g_uv.y)) {
            // uv = (1 - u - v) uv0 + u uv1 + v uv2 with the triangle's three uv vertices (mesh.cpp:232-236)
            const MeshRec &mr = P.S.meshes[its.shape];
            const int f = its.tri - mr.face_offset;
            float bu = v.h.u, bv = v.h.v;
            if (ev_depth0<EV>(B)) { float tt; ray_intersect_triangle(load_tri_geom(P.S, its.tri).p0, load_tri_geom(P.S, its.tri).e1, load_tri_geom(P.S, its.tri).e2, v.ro, v.rd, bu, bv, tt); }
            const float wgt[3] = {1.f - bu - bv, bu, bv};
            float jv = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int j = __ldg(mr.uv_faces + 3 * f + k);
                if (P.S.tri_tangent) jv += wgt[k] * (g_uv.x * __ldg(uv_leaf + 2 * j) + g_uv.y * __ldg(uv_leaf + 2 * j + 1));
                else { atomicAdd(uv_leaf + 2 * j, wgt[k] * g_uv.x); atomicAdd(uv_leaf + 2 * j + 1, wgt[k] * g_uv.y); }
            }
            if (P.S.tri_tangent) jvp_add(P.S, jv);
        }
        const bool want_cam = RC && ev_depth0<EV>(B) && P.S.sensor_grad != nullptr && its.valid;
        if (geom || (RC && geom_rc) || want_cam) {   // chain the vertex adjoints into its triangle (scene.cpp:326-376) and, at the camera vertex, into the sensor pose
            const TriFull t = load_tri_full(P.S, its.tri);
            const bool want_tri = (t.flags & 8) != 0;
            if (want_tri || want_cam) {
                TriGrad tg;
                float gu = 0.f, gv = 0.f;
                if (t.flags & 1) tg.fn += g_shn;
                else shading_normal_vjp(t.n0, t.n1, t.n2, v.h.u, v.h.v, g_shn, tg, gu, gv);
                if (ev_depth0<EV>(B)) {   // solid-angle form: (u, v, t) come from the differentiable ray/triangle test, p = o + t d
                    if (RC && uv_geom && (t.flags & 2) && isfinite(g_uv.x) && isfinite(g_uv.y)) {   // uv = uv0 + u (uv1 - uv0) + v (uv2 - uv0)
                        const float4 *tq = reinterpret_cast<const float4 *>(P.S.tri + its.tri);
                        const float u0x = ldg4(tq + 3).w, u0y = ldg4(tq + 4).w, u1x = ldg4(tq + 5).w, u1y = ldg4(tq + 6).w;
                        const float4 q7 = ldg4(tq + 7);
                        gu += g_uv.x * (u1x - u0x) + g_uv.y * (u1y - u0y);
                        gv += g_uv.x * (q7.x - u0x) + g_uv.y * (q7.y - u0y);
                    }
                    const RayTriGrad r = ray_intersect_triangle_vjp(t.p0, t.e1, t.e2, v.ro, v.rd, gu, gv, pdot(g_p, v.rd));
                    tg.p0 += r.p0; tg.e1 += r.e1; tg.e2 += r.e2;
                    if (want_cam) {
                        const float jv = sensor_ray_adjoint(P.S, P.cam, v.film.x, v.film.y, r.o + g_p, r.d + g_p * its.t + g_a + g_rd_env);
                        if (P.S.tri_tangent) jvp_add(P.S, jv);
                    }
                } else {              // path-space form: barycentrics are frozen, the point rides the triangle
                    tg.p0 += g_p; tg.e1 += g_p * v.h.u; tg.e2 += g_p * v.h.v;
                }
                if (want_tri) tri_grad_scatter(P.S, its.tri, tg);
            }
        }
    }
    if (P.S.tri_tangent) {   // forward mode (uniform over the grid): constant-texture part, then the lane total goes to the derivative image
        int pix = -1;
        float total = 0.f;
        if (i < P.n) {
            if (bsdf_id >= 0) {
                const TexRef &t = P.S.bsdfs[bsdf_id].tex[TEX_REFLECTANCE];
                if (t.grad && t.w == 1 && t.h == 1) jvp_add(P.S, acc.x * __ldg(t.grad) + acc.y * __ldg(t.grad + 1) + acc.z * __ldg(t.grad + 2));
            }
            if (ev_depth0<EV>(B)) { global_lane(P, i, pix); total = P.S.jvp_acc[i]; }
        }
        if (ev_depth0<EV>(B)) film_accumulate1(P.S.jvp_image, pix, P.S.jvp_channel, total);
        return;
    }
    flush_const_tex_grad(P.S, bsdf_id, acc, s_acc);
    if (RC) rc::flush_const_tex_grad(P.S, rc_tex ? bsdf_id : -1, rc_acc, env_scale_acc, s_rc);
}

// adjoint_lin_lane for every lane: diffuse scenes whose only leaves are reflectance textures (no geometry, pose or uv adjoints), reverse mode
template <int EV>
__global__ void __launch_bounds__(256) k_adjoint_lin(RenderParams P, BounceParams B, EventBuffers E, float4 *__restrict__ suffix, const float *__restrict__ dLdI) {
    __shared__ float s_acc[kMaxConstBsdf * 3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 acc = f3(0.f);
    int bsdf_id = -1;
    if (i < P.n) adjoint_lin_lane<EV>(P, B, E, i, suffix, dLdI, acc, bsdf_id);
    flush_const_tex_grad(P.S, bsdf_id, acc, s_acc);
}

// First half of the split geometry adjoint: a lane whose event touches no leaf mesh is finished here from the linearisation; the others are
// listed (order is irrelevant: they only meet in atomics) for k_adjoint, which then runs on full warps of lanes that need it. In lane order a
// warp holds 32 samples of one pixel whose deeper vertices lie anywhere: almost every warp had a few lanes on the bunny and ran the whole kernel.
template <int EV>
__global__ void __launch_bounds__(256) k_adjoint_split(RenderParams P, BounceParams B, EventBuffers E, float4 *__restrict__ suffix, const float *__restrict__ dLdI,
                                                       int *__restrict__ list, unsigned *__restrict__ count) {
    __shared__ float s_acc[kMaxConstBsdf * 3];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 acc = f3(0.f);
    int bsdf_id = -1;
    bool need = false;
    if (i < P.n) {
        need = event_touches_leaf_mesh(P, B, E, i);
        if (!need) adjoint_lin_lane<EV>(P, B, E, i, suffix, dLdI, acc, bsdf_id);
    }
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (m) {
        const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(count, (unsigned)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (need) list[base + __popc(m & ((1u << lane) - 1u))] = i;
    }
    flush_const_tex_grad(P.S, bsdf_id, acc, s_acc);
}

int g_adjoint_lin = 1;
static inline unsigned nblk(long long n, int b) { return (unsigned)((n + b - 1) / b); }

void launch_adjoint(cudaStream_t st, const RenderParams &P, const BounceParams &B, const EventBuffers &E0, float4 *suffix, const float *dLdI,
                    int *split_list, unsigned *split_count) {
    if (P.n <= 0) return;
    const unsigned g = nblk(P.n, 256);
    EventBuffers E = E0;
    E.lane_list = nullptr; E.lane_count = nullptr;
    if (B.rc_grad) { k_adjoint<1, false, true, false, -1><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI); return; }
    if (P.S.simple && g_shade_simple && g_adjoint_lin && E.lin && P.S.tri_grad && !P.S.tri_tangent && !E.inv && split_list && split_count) {
        // vertex leaves on some meshes of a diffuse scene: reflectance adjoint for the lanes that touch none of them, the full kernel for the rest
        cudaMemsetAsync(split_count, 0, sizeof(unsigned), st);
        if (B.depth == 0) k_adjoint_split<3><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI, split_list, split_count);
        else k_adjoint_split<2><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI, split_list, split_count);
        E.lane_list = split_list; E.lane_count = split_count;
        if (B.depth == 0) k_adjoint<3, false, false, true, 3><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        else k_adjoint<3, false, false, true, 2><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        return;
    }
    if (P.S.simple && g_shade_simple && g_adjoint_lin && E.lin && !P.S.tri_grad && !P.S.tri_tangent) {   // reflectance leaves only: from the kept linearisation
        if (B.depth == 0) k_adjoint_lin<3><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        else k_adjoint_lin<2><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        return;
    }
    if (P.S.simple && g_shade_simple) {
        if (g_shade_tune == 6) k_adjoint<3, false, false, true, -1><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        else if (B.depth == 0) k_adjoint<3, false, false, true, 3><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        else k_adjoint<3, false, false, true, 2><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI);
        return;
    }
    switch (g_shade_tune) {
        case 1: k_adjoint<2, true, false, false, -1><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI); break;
        case 4: k_adjoint<2, false, false, false, -1><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI); break;
        case 5: k_adjoint<4, false, false, false, -1><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI); break;
        default: k_adjoint<3, false, false, false, -1><<<g, 256, 0, st>>>(P, B, E, suffix, dLdI); break;
    }
}

}  // namespace pb
