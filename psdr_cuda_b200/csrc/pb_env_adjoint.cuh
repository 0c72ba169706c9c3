// psdr-b200: adjoint of the environment map's radiance lookup, EnvironmentMap::eval / eval_direction<true>
// (src/emitter/envmap.cpp:35-58): Le = bilinear(radiance, uv(from_world * dir)) * scale is attached to the radiance texels
// (EnvironmentMap.radiance.data, src/psdr.cpp:236), to the scale (psdr.cpp:237) and to the direction dir = (q - p)/|q - p| along
// which the emitter is seen, i.e. to the position p of the vertex that looks at it (the bounding mesh point q carries no
// gradient). Sampling and pdfs of the environment map are detached in the reference (envmap.cpp:72-95, 125-143).
#pragma once
#include "pb_wavefront.cuh"

namespace pb {

PB_D bool env_wants_grad(const SceneView &S) {
    if (S.emitter_env < 0) return false;
    const EmitterRec &em = S.emitters[S.emitter_env];
    return em.env_radiance.grad != nullptr || em.env_scale_grad != nullptr || em.env_xf_grad != nullptr;
}

// gLe: dLoss/dLe (per channel). its1: the hit on the bounding mesh, seen from `origin`. Returns the adjoint of `origin`
// (zero unless `want_dir`). The scale's gradient goes to the thread-private `scale_acc` (one address for every lane: the
// kernel reduces it per block before touching global memory).
// `g_dir_out` (optional) receives the adjoint of the looked-up direction itself: the camera ray's direction d = M3 d_cam is not
// re-normalised (perspective.cpp:120-127), so the sensor pose sees the unprojected adjoint.
PB_D float3 env_le_vjp(const SceneView &S, const Its &its1, float3 origin, float3 gLe, bool want_dir, float &scale_acc, float3 *g_dir_out = nullptr) {
    float3 g_origin = f3(0.f);
    if (!its1.valid || !finite3(gLe) || (gLe.x == 0.f && gLe.y == 0.f && gLe.z == 0.f)) return g_origin;
    const int e = S.meshes[its1.shape].emitter;
    if (e < 0) return g_origin;
    const EmitterRec &em = S.emitters[e];
    if (em.type != EMITTER_ENVMAP) return g_origin;
    const float3 dw = -its1.sh.to_world(its1.wi);   // envmap.cpp:36-37
    const float3 v = transform_dir(em.env_from_world, dw);
    float2 uv = make_float2(atan2f(v.x, -v.z) * kInvTwoPi, safe_acos(v.y) * kInvPi);
    uv.x -= floorf(uv.x); uv.y -= floorf(uv.y);
    const TexRef &t = em.env_radiance;
    const TexTap tap = tex_tap(t, uv, false);
    const int idx[4] = {tap.idx, tap.idx + 1, tap.idx + t.w, tap.idx + t.w + 1};
    const float w[4] = {tap.w0y * tap.w0x, tap.w0y * tap.w1x, tap.w1y * tap.w0x, tap.w1y * tap.w1x};
    float tex[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) tex[k][c] = __ldg(t.data + idx[k] * 3 + c);
    const float g[3] = {gLe.x, gLe.y, gLe.z};
    const bool fwd = S.tri_tangent != nullptr;
    float jv = 0.f;
    // radiance texels
    if (t.grad) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float a = g[c] * em.env_scale * w[k];
                if (fwd) jv = fmaf(a, __ldg(t.grad + idx[k] * 3 + c), jv);
                else if (a != 0.f) atomicAdd(t.grad + idx[k] * 3 + c, a);
            }
    }
    // scale
    if (em.env_scale_grad) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) s += g[c] * (w[0] * tex[0][c] + w[1] * tex[1][c] + w[2] * tex[2][c] + w[3] * tex[3][c]);
        if (fwd) jv = fmaf(s, __ldg(em.env_scale_grad), jv);
        else if (isfinite(s)) scale_acc += s;
    }
    // direction -> the looking vertex; from_world -> the environment map's transform (v = F dw)
    if (want_dir || em.env_xf_grad) {
        float gu = 0.f, gv = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float du = tap.w0y * (tex[1][c] - tex[0][c]) + tap.w1y * (tex[3][c] - tex[2][c]);
            const float dv = tap.w0x * (tex[2][c] - tex[0][c]) + tap.w1x * (tex[3][c] - tex[1][c]);
            gu += g[c] * du; gv += g[c] * dv;
        }
        gu *= em.env_scale * (float)(t.w - 1); gv *= em.env_scale * (float)(t.h - 1);
        const float r2 = v.x * v.x + v.z * v.z;
        float3 gvec = f3(0.f);
        if (r2 > 0.f) { gvec.x = gu * kInvTwoPi * (-v.z) / r2; gvec.z = gu * kInvTwoPi * v.x / r2; }
        if (fabsf(v.y) < 1.f) gvec.y = -gv * kInvPi / sqrtf(1.f - v.y * v.y);
        if (em.env_xf_grad && finite3(gvec)) {   // g_F[i][j] = g_v[i] dw[j]
            const float gF[16] = {gvec.x * dw.x, gvec.x * dw.y, gvec.x * dw.z, 0.f, gvec.y * dw.x, gvec.y * dw.y, gvec.y * dw.z, 0.f,
                                  gvec.z * dw.x, gvec.z * dw.y, gvec.z * dw.z, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (fwd) {
#pragma unroll
                for (int k = 0; k < 12; ++k) jv = fmaf(gF[k], __ldg(em.env_xf_grad + k), jv);
            } else {
                sensor_atomic_add16(em.env_xf_grad, gF);
            }
        }
        // dw = from_world^T-applied adjoint: v = M dw  =>  g_dw = M^T g_v
        const Mat4 &M = em.env_from_world;
        if (fwd && jv != 0.f && isfinite(jv)) { S.jvp_acc[blockIdx.x * blockDim.x + threadIdx.x] += jv; jv = 0.f; }
        if (!want_dir) return g_origin;
        const float3 g_dw = f3(M.m[0] * gvec.x + M.m[4] * gvec.y + M.m[8] * gvec.z, M.m[1] * gvec.x + M.m[5] * gvec.y + M.m[9] * gvec.z,
                               M.m[2] * gvec.x + M.m[6] * gvec.y + M.m[10] * gvec.z);
        if (g_dir_out && finite3(g_dw)) *g_dir_out = g_dw;
        const float3 g_dv = normalize_vjp(its1.p - origin, g_dw);   // dw = (q - p)/|q - p|
        if (finite3(g_dv)) g_origin = -g_dv;
    }
    if (fwd && jv != 0.f && isfinite(jv)) S.jvp_acc[blockIdx.x * blockDim.x + threadIdx.x] += jv;
    return g_origin;
}

}  // namespace pb
