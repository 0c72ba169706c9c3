// psdr-b200: device/host math shared by every kernel on the path.
//
// fp32 throughout. Where a result feeds a discrete decision that must agree with the CPU oracle bit for bit
// (world-space vertices, the triangle table, the ray/triangle test) the op order is pinned with the *_rn
// intrinsics so that -fmad contraction cannot change it; everything else is plain arithmetic.
// Reference semantics: include/psdr/core/{frame,warp}.h, include/psdr/utils.h, src/core/sampler.cpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define PB_HD __host__ __device__ __forceinline__
#ifndef PB_D   // tests/native/ref_shade_check.cu compiles the per-lane shading functions for the host and predefines this
#define PB_D __device__ __forceinline__
#endif

namespace pb {

constexpr float kEpsilon = 1e-5f, kRayEpsilon = 1e-3f, kShadowEpsilon = 1e-3f, kEdgeEpsilon = 1e-5f;   // constants.h:8-11
constexpr float kPi = 3.14159265358979323846f, kInvPi = 0.31830988618379067154f;
constexpr float kTwoPi = 6.28318530717958647692f, kInvTwoPi = 0.15915494309189533577f;

// ---- pinned-order scalar ops ---------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
PB_D float mul_rn(float a, float b) { return __fmul_rn(a, b); }
PB_D float add_rn(float a, float b) { return __fadd_rn(a, b); }
PB_D float sub_rn(float a, float b) { return __fsub_rn(a, b); }
PB_D float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
PB_D float div_rn(float a, float b) { return __fdiv_rn(a, b); }
PB_D float sqrt_rn(float a) { return __fsqrt_rn(a); }
#else
inline float mul_rn(float a, float b) { return a * b; }
inline float add_rn(float a, float b) { return a + b; }
inline float sub_rn(float a, float b) { return a - b; }
inline float fma_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float div_rn(float a, float b) { return a / b; }
inline float sqrt_rn(float a) { return sqrtf(a); }
#endif

// ---- float3 ------------------------------------------------------------------------------------------------
PB_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
PB_HD float3 f3(float s) { return make_float3(s, s, s); }
PB_HD float3 f3(const float4 &a) { return make_float3(a.x, a.y, a.z); }
PB_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
PB_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
PB_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
PB_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
PB_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
PB_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
PB_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
PB_HD float3 &operator+=(float3 &a, float3 b) { a = a + b; return a; }
PB_HD float3 &operator-=(float3 &a, float3 b) { a = a - b; return a; }
PB_HD float3 &operator*=(float3 &a, float3 b) { a = a * b; return a; }
PB_HD float3 &operator*=(float3 &a, float s) { a = a * s; return a; }
PB_HD float getc(const float3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

// dot / cross as the oracle writes them: dot = fma(ax,bx, fma(ay,by, az*bz)); cross.x = fma(ay,bz, -(az*by))
PB_HD float dot(float3 a, float3 b) { return fma_rn(a.x, b.x, fma_rn(a.y, b.y, mul_rn(a.z, b.z))); }
PB_HD float3 cross(float3 a, float3 b) {
    return f3(fma_rn(a.y, b.z, -mul_rn(a.z, b.y)), fma_rn(a.z, b.x, -mul_rn(a.x, b.z)), fma_rn(a.x, b.y, -mul_rn(a.y, b.x)));
}
PB_HD float squared_norm(float3 a) { return dot(a, a); }
PB_HD float norm(float3 a) { return sqrt_rn(dot(a, a)); }
PB_HD float3 normalize(float3 a) { float n = norm(a); return f3(div_rn(a.x, n), div_rn(a.y, n), div_rn(a.z, n)); }
PB_HD float3 sub3_rn(float3 a, float3 b) { return f3(sub_rn(a.x, b.x), sub_rn(a.y, b.y), sub_rn(a.z, b.z)); }
// utils.h:49-57: fmadd(e1, s, fmadd(e2, t, p0))
PB_HD float3 bilinear(float3 p0, float3 e1, float3 e2, float s, float t) {
    return f3(fma_rn(e1.x, s, fma_rn(e2.x, t, p0.x)), fma_rn(e1.y, s, fma_rn(e2.y, t, p0.y)), fma_rn(e1.z, s, fma_rn(e2.z, t, p0.z)));
}
PB_HD float safe_sqrt(float x) { return sqrt_rn(x > 0.f ? x : 0.f); }
PB_HD float sqr(float x) { return x * x; }
PB_HD float luminance(float3 c) { return c.x * .2126f + c.y * .7152f + c.z * .0722f; }
PB_HD bool finite3(float3 v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }
PB_HD float3 zero_nonfinite(float3 v) { return f3(isfinite(v.x) ? v.x : 0.f, isfinite(v.y) ? v.y : 0.f, isfinite(v.z) ? v.z : 0.f); }
PB_HD float hmax(float3 v) { return fmaxf(fmaxf(v.x, v.y), v.z); }
PB_HD float mulsign(float a, float b) { return signbit(b) ? -a : a; }

// ---- 4x4 row-major matrices (host-prepared, read from constant/param space) ------------------------------
struct Mat4 { float m[16]; };
// transform.h:85-94 in the oracle's op order: ((m0*x + m1*y) + m2*z) + m3, then /w
PB_HD float3 transform_pos(const Mat4 &M, float3 v) {
    float t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        t[i] = add_rn(add_rn(add_rn(mul_rn(M.m[4 * i], v.x), mul_rn(M.m[4 * i + 1], v.y)), mul_rn(M.m[4 * i + 2], v.z)), M.m[4 * i + 3]);
    return f3(div_rn(t[0], t[3]), div_rn(t[1], t[3]), div_rn(t[2], t[3]));
}
PB_HD float3 transform_dir(const Mat4 &M, float3 v) {
    float t[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = add_rn(add_rn(mul_rn(M.m[4 * i], v.x), mul_rn(M.m[4 * i + 1], v.y)), mul_rn(M.m[4 * i + 2], v.z));
    return f3(t[0], t[1], t[2]);
}

// ---- shading frame (frame.h:9-52, Duff et al. 2017) ---------------------------------------------------------
struct Frame {
    float3 s, t, n;
    PB_HD Frame() {}
    PB_HD explicit Frame(float3 v) : n(v) {
        float sg = copysignf(1.f, v.z);
        float a = -1.f / (sg + v.z);
        float b = v.x * v.y * a;
        s = f3(mulsign(sqr(v.x) * a, v.z) + 1.f, mulsign(b, v.z), mulsign(-v.x, v.z));
        t = f3(b, sg + sqr(v.y) * a, -v.y);
    }
    PB_HD float3 to_local(float3 v) const { return f3(dot(v, s), dot(v, t), dot(v, n)); }
    PB_HD float3 to_world(float3 v) const { return s * v.x + t * v.y + n * v.z; }
};

// ---- warps (warp.h:14-80) -------------------------------------------------------------------------------
PB_HD float2 square_to_uniform_disk_concentric(float sx, float sy) {
    float x = fma_rn(2.f, sx, -1.f), y = fma_rn(2.f, sy, -1.f);
    bool is_zero = (x == 0.f && y == 0.f), q13 = fabsf(x) < fabsf(y);
    float r = q13 ? y : x, rp = q13 ? x : y;
    float phi = .25f * kPi * rp / r;
    if (q13) phi = .5f * kPi - phi;
    if (is_zero) phi = 0.f;
    float s, c;
    sincosf(phi, &s, &c);
    return make_float2(r * c, r * s);
}
PB_HD float3 square_to_cosine_hemisphere(float sx, float sy) {
    float2 p = square_to_uniform_disk_concentric(sx, sy);
    float z = safe_sqrt(1.f - fma_rn(p.x, p.x, mul_rn(p.y, p.y)));
    return f3(p.x, p.y, z);
}
PB_HD float2 square_to_uniform_triangle(float sx, float sy) {
    float t = safe_sqrt(1.f - sx);
    return make_float2(1.f - t, t * sy);
}

// ---- ray / triangle (utils.h:67-77), op order pinned to match the oracle's restatement -----------------------
PB_HD void ray_intersect_triangle(float3 p0, float3 e1, float3 e2, float3 o, float3 d, float &u, float &v, float &t) {
    float3 h = cross(d, e2);
    float a = dot(e1, h);
    float f = div_rn(1.f, a);
    float3 s = sub3_rn(o, p0);
    u = mul_rn(f, dot(s, h));
    float3 q = cross(s, e1);
    v = mul_rn(f, dot(d, q));
    t = mul_rn(f, dot(e2, q));
}

// ---- RNG: src/core/sampler.cpp:8-54 + PCG32 (stateless: seeded from the global lane id, then jumped) ----------
constexpr uint64_t kPCG32DefaultState = 0x853c49e6748fea9bULL;
constexpr uint64_t kPCG32Mult = 0x5851f42d4c957f2dULL;

PB_HD uint64_t sample_tea_64(uint64_t v0, uint64_t v1) {   // 4 rounds, 64-bit lanes, 32-bit running sum
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cULL) ^ (v1 + (uint64_t)sum) ^ ((v1 >> 5) + 0xc8013ea4ULL);
        v1 += ((v0 << 4) + 0xad90777dULL) ^ (v0 + (uint64_t)sum) ^ ((v0 >> 5) + 0x7e95761eULL);
    }
    return v0 + (v1 << 32);
}

// LCG jump coefficients for `n` draws: state' = A*state + inc*B (uniform over lanes; computed on the host)
struct RngJump { uint64_t A, B; };
inline RngJump make_jump(uint64_t n) {
    uint64_t cur_mult = kPCG32Mult, cur_plus = 1, acc_mult = 1, acc_plus = 0;
    while (n > 0) {
        if (n & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
        cur_plus = (cur_mult + 1) * cur_plus;
        cur_mult *= cur_mult;
        n >>= 1;
    }
    return {acc_mult, acc_plus};
}

// seeded generator of stream `lane` before any draw: (state, inc) of PCG32 after Sampler::seed (sampler.cpp:29-40). It depends on the lane
// id alone, so a context computes it once per lane (k_rng_seed) and every kernel skips the two TEA hashes (~270 of its instructions).
struct RngSeed { uint64_t state, inc; };
PB_HD RngSeed rng_seed(uint64_t lane) {
    uint64_t seed_value = lane + kPCG32DefaultState;
    uint64_t initstate = sample_tea_64(seed_value, lane), initseq = sample_tea_64(lane, seed_value);
    RngSeed r;
    r.inc = (initseq << 1) | 1u;
    r.state = r.inc;                                  // state=0; step -> inc
    r.state += initstate;
    r.state = r.state * kPCG32Mult + r.inc;
    return r;
}

struct Rng {
    uint64_t state, inc;
    // stream `lane` of a psdr Sampler seeded with arange(count) (sampler.cpp:29-40), advanced by `jump` draws
    PB_HD Rng() : state(0), inc(1) {}   // placeholder of a kernel instantiation that draws nothing
    PB_HD Rng(uint64_t lane, RngJump jump) {
        const RngSeed s = rng_seed(lane);
        inc = s.inc;
        state = jump.A * s.state + inc * jump.B;
    }
    PB_HD Rng(RngSeed s, RngJump jump) {
        inc = s.inc;
        state = jump.A * s.state + inc * jump.B;
    }
    PB_HD uint32_t next_u32() {
        uint64_t old = state;
        state = old * kPCG32Mult + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    PB_HD float next_1d() {
        uint32_t u = (next_u32() >> 9) | 0x3f800000u;
#ifdef __CUDA_ARCH__
        return __uint_as_float(u) - 1.f;
#else
        float f; memcpy(&f, &u, 4); return f - 1.f;
#endif
    }
    // RNG-dimension order: GCC right-to-left argument evaluation: y first, then x
    PB_HD float2 next_2d() { float y = next_1d(); float x = next_1d(); return make_float2(x, y); }
    PB_HD float3 next_3d() { float z = next_1d(); float y = next_1d(); float x = next_1d(); return f3(x, y, z); }
};

}  // namespace pb
