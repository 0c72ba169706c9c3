// psdr-b200: forward-mode dual numbers with N tangents, used *locally* inside the reverse-mode kernels.
//
// The reference differentiates its BSDFs with Enoki's tape (every Float<ad> in src/bsdf/{roughconductor,ggx}.cpp and
// include/psdr/utils.h:149-164 is attached). A hand-written adjoint of the rough-conductor event (visible-normal sampling,
// attached sampling pdf, MIS weight) would be long and fragile; instead the event's loss-weighted scalar is evaluated once
// on Dual<N> whose N tangents are the event's few local inputs (alpha_u, alpha_v; or the local directions wi, wo). That
// yields the local gradient in one pass, which the kernel then scatters like any other adjoint (reverse mode globally,
// forward mode locally).
#pragma once
#include "pb_math.cuh"

namespace pb {

template <int N>
struct Dual {
    float v;
    float d[N];
    PB_HD Dual() {}
    PB_HD Dual(float x) : v(x) {
#pragma unroll
        for (int i = 0; i < N; ++i) d[i] = 0.f;
    }
    static PB_HD Dual seed(float x, int k) { Dual r(x); r.d[k] = 1.f; return r; }
};

template <int N> PB_HD float val(const Dual<N> &a) { return a.v; }
PB_HD float val(float a) { return a; }

#define PB_DUAL_LOOP _Pragma("unroll") for (int i = 0; i < N; ++i)

template <int N> PB_HD Dual<N> operator-(const Dual<N> &a) { Dual<N> r; r.v = -a.v; PB_DUAL_LOOP r.d[i] = -a.d[i]; return r; }
template <int N> PB_HD Dual<N> operator+(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v + b.v; PB_DUAL_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int N> PB_HD Dual<N> operator-(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v - b.v; PB_DUAL_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int N> PB_HD Dual<N> operator*(const Dual<N> &a, const Dual<N> &b) { Dual<N> r; r.v = a.v * b.v; PB_DUAL_LOOP r.d[i] = fmaf(a.d[i], b.v, a.v * b.d[i]); return r; }
template <int N> PB_HD Dual<N> operator/(const Dual<N> &a, const Dual<N> &b) {
    Dual<N> r;
    const float inv = 1.f / b.v;
    r.v = a.v * inv;
    PB_DUAL_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int N> PB_HD Dual<N> operator+(const Dual<N> &a, float b) { Dual<N> r = a; r.v += b; return r; }
template <int N> PB_HD Dual<N> operator+(float a, const Dual<N> &b) { Dual<N> r = b; r.v += a; return r; }
template <int N> PB_HD Dual<N> operator-(const Dual<N> &a, float b) { Dual<N> r = a; r.v -= b; return r; }
template <int N> PB_HD Dual<N> operator-(float a, const Dual<N> &b) { Dual<N> r; r.v = a - b.v; PB_DUAL_LOOP r.d[i] = -b.d[i]; return r; }
template <int N> PB_HD Dual<N> operator*(const Dual<N> &a, float b) { Dual<N> r; r.v = a.v * b; PB_DUAL_LOOP r.d[i] = a.d[i] * b; return r; }
template <int N> PB_HD Dual<N> operator*(float a, const Dual<N> &b) { return b * a; }
template <int N> PB_HD Dual<N> operator/(const Dual<N> &a, float b) { return a * (1.f / b); }
template <int N> PB_HD Dual<N> operator/(float a, const Dual<N> &b) {
    Dual<N> r;
    const float inv = 1.f / b.v;
    r.v = a * inv;
    const float s = -r.v * inv;
    PB_DUAL_LOOP r.d[i] = s * b.d[i];
    return r;
}
template <int N> PB_HD Dual<N> &operator+=(Dual<N> &a, const Dual<N> &b) { a = a + b; return a; }

template <int N> PB_HD Dual<N> dsqr(const Dual<N> &a) { return a * a; }
PB_HD float dsqr(float a) { return a * a; }
template <int N> PB_HD Dual<N> dsqrt(const Dual<N> &a) {
    Dual<N> r;
    r.v = sqrtf(a.v);
    const float s = .5f / r.v;
    PB_DUAL_LOOP r.d[i] = a.d[i] * s;
    return r;
}
PB_HD float dsqrt(float a) { return sqrtf(a); }
// Enoki's safe_sqrt = sqrt(max(x, 0)): the clamp kills the tangent when it is active
template <int N> PB_HD Dual<N> dsafe_sqrt(const Dual<N> &a) { return a.v > 0.f ? dsqrt(a) : Dual<N>(0.f); }
PB_HD float dsafe_sqrt(float a) { return safe_sqrt(a); }
template <int N> PB_HD Dual<N> dabs(const Dual<N> &a) { return a.v < 0.f ? -a : a; }
PB_HD float dabs(float a) { return fabsf(a); }
template <int N> PB_HD Dual<N> dclamp(const Dual<N> &a, float lo, float hi) { return a.v < lo ? Dual<N>(lo) : (a.v > hi ? Dual<N>(hi) : a); }
PB_HD float dclamp(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
template <int N> PB_HD bool dfinite(const Dual<N> &a) {
    bool ok = isfinite(a.v);
    PB_DUAL_LOOP ok = ok && isfinite(a.d[i]);
    return ok;
}

// ---- 3-vectors over T (float or Dual<N>) ------------------------------------------------------------------------
template <class T> struct V3 {
    T x, y, z;
    PB_HD V3() {}
    PB_HD V3(T a, T b, T c) : x(a), y(b), z(c) {}
    PB_HD explicit V3(float3 a) : x(a.x), y(a.y), z(a.z) {}
};
template <class T> PB_HD V3<T> operator+(const V3<T> &a, const V3<T> &b) { return V3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> PB_HD V3<T> operator-(const V3<T> &a, const V3<T> &b) { return V3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> PB_HD V3<T> operator-(const V3<T> &a) { return V3<T>(-a.x, -a.y, -a.z); }
template <class T> PB_HD V3<T> operator*(const V3<T> &a, const T &s) { return V3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> PB_HD V3<T> operator/(const V3<T> &a, const T &s) { return V3<T>(a.x / s, a.y / s, a.z / s); }
template <class T> PB_HD T vdot(const V3<T> &a, const V3<T> &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> PB_HD V3<T> vnormalize(const V3<T> &a) { return a / dsqrt(vdot(a, a)); }
template <class T> PB_HD float3 vval(const V3<T> &a) { return f3(val(a.x), val(a.y), val(a.z)); }

#undef PB_DUAL_LOOP

}  // namespace pb
