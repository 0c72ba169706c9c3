// ORACLE — TEST INFRASTRUCTURE ONLY. Parity pinned against the reference's own source: psdr-cuda's renderer compiled unmodified from
// /root/reference against CPU stand-ins for its external dependencies Enoki and OptiX (oracle/_ref/libref_render.so, tests/test_ref_render.py,
// tests/test_ref_math.py, tests/test_ref_ingest.py); what stays assumed is Enoki's / OptiX's own rounding and tie-breaking.
//
// Scalar fp32 / forward-mode-dual math used by the CPU restatement of psdr-cuda's hot path.
// Nothing under oracle/ may be imported, linked or executed by the product (psdr_cuda_b200/);
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it.
//
// Follows (semantics, not code): include/psdr/core/frame.h:9-52, include/psdr/core/warp.h:14-95,
// include/psdr/utils.h:32-164, include/psdr/core/transform.h:85-94 and SURVEY.md Appendix D (Enoki semantics).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace orc {

constexpr float kEpsilon = 1e-5f, kRayEpsilon = 1e-3f, kShadowEpsilon = 1e-3f, kEdgeEpsilon = 1e-5f;  // constants.h:8-11
constexpr float kPi = 3.14159265358979323846f, kTwoPi = 6.28318530717958647692f;
constexpr float kInvPi = 0.31830988618379067154f, kInvTwoPi = 0.15915494309189533577f;
constexpr float kInf = std::numeric_limits<float>::infinity();

// ---- forward-mode dual number: value + one tangent --------------------------------------------
struct Dual {
    float v, d;
    Dual() : v(0.f), d(0.f) {}
    Dual(float v_) : v(v_), d(0.f) {}
    Dual(float v_, float d_) : v(v_), d(d_) {}
};
inline float val(float x) { return x; }
inline float val(const Dual &x) { return x.v; }
inline float tan_(float) { return 0.f; }
inline float tan_(const Dual &x) { return x.d; }

inline Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
inline Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
inline Dual operator-(Dual a) { return {-a.v, -a.d}; }
inline Dual operator*(Dual a, Dual b) { return {a.v * b.v, a.d * b.v + a.v * b.d}; }
inline Dual operator/(Dual a, Dual b) {
    float q = a.v / b.v;
    return {q, (a.d - q * b.d) / b.v};
}
inline Dual operator+(Dual a, float b) { return {a.v + b, a.d}; }
inline Dual operator+(float a, Dual b) { return {a + b.v, b.d}; }
inline Dual operator-(Dual a, float b) { return {a.v - b, a.d}; }
inline Dual operator-(float a, Dual b) { return {a - b.v, -b.d}; }
inline Dual operator*(Dual a, float b) { return {a.v * b, a.d * b}; }
inline Dual operator*(float a, Dual b) { return {a * b.v, a * b.d}; }
inline Dual operator/(Dual a, float b) { return {a.v / b, a.d / b}; }
inline Dual operator/(float a, Dual b) { return Dual(a) / b; }
inline Dual &operator+=(Dual &a, Dual b) { a = a + b; return a; }
inline Dual &operator-=(Dual &a, Dual b) { a = a - b; return a; }
inline Dual &operator*=(Dual &a, Dual b) { a = a * b; return a; }
inline Dual &operator/=(Dual &a, Dual b) { a = a / b; return a; }

// fused multiply-add: value part uses fmaf so float and Dual primal paths agree bit for bit
inline float fma_(float a, float b, float c) { return std::fmaf(a, b, c); }
inline Dual fma_(Dual a, Dual b, Dual c) { return {std::fmaf(a.v, b.v, c.v), a.d * b.v + a.v * b.d + c.d}; }
inline Dual fma_(Dual a, float b, Dual c) { return fma_(a, Dual(b), c); }
inline Dual fma_(float a, Dual b, Dual c) { return fma_(Dual(a), b, c); }
inline Dual fma_(Dual a, Dual b, float c) { return fma_(a, b, Dual(c)); }

inline float sqrt_(float x) { return std::sqrt(x); }
// A zero tangent stays zero through an infinite local derivative (sqrt at 0, acos at +-1): Enoki's autodiff multiplies edge weights with a
// "safe" product in which 0 * inf = 0, and the stand-in the reference's own code is run under (oracle/ref_dyn) does the same.
inline Dual sqrt_(Dual x) { float s = std::sqrt(x.v); return {s, x.d == 0.f ? 0.f : x.d / (2.f * s)}; }
inline float abs_(float x) { return std::fabs(x); }
inline Dual abs_(Dual x) { return x.v < 0.f ? Dual(-x.v, -x.d) : x; }
inline float sin_(float x) { return std::sin(x); }
inline Dual sin_(Dual x) { return {std::sin(x.v), std::cos(x.v) * x.d}; }
inline float cos_(float x) { return std::cos(x); }
inline Dual cos_(Dual x) { return {std::cos(x.v), -std::sin(x.v) * x.d}; }
inline float acos_(float x) { return std::acos(x); }
inline Dual acos_(Dual x) { return {std::acos(x.v), x.d == 0.f ? 0.f : -x.d / std::sqrt(1.f - x.v * x.v)}; }
inline float atan2_(float y, float x) { return std::atan2(y, x); }
inline Dual atan2_(Dual y, Dual x) {
    float r2 = x.v * x.v + y.v * y.v;
    return {std::atan2(y.v, x.v), (x.v * y.d - y.v * x.d) / r2};
}
inline float floor_(float x) { return std::floor(x); }
inline Dual floor_(Dual x) { return {std::floor(x.v), 0.f}; }
inline float max_(float a, float b) { return a > b ? a : b; }   // enoki::max(a,b): a if a>b else b
inline Dual max_(Dual a, Dual b) { return a.v > b.v ? a : b; }
inline float min_(float a, float b) { return a < b ? a : b; }
inline Dual min_(Dual a, Dual b) { return a.v < b.v ? a : b; }
template <class R> inline R sqr(R x) { return x * x; }
template <class R> inline R safe_sqrt(R x) { return sqrt_(max_(x, R(0.f))); }
template <class R> inline R safe_acos(R x) { return acos_(min_(max_(x, R(-1.f)), R(1.f))); }
template <class R> inline R clamp_(R x, float lo, float hi) { return min_(max_(x, R(lo)), R(hi)); }
template <class R> inline bool finite_(R x) { return std::isfinite(val(x)) && std::isfinite(tan_(x)); }

inline float sign1(float x) { return std::copysign(1.f, x); }   // enoki::sign
template <class R> inline R mulsign(R a, float b) { return std::signbit(b) ? R(-a) : a; }

// ---- small vectors ---------------------------------------------------------------------------
template <class R> struct V2 {
    R x, y;
    V2() : x(0.f), y(0.f) {}
    V2(R x_, R y_) : x(x_), y(y_) {}
    template <class S> V2(const V2<S> &o) : x(o.x), y(o.y) {}
};
template <class R> struct V3 {
    R x, y, z;
    V3() : x(0.f), y(0.f), z(0.f) {}
    explicit V3(R s) : x(s), y(s), z(s) {}
    V3(R x_, R y_, R z_) : x(x_), y(y_), z(z_) {}
    template <class S> V3(const V3<S> &o) : x(o.x), y(o.y), z(o.z) {}
    R &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const R &operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
using V3f = V3<float>;
using V2f = V2<float>;

template <class R> inline V3<R> operator+(V3<R> a, V3<R> b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <class R> inline V3<R> operator-(V3<R> a, V3<R> b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <class R> inline V3<R> operator-(V3<R> a) { return {-a.x, -a.y, -a.z}; }
template <class R> inline V3<R> operator*(V3<R> a, V3<R> b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <class R> inline V3<R> operator*(V3<R> a, R s) { return {a.x * s, a.y * s, a.z * s}; }
template <class R> inline V3<R> operator*(R s, V3<R> a) { return {a.x * s, a.y * s, a.z * s}; }
template <class R> inline V3<R> operator/(V3<R> a, R s) { return {a.x / s, a.y / s, a.z / s}; }
template <class R> inline V3<R> operator/(V3<R> a, V3<R> b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
template <class R> inline V3<R> &operator+=(V3<R> &a, V3<R> b) { a = a + b; return a; }
template <class R> inline V3<R> &operator*=(V3<R> &a, V3<R> b) { a = a * b; return a; }
template <class R> inline V3<R> &operator*=(V3<R> &a, R s) { a = a * s; return a; }
template <class R> inline V3<R> &operator/=(V3<R> &a, R s) { a = a / s; return a; }
template <class R> inline V2<R> operator+(V2<R> a, V2<R> b) { return {a.x + b.x, a.y + b.y}; }
template <class R> inline V2<R> operator-(V2<R> a, V2<R> b) { return {a.x - b.x, a.y - b.y}; }
template <class R> inline V2<R> operator*(V2<R> a, R s) { return {a.x * s, a.y * s}; }

inline float detach(float x) { return x; }
inline float detach(const Dual &x) { return x.v; }
template <class R> inline V3f detach(const V3<R> &a) { return {val(a.x), val(a.y), val(a.z)}; }
template <class R> inline V2f detach(const V2<R> &a) { return {val(a.x), val(a.y)}; }

// dot / cross use an fma chain so that the product's kernels can reproduce the primal bit for bit
// where that matters (the ray/triangle test): dot = fma(ax,bx, fma(ay,by, az*bz)); cross.x = fma(ay,bz, -(az*by)).
template <class R> inline R dot(const V3<R> &a, const V3<R> &b) { return fma_(a.x, b.x, fma_(a.y, b.y, a.z * b.z)); }
template <class R> inline R dot(const V2<R> &a, const V2<R> &b) { return fma_(a.x, b.x, a.y * b.y); }
template <class R> inline V3<R> cross(const V3<R> &a, const V3<R> &b) {
    return {fma_(a.y, b.z, -(a.z * b.y)), fma_(a.z, b.x, -(a.x * b.z)), fma_(a.x, b.y, -(a.y * b.x))};
}
template <class R> inline R squared_norm(const V3<R> &a) { return dot(a, a); }
template <class R> inline R norm(const V3<R> &a) { return sqrt_(squared_norm(a)); }
template <class R> inline R norm(const V2<R> &a) { return sqrt_(dot(a, a)); }
template <class R> inline V3<R> normalize(const V3<R> &a) { return a / norm(a); }
template <class R> inline bool finite_(const V3<R> &a) { return finite_(a.x) && finite_(a.y) && finite_(a.z); }
// utils.h:49-57: fmadd(e1, s, fmadd(e2, t, p0))
template <class R> inline V3<R> bilinear(const V3<R> &p0, const V3<R> &e1, const V3<R> &e2, const V2<R> &st) {
    return {fma_(e1.x, st.x, fma_(e2.x, st.y, p0.x)), fma_(e1.y, st.x, fma_(e2.y, st.y, p0.y)),
            fma_(e1.z, st.x, fma_(e2.z, st.y, p0.z))};
}
template <class R> inline V2<R> bilinear2(const V2<R> &p0, const V2<R> &e1, const V2<R> &e2, const V2<R> &st) {
    return {fma_(e1.x, st.x, fma_(e2.x, st.y, p0.x)), fma_(e1.y, st.x, fma_(e2.y, st.y, p0.y))};
}
inline float rgb2luminance(const V3f &c) { return c.x * .2126f + c.y * .7152f + c.z * .0722f; }  // utils.h:61-63

// ---- 4x4 matrices (row-major) ------------------------------------------------------------------
template <class R> struct M4 {
    R m[4][4];
    M4() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = R(i == j ? 1.f : 0.f); }
    template <class S> M4(const M4<S> &o) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = R(o.m[i][j]); }
};
using M4f = M4<float>;
template <class R> inline M4<R> operator*(const M4<R> &a, const M4<R> &b) {
    M4<R> c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            R s(0.f);
            for (int k = 0; k < 4; ++k) s = s + a.m[i][k] * b.m[k][j];
            c.m[i][j] = s;
        }
    return c;
}
template <class R> inline M4f detach(const M4<R> &a) {
    M4f c;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c.m[i][j] = val(a.m[i][j]);
    return c;
}
// transform.h:85-94
template <class R> inline V3<R> transform_pos(const M4<R> &M, const V3<R> &v) {
    R t[4];
    for (int i = 0; i < 4; ++i) t[i] = M.m[i][0] * v.x + M.m[i][1] * v.y + M.m[i][2] * v.z + M.m[i][3];
    return {t[0] / t[3], t[1] / t[3], t[2] / t[3]};
}
template <class R> inline V3<R> transform_dir(const M4<R> &M, const V3<R> &v) {
    R t[3];
    for (int i = 0; i < 3; ++i) t[i] = M.m[i][0] * v.x + M.m[i][1] * v.y + M.m[i][2] * v.z;
    return {t[0], t[1], t[2]};
}
// general 4x4 inverse (double precision Gauss-Jordan on the primal; tangent via -A^-1 dA A^-1)
inline M4f inverse(const M4f &A) {
    double a[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = A.m[i][j]; a[i][4 + j] = (i == j); }
    for (int c = 0; c < 4; ++c) {
        int p = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[p][c])) p = r;
        if (p != c) for (int j = 0; j < 8; ++j) std::swap(a[p][j], a[c][j]);
        double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= inv;
        for (int r = 0; r < 4; ++r) if (r != c) { double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
    }
    M4f R;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.m[i][j] = (float)a[i][4 + j];
    return R;
}
inline M4<Dual> inverse(const M4<Dual> &A) {
    M4f Ai = inverse(detach(A)), dA;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) dA.m[i][j] = A.m[i][j].d;
    M4f T = Ai * dA * Ai;
    M4<Dual> R;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) R.m[i][j] = Dual(Ai.m[i][j], -T.m[i][j]);
    return R;
}

// ---- frame (frame.h:9-52; Duff et al. 2017 orthonormal basis) -----------------------------------
template <class R> struct Frame {
    V3<R> s, t, n;
    Frame() {}
    explicit Frame(const V3<R> &v) : n(v) {
        float nz = val(v.z);
        float sg = sign1(nz);
        R a = R(-1.f) / (sg + v.z);
        R b = v.x * v.y * a;
        s = V3<R>(mulsign(sqr(v.x) * a, nz) + 1.f, mulsign(b, nz), mulsign(R(-v.x), nz));
        t = V3<R>(b, sg + sqr(v.y) * a, -v.y);
    }
    V3<R> to_local(const V3<R> &v) const { return {dot(v, s), dot(v, t), dot(v, n)}; }
    V3<R> to_world(const V3<R> &v) const { return s * v.x + t * v.y + n * v.z; }
};

// ---- warps (warp.h:14-80) --------------------------------------------------------------------
template <class R> inline V2<R> square_to_uniform_disk_concentric(const V2<R> &sample) {
    R x = fma_(R(2.f), sample.x, R(-1.f)), y = fma_(R(2.f), sample.y, R(-1.f));
    bool is_zero = (val(x) == 0.f && val(y) == 0.f), q13 = std::fabs(val(x)) < std::fabs(val(y));
    R r = q13 ? y : x, rp = q13 ? x : y;
    R phi = .25f * kPi * rp / r;
    if (q13) phi = .5f * kPi - phi;
    if (is_zero) phi = R(0.f);
    return {r * cos_(phi), r * sin_(phi)};
}
template <class R> inline V3<R> square_to_cosine_hemisphere(const V2<R> &sample) {
    V2<R> p = square_to_uniform_disk_concentric(sample);
    R z = safe_sqrt(R(1.f) - dot(p, p));
    return {p.x, p.y, z};
}
template <class R> inline V2<R> square_to_uniform_triangle(const V2<R> &sample) {
    R t = safe_sqrt(R(1.f) - sample.x);
    return {R(1.f) - t, t * sample.y};
}

// ---- ray / triangle (utils.h:67-77) ------------------------------------------------------------
template <class R> struct Ray {
    V3<R> o, d;
    float tmax = kInf;
    Ray() {}
    Ray(const V3<R> &o_, const V3<R> &d_) : o(o_), d(d_) {}
    V3<R> operator()(R t) const { return {fma_(d.x, t, o.x), fma_(d.y, t, o.y), fma_(d.z, t, o.z)}; }
};
template <class R>
inline void ray_intersect_triangle(const V3<R> &p0, const V3<R> &e1, const V3<R> &e2, const Ray<R> &ray, R &u, R &v, R &t) {
    V3<R> h = cross(ray.d, e2);
    R a = dot(e1, h);
    R f = R(1.f) / a;
    V3<R> s = ray.o - p0;
    u = f * dot(s, h);
    V3<R> q = cross(s, e1);
    v = f * dot(ray.d, q);
    t = f * dot(e2, q);
}

}  // namespace orc
