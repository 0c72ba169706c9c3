// TEST INFRASTRUCTURE (oracle/): a CPU stand-in for the subset of the Enoki API that psdr-cuda's renderer uses, so that the reference's
// OWN source files (src/**/*.cpp, include/psdr/**) compile unmodified, from where they lie under /root/reference, and RUN on the host:
// oracle/build_ref.sh -> oracle/_ref/libref_render.so, used by tests/test_ref_render.py to pin the oracle to the reference's code.
// Enoki itself (and OptiX) are un-vendored external dependencies of the reference; nothing here is copied from them.
//
// Model: CUDAArray<T> is a host vector of lanes (size 1 broadcasts); DiffArray<A> is a value array plus an optional FORWARD-MODE tangent
// array (Enoki's reverse-mode tape is replaced by one directional derivative: seed a parameter's tangent, run renderD, read the
// tangent of the image = what ek.forward(param) leaves in ek.gradient(image)); detach() drops the tangent. Array<T, n> / Matrix<T, n> are
// fixed-size containers of those; ENOKI_STRUCT types are traversed field by field; arrays of object pointers dispatch method calls per
// distinct pointer under a mask (ENOKI_CALL_SUPPORT_*).
// Assumed semantics (not checkable without Enoki): IEEE fp32 with exact 1/x and 1/sqrt(x) for rcp / rsqrt, libm
// transcendentals, sequential fp32 hsum / psum, dot(a, b) = fmadd chain from the last component, masked-out gather lanes read zero.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// ---- struct support: field lists -------------------------------------------------------------------------------------------------------
#define ENOKI_INLINE inline
#define ENOKI_EXPAND(x) x
#define ENOKI_FE_1(m, a) m(a)
#define ENOKI_FE_2(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_1(m, __VA_ARGS__))
#define ENOKI_FE_3(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_2(m, __VA_ARGS__))
#define ENOKI_FE_4(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_3(m, __VA_ARGS__))
#define ENOKI_FE_5(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_4(m, __VA_ARGS__))
#define ENOKI_FE_6(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_5(m, __VA_ARGS__))
#define ENOKI_FE_7(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_6(m, __VA_ARGS__))
#define ENOKI_FE_8(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_7(m, __VA_ARGS__))
#define ENOKI_FE_9(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_8(m, __VA_ARGS__))
#define ENOKI_FE_10(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_9(m, __VA_ARGS__))
#define ENOKI_FE_11(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_10(m, __VA_ARGS__))
#define ENOKI_FE_12(m, a, ...) m(a) ENOKI_EXPAND(ENOKI_FE_11(m, __VA_ARGS__))
#define ENOKI_FE_N(_1, _2, _3, _4, _5, _6, _7, _8, _9, _10, _11, _12, N, ...) N
#define ENOKI_FOR_EACH(m, ...) ENOKI_EXPAND(ENOKI_FE_N(__VA_ARGS__, ENOKI_FE_12, ENOKI_FE_11, ENOKI_FE_10, ENOKI_FE_9, ENOKI_FE_8, ENOKI_FE_7, ENOKI_FE_6, ENOKI_FE_5, ENOKI_FE_4, ENOKI_FE_3, ENOKI_FE_2, ENOKI_FE_1)(m, __VA_ARGS__))
#define ENOKI_FIELD_CALL(f) fn_(os_.f...);
#define ENOKI_FIELDS_FN(...) template <class Fn_, class... Os_> static void enoki_fields(Fn_ &&fn_, Os_ &...os_) { ENOKI_FOR_EACH(ENOKI_FIELD_CALL, __VA_ARGS__) }
#define ENOKI_STRUCT(Name, ...) Name() = default; ENOKI_EXPAND(ENOKI_FIELDS_FN(__VA_ARGS__))
#define ENOKI_BASE_FIELDS(...) __VA_ARGS__
#define ENOKI_DERIVED_FIELDS(...) __VA_ARGS__
#define ENOKI_DERIVED_STRUCT(Name, Base, ...) Name() = default; ENOKI_EXPAND(ENOKI_FIELDS_FN(__VA_ARGS__))
#define ENOKI_STRUCT_SUPPORT(...)
#define ENOKI_PINNED_OPERATOR_NEW(...)
#define ENOKI_PIN_U1(B, a) using B::a;
#define ENOKI_PIN_U2(B, a, ...) using B::a; ENOKI_PIN_U1(B, __VA_ARGS__)
#define ENOKI_PIN_U3(B, a, ...) using B::a; ENOKI_PIN_U2(B, __VA_ARGS__)
#define ENOKI_PIN_U4(B, a, ...) using B::a; ENOKI_PIN_U3(B, __VA_ARGS__)
#define ENOKI_PIN_U5(B, a, ...) using B::a; ENOKI_PIN_U4(B, __VA_ARGS__)
#define ENOKI_PIN_U6(B, a, ...) using B::a; ENOKI_PIN_U5(B, __VA_ARGS__)
#define ENOKI_PIN_U7(B, a, ...) using B::a; ENOKI_PIN_U6(B, __VA_ARGS__)
#define ENOKI_PIN_U8(B, a, ...) using B::a; ENOKI_PIN_U7(B, __VA_ARGS__)
#define ENOKI_PIN_N(_0, _1, _2, _3, _4, _5, _6, _7, _8, N, ...) N
#define ENOKI_PIN_DISPATCH(B, ...) ENOKI_EXPAND(ENOKI_PIN_N(B, __VA_ARGS__, ENOKI_PIN_U8, ENOKI_PIN_U7, ENOKI_PIN_U6, ENOKI_PIN_U5, ENOKI_PIN_U4, ENOKI_PIN_U3, ENOKI_PIN_U2, ENOKI_PIN_U1)(B, __VA_ARGS__))
#define ENOKI_USING_MEMBERS(...) ENOKI_EXPAND(ENOKI_PIN_DISPATCH(__VA_ARGS__))

namespace enoki {

constexpr float Pi = 3.14159265358979323846f, InvPi = 0.31830988618379067154f, TwoPi = 6.28318530717958647692f, InvTwoPi = 0.15915494309189533577f;
template <class T> constexpr T Epsilon = T(1.1920929e-07) / 2;
template <class T> constexpr T Infinity = std::numeric_limits<T>::infinity();

template <class Class, class Storage> struct call_support;
template <class T> using store_t = std::conditional_t<std::is_same_v<T, bool>, uint8_t, T>;

// ---- lanes -----------------------------------------------------------------------------------------------------------------------------
template <class T> struct CUDAArray {
    using Scalar = T;
    static constexpr bool IsDiff = false;
    std::vector<store_t<T>> d;
    CUDAArray() = default;
    CUDAArray(T x) : d(1, (store_t<T>)x) {}
    template <class U, std::enable_if_t<std::is_arithmetic_v<U> && std::is_arithmetic_v<T> && !std::is_same_v<U, T>, int> = 0> CUDAArray(U x) : d(1, (store_t<T>)T(x)) {}
    template <class U, std::enable_if_t<!std::is_same_v<U, T> && std::is_convertible_v<U, T>, int> = 0> CUDAArray(const CUDAArray<U> &o) : d(o.d.size()) { for (size_t i = 0; i < d.size(); ++i) d[i] = (store_t<T>)T(U(o.d[i])); }
    static CUDAArray copy(const T *p, size_t n) { CUDAArray r; r.d.resize(n); for (size_t i = 0; i < n; ++i) r.d[i] = (store_t<T>)p[i]; return r; }
    const store_t<T> *data() const { return d.data(); }
    store_t<T> *data() { return d.data(); }
    size_t size() const { return d.size(); }
    bool empty() const { return d.empty(); }
    T operator[](size_t i) const { return T(d[d.size() == 1 ? 0 : i]); }
    auto operator->() const { return call_support<std::remove_const_t<std::remove_pointer_t<T>>, CUDAArray>{*this}; }
};
template <class A> struct DiffArray {
    using Scalar = typename A::Scalar;
    static constexpr bool IsDiff = true;
    A v;
    CUDAArray<Scalar> g;   // forward-mode tangent (fp32 arrays only); empty = zero
    DiffArray() = default;
    DiffArray(Scalar x) : v(x) {}
    template <class U, std::enable_if_t<std::is_arithmetic_v<U> && std::is_arithmetic_v<Scalar> && !std::is_same_v<U, Scalar>, int> = 0> DiffArray(U x) : v(Scalar(x)) {}
    template <class U, std::enable_if_t<std::is_convertible_v<U, Scalar>, int> = 0> DiffArray(const CUDAArray<U> &a) : v(a) {}
    template <class B, std::enable_if_t<!std::is_same_v<B, A>, int> = 0> DiffArray(const DiffArray<B> &o) : v(o.v) {}
    static DiffArray copy(const Scalar *p, size_t n) { DiffArray r; r.v = A::copy(p, n); return r; }
    const store_t<Scalar> *data() const { return v.data(); }
    size_t size() const { return v.size(); }
    bool empty() const { return v.empty(); }
    Scalar operator[](size_t i) const { return v[i]; }
    auto operator->() const { return call_support<std::remove_const_t<std::remove_pointer_t<Scalar>>, DiffArray>{*this}; }
};
template <class S> struct is_sc : std::false_type {};
template <class T> struct is_sc<CUDAArray<T>> : std::true_type {};
template <class A> struct is_sc<DiffArray<A>> : std::true_type {};
template <class S> constexpr bool is_sc_v = is_sc<std::decay_t<S>>::value;
template <class S> constexpr bool is_plain_v = std::is_arithmetic_v<std::decay_t<S>> || std::is_pointer_v<std::decay_t<S>> || std::is_null_pointer_v<std::decay_t<S>>;
template <class S, class = void> struct scalar_of { using type = std::decay_t<S>; };
template <class S> struct scalar_of<S, std::enable_if_t<is_sc_v<S>>> { using type = typename std::decay_t<S>::Scalar; };
template <class S> using scalar_t = typename scalar_of<S>::type;
template <class S> constexpr bool is_diff_v = [] { if constexpr (is_sc_v<S>) return std::decay_t<S>::IsDiff; else return false; }();
template <class S> constexpr bool is_fdiff_v = is_diff_v<S> && std::is_same_v<scalar_t<S>, float>;
template <class T, bool diff> using leaf_t = std::conditional_t<diff, DiffArray<CUDAArray<T>>, CUDAArray<T>>;

template <class X> size_t lsize(const X &x) { if constexpr (is_sc_v<X>) return x.size(); else return 1; }
template <class X> auto lval(const X &x, size_t i) {   // an empty (never assigned) array reads as zero
    if constexpr (is_sc_v<X>) return x.size() ? x[i] : scalar_t<X>(0); else return x;
}
template <class X> float ltan(const X &x, size_t i) {
    if constexpr (is_fdiff_v<X>) return x.g.d.empty() ? 0.f : x.g.d[x.g.d.size() == 1 ? 0 : i]; else return 0.f;
}
template <class X> bool lhas(const X &x) { if constexpr (is_fdiff_v<X>) return !x.g.d.empty(); else return false; }
inline size_t bsize2(size_t a, size_t b) {
    if (a == b || b <= 1) return std::max<size_t>(a, 1);
    if (a <= 1) return b;
    throw std::runtime_error("enoki stand-in: incompatible array sizes " + std::to_string(a) + " vs " + std::to_string(b));
}
template <class... Xs> size_t bsize(const Xs &...xs) { size_t n = 1; ((n = bsize2(n, lsize(xs))), ...); return n; }
template <class R> R lmake(size_t n) { R r; if constexpr (R::IsDiff) r.v.d.resize(n); else r.d.resize(n); return r; }
template <class R> auto &lstore(R &r) { if constexpr (std::decay_t<R>::IsDiff) return r.v.d; else return r.d; }
template <class R> const auto &lstore(const R &r) { if constexpr (std::decay_t<R>::IsDiff) return r.v.d; else return r.d; }

// result type of a binary op: Diff wins, scalar type by the usual promotion
template <class A, class B, class = void> struct bin {};
template <class A, class B> struct bin<A, B, std::enable_if_t<is_sc_v<A> && is_sc_v<B>>> { using type = leaf_t<decltype(scalar_t<A>() + scalar_t<B>()), is_diff_v<A> || is_diff_v<B>>; };
// array (op) plain scalar: the array's type, except that an integer array meeting a floating-point scalar becomes an fp32 array
template <class S, class P> using promo_t = std::conditional_t<std::is_integral_v<scalar_t<S>> && !std::is_same_v<scalar_t<S>, bool> && std::is_floating_point_v<std::decay_t<P>>, leaf_t<float, is_diff_v<S>>, std::decay_t<S>>;
template <class A, class B> struct bin<A, B, std::enable_if_t<is_sc_v<A> && std::is_arithmetic_v<std::decay_t<B>>>> { using type = promo_t<A, B>; };
template <class A, class B> struct bin<A, B, std::enable_if_t<std::is_arithmetic_v<std::decay_t<A>> && is_sc_v<B>>> { using type = promo_t<B, A>; };
template <class A, class B> using bin_t = typename bin<A, B>::type;
template <class A, class B, class C> using bin3_t = bin_t<bin_t<A, B>, C>;
template <class... Xs> using mask_of_t = leaf_t<bool, (is_diff_v<Xs> || ...)>;

#define ENOKI_DYN_BINOP(op, TEXPR) \
    template <class A, class B, class R = bin_t<A, B>> R operator op(const A &a, const B &b) { \
        using S = scalar_t<R>; const size_t n = bsize(a, b); R r = lmake<R>(n); auto &rs = lstore(r); \
        for (size_t i = 0; i < n; ++i) rs[i] = S(S(lval(a, i)) op S(lval(b, i))); \
        if constexpr (is_fdiff_v<R>) if (lhas(a) || lhas(b)) { r.g.d.resize(n); \
            for (size_t i = 0; i < n; ++i) { const float x = (float)lval(a, i), y = (float)lval(b, i), dx = ltan(a, i), dy = ltan(b, i), z = rs[i]; (void)x; (void)y; (void)z; r.g.d[i] = TEXPR; } } \
        return r; }
ENOKI_DYN_BINOP(+, dx + dy) ENOKI_DYN_BINOP(-, dx - dy) ENOKI_DYN_BINOP(*, dx * y + x * dy) ENOKI_DYN_BINOP(/, (dx - z * dy) / y)
#undef ENOKI_DYN_BINOP
template <class A, class B, class R = bin_t<A, B>, std::enable_if_t<std::is_integral_v<scalar_t<R>>, int> = 0> R operator%(const A &a, const B &b) {
    const size_t n = bsize(a, b); R r = lmake<R>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) rs[i] = lval(a, i) % lval(b, i);
    return r;
}
#define ENOKI_DYN_CMP(name, op) \
    template <class A, class B, std::enable_if_t<(is_sc_v<A> || is_sc_v<B>) && (is_sc_v<A> || is_plain_v<A>) && (is_sc_v<B> || is_plain_v<B>), int> = 0> mask_of_t<A, B> name(const A &a, const B &b) { \
        using R = mask_of_t<A, B>; const size_t n = bsize(a, b); R r = lmake<R>(n); auto &rs = lstore(r); \
        for (size_t i = 0; i < n; ++i) rs[i] = lval(a, i) op lval(b, i); \
        return r; }
ENOKI_DYN_CMP(operator<, <) ENOKI_DYN_CMP(operator<=, <=) ENOKI_DYN_CMP(operator>, >) ENOKI_DYN_CMP(operator>=, >=) ENOKI_DYN_CMP(eq, ==) ENOKI_DYN_CMP(neq, !=)
#undef ENOKI_DYN_CMP
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S operator-(const S &a) {
    S r = a; auto &rs = lstore(r);
    for (auto &x : rs) x = -x;
    if constexpr (is_fdiff_v<S>) for (auto &x : r.g.d) x = -x;
    return r;
}
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S operator+(const S &a) { return a; }
#define ENOKI_DYN_ASSIGN(op) template <class S, class B, std::enable_if_t<is_sc_v<S>, int> = 0> S &operator op##=(S &a, const B &b) { a = S(a op b); return a; }
ENOKI_DYN_ASSIGN(+) ENOKI_DYN_ASSIGN(-) ENOKI_DYN_ASSIGN(*) ENOKI_DYN_ASSIGN(/)
#undef ENOKI_DYN_ASSIGN

// ---- masks -------------------------------------------------------------------------------------------------------------------------------
template <class S> constexpr bool is_mask_v = is_sc_v<S> && std::is_same_v<scalar_t<S>, bool>;
template <class S> constexpr bool is_masklike_v = is_mask_v<S> || std::is_same_v<std::decay_t<S>, bool>;
#define ENOKI_DYN_MASKOP(name, op) \
    template <class A, class B, std::enable_if_t<(is_mask_v<A> || is_mask_v<B>) && is_masklike_v<A> && is_masklike_v<B>, int> = 0> mask_of_t<A, B> name(const A &a, const B &b) { \
        using R = mask_of_t<A, B>; const size_t n = bsize(a, b); R r = lmake<R>(n); auto &rs = lstore(r); \
        for (size_t i = 0; i < n; ++i) rs[i] = bool(lval(a, i)) op bool(lval(b, i)); \
        return r; }
ENOKI_DYN_MASKOP(operator&&, &&) ENOKI_DYN_MASKOP(operator||, ||) ENOKI_DYN_MASKOP(operator&, &&) ENOKI_DYN_MASKOP(operator|, ||) ENOKI_DYN_MASKOP(operator^, !=)
#undef ENOKI_DYN_MASKOP
template <class A, class B, std::enable_if_t<is_mask_v<A> && is_masklike_v<B>, int> = 0> A &operator&=(A &a, const B &b) {
    const size_t n = bsize(a, b); A r = lmake<A>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) rs[i] = bool(lval(a, i)) && bool(lval(b, i));
    a = r; return a;
}
template <class A, class B, std::enable_if_t<is_mask_v<A> && is_masklike_v<B>, int> = 0> A &operator|=(A &a, const B &b) {
    const size_t n = bsize(a, b); A r = lmake<A>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) rs[i] = bool(lval(a, i)) || bool(lval(b, i));
    a = r; return a;
}
template <class A, std::enable_if_t<is_mask_v<A>, int> = 0> A operator!(const A &a) { A r = a; for (auto &x : lstore(r)) x = !x; return r; }
template <class A, std::enable_if_t<is_mask_v<A>, int> = 0> A operator~(const A &a) { return !a; }
// value & mask: zero (value and tangent) where the mask is off
template <class A, class M, std::enable_if_t<is_sc_v<A> && !is_mask_v<A> && is_masklike_v<M>, int> = 0> A operator&(const A &a, const M &m) {
    const size_t n = bsize(a, m); A r = lmake<A>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) rs[i] = bool(lval(m, i)) ? lval(a, i) : scalar_t<A>(0);
    if constexpr (is_fdiff_v<A>) if (lhas(a)) { r.g.d.resize(n); for (size_t i = 0; i < n; ++i) r.g.d[i] = bool(lval(m, i)) ? ltan(a, i) : 0.f; }
    return r;
}
template <class M, std::enable_if_t<is_mask_v<M>, int> = 0> bool any(const M &m) { for (auto x : lstore(m)) if (x) return true; return false; }
template <class M, std::enable_if_t<is_mask_v<M>, int> = 0> bool all(const M &m) { for (auto x : lstore(m)) if (!x) return false; return true; }
template <class M, std::enable_if_t<is_mask_v<M>, int> = 0> bool none(const M &m) { return !any(m); }
inline bool any(bool b) { return b; }
inline bool all(bool b) { return b; }

template <class X> struct mask_of { using type = bool; };
template <class T> struct mask_of<CUDAArray<T>> { using type = CUDAArray<bool>; };
template <class A> struct mask_of<DiffArray<A>> { using type = DiffArray<CUDAArray<bool>>; };
template <class X> using mask_t = typename mask_of<std::decay_t<X>>::type;

inline float fmadd(float a, float b, float c) { return std::fma(a, b, c); }
inline float fmsub(float a, float b, float c) { return std::fma(a, b, -c); }
inline float fnmadd(float a, float b, float c) { return std::fma(-a, b, c); }
inline float fnmsub(float a, float b, float c) { return std::fma(-a, b, -c); }
// ---- math --------------------------------------------------------------------------------------------------------------------------------
#define ENOKI_DYN_UNARY(name, VEXPR, TEXPR) \
    template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S name(const S &a) { \
        using T = scalar_t<S>; const size_t n = lsize(a); S r = lmake<S>(n); auto &rs = lstore(r); \
        for (size_t i = 0; i < n; ++i) { const T x = lval(a, i); rs[i] = T(VEXPR); } \
        if constexpr (is_fdiff_v<S>) if (lhas(a)) { r.g.d.resize(n); for (size_t i = 0; i < n; ++i) { const float x = lval(a, i), dx = ltan(a, i), z = rs[i]; (void)x; (void)z; r.g.d[i] = TEXPR; } } \
        return r; } \
    inline float name(float x) { return (float)(VEXPR); }
ENOKI_DYN_UNARY(sqr, x * x, 2.f * x * dx)
ENOKI_DYN_UNARY(sqrt, std::sqrt(x), dx / (2.f * z))
ENOKI_DYN_UNARY(safe_sqrt, std::sqrt(x > 0 ? x : 0), x > 0 ? dx / (2.f * z) : 0.f)
ENOKI_DYN_UNARY(rcp, 1.f / x, -dx * z * z)
ENOKI_DYN_UNARY(rsqrt, 1.f / std::sqrt(x), -0.5f * dx * z / x)
ENOKI_DYN_UNARY(safe_rsqrt, 1.f / std::sqrt(x > 0 ? x : 0), x > 0 ? -0.5f * dx * z / x : 0.f)
ENOKI_DYN_UNARY(abs, std::abs(x), x < 0 ? -dx : dx)
ENOKI_DYN_UNARY(sin, std::sin(x), dx * std::cos(x))
ENOKI_DYN_UNARY(cos, std::cos(x), -dx * std::sin(x))
ENOKI_DYN_UNARY(tan, std::tan(x), dx * (1.f + z * z))
ENOKI_DYN_UNARY(acos, std::acos(x), -dx / std::sqrt(1.f - x * x))
ENOKI_DYN_UNARY(asin, std::asin(x), dx / std::sqrt(1.f - x * x))
ENOKI_DYN_UNARY(exp, std::exp(x), dx * z)
ENOKI_DYN_UNARY(log, std::log(x), dx / x)
ENOKI_DYN_UNARY(safe_acos, std::acos(x < -1 ? -1 : (x > 1 ? 1 : x)), (x > -1 && x < 1) ? -dx / std::sqrt(1.f - x * x) : 0.f)
ENOKI_DYN_UNARY(floor, std::floor(x), 0.f * dx)
ENOKI_DYN_UNARY(ceil, std::ceil(x), 0.f * dx)
ENOKI_DYN_UNARY(sign, std::copysign(1.f, x), 0.f * dx)
#undef ENOKI_DYN_UNARY
inline float deg_to_rad(float a) { return a * (Pi / 180.f); }
#define ENOKI_DYN_FMA(name, VF, VI, TEXPR) \
    template <class A, class B, class C, class R = bin3_t<A, B, C>> R name(const A &a, const B &b, const C &c) { \
        using S = scalar_t<R>; const size_t n = bsize(a, b, c); R r = lmake<R>(n); auto &rs = lstore(r); \
        for (size_t i = 0; i < n; ++i) { const S x = S(lval(a, i)), y = S(lval(b, i)), w = S(lval(c, i)); if constexpr (std::is_floating_point_v<S>) rs[i] = VF; else rs[i] = VI; } \
        if constexpr (is_fdiff_v<R>) if (lhas(a) || lhas(b) || lhas(c)) { r.g.d.resize(n); \
            for (size_t i = 0; i < n; ++i) { const float x = (float)lval(a, i), y = (float)lval(b, i), dx = ltan(a, i), dy = ltan(b, i), dw = ltan(c, i); r.g.d[i] = TEXPR; } } \
        return r; }
ENOKI_DYN_FMA(fmadd, std::fma(x, y, w), x *y + w, dx *y + x * dy + dw)
ENOKI_DYN_FMA(fmsub, std::fma(x, y, -w), x *y - w, dx *y + x * dy - dw)
ENOKI_DYN_FMA(fnmadd, std::fma(-x, y, w), -x *y + w, -(dx * y + x * dy) + dw)
ENOKI_DYN_FMA(fnmsub, std::fma(-x, y, -w), -x *y - w, -(dx * y + x * dy) - dw)
#undef ENOKI_DYN_FMA
#define ENOKI_DYN_BINFN(name, VEXPR, TEXPR) \
    template <class A, class B, class R = bin_t<A, B>> R name(const A &a, const B &b) { \
        using S = scalar_t<R>; const size_t n = bsize(a, b); R r = lmake<R>(n); auto &rs = lstore(r); \
        for (size_t i = 0; i < n; ++i) { const S x = S(lval(a, i)), y = S(lval(b, i)); rs[i] = S(VEXPR); } \
        if constexpr (is_fdiff_v<R>) if (lhas(a) || lhas(b)) { r.g.d.resize(n); \
            for (size_t i = 0; i < n; ++i) { const float x = (float)lval(a, i), y = (float)lval(b, i), dx = ltan(a, i), dy = ltan(b, i), z = rs[i]; (void)x; (void)y; (void)z; (void)dy; r.g.d[i] = TEXPR; } } \
        return r; }
ENOKI_DYN_BINFN(atan2, std::atan2(x, y), (dx * y - x * dy) / (x * x + y * y))
ENOKI_DYN_BINFN(min, std::min<S>(x, y), x <= y ? dx : dy)
ENOKI_DYN_BINFN(max, std::max<S>(x, y), x >= y ? dx : dy)
ENOKI_DYN_BINFN(pow, std::pow(x, y), y * std::pow(x, y - 1.f) * dx)
ENOKI_DYN_BINFN(mulsign, (std::signbit((float)y) ? -x : x), (std::signbit(y) ? -dx : dx))
ENOKI_DYN_BINFN(mulsign_neg, (std::signbit((float)y) ? x : -x), (std::signbit(y) ? dx : -dx))
#undef ENOKI_DYN_BINFN
inline float min(float a, float b) { return std::min(a, b); }
inline float max(float a, float b) { return std::max(a, b); }
inline int min(int a, int b) { return std::min(a, b); }
inline int max(int a, int b) { return std::max(a, b); }
template <class A, class B, class C, class R = bin3_t<A, B, C>> R clamp(const A &a, const B &lo, const C &hi) { return R(min(max(a, lo), hi)); }
template <class A, class B, class C, class R = bin3_t<A, B, C>> R lerp(const A &a, const B &b, const C &t) { return fmadd(b, t, fnmadd(a, t, a)); }
template <class I, class S, std::enable_if_t<is_sc_v<S> && is_sc_v<I>, int> = 0> I floor2int(const S &a) {
    const size_t n = lsize(a); I r = lmake<I>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) { const float f = std::floor((float)lval(a, i)); rs[i] = std::isfinite(f) && std::abs(f) < 2e9f ? (scalar_t<I>)f : std::numeric_limits<scalar_t<I>>::min(); }
    return r;
}
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> std::pair<S, S> sincos(const S &a) { return {sin(a), cos(a)}; }
inline std::pair<float, float> sincos(float a) { return {std::sin(a), std::cos(a)}; }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> mask_t<S> isfinite(const S &a) {
    const size_t n = lsize(a); mask_t<S> r = lmake<mask_t<S>>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) rs[i] = std::isfinite((float)lval(a, i));
    return r;
}
template <class M, class A, class B, std::enable_if_t<is_masklike_v<M> && (is_sc_v<A> || is_sc_v<B>) && (is_sc_v<A> || is_plain_v<A>) && (is_sc_v<B> || is_plain_v<B>), int> = 0>
auto select(const M &m, const A &a, const B &b) {
    using R0 = std::conditional_t<is_sc_v<A>, std::decay_t<A>, std::decay_t<B>>;
    using R = leaf_t<scalar_t<R0>, is_diff_v<A> || is_diff_v<B> || (is_diff_v<M> && std::is_same_v<scalar_t<R0>, bool>)>;
    using S = scalar_t<R>; const size_t n = bsize(m, a, b); R r = lmake<R>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) rs[i] = bool(lval(m, i)) ? S(lval(a, i)) : S(lval(b, i));
    if constexpr (is_fdiff_v<R>) if (lhas(a) || lhas(b)) { r.g.d.resize(n); for (size_t i = 0; i < n; ++i) r.g.d[i] = bool(lval(m, i)) ? ltan(a, i) : ltan(b, i); }
    return r;
}
template <class A> auto detach(const DiffArray<A> &a) { return a.v; }
template <class T> const CUDAArray<T> &detach(const CUDAArray<T> &a) { return a; }
inline void cuda_eval() {}
inline void cuda_sync() {}
template <class S> void set_requires_gradient(S &, bool = true) {}
template <class S> bool requires_gradient(const S &) { return false; }

// horizontal reductions over the lanes (sequential fp32)
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S hsum(const S &a) {
    scalar_t<S> acc = 0; for (size_t i = 0; i < lsize(a); ++i) acc += lval(a, i);
    S r(acc);
    if constexpr (is_fdiff_v<S>) if (lhas(a)) { float t = 0; for (size_t i = 0; i < lsize(a); ++i) t += ltan(a, i); r.g = CUDAArray<float>(t); }
    return r;
}
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S hmax(const S &a) { scalar_t<S> acc = lval(a, 0); for (size_t i = 1; i < lsize(a); ++i) acc = std::max<scalar_t<S>>(acc, lval(a, i)); return S(acc); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S hmin(const S &a) { scalar_t<S> acc = lval(a, 0); for (size_t i = 1; i < lsize(a); ++i) acc = std::min<scalar_t<S>>(acc, lval(a, i)); return S(acc); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S psum(const S &a) {
    S r = lmake<S>(lsize(a)); auto &rs = lstore(r); scalar_t<S> acc = 0;
    for (size_t i = 0; i < lsize(a); ++i) { acc += lval(a, i); rs[i] = acc; }
    return r;
}
template <class T> struct divisor { T div; divisor(T d) : div(d) {} template <class S> S operator()(const S &a) const { S r = a; for (auto &x : lstore(r)) x /= div; return r; } };

// ---- fixed-size arrays -------------------------------------------------------------------------------------------------------------------
template <class T, size_t n> struct Array {
    using Value = T;
    static constexpr size_t Size = n;
    T d[n];
    Array() { for (size_t i = 0; i < n; ++i) d[i] = T(); }
    template <class U, std::enable_if_t<std::is_constructible_v<T, U> && !std::is_base_of_v<Array, std::decay_t<U>>, int> = 0> Array(const U &s) { for (size_t i = 0; i < n; ++i) d[i] = T(s); }
    template <class U, std::enable_if_t<!std::is_same_v<U, T>, int> = 0> Array(const Array<U, n> &o) { for (size_t i = 0; i < n; ++i) d[i] = T(o.d[i]); }
    template <class A0, class A1, class... Ar, std::enable_if_t<sizeof...(Ar) + 2 == n, int> = 0> Array(const A0 &a0, const A1 &a1, const Ar &...ar) { const T t[n] = {T(a0), T(a1), T(ar)...}; for (size_t i = 0; i < n; ++i) d[i] = t[i]; }
    T &x() { return d[0]; } const T &x() const { return d[0]; }
    T &y() { return d[1]; } const T &y() const { return d[1]; }
    T &z() { return d[2]; } const T &z() const { return d[2]; }
    T &w() { return d[3]; } const T &w() const { return d[3]; }
    T &operator[](size_t i) { return d[i]; } const T &operator[](size_t i) const { return d[i]; }
    T *data() { return d; } const T *data() const { return d; }
};
template <class X> struct is_arr : std::false_type {};
template <class T, size_t n> struct is_arr<Array<T, n>> : std::true_type {};
template <class X> constexpr bool is_arr_v = is_arr<std::decay_t<X>>::value;
template <class X> struct arr_size { static constexpr size_t value = 0; };
template <class T, size_t n> struct arr_size<Array<T, n>> { static constexpr size_t value = n; };
template <class... Xs> constexpr size_t arr_n = std::max({arr_size<std::decay_t<Xs>>::value...});
template <class T, size_t n> struct mask_of<Array<T, n>> { using type = Array<mask_t<T>, n>; };
template <class A> decltype(auto) elem_rt(const A &a, size_t i) { if constexpr (is_arr_v<A>) return (a.d[i]); else return (a); }
template <class A> using value_t = typename std::decay_t<A>::Value;
template <class X> constexpr bool is_elem_v = is_sc_v<X> || is_plain_v<X>;
// elementwise lifting of an n-ary function over Array operands (non-array operands broadcast)
#define ENOKI_DYN_LIFT2(name, expr) \
    template <class A, class B, std::enable_if_t<(is_arr_v<A> || is_arr_v<B>) && (is_arr_v<A> || is_elem_v<A>) && (is_arr_v<B> || is_elem_v<B>), int> = 0> auto name(const A &a, const B &b) { \
        constexpr size_t n = arr_n<A, B>; \
        auto f = [](const auto &x, const auto &y) { return expr; }; \
        Array<std::decay_t<decltype(f(elem_rt(a, 0), elem_rt(b, 0)))>, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = f(elem_rt(a, i), elem_rt(b, i)); return r; }
ENOKI_DYN_LIFT2(operator+, x + y) ENOKI_DYN_LIFT2(operator-, x - y) ENOKI_DYN_LIFT2(operator*, x * y) ENOKI_DYN_LIFT2(operator/, x / y)
ENOKI_DYN_LIFT2(operator<, x < y) ENOKI_DYN_LIFT2(operator<=, x <= y) ENOKI_DYN_LIFT2(operator>, x > y) ENOKI_DYN_LIFT2(operator>=, x >= y)
ENOKI_DYN_LIFT2(eq, eq(x, y)) ENOKI_DYN_LIFT2(neq, neq(x, y)) ENOKI_DYN_LIFT2(operator&&, x && y) ENOKI_DYN_LIFT2(operator||, x || y) ENOKI_DYN_LIFT2(operator&, x & y)
ENOKI_DYN_LIFT2(min, min(x, y)) ENOKI_DYN_LIFT2(max, max(x, y)) ENOKI_DYN_LIFT2(atan2, atan2(x, y)) ENOKI_DYN_LIFT2(pow, pow(x, y))
#undef ENOKI_DYN_LIFT2
inline bool eq(int a, int b) { return a == b; }
inline bool neq(int a, int b) { return a != b; }
inline bool eq(float a, float b) { return a == b; }
inline bool neq(float a, float b) { return a != b; }
template <class T, size_t n, std::enable_if_t<std::is_arithmetic_v<T>, int> = 0> bool operator!=(const Array<T, n> &a, const Array<T, n> &b) { for (size_t i = 0; i < n; ++i) if (a.d[i] != b.d[i]) return true; return false; }
template <class T, size_t n, std::enable_if_t<std::is_arithmetic_v<T>, int> = 0> bool operator==(const Array<T, n> &a, const Array<T, n> &b) { return !(a != b); }
#define ENOKI_DYN_LIFT3(name) \
    template <class A, class B, class C, std::enable_if_t<is_arr_v<A> || is_arr_v<B> || is_arr_v<C>, int> = 0> auto name(const A &a, const B &b, const C &c) { \
        constexpr size_t n = arr_n<A, B, C>; \
        Array<std::decay_t<decltype(name(elem_rt(a, 0), elem_rt(b, 0), elem_rt(c, 0)))>, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = name(elem_rt(a, i), elem_rt(b, i), elem_rt(c, i)); return r; }
ENOKI_DYN_LIFT3(fmadd) ENOKI_DYN_LIFT3(fmsub) ENOKI_DYN_LIFT3(fnmadd) ENOKI_DYN_LIFT3(fnmsub) ENOKI_DYN_LIFT3(select) ENOKI_DYN_LIFT3(clamp)
#undef ENOKI_DYN_LIFT3
#define ENOKI_DYN_LIFT1(name) template <class T, size_t n> auto name(const Array<T, n> &a) { Array<std::decay_t<decltype(name(a.d[0]))>, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = name(a.d[i]); return r; }
ENOKI_DYN_LIFT1(operator-) ENOKI_DYN_LIFT1(operator~) ENOKI_DYN_LIFT1(operator!) ENOKI_DYN_LIFT1(sqr) ENOKI_DYN_LIFT1(sqrt) ENOKI_DYN_LIFT1(abs) ENOKI_DYN_LIFT1(floor) ENOKI_DYN_LIFT1(rcp)
ENOKI_DYN_LIFT1(safe_sqrt) ENOKI_DYN_LIFT1(isfinite) ENOKI_DYN_LIFT1(detach) ENOKI_DYN_LIFT1(sign)
#undef ENOKI_DYN_LIFT1
template <class X, class T, size_t n, std::enable_if_t<std::is_same_v<X, Array<T, n>>, int> = 0> auto isfinite(const Array<T, n> &a) { return isfinite(a); }   // enoki::isfinite<Spectrum<ad>>(value)
#define ENOKI_DYN_AASSIGN(op) template <class T, size_t n, class B> Array<T, n> &operator op##=(Array<T, n> &a, const B &b) { a = Array<T, n>(a op b); return a; }
ENOKI_DYN_AASSIGN(+) ENOKI_DYN_AASSIGN(-) ENOKI_DYN_AASSIGN(*) ENOKI_DYN_AASSIGN(/) ENOKI_DYN_AASSIGN(&)
#undef ENOKI_DYN_AASSIGN
template <class I, class A, std::enable_if_t<is_arr_v<A>, int> = 0> I floor2int(const A &a) { I r; for (size_t i = 0; i < arr_size<A>::value; ++i) r.d[i] = floor2int<typename I::Value>(a.d[i]); return r; }
// dot as an fmadd chain: from the last component down (the oracle's and the CUDA product's form; default) or, with dot_from_first(), from the
// first component up (what Enoki's generic dot over nested arrays is believed to do): a last-bit difference, measured like matvec_plain()
inline bool &dot_from_first() { static bool on = false; return on; }
template <class T, class U, size_t n> auto dot(const Array<T, n> &a, const Array<U, n> &b) {
    if (dot_from_first()) { auto r = a.d[0] * b.d[0]; for (size_t i = 1; i < n; ++i) r = fmadd(a.d[i], b.d[i], r); return r; }
    auto r = a.d[n - 1] * b.d[n - 1]; for (size_t i = n - 1; i-- > 0;) r = fmadd(a.d[i], b.d[i], r); return r;
}
template <class T, size_t n> T squared_norm(const Array<T, n> &a) { return dot(a, a); }
template <class T, size_t n> T norm(const Array<T, n> &a) { return sqrt(squared_norm(a)); }
template <class T, size_t n> Array<T, n> normalize(const Array<T, n> &a) { return a * rsqrt(squared_norm(a)); }
template <class T, class U> auto cross(const Array<T, 3> &a, const Array<U, 3> &b) {
    using R = std::decay_t<decltype(a.d[0] * b.d[0])>;
    return Array<R, 3>(fmsub(a.d[1], b.d[2], a.d[2] * b.d[1]), fmsub(a.d[2], b.d[0], a.d[0] * b.d[2]), fmsub(a.d[0], b.d[1], a.d[1] * b.d[0]));
}
// horizontal reductions across the components
template <class T, size_t n> T hsum(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = r + a.d[i]; return r; }
template <class T, size_t n> T hmax(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = max(r, a.d[i]); return r; }
template <class T, size_t n> T hmin(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = min(r, a.d[i]); return r; }
template <class T, size_t n, std::enable_if_t<is_masklike_v<T>, int> = 0> T all(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = r && a.d[i]; return r; }
template <class T, size_t n, std::enable_if_t<is_masklike_v<T>, int> = 0> T any(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = r || a.d[i]; return r; }
template <size_t k, class T, size_t n> Array<T, k> head(const Array<T, n> &a) { Array<T, k> r; for (size_t i = 0; i < k; ++i) r.d[i] = a.d[i]; return r; }
template <size_t k, class T, size_t n> Array<T, k> tail(const Array<T, n> &a) { Array<T, k> r; for (size_t i = 0; i < k; ++i) r.d[i] = a.d[n - k + i]; return r; }
template <class A, class B> auto concat(const A &a, const B &b) {
    constexpr size_t na = is_arr_v<A> ? arr_size<A>::value : 1, nb = is_arr_v<B> ? arr_size<B>::value : 1;
    using E = std::decay_t<decltype(elem_rt(a, 0))>;
    Array<E, na + nb> r;
    for (size_t i = 0; i < na; ++i) r.d[i] = elem_rt(a, i);
    for (size_t i = 0; i < nb; ++i) r.d[na + i] = E(elem_rt(b, i));
    return r;
}
template <class T, size_t n> std::ostream &operator<<(std::ostream &os, const Array<T, n> &a) { os << "["; for (size_t i = 0; i < n; ++i) os << (i ? ", " : "") << a.d[i]; return os << "]"; }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> std::ostream &operator<<(std::ostream &os, const S &a) {
    os << "["; for (size_t i = 0; i < std::min<size_t>(lsize(a), 8); ++i) os << (i ? ", " : "") << lval(a, i); if (lsize(a) > 8) os << ", ..."; return os << "]";
}

// ---- matrices (entry (i, j) = row i, column j) ----------------------------------------------------------------------------------------------
template <class T, size_t n> struct Matrix {
    using Entry = T;
    static constexpr size_t Size = n;
    T m[n][n];
    Matrix() { for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) m[i][j] = T(i == j ? 1.f : 0.f); }
    template <class U, std::enable_if_t<!std::is_same_v<U, T>, int> = 0> Matrix(const Matrix<U, n> &o) { for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) m[i][j] = T(o.m[i][j]); }
    template <class U, size_t k, std::enable_if_t<(k > n), int> = 0> explicit Matrix(const Matrix<U, k> &o) { for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) m[i][j] = T(o.m[i][j]); }
    T &operator()(size_t i, size_t j) { return m[i][j]; } const T &operator()(size_t i, size_t j) const { return m[i][j]; }
    template <class... C> static Matrix from_cols(const C &...cols) { Matrix r; size_t j = 0; ((void)([&] { for (size_t i = 0; i < n; ++i) r.m[i][j] = T(cols.d[i]); ++j; }()), ...); return r; }
    Matrix &operator*=(const Matrix &o) { *this = *this * o; return *this; }
};
template <class X> struct is_mat : std::false_type {};
template <class T, size_t n> struct is_mat<Matrix<T, n>> : std::true_type {};
template <class X> constexpr bool is_mat_v = is_mat<std::decay_t<X>>::value;
// Default: the fmadd chain over columns that Enoki's column-major matrices lead to (assumed). matvec_plain() switches to plain left-to-right
// sums of products, the form the oracle (and the CUDA product) use, so that tests which need IDENTICAL knife-edge decisions downstream of
// the vertex transform (which coplanar edges pass the 1 - EdgeEpsilon test) can line the two up; the difference is the last bit of a vertex.
inline bool &matvec_plain() { static bool plain = false; return plain; }
template <class R, class A, class B> R mat_row(const A *a, const B *b, size_t stride, size_t n) {
    R s = a[0] * b[0];
    if (matvec_plain()) { for (size_t j = 1; j < n; ++j) s = s + a[j] * b[j * stride]; }
    else { for (size_t j = 1; j < n; ++j) s = fmadd(a[j], b[j * stride], s); }
    return s;
}
template <class T, class U, size_t n> auto operator*(const Matrix<T, n> &a, const Array<U, n> &v) {
    using R = std::decay_t<decltype(a.m[0][0] * v.d[0])>;
    Array<R, n> r;
    for (size_t i = 0; i < n; ++i) r.d[i] = mat_row<R>(a.m[i], v.d, 1, n);
    return r;
}
template <class T, class U, size_t n> auto operator*(const Matrix<T, n> &a, const Matrix<U, n> &b) {
    using R = std::decay_t<decltype(a.m[0][0] * b.m[0][0])>;
    Matrix<R, n> r;
    for (size_t c = 0; c < n; ++c) for (size_t i = 0; i < n; ++i) r.m[i][c] = mat_row<R>(a.m[i], &b.m[0][c], n, n);
    return r;
}
template <class M> M identity() { return M(); }
template <class M, class V> M diag(const V &v) { M r; for (size_t i = 0; i < M::Size; ++i) r.m[i][i] = typename M::Entry(v.d[i]); return r; }
template <class T, size_t n> Matrix<T, n> transpose(const Matrix<T, n> &a) { Matrix<T, n> r; for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) r.m[i][j] = a.m[j][i]; return r; }
template <class M> M load(const float *p) { M r; for (size_t j = 0; j < M::Size; ++j) for (size_t i = 0; i < M::Size; ++i) r.m[i][j] = typename M::Entry(p[j * M::Size + i]); return r; }   // column-major memory
template <class T, size_t n> auto detach(const Matrix<T, n> &a) { Matrix<std::decay_t<decltype(detach(a.m[0][0]))>, n> r; for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) r.m[i][j] = detach(a.m[i][j]); return r; }
template <class T> T det(const Matrix<T, 3> &a) {
    return a.m[0][0] * (a.m[1][1] * a.m[2][2] - a.m[1][2] * a.m[2][1]) - a.m[0][1] * (a.m[1][0] * a.m[2][2] - a.m[1][2] * a.m[2][0]) + a.m[0][2] * (a.m[1][0] * a.m[2][1] - a.m[1][1] * a.m[2][0]);
}
template <class T> Matrix<T, 4> inverse_cofactor(const Matrix<T, 4> &a) {   // cofactor expansion in the entries' own arithmetic (fp32)
    auto M = [&](int i, int j) -> const T & { return a.m[i][j]; };
    T s0 = M(0, 0) * M(1, 1) - M(1, 0) * M(0, 1), s1 = M(0, 0) * M(1, 2) - M(1, 0) * M(0, 2), s2 = M(0, 0) * M(1, 3) - M(1, 0) * M(0, 3);
    T s3 = M(0, 1) * M(1, 2) - M(1, 1) * M(0, 2), s4 = M(0, 1) * M(1, 3) - M(1, 1) * M(0, 3), s5 = M(0, 2) * M(1, 3) - M(1, 2) * M(0, 3);
    T c5 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c4 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3), c3 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    T c2 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3), c1 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c0 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    T inv = rcp(s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0);
    Matrix<T, 4> r;
    r.m[0][0] = (M(1, 1) * c5 - M(1, 2) * c4 + M(1, 3) * c3) * inv; r.m[0][1] = (-M(0, 1) * c5 + M(0, 2) * c4 - M(0, 3) * c3) * inv;
    r.m[0][2] = (M(3, 1) * s5 - M(3, 2) * s4 + M(3, 3) * s3) * inv; r.m[0][3] = (-M(2, 1) * s5 + M(2, 2) * s4 - M(2, 3) * s3) * inv;
    r.m[1][0] = (-M(1, 0) * c5 + M(1, 2) * c2 - M(1, 3) * c1) * inv; r.m[1][1] = (M(0, 0) * c5 - M(0, 2) * c2 + M(0, 3) * c1) * inv;
    r.m[1][2] = (-M(3, 0) * s5 + M(3, 2) * s2 - M(3, 3) * s1) * inv; r.m[1][3] = (M(2, 0) * s5 - M(2, 2) * s2 + M(2, 3) * s1) * inv;
    r.m[2][0] = (M(1, 0) * c4 - M(1, 1) * c2 + M(1, 3) * c0) * inv; r.m[2][1] = (-M(0, 0) * c4 + M(0, 1) * c2 - M(0, 3) * c0) * inv;
    r.m[2][2] = (M(3, 0) * s4 - M(3, 1) * s2 + M(3, 3) * s0) * inv; r.m[2][3] = (-M(2, 0) * s4 + M(2, 1) * s2 - M(2, 3) * s0) * inv;
    r.m[3][0] = (-M(1, 0) * c3 + M(1, 1) * c1 - M(1, 2) * c0) * inv; r.m[3][1] = (M(0, 0) * c3 - M(0, 1) * c1 + M(0, 2) * c0) * inv;
    r.m[3][2] = (-M(3, 0) * s3 + M(3, 1) * s1 - M(3, 2) * s0) * inv; r.m[3][3] = (M(2, 0) * s3 - M(2, 1) * s1 + M(2, 2) * s0) * inv;
    return r;
}
// inverse_rounded(): the 4x4 inverse computed in double precision and rounded to fp32 (tangent: -A^-1 dA A^-1), i.e. the inverse up to the
// last bit, which is also what the oracle and the CUDA product's host code compute. Default: the fp32 cofactor expansion above (Enoki's own
// fp32 formula is not known here). With it on, camera rays agree with the oracle's to the bit and so do all but a few lanes per ten thousand.
inline bool &inverse_rounded() { static bool on = false; return on; }
template <class T> Matrix<T, 4> inverse(const Matrix<T, 4> &a) {
    if (!inverse_rounded()) return inverse_cofactor(a);
    double w[4][8], dA[4][4];
    bool any_t = false;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        w[i][j] = (double)lval(a.m[i][j], 0); w[i][4 + j] = i == j ? 1.0 : 0.0;
        dA[i][j] = (double)ltan(a.m[i][j], 0); any_t = any_t || lhas(a.m[i][j]);
    }
    for (int col = 0; col < 4; ++col) {   // Gauss-Jordan elimination with partial pivoting on [A | I]
        int piv = col;
        for (int r = col + 1; r < 4; ++r) if (std::abs(w[r][col]) > std::abs(w[piv][col])) piv = r;
        if (piv != col) for (int j = 0; j < 8; ++j) std::swap(w[piv][j], w[col][j]);
        const double s = 1.0 / w[col][col];
        for (int j = 0; j < 8; ++j) w[col][j] *= s;
        for (int r = 0; r < 4; ++r) if (r != col) { const double f = w[r][col]; for (int j = 0; j < 8; ++j) w[r][j] -= f * w[col][j]; }
    }
    Matrix<T, 4> r;
    float inv[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { inv[i][j] = (float)w[i][4 + j]; r.m[i][j] = T(inv[i][j]); }
    if constexpr (is_fdiff_v<T>) if (any_t) {
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
            double t = 0;
            for (int k = 0; k < 4; ++k) for (int l = 0; l < 4; ++l) t += (double)inv[i][k] * dA[k][l] * (double)inv[l][j];
            r.m[i][j].g = CUDAArray<float>((float)-t);
        }
    }
    return r;
}
template <class M, class V> M translate(const V &v) { M r; for (size_t i = 0; i < 3; ++i) r.m[i][3] = typename M::Entry(v.d[i]); return r; }
template <class M, class V> M scale(const V &v) { M r; for (size_t i = 0; i < 3; ++i) r.m[i][i] = typename M::Entry(v.d[i]); return r; }
template <class M, class V, class F> M rotate(const V &a, const F &angle) {   // right-handed rotation about a unit axis (Rodrigues; assumed)
    using T = typename M::Entry;
    auto [s, c] = sincos(angle);
    auto cm = 1.f - c;
    M r;
    r.m[0][0] = T(fmadd(a.d[0] * a.d[0], cm, c));          r.m[0][1] = T(fmsub(a.d[0] * a.d[1], cm, a.d[2] * s)); r.m[0][2] = T(fmadd(a.d[0] * a.d[2], cm, a.d[1] * s));
    r.m[1][0] = T(fmadd(a.d[1] * a.d[0], cm, a.d[2] * s)); r.m[1][1] = T(fmadd(a.d[1] * a.d[1], cm, c));          r.m[1][2] = T(fmsub(a.d[1] * a.d[2], cm, a.d[0] * s));
    r.m[2][0] = T(fmsub(a.d[2] * a.d[0], cm, a.d[1] * s)); r.m[2][1] = T(fmadd(a.d[2] * a.d[1], cm, a.d[0] * s)); r.m[2][2] = T(fmadd(a.d[2] * a.d[2], cm, c));
    return r;
}

// ---- generic traversal of leaves: lanes, Array, Matrix, ENOKI_STRUCT types ---------------------------------------------------------------------
template <class X, class = void> struct has_fields : std::false_type {};
template <class X> struct has_fields<X, std::void_t<decltype(&X::template enoki_fields<void (*)(), X>)>> : std::true_type {};
template <class X> constexpr bool has_fields_v = has_fields<std::decay_t<X>>::value;
template <class F, class X0, class... Xs> void traverse(F &&f, X0 &x0, Xs &...xs) {
    using D = std::decay_t<X0>;
    if constexpr (is_sc_v<D> || is_plain_v<D>) f(x0, xs...);
    else if constexpr (is_arr_v<D>) { for (size_t i = 0; i < D::Size; ++i) traverse(f, x0.d[i], xs.d[i]...); }
    else if constexpr (is_mat_v<D>) { for (size_t i = 0; i < D::Size; ++i) for (size_t j = 0; j < D::Size; ++j) traverse(f, x0.m[i][j], xs.m[i][j]...); }
    else D::enoki_fields([&](auto &...fs) { traverse(f, fs...); }, x0, xs...);
}
template <class X> constexpr bool is_traversable_v = is_sc_v<X> || is_plain_v<X> || is_arr_v<X> || is_mat_v<X> || has_fields_v<X>;
template <class X, std::enable_if_t<!is_traversable_v<X>, int> = 0> size_t slices(const X &) { return 1; }   // plain structs (psdr::Ray)
template <class X, std::enable_if_t<is_traversable_v<X>, int> = 0> size_t slices(const X &x) { size_t n = 0; traverse([&](const auto &l) { n = std::max(n, lsize(l)); }, x); return n; }

template <class T> T zero(size_t n = 1) { T r; traverse([&](auto &l) { using L = std::decay_t<decltype(l)>; if constexpr (is_sc_v<L>) { l = lmake<L>(n); } else l = L(0); }, r); return r; }
template <class T> T empty(size_t n = 1) { return zero<T>(n); }
template <class T, class V> T full(const V &v, size_t n = 1) {
    T r; traverse([&](auto &l) { using L = std::decay_t<decltype(l)>; if constexpr (is_sc_v<L>) { l = lmake<L>(n); for (auto &x : lstore(l)) x = (scalar_t<L>)v; } else l = L(v); }, r); return r;
}
template <class T> T arange(size_t n = 1) { T r = lmake<T>(n); auto &rs = lstore(r); for (size_t i = 0; i < n; ++i) rs[i] = (scalar_t<T>)i; return r; }
template <class T> Array<T, 2> meshgrid(const T &x, const T &y) {   // index = iy * |x| + ix
    const size_t nx = lsize(x), ny = lsize(y); Array<T, 2> r; r.d[0] = lmake<T>(nx * ny); r.d[1] = lmake<T>(nx * ny);
    for (size_t j = 0; j < ny; ++j) for (size_t i = 0; i < nx; ++i) { lstore(r.d[0])[j * nx + i] = lval(x, i); lstore(r.d[1])[j * nx + i] = lval(y, j); }
    return r;
}

// detached type of a leaf / Array / Matrix / struct template instantiated on the float type
template <class X> struct detached { using type = X; };
template <class A> struct detached<DiffArray<A>> { using type = A; };
template <class T, size_t n> struct detached<Array<T, n>> { using type = Array<typename detached<T>::type, n>; };
template <template <class> class S, class F> struct detached<S<F>> { using type = S<typename detached<F>::type>; };
template <class T> struct detached<CUDAArray<T>> { using type = CUDAArray<T>; };
template <class X, std::enable_if_t<has_fields_v<X>, int> = 0> auto detach(const X &x) {
    typename detached<X>::type r;
    traverse([](auto &dst, const auto &src) { dst = detach(src); }, r, x);
    return r;
}

template <class L, class S, class I, class M> L gather_leaf(const S &src, const I &idx, const M &mask) {
    const size_t n = bsize(idx, mask), m = lsize(src); L r = lmake<L>(n); auto &rs = lstore(r);
    const bool t = is_fdiff_v<L> && lhas(src);
    if constexpr (is_fdiff_v<L>) if (t) r.g.d.assign(n, 0.f);
    for (size_t i = 0; i < n; ++i) {
        const int64_t k = (int64_t)lval(idx, i);
        const bool ok = bool(lval(mask, i)) && k >= 0 && (size_t)k < m;
        rs[i] = ok ? (scalar_t<L>)lval(src, (size_t)k) : scalar_t<L>(0);
        if constexpr (is_fdiff_v<L>) if (t && ok) r.g.d[i] = ltan(src, (size_t)k);
    }
    return r;
}
template <class T, class S, class I, class M = bool> T gather(const S &src, const I &idx, const M &mask = true) {
    T r; traverse([&](auto &dst, const auto &s) { dst = gather_leaf<std::decay_t<decltype(dst)>>(s, idx, mask); }, r, src); return r;
}
template <class D, class V, class I, class M> void scatter_leaf(D &dst, const V &v, const I &idx, const M &mask, bool add) {
    const size_t n = bsize(v, idx, mask); auto &ds = lstore(dst);
    bool t = false;
    if constexpr (is_fdiff_v<D>) { t = lhas(v) || lhas(dst); if (t && dst.g.d.size() != ds.size()) { const float g0 = dst.g.d.size() == 1 ? dst.g.d[0] : 0.f; dst.g.d.assign(ds.size(), g0); } }
    for (size_t i = 0; i < n; ++i) {
        const int64_t k = (int64_t)lval(idx, i);
        if (!bool(lval(mask, i)) || k < 0 || (size_t)k >= ds.size()) continue;
        if (add) ds[k] = ds[k] + (scalar_t<D>)lval(v, i); else ds[k] = (scalar_t<D>)lval(v, i);
        if constexpr (is_fdiff_v<D>) if (t) { if (add) dst.g.d[k] += ltan(v, i); else dst.g.d[k] = ltan(v, i); }
    }
}
template <class D, class V, class I, class M = bool> void scatter(D &dst, const V &v, const I &idx, const M &mask = true) {
    if constexpr (is_sc_v<D>) scatter_leaf(dst, v, idx, mask, false);
    else if constexpr (is_arr_v<D> && !is_arr_v<V> && !has_fields_v<V>) { for (size_t i = 0; i < D::Size; ++i) scatter(dst.d[i], v, idx, mask); }
    else traverse([&](auto &d, const auto &s) { scatter_leaf(d, s, idx, mask, false); }, dst, v);
}
template <class D, class V, class I, class M = bool> void scatter_add(D &dst, const V &v, const I &idx, const M &mask = true) {
    traverse([&](auto &d, const auto &s) { scatter_leaf(d, s, idx, mask, true); }, dst, v);
}
template <class A, class M> A compress(const A &a, const M &mask) {
    A r; auto &rs = lstore(r);
    for (size_t i = 0; i < bsize(a, mask); ++i) if (bool(lval(mask, i))) rs.push_back(lval(a, i));
    return r;
}
template <class P> CUDAArray<int> binary_search(int start_, int end_, const P &pred) {   // first index in [start, end] where pred is false (end if none)
    using I = CUDAArray<int>;
    I start(start_), end(end_);
    int iterations = 0;
    if (start_ < end_) { uint32_t range = (uint32_t)(end_ - start_); while (range) { ++iterations; range >>= 1; } }
    for (int it = 0; it < iterations; ++it) {
        I middle = (start + end) / 2;
        auto cond = pred(middle);
        I next = select(cond, min(middle + 1, end), start);
        end = select(cond, end, middle);
        start = next;
    }
    return start;
}

// masked assignment
template <class X, class M> struct masked_ref {
    X &ref; const M &m;
    template <class V> void operator=(const V &v) { ref = X(select(m, X(v), ref)); }
    template <class V> void operator+=(const V &v) { ref = X(select(m, X(ref + v), ref)); }
    template <class V> void operator-=(const V &v) { ref = X(select(m, X(ref - v), ref)); }
    template <class V> void operator*=(const V &v) { ref = X(select(m, X(ref * v), ref)); }
    template <class V> void operator/=(const V &v) { ref = X(select(m, X(ref / v), ref)); }
};
template <class X, class M> masked_ref<X, M> masked(X &x, const M &m) { return {x, m}; }

inline void cuda_memcpy_from_device(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }
inline void cuda_memcpy_from_device_async(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }
inline void cuda_memcpy_to_device(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }

// ---- method calls on arrays of object pointers ---------------------------------------------------------------------------------------------
template <class Storage> std::vector<scalar_t<Storage>> unique_pointers(const Storage &self) {
    std::vector<scalar_t<Storage>> u;
    for (size_t i = 0; i < lsize(self); ++i) { auto p = lval(self, i); if (p && std::find(u.begin(), u.end(), p) == u.end()) u.push_back(p); }
    return u;
}
template <class Class, class Storage, class F, class Tup, size_t... Is> auto call_dispatch_impl(const Storage &self, F &f, const Tup &args, std::index_sequence<Is...>) {
    constexpr size_t last = std::tuple_size_v<Tup> - 1;
    const auto &mask = std::get<last>(args);
    using R = decltype(f((const Class *)nullptr, std::get<Is>(args)..., mask));
    const size_t n = bsize(self, mask);
    R result = zero<R>(n);
    for (auto p : unique_pointers(self)) {
        auto mp = mask && eq(self, p);
        R r = f((const Class *)p, std::get<Is>(args)..., mp);
        traverse([&](auto &acc, const auto &v) { acc = select(mp, v, acc); }, result, r);
    }
    return result;
}
template <class Class, class Storage, class F, class... Args> auto call_dispatch(const Storage &self, F f, const Args &...args) {
    return call_dispatch_impl<Class>(self, f, std::tie(args...), std::make_index_sequence<sizeof...(Args) - 1>());
}
template <class Class, class Storage, class F, class M> auto call_getter(const Storage &self, F f, const M &mask) {
    using P = std::decay_t<decltype(f((const Class *)nullptr))>;
    using Q = std::remove_const_t<std::remove_pointer_t<P>> *;
    using R = leaf_t<Q, is_diff_v<Storage>>;
    const size_t n = bsize(self, mask); R r = lmake<R>(n); auto &rs = lstore(r);
    for (size_t i = 0; i < n; ++i) { auto p = lval(self, i); rs[i] = (p && bool(lval(mask, i))) ? const_cast<Q>(f((const Class *)p)) : nullptr; }
    return r;
}

// ---- integers, PCG32 (enoki/random.h; the generator itself is Enoki's, restated per SURVEY App. D = canonical pcg32) -------------------------
template <class S> using uint64_array_t = leaf_t<uint64_t, is_diff_v<S>>;
template <class S> using uint32_array_t = leaf_t<uint32_t, is_diff_v<S>>;
template <size_t k, class S, std::enable_if_t<is_sc_v<S>, int> = 0> S sl(const S &a) { S r = a; for (auto &x : lstore(r)) x = scalar_t<S>(x << k); return r; }
template <size_t k, class S, std::enable_if_t<is_sc_v<S>, int> = 0> S sr(const S &a) { S r = a; for (auto &x : lstore(r)) x = scalar_t<S>(x >> k); return r; }
template <class A, class B, class R = bin_t<A, B>, std::enable_if_t<std::is_integral_v<scalar_t<R>> && !is_masklike_v<A> && !is_masklike_v<B>, int> = 0>
R operator^(const A &a, const B &b) { const size_t n = bsize(a, b); R r = lmake<R>(n); auto &rs = lstore(r); for (size_t i = 0; i < n; ++i) rs[i] = scalar_t<R>(scalar_t<R>(lval(a, i)) ^ scalar_t<R>(lval(b, i))); return r; }
constexpr uint64_t PCG32_DEFAULT_STATE = 0x853c49e6748fea9bULL, PCG32_DEFAULT_STREAM = 0xda3e39cb94b95bdbULL, PCG32_MULT = 0x5851f42d4c957f2dULL;
template <class UInt32> struct PCG32 {
    using UInt64 = uint64_array_t<UInt32>;
    using Float32 = leaf_t<float, is_diff_v<UInt32>>;
    std::vector<uint64_t> state, inc;
    PCG32(uint64_t initstate = PCG32_DEFAULT_STATE, uint64_t initseq = PCG32_DEFAULT_STREAM) { seed(UInt64(initstate), UInt64(initseq)); }
    void seed(const UInt64 &initstate, const UInt64 &initseq) {
        const size_t n = bsize(initstate, initseq);
        state.assign(n, 0); inc.resize(n);
        for (size_t i = 0; i < n; ++i) inc[i] = (lval(initseq, i) << 1) | 1u;
        next_uint32();
        for (size_t i = 0; i < n; ++i) state[i] += lval(initstate, i);
        next_uint32();
    }
    UInt32 next_uint32() {
        UInt32 r = lmake<UInt32>(state.size()); auto &rs = lstore(r);
        for (size_t i = 0; i < state.size(); ++i) {
            const uint64_t old = state[i];
            state[i] = old * PCG32_MULT + inc[i];
            const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
            rs[i] = (xs >> rot) | (xs << ((~rot + 1u) & 31));
        }
        return r;
    }
    Float32 next_float32() {
        UInt32 u = next_uint32(); Float32 r = lmake<Float32>(state.size()); auto &rs = lstore(r);
        for (size_t i = 0; i < state.size(); ++i) { const uint32_t b = (lval(u, i) >> 9) | 0x3f800000u; float f; std::memcpy(&f, &b, 4); rs[i] = f - 1.f; }
        return r;
    }
};

}  // namespace enoki

#define ENOKI_CALL_SUPPORT_BEGIN(Class) namespace enoki { template <class Storage> struct call_support<Class, Storage> { using Class_ = Class; const Storage &self; const call_support *operator->() const { return this; }
#define ENOKI_CALL_SUPPORT_METHOD(name) template <class... Args> auto name(const Args &...args) const { return enoki::call_dispatch<Class_>(self, [](const Class_ *p, const auto &...a) { return p->name(a...); }, args...); }
#define ENOKI_CALL_SUPPORT_GETTER(name, field) template <class M = bool> auto name(const M &m = true) const { return enoki::call_getter<Class_>(self, [](const Class_ *p) { return p->field; }, m); }
#define ENOKI_CALL_SUPPORT_GETTER_TYPE(...)
#define ENOKI_CALL_SUPPORT_END(Class) }; }
