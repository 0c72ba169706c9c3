// TEST INFRASTRUCTURE (oracle/): a *scalar* stand-in for the subset of the Enoki API that psdr-cuda's header-only math uses
// (include/psdr/core/{warp,frame}.h, include/psdr/utils.h, src/bsdf/ggx.cpp). Enoki itself is an un-vendored external dependency of
// the reference, so the reference cannot be built as shipped; with this stub on the include path the reference's OWN
// source files compile unmodified, from where they lie under /root/reference, as plain fp32 scalar code (one "lane" per array), and
// oracle/_ref/libref_math.so exposes them to the tests that pin the oracle's restatement of the same formulas.
// Semantics follow SURVEY App. D: IEEE fp32, exact 1/x and 1/sqrt(x) for rcp / rsqrt (Enoki's CUDA backend uses the approximate
// instructions), select/masked on the primal, no autodiff tape (DiffArray carries the value only; detach is the identity).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <memory>
#include <tuple>
#include <type_traits>
#include <utility>

#define ENOKI_INLINE inline
#define ENOKI_STRUCT(Name, ...) Name() = default;
#define ENOKI_DERIVED_STRUCT(Name, Base, ...) Name() = default;
#define ENOKI_BASE_FIELDS(...)
#define ENOKI_DERIVED_FIELDS(...)
#define ENOKI_STRUCT_SUPPORT(...)
#define ENOKI_CALL_SUPPORT_BEGIN(...)
#define ENOKI_CALL_SUPPORT_METHOD(...)
#define ENOKI_CALL_SUPPORT_GETTER(...)
#define ENOKI_CALL_SUPPORT_GETTER_TYPE(...)
#define ENOKI_CALL_SUPPORT_END(...)
#define ENOKI_PIN_U1(B, a) using B::a;
#define ENOKI_PIN_U2(B, a, ...) using B::a; ENOKI_PIN_U1(B, __VA_ARGS__)
#define ENOKI_PIN_U3(B, a, ...) using B::a; ENOKI_PIN_U2(B, __VA_ARGS__)
#define ENOKI_PIN_U4(B, a, ...) using B::a; ENOKI_PIN_U3(B, __VA_ARGS__)
#define ENOKI_PIN_U5(B, a, ...) using B::a; ENOKI_PIN_U4(B, __VA_ARGS__)
#define ENOKI_PIN_U6(B, a, ...) using B::a; ENOKI_PIN_U5(B, __VA_ARGS__)
#define ENOKI_PIN_U7(B, a, ...) using B::a; ENOKI_PIN_U6(B, __VA_ARGS__)
#define ENOKI_PIN_U8(B, a, ...) using B::a; ENOKI_PIN_U7(B, __VA_ARGS__)
#define ENOKI_PIN_U9(B, a, ...) using B::a; ENOKI_PIN_U8(B, __VA_ARGS__)
#define ENOKI_PIN_U10(B, a, ...) using B::a; ENOKI_PIN_U9(B, __VA_ARGS__)
#define ENOKI_PIN_U11(B, a, ...) using B::a; ENOKI_PIN_U10(B, __VA_ARGS__)
#define ENOKI_PIN_U12(B, a, ...) using B::a; ENOKI_PIN_U11(B, __VA_ARGS__)
#define ENOKI_PIN_N(_0, _1, _2, _3, _4, _5, _6, _7, _8, _9, _10, _11, _12, N, ...) N
#define ENOKI_PIN_EXPAND(x) x
#define ENOKI_PIN_DISPATCH(B, ...) ENOKI_PIN_EXPAND(ENOKI_PIN_N(B, __VA_ARGS__, ENOKI_PIN_U12, ENOKI_PIN_U11, ENOKI_PIN_U10, ENOKI_PIN_U9, ENOKI_PIN_U8, ENOKI_PIN_U7, ENOKI_PIN_U6, ENOKI_PIN_U5, ENOKI_PIN_U4, ENOKI_PIN_U3, ENOKI_PIN_U2, ENOKI_PIN_U1)(B, __VA_ARGS__))
// the reference passes a macro that itself expands to "Base, a, b, ..." (PSDR_IMPORT_BASE_HELPER): one more expansion level
#define ENOKI_USING_MEMBERS(...) ENOKI_PIN_EXPAND(ENOKI_PIN_DISPATCH(__VA_ARGS__))

namespace enoki {

constexpr float Pi = 3.14159265358979323846f, InvPi = 0.31830988618379067154f, TwoPi = 6.28318530717958647692f, InvTwoPi = 0.15915494309189533577f;
template <class T> constexpr T Epsilon = T(1.1920929e-07) / 2;
template <class T> constexpr T Infinity = std::numeric_limits<T>::infinity();

// ---- one-lane "arrays" ---------------------------------------------------------------------------------------------------
template <class T> struct CUDAArray {
    using Scalar = T;
    static constexpr bool IsDiff = false;
    T v{};
    CUDAArray() = default;
    CUDAArray(T x) : v(x) {}
    template <class U, std::enable_if_t<std::is_arithmetic_v<U> && !std::is_same_v<U, T>, int> = 0> CUDAArray(U x) : v(T(x)) {}
    template <class U> CUDAArray(const CUDAArray<U> &o) : v(T(o.v)) {}
    static CUDAArray copy(const T *p, size_t) { return CUDAArray(p[0]); }
    const T *data() const { return &v; }
    T *data() { return &v; }
    size_t size() const { return 1; }
    T operator[](size_t) const { return v; }
};
template <class A> struct DiffArray {
    using Scalar = typename A::Scalar;
    static constexpr bool IsDiff = true;
    Scalar v{};
    DiffArray() = default;
    DiffArray(Scalar x) : v(x) {}
    template <class U, std::enable_if_t<std::is_arithmetic_v<U> && !std::is_same_v<U, Scalar>, int> = 0> DiffArray(U x) : v(Scalar(x)) {}
    DiffArray(const A &a) : v(a.v) {}
    template <class B> DiffArray(const DiffArray<B> &o) : v(Scalar(o.v)) {}
    static DiffArray copy(const Scalar *p, size_t) { return DiffArray(p[0]); }
    const Scalar *data() const { return &v; }
    size_t size() const { return 1; }
    Scalar operator[](size_t) const { return v; }
};
template <class S> struct is_sc : std::false_type {};
template <class T> struct is_sc<CUDAArray<T>> : std::true_type {};
template <class A> struct is_sc<DiffArray<A>> : std::true_type {};
template <class S> constexpr bool is_sc_v = is_sc<std::decay_t<S>>::value;
template <class S> using scalar_t = typename std::decay_t<S>::Scalar;

// result type of a binary op: Diff wins, scalar type by usual promotion
template <class S, class T> struct rebind;
template <class U, class T> struct rebind<CUDAArray<U>, T> { using type = CUDAArray<T>; };
template <class A, class T> struct rebind<DiffArray<A>, T> { using type = DiffArray<CUDAArray<T>>; };
template <class S, class T> using rebind_t = typename rebind<std::decay_t<S>, T>::type;
template <class A, class B, class = void> struct bin { };
template <class A, class B> struct bin<A, B, std::enable_if_t<is_sc_v<A> && is_sc_v<B>>> {
    using T = decltype(scalar_t<A>() + scalar_t<B>());
    using type = std::conditional_t<std::decay_t<A>::IsDiff || std::decay_t<B>::IsDiff, DiffArray<CUDAArray<T>>, CUDAArray<T>>;
};
template <class A, class B> struct bin<A, B, std::enable_if_t<is_sc_v<A> && std::is_arithmetic_v<std::decay_t<B>>>> { using type = std::decay_t<A>; };
template <class A, class B> struct bin<A, B, std::enable_if_t<std::is_arithmetic_v<std::decay_t<A>> && is_sc_v<B>>> { using type = std::decay_t<B>; };
template <class A, class B> using bin_t = typename bin<A, B>::type;
template <class S> using mask_t_sc = rebind_t<S, bool>;
template <class X> auto raw(const X &x) { if constexpr (is_sc_v<X>) return x.v; else return x; }

#define ENOKI_STUB_BINOP(op) \
    template <class A, class B, class R = bin_t<A, B>> R operator op(const A &a, const B &b) { return R(scalar_t<R>(scalar_t<R>(raw(a)) op scalar_t<R>(raw(b)))); }
ENOKI_STUB_BINOP(+) ENOKI_STUB_BINOP(-) ENOKI_STUB_BINOP(*) ENOKI_STUB_BINOP(/)
#undef ENOKI_STUB_BINOP
template <class A, class B, class R = bin_t<A, B>, std::enable_if_t<std::is_integral_v<scalar_t<R>>, int> = 0> R operator%(const A &a, const B &b) { return R(raw(a) % raw(b)); }
#define ENOKI_STUB_CMP(op) \
    template <class A, class B, class R = bin_t<A, B>> mask_t_sc<R> operator op(const A &a, const B &b) { return mask_t_sc<R>(scalar_t<R>(raw(a)) op scalar_t<R>(raw(b))); }
ENOKI_STUB_CMP(<) ENOKI_STUB_CMP(<=) ENOKI_STUB_CMP(>) ENOKI_STUB_CMP(>=)
#undef ENOKI_STUB_CMP
template <class A, class B, class R = bin_t<A, B>> mask_t_sc<R> eq(const A &a, const B &b) { return mask_t_sc<R>(scalar_t<R>(raw(a)) == scalar_t<R>(raw(b))); }
template <class A, class B, class R = bin_t<A, B>> mask_t_sc<R> neq(const A &a, const B &b) { return mask_t_sc<R>(scalar_t<R>(raw(a)) != scalar_t<R>(raw(b))); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> mask_t_sc<S> neq(const S &a, std::nullptr_t) { return mask_t_sc<S>(a.v != nullptr); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> mask_t_sc<S> eq(const S &a, std::nullptr_t) { return mask_t_sc<S>(a.v == nullptr); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S operator-(const S &a) { return S(scalar_t<S>(-a.v)); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S operator+(const S &a) { return a; }
#define ENOKI_STUB_ASSIGN(op) template <class S, class B, std::enable_if_t<is_sc_v<S>, int> = 0> S &operator op##=(S &a, const B &b) { a = S(a op b); return a; }
ENOKI_STUB_ASSIGN(+) ENOKI_STUB_ASSIGN(-) ENOKI_STUB_ASSIGN(*) ENOKI_STUB_ASSIGN(/)
#undef ENOKI_STUB_ASSIGN
// masks (scalars of bool) and value & mask
template <class A, class B> using mbin_t = std::conditional_t<std::decay_t<A>::IsDiff || std::decay_t<B>::IsDiff, DiffArray<CUDAArray<bool>>, CUDAArray<bool>>;
template <class A, class B, std::enable_if_t<is_sc_v<A> && is_sc_v<B> && std::is_same_v<scalar_t<A>, bool> && std::is_same_v<scalar_t<B>, bool>, int> = 0>
mbin_t<A, B> operator&&(const A &a, const B &b) { return mbin_t<A, B>(a.v && b.v); }
template <class A, class B, std::enable_if_t<is_sc_v<A> && is_sc_v<B> && std::is_same_v<scalar_t<A>, bool> && std::is_same_v<scalar_t<B>, bool>, int> = 0>
mbin_t<A, B> operator||(const A &a, const B &b) { return mbin_t<A, B>(a.v || b.v); }
template <class A, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A operator&&(const A &a, bool b) { return A(a.v && b); }
template <class A, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A operator&&(bool b, const A &a) { return A(a.v && b); }
template <class A, class B, std::enable_if_t<is_sc_v<A> && is_sc_v<B> && std::is_same_v<scalar_t<B>, bool>, int> = 0>
std::decay_t<A> operator&(const A &a, const B &m) { return m.v ? a : std::decay_t<A>(scalar_t<A>(0)); }
template <class A, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A operator&(const A &a, bool m) { return A(a.v && m); }
template <class A, class B, std::enable_if_t<is_sc_v<A> && is_sc_v<B> && std::is_same_v<scalar_t<A>, bool> && std::is_same_v<scalar_t<B>, bool>, int> = 0>
mbin_t<A, B> operator|(const A &a, const B &b) { return mbin_t<A, B>(a.v || b.v); }
template <class A, class B, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A &operator&=(A &a, const B &b) { a.v = a.v && bool(raw(b)); return a; }
template <class A, class B, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A &operator|=(A &a, const B &b) { a.v = a.v || bool(raw(b)); return a; }
template <class A, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A operator!(const A &a) { return A(!a.v); }
template <class A, std::enable_if_t<is_sc_v<A> && std::is_same_v<scalar_t<A>, bool>, int> = 0> A operator~(const A &a) { return A(!a.v); }
template <class M> bool any(const M &m) { return bool(raw(m)); }
template <class M> bool all(const M &m) { return bool(raw(m)); }
template <class M> bool none(const M &m) { return !bool(raw(m)); }

template <class X> struct mask_of { using type = bool; };
template <class T> struct mask_of<CUDAArray<T>> { using type = CUDAArray<bool>; };
template <class A> struct mask_of<DiffArray<A>> { using type = DiffArray<CUDAArray<bool>>; };
template <class X> using mask_t = typename mask_of<std::decay_t<X>>::type;

// ---- scalar math --------------------------------------------------------------------------------------------------------------
template <class S> S mk(const S &, scalar_t<S> x) { return S(x); }
#define ENOKI_STUB_UNARY(name, expr) template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S name(const S &a) { const scalar_t<S> x = a.v; return S(scalar_t<S>(expr)); } \
                                     inline float name(float x) { return (float)(expr); }
ENOKI_STUB_UNARY(sqr, x * x) ENOKI_STUB_UNARY(sqrt, std::sqrt(x)) ENOKI_STUB_UNARY(safe_sqrt, std::sqrt(x > 0 ? x : 0)) ENOKI_STUB_UNARY(rcp, 1.f / x)
ENOKI_STUB_UNARY(rsqrt, 1.f / std::sqrt(x)) ENOKI_STUB_UNARY(safe_rsqrt, 1.f / std::sqrt(x > 0 ? x : 0)) ENOKI_STUB_UNARY(abs, std::fabs(x)) ENOKI_STUB_UNARY(sin, std::sin(x))
ENOKI_STUB_UNARY(cos, std::cos(x)) ENOKI_STUB_UNARY(tan, std::tan(x)) ENOKI_STUB_UNARY(acos, std::acos(x)) ENOKI_STUB_UNARY(asin, std::asin(x)) ENOKI_STUB_UNARY(exp, std::exp(x)) ENOKI_STUB_UNARY(log, std::log(x))
ENOKI_STUB_UNARY(safe_acos, std::acos(x < -1 ? -1 : (x > 1 ? 1 : x))) ENOKI_STUB_UNARY(floor, std::floor(x)) ENOKI_STUB_UNARY(ceil, std::ceil(x))
#undef ENOKI_STUB_UNARY
template <class A, class B, class C, class R = bin_t<bin_t<A, B>, C>> R fmadd(const A &a, const B &b, const C &c) { return R(std::fmaf(raw(a), raw(b), raw(c))); }
template <class A, class B, class C, class R = bin_t<bin_t<A, B>, C>> R fmsub(const A &a, const B &b, const C &c) { return R(std::fmaf(raw(a), raw(b), -raw(c))); }
template <class A, class B, class C, class R = bin_t<bin_t<A, B>, C>> R fnmadd(const A &a, const B &b, const C &c) { return R(std::fmaf(-raw(a), raw(b), raw(c))); }
template <class A, class B, class C, class R = bin_t<bin_t<A, B>, C>> R fnmsub(const A &a, const B &b, const C &c) { return R(std::fmaf(-raw(a), raw(b), -raw(c))); }
template <class A, class B, class R = bin_t<A, B>> R atan2(const A &y, const B &x) { return R(std::atan2(raw(y), raw(x))); }
template <class A, class B, class R = bin_t<A, B>> R min(const A &a, const B &b) { return R(std::min<scalar_t<R>>(raw(a), raw(b))); }
template <class A, class B, class R = bin_t<A, B>> R max(const A &a, const B &b) { return R(std::max<scalar_t<R>>(raw(a), raw(b))); }
template <class A, class B, class C, class R = bin_t<bin_t<A, B>, C>> R clamp(const A &a, const B &lo, const C &hi) { return R(std::min<scalar_t<R>>(std::max<scalar_t<R>>(raw(a), raw(lo)), raw(hi))); }
template <class A, class B, class C, class R = bin_t<bin_t<A, B>, C>> R lerp(const A &a, const B &b, const C &t) { return fmadd(b, t, fnmadd(a, t, a)); }
template <class I, class S, std::enable_if_t<is_sc_v<S>, int> = 0> I floor2int(const S &a) { return I((scalar_t<I>)std::floor(a.v)); }
template <class I, class S, std::enable_if_t<is_sc_v<S>, int> = 0> I ceil2int(const S &a) { return I((scalar_t<I>)std::ceil(a.v)); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> std::pair<S, S> sincos(const S &a) { return {S(std::sin(a.v)), S(std::cos(a.v))}; }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S sign(const S &a) { return S(std::copysign(scalar_t<S>(1), a.v)); }
template <class A, class B, class R = bin_t<A, B>> R mulsign(const A &a, const B &b) { return R(std::signbit((float)raw(b)) ? -raw(a) : raw(a)); }
template <class A, class B, class R = bin_t<A, B>> R mulsign_neg(const A &a, const B &b) { return R(std::signbit((float)raw(b)) ? raw(a) : -raw(a)); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> mask_t_sc<S> isfinite(const S &a) { return mask_t_sc<S>(std::isfinite(a.v)); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> mask_t_sc<S> isnan(const S &a) { return mask_t_sc<S>(std::isnan(a.v)); }
template <class M, class A, class B, class R = bin_t<A, B>, std::enable_if_t<is_sc_v<M>, int> = 0> R select(const M &m, const A &a, const B &b) { return m.v ? R(raw(a)) : R(raw(b)); }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> size_t slices(const S &) { return 1; }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S hsum(const S &a) { return a; }
template <class S, std::enable_if_t<is_sc_v<S>, int> = 0> S hmax(const S &a) { return a; }
template <class A> CUDAArray<typename A::Scalar> detach(const DiffArray<A> &a) { return CUDAArray<typename A::Scalar>(a.v); }
template <class T> const CUDAArray<T> &detach(const CUDAArray<T> &a) { return a; }
template <class S> struct masked_ref {
    S &ref; bool m;
    template <class V> void operator=(const V &v) { if (m) ref = S(v); }
    template <class V> void operator+=(const V &v) { if (m) ref = S(ref + v); }
    template <class V> void operator-=(const V &v) { if (m) ref = S(ref - v); }
    template <class V> void operator*=(const V &v) { if (m) ref = S(ref * v); }
};
template <class S, class M> masked_ref<S> masked(S &x, const M &m) { return {x, bool(raw(m))}; }
template <class T> T zero(size_t = 1) { return T(0); }
template <class T, class V> T full(const V &v, size_t = 1) { return T(v); }
template <class T> T empty(size_t = 1) { return T(); }
// the lane this one-lane "array" stands for: arange<T>(n) is its index (set by the shim before a call that seeds per lane)
inline uint64_t &stub_lane() { static thread_local uint64_t lane = 0; return lane; }
template <class T> T arange(size_t = 1) { return T((scalar_t<T>)stub_lane()); }
inline void cuda_eval() {}
inline void cuda_sync() {}
template <class S> void set_requires_gradient(S &, bool = true) {}
template <class S> bool requires_gradient(const S &) { return false; }

// ---- fixed-size arrays ------------------------------------------------------------------------------------------------------------
template <class T, size_t n> struct Array {
    using Value = T;
    static constexpr size_t Size = n;
    T d[n];
    Array() { for (size_t i = 0; i < n; ++i) d[i] = T(); }
    template <class U, std::enable_if_t<std::is_constructible_v<T, U> && !std::is_base_of_v<Array, std::decay_t<U>>, int> = 0> Array(const U &s) { for (size_t i = 0; i < n; ++i) d[i] = T(s); }
    template <class U> Array(const Array<U, n> &o) { for (size_t i = 0; i < n; ++i) d[i] = T(o.d[i]); }
    template <class A0, class A1, class... Ar, std::enable_if_t<sizeof...(Ar) + 2 == n, int> = 0> Array(const A0 &a0, const A1 &a1, const Ar &...ar) { const T t[n] = {T(a0), T(a1), T(ar)...}; for (size_t i = 0; i < n; ++i) d[i] = t[i]; }
    T &x() { return d[0]; } const T &x() const { return d[0]; }
    T &y() { return d[1]; } const T &y() const { return d[1]; }
    T &z() { return d[2]; } const T &z() const { return d[2]; }
    T &w() { return d[3]; } const T &w() const { return d[3]; }
    T &operator[](size_t i) { return d[i]; } const T &operator[](size_t i) const { return d[i]; }
};
template <class X> struct is_arr : std::false_type {};
template <class T, size_t n> struct is_arr<Array<T, n>> : std::true_type {};
template <class X> constexpr bool is_arr_v = is_arr<std::decay_t<X>>::value;
template <class T, size_t n> struct mask_of<Array<T, n>> { using type = Array<mask_t<T>, n>; };
template <class A, class B, class = void> struct abin {};
template <class T, class U, size_t n> struct abin<Array<T, n>, Array<U, n>> { using type = Array<decltype(std::declval<T>() + std::declval<U>()), n>; };
template <class T, size_t n, class U> struct abin<Array<T, n>, U, std::enable_if_t<!is_arr_v<U>>> { using type = Array<decltype(std::declval<T>() + std::declval<U>()), n>; };
template <class U, class T, size_t n> struct abin<U, Array<T, n>, std::enable_if_t<!is_arr_v<U>>> { using type = Array<decltype(std::declval<U>() + std::declval<T>()), n>; };
template <class A, class B> using abin_t = typename abin<std::decay_t<A>, std::decay_t<B>>::type;
template <class A, size_t i> decltype(auto) elem(const A &a) { if constexpr (is_arr_v<A>) return a.d[i]; else return a; }
template <class A> decltype(auto) elem_rt(const A &a, size_t i) { if constexpr (is_arr_v<A>) return a.d[i]; else return a; }
#define ENOKI_STUB_ABIN(op) \
    template <class A, class B, std::enable_if_t<is_arr_v<A> || is_arr_v<B>, int> = 0, class R = abin_t<A, B>> R operator op(const A &a, const B &b) { R r; for (size_t i = 0; i < R::Size; ++i) r.d[i] = typename R::Value(elem_rt(a, i) op elem_rt(b, i)); return r; }
ENOKI_STUB_ABIN(+) ENOKI_STUB_ABIN(-) ENOKI_STUB_ABIN(*) ENOKI_STUB_ABIN(/)
#undef ENOKI_STUB_ABIN
template <class T, size_t n> Array<T, n> operator-(const Array<T, n> &a) { Array<T, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = -a.d[i]; return r; }
#define ENOKI_STUB_AASSIGN(op) template <class T, size_t n, class B> Array<T, n> &operator op##=(Array<T, n> &a, const B &b) { a = Array<T, n>(a op b); return a; }
ENOKI_STUB_AASSIGN(+) ENOKI_STUB_AASSIGN(-) ENOKI_STUB_AASSIGN(*) ENOKI_STUB_AASSIGN(/)
#undef ENOKI_STUB_AASSIGN
template <class T, size_t n, class M, std::enable_if_t<is_sc_v<M>, int> = 0> Array<T, n> operator&(const Array<T, n> &a, const M &m) { Array<T, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = a.d[i] & m; return r; }
template <class T, size_t n, class M, std::enable_if_t<is_sc_v<M>, int> = 0> Array<T, n> &operator&=(Array<T, n> &a, const M &m) { a = a & m; return a; }
// assumed (Enoki is not available to check): dot = fmadd(a0, b0, fmadd(a1, b1, a2 * b2)), the form the oracle uses
template <class T, class U, size_t n> auto dot(const Array<T, n> &a, const Array<U, n> &b) { auto r = a.d[n - 1] * b.d[n - 1]; for (size_t i = n - 1; i-- > 0;) r = fmadd(a.d[i], b.d[i], r); return r; }
template <class T, size_t n> T squared_norm(const Array<T, n> &a) { return dot(a, a); }
template <class T, size_t n> T norm(const Array<T, n> &a) { return sqrt(squared_norm(a)); }
template <class T, size_t n> Array<T, n> normalize(const Array<T, n> &a) { return a * rsqrt(squared_norm(a)); }
template <class T, class U> auto cross(const Array<T, 3> &a, const Array<U, 3> &b) {
    using R = decltype(a.d[0] * b.d[0]);
    return Array<R, 3>(fmsub(a.d[1], b.d[2], a.d[2] * b.d[1]), fmsub(a.d[2], b.d[0], a.d[0] * b.d[2]), fmsub(a.d[0], b.d[1], a.d[1] * b.d[0]));
}
template <class T, size_t n> T hsum(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = r + a.d[i]; return r; }
template <class T, size_t n> T hmax(const Array<T, n> &a) { T r = a.d[0]; for (size_t i = 1; i < n; ++i) r = max(r, a.d[i]); return r; }
template <size_t k, class T, size_t n> Array<T, k> head(const Array<T, n> &a) { Array<T, k> r; for (size_t i = 0; i < k; ++i) r.d[i] = a.d[i]; return r; }
template <size_t k, class T, size_t n> Array<T, k> tail(const Array<T, n> &a) { Array<T, k> r; for (size_t i = 0; i < k; ++i) r.d[i] = a.d[n - k + i]; return r; }
template <class T, size_t n> auto detach(const Array<T, n> &a) { Array<std::decay_t<decltype(detach(a.d[0]))>, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = detach(a.d[i]); return r; }
template <class M, class T, class U, size_t n, std::enable_if_t<is_sc_v<M>, int> = 0> auto select(const M &m, const Array<T, n> &a, const Array<U, n> &b) { abin_t<Array<T, n>, Array<U, n>> r; for (size_t i = 0; i < n; ++i) r.d[i] = select(m, a.d[i], b.d[i]); return r; }
template <class T, size_t n> size_t slices(const Array<T, n> &) { return 1; }
#define ENOKI_STUB_AUNARY(name) template <class T, size_t n> Array<T, n> name(const Array<T, n> &a) { Array<T, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = name(a.d[i]); return r; }
ENOKI_STUB_AUNARY(sqr) ENOKI_STUB_AUNARY(sqrt) ENOKI_STUB_AUNARY(abs) ENOKI_STUB_AUNARY(floor) ENOKI_STUB_AUNARY(rcp) ENOKI_STUB_AUNARY(safe_sqrt)
#undef ENOKI_STUB_AUNARY
template <class T, size_t n> auto isfinite(const Array<T, n> &a) { Array<mask_t<T>, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = isfinite(a.d[i]); return r; }
template <class T, size_t n> struct masked_aref { Array<T, n> &ref; bool m; template <class V> void operator=(const V &v) { if (m) ref = Array<T, n>(v); } template <class V> void operator+=(const V &v) { if (m) ref = Array<T, n>(ref + v); } };
template <class T, size_t n, class M> masked_aref<T, n> masked(Array<T, n> &x, const M &m) { return {x, bool(raw(m))}; }

// fused multiply-adds with array operands (elementwise, scalars broadcast)
#define ENOKI_STUB_AFMA(name) \
    template <class A, class B, class C, std::enable_if_t<is_arr_v<A> || is_arr_v<B> || is_arr_v<C>, int> = 0> auto name(const A &a, const B &b, const C &c) { \
        constexpr size_t n = is_arr_v<A> ? arr_size<A>::value : (is_arr_v<B> ? arr_size<B>::value : arr_size<C>::value); \
        using E = decltype(name(elem_rt(a, 0), elem_rt(b, 0), elem_rt(c, 0))); \
        Array<E, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = name(elem_rt(a, i), elem_rt(b, i), elem_rt(c, i)); return r; }
template <class X> struct arr_size { static constexpr size_t value = 0; };
template <class T, size_t n> struct arr_size<Array<T, n>> { static constexpr size_t value = n; };
ENOKI_STUB_AFMA(fmadd) ENOKI_STUB_AFMA(fmsub) ENOKI_STUB_AFMA(fnmadd) ENOKI_STUB_AFMA(fnmsub)
#undef ENOKI_STUB_AFMA
template <class T, class U, size_t n> auto max(const Array<T, n> &a, const Array<U, n> &b) { abin_t<Array<T, n>, Array<U, n>> r; for (size_t i = 0; i < n; ++i) r.d[i] = max(a.d[i], b.d[i]); return r; }
template <class T, class U, size_t n> auto min(const Array<T, n> &a, const Array<U, n> &b) { abin_t<Array<T, n>, Array<U, n>> r; for (size_t i = 0; i < n; ++i) r.d[i] = min(a.d[i], b.d[i]); return r; }
template <class X, std::enable_if_t<!is_sc_v<X> && !is_arr_v<X>, int> = 0> size_t slices(const X &) { return 1; }   // plain structs (psdr::Ray)
template <class I, class A, std::enable_if_t<is_arr_v<A>, int> = 0> I floor2int(const A &a) { I r; for (size_t i = 0; i < arr_size<A>::value; ++i) r.d[i] = floor2int<typename I::Value>(a.d[i]); return r; }
template <class A> using value_t = typename std::decay_t<A>::Value;
template <class T, size_t n, class L, class H> Array<T, n> clamp(const Array<T, n> &a, const L &lo, const H &hi) { Array<T, n> r; for (size_t i = 0; i < n; ++i) r.d[i] = T(clamp(a.d[i], lo, hi)); return r; }
// wavefront plumbing the math never reaches on one lane
template <class T, class S, class I> T gather(const S &src, const I &, bool = true) { return T(src); }
template <class T, class S, class I, class M> T gather(const S &src, const I &, const M &) { return T(src); }
template <class A, class M> A compress(const A &a, const M &) { return a; }
template <class D, class V, class I> void scatter(D &dst, const V &v, const I &) { dst = D(v); }
template <class D, class V, class I, class M> void scatter(D &dst, const V &v, const I &, const M &m) { if (bool(raw(m))) dst = D(v); }
template <class D, class V, class I> void scatter_add(D &dst, const V &v, const I &) { dst = D(dst + v); }
template <class D, class V, class I, class M> void scatter_add(D &dst, const V &v, const I &, const M &m) { if (bool(raw(m))) dst = D(dst + v); }
inline void cuda_memcpy_from_device(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }
inline void cuda_memcpy_from_device_async(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }
inline void cuda_memcpy_to_device(void *dst, const void *src, size_t n) { std::memcpy(dst, src, n); }

// ---- integers, PCG32 (enoki/random.h; the generator itself is Enoki's, restated per SURVEY App. D = canonical pcg32) -------------------
template <class S> using uint64_array_t = rebind_t<S, uint64_t>;
template <class S> using uint32_array_t = rebind_t<S, uint32_t>;
template <size_t k, class S, std::enable_if_t<is_sc_v<S>, int> = 0> S sl(const S &a) { return S(scalar_t<S>(a.v << k)); }
template <size_t k, class S, std::enable_if_t<is_sc_v<S>, int> = 0> S sr(const S &a) { return S(scalar_t<S>(a.v >> k)); }
template <class A, class B, class R = bin_t<A, B>, std::enable_if_t<std::is_integral_v<scalar_t<R>> && !std::is_same_v<scalar_t<R>, bool>, int> = 0>
R operator^(const A &a, const B &b) { return R(scalar_t<R>(scalar_t<R>(raw(a)) ^ scalar_t<R>(raw(b)))); }
template <class A, class B, class R = bin_t<A, B>, std::enable_if_t<std::is_integral_v<scalar_t<R>> && !std::is_same_v<scalar_t<R>, bool> && !std::is_same_v<scalar_t<std::conditional_t<is_sc_v<B>, B, A>>, bool>, int> = 0>
R operator|(const A &a, const B &b) { return R(scalar_t<R>(scalar_t<R>(raw(a)) | scalar_t<R>(raw(b)))); }
constexpr uint64_t PCG32_DEFAULT_STATE = 0x853c49e6748fea9bULL, PCG32_DEFAULT_STREAM = 0xda3e39cb94b95bdbULL, PCG32_MULT = 0x5851f42d4c957f2dULL;
template <class UInt32> struct PCG32 {
    using UInt64 = uint64_array_t<UInt32>;
    using Float32 = rebind_t<UInt32, float>;
    uint64_t state = 0, inc = 0;
    PCG32(uint64_t initstate = PCG32_DEFAULT_STATE, uint64_t initseq = PCG32_DEFAULT_STREAM) { seed(UInt64(initstate), UInt64(initseq)); }
    void seed(const UInt64 &initstate, const UInt64 &initseq) { state = 0; inc = (initseq.v << 1) | 1u; next_uint32(); state += initstate.v; next_uint32(); }
    UInt32 next_uint32() {
        const uint64_t old = state;
        state = old * PCG32_MULT + inc;
        const uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u), rot = (uint32_t)(old >> 59u);
        return UInt32((xs >> rot) | (xs << ((~rot + 1u) & 31)));
    }
    Float32 next_float32() { const uint32_t u = (next_uint32().v >> 9) | 0x3f800000u; float f; std::memcpy(&f, &u, 4); return Float32(f - 1.f); }
};
// concat of scalars / arrays into one array
template <class A, class B> auto concat(const A &a, const B &b) {
    constexpr size_t na = is_arr_v<A> ? arr_size<A>::value : 1, nb = is_arr_v<B> ? arr_size<B>::value : 1;
    using E = std::decay_t<decltype(elem_rt(a, 0))>;
    Array<E, na + nb> r;
    for (size_t i = 0; i < na; ++i) r.d[i] = elem_rt(a, i);
    for (size_t i = 0; i < nb; ++i) r.d[na + i] = elem_rt(b, i);
    return r;
}

template <class T, size_t n> struct Matrix {
    T m[n][n];
    Matrix() { for (size_t i = 0; i < n; ++i) for (size_t j = 0; j < n; ++j) m[i][j] = T(i == j ? 1 : 0); }
    T &operator()(size_t i, size_t j) { return m[i][j]; } const T &operator()(size_t i, size_t j) const { return m[i][j]; }
};

}  // namespace enoki
