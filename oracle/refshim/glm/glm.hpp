// Stand-in for the subset of glm (0.9.5-era semantics) the reference uses. Written for this repository; see ../README.md.
#pragma once
#include <algorithm>
#include <cassert>   // the reference uses assert() without including <cassert> itself (real glm pulls it in)
#include <cmath>
#include <cstddef>

namespace glm {

template <class T> struct tvec2 {
    union { T x, r, s; }; union { T y, g, t; };
    tvec2() : x(0), y(0) {}
    explicit tvec2(T v) : x(v), y(v) {}
    tvec2(T a, T b) : x(a), y(b) {}
    template <class U> explicit tvec2(const tvec2<U>& o) : x((T)o.x), y((T)o.y) {}
    T& operator[](int i) { return i == 0 ? x : y; }
    const T& operator[](int i) const { return i == 0 ? x : y; }
    static constexpr int N = 2;
};
template <class T> struct tvec3 {
    union { T x, r, s; }; union { T y, g, t; }; union { T z, b, p; };
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T v) : x(v), y(v), z(v) {}
    tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
    tvec3(const tvec2<T>& ab, T c) : x(ab.x), y(ab.y), z(c) {}
    template <class U> explicit tvec3(const tvec3<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    static constexpr int N = 3;
};
typedef tvec2<double> dvec2; typedef tvec3<double> dvec3; typedef tvec2<float> vec2; typedef tvec3<float> vec3;
typedef tvec2<int> ivec2; typedef tvec3<int> ivec3;

#define NGI_GLM_VEC_OPS(V, BODY2, BODY3)
// ---- vec2 ----
template <class T> inline tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <class T> inline tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <class T> inline tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <class T> inline tvec2<T> operator/(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x / b.x, a.y / b.y); }
template <class T> inline tvec2<T> operator+(const tvec2<T>& a, T s) { return tvec2<T>(a.x + s, a.y + s); }
template <class T> inline tvec2<T> operator-(const tvec2<T>& a, T s) { return tvec2<T>(a.x - s, a.y - s); }
template <class T> inline tvec2<T> operator*(const tvec2<T>& a, T s) { return tvec2<T>(a.x * s, a.y * s); }
template <class T> inline tvec2<T> operator/(const tvec2<T>& a, T s) { return tvec2<T>(a.x / s, a.y / s); }
template <class T> inline tvec2<T> operator*(T s, const tvec2<T>& a) { return tvec2<T>(s * a.x, s * a.y); }
template <class T> inline tvec2<T> operator-(const tvec2<T>& a) { return tvec2<T>(-a.x, -a.y); }
template <class T> inline tvec2<T>& operator+=(tvec2<T>& a, const tvec2<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <class T> inline tvec2<T>& operator-=(tvec2<T>& a, const tvec2<T>& b) { a.x -= b.x; a.y -= b.y; return a; }
template <class T> inline tvec2<T>& operator*=(tvec2<T>& a, T s) { a.x *= s; a.y *= s; return a; }
template <class T> inline tvec2<T>& operator/=(tvec2<T>& a, T s) { a.x /= s; a.y /= s; return a; }
template <class T> inline bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <class T> inline bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }
// ---- vec3 ----
template <class T> inline tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> inline tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> inline tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <class T> inline tvec3<T> operator/(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x / b.x, a.y / b.y, a.z / b.z); }
template <class T> inline tvec3<T> operator+(const tvec3<T>& a, T s) { return tvec3<T>(a.x + s, a.y + s, a.z + s); }
template <class T> inline tvec3<T> operator-(const tvec3<T>& a, T s) { return tvec3<T>(a.x - s, a.y - s, a.z - s); }
template <class T> inline tvec3<T> operator*(const tvec3<T>& a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template <class T> inline tvec3<T> operator/(const tvec3<T>& a, T s) { return tvec3<T>(a.x / s, a.y / s, a.z / s); }
template <class T> inline tvec3<T> operator*(T s, const tvec3<T>& a) { return tvec3<T>(s * a.x, s * a.y, s * a.z); }
template <class T> inline tvec3<T> operator+(T s, const tvec3<T>& a) { return tvec3<T>(s + a.x, s + a.y, s + a.z); }
template <class T> inline tvec3<T> operator-(T s, const tvec3<T>& a) { return tvec3<T>(s - a.x, s - a.y, s - a.z); }
template <class T> inline tvec3<T> operator/(T s, const tvec3<T>& a) { return tvec3<T>(s / a.x, s / a.y, s / a.z); }
template <class T> inline tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <class T> inline tvec3<T>& operator+=(tvec3<T>& a, const tvec3<T>& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
template <class T> inline tvec3<T>& operator-=(tvec3<T>& a, const tvec3<T>& b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
template <class T> inline tvec3<T>& operator*=(tvec3<T>& a, const tvec3<T>& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; return a; }
template <class T> inline tvec3<T>& operator/=(tvec3<T>& a, const tvec3<T>& b) { a.x /= b.x; a.y /= b.y; a.z /= b.z; return a; }
template <class T> inline tvec3<T>& operator*=(tvec3<T>& a, T s) { a.x *= s; a.y *= s; a.z *= s; return a; }
template <class T> inline tvec3<T>& operator/=(tvec3<T>& a, T s) { a.x /= s; a.y /= s; a.z /= s; return a; }
template <class T> inline bool operator==(const tvec3<T>& a, const tvec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <class T> inline bool operator!=(const tvec3<T>& a, const tvec3<T>& b) { return !(a == b); }

// ---- scalar functions ----
inline double sqrt(double x) { return std::sqrt(x); }   inline float sqrt(float x) { return std::sqrt(x); }
inline double abs(double x) { return std::fabs(x); }     inline float abs(float x) { return std::fabs(x); }   inline int abs(int x) { return x < 0 ? -x : x; }
inline double sin(double x) { return std::sin(x); }      inline double cos(double x) { return std::cos(x); }   inline double tan(double x) { return std::tan(x); }
inline double exp(double x) { return std::exp(x); }      inline double pow(double x, double y) { return std::pow(x, y); }
inline double floor(double x) { return std::floor(x); }  inline double fract(double x) { return x - std::floor(x); }
inline double radians(double deg) { return deg * 0.01745329251994329576923690768489; }
template <class T> inline T min(T a, T b) { return b < a ? b : a; }
template <class T> inline T max(T a, T b) { return a < b ? b : a; }
template <class T> inline T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
// ---- vector functions ----
template <class T> inline T dot(const tvec2<T>& a, const tvec2<T>& b) { return a.x * b.x + a.y * b.y; }
template <class T> inline T dot(const tvec3<T>& a, const tvec3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <class T> inline tvec3<T> cross(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
template <class V> inline auto length(const V& v) -> decltype(dot(v, v)) { return std::sqrt(dot(v, v)); }
template <class V> inline auto length2(const V& v) -> decltype(dot(v, v)) { return dot(v, v); }
template <class V> inline V normalize(const V& v) { return v * (decltype(dot(v, v))(1) / std::sqrt(dot(v, v))); }   // v * inversesqrt(dot(v, v))
template <class T> inline tvec3<T> min(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <class T> inline tvec3<T> max(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <class T> inline tvec3<T> abs(const tvec3<T>& a) { return tvec3<T>(abs(a.x), abs(a.y), abs(a.z)); }
template <class T> inline tvec3<T> sqrt(const tvec3<T>& a) { return tvec3<T>(std::sqrt(a.x), std::sqrt(a.y), std::sqrt(a.z)); }
template <class T> inline tvec3<T> exp(const tvec3<T>& a) { return tvec3<T>(std::exp(a.x), std::exp(a.y), std::exp(a.z)); }
template <class T> inline tvec3<T> clamp(const tvec3<T>& a, T lo, T hi) { return tvec3<T>(clamp(a.x, lo, hi), clamp(a.y, lo, hi), clamp(a.z, lo, hi)); }

// ---- matrices: C columns of R-component vectors, column-major like glm (m[col][row]) ----
template <int R> struct colvec;
template <> struct colvec<2> { typedef dvec2 type; };
template <> struct colvec<3> { typedef dvec3 type; };
template <int C, int R> struct dmat {
    typedef typename colvec<R>::type col_type;
    col_type c[C];
    dmat() {}                                                      // zero (columns zero-initialise)
    explicit dmat(double d) { for (int i = 0; i < C && i < R; i++) c[i][i] = d; }
    dmat(const col_type& a, const col_type& b) { static_assert(C == 2, "2 columns"); c[0] = a; c[1] = b; }
    dmat(const col_type& a, const col_type& b, const col_type& d) { static_assert(C == 3, "3 columns"); c[0] = a; c[1] = b; c[2] = d; }
    dmat(double a, double b, double d, double e) { static_assert(C == 2 && R == 2, "mat2"); c[0] = col_type(a, b); c[1] = col_type(d, e); }
    col_type& operator[](int i) { return c[i]; }
    const col_type& operator[](int i) const { return c[i]; }
};
typedef dmat<2, 2> dmat2; typedef dmat<3, 3> dmat3; typedef dmat<2, 3> dmat2x3; typedef dmat<3, 2> dmat3x2;
template <int C, int R> inline dmat<R, C> transpose(const dmat<C, R>& m) { dmat<R, C> t; for (int i = 0; i < C; i++) for (int j = 0; j < R; j++) t[j][i] = m[i][j]; return t; }
template <int C, int R> inline typename colvec<R>::type operator*(const dmat<C, R>& m, const typename colvec<C>::type& v) {
    typename colvec<R>::type r;
    for (int j = 0; j < R; j++) { double s = m[0][j] * v[0]; for (int i = 1; i < C; i++) s += m[i][j] * v[i]; r[j] = s; }
    return r;
}
template <int C, int R> inline typename colvec<C>::type operator*(const typename colvec<R>::type& v, const dmat<C, R>& m) {
    typename colvec<C>::type r;
    for (int i = 0; i < C; i++) r[i] = dot(v, m[i]);
    return r;
}
template <int K, int R, int C> inline dmat<C, R> operator*(const dmat<K, R>& a, const dmat<C, K>& b) {   // (R x K) * (K x C)
    dmat<C, R> r;
    for (int i = 0; i < C; i++) for (int j = 0; j < R; j++) { double s = a[0][j] * b[i][0]; for (int k = 1; k < K; k++) s += a[k][j] * b[i][k]; r[i][j] = s; }
    return r;
}
template <int C, int R> inline dmat<C, R> operator*(const dmat<C, R>& m, double s) { dmat<C, R> r; for (int i = 0; i < C; i++) r[i] = m[i] * s; return r; }
template <int C, int R> inline dmat<C, R> operator*(double s, const dmat<C, R>& m) { return m * s; }
template <int C, int R> inline dmat<C, R> operator/(const dmat<C, R>& m, double s) { dmat<C, R> r; for (int i = 0; i < C; i++) r[i] = m[i] / s; return r; }
template <int C, int R> inline dmat<C, R> operator+(const dmat<C, R>& a, const dmat<C, R>& b) { dmat<C, R> r; for (int i = 0; i < C; i++) r[i] = a[i] + b[i]; return r; }
template <int C, int R> inline dmat<C, R> operator-(const dmat<C, R>& a, const dmat<C, R>& b) { dmat<C, R> r; for (int i = 0; i < C; i++) r[i] = a[i] - b[i]; return r; }
template <int C, int R> inline dmat<C, R> operator-(const dmat<C, R>& a) { dmat<C, R> r; for (int i = 0; i < C; i++) r[i] = -a[i]; return r; }
inline double determinant(const dmat2& m) { return m[0][0] * m[1][1] - m[1][0] * m[0][1]; }
inline dmat2 inverse(const dmat2& m) {
    const double inv = 1.0 / determinant(m);
    return dmat2(m[1][1] * inv, -m[0][1] * inv, -m[1][0] * inv, m[0][0] * inv);
}
inline double determinant(const dmat3& m) {
    return m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) + m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]);
}

}  // namespace glm
