/*
 * Shim of MVE math/vector.h: fixed-size dense vector with the members the
 * SMVS hot-path sources use. TEST INFRASTRUCTURE ONLY (oracle build).
 */
#ifndef SHIM_MATH_VECTOR_HEADER
#define SHIM_MATH_VECTOR_HEADER

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <numeric>

#include "math/defines.h"

MATH_NAMESPACE_BEGIN

template <typename T, int N>
class Vector
{
public:
    static int constexpr dim = N;

    Vector (void) {}
    explicit Vector (T const* values) { std::copy(values, values + N, v); }
    explicit Vector (T const& value) { std::fill(v, v + N, value); }
    Vector (T const& v1, T const& v2) { v[0] = v1; v[1] = v2; }
    Vector (T const& v1, T const& v2, T const& v3)
    { v[0] = v1; v[1] = v2; v[2] = v3; }
    Vector (T const& v1, T const& v2, T const& v3, T const& v4)
    { v[0] = v1; v[1] = v2; v[2] = v3; v[3] = v4; }
    template <typename O>
    Vector (Vector<O,N> const& other)
    { for (int i = 0; i < N; ++i) v[i] = static_cast<T>(other[i]); }

    Vector& fill (T const& value) { std::fill(v, v + N, value); return *this; }

    T* begin (void) { return v; }
    T const* begin (void) const { return v; }
    T* end (void) { return v + N; }
    T const* end (void) const { return v + N; }
    T* operator* (void) { return v; }
    T const* operator* (void) const { return v; }
    T& operator[] (int index) { return v[index]; }
    T const& operator[] (int index) const { return v[index]; }
    T& operator() (int index) { return v[index]; }
    T const& operator() (int index) const { return v[index]; }

    T dot (Vector const& o) const
    { return std::inner_product(v, v + N, o.v, T(0)); }
    T square_norm (void) const { return this->dot(*this); }
    T norm (void) const { return std::sqrt(this->square_norm()); }
    T abs_sum (void) const
    { T r(0); for (int i = 0; i < N; ++i) r += std::abs(v[i]); return r; }
    T sum (void) const { return std::accumulate(v, v + N, T(0)); }
    T minimum (void) const { return *std::min_element(v, v + N); }
    T maximum (void) const { return *std::max_element(v, v + N); }
    Vector& normalize (void)
    { T const n = this->norm(); for (int i = 0; i < N; ++i) v[i] /= n;
      return *this; }
    Vector normalized (void) const { return Vector(*this).normalize(); }
    Vector cross (Vector const& o) const
    {
        return Vector(v[1] * o.v[2] - v[2] * o.v[1],
            v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }

    Vector operator- (void) const
    { Vector r; for (int i = 0; i < N; ++i) r.v[i] = -v[i]; return r; }
    Vector& operator+= (Vector const& o)
    { for (int i = 0; i < N; ++i) v[i] += o.v[i]; return *this; }
    Vector& operator-= (Vector const& o)
    { for (int i = 0; i < N; ++i) v[i] -= o.v[i]; return *this; }
    Vector& operator+= (T const& s)
    { for (int i = 0; i < N; ++i) v[i] += s; return *this; }
    Vector& operator-= (T const& s)
    { for (int i = 0; i < N; ++i) v[i] -= s; return *this; }
    Vector& operator*= (T const& s)
    { for (int i = 0; i < N; ++i) v[i] *= s; return *this; }
    Vector& operator/= (T const& s)
    { for (int i = 0; i < N; ++i) v[i] /= s; return *this; }
    Vector operator+ (Vector const& o) const { return Vector(*this) += o; }
    Vector operator- (Vector const& o) const { return Vector(*this) -= o; }
    Vector operator+ (T const& s) const { return Vector(*this) += s; }
    Vector operator- (T const& s) const { return Vector(*this) -= s; }
    Vector operator* (T const& s) const { return Vector(*this) *= s; }
    Vector operator/ (T const& s) const { return Vector(*this) /= s; }
    bool operator== (Vector const& o) const
    { return std::equal(v, v + N, o.v); }
    bool operator!= (Vector const& o) const { return !(*this == o); }

protected:
    T v[N];
};

template <typename T, int N>
inline Vector<T,N>
operator* (T const& s, Vector<T,N> const& v)
{
    return v * s;
}

typedef Vector<float,2> Vec2f;
typedef Vector<float,3> Vec3f;
typedef Vector<float,4> Vec4f;
typedef Vector<double,2> Vec2d;
typedef Vector<double,3> Vec3d;
typedef Vector<double,4> Vec4d;
typedef Vector<int,2> Vec2i;
typedef Vector<int,3> Vec3i;
typedef Vector<unsigned int,3> Vec3ui;

MATH_NAMESPACE_END

#endif
