/*
 * Shim of MVE math/matrix.h: fixed-size row-major matrix (subset).
 * Products accumulate left to right starting from T(0), as
 * std::inner_product does in MVE. TEST INFRASTRUCTURE ONLY (oracle build).
 */
#ifndef SHIM_MATH_MATRIX_HEADER
#define SHIM_MATH_MATRIX_HEADER

#include <algorithm>

#include "math/defines.h"
#include "math/vector.h"

MATH_NAMESPACE_BEGIN

template <typename T, int N, int M>
class Matrix
{
public:
    static int constexpr rows = N;
    static int constexpr cols = M;

    Matrix (void) {}
    explicit Matrix (T const* values) { std::copy(values, values + N * M, m); }
    explicit Matrix (T const& value) { std::fill(m, m + N * M, value); }
    template <typename O>
    Matrix (Matrix<O,N,M> const& other)
    { for (int i = 0; i < N * M; ++i) m[i] = static_cast<T>(other[i]); }

    Matrix& fill (T const& value)
    { std::fill(m, m + N * M, value); return *this; }

    T* begin (void) { return m; }
    T const* begin (void) const { return m; }
    T* end (void) { return m + N * M; }
    T const* end (void) const { return m + N * M; }
    T* operator* (void) { return m; }
    T const* operator* (void) const { return m; }
    T& operator() (int row, int col) { return m[row * M + col]; }
    T const& operator() (int row, int col) const { return m[row * M + col]; }
    T& operator[] (unsigned int i) { return m[i]; }
    T const& operator[] (unsigned int i) const { return m[i]; }

    Vector<T,M> row (int index) const { return Vector<T,M>(m + index * M); }
    Vector<T,N> col (int index) const
    { Vector<T,N> r; for (int i = 0; i < N; ++i) r[i] = m[i * M + index];
      return r; }

    Matrix<T,M,N> transposed (void) const
    {
        Matrix<T,M,N> r;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < M; ++j)
                r(j, i) = (*this)(i, j);
        return r;
    }

    template <int U>
    Matrix<T,N,U> mult (Matrix<T,M,U> const& rhs) const
    {
        Matrix<T,N,U> r;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < U; ++j)
            {
                T sum(0);
                for (int k = 0; k < M; ++k)
                    sum += m[i * M + k] * rhs[k * U + j];
                r(i, j) = sum;
            }
        return r;
    }

    Vector<T,N> mult (Vector<T,M> const& rhs) const
    {
        Vector<T,N> r;
        for (int i = 0; i < N; ++i)
        {
            T sum(0);
            for (int k = 0; k < M; ++k)
                sum += m[i * M + k] * rhs[k];
            r[i] = sum;
        }
        return r;
    }

    /* homogeneous product: (M-1)-vector extended by v, first N-1 rows
     * (MVE, restated) */
    Vector<T,N-1> mult (Vector<T,M-1> const& rhs, T const& v) const
    {
        Vector<T,N-1> r;
        for (int i = 0; i < N - 1; ++i)
        {
            T sum(0);
            for (int k = 0; k < M - 1; ++k)
                sum += m[i * M + k] * rhs[k];
            r[i] = sum + v * m[i * M + M - 1];
        }
        return r;
    }

    template <int U>
    Matrix<T,N,U> operator* (Matrix<T,M,U> const& rhs) const
    { return this->mult(rhs); }
    Vector<T,N> operator* (Vector<T,M> const& rhs) const
    { return this->mult(rhs); }

    Matrix& operator+= (Matrix const& o)
    { for (int i = 0; i < N * M; ++i) m[i] += o.m[i]; return *this; }
    Matrix& operator-= (Matrix const& o)
    { for (int i = 0; i < N * M; ++i) m[i] -= o.m[i]; return *this; }
    Matrix& operator*= (T const& s)
    { for (int i = 0; i < N * M; ++i) m[i] *= s; return *this; }
    Matrix& operator/= (T const& s)
    { for (int i = 0; i < N * M; ++i) m[i] /= s; return *this; }
    Matrix operator+ (Matrix const& o) const { return Matrix(*this) += o; }
    Matrix operator- (Matrix const& o) const { return Matrix(*this) -= o; }
    Matrix operator* (T const& s) const { return Matrix(*this) *= s; }
    Matrix operator/ (T const& s) const { return Matrix(*this) /= s; }
    Matrix operator- (void) const
    { Matrix r; for (int i = 0; i < N * M; ++i) r.m[i] = -m[i]; return r; }

protected:
    T m[N * M];
};

typedef Matrix<float,2,2> Matrix2f;
typedef Matrix<float,3,3> Matrix3f;
typedef Matrix<float,4,4> Matrix4f;
typedef Matrix<double,2,2> Matrix2d;
typedef Matrix<double,3,3> Matrix3d;
typedef Matrix<double,4,4> Matrix4d;

MATH_NAMESPACE_END

#endif
