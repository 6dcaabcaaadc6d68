/*
 * Shim of MVE math/matrix_svd.h: SVD + pseudo inverse. MVE uses a
 * Golub-Kahan SVD; this restatement uses one-sided Jacobi (same result up
 * to rounding). Singular values within epsilon of zero are dropped, as
 * matrix_pseudo_inverse does in MVE. TEST INFRASTRUCTURE ONLY.
 */
#ifndef SHIM_MATH_MATRIX_SVD_HEADER
#define SHIM_MATH_MATRIX_SVD_HEADER

#include <cmath>
#include <vector>

#include "math/matrix.h"
#include "math/matrix_tools.h"

MATH_NAMESPACE_BEGIN

/* A (rows x cols, rows >= cols) = U * diag(S) * V^T, raw row-major arrays. */
template <typename T>
inline void
matrix_svd (T const* mat_a, int rows, int cols,
    T* mat_u, T* vec_s, T* mat_v, T const& epsilon = T(1e-12))
{
    (void)epsilon;
    std::vector<T> U(mat_a, mat_a + rows * cols);
    std::vector<T> V(cols * cols, T(0));
    for (int i = 0; i < cols; ++i)
        V[i * cols + i] = T(1);

    for (int sweep = 0; sweep < 60; ++sweep)
    {
        T off(0);
        for (int p = 0; p < cols - 1; ++p)
            for (int q = p + 1; q < cols; ++q)
            {
                T alpha(0), beta(0), gamma(0);
                for (int i = 0; i < rows; ++i)
                {
                    alpha += U[i * cols + p] * U[i * cols + p];
                    beta += U[i * cols + q] * U[i * cols + q];
                    gamma += U[i * cols + p] * U[i * cols + q];
                }
                if (gamma == T(0))
                    continue;
                T const lim = std::sqrt(alpha * beta);
                if (std::abs(gamma) <= T(1e-16) * lim)
                    continue;
                off = std::max(off, std::abs(gamma) / (lim > T(0) ? lim : T(1)));
                T const zeta = (beta - alpha) / (T(2) * gamma);
                T const t = (zeta >= T(0) ? T(1) : T(-1))
                    / (std::abs(zeta) + std::sqrt(T(1) + zeta * zeta));
                T const c = T(1) / std::sqrt(T(1) + t * t);
                T const s = c * t;
                for (int i = 0; i < rows; ++i)
                {
                    T const up = U[i * cols + p], uq = U[i * cols + q];
                    U[i * cols + p] = c * up - s * uq;
                    U[i * cols + q] = s * up + c * uq;
                }
                for (int i = 0; i < cols; ++i)
                {
                    T const vp = V[i * cols + p], vq = V[i * cols + q];
                    V[i * cols + p] = c * vp - s * vq;
                    V[i * cols + q] = s * vp + c * vq;
                }
            }
        if (off < T(1e-15))
            break;
    }

    for (int j = 0; j < cols; ++j)
    {
        T n(0);
        for (int i = 0; i < rows; ++i)
            n += U[i * cols + j] * U[i * cols + j];
        n = std::sqrt(n);
        vec_s[j] = n;
        for (int i = 0; i < rows; ++i)
            mat_u[i * cols + j] = (n > T(0) ? U[i * cols + j] / n : T(0));
    }
    std::copy(V.begin(), V.end(), mat_v);
}

template <typename T, int M, int N>
inline void
matrix_svd (Matrix<T,M,N> const& mat_a, Matrix<T,M,N>* mat_u,
    Matrix<T,N,N>* mat_s, Matrix<T,N,N>* mat_v, T const& epsilon = T(1e-12))
{
    T s[N];
    matrix_svd<T>(mat_a.begin(), M, N, mat_u->begin(), s, mat_v->begin(),
        epsilon);
    mat_s->fill(T(0));
    for (int i = 0; i < N; ++i)
        (*mat_s)(i, i) = s[i];
}

template <typename T, int M, int N>
inline void
matrix_pseudo_inverse (Matrix<T,M,N> const& A, Matrix<T,N,M>* result,
    T const& epsilon = T(1e-12))
{
    Matrix<T,M,N> U;
    Matrix<T,N,N> S;
    Matrix<T,N,N> V;
    matrix_svd(A, &U, &S, &V, epsilon);
    for (int i = 0; i < N; ++i)
    {
        if (MATH_EPSILON_EQ(S(i, i), T(0), epsilon))
            S(i, i) = T(0);
        else
            S(i, i) = T(1) / S(i, i);
    }
    *result = V * S * U.transposed();
}

MATH_NAMESPACE_END

#endif
