/* Shim of MVE math/matrix_tools.h (raw-array helpers). TEST INFRA ONLY. */
#ifndef SHIM_MATH_MATRIX_TOOLS_HEADER
#define SHIM_MATH_MATRIX_TOOLS_HEADER

#include <algorithm>
#include <vector>

#include "math/matrix.h"

MATH_NAMESPACE_BEGIN

/* In-place transpose of a rows x cols row-major array. */
template <typename T>
inline void
matrix_transpose (T* mat, int rows, int cols)
{
    std::vector<T> tmp(mat, mat + rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c)
            mat[c * rows + r] = tmp[r * cols + c];
}

/* R = A (rows_a x cols_a) * B (cols_a x cols_b). */
template <typename T>
inline void
matrix_multiply (T const* mat_a, int rows_a, int cols_a,
    T const* mat_b, int cols_b, T* mat_res)
{
    std::fill(mat_res, mat_res + rows_a * cols_b, T(0));
    for (int j = 0; j < cols_b; ++j)
        for (int i = 0; i < rows_a; ++i)
            for (int k = 0; k < cols_a; ++k)
                mat_res[i * cols_b + j] +=
                    mat_a[i * cols_a + k] * mat_b[k * cols_b + j];
}

MATH_NAMESPACE_END

#endif
