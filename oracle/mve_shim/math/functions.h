/* Shim of MVE math/functions.h (subset). TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_MATH_FUNCTIONS_HEADER
#define SHIM_MATH_FUNCTIONS_HEADER

#include <cmath>
#include <cstdint>
#include "math/defines.h"

MATH_NAMESPACE_BEGIN

template <typename T>
inline T const&
clamp (T const& v, T const& min = T(0), T const& max = T(1))
{
    return (v < min ? min : (v > max ? max : v));
}

template <typename T>
inline T
gaussian (T const& x, T const& sigma)
{
    return std::exp(-((x * x) / (T(2) * sigma * sigma)));
}

template <typename T>
inline T
gaussian_xx (T const& xx, T const& sigma)
{
    return std::exp(-(xx / (T(2) * sigma * sigma)));
}

template <typename T>
inline T
gaussian_2d (T const& x, T const& y, T const& sigma_x, T const& sigma_y)
{
    return std::exp(-(x * x) / (T(2) * sigma_x * sigma_x)
        - (y * y) / (T(2) * sigma_y * sigma_y));
}

/* Weighted sums; the byte specialisation rounds to nearest. */
template <typename T>
inline T
interpolate (T const& v1, T const& v2, T const& v3,
    float w1, float w2, float w3)
{
    return v1 * w1 + v2 * w2 + v3 * w3;
}

template <typename T>
inline T
interpolate (T const& v1, T const& v2, T const& v3, T const& v4,
    float w1, float w2, float w3, float w4)
{
    return v1 * w1 + v2 * w2 + v3 * w3 + v4 * w4;
}

template <>
inline unsigned char
interpolate (unsigned char const& v1, unsigned char const& v2,
    unsigned char const& v3, float w1, float w2, float w3)
{
    return (unsigned char)((float)v1 * w1 + (float)v2 * w2
        + (float)v3 * w3 + 0.5f);
}

template <>
inline unsigned char
interpolate (unsigned char const& v1, unsigned char const& v2,
    unsigned char const& v3, unsigned char const& v4,
    float w1, float w2, float w3, float w4)
{
    return (unsigned char)((float)v1 * w1 + (float)v2 * w2
        + (float)v3 * w3 + (float)v4 * w4 + 0.5f);
}

MATH_NAMESPACE_END

#endif
