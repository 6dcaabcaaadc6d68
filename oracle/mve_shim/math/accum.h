/* Shim of MVE math/accum.h. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_MATH_ACCUM_HEADER
#define SHIM_MATH_ACCUM_HEADER

#include "math/defines.h"

MATH_NAMESPACE_BEGIN

template <typename T>
class Accum
{
public:
    T v;
    float w;

public:
    Accum (void) : w(0.0f) {}
    Accum (T const& init) : v(init), w(0.0f) {}
    void add (T const& value, float weight)
    {
        this->v += value * weight;
        this->w += weight;
    }
    void sub (T const& value, float weight)
    {
        this->v -= value * weight;
        this->w -= weight;
    }
    T normalized (float weight) const { return this->v / weight; }
    T normalized (void) const { return this->v / this->w; }
};

MATH_NAMESPACE_END

#endif
