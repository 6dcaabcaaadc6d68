/*
 * Minimal restatement of the part of MVE's math/defines.h that the SMVS hot
 * path uses. TEST INFRASTRUCTURE ONLY (oracle build). Written from the
 * documented behaviour of simonfuhrmann/mve; MVE itself is not in this tree.
 */
#ifndef SHIM_MATH_DEFINES_HEADER
#define SHIM_MATH_DEFINES_HEADER

#define MATH_NAMESPACE_BEGIN namespace math {
#define MATH_NAMESPACE_END }

#define MATH_PI         3.14159265358979323846264338327950288
#define MATH_SQRT2      1.41421356237309504880168872420969808
#define MATH_SQRT3      1.7320508075688772935274463415058723669
#define MATH_SQRT_2PI   2.506628274631000502415765284811045253

#define MATH_POW2(x) ((x) * (x))
#define MATH_POW3(x) (MATH_POW2(x) * (x))
#define MATH_POW4(x) (MATH_POW2(MATH_POW2(x)))

#define MATH_FLT_EPS (1e-5f)
#define MATH_DBL_EPS (1e-14)
#define MATH_EPSILON_EQ(x,v,eps) (((v - eps) <= x) && (x <= (v + eps)))
#define MATH_EPSILON_LESS(x,v,eps) ((x + eps) < v)
#define MATH_FLOAT_EQ(x,v) MATH_EPSILON_EQ(x,v,MATH_FLT_EPS)
#define MATH_DOUBLE_EQ(x,v) MATH_EPSILON_EQ(x,v,MATH_DBL_EPS)

#endif
