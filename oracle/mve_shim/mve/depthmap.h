/* Shim of MVE mve/depthmap.h (depth convention change). TEST INFRA ONLY. */
#ifndef SHIM_MVE_DEPTHMAP_HEADER
#define SHIM_MVE_DEPTHMAP_HEADER

#include "math/matrix.h"
#include "math/vector.h"
#include "mve/image.h"

MVE_NAMESPACE_BEGIN
MVE_IMAGE_NAMESPACE_BEGIN

/* MVE depth = distance along the viewing ray; SMVS depth = z. */
template <typename T>
inline void
depthmap_convert_conventions (typename Image<T>::Ptr dm,
    math::Matrix3f const& invproj, bool to_mve)
{
    std::size_t const width = dm->width();
    std::size_t const height = dm->height();
    std::size_t pos = 0;
    for (std::size_t y = 0; y < height; ++y)
        for (std::size_t x = 0; x < width; ++x, ++pos)
        {
            math::Vec3f px((float)x + 0.5f, (float)y + 0.5f, 1.0f);
            px = invproj * px;
            double len = px.norm();
            dm->at(pos) *= (to_mve ? len : 1.0 / len);
        }
}

MVE_IMAGE_NAMESPACE_END

MVE_GEOM_NAMESPACE_BEGIN

/* 3-D position of pixel (x, y) at MVE depth (distance along the viewing ray)
 * in camera coordinates (MVE, restated). */
inline math::Vec3f
pixel_3dpos (std::size_t x, std::size_t y, float depth,
    math::Matrix3f const& invproj)
{
    math::Vec3f ray = invproj * math::Vec3f((float)x + 0.5f,
        (float)y + 0.5f, 1.0f);
    return ray.normalized() * depth;
}

MVE_GEOM_NAMESPACE_END
MVE_NAMESPACE_END

#endif
