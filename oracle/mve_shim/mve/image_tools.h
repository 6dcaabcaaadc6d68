/*
 * Shim of MVE mve/image_tools.h (subset used by stereo_view.cc, sgm_stereo.cc,
 * depth_optimizer.cc). Semantics restated from MVE's documented behaviour
 * (separable Gaussian with radius ceil(2.884 sigma) and clamped borders,
 * 2x2 box half-size, luminance desaturation). TEST INFRASTRUCTURE ONLY.
 */
#ifndef SHIM_MVE_IMAGE_TOOLS_HEADER
#define SHIM_MVE_IMAGE_TOOLS_HEADER

#include <cmath>
#include <stdexcept>
#include <vector>

#include "math/accum.h"
#include "math/functions.h"
#include "mve/image.h"

MVE_NAMESPACE_BEGIN
MVE_IMAGE_NAMESPACE_BEGIN

enum DesaturateType
{
    DESATURATE_MAXIMUM,
    DESATURATE_LIGHTNESS,
    DESATURATE_LUMINOSITY,
    DESATURATE_LUMINANCE,
    DESATURATE_AVERAGE
};

inline FloatImage::Ptr
byte_to_float_image (ByteImage::ConstPtr image)
{
    if (image == nullptr)
        throw std::invalid_argument("Null image given");
    FloatImage::Ptr img = FloatImage::create(image->width(),
        image->height(), image->channels());
    for (int64_t i = 0; i < image->get_value_amount(); ++i)
    {
        float value = static_cast<float>(image->at(i)) / 255.0f;
        img->at(i) = std::min(1.0f, std::max(0.0f, value));
    }
    return img;
}

inline ByteImage::Ptr
float_to_byte_image (FloatImage::ConstPtr image, float vmin = 0.0f,
    float vmax = 1.0f)
{
    ByteImage::Ptr img = ByteImage::create(image->width(),
        image->height(), image->channels());
    for (int64_t i = 0; i < image->get_value_amount(); ++i)
    {
        float value = std::min(vmax, std::max(vmin, image->at(i)));
        value = 255.0f * (value - vmin) / (vmax - vmin);
        img->at(i) = static_cast<uint8_t>(value + 0.5f);
    }
    return img;
}

template <typename T>
inline typename Image<T>::Ptr
desaturate (typename Image<T>::ConstPtr img, DesaturateType type)
{
    if (img == nullptr)
        throw std::invalid_argument("Null image given");
    int64_t const ic = img->channels();
    if (ic != 3 && ic != 4)
        throw std::invalid_argument("Image must be RGB or RGBA");
    if (type != DESATURATE_LUMINANCE)
        throw std::invalid_argument("shim: only DESATURATE_LUMINANCE");
    bool const has_alpha = (ic == 4);
    typename Image<T>::Ptr out = Image<T>::create(img->width(),
        img->height(), 1 + has_alpha);
    int64_t outpos = 0, inpos = 0;
    for (int64_t i = 0; i < img->get_pixel_amount(); ++i)
    {
        T const* v = &img->at(inpos);
        out->at(outpos) = math::interpolate<T>(v[0], v[1], v[2],
            0.21f, 0.72f, 0.07f);
        if (has_alpha)
            out->at(outpos + 1) = img->at(inpos + 3);
        outpos += 1 + has_alpha;
        inpos += ic;
    }
    return out;
}

template <typename T>
inline typename Image<T>::Ptr
blur_gaussian (typename Image<T>::ConstPtr in, float sigma)
{
    if (in == nullptr)
        throw std::invalid_argument("Null image given");
    if (MATH_EPSILON_EQ(sigma, 0.0f, 0.1f))
        return in->duplicate();

    int64_t const w = in->width(), h = in->height(), c = in->channels();
    int const ks = static_cast<int>(std::ceil(sigma * 2.884f));
    std::vector<float> kernel(ks + 1);
    for (int i = 0; i < ks + 1; ++i)
        kernel[i] = math::gaussian((float)i, sigma);

    typename Image<T>::Ptr sep(Image<T>::create(w, h, c));
    int64_t px = 0;
    for (int64_t y = 0; y < h; ++y)
        for (int64_t x = 0; x < w; ++x, ++px)
            for (int64_t cc = 0; cc < c; ++cc)
            {
                math::Accum<T> accum(T(0));
                for (int i = -ks; i <= ks; ++i)
                {
                    int64_t idx = math::clamp<int64_t>(x + i, 0, w - 1);
                    accum.add(in->at(y * w + idx, cc), kernel[std::abs(i)]);
                }
                sep->at(px, cc) = accum.normalized();
            }

    typename Image<T>::Ptr out(Image<T>::create(w, h, c));
    px = 0;
    for (int64_t y = 0; y < h; ++y)
        for (int64_t x = 0; x < w; ++x, ++px)
            for (int64_t cc = 0; cc < c; ++cc)
            {
                math::Accum<T> accum(T(0));
                for (int i = -ks; i <= ks; ++i)
                {
                    int64_t idx = math::clamp<int64_t>(y + i, 0, h - 1);
                    accum.add(sep->at(idx * w + x, cc), kernel[std::abs(i)]);
                }
                out->at(px, cc) = accum.normalized();
            }
    return out;
}

template <typename T>
inline typename Image<T>::Ptr
rescale_half_size (typename Image<T>::ConstPtr img)
{
    if (img == nullptr)
        throw std::invalid_argument("Null image given");
    int64_t const iw = img->width(), ih = img->height(), ic = img->channels();
    int64_t const ow = (iw + 1) >> 1, oh = (ih + 1) >> 1;
    if (iw < 2 || ih < 2)
        throw std::invalid_argument("Input image too small for half-sizing");
    typename Image<T>::Ptr out(Image<T>::create(ow, oh, ic));
    int64_t outpos = 0;
    int64_t const rowstride = iw * ic;
    for (int64_t y = 0; y < oh; ++y)
    {
        int64_t irow1 = y * 2 * rowstride;
        int64_t irow2 = irow1 + rowstride * (y * 2 + 1 < ih);
        for (int64_t x = 0; x < ow; ++x)
        {
            int64_t ipix1 = irow1 + x * 2 * ic;
            int64_t ipix2 = irow2 + x * 2 * ic;
            int64_t hasnext = (x * 2 + 1 < iw);
            for (int64_t cc = 0; cc < ic; ++cc)
                out->at(outpos++) = math::interpolate<T>(
                    img->at(ipix1 + cc), img->at(ipix1 + ic * hasnext + cc),
                    img->at(ipix2 + cc), img->at(ipix2 + ic * hasnext + cc),
                    0.25f, 0.25f, 0.25f, 0.25f);
        }
    }
    return out;
}

template <typename T>
inline void
gamma_correct_inv_srgb (typename Image<T>::Ptr image)
{
    for (T* p = image->begin(); p != image->end(); ++p)
    {
        T const v = *p;
        *p = (v <= T(0.04045) ? v / T(12.92)
            : std::pow((v + T(0.055)) / T(1.055), T(2.4)));
    }
}

MVE_IMAGE_NAMESPACE_END
MVE_NAMESPACE_END

#endif
