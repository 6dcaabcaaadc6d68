/*
 * Shim of MVE mve/image.h: interleaved W x H x C image,
 * data[(y * w + x) * c + channel]. linear_at() clamps the coordinates to the
 * image, uses fp32 weights, and rounds to nearest for byte images (through
 * math::interpolate). TEST INFRASTRUCTURE ONLY (oracle build).
 */
#ifndef SHIM_MVE_IMAGE_HEADER
#define SHIM_MVE_IMAGE_HEADER

#include <algorithm>
#include <cstdint>
#include <memory>
#include <vector>

#include "math/functions.h"
#include "math/vector.h"
#include "math/matrix.h"
#include "mve/defines.h"
#include "mve/image_base.h"

MVE_NAMESPACE_BEGIN

template <typename T>
class Image : public ImageBase
{
public:
    typedef std::shared_ptr<Image<T> > Ptr;
    typedef std::shared_ptr<Image<T> const> ConstPtr;
    typedef T ValueType;

    Image (void) {}
    Image (int64_t width, int64_t height, int64_t chans)
    { this->allocate(width, height, chans); }

    static Ptr create (void) { return Ptr(new Image<T>()); }
    static Ptr create (int64_t width, int64_t height, int64_t chans)
    { return Ptr(new Image<T>(width, height, chans)); }
    static Ptr create (Image<T> const& other)
    { return Ptr(new Image<T>(other)); }

    Ptr duplicate (void) const { return Ptr(new Image<T>(*this)); }

    void allocate (int64_t width, int64_t height, int64_t chans)
    {
        this->w = width; this->h = height; this->c = chans;
        this->data.assign(width * height * chans, T(0));
    }
    void clear (void) { this->w = this->h = this->c = 0; this->data.clear(); }
    void fill (T const& value)
    { std::fill(this->data.begin(), this->data.end(), value); }

    int64_t get_pixel_amount (void) const { return this->w * this->h; }
    int64_t get_value_amount (void) const
    { return static_cast<int64_t>(this->data.size()); }

    T* begin (void) { return this->data.data(); }
    T const* begin (void) const { return this->data.data(); }
    T* end (void) { return this->data.data() + this->data.size(); }
    T const* end (void) const { return this->data.data() + this->data.size(); }
    T* get_data_pointer (void) { return this->data.data(); }
    T const* get_data_pointer (void) const { return this->data.data(); }

    T const& at (int64_t index) const { return this->data[index]; }
    T& at (int64_t index) { return this->data[index]; }
    T const& at (int64_t index, int64_t channel) const
    { return this->data[index * this->c + channel]; }
    T& at (int64_t index, int64_t channel)
    { return this->data[index * this->c + channel]; }
    T const& at (int64_t x, int64_t y, int64_t channel) const
    { return this->data[(y * this->w + x) * this->c + channel]; }
    T& at (int64_t x, int64_t y, int64_t channel)
    { return this->data[(y * this->w + x) * this->c + channel]; }
    T const& operator[] (int64_t index) const { return this->data[index]; }
    T& operator[] (int64_t index) { return this->data[index]; }

    T linear_at (float x, float y, int64_t channel) const
    {
        x = std::max(0.0f, std::min(static_cast<float>(this->w - 1), x));
        y = std::max(0.0f, std::min(static_cast<float>(this->h - 1), y));

        int64_t const floor_x = static_cast<int64_t>(x);
        int64_t const floor_y = static_cast<int64_t>(y);
        int64_t const floor_xp1 = std::min(floor_x + 1, this->w - 1);
        int64_t const floor_yp1 = std::min(floor_y + 1, this->h - 1);

        float const w1 = x - static_cast<float>(floor_x);
        float const w0 = 1.0f - w1;
        float const w3 = y - static_cast<float>(floor_y);
        float const w2 = 1.0f - w3;

        int64_t const rowstride = this->w * this->c;
        int64_t const row1 = floor_y * rowstride;
        int64_t const row2 = floor_yp1 * rowstride;
        int64_t const col1 = floor_x * this->c;
        int64_t const col2 = floor_xp1 * this->c;

        return math::interpolate<T>(
            this->at(row1 + col1 + channel), this->at(row1 + col2 + channel),
            this->at(row2 + col1 + channel), this->at(row2 + col2 + channel),
            w0 * w2, w1 * w2, w0 * w3, w1 * w3);
    }

    void linear_at (float x, float y, T* px) const
    {
        for (int64_t cc = 0; cc < this->c; ++cc)
            px[cc] = this->linear_at(x, y, cc);
    }

protected:
    std::vector<T> data;
};

typedef Image<uint8_t> ByteImage;
typedef Image<uint16_t> RawImage;
typedef Image<float> FloatImage;
typedef Image<double> DoubleImage;
typedef Image<int> IntImage;

MVE_NAMESPACE_END

#endif
