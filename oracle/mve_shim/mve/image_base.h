/* Shim of MVE mve/image_base.h. TEST INFRASTRUCTURE ONLY (oracle build). */
#ifndef SHIM_MVE_IMAGE_BASE_HEADER
#define SHIM_MVE_IMAGE_BASE_HEADER

#include <cstdint>
#include <memory>
#include <vector>

#include "mve/defines.h"

MVE_NAMESPACE_BEGIN

class ImageBase
{
public:
    typedef std::shared_ptr<ImageBase> Ptr;
    typedef std::shared_ptr<ImageBase const> ConstPtr;

    ImageBase (void) : w(0), h(0), c(0) {}
    virtual ~ImageBase (void) {}

    int64_t width (void) const { return w; }
    int64_t height (void) const { return h; }
    int64_t channels (void) const { return c; }
    bool valid (void) const { return w && h && c; }

protected:
    int64_t w, h, c;
};

MVE_NAMESPACE_END

#endif
