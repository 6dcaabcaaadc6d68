#ifndef SHIM_MVE_DEFINES_HEADER
#define SHIM_MVE_DEFINES_HEADER
#define MVE_NAMESPACE_BEGIN namespace mve {
#define MVE_NAMESPACE_END }
#define MVE_IMAGE_NAMESPACE_BEGIN namespace image {
#define MVE_IMAGE_NAMESPACE_END }
#define MVE_GEOM_NAMESPACE_BEGIN namespace geom {
#define MVE_GEOM_NAMESPACE_END }
#endif
