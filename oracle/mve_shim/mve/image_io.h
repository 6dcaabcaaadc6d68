/* Shim: the hot path never touches image files. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_MVE_IMAGE_IO_HEADER
#define SHIM_MVE_IMAGE_IO_HEADER
#include "mve/image.h"
#endif
