/* Shim of MVE mve/mesh_io.h: nothing of it is used by the oracle. */
#ifndef SHIM_MVE_MESH_IO_HEADER
#define SHIM_MVE_MESH_IO_HEADER
#include "mve/mesh.h"
#endif
