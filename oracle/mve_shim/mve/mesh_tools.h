/* Shim of MVE mve/mesh_tools.h (declarations only, see mve/mesh.h). */
#ifndef SHIM_MVE_MESH_TOOLS_HEADER
#define SHIM_MVE_MESH_TOOLS_HEADER

#include "mve/mesh.h"

MVE_NAMESPACE_BEGIN
MVE_GEOM_NAMESPACE_BEGIN

void mesh_merge (TriangleMesh::ConstPtr mesh1, TriangleMesh::Ptr mesh2);
void depthmap_mesh_confidences (TriangleMesh::Ptr mesh, int iterations = 3);

MVE_GEOM_NAMESPACE_END
MVE_NAMESPACE_END

#endif
