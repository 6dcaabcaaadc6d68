/* Shim of MVE mve/mesh_info.h (declarations only, see mve/mesh.h). */
#ifndef SHIM_MVE_MESH_INFO_HEADER
#define SHIM_MVE_MESH_INFO_HEADER

#include <vector>

#include "mve/mesh.h"

MVE_NAMESPACE_BEGIN

class MeshInfo
{
public:
    struct VertexInfo
    {
        int vclass;
        std::vector<std::size_t> verts;
        std::vector<std::size_t> faces;
    };

    MeshInfo (void) {}
    MeshInfo (TriangleMesh::ConstPtr mesh);      /* never defined / called */
    std::size_t size (void) const { return info.size(); }
    VertexInfo& operator[] (std::size_t i) { return info[i]; }
    VertexInfo const& operator[] (std::size_t i) const { return info[i]; }
    VertexInfo& at (std::size_t i) { return info[i]; }

private:
    std::vector<VertexInfo> info;
};

MVE_NAMESPACE_END

#endif
