/* Shim of MVE mve/mesh.h: the container interface lib/mesh_generator.cc,
 * lib/depth_triangulator.h and lib/mesh_simplifier.h name. Only
 * MeshGenerator::cut_depth_maps of that translation unit is ever CALLED by the
 * oracle; the meshing functions compile against these declarations and are
 * never linked to an implementation. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_MVE_MESH_HEADER
#define SHIM_MVE_MESH_HEADER

#include <memory>
#include <vector>

#include "math/vector.h"
#include "mve/defines.h"

MVE_NAMESPACE_BEGIN

class TriangleMesh
{
public:
    typedef std::shared_ptr<TriangleMesh> Ptr;
    typedef std::shared_ptr<TriangleMesh const> ConstPtr;
    typedef unsigned int VertexID;
    typedef std::vector<math::Vec3f> VertexList;
    typedef std::vector<math::Vec3f> NormalList;
    typedef std::vector<math::Vec4f> ColorList;
    typedef std::vector<float> ValueList;
    typedef std::vector<float> ConfidenceList;
    typedef std::vector<VertexID> FaceList;

    static Ptr create (void) { return Ptr(new TriangleMesh()); }
    Ptr duplicate (void) const { return Ptr(new TriangleMesh(*this)); }

    VertexList& get_vertices (void) { return vertices; }
    VertexList const& get_vertices (void) const { return vertices; }
    NormalList& get_vertex_normals (void) { return normals; }
    NormalList const& get_vertex_normals (void) const { return normals; }
    ColorList& get_vertex_colors (void) { return colors; }
    ColorList const& get_vertex_colors (void) const { return colors; }
    ValueList& get_vertex_values (void) { return values; }
    ValueList const& get_vertex_values (void) const { return values; }
    ConfidenceList& get_vertex_confidences (void) { return confidences; }
    ConfidenceList const& get_vertex_confidences (void) const
    { return confidences; }
    FaceList& get_faces (void) { return faces; }
    FaceList const& get_faces (void) const { return faces; }
    void recalc_normals (bool = true, bool = true) {}
    void delete_vertices_fix_faces (std::vector<bool> const&) {}

private:
    VertexList vertices;
    NormalList normals;
    ColorList colors;
    ValueList values;
    ConfidenceList confidences;
    FaceList faces;
};

MVE_NAMESPACE_END

#endif
