/* Shim of MVE mve/scene.h: only the view list type lib/mesh_generator.h
 * names. TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_MVE_SCENE_HEADER
#define SHIM_MVE_SCENE_HEADER

#include <memory>
#include <vector>

#include "mve/view.h"
#include "mve/mesh.h"

MVE_NAMESPACE_BEGIN

class Scene
{
public:
    typedef std::shared_ptr<Scene> Ptr;
    typedef std::vector<View::Ptr> ViewList;
};

MVE_NAMESPACE_END

#endif
