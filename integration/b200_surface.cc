/*
 * integration/b200_surface.cc
 *
 * Drop-in bodies for smvs::Surface::get_depth_map and get_normal_map
 * (reference: lib/surface.cc:155-183) -- the only two doors through which a
 * caller of DepthOptimizer can reach the optimizer's private surface
 * (DepthOptimizer::get_depth / get_normals, inline in lib/depth_optimizer.h:
 * 136-148). After the resident optimize() the final surface exists on the
 * GPU, and so do its rendered depth and normal maps (bit-identical to the
 * reference's renderer, tests/test_gpu_visibility.py); building the host
 * Surface object for it -- 128 104 shared_ptr patches at 2 MP -- takes seven
 * times as long as the whole optimisation, and smvsrecon never looks at it
 * (app/smvsrecon.cc:722-724). So optimize() leaves a small stand-in Surface
 * and registers the two maps for it here; a Surface that is not registered
 * runs the reference's own renderer, kept as smvs_ref_surface_get_*_map by
 * integration/Makefile. lib/surface.h is untouched.
 */
#include <mutex>
#include <stdexcept>
#include <unordered_map>

#include "surface.h"

#include "b200_context.h"

SMVS_NAMESPACE_BEGIN

extern "C" mve::FloatImage::Ptr smvs_ref_surface_get_depth_map (Surface* self);
extern "C" mve::FloatImage::Ptr smvs_ref_surface_get_normal_map (Surface* self,
    float inv_flen);

namespace
{
    struct HeldMaps
    {
        std::weak_ptr<Surface> owner;     /* expired = the address was reused */
        mve::FloatImage::Ptr depth, normals;
        float inv_flen;
    };
    std::mutex g_mutex;
    std::unordered_map<Surface const*, HeldMaps> g_held;

    bool
    held_for (Surface const* surface, HeldMaps* out)
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        auto it = g_held.find(surface);
        if (it == g_held.end())
            return false;
        std::shared_ptr<Surface> alive = it->second.owner.lock();
        if (alive.get() != surface)
        {
            g_held.erase(it);
            return false;
        }
        *out = it->second;
        return true;
    }
}

mve::FloatImage::Ptr
Surface::get_depth_map (void)
{
    HeldMaps held;
    if (held_for(this, &held))
        return held.depth->duplicate();
    return smvs_ref_surface_get_depth_map(this);
}

mve::FloatImage::Ptr
Surface::get_normal_map (float inv_flen)
{
    HeldMaps held;
    if (held_for(this, &held))
    {
        if (inv_flen != held.inv_flen)
            throw std::runtime_error("smvs_b200: normal map of the resident "
                "surface asked for another focal length than it was "
                "rendered with (set SMVSB_REBUILD_SURFACE=1)");
        return held.normals->duplicate();
    }
    return smvs_ref_surface_get_normal_map(this, inv_flen);
}

SMVS_NAMESPACE_END

namespace smvs_b200_integration {

void
hold_maps (smvs::Surface::Ptr const& surface, mve::FloatImage::Ptr depth,
    mve::FloatImage::Ptr normals, float inv_flen)
{
    std::lock_guard<std::mutex> lock(smvs::g_mutex);
    /* entries of surfaces that are gone */
    for (auto it = smvs::g_held.begin(); it != smvs::g_held.end();)
        if (it->second.owner.expired())
            it = smvs::g_held.erase(it);
        else
            ++it;
    smvs::HeldMaps h;
    h.owner = surface;
    h.depth = depth;
    h.normals = normals;
    h.inv_flen = inv_flen;
    smvs::g_held[surface.get()] = h;
}

}
