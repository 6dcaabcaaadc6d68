/* Shim of MVE mve/bundle.h (SfM feature list). TEST INFRASTRUCTURE ONLY. */
#ifndef SHIM_MVE_BUNDLE_HEADER
#define SHIM_MVE_BUNDLE_HEADER

#include <memory>
#include <vector>

#include "mve/camera.h"

MVE_NAMESPACE_BEGIN

class Bundle
{
public:
    struct Feature2D
    {
        int view_id;
        int feature_id;
        float pos[2];
    };
    struct Feature3D
    {
        float pos[3];
        float color[3];
        std::vector<Feature2D> refs;
    };
    typedef std::shared_ptr<Bundle> Ptr;
    typedef std::shared_ptr<Bundle const> ConstPtr;
    typedef std::vector<CameraInfo> Cameras;
    typedef std::vector<Feature3D> Features;

    static Ptr create (void) { return Ptr(new Bundle()); }
    Cameras const& get_cameras (void) const { return cameras; }
    Cameras& get_cameras (void) { return cameras; }
    Features const& get_features (void) const { return features; }
    Features& get_features (void) { return features; }

private:
    Cameras cameras;
    Features features;
};

MVE_NAMESPACE_END

#endif
