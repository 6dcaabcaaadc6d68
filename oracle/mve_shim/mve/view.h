/*
 * Shim of MVE mve/view.h: an in-memory map embedding-name -> image plus a
 * camera. save_view/cache_cleanup are no-ops. TEST INFRASTRUCTURE ONLY.
 */
#ifndef SHIM_MVE_VIEW_HEADER
#define SHIM_MVE_VIEW_HEADER

#include <map>
#include <memory>
#include <string>

#include "util/strings.h"
#include "mve/camera.h"
#include "mve/image.h"

MVE_NAMESPACE_BEGIN

class View
{
public:
    typedef std::shared_ptr<View> Ptr;
    typedef std::shared_ptr<View const> ConstPtr;

    static Ptr create (void) { return Ptr(new View()); }

    View (void) : id(0) {}

    void set_id (int view_id) { this->id = view_id; }
    int get_id (void) const { return this->id; }
    void set_camera (CameraInfo const& camera) { this->cam = camera; }
    CameraInfo const& get_camera (void) const { return this->cam; }

    void set_image (ImageBase::Ptr image, std::string const& name)
    { this->images[name] = image; }
    bool has_image (std::string const& name) const
    { return this->images.count(name) > 0; }
    ImageBase::Ptr get_image (std::string const& name)
    {
        auto it = this->images.find(name);
        return it == this->images.end() ? ImageBase::Ptr() : it->second;
    }
    ByteImage::Ptr get_byte_image (std::string const& name)
    { return std::dynamic_pointer_cast<ByteImage>(this->get_image(name)); }
    FloatImage::Ptr get_float_image (std::string const& name)
    {
        /* Callers modify the returned image (depth conventions): hand out a
         * copy, like MVE handing out a freshly loaded embedding. */
        FloatImage::Ptr img = std::dynamic_pointer_cast<FloatImage>(
            this->get_image(name));
        return img == nullptr ? img : img->duplicate();
    }
    void remove_image (std::string const& name) { this->images.erase(name); }

    /* what lib/mesh_generator.cc:176-178 reads of MVE's image proxies */
    struct ImageProxy
    {
        int width, height, channels;
    };
    ImageProxy const* get_image_proxy (std::string const& name)
    {
        ImageBase::Ptr img = this->get_image(name);
        if (img == nullptr)
            return nullptr;
        this->proxy.width = img->width();
        this->proxy.height = img->height();
        this->proxy.channels = img->channels();
        return &this->proxy;
    }

    void save_view (void) {}
    int cache_cleanup (void) { return 0; }

private:
    int id;
    CameraInfo cam;
    std::map<std::string, ImageBase::Ptr> images;
    ImageProxy proxy;
};

MVE_NAMESPACE_END

#endif
