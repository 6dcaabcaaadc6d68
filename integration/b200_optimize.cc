/*
 * integration/b200_optimize.cc
 *
 * Drop-in body for smvs::DepthOptimizer::optimize (reference:
 * lib/depth_optimizer.cc:54-162). In the use_sgm mode without debug output the
 * whole coarse-to-fine ladder of the view runs on the GPU through ONE call:
 * smvsb_optimize for single-channel views (uploaded as byte images),
 * smvsb_optimize_rgb_f32 for three-channel views (uploaded as the float
 * images StereoView::get_image() holds -- what real MVE scenes contain). The
 * depth and normal maps come back, and what optimize() has to leave behind --
 * the two view embeddings (:158-161), the final surface for get_depth() /
 * get_normals(), the fitted lighting -- is put in place. use_sgm = false
 * (--no-sgm; colour views and a bundle) takes the same call with the sparse
 * depth of the bundle's features as initial depth. Every other configuration
 * (views of mixed or other channel counts, debug levels that write
 * intermediate images) runs the reference's own optimize(), whose members are the per-call
 * drop-ins of b200_depth_optimizer.cc. lib/depth_optimizer.h is untouched.
 *
 *   SMVSB_MEMBERWISE=1        forces the reference's optimize() (per-member path)
 *   SMVSB_REBUILD_SURFACE=1   rebuilds the final surface as a host object
 *                             (default: a stand-in whose get_depth_map /
 *                             get_normal_map return the device's maps,
 *                             b200_surface.cc)
 */
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "depth_optimizer.h"

#include "b200_context.h"

namespace smvs_b200_integration {
/* b200_surface.cc */
void hold_maps (smvs::Surface::Ptr const& surface, mve::FloatImage::Ptr depth,
    mve::FloatImage::Ptr normals, float inv_flen);
}

SMVS_NAMESPACE_BEGIN

/* The reference's own optimize(), kept under this name by integration/Makefile
 * (objcopy --redefine-sym on the all-weak copy of the object). */
extern "C" void smvs_ref_optimize (DepthOptimizer* self);

namespace
{
    bool
    has_channels (StereoView::Ptr const& v, int channels)
    {
        return v->get_image() != nullptr
            && v->get_image()->channels() == channels;
    }
}

void
DepthOptimizer::optimize (void)
{
    bool const colour = has_channels(this->main_view, 3);
    int const channels = colour ? 3 : 1;
    /* use_sgm = false: the NCC filter needs colour views, the initial
     * surface the bundle's features */
    bool const no_sgm = !this->opts.use_sgm;
    bool resident = (!no_sgm || (colour && this->bundle != nullptr))
        && this->opts.debug_lvl == 0
        && std::getenv("SMVSB_MEMBERWISE") == nullptr
        && has_channels(this->main_view, channels) && !this->sub_views.empty()
        && this->sub_views.size() <= 32;
    for (auto const& v : this->sub_views)
        resident = resident && has_channels(v, channels);
    if (this->opts.use_shading)
        resident = resident && this->main_view->get_shading_image() != nullptr;
    if (!resident)
    {
        smvs_ref_optimize(this);
        return;
    }

    smvsb::Context& gpu = smvs_b200_integration::thread_context();
    int const w = this->main_view->get_width();
    int const h = this->main_view->get_height();

    /* the view as the reference's StereoViews hold it */
    std::size_t const n = this->sub_views.size();
    mve::ByteImage::ConstPtr main_img;
    std::vector<mve::ByteImage::ConstPtr> sub_keep(n);
    std::vector<int> sw(n), sh(n);
    std::vector<uint8_t const*> simg(n);
    std::vector<float const*> simg_rgb(n);
    std::vector<double> M(9 * n), t(3 * n);
    if (!colour)
        main_img = this->main_view->get_byte_image();
    for (std::size_t k = 0; k < n; ++k)
    {
        if (colour)
        {
            mve::FloatImage::ConstPtr img = this->sub_views[k]->get_image();
            sw[k] = img->width();
            sh[k] = img->height();
            simg_rgb[k] = img->begin();
        }
        else
        {
            sub_keep[k] = this->sub_views[k]->get_byte_image();
            sw[k] = sub_keep[k]->width();
            sh[k] = sub_keep[k]->height();
            simg[k] = sub_keep[k]->begin();
        }
        for (int j = 0; j < 9; ++j) M[9 * k + j] = this->Mi[k][j];
        for (int j = 0; j < 3; ++j) t[3 * k + j] = this->ti[k][j];
    }
    math::Matrix3f invproj;
    this->main_view->get_camera().fill_inverse_calibration(*invproj, w, h);
    mve::FloatImage::Ptr sgm;
    if (no_sgm)
    {
        /* the depth image Surface::create makes of the bundle
         * (lib/surface.cc:43-46 with initialize_depth_from_bundle, :91-128):
         * every feature the main view observes, projected with the view's
         * camera in the reference's types and order of operations; its depth
         * in the pixel it falls into, later features overwrite earlier ones */
        sgm = mve::FloatImage::create(w, h, 1);
        sgm->fill(0.0f);
        mve::CameraInfo const& cam = this->main_view->get_camera();
        int const view_id = this->main_view->get_view_id();
        math::Matrix3f const rot(cam.rot);
        math::Vec3f const trans(cam.trans);
        float const flen = cam.flen;
        double const half_w = static_cast<double>(w) / 2.0;
        double const half_h = static_cast<double>(h) / 2.0;
        double const norm = static_cast<double>(std::max(w, h));
        for (mve::Bundle::Feature3D const& feat
            : this->bundle->get_features())
            for (mve::Bundle::Feature2D const& ref : feat.refs)
            {
                if (ref.view_id != view_id)
                    continue;
                math::Vec3f const fpos(feat.pos);
                math::Vec3f proj = rot * fpos + trans;
                float const depth = proj[2];
                proj[0] = proj[0] * flen / proj[2];
                proj[1] = proj[1] * flen / proj[2];
                float const ix = proj[0] * norm + half_w;
                float const iy = proj[1] * norm + half_h;
                int const x = std::floor(ix), y = std::floor(iy);
                if (x >= 0 && x < w && y >= 0 && y < h)
                    sgm->at(x, y, 0) = depth;
                break;
            }
    }
    else
        sgm = this->main_view->get_sgm_depth();                     /* :41 */
    bool const lit = this->opts.use_shading;

    smvsb_optimize_options o;
    o.regularization = this->opts.regularization;
    o.light_surf_regularization = this->opts.light_surf_regularization;
    o.num_iterations = this->opts.num_iterations;
    o.min_scale = this->opts.min_scale;
    o.use_shading = lit ? 1 : 0;
    o.full_optimization = this->opts.full_optimization ? 1 : 0;
    o.no_sgm = no_sgm ? 1 : 0;
    o.reserved = 0;

    mve::FloatImage::Ptr depth = mve::FloatImage::create(w, h, 1);
    mve::FloatImage::Ptr normals = mve::FloatImage::create(w, h, 3);
    double light[16];
    smvsb_optimize_stats st;
    float const* shading = lit
        ? this->main_view->get_shading_image()->begin() : nullptr;
    float const* shading_grad = lit
        ? this->main_view->get_shading_gradients()->begin() : nullptr;
    if (colour)
        gpu.check(smvsb_optimize_rgb_f32(gpu.get(), w, h,
            this->main_view->get_flen(), this->main_view->get_inverse_flen(),
            *invproj, this->main_view->get_image()->begin(),
            static_cast<int>(n), sw.data(), sh.data(), simg_rgb.data(),
            M.data(), t.data(), shading, shading_grad, sgm->width(),
            sgm->height(), sgm->begin(), &o, depth->begin(), normals->begin(),
            light, &st));
    else
        gpu.check(smvsb_optimize(gpu.get(), w, h, this->main_view->get_flen(),
            this->main_view->get_inverse_flen(), *invproj, main_img->begin(),
            static_cast<int>(n), sw.data(), sh.data(), simg.data(), M.data(),
            t.data(), shading, shading_grad, sgm->width(), sgm->height(),
            sgm->begin(), &o, depth->begin(), normals->begin(), light, &st));

    /* ---- what optimize() leaves behind -------------------------------- */
    int const init_scale = st.final_scale + (st.scales - 1);
    if (std::getenv("SMVSB_REBUILD_SURFACE") == nullptr)
    {
        /* get_depth() / get_normals() (the only public doors to the private
         * surface) are served from the maps the device rendered, see
         * b200_surface.cc; the member itself is a small stand-in */
        mve::FloatImage::Ptr ones = mve::FloatImage::create(w, h, 1);
        ones->fill(1.0f);
        this->surface = Surface::create(nullptr, this->main_view, init_scale,
            ones);
        smvs_b200_integration::hold_maps(this->surface, depth->duplicate(),
            normals->duplicate(), this->main_view->get_inverse_flen());
    }
    else
    {
    /* SMVSB_REBUILD_SURFACE=1: the final surface as a host object, for a
     * host that reaches the surface some other way: the grid geometry
     * of the ladder from a stand-in created and subdivided by the reference's
     * own code (a constant depth image makes every node and patch exist),
     * then the device's nodes and validity */
    {
        mve::FloatImage::Ptr ones = mve::FloatImage::create(w, h, 1);
        ones->fill(1.0f);
        this->surface = Surface::create(nullptr, this->main_view, init_scale,
            ones);
    }
    while (this->surface->get_scale() > st.final_scale)
    {
        this->surface->subdivide_patches();
        /* the ring a subdivision adds around the grid (:993-1014) is filled
         * from the depth, as optimize() does (:99) */
        this->surface->fill_patches_from_depth();
    }
    {
        Surface::NodeList const& nodes = this->surface->get_nodes();
        Surface::PatchList const& patches = this->surface->get_patches();
        std::vector<double> values(nodes.size() * 4);
        std::vector<uint8_t> nvalid(nodes.size()), pvalid(patches.size());
        gpu.check(smvsb_get_nodes(gpu.get(), values.data()));
        gpu.check(smvsb_get_surface_state(gpu.get(), nvalid.data(),
            pvalid.data(), nullptr, nullptr, 0));
        for (std::size_t p = 0; p < patches.size(); ++p)
            if (!pvalid[p] && patches[p] != nullptr)
                this->surface->delete_patch(p);
        this->surface->remove_nodes_without_patch();
        std::vector<double> zero(nodes.size() * 4, 0.0), unused;
        this->surface->update_nodes(zero, &unused);    /* resets patch caches */
        for (std::size_t i = 0; i < nodes.size(); ++i)
        {
            if (nodes[i] == nullptr)
                continue;
            nodes[i]->f = values[4 * i + 0];
            nodes[i]->dx = values[4 * i + 1];
            nodes[i]->dy = values[4 * i + 2];
            nodes[i]->dxy = values[4 * i + 3];
        }
    }
    }
    if (lit && st.final_scale < 4)
    {
        GlobalLighting::Params p;
        for (int i = 0; i < 16; ++i)
            p[i] = light[i];
        this->lighting = GlobalLighting::create(p);
    }
    /* :158-161 */
    this->main_view->write_depth_to_view(depth, this->opts.output_name);
    this->main_view->write_image_to_view(normals,
        this->opts.output_name + "N");
}

SMVS_NAMESPACE_END
