/*
 * integration/b200_depth_optimizer.cc
 *
 * Drop-in bodies for four members of smvs::DepthOptimizer that run on the
 * GPU through the C ABI of libsmvs_b200.so:
 *   run_newton_iterations    lib/depth_optimizer.cc:164-358 (inner Newton
 *                            loop :204-304 -> smvsb_newton_loop)
 *   create_subview_surfaces  :433-604 -> smvsb_visibility (both modes; the
 *                            use_sgm = false mode with its NCC filter
 *                            ncc_for_patch :795-912 needs 3-channel images,
 *                            otherwise the reference's body runs)
 *   cut_boundaries           :360-431 -> smvsb_cut_boundaries
 *   depthmap_bilateral_filter :957-1004 -> smvsb_bilateral_filter
 * Everything else -- surface expansion / subdivision, isolated-patch removal,
 * the patch-count convergence test -- still calls the reference's own member
 * functions, and the host Surface stays the owner of the state (uploaded
 * before, read back after every call). lib/depth_optimizer.h is untouched:
 * this file only DEFINES members, so it compiles against the unmodified
 * header.
 *
 * Built by integration/Makefile together with the reference's unmodified
 * objects (its own definitions of these symbols weakened with objcopy) into
 * integration/_build/libsmvs_ref_b200.so, which tests/test_integration.py
 * drives side by side with the pure-CPU build.
 */
#include <cmath>
#include <cstdint>
#include <iostream>
#include <limits>
#include <memory>
#include <chrono>
#include <cstdlib>
#include <vector>

#include "depth_optimizer.h"

#include "b200_context.h"

SMVS_NAMESPACE_BEGIN

namespace
{
    using smvs_b200_integration::thread_context;

    int
    count_patches (Surface::Ptr surface)
    {
        int n = 0;
        for (auto const& p : surface->get_patches())
            n += (p != nullptr);
        return n;
    }

    /* Which optimizer / scale the context's views belong to. */
    struct ViewsKey
    {
        void const* owner = nullptr;
        void const* gradients = nullptr;
        int scale = -1;
        std::uint64_t generation = 0;   /* b200_context.h: views_generation */
    };
    ViewsKey&
    views_key (void)
    {
        static thread_local ViewsKey key;
        return key;
    }

    /* StereoView images + reprojections -> smvsb_set_views. */
    void
    upload_views (smvsb::Context& gpu, void const* owner, int scale,
        StereoView::Ptr main_view, std::vector<StereoView::Ptr> const& subs,
        std::vector<math::Matrix3d> const& Mi,
        std::vector<math::Vec3d> const& ti)
    {
        std::size_t const n = subs.size();
        std::vector<int> sw(n), sh(n);
        std::vector<float const*> sg(n), shess(n);
        std::vector<double> M(9 * n), t(3 * n);
        for (std::size_t k = 0; k < n; ++k)
        {
            sw[k] = subs[k]->get_width();
            sh[k] = subs[k]->get_height();
            sg[k] = subs[k]->get_image_gradients()->begin();
            shess[k] = subs[k]->get_image_hessian()->begin();
            for (int j = 0; j < 9; ++j) M[9 * k + j] = Mi[k][j];
            for (int j = 0; j < 3; ++j) t[3 * k + j] = ti[k][j];
        }
        bool const lit = (main_view->get_shading_image() != nullptr);
        gpu.check(smvsb_set_views(gpu.get(), main_view->get_width(),
            main_view->get_height(), main_view->get_flen(),
            main_view->get_inverse_flen(),
            main_view->get_image_gradients()->begin(),
            lit ? main_view->get_shading_image()->begin() : nullptr,
            lit ? main_view->get_shading_gradients()->begin() : nullptr,
            static_cast<int>(n), sw.data(), sh.data(), sg.data(),
            shess.data(), M.data(), t.data()));
        views_key().owner = owner;
        views_key().gradients = main_view->get_image_gradients().get();
        views_key().scale = scale;
        views_key().generation = smvs_b200_integration::views_generation();
    }

    bool
    views_current (void const* owner, int scale, StereoView::Ptr main_view)
    {
        ViewsKey const& k = views_key();
        return k.owner == owner && k.scale == scale
            && k.generation == smvs_b200_integration::views_generation()
            && k.gradients == main_view->get_image_gradients().get();
    }

    /* Surface (+ visibility lists, if given) -> smvsb_set_surface. */
    struct PackedSurface
    {
        int npx = 0, npy = 0;
        std::vector<double> node_values;
        std::vector<uint8_t> node_valid, patch_valid;
    };

    void
    upload_surface (smvsb::Context& gpu, Surface::Ptr surface,
        std::vector<std::vector<std::size_t>> const* subsurfaces,
        PackedSurface* out)
    {
        Surface::NodeList const& nodes = surface->get_nodes();
        Surface::PatchList const& patches = surface->get_patches();
        std::size_t ids[4];
        surface->fill_node_ids_for_patch(0, ids);
        int const npx = static_cast<int>(ids[2]) - 1;
        int const npy = static_cast<int>(patches.size()) / npx;
        int const ps = surface->get_patchsize();
        out->npx = npx;
        out->npy = npy;
        out->node_values.assign(nodes.size() * 4, 0.0);
        out->node_valid.assign(nodes.size(), 0);
        out->patch_valid.assign(patches.size(), 0);
        std::vector<uint32_t> vis_off(patches.size() + 1, 0);
        std::vector<uint8_t> vis_ids;
        for (std::size_t i = 0; i < nodes.size(); ++i)
        {
            if (nodes[i] == nullptr)
                continue;
            out->node_valid[i] = 1;
            out->node_values[4 * i + 0] = nodes[i]->f;
            out->node_values[4 * i + 1] = nodes[i]->dx;
            out->node_values[4 * i + 2] = nodes[i]->dy;
            out->node_values[4 * i + 3] = nodes[i]->dxy;
        }
        int start_x = 0, start_y = 0;
        for (std::size_t p = 0; p < patches.size(); ++p)
        {
            vis_off[p] = static_cast<uint32_t>(vis_ids.size());
            if (patches[p] == nullptr)
                continue;
            out->patch_valid[p] = 1;
            start_x = patches[p]->get_x() - static_cast<int>(p % npx) * ps;
            start_y = patches[p]->get_y() - static_cast<int>(p / npx) * ps;
            if (subsurfaces != nullptr && p < subsurfaces->size())
                for (std::size_t id : (*subsurfaces)[p])
                    vis_ids.push_back(static_cast<uint8_t>(id));
        }
        vis_off[patches.size()] = static_cast<uint32_t>(vis_ids.size());
        if (vis_ids.empty())
            vis_ids.push_back(0);
        gpu.check(smvsb_set_surface(gpu.get(), surface->get_scale(),
            npx, npy, start_x, start_y, out->node_values.data(),
            out->node_valid.data(), out->patch_valid.data(),
            subsurfaces != nullptr ? vis_off.data() : nullptr,
            subsurfaces != nullptr ? vis_ids.data() : nullptr));
    }

    /* Patches the device deleted -> Surface::delete_patch, then the
     * reference's own node clean-up. Returns the number deleted. */
    int
    apply_deletions (smvsb::Context& gpu, Surface::Ptr surface,
        PackedSurface const& before)
    {
        std::vector<uint8_t> now(before.patch_valid.size());
        gpu.check(smvsb_get_surface_state(gpu.get(), nullptr, now.data(),
            nullptr, nullptr, 0));
        int deleted = 0;
        for (std::size_t p = 0; p < now.size(); ++p)
            if (before.patch_valid[p] && !now[p])
            {
                surface->delete_patch(p);
                deleted += 1;
            }
        if (deleted > 0)
            surface->remove_nodes_without_patch();
        return deleted;
    }
}

/* The reference's own create_subview_surfaces, kept under this name by
 * integration/Makefile (objcopy --redefine-sym on a private copy of the
 * object): the use_sgm = false mode on images that are not 3-channel (where
 * the reference's ncc_for_patch indexes channels 1 and 2 of whatever it is
 * given, lib/depth_optimizer.cc:884-889). */
extern "C" void smvs_ref_create_subview_surfaces (DepthOptimizer* self);

namespace
{
    /* whose colour images the context holds (they do not depend on the
     * scale: StereoView::get_image() is the unscaled image) */
    struct ColorKey
    {
        void const* owner = nullptr;
        std::uint64_t generation = 0;
    };

    bool
    upload_color_images (smvsb::Context& gpu, void const* owner,
        StereoView::Ptr main_view, std::vector<StereoView::Ptr> const& subs)
    {
        static thread_local ColorKey key;
        if (main_view->get_image()->channels() != 3)
            return false;
        std::vector<float const*> ptrs(subs.size());
        for (std::size_t k = 0; k < subs.size(); ++k)
        {
            if (subs[k]->get_image()->channels() != 3)
                return false;
            ptrs[k] = subs[k]->get_image()->begin();
        }
        std::uint64_t const gen = smvs_b200_integration::views_generation();
        if (key.owner == owner && key.generation == gen)
            return true;
        gpu.check(smvsb_set_color_images(gpu.get(),
            main_view->get_image()->begin(), static_cast<int>(subs.size()),
            ptrs.data()));
        key.owner = owner;
        key.generation = gen;
        return true;
    }
}

void
DepthOptimizer::create_subview_surfaces (void)
{
    smvsb::Context& gpu = thread_context();
    int const scale = this->surface->get_scale();
    if (!this->opts.use_sgm
        && (this->main_view->get_image()->channels() != 3
            || this->sub_views.empty()))
    {
        smvs_ref_create_subview_surfaces(this);
        return;
    }
    if (!views_current(this, scale, this->main_view))
        upload_views(gpu, this, scale, this->main_view, this->sub_views,
            this->Mi, this->ti);
    if (!this->opts.use_sgm && !upload_color_images(gpu, this,
        this->main_view, this->sub_views))
    {
        smvs_ref_create_subview_surfaces(this);
        return;
    }

    PackedSurface packed;
    upload_surface(gpu, this->surface, nullptr, &packed);
    uint64_t removed = 0;
    gpu.check(smvsb_visibility(gpu.get(), this->opts.use_sgm
        ? this->sgm_depth->begin() : nullptr, &removed));

    std::size_t const np = packed.patch_valid.size();
    std::vector<uint32_t> vis_off(np + 1);
    std::vector<uint8_t> vis_ids(np * this->sub_views.size() + 1);
    gpu.check(smvsb_get_surface_state(gpu.get(), nullptr, nullptr,
        vis_off.data(), vis_ids.data(), vis_ids.size()));
    this->subsurfaces.clear();
    this->subsurfaces.resize(np);
    for (std::size_t p = 0; p < np; ++p)
        for (uint32_t k = vis_off[p]; k < vis_off[p + 1]; ++k)
            this->subsurfaces[p].push_back(vis_ids[k]);
    int const deleted = apply_deletions(gpu, this->surface, packed);
    if (this->opts.debug_lvl > 0)
        std::cout << "Removed " << deleted << " patches "
            "due to occlusions." << std::endl;
}

mve::FloatImage::Ptr
DepthOptimizer::depthmap_bilateral_filter (mve::FloatImage::ConstPtr dm,
    mve::FloatImage::ConstPtr ci, float sigma, int kernel_size)
{
    smvsb::Context& gpu = thread_context();
    mve::FloatImage::Ptr out = mve::FloatImage::create(ci->width(),
        ci->height(), 1);
    gpu.check(smvsb_bilateral_filter(gpu.get(), ci->width(), ci->height(),
        ci->channels(), ci->begin(), dm->width(), dm->height(), dm->begin(),
        sigma, kernel_size, out->begin()));
    return out;
}

int
DepthOptimizer::cut_boundaries (void)
{
    smvsb::Context& gpu = thread_context();
    int const scale = this->surface->get_scale();
    if (!views_current(this, scale, this->main_view))
        upload_views(gpu, this, scale, this->main_view, this->sub_views,
            this->Mi, this->ti);
    PackedSurface packed;
    upload_surface(gpu, this->surface, &this->subsurfaces, &packed);
    math::Matrix3f invproj;
    this->main_view->get_camera().fill_inverse_calibration(*invproj,
        this->main_view->get_width(), this->main_view->get_height());
    int deleted = 0;
    gpu.check(smvsb_cut_boundaries(gpu.get(), *invproj, &deleted));
    apply_deletions(gpu, this->surface, packed);
    return deleted;
}

void
DepthOptimizer::run_newton_iterations (int num_iters)
{
    smvsb::Context& gpu = thread_context();
    this->main_gradients = this->main_view->get_image_gradients();

    /* ---- images of this scale: once per call (set_scale precedes it) ---- */
    upload_views(gpu, this, this->surface->get_scale(), this->main_view,
        this->sub_views, this->Mi, this->ti);

    /* SMVSB_TIMING=1: where this call's wall time goes (host code of the
     * reference vs. the path on the GPU), one line per call */
    typedef std::chrono::steady_clock Clock;
    double t_vis = 0, t_cut = 0, t_pack = 0, t_gpu = 0, t_unpack = 0, t_topo = 0;
    Clock::time_point tick = Clock::now();
    auto lap = [&tick] (double& acc) {
        Clock::time_point const now = Clock::now();
        acc += std::chrono::duration<double>(now - tick).count();
        tick = now;
    };
    lap(t_pack);   /* smvsb_set_views above */

    bool converged = false;
    for (int iter = 0; iter < num_iters; ++iter)
    {
        int const patches_before = count_patches(this->surface);
        tick = Clock::now();
        if (iter == 0)
        {
            /* :189-195, reference code */
            this->create_subview_surfaces();
            lap(t_vis);
            for (int del = std::numeric_limits<int>::max(); del > 10;)
                del = this->cut_boundaries();
            lap(t_cut);
        }

        /* ---- Surface + visibility -> device (replaces :197-213) -------- */
        Surface::NodeList const& nodes = this->surface->get_nodes();
        PackedSurface packed;
        upload_surface(gpu, this->surface, &this->subsurfaces, &packed);
        std::vector<double>& node_values = packed.node_values;

        lap(t_pack);
        /* ---- the inner Newton loop, :219-304, on the GPU ---------------- */
        double light[16];
        if (this->lighting != nullptr)
        {
            GlobalLighting::Params const p = this->lighting->get_parameters();
            for (int l = 0; l < 16; ++l)
                light[l] = p[l];
        }
        smvsb_newton_stats st;
        gpu.check(smvsb_newton_loop(gpu.get(),
            this->lighting != nullptr ? light : nullptr,
            this->opts.regularization, this->opts.light_surf_regularization,
            200, this->opts.full_optimization ? 1 : 0, &st));
        if (this->opts.debug_lvl > 0)
            std::cout << "### Finished iteration: " << iter
                << " (B200) Newton steps: " << st.newton_steps
                << " CG iterations: " << st.cg_iterations
                << " construct/solve/update ms: " << st.ms_construct << " / "
                << st.ms_solve << " / " << st.ms_update << std::endl;

        lap(t_gpu);
        /* ---- nodes back into the Surface (Surface::update_nodes' job) --- */
        gpu.check(smvsb_get_nodes(gpu.get(), node_values.data()));
        std::vector<double> delta(node_values.size(), 0.0), unused;
        for (std::size_t i = 0; i < nodes.size(); ++i)
        {
            if (nodes[i] == nullptr)
                continue;
            delta[4 * i + 0] = node_values[4 * i + 0] - nodes[i]->f;
            delta[4 * i + 1] = node_values[4 * i + 1] - nodes[i]->dx;
            delta[4 * i + 2] = node_values[4 * i + 2] - nodes[i]->dy;
            delta[4 * i + 3] = node_values[4 * i + 3] - nodes[i]->dxy;
        }
        this->surface->update_nodes(delta, &unused);   /* resets patch caches */
        for (std::size_t i = 0; i < nodes.size(); ++i)
        {
            if (nodes[i] == nullptr)
                continue;
            nodes[i]->f = node_values[4 * i + 0];      /* exact device values */
            nodes[i]->dx = node_values[4 * i + 1];
            nodes[i]->dy = node_values[4 * i + 2];
            nodes[i]->dxy = node_values[4 * i + 3];
        }

        lap(t_unpack);
        /* ---- :318-356, reference code ---------------------------------- */
        if (converged)
            break;
        for (int del = std::numeric_limits<int>::max(); del > 10;)
            del = this->cut_boundaries();
        lap(t_cut);
        if (!this->opts.use_sgm)
        {
            this->surface->expand();
            lap(t_topo);
            this->create_subview_surfaces();
            lap(t_vis);
            for (int del = std::numeric_limits<int>::max(); del > 10;)
                del = this->cut_boundaries();
            lap(t_cut);
        }
        this->surface->remove_isolated_patches();
        lap(t_topo);
        int const patches_after = count_patches(this->surface);
        double const change = 1.0
            - static_cast<double>(std::min(patches_after, patches_before))
            / static_cast<double>(std::max(patches_after, patches_before));
        if (iter > 0 && (patches_after <= patches_before
            || change < 0.05 * this->surface->get_scale()))
            converged = true;
    }
    if (std::getenv("SMVSB_TIMING") != nullptr)
        std::cerr << "[b200] scale " << this->surface->get_scale()
            << ": create_subview_surfaces " << t_vis << " s, cut_boundaries "
            << t_cut << " s, expand/remove_isolated " << t_topo
            << " s, pack+upload " << t_pack << " s, GPU newton loops "
            << t_gpu << " s, download+unpack " << t_unpack << " s" << std::endl;
}

SMVS_NAMESPACE_END
