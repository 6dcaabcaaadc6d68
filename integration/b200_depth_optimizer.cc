/*
 * integration/b200_depth_optimizer.cc
 *
 * Drop-in body for smvs::DepthOptimizer::run_newton_iterations
 * (reference: lib/depth_optimizer.cc:164-358) that runs the inner Newton loop
 * (:204-304) on the GPU through the C ABI of libsmvs_b200.so. Everything that
 * is not the hot path -- visibility lists, boundary cutting, surface
 * expansion, the patch-count convergence test -- still calls the reference's
 * own member functions. lib/depth_optimizer.h is untouched: this file only
 * DEFINES the member, so it compiles against the unmodified header.
 *
 * Built by integration/Makefile together with the reference's unmodified
 * objects (its own definition of this one symbol weakened with objcopy) into
 * integration/_build/libsmvs_ref_b200.so, which tests/test_integration.py
 * drives side by side with the pure-CPU build.
 */
#include <cmath>
#include <cstdint>
#include <iostream>
#include <limits>
#include <memory>
#include <vector>

#include "depth_optimizer.h"

#include "smvs_b200.hpp"

SMVS_NAMESPACE_BEGIN

namespace
{
    /* One GPU context per host thread: the reference runs one DepthOptimizer
     * per pool thread (app/smvsrecon.cc:658-733). */
    smvsb::Context&
    thread_context (void)
    {
        static thread_local std::unique_ptr<smvsb::Context> ctx;
        if (!ctx)
            ctx.reset(new smvsb::Context(0));
        return *ctx;
    }

    int
    count_patches (Surface::Ptr surface)
    {
        int n = 0;
        for (auto const& p : surface->get_patches())
            n += (p != nullptr);
        return n;
    }
}

void
DepthOptimizer::run_newton_iterations (int num_iters)
{
    smvsb::Context& gpu = thread_context();
    this->main_gradients = this->main_view->get_image_gradients();

    /* ---- images of this scale: once per call (set_scale precedes it) ---- */
    {
        std::size_t const n = this->sub_views.size();
        std::vector<int> sw(n), sh(n);
        std::vector<float const*> sg(n), shess(n);
        std::vector<double> M(9 * n), t(3 * n);
        for (std::size_t k = 0; k < n; ++k)
        {
            sw[k] = this->sub_views[k]->get_width();
            sh[k] = this->sub_views[k]->get_height();
            sg[k] = this->sub_views[k]->get_image_gradients()->begin();
            shess[k] = this->sub_views[k]->get_image_hessian()->begin();
            for (int j = 0; j < 9; ++j) M[9 * k + j] = this->Mi[k][j];
            for (int j = 0; j < 3; ++j) t[3 * k + j] = this->ti[k][j];
        }
        bool const lit = (this->main_view->get_shading_image() != nullptr);
        gpu.check(smvsb_set_views(gpu.get(), this->main_view->get_width(),
            this->main_view->get_height(), this->main_view->get_flen(),
            this->main_view->get_inverse_flen(),
            this->main_gradients->begin(),
            lit ? this->main_view->get_shading_image()->begin() : nullptr,
            lit ? this->main_view->get_shading_gradients()->begin() : nullptr,
            static_cast<int>(n), sw.data(), sh.data(), sg.data(),
            shess.data(), M.data(), t.data()));
    }

    bool converged = false;
    for (int iter = 0; iter < num_iters; ++iter)
    {
        int const patches_before = count_patches(this->surface);
        if (iter == 0)
        {
            /* :189-195, reference code */
            this->create_subview_surfaces();
            for (int del = std::numeric_limits<int>::max(); del > 10;)
                del = this->cut_boundaries();
        }

        /* ---- Surface + visibility -> device (replaces :197-213) -------- */
        Surface::NodeList const& nodes = this->surface->get_nodes();
        Surface::PatchList const& patches = this->surface->get_patches();
        std::size_t ids[4];
        this->surface->fill_node_ids_for_patch(0, ids);
        int const npx = static_cast<int>(ids[2]) - 1;
        int const npy = static_cast<int>(patches.size()) / npx;
        int const ps = this->surface->get_patchsize();
        std::vector<double> node_values(nodes.size() * 4, 0.0);
        std::vector<uint8_t> node_valid(nodes.size(), 0);
        std::vector<uint8_t> patch_valid(patches.size(), 0);
        std::vector<uint32_t> vis_off(patches.size() + 1, 0);
        std::vector<uint8_t> vis_ids;
        for (std::size_t i = 0; i < nodes.size(); ++i)
        {
            if (nodes[i] == nullptr)
                continue;
            node_valid[i] = 1;
            node_values[4 * i + 0] = nodes[i]->f;
            node_values[4 * i + 1] = nodes[i]->dx;
            node_values[4 * i + 2] = nodes[i]->dy;
            node_values[4 * i + 3] = nodes[i]->dxy;
        }
        int start_x = 0, start_y = 0;
        for (std::size_t p = 0; p < patches.size(); ++p)
        {
            vis_off[p] = static_cast<uint32_t>(vis_ids.size());
            if (patches[p] == nullptr)
                continue;
            patch_valid[p] = 1;
            start_x = patches[p]->get_x() - static_cast<int>(p % npx) * ps;
            start_y = patches[p]->get_y() - static_cast<int>(p / npx) * ps;
            for (std::size_t id : this->subsurfaces[p])
                vis_ids.push_back(static_cast<uint8_t>(id));
        }
        vis_off[patches.size()] = static_cast<uint32_t>(vis_ids.size());
        if (vis_ids.empty())
            vis_ids.push_back(0);
        gpu.check(smvsb_set_surface(gpu.get(), this->surface->get_scale(),
            npx, npy, start_x, start_y, node_values.data(), node_valid.data(),
            patch_valid.data(), vis_off.data(), vis_ids.data()));

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

        /* ---- :318-356, reference code ---------------------------------- */
        if (converged)
            break;
        for (int del = std::numeric_limits<int>::max(); del > 10;)
            del = this->cut_boundaries();
        if (!this->opts.use_sgm)
        {
            this->surface->expand();
            this->create_subview_surfaces();
            for (int del = std::numeric_limits<int>::max(); del > 10;)
                del = this->cut_boundaries();
        }
        this->surface->remove_isolated_patches();
        int const patches_after = count_patches(this->surface);
        double const change = 1.0
            - static_cast<double>(std::min(patches_after, patches_before))
            / static_cast<double>(std::max(patches_after, patches_before));
        if (iter > 0 && (patches_after <= patches_before
            || change < 0.05 * this->surface->get_scale()))
            converged = true;
    }
}

SMVS_NAMESPACE_END
