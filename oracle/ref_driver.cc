/*
 * oracle/ref_driver.cc -- C entry points over the UNMODIFIED reference.
 *
 * TEST INFRASTRUCTURE ONLY. This file is linked with the reference's own
 * hot-path translation units, compiled verbatim from /root/reference/lib
 * against the header-only MVE shim in oracle/mve_shim (see oracle/Makefile),
 * into oracle/_ref/libsmvs_ref.so. Only tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py may load it. Nothing
 * in the product path (smvs_b200/, include/) may.
 *
 * It exposes the reference classes (StereoView, Surface, GaussNewtonStep,
 * ConjugateGradient, DepthOptimizer, LightOptimizer, SGMStereo) to ctypes
 * with plain arrays. Private members are reached by compiling THIS file
 * (only) with `private` spelled `public`; access specifiers do not change
 * the object layout, so the verbatim objects stay ABI compatible.
 *
 * The only control flow restated here is the instrumented inner Newton loop
 * (ref_newton_loop), which follows lib/depth_optimizer.cc:197-304 line by
 * line and calls the reference's own functions for every numeric step.
 */

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#define private public
#define protected public
#include "stereo_view.h"
#include "surface.h"
#include "gauss_newton_step.h"
#include "conjugate_gradient.h"
#include "depth_optimizer.h"
#include "light_optimizer.h"
#include "global_lighting.h"
#include "sgm_stereo.h"
#include "correspondence.h"
#include "bicubic_patch.h"
#include "surface_derivative.h"
#include "spherical_harmonics.h"
#include "ldl_decomposition.h"
#include "mesh_generator.h"
#include "mve/mesh_tools.h"
#include "mve/mesh_info.h"
#include "depth_triangulator.h"
#include "sse_vector.h"
#include "block_sparse_matrix.h"
#undef private
#undef protected

namespace {

typedef std::chrono::high_resolution_clock Clock;

double
seconds_since (Clock::time_point const& t0)
{
    return std::chrono::duration<double>(Clock::now() - t0).count();
}

struct RefScene
{
    std::vector<mve::View::Ptr> views;
    smvs::StereoView::Ptr main_view;
    std::vector<smvs::StereoView::Ptr> sub_views;
    smvs::DepthOptimizer::Options opts;
    std::unique_ptr<smvs::DepthOptimizer> optimizer;

    /* last constructed linear system (ref_gn_construct) */
    smvs::GaussNewtonStep::SparseMatrix hessian;
    smvs::GaussNewtonStep::SparseMatrix precond;
    smvs::GaussNewtonStep::DenseVector gradient;
};

void
silence_cout (bool on)
{
    static std::streambuf* saved = nullptr;
    static std::ostringstream sink;
    if (on && saved == nullptr)
        saved = std::cout.rdbuf(sink.rdbuf());
    else if (!on && saved != nullptr)
    {
        std::cout.rdbuf(saved);
        saved = nullptr;
        sink.str("");
    }
}

int
sampling_for_scale (int scale)
{
    /* lib/gauss_newton_step.cc:157-161 */
    int sampling = 4;
    if (scale < 5) sampling = 2;
    if (scale < 3) sampling = 1;
    return sampling;
}

} /* namespace */

/* lib/mesh_generator.cc is compiled whole; of it only cut_depth_maps runs.
 * What generate_mesh would call of MVE's and the reference's meshing code is
 * satisfied with definitions that refuse to run. */
namespace mve {
namespace geom {
void mesh_merge (TriangleMesh::ConstPtr, TriangleMesh::Ptr)
{ throw std::logic_error("oracle: meshing is not part of the hot path"); }
void depthmap_mesh_confidences (TriangleMesh::Ptr, int)
{ throw std::logic_error("oracle: meshing is not part of the hot path"); }
}
MeshInfo::MeshInfo (TriangleMesh::ConstPtr)
{ throw std::logic_error("oracle: meshing is not part of the hot path"); }
}
namespace smvs {
mve::TriangleMesh::Ptr DepthTriangulator::full_triangulation (void)
{ throw std::logic_error("oracle: meshing is not part of the hot path"); }
mve::TriangleMesh::Ptr DepthTriangulator::approximate_triangulation (int,
    double)
{ throw std::logic_error("oracle: meshing is not part of the hot path"); }
}

extern "C" {

/* ------------------------------------------------------------------ */
/* Scene                                                              */
/* ------------------------------------------------------------------ */

/* View 0 is the reference (main) view, views 1.. are the neighbours. */
void*
ref_scene_create (int n_views, int const* w, int const* h, int const* ch,
    uint8_t const* const* img, float const* flen, float const* rot,
    float const* trans, int init_linear)
{
    RefScene* s = new RefScene();
    for (int v = 0; v < n_views; ++v)
    {
        mve::View::Ptr view = mve::View::create();
        view->set_id(v);
        mve::CameraInfo cam;
        cam.flen = flen[v];
        std::copy(rot + 9 * v, rot + 9 * v + 9, cam.rot);
        std::copy(trans + 3 * v, trans + 3 * v + 3, cam.trans);
        view->set_camera(cam);
        mve::ByteImage::Ptr image = mve::ByteImage::create(w[v], h[v], ch[v]);
        std::copy(img[v], img[v] + (std::size_t)w[v] * h[v] * ch[v],
            image->begin());
        view->set_image(image, "undistorted");
        s->views.push_back(view);
        if (v == 0)
            s->main_view = smvs::StereoView::create(view, "undistorted",
                init_linear != 0, false);
        else
            s->sub_views.push_back(
                smvs::StereoView::create(view, "undistorted"));
    }
    s->opts.use_sgm = true;
    s->optimizer.reset(new smvs::DepthOptimizer(s->main_view, s->sub_views,
        mve::Bundle::ConstPtr(), s->opts));
    return s;
}

void
ref_scene_destroy (void* scene)
{
    delete static_cast<RefScene*>(scene);
}

/* lib/stereo_view.cc:24-46 on every view. */
void
ref_scene_set_scale (void* scene, int scale)
{
    RefScene* s = static_cast<RefScene*>(scene);
    s->main_view->set_scale(scale);
    for (auto& v : s->sub_views)
        v->set_scale(scale);
}

static smvs::StereoView::Ptr
view_of (RefScene* s, int v)
{
    return v == 0 ? s->main_view : s->sub_views[v - 1];
}

void
ref_view_get_gradients (void* scene, int v, float* out)
{
    mve::FloatImage::ConstPtr g = view_of(static_cast<RefScene*>(scene), v)
        ->get_image_gradients();
    std::copy(g->begin(), g->end(), out);
}

void
ref_view_get_hessian (void* scene, int v, float* out)
{
    mve::FloatImage::ConstPtr g = view_of(static_cast<RefScene*>(scene), v)
        ->get_image_hessian();
    std::copy(g->begin(), g->end(), out);
}

void
ref_view_get_scaleimage (void* scene, int v, float* out)
{
    mve::FloatImage::ConstPtr g = view_of(static_cast<RefScene*>(scene), v)
        ->get_scaleimage();
    std::copy(g->begin(), g->end(), out);
}

/* Shading image (1 ch) and its gradient (2 ch) of the main view. */
int
ref_view_get_shading (void* scene, float* img, float* grad)
{
    RefScene* s = static_cast<RefScene*>(scene);
    if (s->main_view->get_shading_image() == nullptr)
        return -1;
    mve::FloatImage::ConstPtr a = s->main_view->get_shading_image();
    mve::FloatImage::ConstPtr b = s->main_view->get_shading_gradients();
    std::copy(a->begin(), a->end(), img);
    std::copy(b->begin(), b->end(), grad);
    return 0;
}

/* Inject prepared gradient (2 ch) / Hessian (3 ch) images into view v so
 * that both arms of a comparison consume bit-identical inputs. */
void
ref_view_set_arrays (void* scene, int v, float const* grad, float const* hess)
{
    smvs::StereoView::Ptr view = view_of(static_cast<RefScene*>(scene), v);
    int const w = view->get_width(), h = view->get_height();
    view->image_grad = mve::FloatImage::create(w, h, 2);
    std::copy(grad, grad + (std::size_t)w * h * 2, view->image_grad->begin());
    view->image_hessian = mve::FloatImage::create(w, h, 3);
    if (hess != nullptr)
        std::copy(hess, hess + (std::size_t)w * h * 3,
            view->image_hessian->begin());
}

void
ref_view_set_shading (void* scene, float const* img, float const* grad)
{
    RefScene* s = static_cast<RefScene*>(scene);
    int const w = s->main_view->get_width(), h = s->main_view->get_height();
    s->main_view->shading = mve::FloatImage::create(w, h, 1);
    std::copy(img, img + (std::size_t)w * h, s->main_view->shading->begin());
    s->main_view->shading_grad = mve::FloatImage::create(w, h, 2);
    std::copy(grad, grad + (std::size_t)w * h * 2,
        s->main_view->shading_grad->begin());
}

float
ref_view_get_flen (void* scene, int v)
{
    return view_of(static_cast<RefScene*>(scene), v)->get_flen();
}

float
ref_view_get_inverse_flen (void* scene, int v)
{
    return view_of(static_cast<RefScene*>(scene), v)->get_inverse_flen();
}

/* Mi (n_sub x 9, row-major) and ti (n_sub x 3): fp32 reprojection widened to
 * double, lib/depth_optimizer.cc:679-699. */
void
ref_scene_get_Mt (void* scene, double* Mi, double* ti)
{
    RefScene* s = static_cast<RefScene*>(scene);
    for (std::size_t i = 0; i < s->sub_views.size(); ++i)
    {
        for (int j = 0; j < 9; ++j) Mi[i * 9 + j] = s->optimizer->Mi[i][j];
        for (int j = 0; j < 3; ++j) ti[i * 3 + j] = s->optimizer->ti[i][j];
    }
}

/* ------------------------------------------------------------------ */
/* Surface                                                            */
/* ------------------------------------------------------------------ */

/* Surface::create(nullptr, main, scale, init_depth), lib/surface.cc:19-53. */
void
ref_surface_create (void* scene, int scale, float const* init_depth)
{
    RefScene* s = static_cast<RefScene*>(scene);
    int const w = s->main_view->get_width(), h = s->main_view->get_height();
    mve::FloatImage::Ptr init = mve::FloatImage::create(w, h, 1);
    std::copy(init_depth, init_depth + (std::size_t)w * h, init->begin());
    s->optimizer->surface = smvs::Surface::create(mve::Bundle::ConstPtr(),
        s->main_view, scale, init);
    s->optimizer->sgm_depth = init;
    s->optimizer->subsurfaces.clear();
}

/* info[6] = scale, num_patches_x, num_patches_y, start_x, start_y, patchsize */
void
ref_surface_info (void* scene, int* info)
{
    smvs::Surface::Ptr sf = static_cast<RefScene*>(scene)->optimizer->surface;
    info[0] = sf->scale;
    info[1] = sf->num_patches_x;
    info[2] = sf->num_patches_y;
    info[3] = sf->pixel_start_x;
    info[4] = sf->pixel_start_y;
    info[5] = sf->patchsize;
}

void
ref_surface_get (void* scene, double* nodes, uint8_t* node_valid,
    uint8_t* patch_valid)
{
    smvs::Surface::Ptr sf = static_cast<RefScene*>(scene)->optimizer->surface;
    for (std::size_t i = 0; i < sf->nodes.size(); ++i)
    {
        smvs::Surface::Node::Ptr n = sf->nodes[i];
        node_valid[i] = (n != nullptr);
        nodes[i * 4 + 0] = n ? n->f : 0.0;
        nodes[i * 4 + 1] = n ? n->dx : 0.0;
        nodes[i * 4 + 2] = n ? n->dy : 0.0;
        nodes[i * 4 + 3] = n ? n->dxy : 0.0;
    }
    for (std::size_t i = 0; i < sf->patches.size(); ++i)
        patch_valid[i] = (sf->patches[i] != nullptr);
}

/* Overwrite the surface with the given node values / validity masks on the
 * grid the surface already has. */
void
ref_surface_set (void* scene, double const* nodes, uint8_t const* node_valid,
    uint8_t const* patch_valid)
{
    smvs::Surface::Ptr sf = static_cast<RefScene*>(scene)->optimizer->surface;
    for (std::size_t i = 0; i < sf->nodes.size(); ++i)
    {
        if (!node_valid[i])
        {
            sf->nodes[i].reset();
            continue;
        }
        if (sf->nodes[i] == nullptr)
            sf->nodes[i] = smvs::Surface::Node::create();
        sf->nodes[i]->f = nodes[i * 4 + 0];
        sf->nodes[i]->dx = nodes[i * 4 + 1];
        sf->nodes[i]->dy = nodes[i * 4 + 2];
        sf->nodes[i]->dxy = nodes[i * 4 + 3];
    }
    for (std::size_t i = 0; i < sf->patches.size(); ++i)
    {
        std::size_t const idx = i % sf->num_patches_x;
        std::size_t const idy = i / sf->num_patches_x;
        sf->patches[i].reset();
        if (patch_valid[i])
            sf->create_patch(idx, idy);
    }
}

void
ref_surface_subdivide (void* scene)
{
    static_cast<RefScene*>(scene)->optimizer->surface->subdivide_patches();
}

void
ref_surface_fill_from_depth (void* scene)
{
    static_cast<RefScene*>(scene)->optimizer->surface
        ->fill_patches_from_depth();
}

int
ref_surface_expand (void* scene)
{
    return static_cast<RefScene*>(scene)->optimizer->surface->expand();
}

void
ref_surface_remove_isolated (void* scene)
{
    static_cast<RefScene*>(scene)->optimizer->surface
        ->remove_isolated_patches();
}

/* Surface::get_depth_map, lib/surface.cc:155-168. */
void
ref_surface_get_depth (void* scene, float* depth)
{
    mve::FloatImage::Ptr d =
        static_cast<RefScene*>(scene)->optimizer->surface->get_depth_map();
    std::copy(d->begin(), d->end(), depth);
}

void
ref_surface_get_normals (void* scene, float* normals)
{
    RefScene* s = static_cast<RefScene*>(scene);
    mve::FloatImage::Ptr d = s->optimizer->surface->get_normal_map(
        s->main_view->get_inverse_flen());
    std::copy(d->begin(), d->end(), normals);
}

/* node_derivatives table, lib/gauss_newton_step.cc:43-51: ps^2 x 96. */
void
ref_node_derivative_table (void* scene, double* table)
{
    smvs::Surface::Ptr sf = static_cast<RefScene*>(scene)->optimizer->surface;
    int const n = sf->get_patchsize() * sf->get_patchsize();
    for (int i = 0; i < n; ++i)
        sf->fill_node_derivatives_for_pixel(i, table + i * 96,
            table + i * 96 + 24, table + i * 96 + 48, table + i * 96 + 72);
}

/* ------------------------------------------------------------------ */
/* Visibility                                                         */
/* ------------------------------------------------------------------ */

/* create_subview_surfaces + cut_boundaries until < 10 deletions,
 * lib/depth_optimizer.cc:189-195. Returns number of valid patches. */
int
ref_compute_visibility (void* scene)
{
    RefScene* s = static_cast<RefScene*>(scene);
    s->optimizer->main_gradients = s->main_view->get_image_gradients();
    silence_cout(true);
    s->optimizer->create_subview_surfaces();
    int deleted = std::numeric_limits<int>::max();
    while (deleted > 10)
        deleted = s->optimizer->cut_boundaries();
    silence_cout(false);
    int valid = 0;
    for (auto const& p : s->optimizer->surface->get_patches())
        valid += (p != nullptr);
    return valid;
}

/* DepthOptimizer::create_subview_surfaces alone (lib/depth_optimizer.cc:433-604);
 * returns the number of patches left. */
int
ref_create_subview_surfaces (void* scene, int use_sgm)
{
    RefScene* s = static_cast<RefScene*>(scene);
    s->opts.use_sgm = (use_sgm != 0);
    s->optimizer->main_gradients = s->main_view->get_image_gradients();
    silence_cout(true);
    s->optimizer->create_subview_surfaces();
    silence_cout(false);
    int valid = 0;
    for (auto const& p : s->optimizer->surface->get_patches())
        valid += (p != nullptr);
    return valid;
}

/* One DepthOptimizer::cut_boundaries() (lib/depth_optimizer.cc:360-431);
 * returns its return value (patches deleted). */
int
ref_cut_boundaries (void* scene)
{
    RefScene* s = static_cast<RefScene*>(scene);
    s->optimizer->main_gradients = s->main_view->get_image_gradients();
    return s->optimizer->cut_boundaries();
}

/* The depth image create_subview_surfaces adds to the z-buffers in the
 * use_sgm mode (DepthOptimizer::sgm_depth). */
void
ref_set_sgm_depth (void* scene, float const* depth)
{
    RefScene* s = static_cast<RefScene*>(scene);
    int const w = s->main_view->get_width(), h = s->main_view->get_height();
    mve::FloatImage::Ptr img = mve::FloatImage::create(w, h, 1);
    std::copy(depth, depth + (std::size_t)w * h, img->begin());
    s->optimizer->sgm_depth = img;
}

/* The matrix cut_boundaries builds at lib/depth_optimizer.cc:377-379. */
void
ref_main_inverse_calibration (void* scene, float* out9)
{
    RefScene* s = static_cast<RefScene*>(scene);
    s->main_view->get_camera().fill_inverse_calibration(out9,
        s->main_view->get_width(), s->main_view->get_height());
}

/* DepthOptimizer::depthmap_bilateral_filter(dm, main_view->get_image()) with
 * its default sigma = 5, kernel_size = 5 (lib/depth_optimizer.cc:42-43,
 * 957-1004). dm: dm_w x dm_h; out: main view size. */
void
ref_bilateral_filter (void* scene, float const* dm, int dm_w, int dm_h,
    float* out)
{
    RefScene* s = static_cast<RefScene*>(scene);
    mve::FloatImage::Ptr d = mve::FloatImage::create(dm_w, dm_h, 1);
    std::copy(dm, dm + (std::size_t)dm_w * dm_h, d->begin());
    mve::FloatImage::Ptr r = s->optimizer->depthmap_bilateral_filter(d,
        s->main_view->get_image());
    std::copy(r->begin(), r->end(), out);
}

/* StereoView::get_image(): the unblurred float image; returns its channels. */
int
ref_view_get_image (void* scene, int v, float* out)
{
    mve::FloatImage::ConstPtr img = view_of(static_cast<RefScene*>(scene), v)
        ->get_image();
    if (out != nullptr)
        std::copy(img->begin(), img->end(), out);
    return img->channels();
}

/* vis_off has num_patches + 1 entries; pass vis_ids = NULL to query size. */
uint64_t
ref_get_visibility (void* scene, uint32_t* vis_off, uint8_t* vis_ids)
{
    RefScene* s = static_cast<RefScene*>(scene);
    std::size_t const np = s->optimizer->surface->get_patches().size();
    uint64_t total = 0;
    for (std::size_t p = 0; p < np; ++p)
    {
        if (vis_off != nullptr) vis_off[p] = static_cast<uint32_t>(total);
        if (p < s->optimizer->subsurfaces.size())
            for (std::size_t id : s->optimizer->subsurfaces[p])
            {
                if (vis_ids != nullptr) vis_ids[total] = (uint8_t)id;
                total += 1;
            }
    }
    if (vis_off != nullptr) vis_off[np] = static_cast<uint32_t>(total);
    return total;
}

void
ref_set_visibility (void* scene, uint32_t const* vis_off,
    uint8_t const* vis_ids)
{
    RefScene* s = static_cast<RefScene*>(scene);
    std::size_t const np = s->optimizer->surface->get_patches().size();
    s->optimizer->subsurfaces.assign(np, std::vector<std::size_t>());
    for (std::size_t p = 0; p < np; ++p)
        for (uint32_t k = vis_off[p]; k < vis_off[p + 1]; ++k)
            s->optimizer->subsurfaces[p].push_back(vis_ids[k]);
}

/* ------------------------------------------------------------------ */
/* Gauss-Newton step, CG, update                                      */
/* ------------------------------------------------------------------ */

static smvs::GlobalLighting::Ptr
lighting_from (double const* light16)
{
    if (light16 == nullptr)
        return nullptr;
    smvs::GlobalLighting::Params p;
    for (int i = 0; i < 16; ++i)
        p[i] = light16[i];
    return smvs::GlobalLighting::create(p);
}

/* GaussNewtonStep::construct, lib/gauss_newton_step.cc:33-143. The system is
 * kept in the scene; returns the number of Hessian blocks. */
int64_t
ref_gn_construct (void* scene, uint8_t const* active_nodes,
    double const* light16, double regularization,
    double light_surf_regularization)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::GaussNewtonStep::Options o;
    o.regularization = regularization;
    o.light_surf_regularization = light_surf_regularization;
    smvs::GaussNewtonStep step(o, s->main_view, s->sub_views,
        s->optimizer->Mi, s->optimizer->ti);
    std::size_t const nn = s->optimizer->surface->get_nodes().size();
    std::vector<char> active(active_nodes, active_nodes + nn);
    step.construct(s->optimizer->surface, s->optimizer->subsurfaces, active,
        lighting_from(light16), &s->hessian, &s->gradient, &s->precond);
    return static_cast<int64_t>(s->hessian.num_non_zero());
}

/* sizes[3] = num_params, nnzb(H), nnzb(P) */
void
ref_get_system_sizes (void* scene, uint64_t* sizes)
{
    RefScene* s = static_cast<RefScene*>(scene);
    sizes[0] = s->gradient.size();
    sizes[1] = s->hessian.num_non_zero();
    sizes[2] = s->precond.num_non_zero();
}

/* BSC arrays exactly as lib/block_sparse_matrix.h:95-97 holds them: blocks
 * sorted by column block, each block row-major 4x4, inner = 4 * block row. */
void
ref_get_system (void* scene, double* g, double* Hvals, uint64_t* Houter,
    uint64_t* Hinner, double* Pvals, uint64_t* Pouter, uint64_t* Pinner)
{
    RefScene* s = static_cast<RefScene*>(scene);
    if (g) std::copy(s->gradient.begin(), s->gradient.end(), g);
    if (Hvals)
        for (std::size_t i = 0; i < s->hessian.values.size(); ++i)
            std::copy(s->hessian.values[i].begin(),
                s->hessian.values[i].end(), Hvals + 16 * i);
    if (Houter) std::copy(s->hessian.outer.begin(), s->hessian.outer.end(),
        Houter);
    if (Hinner) std::copy(s->hessian.inner.begin(), s->hessian.inner.end(),
        Hinner);
    if (Pvals)
        for (std::size_t i = 0; i < s->precond.values.size(); ++i)
            std::copy(s->precond.values[i].begin(),
                s->precond.values[i].end(), Pvals + 16 * i);
    if (Pouter) std::copy(s->precond.outer.begin(), s->precond.outer.end(),
        Pouter);
    if (Pinner) std::copy(s->precond.inner.begin(), s->precond.inner.end(),
        Pinner);
}

/* y = H * x with BlockSparseMatrix::multiply. */
void
ref_hessian_multiply (void* scene, double const* x, double* y)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::SSEVector v(s->hessian.num_cols());
    std::copy(x, x + v.size(), v.begin());
    smvs::SSEVector r = s->hessian.multiply(v);
    std::copy(r.begin(), r.end(), y);
}

/* ConjugateGradient::solve(H, -g, &x, &P), lib/conjugate_gradient.h:72-202,
 * with the options of lib/depth_optimizer.cc:245-254 when err_tol < 0
 * (error_tolerance = 0.01 * ||g||). */
int
ref_cg_solve (void* scene, int max_iter, double err_tol, double q_tol,
    double* x_out, int* iters, int* info)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::ConjugateGradient::Options o;
    o.max_iterations = max_iter;
    o.error_tolerance = (err_tol < 0.0 ? s->gradient.norm() * 0.01 : err_tol);
    o.q_tolerance = q_tol;
    smvs::ConjugateGradient cg(o);
    smvs::SSEVector b = s->gradient;
    b.negate_self();
    smvs::SSEVector x;
    smvs::ConjugateGradient::Status st = cg.solve(s->hessian, b, &x,
        &s->precond);
    std::copy(x.begin(), x.end(), x_out);
    *iters = st.num_iterations;
    *info = static_cast<int>(st.info);
    return 0;
}

/* lib/depth_optimizer.cc:271-303: reprojections before/after
 * Surface::update_nodes(delta), then the new active set (or, with full_opt,
 * the mean shift). active_inout is read (old set) and overwritten. */
int
ref_update_nodes (void* scene, double const* delta, double reproj_thresh,
    int full_opt, uint8_t* active_inout, uint64_t* n_active,
    double* mean_shift)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::Surface::Ptr sf = s->optimizer->surface;
    std::size_t const nn = sf->get_nodes().size();
    std::vector<char> active(active_inout, active_inout + nn);
    std::vector<double> d(delta, delta + nn * 4);
    std::vector<double> depth_updates;
    std::vector<std::pair<std::size_t, math::Vec2d>> p1, p2;
    s->optimizer->fill_node_reprojections(active, &p1);
    sf->update_nodes(d, &depth_updates);
    s->optimizer->fill_node_reprojections(active, &p2);

    double sum_diff = 0.0;
    for (std::size_t p = 0; p < p1.size(); ++p)
        sum_diff += (p1[p].second - p2[p].second).norm();
    if (mean_shift)
        *mean_shift = sum_diff / (double)p1.size();
    if (!full_opt)
    {
        std::fill(active.begin(), active.end(), 0);
        for (std::size_t p = 0; p < p1.size(); ++p)
        {
            double diff = (p1[p].second - p2[p].second).norm();
            if (diff > reproj_thresh)
                active[p1[p].first] = 1;
        }
    }
    uint64_t cnt = 0;
    for (std::size_t i = 0; i < nn; ++i)
    {
        active_inout[i] = active[i];
        cnt += (active[i] == 1);
    }
    if (n_active) *n_active = cnt;
    return 0;
}

/* Instrumented replica of the inner Newton loop,
 * lib/depth_optimizer.cc:197-304 (no full_optimization branch).
 * stats[8] = newton steps, sum CG iterations, pixel-iterations,
 *            construct s, solve s, update s, final active nodes, nan flag */
int
ref_newton_loop (void* scene, double const* light16, double regularization,
    double light_surf_regularization, int max_steps, double* stats)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::Surface::Ptr sf = s->optimizer->surface;
    smvs::GaussNewtonStep::Options o;
    o.regularization = regularization;
    o.light_surf_regularization = light_surf_regularization;
    smvs::GaussNewtonStep step(o, s->main_view, s->sub_views,
        s->optimizer->Mi, s->optimizer->ti);
    smvs::GlobalLighting::Ptr lighting = lighting_from(light16);

    std::size_t num_initial_active_nodes = 0;
    std::vector<char> active_nodes(sf->get_nodes().size(), 0);
    for (std::size_t i = 0; i < sf->get_nodes().size(); ++i)
        if (sf->get_nodes()[i] != nullptr)
        {
            active_nodes[i] = 1;
            num_initial_active_nodes += 1;
        }
    std::size_t num_active_nodes = num_initial_active_nodes;

    int const ps = sf->get_patchsize();
    int const sampling = sampling_for_scale(sf->get_scale());
    double const samples_per_patch = double(ps * ps) / (sampling * sampling);

    unsigned int newton_step = 0;
    double cg_iters = 0, pixiters = 0, t_build = 0, t_solve = 0, t_upd = 0;
    double nan_flag = 0;
    std::vector<double> delta, depth_updates;
    std::vector<std::pair<std::size_t, math::Vec2d>> p1, p2;

    for (; newton_step < (unsigned)max_steps
        && num_active_nodes > num_initial_active_nodes / 20;)
    {
        newton_step += 1;
        for (std::size_t p = 0; p < sf->get_patches().size(); ++p)
        {
            if (sf->get_patches()[p] == nullptr) continue;
            std::size_t ids[4];
            sf->fill_node_ids_for_patch(p, ids);
            if (active_nodes[ids[0]] || active_nodes[ids[1]]
                || active_nodes[ids[2]] || active_nodes[ids[3]])
                pixiters += samples_per_patch;
        }

        Clock::time_point t0 = Clock::now();
        step.construct(sf, s->optimizer->subsurfaces, active_nodes, lighting,
            &s->hessian, &s->gradient, &s->precond);
        t_build += seconds_since(t0);

        smvs::ConjugateGradient::Options cg_opts;
        cg_opts.max_iterations = 200;
        cg_opts.error_tolerance = s->gradient.norm() * 0.01;
        smvs::ConjugateGradient cg_solver(cg_opts);
        smvs::SSEVector x;
        s->gradient.negate_self();
        t0 = Clock::now();
        smvs::ConjugateGradient::Status st = cg_solver.solve(s->hessian,
            s->gradient, &x, &s->precond);
        t_solve += seconds_since(t0);
        cg_iters += st.num_iterations;

        delta.resize(x.size());
        std::copy(x.begin(), x.end(), delta.begin());
        if (std::isnan(delta[0]))
        {
            nan_flag = 1;
            break;
        }

        t0 = Clock::now();
        s->optimizer->fill_node_reprojections(active_nodes, &p1);
        sf->update_nodes(delta, &depth_updates);
        s->optimizer->fill_node_reprojections(active_nodes, &p2);
        std::fill(active_nodes.begin(), active_nodes.end(), 0);
        for (std::size_t p = 0; p < p1.size(); ++p)
        {
            double diff = (p1[p].second - p2[p].second).norm();
            if (diff > 0.15)
                active_nodes[p1[p].first] = 1;
        }
        num_active_nodes = 0;
        for (auto& node : active_nodes)
            if (node == 1)
                num_active_nodes += 1;
        t_upd += seconds_since(t0);
    }
    stats[0] = newton_step; stats[1] = cg_iters; stats[2] = pixiters;
    stats[3] = t_build; stats[4] = t_solve; stats[5] = t_upd;
    stats[6] = (double)num_active_nodes; stats[7] = nan_flag;
    return 0;
}

/* LightOptimizer::fit_lighting_to_image, lib/light_optimizer.cc:22-55. */
int
ref_fit_lighting (void* scene, double* params16)
{
    RefScene* s = static_cast<RefScene*>(scene);
    if (s->main_view->get_shading_image() == nullptr)
        return -1;
    smvs::LightOptimizer lo(s->optimizer->surface, s->main_view);
    smvs::GlobalLighting::Ptr l = lo.fit_lighting_to_image(
        s->main_view->get_shading_image());
    smvs::GlobalLighting::Params p = l->get_parameters();
    for (int i = 0; i < 16; ++i)
        params16[i] = p[i];
    return 0;
}

/* What StereoView::get_sgm_depth() returns for a depth written as the
 * "smvs-sgm" embedding (the two convention changes cost fp32 rounding). */
int
ref_sgm_roundtrip (void* scene, float const* sgm_depth, float* out)
{
    RefScene* s = static_cast<RefScene*>(scene);
    int const w = s->main_view->get_width(), h = s->main_view->get_height();
    mve::FloatImage::Ptr init = mve::FloatImage::create(w, h, 1);
    std::copy(sgm_depth, sgm_depth + (std::size_t)w * h, init->begin());
    s->main_view->write_depth_to_view(init, "smvs-sgm");
    mve::FloatImage::Ptr back = s->main_view->get_sgm_depth();
    std::copy(back->begin(), back->end(), out);
    return 0;
}

/* Full DepthOptimizer::optimize(), lib/depth_optimizer.cc:53-162, with the
 * options the CLI sets (app/smvsrecon.cc:711-720). The initial depth goes in
 * as the "smvs-sgm" embedding (z-depth; stored in MVE convention so that
 * StereoView::get_sgm_depth returns it unchanged up to fp32 rounding). */
int
ref_optimize (void* scene, float const* sgm_depth, double regularization,
    int num_iterations, int min_scale, int use_shading, int debug_lvl,
    float* depth_out, float* normals_out, double* light16_out)
{
    RefScene* s = static_cast<RefScene*>(scene);
    int const w = s->main_view->get_width(), h = s->main_view->get_height();
    mve::FloatImage::Ptr init = mve::FloatImage::create(w, h, 1);
    std::copy(sgm_depth, sgm_depth + (std::size_t)w * h, init->begin());
    s->main_view->write_depth_to_view(init, "smvs-sgm");

    s->opts.regularization = regularization;
    s->opts.num_iterations = num_iterations;
    s->opts.min_scale = min_scale;
    s->opts.use_shading = (use_shading != 0);
    s->opts.use_sgm = true;
    s->opts.debug_lvl = debug_lvl;
    s->opts.output_name = "smvs-out";
    s->optimizer.reset(new smvs::DepthOptimizer(s->main_view, s->sub_views,
        mve::Bundle::ConstPtr(), s->opts));
    if (debug_lvl == 0) silence_cout(true);
    s->optimizer->optimize();
    silence_cout(false);
    if (depth_out)
    {
        mve::FloatImage::Ptr d = s->optimizer->surface->get_depth_map();
        std::copy(d->begin(), d->end(), depth_out);
    }
    if (normals_out)
    {
        mve::FloatImage::Ptr n = s->optimizer->get_normals();
        std::copy(n->begin(), n->end(), normals_out);
    }
    if (light16_out && s->optimizer->lighting != nullptr)
    {
        smvs::GlobalLighting::Params p =
            s->optimizer->lighting->get_parameters();
        for (int i = 0; i < 16; ++i)
            light16_out[i] = p[i];
    }
    return 0;
}

/* DepthOptimizer::optimize() with use_sgm = false (app/smvsrecon.cc's
 * --no-sgm): the initial surface comes from the bundle's features
 * (lib/surface.cc:43-46, 91-128). n_feat points `pos3` (world coordinates),
 * each observed by the main view. sparse_out (w*h, may be NULL) receives the
 * depth image Surface::create makes of them -- what the resident GPU path is
 * given as its initial depth. */
int
ref_optimize_nosgm (void* scene, int n_feat, float const* pos3,
    double regularization, int num_iterations, int min_scale,
    float* sparse_out, float* depth_out, float* normals_out)
{
    RefScene* s = static_cast<RefScene*>(scene);
    mve::Bundle::Ptr bundle = mve::Bundle::create();
    for (int i = 0; i < n_feat; ++i)
    {
        mve::Bundle::Feature3D f;
        for (int k = 0; k < 3; ++k)
        {
            f.pos[k] = pos3[3 * i + k];
            f.color[k] = 0.5f;
        }
        mve::Bundle::Feature2D r;
        r.view_id = s->main_view->get_view_id();
        r.feature_id = i;
        r.pos[0] = r.pos[1] = 0.0f;
        f.refs.push_back(r);
        bundle->get_features().push_back(f);
    }
    if (sparse_out)
    {
        int const w = s->main_view->get_width(), h = s->main_view->get_height();
        int const init_scale = (int)std::max(std::ceil(std::log2(w * h / 1.7e6)
            / 2) + 4, 4.0);
        smvs::Surface::Ptr tmp = smvs::Surface::create(bundle, s->main_view,
            init_scale + 1);
        std::copy(tmp->depth->begin(), tmp->depth->end(), sparse_out);
    }
    s->opts.regularization = regularization;
    s->opts.num_iterations = num_iterations;
    s->opts.min_scale = min_scale;
    s->opts.use_shading = false;
    s->opts.use_sgm = false;
    s->opts.debug_lvl = 0;
    s->opts.output_name = "smvs-out";
    s->optimizer.reset(new smvs::DepthOptimizer(s->main_view, s->sub_views,
        bundle, s->opts));
    silence_cout(true);
    s->optimizer->optimize();
    silence_cout(false);
    s->opts.use_sgm = true;
    if (depth_out)
    {
        mve::FloatImage::Ptr d = s->optimizer->surface->get_depth_map();
        std::copy(d->begin(), d->end(), depth_out);
    }
    if (normals_out)
    {
        mve::FloatImage::Ptr n = s->optimizer->get_normals();
        std::copy(n->begin(), n->end(), normals_out);
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* SGM                                                                */
/* ------------------------------------------------------------------ */

/* SGMStereo(main = view a, neighbour = view b).run_sgm(min, max),
 * lib/sgm_stereo.cc:98-124; optionally dumps the cost volume and the
 * aggregated volume (uint16, pixel-major, num_steps contiguous).
 * times[3] = cost volume s, aggregation s, winner-takes-all s. */
int
ref_sgm_run (void* scene, int a, int b, int scale, int num_steps,
    float min_depth, float max_depth, int penalty1, int penalty2,
    float* depth_out, uint16_t* cost_out, uint16_t* sgm_out, double* times)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::SGMStereo::Options o;
    o.scale = scale;
    o.num_steps = num_steps;
    o.min_depth = min_depth;
    o.max_depth = max_depth;
    o.penalty1 = (uint16_t)penalty1;
    o.penalty2 = (uint16_t)penalty2;
    smvs::SGMStereo sgm(o, view_of(s, a), view_of(s, b));
    Clock::time_point t0 = Clock::now();
    sgm.create_cost_volume(min_depth, max_depth, num_steps);
    double const t_cost = seconds_since(t0);
    t0 = Clock::now();
    sgm.aggregate_sgm_costs();
    double const t_agg = seconds_since(t0);
    t0 = Clock::now();
    mve::FloatImage::Ptr d = sgm.depth_from_sgm_volume();
    double const t_wta = seconds_since(t0);
    if (depth_out) std::copy(d->begin(), d->end(), depth_out);
    if (cost_out) std::copy(sgm.sse_cost_volume.begin(),
        sgm.sse_cost_volume.end(), cost_out);
    if (sgm_out) std::copy(sgm.sse_sgm_volume.begin(),
        sgm.sse_sgm_volume.end(), sgm_out);
    if (times) { times[0] = t_cost; times[1] = t_agg; times[2] = t_wta; }
    return 0;
}

/* info[2] = width, height of the SGM working image of view v at `scale`. */
void
ref_sgm_dims (void* scene, int v, int scale, int* info)
{
    RefScene* s = static_cast<RefScene*>(scene);
    int w = view_of(s, v)->get_width(), h = view_of(s, v)->get_height();
    for (int i = 0; i < scale; ++i) { w = (w + 1) / 2; h = (h + 1) / 2; }
    info[0] = w; info[1] = h;
}

/* SGMStereo::reconstruct (both directions + consistency check),
 * lib/sgm_stereo.cc:46-96, with an explicit depth range. */
int
ref_sgm_reconstruct (void* scene, int a, int b, int scale, int num_steps,
    float min_depth, float max_depth, float* depth_out)
{
    RefScene* s = static_cast<RefScene*>(scene);
    smvs::SGMStereo::Options o;
    o.scale = scale;
    o.num_steps = num_steps;
    o.min_depth = min_depth;
    o.max_depth = max_depth;
    mve::FloatImage::Ptr d = smvs::SGMStereo::reconstruct(o, view_of(s, a),
        view_of(s, b), mve::Bundle::ConstPtr());
    std::copy(d->begin(), d->end(), depth_out);
    return 0;
}

/* fp32 reprojection view a -> view b at the given working sizes. */
void
ref_reprojection (void* scene, int a, int b, int aw, int ah, int bw, int bh,
    float* M, float* t)
{
    RefScene* s = static_cast<RefScene*>(scene);
    view_of(s, a)->get_camera().fill_reprojection(
        view_of(s, b)->get_camera(), aw, ah, bw, bh, M, t);
}

/* ------------------------------------------------------------------ */
/* Unit-level entry points for the reference's own known-answer tests  */
/* ------------------------------------------------------------------ */

/* BicubicPatch from 4 nodes x (f,dx,dy,dxy): out[6] = f,dx,dy,dxy,dxx,dyy. */
void
ref_bicubic_eval (double const* nodes16, double x, double y, double* out)
{
    smvs::BicubicPatch::Node::Ptr n[4];
    for (int i = 0; i < 4; ++i)
    {
        n[i] = smvs::BicubicPatch::Node::create();
        n[i]->f = nodes16[i * 4]; n[i]->dx = nodes16[i * 4 + 1];
        n[i]->dy = nodes16[i * 4 + 2]; n[i]->dxy = nodes16[i * 4 + 3];
    }
    smvs::BicubicPatch::Ptr p = smvs::BicubicPatch::create(n[0], n[1], n[2],
        n[3]);
    out[0] = p->evaluate_f(x, y); out[1] = p->evaluate_dx(x, y);
    out[2] = p->evaluate_dy(x, y); out[3] = p->evaluate_dxy(x, y);
    out[4] = p->evaluate_dxx(x, y); out[5] = p->evaluate_dyy(x, y);
}

/* out[96]: 4 nodes x 24, lib/bicubic_patch.cc:302-316 (+ patchsize). */
void
ref_bicubic_node_derivatives (double x, double y, double patchsize,
    double* out)
{
    if (patchsize > 0.0)
        smvs::BicubicPatch::node_derivatives_for_patchsize(x, y, patchsize,
            out, out + 24, out + 48, out + 72);
    else
        smvs::BicubicPatch::node_derivatives(x, y, out, out + 24, out + 48,
            out + 72);
}

/* Correspondence: proj[2], jac[4], c_dn[32], jac_dn[32] for given dn[96]. */
void
ref_correspondence (double const* M9, double const* t3, double u, double v,
    double w, double wx, double wy, double const* grad2, double const* dn96,
    double* proj, double* jac, double* c_dn, double* jac_dn, double* depth)
{
    math::Matrix3d M(M9);
    math::Vec3d t(t3);
    smvs::Correspondence C(M, t, u, v, w, wx, wy);
    C.fill(proj);
    C.fill_jacobian(jac);
    math::Vec2d a[16], b[16];
    C.fill_derivative(dn96, a);
    C.fill_jacobian_derivative_grad(grad2, dn96, b);
    for (int i = 0; i < 16; ++i)
    {
        c_dn[i * 2] = a[i][0]; c_dn[i * 2 + 1] = a[i][1];
        jac_dn[i * 2] = b[i][0]; jac_dn[i * 2 + 1] = b[i][1];
    }
    *depth = C.get_depth();
}

/* surfderiv: normal[3] (inv_flen = 1/f), div[6], div_deriv[96], ndiv[48]. */
void
ref_surface_derivatives (double const* dn96, double x, double y, double f,
    double w, double dx, double dy, double dxy, double dxx, double dyy,
    double* normal, double* div, double* div_deriv, double* normal_deriv)
{
    smvs::surfderiv::fill_normal(x, y, 1.0 / f, w, dx, dy, normal);
    smvs::surfderiv::normal_divergence(x, y, f, w, dx, dy, dxy, dxx, dyy, div);
    smvs::surfderiv::normal_divergence_deriv(dn96, x, y, f, w, dx, dy, dxy,
        dxx, dyy, div_deriv);
    smvs::surfderiv::normal_derivative(dn96, x, y, f, w, dx, dy,
        normal_deriv);
}

void
ref_sh_4band (double const* normal, double* sh16, double* deriv48)
{
    smvs::sh::evaluate_4_band(normal, sh16);
    smvs::sh::derivative_4_band(normal, deriv48);
}

void
ref_ldl_inverse (double* A, int n)
{
    smvs::ldl_inverse(A, n);
}

/* MeshGenerator::cut_depth_maps (lib/mesh_generator.cc:25-158) on n views
 * given by their cameras (flen, world-to-camera rotation and translation),
 * depth maps (MVE convention) and world-space normal maps. Also returns the
 * per-view matrices the function derives from the cameras, so that the device
 * twin consumes identical inputs: invproj (9), cam-to-world (16), KR (9), t (3).
 * lib/mesh_generator.cc is compiled verbatim; only cut_depth_maps is called. */
int
ref_cut_depth_maps (int n, int const* w, int const* h, float const* flen,
    float const* rot9, float const* trans3, float const* const* depth,
    float const* const* normals, float* const* depth_out, float* invproj9,
    float* ctw16, float* KR9, float* t3)
{
    smvs::MeshGenerator::Options o;
    o.num_threads = 4;
    smvs::MeshGenerator mg(o);
    std::vector<mve::FloatImage::Ptr> depthmaps(n), normalmaps(n);
    for (int i = 0; i < n; ++i)
    {
        mve::CameraInfo cam;
        cam.flen = flen[i];
        std::copy(rot9 + 9 * i, rot9 + 9 * i + 9, cam.rot);
        std::copy(trans3 + 3 * i, trans3 + 3 * i + 3, cam.trans);
        mve::View::Ptr view = mve::View::create();
        view->set_id(i);
        view->set_camera(cam);
        mg.views.push_back(view);
        mg.view_projs.emplace_back(cam, w[i], h[i]);
        depthmaps[i] = mve::FloatImage::create(w[i], h[i], 1);
        std::copy(depth[i], depth[i] + (std::size_t)w[i] * h[i],
            depthmaps[i]->begin());
        normalmaps[i] = mve::FloatImage::create(w[i], h[i], 3);
        std::copy(normals[i], normals[i] + (std::size_t)w[i] * h[i] * 3,
            normalmaps[i]->begin());
        if (invproj9)
            cam.fill_inverse_calibration(invproj9 + 9 * i, w[i], h[i]);
        if (ctw16)
            cam.fill_cam_to_world(ctw16 + 16 * i);
        if (KR9)
            std::copy(mg.view_projs[i].KR.begin(), mg.view_projs[i].KR.end(),
                KR9 + 9 * i);
        if (t3)
            std::copy(mg.view_projs[i].t.begin(), mg.view_projs[i].t.end(),
                t3 + 3 * i);
    }
    if (depth_out != nullptr)
    {
        mg.cut_depth_maps(&depthmaps, &normalmaps);
        for (int i = 0; i < n; ++i)
            std::copy(depthmaps[i]->begin(), depthmaps[i]->end(),
                depth_out[i]);
    }
    return 0;
}

/* SSEVector operations (lib/sse_vector.cc:19-205), for the reference's own
 * known answers (tests/gtest_matrix_vector.cc:33-195).
 * op: 0 dot (out[0]), 1 add, 2 subtract, 3 multiply (a * factor),
 *     4 multiply_add (a + b * factor), 5 multiply_sub (a - b * factor). */
void
ref_ssevector_op (int op, int n, double const* a, double const* b,
    double factor, double* out)
{
    smvs::SSEVector va(n), vb(n);
    for (int i = 0; i < n; ++i)
    {
        va[i] = a[i];
        vb[i] = b ? b[i] : 0.0;
    }
    if (op == 0)
    {
        out[0] = va.dot(vb);
        return;
    }
    smvs::SSEVector c;
    switch (op)
    {
    case 1: c = va.add(vb); break;
    case 2: c = va.subtract(vb); break;
    case 3: c = va.multiply(factor); break;
    case 4: c = va.multiply_add(vb, factor); break;
    default: c = va.multiply_sub(vb, factor); break;
    }
    for (int i = 0; i < n; ++i)
        out[i] = c[i];
}

/* BlockSparseMatrix<2> (lib/block_sparse_matrix.h) from blocks (row, col,
 * 4 row-major values each) or from scalar triplets; optionally
 * invert_blocks_inplace(); y = A x; returns num_non_zero()
 * (tests/gtest_matrix_vector.cc:197-356). */
int
ref_bsm2 (int dim, int n_blocks, int const* block_rc, double const* block_vals,
    int n_triplets, int const* trip_rc, double const* trip_vals, int invert,
    double const* x, double* y)
{
    typedef smvs::BlockSparseMatrix<2> BSMatrix;
    BSMatrix m(dim, dim);
    if (n_triplets > 0)
    {
        BSMatrix::Triplets trips;
        for (int i = 0; i < n_triplets; ++i)
            trips.emplace_back(trip_rc[2 * i], trip_rc[2 * i + 1],
                trip_vals[i]);
        m.set_from_triplets(trips);
    }
    else
    {
        BSMatrix::Blocks blocks;
        for (int i = 0; i < n_blocks; ++i)
            blocks.emplace_back(block_rc[2 * i], block_rc[2 * i + 1],
                block_vals + 4 * i);
        m.set_from_blocks(blocks);
    }
    if (invert)
        m.invert_blocks_inplace();
    if (x != nullptr && y != nullptr)
    {
        smvs::SSEVector vx(dim);
        for (int i = 0; i < dim; ++i)
            vx[i] = x[i];
        smvs::SSEVector r = m.multiply(vx);
        for (int i = 0; i < dim; ++i)
            y[i] = r[i];
    }
    return static_cast<int>(m.num_non_zero());
}

} /* extern "C" */
