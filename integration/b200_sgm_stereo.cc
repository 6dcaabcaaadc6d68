/*
 * integration/b200_sgm_stereo.cc
 *
 * Drop-in body for smvs::SGMStereo::run_sgm (reference:
 * lib/sgm_stereo.cc:98-124): cost volume, 8-path aggregation and
 * winner-takes-all on the GPU through smvsb_sgm. lib/sgm_stereo.h untouched.
 */
#include <iostream>
#include <stdexcept>
#include <string>

#include "sgm_stereo.h"

#include "b200_context.h"

SMVS_NAMESPACE_BEGIN

mve::FloatImage::Ptr
SGMStereo::run_sgm (float min_depth, float max_depth)
{
    /* lib/sgm_stereo.cc:153-160: reprojection at SGM working resolution */
    math::Matrix3f M;
    math::Vec3f t;
    mve::CameraInfo n_cam = this->neighbor->get_camera();
    this->main->get_camera().fill_reprojection(n_cam,
        this->main_image->width(), this->main_image->height(),
        this->neighbor_image->width(), this->neighbor_image->height(), *M, *t);

    mve::FloatImage::Ptr depth = mve::FloatImage::create(
        this->main_image->width(), this->main_image->height(), 1);
    int const rc = smvsb_sgm(smvs_b200_integration::thread_device(),
        this->main_image->width(),
        this->main_image->height(), this->main_image->begin(),
        this->neighbor_image->width(), this->neighbor_image->height(),
        this->neighbor_image->begin(), *M, *t, min_depth, max_depth,
        this->opts.num_steps, this->opts.penalty1, this->opts.penalty2,
        depth->begin(), nullptr, nullptr, nullptr);
    if (rc != SMVSB_OK)
        throw std::runtime_error(std::string("smvs_b200: ")
            + smvsb_last_error(nullptr));
    return depth;
}

/*
 * Drop-in body for the static SGMStereo::reconstruct (lib/sgm_stereo.cc:
 * 45-96): both run_sgm directions, the consistency check and nothing else on
 * the host but the image pyramids the constructor builds (:27-39) and the
 * depth ranges from the bundle (:53-54, 59-60, reference code).
 */
mve::FloatImage::Ptr
SGMStereo::reconstruct (SGMStereo::Options sgm_opts, StereoView::Ptr main_view,
    StereoView::Ptr neighbor, mve::Bundle::ConstPtr bundle)
{
    float range_main[2] = { sgm_opts.min_depth, sgm_opts.max_depth };
    float range_neig[2] = { sgm_opts.min_depth, sgm_opts.max_depth };
    if (bundle != nullptr && sgm_opts.max_depth == 0.0)
    {
        fill_depth_range_for_view(bundle, main_view, range_main);
        /* :59-60 overwrites the SAME array: a range the second call leaves
         * untouched keeps the first view's value */
        range_neig[0] = range_main[0];
        range_neig[1] = range_main[1];
        fill_depth_range_for_view(bundle, neighbor, range_neig);
    }
    SGMStereo sgm1(sgm_opts, main_view, neighbor);    /* image pyramids */
    mve::ByteImage::ConstPtr a = sgm1.main_image, b = sgm1.neighbor_image;

    math::Matrix3f M_mn, M_nm;
    math::Vec3f t_mn, t_nm;
    main_view->get_camera().fill_reprojection(neighbor->get_camera(),
        a->width(), a->height(), b->width(), b->height(), *M_mn, *t_mn);
    neighbor->get_camera().fill_reprojection(main_view->get_camera(),
        b->width(), b->height(), a->width(), a->height(), *M_nm, *t_nm);

    mve::FloatImage::Ptr depth = mve::FloatImage::create(a->width(),
        a->height(), 1);
    int const rc = smvsb_sgm_reconstruct(
        smvs_b200_integration::thread_device(), a->width(), a->height(),
        a->begin(), b->width(), b->height(), b->begin(), *M_mn, *t_mn, *M_nm,
        *t_nm, range_main, range_neig, sgm_opts.num_steps, sgm_opts.penalty1,
        sgm_opts.penalty2, nullptr, depth->begin(), nullptr);
    if (rc != SMVSB_OK)
        throw std::runtime_error(std::string("smvs_b200: ")
            + smvsb_last_error(nullptr));
    if (sgm_opts.debug_lvl > 1)
        std::cout << "SGM finished." << std::endl;
    return depth;
}

SMVS_NAMESPACE_END
