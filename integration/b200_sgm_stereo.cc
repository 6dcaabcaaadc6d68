/*
 * integration/b200_sgm_stereo.cc
 *
 * Drop-in body for smvs::SGMStereo::run_sgm (reference:
 * lib/sgm_stereo.cc:98-124): cost volume, 8-path aggregation and
 * winner-takes-all on the GPU through smvsb_sgm. lib/sgm_stereo.h untouched.
 */
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

SMVS_NAMESPACE_END
