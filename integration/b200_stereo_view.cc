/*
 * integration/b200_stereo_view.cc
 *
 * Drop-in body for smvs::StereoView::set_scale (reference:
 * lib/stereo_view.cc:24-46, with initialize_image_gradients :48-62 and
 * compute_gradients_and_hessian :97-188): Gaussian blur (+ luminance of a
 * colour view) + gradient / Hessian images on the GPU through
 * smvsb_view_set_scale_c, results written into the members every other
 * reference function reads (scaleimage, image_grad, image_hessian). Images
 * with 2 or 4 channels and the debug variant keep the reference's own body.
 * lib/stereo_view.h untouched.
 */
#include "stereo_view.h"

#include "b200_context.h"

SMVS_NAMESPACE_BEGIN

/* The reference's own set_scale, kept under this name by integration/Makefile
 * (objcopy --redefine-sym on a second, all-weak copy of the object). */
extern "C" void smvs_ref_stereo_view_set_scale (StereoView* self, int scale,
    bool debug);

void
StereoView::set_scale (int scale, bool debug)
{
    /* whatever a context of this thread holds of older images is stale now */
    smvs_b200_integration::views_generation() += 1;
    int const ch = this->image->channels();
    if (debug || (ch != 1 && ch != 3))
    {
        smvs_ref_stereo_view_set_scale(this, scale, debug);
        return;
    }
    smvsb::Context& gpu = smvs_b200_integration::thread_context();
    int const w = this->image->width(), h = this->image->height();
    this->scaleimage = mve::FloatImage::create(w, h, ch);
    this->image_grad = mve::FloatImage::create(w, h, 2);
    this->image_hessian = mve::FloatImage::create(w, h, 3);
    gpu.check(smvsb_view_set_scale_c(gpu.get(), w, h, ch,
        this->image->begin(), scale, this->scaleimage->begin(),
        this->image_grad->begin(), this->image_hessian->begin()));
    this->view->cache_cleanup();
}

SMVS_NAMESPACE_END
