/*
 * Shim of MVE mve/camera.h: pinhole camera with normalised focal length,
 * all math in fp32 as in MVE. TEST INFRASTRUCTURE ONLY (oracle build).
 */
#ifndef SHIM_MVE_CAMERA_HEADER
#define SHIM_MVE_CAMERA_HEADER

#include <algorithm>

#include "math/matrix.h"
#include "math/vector.h"
#include "mve/defines.h"

MVE_NAMESPACE_BEGIN

struct CameraInfo
{
    CameraInfo (void)
        : flen(0.0f), paspect(1.0f)
    {
        ppoint[0] = ppoint[1] = 0.5f;
        dist[0] = dist[1] = 0.0f;
        trans[0] = trans[1] = trans[2] = 0.0f;
        std::fill(rot, rot + 9, 0.0f);
        rot[0] = rot[4] = rot[8] = 1.0f;
    }

    void fill_calibration (float* mat, float width, float height) const
    {
        float dim_aspect = width / height;
        float image_aspect = dim_aspect * this->paspect;
        float ax, ay;
        if (image_aspect < 1.0f)
        {
            ax = this->flen * height / this->paspect;
            ay = this->flen * height;
        }
        else
        {
            ax = this->flen * width;
            ay = this->flen * width * this->paspect;
        }
        mat[0] = ax;   mat[1] = 0.0f; mat[2] = width * this->ppoint[0];
        mat[3] = 0.0f; mat[4] = ay;   mat[5] = height * this->ppoint[1];
        mat[6] = 0.0f; mat[7] = 0.0f; mat[8] = 1.0f;
    }

    void fill_inverse_calibration (float* mat, float width, float height) const
    {
        float dim_aspect = width / height;
        float image_aspect = dim_aspect * this->paspect;
        float ax, ay;
        if (image_aspect < 1.0f)
        {
            ax = this->flen * height / this->paspect;
            ay = this->flen * height;
        }
        else
        {
            ax = this->flen * width;
            ay = this->flen * width * this->paspect;
        }
        mat[0] = 1.0f / ax; mat[1] = 0.0f; mat[2] = -width * ppoint[0] / ax;
        mat[3] = 0.0f; mat[4] = 1.0f / ay; mat[5] = -height * ppoint[1] / ay;
        mat[6] = 0.0f; mat[7] = 0.0f;      mat[8] = 1.0f;
    }

    void fill_world_to_cam_rot (float* mat) const
    { std::copy(rot, rot + 9, mat); }
    void fill_cam_to_world_rot (float* mat) const
    {
        mat[0] = rot[0]; mat[1] = rot[3]; mat[2] = rot[6];
        mat[3] = rot[1]; mat[4] = rot[4]; mat[5] = rot[7];
        mat[6] = rot[2]; mat[7] = rot[5]; mat[8] = rot[8];
    }
    void fill_camera_translation (float* vec) const
    { std::copy(trans, trans + 3, vec); }
    /* camera position = -R^T t (MVE, restated) */
    void fill_camera_pos (float* pos) const
    {
        pos[0] = -rot[0] * trans[0] - rot[3] * trans[1] - rot[6] * trans[2];
        pos[1] = -rot[1] * trans[0] - rot[4] * trans[1] - rot[7] * trans[2];
        pos[2] = -rot[2] * trans[0] - rot[5] * trans[1] - rot[8] * trans[2];
    }
    /* 4x4 camera-to-world: [R^T | camera position] (MVE, restated) */
    void fill_cam_to_world (float* mat) const
    {
        mat[0]  = rot[0]; mat[1]  = rot[3]; mat[2]  = rot[6];
        mat[4]  = rot[1]; mat[5]  = rot[4]; mat[6]  = rot[7];
        mat[8]  = rot[2]; mat[9]  = rot[5]; mat[10] = rot[8];
        mat[3]  = -(rot[0] * trans[0] + rot[3] * trans[1] + rot[6] * trans[2]);
        mat[7]  = -(rot[1] * trans[0] + rot[4] * trans[1] + rot[7] * trans[2]);
        mat[11] = -(rot[2] * trans[0] + rot[5] * trans[1] + rot[8] * trans[2]);
        mat[12] = 0.0f; mat[13] = 0.0f; mat[14] = 0.0f; mat[15] = 1.0f;
    }

    void fill_reprojection (CameraInfo const& destination,
        float src_width, float src_height, float dst_width, float dst_height,
        float* mat, float* vec) const
    {
        math::Matrix3f dst_K, dst_R, src_Ri, src_Ki;
        math::Vec3f dst_t, src_t;
        destination.fill_calibration(dst_K.begin(), dst_width, dst_height);
        destination.fill_world_to_cam_rot(dst_R.begin());
        destination.fill_camera_translation(dst_t.begin());
        this->fill_cam_to_world_rot(src_Ri.begin());
        this->fill_inverse_calibration(src_Ki.begin(), src_width, src_height);
        this->fill_camera_translation(src_t.begin());

        math::Matrix3f ret_mat = dst_K * dst_R * src_Ri * src_Ki;
        math::Vec3f ret_vec = dst_K * (dst_t - dst_R * src_Ri * src_t);
        std::copy(ret_mat.begin(), ret_mat.end(), mat);
        std::copy(ret_vec.begin(), ret_vec.end(), vec);
    }

    float flen;
    float ppoint[2];
    float paspect;
    float dist[2];
    float trans[3];
    float rot[9];
};

MVE_NAMESPACE_END

#endif
