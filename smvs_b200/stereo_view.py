"""Host-side mirror of StereoView::set_scale (lib/stereo_view.cc:24-46,
97-188): per-scale Gaussian blur, then the 3x3 quadratic-fit gradient (2 ch)
and Hessian (3 ch: xx, xy, yy). This is the producer of the hot path's image
inputs (SURVEY.md section 8f "next #1"); it runs on the host in numpy, with
the fp32 operation order of mve::image::blur_gaussian (separable, radius
ceil(2.884 sigma), clamped borders, normalised per pixel) and the fp64 order
of the reference's 6x9 stencil, so that its output is interchangeable with
the reference's.
"""
from __future__ import annotations

import math

import numpy as np

# lib/stereo_view.cc:104-162
_M = np.array([
    [1 / 6, 1 / 6, 1 / 6, -1 / 3, -1 / 3, -1 / 3, 1 / 6, 1 / 6, 1 / 6],
    [1 / 6, -1 / 3, 1 / 6, 1 / 6, -1 / 3, 1 / 6, 1 / 6, -1 / 3, 1 / 6],
    [1 / 4, 0, -1 / 4, 0, 0, 0, -1 / 4, 0, 1 / 4],
    [-1 / 6, -1 / 6, -1 / 6, 0, 0, 0, 1 / 6, 1 / 6, 1 / 6],
    [-1 / 6, 0, 1 / 6, -1 / 6, 0, 1 / 6, -1 / 6, 0, 1 / 6],
    [-1 / 9, 2 / 9, -1 / 9, 2 / 9, 5 / 9, 2 / 9, -1 / 9, 2 / 9, -1 / 9]],
    dtype=np.float64)
_M[0] = [1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0, -1.0 / 3.0, -1.0 / 3.0, -1.0 / 3.0,
         1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0]


def byte_to_float(img_u8):
    """mve::image::byte_to_float_image."""
    return np.clip(img_u8.astype(np.float32) / np.float32(255.0), 0.0, 1.0).astype(np.float32)


def blur_gaussian(img, sigma):
    """mve::image::blur_gaussian<float> on a single-channel fp32 image."""
    sigma = np.float32(sigma)
    if -0.1 <= sigma <= 0.1:
        return img.copy()
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    ks = int(math.ceil(float(sigma * np.float32(2.884))))
    xs = np.arange(ks + 1, dtype=np.float32)
    # expf of glibc is correctly rounded; numpy's vectorised float32 exp is
    # not, so go through a double exp and round once.
    arg = (-((xs * xs) / (np.float32(2) * sigma * sigma))).astype(np.float32)
    kernel = np.array([np.float32(math.exp(float(a))) for a in arg], dtype=np.float32)

    def conv(src, axis):
        n = src.shape[axis]
        acc = np.zeros_like(src)
        wsum = np.float32(0)
        idx = np.arange(n)
        for i in range(-ks, ks + 1):
            j = np.clip(idx + i, 0, n - 1)
            k = kernel[abs(i)]
            acc = (acc + np.take(src, j, axis=axis) * k).astype(np.float32)
            wsum = np.float32(wsum + k)
        return (acc / wsum).astype(np.float32)

    return conv(conv(img, 1), 0)


def gradients_and_hessian(img):
    """StereoView::compute_gradients_and_hessian: (H,W,2) and (H,W,3) fp32,
    zero on the one-pixel border."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    h, w = img.shape
    v = []
    for a in (-1, 0, 1):          # x offset (outer loop of :170-172)
        for b in (-1, 0, 1):      # y offset
            v.append(img[1 + b:h - 1 + b, 1 + a:w - 1 + a].astype(np.float64))
    r = []
    for row in range(6):
        s = np.zeros_like(v[0])
        for k in range(9):
            s = s + _M[row, k] * v[k]
        r.append(s)
    grad = np.zeros((h, w, 2), dtype=np.float32)
    hess = np.zeros((h, w, 3), dtype=np.float32)
    grad[1:-1, 1:-1, 0] = r[3]
    grad[1:-1, 1:-1, 1] = r[4]
    hess[1:-1, 1:-1, 0] = 2.0 * r[0]
    hess[1:-1, 1:-1, 1] = r[2]
    hess[1:-1, 1:-1, 2] = 2.0 * r[1]
    return grad, hess


def desaturate_luminance(img):
    """mve::image::desaturate<float>(img, DESATURATE_LUMINANCE) of an
    (H, W, 3) image: v0 * 0.21f + v1 * 0.72f + v2 * 0.07f, left to right."""
    img = np.ascontiguousarray(img, dtype=np.float32)
    a = (img[:, :, 0] * np.float32(0.21)).astype(np.float32)
    b = (img[:, :, 1] * np.float32(0.72)).astype(np.float32)
    c = (img[:, :, 2] * np.float32(0.07)).astype(np.float32)
    return ((a + b).astype(np.float32) + c).astype(np.float32)


def set_scale(img_u8, scale):
    """Blurred image, gradient and Hessian of one view at `scale`
    (sigma = 0.12 * 2^scale + 0.2, lib/stereo_view.cc:28). A three-channel
    image is blurred channel by channel (mve::image::blur_gaussian) and the
    blurred image desaturated before the stencil (:48-62); the blurred image
    returned keeps its channels."""
    sigma = 0.12 * math.pow(2.0, scale) + 0.2
    f = byte_to_float(img_u8)
    if f.ndim == 3:
        blurred = np.stack([blur_gaussian(f[:, :, c], sigma) for c in range(f.shape[2])],
                           axis=2)
        grad, hess = gradients_and_hessian(desaturate_luminance(blurred))
        return blurred, grad, hess
    blurred = blur_gaussian(f, sigma)
    grad, hess = gradients_and_hessian(blurred)
    return blurred, grad, hess


def shading_inputs(img_u8):
    """StereoView::initialize_linear without gamma (lib/stereo_view.cc:64-84):
    shading image = unblurred float image, plus its gradient."""
    shading = byte_to_float(img_u8)
    grad, _ = gradients_and_hessian(shading)
    return shading, grad
