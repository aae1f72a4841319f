"""PSNR / SSIM with the reference's signatures (SRModels/metrics.py:3-7).

``psnr(y_true, y_pred)`` and ``ssim(y_true, y_pred)`` replace ``tf.image.psnr/ssim(..., max_val=1.0)``:
per-image float32 vectors ``[B]`` (a scalar for a single ``[H, W, C]`` image).  Inputs may be numpy
arrays (copied to the device, result returned as numpy) or CUDA torch tensors (result stays on
the device).  Both functions run the same fused one-pass kernel; ``psnr_ssim`` returns both at once.
``peak_signal_noise_ratio`` / ``structural_similarity`` are the skimage.metrics definitions the classical
benchmark notebook uses (7 x 7 uniform window), on the same kernel.
"""
from __future__ import annotations

import numpy as np

from . import _capi as capi
from . import ops


def _to_device(a):
    torch = capi.require_cuda()
    if isinstance(a, torch.Tensor):
        t, was_numpy = a, False
        if not t.is_cuda:
            t = t.cuda()
    else:
        t, was_numpy = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda(), True
    if t.dtype != torch.float32:
        t = t.float()
    single = t.dim() == 3
    if single:
        t = t.unsqueeze(0)
    if t.dim() != 4:
        raise ValueError(f"expected [B, H, W, C] or [H, W, C], got shape {tuple(t.shape)}")
    return t.contiguous(), was_numpy, single


def psnr_ssim(y_true, y_pred, max_val=1.0, exact=False):
    """Both metrics from one pass.  ``exact=True`` keeps the float32 Gaussian on the CUDA cores for every image size (by
    default wide 1- / 3-channel images use the tensor-path kernel: fp16 window summing to 1, |dSSIM| <= ~3e-6)."""
    a, np_a, single = _to_device(y_true)
    b, np_b, _ = _to_device(y_pred)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.shape[1] < 11 or a.shape[2] < 11:
        raise ValueError(f"image dimensions must be at least 11x11 for SSIM, got {a.shape[1]}x{a.shape[2]}")
    p, s = ops.psnr_ssim(a, b, max_val, window=capi.SSIM_TF_EXACT if exact else capi.SSIM_TF)
    if single:
        p, s = p[0], s[0]
    if np_a and np_b:
        return p.cpu().numpy(), s.cpu().numpy()
    return p, s


def psnr(y_true, y_pred):
    """tf.image.psnr(y_true, y_pred, max_val=1.0) (metrics.py:3-4): the squared-error reduction alone, any image size."""
    a, np_a, single = _to_device(y_true)
    b, np_b, _ = _to_device(y_pred)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
    p = ops.psnr(a, b, 1.0)
    if single:
        p = p[0]
    return p.cpu().numpy() if (np_a and np_b) else p


def ssim(y_true, y_pred):
    return psnr_ssim(y_true, y_pred, 1.0)[1]


# ---------------------------------------------------------------------------------------------
# skimage.metrics definitions, as the classical benchmark calls them
# (super_resolucion_clasica.ipynb cell 7: psnr(hr_f, sr_f, data_range=1.0), ssim(hr_f, sr_f, channel_axis=2,
#  data_range=1.0), grayscale: ssim(hr_g, sr_g, data_range=dr); EDA.ipynb: data_range=255 on uint8)
# ---------------------------------------------------------------------------------------------
def _skimage_pair(im1, im2, data_range, channel_axis, metric="ssim"):
    torch = capi.require_cuda()
    tensors = []
    for im in (im1, im2):
        t = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im))
        t = t.cuda() if not t.is_cuda else t
        if t.dim() == 2 and channel_axis is None:
            t = t[:, :, None]
        elif t.dim() == 3 and channel_axis is not None:
            t = t.movedim(channel_axis, -1)
        else:
            raise ValueError("expected a 2-D image (channel_axis=None) or a 3-D image with channel_axis given, "
                             f"got shape {tuple(t.shape)} with channel_axis={channel_axis}")
        tensors.append(t)
    a, b = tensors
    if a.shape != b.shape:
        raise ValueError("Input images must have the same dimensions.")
    if data_range is None:
        if metric == "ssim":
            if a.dtype.is_floating_point:
                raise ValueError("Since image dtype is floating point, you must specify the data_range parameter. Please read "
                                 "the documentation carefully (including the note). It is recommended that you always "
                                 "specify the data_range anyway.")
            data_range = 255.0
        else:
            # skimage.metrics.peak_signal_noise_ratio: the dtype's range - (-1, 1) for floats, so 1 for non-negative float
            # images and 2 otherwise; integer images 0 .. dtype max
            if a.dtype.is_floating_point:
                lo, hi = float(a.min().item()), float(a.max().item())
                if hi > 1.0 or lo < -1.0:
                    raise ValueError("image_true has intensity values outside the range expected for its data type. Please "
                                     "manually specify the data_range.")
                data_range = 1.0 if lo >= 0.0 else 2.0
            else:
                data_range = float(torch.iinfo(a.dtype).max) if a.dtype != torch.bool else 1.0
    # both metrics are invariant under a common scale when data_range scales with it: evaluate on [0, 1]
    scale = 1.0 / float(data_range)
    a = (a.to(torch.float32) * scale).contiguous()[None]
    b = (b.to(torch.float32) * scale).contiguous()[None]
    return a, b


def peak_signal_noise_ratio(image_true, image_test, *, data_range=None):
    """skimage.metrics.peak_signal_noise_ratio: 10 log10(data_range^2 / mse) over the whole array -> float.
    ``data_range=None`` falls back to the dtype's range as skimage does.  The squared error is summed in float64 from
    float32 differences of the [0, 1]-scaled images (skimage subtracts in float64: relative difference ~1e-7)."""
    a, b = _skimage_pair(image_true, image_test, data_range, None if np.ndim(image_true) == 2 else -1, metric="psnr")
    return float(ops.psnr(a, b, 1.0, window=capi.SSIM_SKIMAGE)[0].item())


def structural_similarity(im1, im2, *, win_size=None, data_range=None, channel_axis=None, gaussian_weights=False,
                          full=False):
    """skimage.metrics.structural_similarity with its defaults (7 x 7 uniform window, sample covariance, K1 = 0.01,
    K2 = 0.03, mean over the border-cropped map and then over channels) -> float."""
    if gaussian_weights or full or (win_size not in (None, 7)):
        raise NotImplementedError("only skimage's default window (win_size=7, uniform, full=False) is built")
    a, b = _skimage_pair(im1, im2, data_range, channel_axis)
    if a.shape[1] < 7 or a.shape[2] < 7:
        raise ValueError("win_size exceeds image extent. Either ensure that your images are at least 7x7; or pass "
                         "win_size explicitly in the function call, with an odd value less than or equal to the "
                         "smaller side of your images.")
    _, s = ops.psnr_ssim(a, b, 1.0, window=capi.SSIM_SKIMAGE)
    return float(s[0].item())
