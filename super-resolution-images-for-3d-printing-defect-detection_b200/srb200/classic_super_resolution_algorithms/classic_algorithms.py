"""``interpolate_bicubic`` / ``interpolate_bilinear`` / ``interpolate_area`` / ``interpolate_lanczos`` with the reference's
signatures (SRModels/classic_super_resolution_algorithms/classic_algorithms.py:7-21), computed on the GPU.

``target_shape`` is ``(width, height)`` exactly as for ``cv2.resize``; dtype is preserved (uint8 or
float32); float results are not clipped.  uint8 images take OpenCV's fixed-point paths (linear, area and Lanczos-4
bit-exact; bicubic as OpenCV's default dispatch, <= 1 LSB).  Accepts ``[H, W, C]`` / ``[H, W]`` images or an
``[B, H, W, C]`` batch; numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .. import _capi as capi
from .. import ops

INTER_LINEAR, INTER_CUBIC, INTER_AREA, INTER_LANCZOS4 = 1, 2, 3, 4   # cv2 codes


def resize_cubic(img, target_shape: Tuple[int, int], clip01=False, fixed_point=False):
    torch = capi.require_cuda()
    was_numpy = not isinstance(img, torch.Tensor)
    if was_numpy:
        arr = np.ascontiguousarray(img)
        if arr.dtype not in (np.uint8, np.float32):
            arr = arr.astype(np.float32)
        t = torch.from_numpy(arr).cuda()
    else:
        t = img if img.is_cuda else img.cuda()
    shape_in = t.dim()
    if shape_in == 2:
        t = t[None, :, :, None]
    elif shape_in == 3:
        t = t[None]
    elif shape_in != 4:
        raise ValueError(f"expected an image or NHWC batch, got shape {tuple(t.shape)}")
    w, h = int(target_shape[0]), int(target_shape[1])
    if w <= 0 or h <= 0:
        raise ValueError("target_shape must be (width, height) with positive entries")
    out = ops.bicubic(t.contiguous(), h, w, clip01=clip01, fixed_point=fixed_point)
    if shape_in == 2:
        out = out[0, :, :, 0]
    elif shape_in == 3:
        out = out[0]
    return out.cpu().numpy() if was_numpy else out


def interpolate_bicubic(lr_img, target_shape: Tuple[int, int]):
    """Bicubic upscaling (== cv2.resize(lr_img, target_shape, interpolation=cv2.INTER_CUBIC))."""
    return resize_cubic(lr_img, target_shape)


def _not_on_path(name):
    def f(*_a, **_k):
        raise NotImplementedError(
            f"{name} is outside the B200 hot path (SURVEY.md section 8f): only the four cv2.resize interpolators are built")
    f.__name__ = name
    return f


# the reference's other classical upscalers (classic_algorithms.py:23-110) are outside the hot path; the names stay
# importable so that ``from ...classic_algorithms import *`` style notebooks fail at the call, with a reason
back_projection = _not_on_path("back_projection")
non_local_means = _not_on_path("non_local_means")
edge_guided_interpolation = _not_on_path("edge_guided_interpolation")
frequency_extrapolation = _not_on_path("frequency_extrapolation")


def _resize_float(img, target_shape, code):
    torch = capi.require_cuda()
    was_numpy = not isinstance(img, torch.Tensor)
    if was_numpy:
        arr = np.ascontiguousarray(img)
        if arr.dtype not in (np.uint8, np.float32):
            arr = arr.astype(np.float32)
        t = torch.from_numpy(arr).cuda()
    else:
        t = img if img.is_cuda else img.cuda()
    if t.dtype not in (torch.float32, torch.uint8):
        raise TypeError(f"resize supports float32 and uint8 images, got {t.dtype}")
    nd = t.dim()
    t4 = t[None, :, :, None] if nd == 2 else t[None] if nd == 3 else t
    w, h = int(target_shape[0]), int(target_shape[1])
    if w <= 0 or h <= 0:
        raise ValueError("target_shape must be (width, height) with positive entries")
    out = ops.resize(t4.contiguous(), h, w, interpolation=code)
    out = out[0, :, :, 0] if nd == 2 else out[0] if nd == 3 else out
    return out.cpu().numpy() if was_numpy else out


def interpolate_bilinear(lr_img, target_shape: Tuple[int, int]):
    """Bilinear upscaling (== cv2.resize(lr_img, target_shape, interpolation=cv2.INTER_LINEAR), classic_algorithms.py:7-9)."""
    return _resize_float(lr_img, target_shape, capi.INTER_LINEAR)


def interpolate_area(lr_img, target_shape: Tuple[int, int]):
    """Area upscaling (== cv2.resize(..., interpolation=cv2.INTER_AREA) when the image grows, classic_algorithms.py:15-17)."""
    return _resize_float(lr_img, target_shape, capi.INTER_AREA)



def interpolate_lanczos(lr_img, target_shape: Tuple[int, int]):
    """Lanczos-4 upscaling (== cv2.resize(..., interpolation=cv2.INTER_LANCZOS4), classic_algorithms.py:19-21)."""
    return _resize_float(lr_img, target_shape, capi.INTER_LANCZOS4)
