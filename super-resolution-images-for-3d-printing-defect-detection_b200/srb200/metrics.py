"""PSNR / SSIM with the reference's signatures (SRModels/metrics.py:3-7).

``psnr(y_true, y_pred)`` and ``ssim(y_true, y_pred)`` replace ``tf.image.psnr/ssim(..., max_val=1.0)``:
per-image float32 vectors ``[B]`` (a scalar for a single ``[H, W, C]`` image).  Inputs may be numpy
arrays (copied to the device, result returned as numpy) or CUDA torch tensors (result stays on
the device).  Both functions run the same fused one-pass kernel; ``psnr_ssim`` returns both at once.
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


def psnr_ssim(y_true, y_pred, max_val=1.0):
    a, np_a, single = _to_device(y_true)
    b, np_b, _ = _to_device(y_pred)
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.shape[1] < 11 or a.shape[2] < 11:
        raise ValueError(f"image dimensions must be at least 11x11 for SSIM, got {a.shape[1]}x{a.shape[2]}")
    p, s = ops.psnr_ssim(a, b, max_val)
    if single:
        p, s = p[0], s[0]
    if np_a and np_b:
        return p.cpu().numpy(), s.cpu().numpy()
    return p, s


def psnr(y_true, y_pred):
    return psnr_ssim(y_true, y_pred, 1.0)[0]


def ssim(y_true, y_pred):
    return psnr_ssim(y_true, y_pred, 1.0)[1]
