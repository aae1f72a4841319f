"""Shared host logic of the model wrappers: weight files, tiled inference and evaluation.

Mirrors the flow the reference repeats in every ``super_resolve_image`` (SRCNN_model.py:111-247,
EDSR_model.py:189-315, ESRGAN_model.py:858-979) and ``evaluate`` (SRCNN_model.py:100-109,
EDSR_model.py:178-187) with the Python patch loops replaced by device kernels.
"""
from __future__ import annotations

import os

import numpy as np

from .. import _capi as capi
from .. import engine, ops


def load_weight_file(path):
    """``.npz`` with Keras-layout arrays (``<layer>/kernel`` HWIO, ``<layer>/bias``).  Keys may be the internal layer
    names or the reference model's own variable names (``conv2d_3/kernel:0`` ...: ``weights.adopt`` maps the auto-named
    layers in creation order and validates every shape when the network is constructed).  The reference
    stores ``.h5`` (SRCNN_model.py:258); reading those needs h5py, which this image lacks."""
    if path is None or not os.path.isfile(path):
        raise FileNotFoundError(f"Pretrained model file not found at {path}")
    if path.endswith(".npz"):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    if path.endswith((".h5", ".hdf5", ".keras")):
        try:
            import h5py  # noqa: F401
        except ImportError as e:
            raise NotImplementedError(
                "Keras .h5 checkpoints need h5py, which is not installed; export the weights to .npz "
                "(np.savez(path, **{v.name: v.numpy() for v in model.weights}))") from e
        raise NotImplementedError("Keras .h5 import is not implemented; use .npz")
    raise ValueError(f"unsupported weight file {path}")


def save_weight_file(path, weights):
    np.savez(path, **weights)


def as_device_image(img):
    """HxWxC numpy (uint8 / float) or CUDA tensor -> float32 CUDA tensor, values as given."""
    torch = capi.require_cuda()
    if isinstance(img, torch.Tensor):
        t = img if img.is_cuda else img.cuda()
        return t.float().contiguous()
    if img is None or not isinstance(img, np.ndarray):
        raise ValueError("lr_img must be a numpy array (RGB).")
    return torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).cuda()


def tiled_super_resolve(model: engine.DeviceModel, image_dev, patch, stride, scale, pre=None, post=None):
    """pad (reflect, bottom/right) -> patches -> network -> overlap-average -> crop -> clip(0,1).

    image_dev: [H,W,C] float32 CUDA.  ``pre``/``post`` are (scale, shift) pairs applied to the
    patches before / after the network (ESRGAN's [-1,1] mapping, ESRGAN_model.py:929,946).
    Returns (sr [H*scale, W*scale, C] float32 CUDA, inference_metrics)."""
    h, w, _ = image_dev.shape
    patches, (ph, pw, ny, nx) = ops.pad_extract(image_dev, patch, stride)
    if pre is not None:
        patches = ops.cast(patches, patches.dtype, pre[0], pre[1])
    preds, metrics = engine.timed_predict(model, patches)
    if post is not None:
        preds = ops.cast(preds, preds.dtype, post[0], post[1])
    sr = ops.overlap_add(preds.contiguous(), ny, nx, stride * scale, h * scale, w * scale)
    return sr, metrics


def whole_image_super_resolve(model: engine.DeviceModel, image_dev, pre=None, post=None):
    """Fully-convolutional fast path: the network runs once over the whole image instead of
    over overlapping patches - no padding, no patch loop, no overlap averaging - then ``clip(0, 1)``.  Every reference
    network is built with ``input_shape=(None, None, C)`` (EDSR_model.py:98), so this is the same function of the image
    wherever a pixel's receptive field lies inside one patch of the tiled flow; near patch borders the tiled flow sees
    zero padding and averages overlapping predictions, which this path does not reproduce (it is an additional entry
    point, not a replacement for ``super_resolve_image``).
    Returns (sr [H*scale, W*scale, C] float32 CUDA, inference_metrics)."""
    x = image_dev[None].contiguous()
    if pre is not None:
        x = ops.cast(x, x.dtype, pre[0], pre[1])
    preds, metrics = engine.timed_predict(model, x)
    if post is not None:
        preds = ops.cast(preds, preds.dtype, post[0], post[1])
    return preds[0].clamp_(0.0, 1.0), metrics


def evaluate_arrays(model: engine.DeviceModel, X, Y, micro_batch=64, sums=None):
    """Forward X in micro-batches and accumulate (sum psnr, sum ssim, count, sum mse) on the device.

    X, Y: numpy NHWC float32 (or CUDA tensors).  Returns the float64[4] CUDA tensor ``sums``."""
    torch = capi.require_cuda()
    n = len(X)
    if len(Y) != n:
        raise ValueError("X and Y must have the same number of samples")
    if sums is None:
        sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    for i in range(0, n, micro_batch):
        xb, yb = X[i:i + micro_batch], Y[i:i + micro_batch]
        if not isinstance(xb, torch.Tensor):
            xb = torch.from_numpy(np.ascontiguousarray(xb, dtype=np.float32)).cuda(non_blocking=True)
        if not isinstance(yb, torch.Tensor):
            yb = torch.from_numpy(np.ascontiguousarray(yb, dtype=np.float32)).cuda(non_blocking=True)
        pred = model.predict_device(xb.float().contiguous())
        ops.psnr_ssim(yb.float().contiguous(), pred.contiguous(), 1.0, sums=sums)
    return sums


def finish_evaluation(sums, group=None):
    """All-reduce the 4 sums when torch.distributed is initialised, then form the Keras-style sample
    means -> [loss, psnr, ssim] (row A12 of SURVEY.md section 8)."""
    from .. import distributed as dist
    total = dist.allreduce_sums(sums, group)
    s = total.cpu().numpy()
    cnt = max(s[2], 1.0)
    return [float(s[3] / cnt), float(s[0] / cnt), float(s[1] / cnt)]
