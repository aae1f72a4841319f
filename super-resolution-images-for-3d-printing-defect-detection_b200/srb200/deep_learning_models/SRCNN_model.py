"""``SRCNNModel`` with the reference's inference surface (SRModels/deep_learning_models/SRCNN_model.py).

Kept: ``setup_model`` (:23-43), ``evaluate`` (:100-109), ``super_resolve_image`` (:111-247), the
``_trained`` gate and the exceptions.  Not built: ``fit`` / ``save`` to .h5 (training is out of scope).
"""
from __future__ import annotations

import numpy as np

from .. import _capi as capi
from .. import engine, ops, weights as W
from ..classic_super_resolution_algorithms.classic_algorithms import INTER_CUBIC
from . import _common as common


class SRCNNModel:
    def __init__(self):
        self.model = None
        self._trained = False

    def setup_model(self, input_shape=None, learning_rate=1e-4, from_pretrained=False, pretrained_path=None,
                    precision="fp32", seed=1234):
        """Loads pretrained weights (.npz) or builds a Keras-default-initialised 9-1-5 network."""
        if from_pretrained:
            w = common.load_weight_file(pretrained_path)
            self.model = engine.SRCNNNet(w, precision=precision)
            print(f"Loaded pretrained model from {pretrained_path}")
            self._trained = True
        else:
            if input_shape is None:
                raise ValueError("input_shape must be provided when not using a pretrained model.")
            channels = int(input_shape[-1])
            self.model = engine.SRCNNNet(W.srcnn_weights(seed=seed, channels=channels), precision=precision)
            self.model.summary()

    def load_weights(self, weights: dict, precision=None):
        """Install a Keras-layout weight dict and mark the model usable (what ``from_pretrained`` does)."""
        self.model = engine.SRCNNNet(weights, precision=precision or (self.model.precision if self.model else "fp32"))
        self._trained = True

    def fit(self, *a, **k):
        raise NotImplementedError("training is outside the B200 inference path (SURVEY.md section 2)")

    def evaluate(self, X_test, Y_test):
        """Evaluates the model -> [loss (MSE), psnr, ssim] sample means."""
        if not self._trained:
            raise RuntimeError("Model has not been trained.")
        sums = common.evaluate_arrays(self.model, X_test, Y_test)
        results = common.finish_evaluation(sums)
        print(f"Loss: {results[0]:.4f}, PSNR: {results[1]:.2f} dB, SSIM: {results[2]:.4f}")
        return results

    def super_resolve_image(self, lr_img, hr_h, hr_w, patch_size=33, stride=14, interpolation=INTER_CUBIC):
        """Upscale the LR image to (hr_w, hr_h) with the given cv2 interpolation (bicubic by default), then patch-wise SRCNN with overlap averaging.
        Returns (float32 RGB in [0,1] of shape (hr_h, hr_w, 3), inference_metrics)."""
        if not self._trained:
            raise RuntimeError("Model has not been trained.")
        if lr_img is None or not isinstance(lr_img, np.ndarray):
            raise ValueError("lr_img must be a numpy array (RGB).")
        if interpolation not in (capi.INTER_LINEAR, capi.INTER_CUBIC, capi.INTER_AREA, capi.INTER_LANCZOS4):
            raise NotImplementedError(f"cv2 interpolation code {interpolation} is not built on the device "
                                      "(INTER_LINEAR, INTER_CUBIC, INTER_AREA and INTER_LANCZOS4 are)")
        if interpolation != INTER_CUBIC and lr_img.dtype == np.uint8:
            raise NotImplementedError("uint8 input is built for cv2.INTER_CUBIC only (OpenCV's fixed-point filters)")
        torch = capi.require_cuda()
        if lr_img.dtype == np.uint8:
            # cv2.resize keeps uint8; the network then sees 0..255 values, exactly as in the reference
            src = torch.from_numpy(np.ascontiguousarray(lr_img)).cuda()
        else:
            src = common.as_device_image(lr_img)
        if interpolation == INTER_CUBIC:
            up = ops.bicubic(src[None], hr_h, hr_w)[0].float()
        else:
            up = ops.resize(src[None], hr_h, hr_w, interpolation=interpolation)[0]
        sr, metrics = common.tiled_super_resolve(self.model, up, patch_size, stride, 1)
        return sr.cpu().numpy(), metrics

    def super_resolve_image_whole(self, lr_img, hr_h, hr_w, interpolation=INTER_CUBIC):
        """Fast path without tiling: pre-upsample, then one fully-convolutional pass over the whole image."""
        if not self._trained:
            raise RuntimeError("Model has not been trained.")
        if lr_img is None or not isinstance(lr_img, np.ndarray):
            raise ValueError("lr_img must be a numpy array (RGB).")
        src = common.as_device_image(lr_img)
        up = ops.resize(src[None], hr_h, hr_w, interpolation=interpolation)[0]
        sr, metrics = common.whole_image_super_resolve(self.model, up)
        return sr.cpu().numpy(), metrics

    def save(self, directory, timestamp):
        if not self._trained:
            raise RuntimeError("Cannot save an untrained model.")
        if not directory:
            raise ValueError("Directory path must be provided.")
        import os
        os.makedirs(directory, exist_ok=True)
        filepath = os.path.join(directory, f"SRCNN_{timestamp}.npz")
        common.save_weight_file(filepath, self.model.get_weights_dict())
        print(f"Model saved to {filepath}")
