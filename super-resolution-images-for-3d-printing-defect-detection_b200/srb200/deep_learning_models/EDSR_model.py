"""``EDSR`` with the reference's inference surface (SRModels/deep_learning_models/EDSR_model.py).

Kept: ``setup_model`` (:29-53), ``evaluate`` (:178-187), ``super_resolve_image`` (:189-315), the
``trained`` gate and the exceptions.  ``fit`` is out of scope.
"""
from __future__ import annotations

import os

import numpy as np

from .. import engine, weights as W
from . import _common as common


class EDSR:
    def __init__(self):
        self.model = None
        self.scale_factor = None
        self.trained = False
        self._arch = {}

    def setup_model(self, scale_factor=2, channels=3, num_res_blocks=16, num_filters=64, res_scaling=0.1,
                    learning_rate=1e-4, loss="mean_absolute_error", from_pretrained=False, pretrained_path=None,
                    precision="fp16", seed=1234):
        """Set up the EDSR model, either by loading pretrained weights (.npz) or building a new one."""
        self.scale_factor = scale_factor
        self._arch = dict(scale_factor=scale_factor, num_res_blocks=num_res_blocks, res_scaling=res_scaling)
        if from_pretrained:
            w = common.load_weight_file(pretrained_path)
            n_named = sum(1 for k in w if k.endswith("_c1/kernel"))
            if n_named:
                self._arch["num_res_blocks"] = n_named
            else:
                # a file exported from the reference's Keras model (auto-named conv2d, conv2d_1, ...): head + 2 per block +
                # body end + one up-sampling conv per x2 stage (two for x4) + tail (EDSR_model.py:96-125)
                n_conv = len(W.keras_conv_layers(W.normalize_keras_names(w)))
                n_fixed = 3 + (2 if scale_factor == 4 else 1)
                if n_conv < n_fixed or (n_conv - n_fixed) % 2:
                    raise ValueError(f"{pretrained_path}: {n_conv} Conv2D layers do not form an EDSR x{scale_factor} graph")
                self._arch["num_res_blocks"] = (n_conv - n_fixed) // 2
            self.model = engine.EDSRNet(w, precision=precision, **self._arch)
            self.trained = True
            print(f"Loaded pretrained model from {pretrained_path}")
        else:
            w = W.edsr_weights(scale_factor, channels, num_res_blocks, num_filters, seed=seed)
            self.model = engine.EDSRNet(w, precision=precision, **self._arch)
            self.model.summary()

    def load_weights(self, weights: dict, precision=None):
        if self.scale_factor is None:
            raise ValueError("scale_factor is not set. Call setup_model first.")
        self.model = engine.EDSRNet(weights, precision=precision or self.model.precision, **self._arch)
        self.trained = True

    def fit(self, *a, **k):
        raise NotImplementedError("training is outside the B200 inference path (SURVEY.md section 2)")

    def evaluate(self, X_test, Y_test):
        """Evaluate the model on test data and print loss, PSNR, and SSIM."""
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        sums = common.evaluate_arrays(self.model, X_test, Y_test)
        results = common.finish_evaluation(sums)
        print(f"Loss: {results[0]:.4f}, PSNR: {results[1]:.2f} dB, SSIM: {results[2]:.4f}")
        return results

    def super_resolve_image(self, lr_img, patch_size_lr=48, stride=24):
        """Patch-based SR of an in-memory LR array: pad, extract LR patches, predict HR patches,
        overlap-average, crop to (h*scale, w*scale).  Returns (sr_img float32 [0,1], inference_metrics)."""
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        if self.scale_factor is None:
            raise ValueError("scale_factor is not set. Call setup_model first.")
        img = common.as_device_image(lr_img)
        sr, metrics = common.tiled_super_resolve(self.model, img, patch_size_lr, stride, self.scale_factor)
        return sr.cpu().numpy(), metrics

    def super_resolve_image_whole(self, lr_img):
        """Fast path without tiling: one fully-convolutional pass over the whole LR image (see
        ``_common.whole_image_super_resolve`` for how it relates to the tiled flow).  Same return convention."""
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        if self.scale_factor is None:
            raise ValueError("scale_factor is not set. Call setup_model first.")
        sr, metrics = common.whole_image_super_resolve(self.model, common.as_device_image(lr_img))
        return sr.cpu().numpy(), metrics

    def super_resolve_batch(self, lr_batch):
        """Whole-image fully-convolutional fast path (the network's input is (None, None, 3),
        EDSR_model.py:98): NHWC float32 batch in, NHWC float32 batch out, no tiling."""
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        return self.model.predict(lr_batch)

    def save(self, directory, timestamp):
        if not self.trained:
            raise RuntimeError("Cannot save an untrained model.")
        if not directory:
            raise ValueError("Directory path must be provided.")
        os.makedirs(directory, exist_ok=True)
        path = os.path.join(directory, f"EDSR_x{self.scale_factor}_{timestamp}.npz")
        common.save_weight_file(path, self.model.get_weights_dict())
        print(f"Model saved to {path}")
