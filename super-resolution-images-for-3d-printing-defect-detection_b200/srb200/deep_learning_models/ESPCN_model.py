"""ESPCN x r (BASELINE config 2).  Not in the reference: composed from its Conv2D(padding="same") +
``tf.nn.depth_to_space`` layer semantics, with the same wrapper surface
as ``EDSR``."""
from __future__ import annotations

from .. import engine, weights as W
from . import _common as common


class ESPCN:
    def __init__(self):
        self.model = None
        self.scale_factor = None
        self.trained = False

    def setup_model(self, scale_factor=4, channels=3, activation="relu", from_pretrained=False, pretrained_path=None,
                    precision="fp16", seed=1234):
        self.scale_factor = scale_factor
        if from_pretrained:
            w = common.load_weight_file(pretrained_path)
            self.trained = True
        else:
            w = W.espcn_weights(scale_factor, channels, seed=seed)
        self.model = engine.ESPCNNet(w, scale_factor, activation, precision)

    def load_weights(self, weights, precision=None):
        self.model = engine.ESPCNNet(weights, self.scale_factor, self.model.activation, precision or self.model.precision)
        self.trained = True

    def evaluate(self, X_test, Y_test):
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        results = common.finish_evaluation(common.evaluate_arrays(self.model, X_test, Y_test))
        print(f"Loss: {results[0]:.4f}, PSNR: {results[1]:.2f} dB, SSIM: {results[2]:.4f}")
        return results

    def super_resolve_image(self, lr_img, patch_size_lr=48, stride=24):
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        sr, metrics = common.tiled_super_resolve(self.model, common.as_device_image(lr_img), patch_size_lr, stride,
                                                 self.scale_factor)
        return sr.cpu().numpy(), metrics
