"""SRResNet / SRGAN generator x4 (BASELINE config 4).  Not in the reference: composed from its layer
semantics with BatchNorm folded; same wrapper surface as ``EDSR``."""
from __future__ import annotations

from .. import engine, weights as W
from . import _common as common


class SRResNet:
    def __init__(self):
        self.model = None
        self.scale_factor = None
        self.trained = False
        self._blocks = 16

    def setup_model(self, scale_factor=4, channels=3, num_res_blocks=16, num_filters=64, from_pretrained=False,
                    pretrained_path=None, precision="fp16", seed=1234):
        self.scale_factor, self._blocks = scale_factor, num_res_blocks
        if from_pretrained:
            w = common.load_weight_file(pretrained_path)
            self.trained = True
        else:
            w = W.srresnet_weights(scale_factor, channels, num_res_blocks, num_filters, seed=seed)
        self.model = engine.SRResNetNet(w, scale_factor, num_res_blocks, precision)

    def load_weights(self, weights, precision=None):
        self.model = engine.SRResNetNet(weights, self.scale_factor, self._blocks, precision or self.model.precision)
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
