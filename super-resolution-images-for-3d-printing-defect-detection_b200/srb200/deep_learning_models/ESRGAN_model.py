"""``ESRGAN`` generator inference (SRModels/deep_learning_models/ESRGAN_model.py).

Kept: generator forward (:212-345 incl. the two SelfAttention layers :30-79), ``super_resolve_image``
(:858-979) with its [-1,1] mapping.  Discriminator, VGG19 perceptual loss, FFT loss and the GAN loop
are training-only and out of scope."""
from __future__ import annotations

from .. import engine, weights as W
from . import _common as common


class ESRGAN:
    def __init__(self):
        self.generator = None
        self.scale_factor = None
        self.trained = False
        self._arch = {}

    def setup_model(self, scale_factor=2, growth_channels=32, num_rrdb_blocks=23, input_shape=None, output_shape=None,
                    from_trained=False, generator_pretrained_path=None, precision="fp32", seed=1234, **_ignored):
        self.scale_factor = scale_factor
        self._arch = dict(scale_factor=scale_factor, growth_channels=growth_channels, num_rrdb_blocks=num_rrdb_blocks)
        if from_trained:
            w = common.load_weight_file(generator_pretrained_path)
            self.trained = True
        else:
            w = W.esrgan_generator_weights(scale_factor, growth_channels, num_rrdb_blocks, seed=seed)
        self.generator = engine.ESRGANGeneratorNet(w, precision=precision, **self._arch)

    def load_weights(self, weights, precision=None):
        self.generator = engine.ESRGANGeneratorNet(weights, precision=precision or self.generator.precision, **self._arch)
        self.trained = True

    def super_resolve_image(self, lr_img, patch_size_lr=48, stride=24, batch_size=16):
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        if self.scale_factor is None:
            raise ValueError("scale_factor is not set. Call setup_model first.")
        img = common.as_device_image(lr_img)
        sr, metrics = common.tiled_super_resolve(self.generator, img, patch_size_lr, stride, self.scale_factor,
                                                 pre=(2.0, -1.0), post=(0.5, 0.5))
        return sr.cpu().numpy(), metrics

    def super_resolve_image_whole(self, lr_img):
        """Fast path without tiling: one pass of the generator over the whole LR image ([-1, 1] mapping as in the tiled
        flow).  The two SelfAttention layers then attend over the whole image instead of one patch."""
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        if self.scale_factor is None:
            raise ValueError("scale_factor is not set. Call setup_model first.")
        sr, metrics = common.whole_image_super_resolve(self.generator, common.as_device_image(lr_img),
                                                       pre=(2.0, -1.0), post=(0.5, 0.5))
        return sr.cpu().numpy(), metrics
