"""``ESRGAN`` generator inference (SRModels/deep_learning_models/ESRGAN_model.py).

Kept: generator forward (:212-345 incl. the two SelfAttention layers :30-79), ``super_resolve_image``
(:858-979) with its [-1,1] mapping, ``evaluate`` (:782-856) with its mean-of-per-batch-means aggregation and dict keys.
Discriminator, VGG19 perceptual loss, FFT loss and the GAN loop are training-only and out of scope."""
from __future__ import annotations

import numpy as np

from .. import _capi as capi
from .. import engine, ops, weights as W
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

    def evaluate(self, test_dataset):
        """Evaluate the generator on an iterable of ``(lr_batch, hr_batch)`` pairs in [-1, 1] (a ``tf.data.Dataset`` in the
        reference; any iterable of NHWC numpy arrays / CUDA tensors here).

        Aggregation is the reference's (ESRGAN_model.py:828-843): PSNR and SSIM are averaged per batch, and the per-batch
        means are averaged over batches - NOT the sample mean Keras ``evaluate`` reports for SRCNN / EDSR (the two differ
        when the last batch is short).  Returns ``{"avg_psnr", "avg_ssim", "avg_g_loss"}``.  ``avg_g_loss`` needs the
        discriminator and the VGG19 perceptual network (training-only, outside this path): it is ``nan`` here, and the one
        term of it this path can evaluate, the L1 pixel loss on the [-1, 1] images (:433-445), is returned as the extra
        key ``avg_pixel_loss``.  All sums stay on the device; one read-back at the end."""
        if not self.trained:
            raise RuntimeError("Model has not been trained.")
        torch = capi.require_cuda()
        print("Evaluating model on test dataset...")
        totals = torch.zeros(3, dtype=torch.float64, device="cuda")     # sum of per-batch mean psnr / ssim / pixel loss
        num_batches = 0
        for lr_batch, hr_batch in test_dataset:
            lr = self._as_batch(lr_batch)
            hr = self._as_batch(hr_batch)
            gen = self.generator.predict_device(lr)
            if tuple(gen.shape) != tuple(hr.shape):
                raise ValueError(f"generator output {tuple(gen.shape)} does not match the HR batch {tuple(hr.shape)}")
            p, q = ops.psnr_ssim(ops.cast(hr, torch.float32, 0.5, 0.5), ops.cast(gen, torch.float32, 0.5, 0.5), 1.0)
            totals[0] += p.double().mean()
            totals[1] += q.double().mean()
            totals[2] += (hr - gen).abs().double().mean()
            num_batches += 1
        if num_batches == 0:
            raise ValueError("test_dataset is empty")
        t = (totals / num_batches).cpu().numpy()
        metrics = {"avg_psnr": float(t[0]), "avg_ssim": float(t[1]), "avg_g_loss": float("nan"),
                   "avg_pixel_loss": float(t[2])}
        print("Evaluation Results:")
        print(f"  Average PSNR: {metrics['avg_psnr']:.4f}")
        print(f"  Average SSIM: {metrics['avg_ssim']:.4f}")
        print(f"  Average G Loss: {metrics['avg_g_loss']:.4f}")
        return metrics

    @staticmethod
    def _as_batch(a):
        torch = capi.require_cuda()
        if isinstance(a, torch.Tensor):
            t = a if a.is_cuda else a.cuda()
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.float32)).cuda()
        if t.dim() != 4:
            raise ValueError(f"expected NHWC batches, got shape {tuple(t.shape)}")
        return t.float().contiguous()

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
