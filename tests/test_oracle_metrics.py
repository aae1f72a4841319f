"""Analytic known-answers for the tf.image.psnr / ssim restatement."""
import os

import numpy as np
import pytest

from oracle import metrics as om


def test_gaussian_window():
    t = om.gaussian_taps()
    assert abs(t.sum() - 1) < 1e-12
    assert np.allclose(t[:6], [0.00102838, 0.00759876, 0.03600077, 0.10936069, 0.21300554, 0.26601172], atol=1e-8)
    assert abs(np.outer(t, t)[5, 5] - 0.07076224) < 1e-8


def test_constant_images():
    a = np.full((2, 16, 20, 3), 0.3, np.float32)
    b = np.full((2, 16, 20, 3), 0.6, np.float32)
    assert np.allclose(om.ssim(a, b, dtype=np.float64), 0.80004443, atol=1e-8)
    # float32 (what tf.image.ssim computes in): E[xy] - mu_x*mu_y cancels against c2 = 9e-4, so
    # a zero-variance image carries ~1e-4 of rounding noise in any float32 evaluation order.
    assert np.allclose(om.ssim(a, b), 0.80004443, atol=2e-4)
    assert np.allclose(om.psnr(a, b), 10.4575749, atol=1e-4)
    assert np.allclose(om.ssim(a, a), 1.0, atol=1e-6)
    assert np.all(np.isinf(om.psnr(a, a)))


def test_small_images_raise():
    a = np.zeros((1, 10, 32, 3), np.float32)
    with pytest.raises(ValueError):
        om.ssim(a, a)


def test_separable_equals_2d_and_fp64(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_oracle.npz"))
    a, b = g["a"], g["b"]
    s_sep = om.ssim(a, b)
    s_2d = om.ssim(a, b, separable=False)
    assert np.abs(s_sep - s_2d).max() < 2e-6
    assert np.abs(s_sep - g["ssim64"]).max() < 1e-5
    assert np.abs(om.psnr(a, b) - g["psnr64"]).max() < 1e-4
    assert np.array_equal(s_sep, g["ssim32"]) and np.array_equal(om.psnr(a, b), g["psnr32"])


def test_evaluate_means_shape():
    rng = np.random.default_rng(0)
    a = rng.random((3, 12, 12, 3), dtype=np.float32)
    b = rng.random((3, 12, 12, 3), dtype=np.float32)
    loss, p, s = om.evaluate_means(a, b)
    assert abs(loss - np.mean((a - b) ** 2)) < 1e-7 and -1 < s < 1 and p > 0


def test_ssim_against_independent_scipy_evaluation():
    """Independent code path: the Wang et al. SSIM with tf.image.ssim's constants, written directly on
    scipy.ndimage's separable correlation (no shared code with oracle/metrics.py), float64.  Pins the restatement's
    window, VALID cropping, c1 / c2 and the (mean over pixels, then over channels) reduction."""
    from scipy import ndimage
    rng = np.random.default_rng(3)
    a = rng.random((2, 40, 37, 3))
    b = np.clip(a + 0.1 * rng.standard_normal(a.shape), 0, 1)
    x = np.arange(11) - 5.0
    w = np.exp(-x * x / (2 * 1.5 ** 2)); w /= w.sum()

    def filt(img):                                   # [H, W] -> VALID 11x11 Gaussian
        f = ndimage.correlate1d(ndimage.correlate1d(img, w, axis=0, mode="constant"), w, axis=1, mode="constant")
        return f[5:-5, 5:-5]
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    want = []
    for i in range(a.shape[0]):
        per_c = []
        for c in range(3):
            p, q = a[i, :, :, c], b[i, :, :, c]
            mp, mq = filt(p), filt(q)
            spp, sqq, spq = filt(p * p) - mp * mp, filt(q * q) - mq * mq, filt(p * q) - mp * mq
            per_c.append((((2 * mp * mq + c1) * (2 * spq + c2)) / ((mp * mp + mq * mq + c1) * (spp + sqq + c2))).mean())
        want.append(np.mean(per_c))
    got = om.ssim(a.astype(np.float32), b.astype(np.float32), dtype=np.float64)
    assert np.abs(got - np.array(want)).max() < 1e-6
    mse = ((a.astype(np.float32).astype(np.float64) - b.astype(np.float32).astype(np.float64)) ** 2).mean(axis=(1, 2, 3))
    assert np.abs(om.psnr(a.astype(np.float32), b.astype(np.float32), dtype=np.float64) - (-10 * np.log10(mse))).max() < 1e-9


def test_skimage_definitions_known_answers_and_windowed_evaluation():
    """skimage.metrics restatement (super_resolucion_clasica.ipynb cell 7): known answers plus a direct evaluation of
    every valid 7 x 7 window with numpy's own sample (co)variance."""
    a = np.full((16, 20), 0.3)
    b = np.full((16, 20), 0.6)
    assert abs(om.skimage_ssim(a, b, 1.0) - 0.80004443) <= 1e-7
    assert abs(om.skimage_psnr(a, b, 1.0) - 10.4575749) <= 1e-6
    rng = np.random.default_rng(5)
    x = rng.random((19, 23, 3))
    y = np.clip(x + 0.1 * rng.standard_normal(x.shape), 0, 1)
    assert om.skimage_ssim(x, x, 1.0, channel_axis=2) == 1.0 and np.isinf(om.skimage_psnr(x, x, 1.0))
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    per_channel = []
    for c in range(3):
        vals = []
        for i in range(19 - 6):
            for j in range(23 - 6):
                wx, wy = x[i:i + 7, j:j + 7, c].ravel(), y[i:i + 7, j:j + 7, c].ravel()
                cov = np.cov(wx, wy, ddof=1)
                mx, my = wx.mean(), wy.mean()
                vals.append((2 * mx * my + c1) * (2 * cov[0, 1] + c2) / ((mx * mx + my * my + c1) * (cov[0, 0] + cov[1, 1] + c2)))
        per_channel.append(np.mean(vals))
    assert abs(om.skimage_ssim(x, y, 1.0, channel_axis=2) - np.mean(per_channel)) <= 1e-12
    # uint8 with data_range = 255 (EDA.ipynb) is the same function of the scaled images
    xu, yu = np.rint(x * 255).astype(np.uint8), np.rint(y * 255).astype(np.uint8)
    assert abs(om.skimage_ssim(xu, yu, 255, channel_axis=2) - om.skimage_ssim(xu / 255.0, yu / 255.0, 1.0, channel_axis=2)) <= 1e-12
    assert abs(om.skimage_psnr(xu, yu, 255) - om.skimage_psnr(xu / 255.0, yu / 255.0, 1.0)) <= 1e-9
    with pytest.raises(ValueError):
        om.skimage_ssim(np.zeros((6, 9)), np.zeros((6, 9)), 1.0)
