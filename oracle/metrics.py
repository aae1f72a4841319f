"""Oracle: ``tf.image.psnr`` / ``tf.image.ssim`` restated in numpy.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Reference call sites: /root/reference/SRModels/metrics.py:3-7
(``tf.image.psnr(y_true, y_pred, max_val=1.0)``, ``tf.image.ssim(..., max_val=1.0)`` with
TF 2.10 defaults filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03).  The algorithm lives
in tensorflow/python/ops/image_ops_impl.py (TensorFlow 2.10.0, not vendored, not
installable here) and is restated from its published definition:

* psnr = 20*log10(max_val) - 10*log10(mean((a-b)^2)) over (H, W, C), per image.
* ssim: Gaussian window g[i,j] = softmax(-(i^2+j^2)/(2 sigma^2)) on coords -5..5;
  mu_x, mu_y, E[xy], E[x^2+y^2] by VALID depthwise correlation;
  l = (2 mu_x mu_y + c1)/(mu_x^2 + mu_y^2 + c1);
  cs = (2 E[xy] - 2 mu_x mu_y + c2)/(E[x^2+y^2] - mu_x^2 - mu_y^2 + c2);
  mean over the (H-10) x (W-10) map, then over channels.  Returns ``[B]``.
"""
from __future__ import annotations

import numpy as np


def gaussian_taps(size=11, sigma=1.5, dtype=np.float64):
    """1-D factor of TF's _fspecial_gauss; the 2-D window is its outer product."""
    coords = np.arange(size, dtype=np.float64) - (size - 1) / 2.0
    g = np.exp(-0.5 * (coords / sigma) ** 2)
    return (g / g.sum()).astype(dtype)


def psnr(a, b, max_val=1.0, dtype=np.float32):
    a = np.asarray(a, dtype=dtype)
    b = np.asarray(b, dtype=dtype)
    if a.ndim == 3:
        a, b = a[None], b[None]
    mse = np.mean((a - b) ** 2, axis=(1, 2, 3), dtype=dtype)
    with np.errstate(divide="ignore"):
        return (20.0 * np.log10(dtype(max_val)) - 10.0 * np.log10(mse)).astype(dtype)


def _valid_filter(x, taps):
    """Separable VALID correlation of NHWC x with the outer product of taps."""
    k = len(taps)
    n, h, w, c = x.shape
    tmp = np.zeros((n, h - k + 1, w, c), dtype=x.dtype)
    for i in range(k):
        tmp += taps[i] * x[:, i:i + h - k + 1]
    out = np.zeros((n, h - k + 1, w - k + 1, c), dtype=x.dtype)
    for j in range(k):
        out += taps[j] * tmp[:, :, j:j + w - k + 1]
    return out


def _valid_filter_2d(x, taps):
    """Non-separable 121-tap form (what TF's depthwise_conv2d evaluates); used to confirm
    that the separable form is the same function up to rounding."""
    k = len(taps)
    g2 = np.outer(taps, taps)
    n, h, w, c = x.shape
    out = np.zeros((n, h - k + 1, w - k + 1, c), dtype=x.dtype)
    for i in range(k):
        for j in range(k):
            out += g2[i, j] * x[:, i:i + h - k + 1, j:j + w - k + 1]
    return out


def ssim(a, b, max_val=1.0, filter_size=11, filter_sigma=1.5, k1=0.01, k2=0.03,
         dtype=np.float32, separable=True):
    a = np.asarray(a, dtype=dtype)
    b = np.asarray(b, dtype=dtype)
    if a.ndim == 3:
        a, b = a[None], b[None]
    if a.shape[1] < filter_size or a.shape[2] < filter_size:
        raise ValueError(f"image dimensions must be at least {filter_size}x{filter_size}")
    taps = gaussian_taps(filter_size, filter_sigma, dtype)
    filt = _valid_filter if separable else _valid_filter_2d
    c1 = dtype((k1 * max_val) ** 2)
    c2 = dtype((k2 * max_val) ** 2)
    mu_a, mu_b = filt(a, taps), filt(b, taps)
    num0 = mu_a * mu_b * dtype(2.0)
    den0 = mu_a * mu_a + mu_b * mu_b
    lum = (num0 + c1) / (den0 + c1)
    num1 = filt(a * b, taps) * dtype(2.0)
    den1 = filt(a * a + b * b, taps)
    cs = (num1 - num0 + c2) / (den1 - den0 + c2)
    per_channel = np.mean(lum * cs, axis=(1, 2), dtype=dtype)
    return np.mean(per_channel, axis=-1, dtype=dtype).astype(dtype)


def evaluate_means(y_true, y_pred):
    """Keras ``Model.evaluate`` aggregation (SRCNN_model.py:100-109, EDSR_model.py:178-187):
    sample means of MSE loss, PSNR and SSIM -> [loss, psnr, ssim]."""
    y_true = np.asarray(y_true, dtype=np.float32)
    y_pred = np.asarray(y_pred, dtype=np.float32)
    loss = np.mean((y_true - y_pred) ** 2, dtype=np.float64)
    return [float(loss), float(np.mean(psnr(y_true, y_pred), dtype=np.float64)),
            float(np.mean(ssim(y_true, y_pred), dtype=np.float64))]


# ---- skimage.metrics definitions (super_resolucion_clasica.ipynb cell 7 lines 213-214 / 278-279, EDA.ipynb:235,254) ----
# scikit-image is unpinned by the reference and absent here ("parity unpinned"): restated from its published algorithm
# (skimage/metrics/_structural_similarity.py, simple_metrics.py), on the same scipy.ndimage.uniform_filter primitive
# skimage itself calls; pinned by known answers and an independent windowed evaluation in tests/test_oracle_metrics.py.
def skimage_psnr(image_true, image_test, data_range):
    """peak_signal_noise_ratio: 10 log10(data_range^2 / mean((a - b)^2)) in float64 over the whole array."""
    a = np.asarray(image_true, dtype=np.float64)
    b = np.asarray(image_test, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return float(10.0 * np.log10(float(data_range) ** 2 / np.mean((a - b) ** 2)))


def skimage_ssim(im1, im2, data_range, channel_axis=None, win_size=7, k1=0.01, k2=0.03):
    """structural_similarity with its defaults: uniform win_size^2 window (uniform_filter, reflect borders), sample
    covariance (N / (N - 1)), mean over the map cropped by (win_size - 1) // 2, then over channels."""
    from scipy.ndimage import uniform_filter
    a = np.asarray(im1, dtype=np.float64)
    b = np.asarray(im2, dtype=np.float64)
    if channel_axis is not None:
        a, b = np.moveaxis(a, channel_axis, -1), np.moveaxis(b, channel_axis, -1)
        return float(np.mean([skimage_ssim(a[..., c], b[..., c], data_range, None, win_size, k1, k2)
                              for c in range(a.shape[-1])]))
    if min(a.shape) < win_size:
        raise ValueError("win_size exceeds image extent")
    n = win_size ** a.ndim
    cov_norm = n / (n - 1.0)
    ux, uy = uniform_filter(a, size=win_size), uniform_filter(b, size=win_size)
    uxx, uyy, uxy = uniform_filter(a * a, size=win_size), uniform_filter(b * b, size=win_size), uniform_filter(a * b, size=win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
    pad = (win_size - 1) // 2
    return float(s[pad:a.shape[0] - pad, pad:a.shape[1] - pad].mean(dtype=np.float64))
