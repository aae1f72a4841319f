"""Composition of EDSR's activation-free up-sampling tail into one 5 x 5 convolution.

``EDSR_model.py:76-95, 117-123`` builds the tail as ``Conv2D(64 -> 256, 3x3) -> depth_to_space(2)`` (once for x2, twice
for x4; ``64 -> 576`` and ``depth_to_space(3)`` for x3) followed by ``Conv2D(64 -> channels, 3x3)`` with **no activation in
between**, so the whole tail is one linear map from the 64-channel low-resolution feature image to the r x r x channels
block of every low-resolution pixel.  Its footprint is 5 x 5 low-resolution pixels:

    out[b, r*y + i, r*x + j, c] = bias[(i*r + j)*C + c]
                                  + sum over qy, qx in 0..4, ci of  X[b, y + qy - 2, x + qx - 2, ci] * Wc[qy, qx, ci, (i*r + j)*C + c]

with 153,600 FLOP per low-resolution pixel at x4 instead of 1,529,856 for the three layers, and neither of the two 64-channel
full-resolution intermediates (2.4 GB per 32 tiles of 192 x 192) is ever written.

The zero padding of the *intermediate* layers is what makes this more than a product of kernels: ``same`` padding
zero-fills each intermediate image outside its borders, whereas a composed kernel would implicitly evaluate the inner
layers there.  The exclusions only reach low-resolution pixels on the image border (row 0 / H-1, column 0 / W-1), so the
exact map has nine variants: (top, interior, bottom) x (left, interior, right).  They are obtained by probing the exact
float64 layer-by-layer map on a 5 x 5 image (every class has a representative there: rows / columns 0, 2 and 4) with unit
impulses, which needs no case analysis and is exact by linearity; the bias vector of each variant is the response to the
zero image.  The natural zero padding of the tail's *input* is the TMA out-of-bounds fill of the kernel as usual.

Host-side weight preprocessing (like the K-major repacking in ``srb_conv_weights_create``); the result feeds
``srb_upsampler_create`` and the convolution itself runs in ``upsample5_fold_kernel`` on the tensor cores.
"""
from __future__ import annotations

import hashlib

import numpy as np

FOOT = 5          # footprint of the composed kernel in low-resolution pixels
_PROBE = 5        # probe image size: rows / columns 0, 2, 4 represent border / interior / border


def _conv_same(x, k, b=None):
    """Keras Conv2D(padding="same", stride 1) in float64: x [N, H, W, Cin], k HWIO."""
    kh, kw = k.shape[:2]
    n, h, w, _ = x.shape
    xp = np.pad(x, ((0, 0), (kh // 2, kh // 2), (kw // 2, kw // 2), (0, 0)))
    y = np.zeros((n, h, w, k.shape[3]))
    for dy in range(kh):
        for dx in range(kw):
            y += xp[:, dy:dy + h, dx:dx + w, :] @ k[dy, dx]
    return y if b is None else y + b


def _depth_to_space(x, r):
    """tf.nn.depth_to_space (DCR): out[n, y*r + i, x*r + j, c] = in[n, y, x, (i*r + j)*C + c]."""
    n, h, w, c = x.shape
    co = c // (r * r)
    return x.reshape(n, h, w, r, r, co).transpose(0, 1, 3, 2, 4, 5).reshape(n, h * r, w * r, co)


def _stages(weights, scale_factor):
    names = ["up0"] if scale_factor in (2, 3) else ["up0", "up1"]
    shuffles = [scale_factor] if scale_factor in (2, 3) else [2, 2]
    return names, shuffles


def layered_tail(weights, x, scale_factor, with_bias=True):
    """The tail exactly as the reference builds it, in float64 (no clip): x [N, H, W, 64] -> [N, H*r, W*r, C]."""
    names, shuffles = _stages(weights, scale_factor)
    x = np.asarray(x, dtype=np.float64)
    for name, r in zip(names, shuffles):
        b = weights.get(name + "/bias") if with_bias else None
        x = _depth_to_space(_conv_same(x, np.asarray(weights[name + "/kernel"], np.float64),
                                       None if b is None else np.asarray(b, np.float64)), r)
    b = weights.get("tail/bias") if with_bias else None
    return _conv_same(x, np.asarray(weights["tail/kernel"], np.float64), None if b is None else np.asarray(b, np.float64))


_CACHE = {}          # fingerprint of the tail's layers -> (w, bias); a handful of entries (one per distinct weight set)


def _fingerprint(weights, scale_factor):
    h = hashlib.sha1(str(int(scale_factor)).encode())
    names, _ = _stages(weights, scale_factor)
    for name in names + ["tail"]:
        for part in ("/kernel", "/bias"):
            a = weights.get(name + part)
            h.update(b"-" if a is None else np.ascontiguousarray(a, dtype=np.float32).tobytes())
    return h.hexdigest()


def compose_edsr_tail(weights, scale_factor):
    """-> (w [3, 3, 5, 5, cin, r*r*C] float64, bias [3, 3, r*r*C] float64): variant (vy, vx) with vy / vx = 0 for the first
    row / column of the image, 1 for the interior, 2 for the last; output channel (i*r + j)*C + c.  Results are cached by
    a hash of the layers' values (the float64 probing takes 1-3 s)."""
    key = _fingerprint(weights, scale_factor)
    if key not in _CACHE:
        if len(_CACHE) >= 8:
            _CACHE.pop(next(iter(_CACHE)))
        _CACHE[key] = _compose(weights, scale_factor)
    return _CACHE[key]


def _compose(weights, scale_factor):
    r = int(scale_factor)
    cin = int(np.asarray(weights["up0/kernel"]).shape[2])
    c_img = int(np.asarray(weights["tail/kernel"]).shape[3])
    cout = r * r * c_img
    P = _PROBE

    def blocks(y):      # [N, P*r, P*r, C] -> [N, P, P, r*r*C] (the depth_to_space channel order)
        n = y.shape[0]
        return y.reshape(n, P, r, P, r, c_img).transpose(0, 1, 3, 2, 4, 5).reshape(n, P, P, cout)

    bias_map = blocks(layered_tail(weights, np.zeros((1, P, P, cin)), r, with_bias=True))[0]
    w = np.zeros((3, 3, FOOT, FOOT, cin, cout))
    eye = np.eye(cin)
    for py in range(P):
        for px in range(P):
            x = np.zeros((cin, P, P, cin))
            x[:, py, px, :] = eye                       # image n = unit impulse in channel n at (py, px)
            resp = blocks(layered_tail(weights, x, r, with_bias=False))   # [cin, P, P, cout]
            for vy in range(3):
                for vx in range(3):
                    qy, qx = py - 2 * vy + 2, px - 2 * vx + 2      # tap of the variant's pixel (2 vy, 2 vx) that reads (py, px)
                    if 0 <= qy < FOOT and 0 <= qx < FOOT:
                        w[vy, vx, qy, qx] = resp[:, 2 * vy, 2 * vx, :]
    bias = np.stack([np.stack([bias_map[2 * vy, 2 * vx] for vx in range(3)]) for vy in range(3)])
    return w, bias


def apply_composed(w, bias, x, scale_factor):
    """Float64 evaluation of the composed map (host check of the composition; the product path runs the CUDA kernel):
    x [N, H, W, cin] -> [N, H*r, W*r, C].  Needs H, W >= 2."""
    r = int(scale_factor)
    x = np.asarray(x, dtype=np.float64)
    n, h, wd, _ = x.shape
    if h < 2 or wd < 2:
        raise ValueError("the composed up-sampler needs images of at least 2 x 2 pixels")
    cout = w.shape[-1]
    out = np.zeros((n, h, wd, cout))
    cls_y = np.ones(h, dtype=int); cls_y[0] = 0; cls_y[-1] = 2
    cls_x = np.ones(wd, dtype=int); cls_x[0] = 0; cls_x[-1] = 2
    for vy in range(3):
        for vx in range(3):
            full = _conv_same(x, w[vy, vx]) + bias[vy, vx]
            mask = (cls_y[:, None] == vy) & (cls_x[None, :] == vx)
            out[:, mask] = full[:, mask]
    return _depth_to_space(out, r)


def weight_scale(w):
    """Power of two that brings the largest composed weight into [128, 256): the 16-bit copies then keep full relative
    precision for weights down to 2^-22 of the largest; the kernel multiplies the accumulators by the inverse."""
    m = float(np.abs(w).max())
    if not np.isfinite(m) or m == 0.0:
        return 1.0
    return float(2.0 ** (7 - int(np.floor(np.log2(m)))))
