"""Oracle: ``cv2.resize(src, (W, H), interpolation=cv2.INTER_CUBIC)`` restated in numpy.

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Reference call sites: classic_algorithms.py:11-13 (``interpolate_bicubic``),
loading_methods.py:147-148, SRCNN_model.py:191.  The arithmetic lives in OpenCV
(modules/imgproc/src/resize.cpp; unpinned by the reference, 4.13.0 in this image) and is
restated from its published algorithm.  ``cv2_resize`` below calls the real thing and is
what the restatement is pinned against in tests/test_oracle_bicubic.py.

Algorithm (per axis, separable, horizontal pass first):
  scale = 1 / (dst / src)  (double);  f = (d + 0.5) * scale - 0.5;  s = floor(f);
  t = f - s;  Keys cubic with A = -0.75 (OpenCV ``interpolateCubic``); taps s-1 .. s+2 with
  the *index* clamped to [0, n-1]; float32 result is not clipped.

OpenCV 4.13 has two observable behaviours (measured here, see DESIGN.md "bicubic"):

* ``mode="scalar"``  == ``cv2.setUseOptimized(False)``: f and t in float32, coefficients in
  plain float32; uint8 goes through 11-bit fixed-point coefficients (rint(c * 2048) as int16),
  an int32 horizontal pass and a float32 vertical pass
  ``sum_k float(row_k) * (float(beta_k) * 2^-22)``, round-half-even, saturate.  The
  restatement is bit-exact against it (uint8) / <= 2.4e-7 (float32).
* ``mode="default"`` == OpenCV's default dispatch (what a reference user gets): t is
  position-independent (computed in double), float32 results agree to <= 1e-6, and uint8
  output equals ``saturate(rint(float32 path))`` (so it is NOT the fixed-point path for
  non-dyadic ratios).  The restatement matches uint8 to <= 1 LSB on < 0.5 % of pixels for x3 /
  non-integer ratios and bit-exactly on all but ~5e-6 of pixels for x2 / x4.
"""
from __future__ import annotations

import numpy as np

_A = np.float32(-0.75)
_f32, _f64 = np.float32, np.float64


def _fma(a, b, c):
    """float32 fused multiply-add (exact product in float64, one rounding)."""
    return (np.asarray(a, _f64) * np.asarray(b, _f64) + np.asarray(c, _f64)).astype(_f32)


def cubic_coeffs(t):
    """OpenCV interpolateCubic, plain float32 arithmetic in source order."""
    t = np.asarray(t, dtype=_f32)
    one = _f32(1.0)
    x1 = t + one
    c0 = ((_A * x1 - _f32(5) * _A) * x1 + _f32(8) * _A) * x1 - _f32(4) * _A
    c1 = ((_A + _f32(2)) * t - (_A + _f32(3))) * t * t + one
    u = one - t
    c2 = ((_A + _f32(2)) * u - (_A + _f32(3))) * u * u + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], axis=-1).astype(_f32)


def cubic_coeffs_fma(t):
    """Same polynomial with the multiply-adds contracted to FMAs (what nvcc emits for the
    CUDA kernel and, to within 2.3e-7, what OpenCV's default build evaluates)."""
    t = np.asarray(t, dtype=_f32)
    one = _f32(1.0)
    a = np.full_like(t, _A)
    x1 = (t + one).astype(_f32)
    c0 = _fma(_fma(_fma(a, x1, _f32(-5) * _A), x1, _f32(8) * _A), x1, _f32(-4) * _A)
    a2 = np.full_like(t, _A + _f32(2))
    c1 = _fma((_fma(a2, t, -(_A + _f32(3))) * t).astype(_f32), t, one)
    u = (one - t).astype(_f32)
    c2 = _fma((_fma(a2, u, -(_A + _f32(3))) * u).astype(_f32), u, one)
    c3 = (((one - c0).astype(_f32) - c1).astype(_f32) - c2).astype(_f32)
    return np.stack([c0, c1, c2, c3], axis=-1)


def axis_table(n_src, n_dst, mode="default"):
    """-> (idx [n_dst, 4] int64 clamped tap indices, coef [n_dst, 4] float32)."""
    scale = 1.0 / (float(n_dst) / float(n_src))
    d = np.arange(n_dst, dtype=_f64)
    if mode == "scalar":
        f = ((d + 0.5) * scale - 0.5).astype(_f32)
        s = np.floor(f).astype(np.int64)
        coef = cubic_coeffs((f - s.astype(_f32)).astype(_f32))
    elif mode == "default":
        f = (d + 0.5) * scale - 0.5
        s = np.floor(f).astype(np.int64)
        coef = cubic_coeffs_fma((f - s).astype(_f32))
    else:
        raise ValueError(mode)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, n_src - 1)
    return idx, coef


def _as_hwc(src):
    return (src[:, :, None], True) if src.ndim == 2 else (src, False)


def resize_cubic_f32(src, dsize, mode="default"):
    """src: HWC float32; dsize = (W, H) like cv2.  float32 out, unclipped."""
    src, squeeze = _as_hwc(np.asarray(src, dtype=_f32))
    dw, dh = int(dsize[0]), int(dsize[1])
    h, w, c = src.shape
    xi, xa = axis_table(w, dw, mode)
    yi, yb = axis_table(h, dh, mode)
    fused = mode == "default"
    rows = np.zeros((h, dw, c), dtype=_f32)
    for k in range(4):                      # horizontal pass, taps in order
        if fused:
            rows = _fma(src[:, xi[:, k], :], xa[None, :, k, None], rows)
        else:
            rows += src[:, xi[:, k], :] * xa[None, :, k, None]
    out = np.zeros((dh, dw, c), dtype=_f32)
    for k in range(4):                      # vertical pass
        if fused:
            out = _fma(rows[yi[:, k]], yb[:, k, None, None], out)
        else:
            out += rows[yi[:, k]] * yb[:, k, None, None]
    return out[:, :, 0] if squeeze else out


def _fix_coeffs(c):
    return np.clip(np.rint(c * _f32(2048.0)), -32768, 32767).astype(np.int32)


def resize_cubic_u8_fixed(src, dsize, mode="scalar", simd_lanes=0):
    """OpenCV's fixed-point uint8 path (INTER_RESIZE_COEF_BITS = 11)."""
    src, squeeze = _as_hwc(np.asarray(src, dtype=np.uint8))
    dw, dh = int(dsize[0]), int(dsize[1])
    h, w, c = src.shape
    xi, xa = axis_table(w, dw, mode)
    yi, yb = axis_table(h, dh, mode)
    ia, ib = _fix_coeffs(xa), _fix_coeffs(yb)
    s32 = src.astype(np.int32)
    rows = np.zeros((h, dw, c), dtype=np.int32)
    for k in range(4):
        rows += s32[:, xi[:, k], :] * ia[None, :, k, None]
    fb = ib.astype(_f32) * _f32(1.0 / (2048.0 * 2048.0))
    acc = np.zeros((dh, dw, c), dtype=_f32)
    for k in range(4):
        acc = acc + rows[yi[:, k]].astype(_f32) * fb[:, k, None, None]
    out = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    tail = (dw * c) % simd_lanes if simd_lanes else 0
    if tail:                                # OpenCV's scalar tail: (v + 2^21) >> 22
        iacc = np.zeros((dh, dw, c), dtype=np.int64)
        for k in range(4):
            iacc += rows[yi[:, k]].astype(np.int64) * ib[:, k, None, None]
        iout = np.clip((iacc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
        flat, iflat = out.reshape(dh, dw * c), iout.reshape(dh, dw * c)
        flat[:, dw * c - tail:] = iflat[:, dw * c - tail:]
        out = flat.reshape(dh, dw, c)
    return out[:, :, 0] if squeeze else out


def resize_cubic_u8(src, dsize, mode="default"):
    """uint8 in / uint8 out.  default: saturate(rint(float32 path)); scalar: fixed-point."""
    if mode == "scalar":
        return resize_cubic_u8_fixed(src, dsize, "scalar")
    f = resize_cubic_f32(np.asarray(src, dtype=np.uint8).astype(_f32), dsize, "default")
    return np.clip(np.rint(f), 0, 255).astype(np.uint8)


def resize_cubic(src, dsize, mode="default"):
    src = np.asarray(src)
    if src.dtype == np.uint8:
        return resize_cubic_u8(src, dsize, mode)
    return resize_cubic_f32(src, dsize, mode)


# ---- cv2.resize(..., INTER_LANCZOS4): classic_algorithms.py:19-21 (interpolate_lanczos), :58-62 -------------------
def lanczos4_coeffs(t):
    """OpenCV interpolateLanczos4: eight sinc weights from one sin/cos pair (double) through the pi/4 rotation table,
    1e30 at a zero argument, normalised in float32 by the reciprocal of the float32 sum."""
    t = np.asarray(t, dtype=_f32)
    s45 = 0.70710678118654752440084436210485
    cs = np.array([[1, 0], [-s45, -s45], [0, 1], [s45, -s45], [-1, 0], [s45, s45], [0, -1], [-s45, s45]], _f64)
    y0 = -(t.astype(_f64) + 3.0) * np.pi * 0.25
    s0, c0 = np.sin(y0), np.cos(y0)
    co = np.zeros(t.shape + (8,), _f32)
    for i in range(8):
        yi = ((t + _f32(3)).astype(_f32) - _f32(i)).astype(_f32)
        y = -yi.astype(_f64) * np.pi * 0.25
        with np.errstate(divide="ignore", invalid="ignore"):
            v = ((cs[i, 0] * s0 + cs[i, 1] * c0) / (y * y)).astype(_f32)
        co[..., i] = np.where(np.abs(yi) >= _f32(1e-6), v, _f32(1e30))
    total = np.zeros(t.shape, _f32)
    for i in range(8):
        total = (total + co[..., i]).astype(_f32)
    return (co * (_f32(1) / total).astype(_f32)[..., None]).astype(_f32)


def lanczos4_axis_table(n_src, n_dst):
    """-> (idx [n_dst, 8] clamped tap indices floor(f) - 3 .. floor(f) + 4, coef [n_dst, 8] float32); f, t in float32."""
    scale = 1.0 / (float(n_dst) / float(n_src))
    f = ((np.arange(n_dst, dtype=_f64) + 0.5) * scale - 0.5).astype(_f32)
    s = np.floor(f).astype(np.int64)
    coef = lanczos4_coeffs((f - s.astype(_f32)).astype(_f32))
    return np.clip(s[:, None] + np.arange(-3, 5)[None, :], 0, n_src - 1), coef


def resize_lanczos4_f32(src, dsize):
    """src: HWC float32; dsize = (W, H).  Horizontal pass first, taps accumulated in order; <= 5e-7 from cv2 4.13."""
    src, squeeze = _as_hwc(np.asarray(src, dtype=_f32))
    dw, dh = int(dsize[0]), int(dsize[1])
    h, w, c = src.shape
    xi, xa = lanczos4_axis_table(w, dw)
    yi, yb = lanczos4_axis_table(h, dh)
    rows = np.zeros((h, dw, c), dtype=_f32)
    for k in range(8):
        rows = _fma(src[:, xi[:, k], :], xa[None, :, k, None], rows)
    out = np.zeros((dh, dw, c), dtype=_f32)
    for k in range(8):
        out = _fma(rows[yi[:, k]], yb[:, k, None, None], out)
    return out[:, :, 0] if squeeze else out


# ---- uint8 INTER_LINEAR / INTER_AREA (up-scaling) / INTER_LANCZOS4: OpenCV's 11-bit fixed-point paths ----------------
# (classic_algorithms.py:7-9, 15-21 called on uint8 images by super_resolucion_clasica.ipynb cell 7)
def linear_axis_table_u8(n_src, n_dst, area=False, clamp_t=True):
    """-> (idx [n_dst, 2], icoef [n_dst, 2] = rint(c * 2048) as int).  OpenCV builds the x table with the fraction forced
    to 0 where the tap pair leaves the image (``clamp_t``); the y table keeps the fraction and clamps the ROW INDEX in the
    row loop instead, so a border row is blended with itself through both (separately truncated) products."""
    scale = 1.0 / (float(n_dst) / float(n_src))
    inv = float(n_dst) / float(n_src)
    d = np.arange(n_dst, dtype=_f64)
    if not area:
        f = ((d + 0.5) * scale - 0.5).astype(_f32)
        s = np.floor(f).astype(np.int64)
        t = (f - s.astype(_f32)).astype(_f32)
    else:
        s = np.floor(d * scale).astype(np.int64)
        t = ((d + 1) - (s + 1) * inv).astype(_f32)
        t = np.where(t <= 0, _f32(0), (t - np.floor(t)).astype(_f32)).astype(_f32)
    if clamp_t:
        lo = s < 0
        t, s = np.where(lo, _f32(0), t), np.where(lo, 0, s)
        hi = s >= n_src - 1
        t, s = np.where(hi, _f32(0), t), np.where(hi, n_src - 1, s)
    c = np.stack([(_f32(1) - t).astype(_f32), t.astype(_f32)], -1)
    idx = np.clip(np.stack([s, s + 1], -1), 0, n_src - 1)
    return idx, _fix_coeffs(c).astype(np.int64)


def resize_linear_u8(src, dsize, area=False):
    """cv2.resize(uint8, INTER_LINEAR) and - when the image grows - INTER_AREA, bit-exact (HResizeLinear int32 pass, then
    VResizeLinear<uchar>: ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2)."""
    src, squeeze = _as_hwc(np.asarray(src, dtype=np.uint8))
    dw, dh = int(dsize[0]), int(dsize[1])
    h, w, c = src.shape
    xi, xa = linear_axis_table_u8(w, dw, area, clamp_t=True)
    yi, yb = linear_axis_table_u8(h, dh, area, clamp_t=False)
    s = src.astype(np.int64)
    rows = s[:, xi[:, 0], :] * xa[None, :, 0, None] + s[:, xi[:, 1], :] * xa[None, :, 1, None]
    s0, s1 = rows[yi[:, 0]], rows[yi[:, 1]]
    b0, b1 = yb[:, 0, None, None], yb[:, 1, None, None]
    out = np.clip((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def resize_lanczos4_u8(src, dsize):
    """cv2.resize(uint8, INTER_LANCZOS4): coefficients rint(c * 2048) as int16, int32 horizontal pass, integer vertical pass
    (sum + 2^21) >> 22, saturate."""
    src, squeeze = _as_hwc(np.asarray(src, dtype=np.uint8))
    dw, dh = int(dsize[0]), int(dsize[1])
    h, w, c = src.shape
    xi, xa = lanczos4_axis_table(w, dw)
    yi, yb = lanczos4_axis_table(h, dh)
    ia, ib = _fix_coeffs(xa).astype(np.int64), _fix_coeffs(yb).astype(np.int64)
    s = src.astype(np.int64)
    rows = np.zeros((h, dw, c), np.int64)
    for k in range(8):
        rows += s[:, xi[:, k], :] * ia[None, :, k, None]
    acc = np.zeros((dh, dw, c), np.int64)
    for k in range(8):
        acc += rows[yi[:, k]] * ib[:, k, None, None]
    out = np.clip((acc + (1 << 21)) >> 22, 0, 255).astype(np.uint8)
    return out[:, :, 0] if squeeze else out


def cv2_resize(src, dsize, optimized=True):
    """The reference's actual implementation of this step (classic_algorithms.py:13)."""
    import cv2
    prev = cv2.useOptimized()
    cv2.setUseOptimized(bool(optimized))
    try:
        return cv2.resize(src, (int(dsize[0]), int(dsize[1])), interpolation=cv2.INTER_CUBIC)
    finally:
        cv2.setUseOptimized(prev)
