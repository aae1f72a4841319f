"""GPU parity: bicubic kernel vs cv2.resize(INTER_CUBIC) golden outputs and the numpy restatement."""
import os

import numpy as np
import pytest

from oracle import bicubic as ob

pytestmark = pytest.mark.gpu


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "bicubic_cv2.npz"))
    for n in range(int(g["n_cases"])):
        h, w, dh, dw = (int(v) for v in g[f"c{n}_shape"])
        yield n, g, (dw, dh)


def test_float32_vs_cv2_golden(golden_dir):
    from srb200.classic_super_resolution_algorithms.classic_algorithms import interpolate_bicubic
    for n, g, dsize in _cases(golden_dir):
        out = interpolate_bicubic(g[f"c{n}_f32_in"], dsize)
        assert out.dtype == np.float32 and out.shape == g[f"c{n}_f32_default"].shape
        assert np.abs(out - g[f"c{n}_f32_default"]).max() <= 2e-6, n      # OpenCV default dispatch
        assert np.abs(out - g[f"c{n}_f32_scalar"]).max() <= 2e-6, n


def test_uint8_vs_cv2_golden(golden_dir):
    from srb200.classic_super_resolution_algorithms.classic_algorithms import interpolate_bicubic, resize_cubic
    for n, g, dsize in _cases(golden_dir):
        src = g[f"c{n}_u8_in"]
        fixed = resize_cubic(src, dsize, fixed_point=True)
        assert fixed.dtype == np.uint8
        assert np.array_equal(fixed, g[f"c{n}_u8_scalar"]), n                # bit-exact fixed-point path
        dflt = interpolate_bicubic(src, dsize)
        diff = np.abs(dflt.astype(np.int32) - g[f"c{n}_u8_default"].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 0.01, n              # <= 1 LSB on < 1 % of pixels


@pytest.mark.parametrize("shape,dsize", [((64, 64, 3), (128, 128)), ((239, 239, 3), (478, 478)),
                                         ((100, 75, 3), (225, 300)), ((341, 341, 3), (1024, 1024)),
                                         ((50, 60, 1), (240, 200)), ((30, 20, 4), (35, 41))])
def test_against_restatement(shape, dsize):
    from srb200.classic_super_resolution_algorithms.classic_algorithms import interpolate_bicubic, resize_cubic
    rng = np.random.default_rng(shape[0])
    f = rng.random(shape, dtype=np.float32)
    assert np.abs(interpolate_bicubic(f, dsize) - ob.resize_cubic_f32(f, dsize, "default")).max() <= 1e-6
    u = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(resize_cubic(u, dsize, fixed_point=True), ob.resize_cubic_u8(u, dsize, "scalar"))


def test_clip_batch_and_identity():
    import torch
    from srb200 import ops
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.random((3, 20, 24, 3), dtype=np.float32)).cuda()
    y = ops.bicubic(x, 40, 48)
    yc = ops.bicubic(x, 40, 48, clip01=True)
    assert y.min() < 0 or y.max() > 1                       # cubic overshoot exists and is kept
    assert torch.equal(yc, y.clamp(0, 1))
    for i in range(3):
        assert torch.equal(ops.bicubic(x[i:i + 1], 40, 48)[0], y[i])
    assert torch.allclose(ops.bicubic(x, 20, 24), x, atol=1e-6)   # same size = identity


def test_full_size_linearity():
    """BASELINE config 5 size: resampling is linear, so B(a*x + b*y) == a*B(x) + b*B(y)."""
    import torch
    from srb200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand((1, 1080, 1920, 3), device="cuda", generator=g)
    y = torch.rand((1, 1080, 1920, 3), device="cuda", generator=g)
    lhs = ops.bicubic(0.25 * x + 0.75 * y, 2160, 3840)
    rhs = 0.25 * ops.bicubic(x, 2160, 3840) + 0.75 * ops.bicubic(y, 2160, 3840)
    assert (lhs - rhs).abs().max() <= 1e-5
    ones = torch.full((1, 1080, 1920, 3), 0.5, device="cuda")
    assert (ops.bicubic(ones, 2160, 3840) - 0.5).abs().max() <= 1e-6     # taps sum to one


def test_downscale_uses_streaming_fallback():
    """Strong down-scaling: the source footprint of a tile exceeds shared memory -> global-memory streaming kernel."""
    from srb200.classic_super_resolution_algorithms.classic_algorithms import interpolate_bicubic, resize_cubic
    rng = np.random.default_rng(5)
    f = rng.random((400, 500, 3), dtype=np.float32)
    assert np.abs(interpolate_bicubic(f, (31, 20)) - ob.resize_cubic_f32(f, (31, 20), "default")).max() <= 1e-6
    u = rng.integers(0, 256, (300, 300, 3), dtype=np.uint8)
    assert np.array_equal(resize_cubic(u, (23, 17), fixed_point=True), ob.resize_cubic_u8(u, (23, 17), "scalar"))
    assert np.abs(interpolate_bicubic(f, (250, 200)) - ob.cv2_resize(f, (250, 200))).max() <= 2e-6   # mild x0.5 (smem path)


def test_bilinear_and_area_vs_cv2_golden(golden_dir):
    """interpolate_bilinear / interpolate_area (classic_algorithms.py:7-17) against cv2.resize outputs: the bilinear taps
    ride in the four-tap table of the bicubic kernels."""
    from srb200.classic_super_resolution_algorithms.classic_algorithms import interpolate_area, interpolate_bilinear
    g = np.load(os.path.join(golden_dir, "resize_cv2.npz"))
    for n in range(int(g["n_cases"])):
        h, w, dh, dw = (int(v) for v in g[f"c{n}_shape"])
        src = g[f"c{n}_in"]
        lin = interpolate_bilinear(src, (dw, dh))
        assert lin.dtype == np.float32 and lin.shape == (dh, dw, 3)
        assert np.abs(lin - g[f"c{n}_linear"]).max() <= 1e-6, n
        assert np.abs(interpolate_area(src, (dw, dh)) - g[f"c{n}_area"]).max() <= 1e-6, n
    with pytest.raises(ValueError):
        interpolate_area(np.zeros((16, 16, 3), np.float32), (8, 8))          # down-scaling area resampling is not built


def test_lanczos4_vs_cv2_golden_and_oracle(golden_dir):
    """interpolate_lanczos (classic_algorithms.py:19-21) against cv2.resize(INTER_LANCZOS4) outputs and the restatement;
    batch, grayscale, strip boundaries (more rows than one block owns) and the clip of the loaders."""
    import torch
    from srb200 import ops, _capi
    from srb200.classic_super_resolution_algorithms.classic_algorithms import interpolate_lanczos
    g = np.load(os.path.join(golden_dir, "resize_cv2.npz"))
    for n in range(int(g["n_cases"])):
        h, w, dh, dw = (int(v) for v in g[f"c{n}_shape"])
        got = interpolate_lanczos(g[f"c{n}_in"], (dw, dh))
        assert got.dtype == np.float32 and got.shape == (dh, dw, 3)
        assert np.abs(got - g[f"c{n}_lanczos4"]).max() <= 2e-6, n
    rng = np.random.default_rng(3)
    f = rng.random((3, 41, 29, 3), dtype=np.float32)
    got = interpolate_lanczos(f, (87, 300))                                  # 300 output rows: several row strips
    for i in range(3):
        assert np.abs(got[i] - ob.resize_lanczos4_f32(f[i], (87, 300))).max() <= 2e-6
    gray = rng.random((20, 33), dtype=np.float32)
    assert np.abs(interpolate_lanczos(gray, (66, 40)) - ob.resize_lanczos4_f32(gray, (66, 40))).max() <= 2e-6
    t = torch.from_numpy(f).cuda()
    clipped = ops.resize(t, 82, 58, interpolation=_capi.INTER_LANCZOS4, clip01=True).cpu().numpy()
    assert np.abs(clipped - np.clip(np.stack([ob.resize_lanczos4_f32(x, (58, 82)) for x in f]), 0, 1)).max() <= 2e-6


def test_uint8_linear_area_lanczos_bit_exact_vs_cv2_golden(golden_dir):
    """The classical benchmark (super_resolucion_clasica.ipynb cell 7) calls all four interpolators on uint8 images:
    OpenCV's fixed-point linear / area / Lanczos-4 paths, bit-exact against cv2.resize outputs and the restatement."""
    from srb200.classic_super_resolution_algorithms.classic_algorithms import (interpolate_area, interpolate_bilinear,
                                                                               interpolate_lanczos)
    g = np.load(os.path.join(golden_dir, "resize_u8_cv2.npz"))
    for n in range(int(g["n_cases"])):
        h, w, dh, dw, c = (int(v) for v in g[f"c{n}_shape"])
        src = g[f"c{n}_in"]
        for fn, key in ((interpolate_bilinear, "linear"), (interpolate_area, "area"), (interpolate_lanczos, "lanczos4")):
            got = fn(src, (dw, dh))
            assert got.dtype == np.uint8 and got.shape == g[f"c{n}_{key}"].shape
            assert np.array_equal(got, g[f"c{n}_{key}"]), (n, key)
    rng = np.random.default_rng(5)
    u = rng.integers(0, 256, (2, 90, 70, 3), dtype=np.uint8)                 # batch, several row strips
    for fn, ref in ((interpolate_bilinear, lambda x, d: ob.resize_linear_u8(x, d)),
                    (interpolate_area, lambda x, d: ob.resize_linear_u8(x, d, area=True)),
                    (interpolate_lanczos, ob.resize_lanczos4_u8)):
        got = fn(u, (211, 300))
        for i in range(2):
            assert np.array_equal(got[i], ref(u[i], (211, 300)))
    with pytest.raises(ValueError):
        interpolate_area(np.zeros((16, 16, 3), np.uint8), (8, 8))
