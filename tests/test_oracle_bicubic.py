"""Pin the bicubic oracle against cv2 (the reference's real implementation) and the fixtures."""
import os

import numpy as np
import pytest

from oracle import bicubic as ob


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "bicubic_cv2.npz"))
    for n in range(int(g["n_cases"])):
        yield n, g


def test_scalar_mode_is_bit_exact_against_fixtures(golden_dir):
    for n, g in _cases(golden_dir):
        h, w, dh, dw = g[f"c{n}_shape"]
        o8 = ob.resize_cubic_u8(g[f"c{n}_u8_in"], (dw, dh), mode="scalar")
        assert np.array_equal(o8, g[f"c{n}_u8_scalar"]), n
        of = ob.resize_cubic_f32(g[f"c{n}_f32_in"], (dw, dh), mode="scalar")
        assert np.abs(of - g[f"c{n}_f32_scalar"]).max() <= 5e-7, n


def test_default_mode_against_fixtures(golden_dir):
    for n, g in _cases(golden_dir):
        h, w, dh, dw = g[f"c{n}_shape"]
        of = ob.resize_cubic_f32(g[f"c{n}_f32_in"], (dw, dh))
        assert np.abs(of - g[f"c{n}_f32_default"]).max() <= 1e-6, n
        o8 = ob.resize_cubic_u8(g[f"c{n}_u8_in"], (dw, dh)).astype(int)
        d = np.abs(o8 - g[f"c{n}_u8_default"].astype(int))
        assert d.max() <= 1 and (d != 0).mean() < 0.01, n


@pytest.mark.parametrize("shape", [(37, 53, 74, 106), (37, 53, 148, 212), (239, 239, 478, 478)])
def test_dyadic_ratios_bit_exact_vs_live_cv2(shape):
    """x2 / x4 (the reference's own 239 -> 478 case): every model agrees with cv2."""
    cv2 = pytest.importorskip("cv2")
    h, w, dh, dw = shape
    rng = np.random.default_rng(5)
    u = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = ob.cv2_resize(u, (dw, dh))
    for mode in ("default", "scalar"):
        got = ob.resize_cubic_u8(u, (dw, dh), mode=mode)
        assert (got != ref).mean() < 2e-5 and np.abs(got.astype(int) - ref).max() <= 1


def test_general_ratio_vs_live_cv2():
    pytest.importorskip("cv2")
    rng = np.random.default_rng(6)
    f = rng.random((41, 67, 3), dtype=np.float32)
    for dsize in [(201, 123), (100, 64), (1024 // 3, 100)]:
        assert np.abs(ob.resize_cubic_f32(f, dsize) - ob.cv2_resize(f, dsize)).max() <= 1e-5
        assert np.abs(ob.resize_cubic_f32(f, dsize, "scalar")
                      - ob.cv2_resize(f, dsize, optimized=False)).max() <= 1e-6


def test_x2_phase_table_and_overshoot():
    _, co = ob.axis_table(8, 16)
    assert np.allclose(co[2] * 2048, [-72, 536, 1800, -216]) or np.allclose(co[2] * 2048, [-216, 1800, 536, -72])
    rng = np.random.default_rng(1)
    out = ob.resize_cubic_f32(rng.random((32, 32, 3), dtype=np.float32), (64, 64))
    assert out.min() < 0.0 and out.max() > 1.0          # float result is not clipped


def test_grayscale_and_identity():
    rng = np.random.default_rng(2)
    g = rng.random((9, 11), dtype=np.float32)
    assert ob.resize_cubic_f32(g, (22, 18)).shape == (18, 22)
    assert np.array_equal(ob.resize_cubic_f32(g, (11, 9)), g)   # t == 0 -> taps (0,1,0,0)


def test_lanczos4_restatement_against_cv2_golden(golden_dir):
    """interpolate_lanczos (classic_algorithms.py:19-21): the 8-tap restatement against cv2.resize(INTER_LANCZOS4) outputs
    (up-scaling, down-scaling along one axis, identity) and its known answers (one-hot rows up to the 1e30 sentinel, symmetry, unit sum)."""
    g = np.load(os.path.join(golden_dir, "resize_cv2.npz"))
    for n in range(int(g["n_cases"])):
        h, w, dh, dw = (int(v) for v in g[f"c{n}_shape"])
        got = ob.resize_lanczos4_f32(g[f"c{n}_in"], (dw, dh))
        assert got.shape == (dh, dw, 3)
        assert np.abs(got - g[f"c{n}_lanczos4"]).max() <= 2e-6, n
    c = ob.lanczos4_coeffs(np.array([0.0, 0.5, 1e-7], np.float32))
    assert np.allclose(c[0], np.eye(8, dtype=np.float32)[3], atol=1e-29) and np.allclose(c[2], c[0], atol=1e-29)
    assert np.allclose(c[1], c[1][::-1], atol=1e-7) and abs(float(c[1].sum()) - 1.0) <= 1e-6
    idx, _ = ob.lanczos4_axis_table(10, 20)
    assert idx.min() == 0 and idx.max() == 9 and idx.shape == (20, 8)


def test_uint8_linear_area_lanczos_restatements_against_cv2_golden(golden_dir):
    """OpenCV's fixed-point uint8 paths for INTER_LINEAR, up-scaling INTER_AREA and INTER_LANCZOS4 (what
    classic_algorithms.py:7-21 return on the uint8 images of the classical benchmark), bit-exact against cv2 outputs."""
    g = np.load(os.path.join(golden_dir, "resize_u8_cv2.npz"))
    for n in range(int(g["n_cases"])):
        h, w, dh, dw, c = (int(v) for v in g[f"c{n}_shape"])
        src = g[f"c{n}_in"]
        assert np.array_equal(ob.resize_linear_u8(src, (dw, dh)), g[f"c{n}_linear"]), n
        assert np.array_equal(ob.resize_linear_u8(src, (dw, dh), area=True), g[f"c{n}_area"]), n
        assert np.array_equal(ob.resize_lanczos4_u8(src, (dw, dh)), g[f"c{n}_lanczos4"]), n
    # the y table keeps the fraction at the image ends (row index clamped instead): a x2 border row is a 1/4 : 3/4 blend
    idx, co = ob.linear_axis_table_u8(8, 16, clamp_t=False)
    assert idx[0].tolist() == [0, 0] and co[0].tolist() == [512, 1536]
    idx, co = ob.linear_axis_table_u8(8, 16, clamp_t=True)
    assert idx[0].tolist() == [0, 1] and co[0].tolist() == [2048, 0]
