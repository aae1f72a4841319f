"""GPU parity: fused PSNR+SSIM kernel vs the tf.image restatement (oracle/metrics.py).
Tolerances are BASELINE.json's: PSNR within 0.01 dB, SSIM within 1e-4."""
import os

import numpy as np
import pytest

from oracle import metrics as om

pytestmark = pytest.mark.gpu
PSNR_TOL, SSIM_TOL = 0.01, 1e-4


def _pair(shape, seed, noise=0.05):
    from srb200 import synth
    n, h, w, c = shape
    hr = synth.hr_batch(n, h, w, channels=c, first_index=seed)
    rng = np.random.default_rng(seed)
    pred = np.clip(hr + noise * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    return hr, pred


def test_golden_vectors(golden_dir):
    from srb200 import metrics
    g = np.load(os.path.join(golden_dir, "metrics_oracle.npz"))
    p, s = metrics.psnr(g["a"], g["b"]), metrics.ssim(g["a"], g["b"])
    assert p.dtype == np.float32 and p.shape == (4,)
    assert np.abs(p - g["psnr64"]).max() <= PSNR_TOL
    assert np.abs(s - g["ssim64"]).max() <= SSIM_TOL


@pytest.mark.parametrize("shape", [(3, 11, 11, 3), (2, 24, 24, 3), (2, 37, 150, 3), (1, 130, 47, 1), (2, 64, 300, 4),
                                   (1, 478, 478, 3)])
def test_against_oracle(shape):
    from srb200 import metrics
    a, b = _pair(shape, seed=shape[1])
    p, s = metrics.psnr_ssim(a, b)
    assert np.abs(p - om.psnr(a, b, dtype=np.float64)).max() <= PSNR_TOL
    assert np.abs(s - om.ssim(a, b, dtype=np.float64)).max() <= SSIM_TOL


# shapes the tensor-path kernel takes (1 / 3 channels, map at least 96 wide, 16-byte rows): ragged strips (22 columns) and
# steps (16 rows), a single step, the minimum height, several chunks
MMA_SHAPES = [(2, 130, 300, 3), (1, 27, 116, 3), (1, 11, 108, 3), (2, 200, 256, 1), (1, 478, 480, 3), (1, 300, 1000, 3),
              (1, 1100, 128, 3), (3, 26, 106, 1)]


@pytest.mark.parametrize("shape", MMA_SHAPES)
def test_tensor_path_against_oracle(shape):
    """mma.sync kernel (fp16 window summing to 1, hi/lo-split data) vs the float64 oracle and vs the float32-Gaussian kernel."""
    from srb200 import metrics
    a, b = _pair(shape, seed=shape[1] + shape[2])
    p, s = metrics.psnr_ssim(a, b)
    pe, se = metrics.psnr_ssim(a, b, exact=True)
    assert np.abs(p - om.psnr(a, b, dtype=np.float64)).max() <= PSNR_TOL
    assert np.abs(s - om.ssim(a, b, dtype=np.float64)).max() <= SSIM_TOL
    assert np.abs(p - pe).max() <= 1e-4            # same squared error, other summation order
    assert np.abs(s - se).max() <= 2e-5            # window rounding + 22-bit data (measured <= 3e-6)


@pytest.mark.parametrize("noise", [0.0, 0.002, 0.3])
def test_tensor_path_flat_and_identical_images(noise):
    """Variance cancellation: flat images with tiny / no noise, where E[a^2] - mu^2 is ~1e-6 against c2 = 9e-4."""
    from srb200 import metrics
    rng = np.random.default_rng(3)
    a = np.full((2, 96, 160, 3), 0.5, np.float32)
    a[1] = 0.9
    b = np.clip(a + noise * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    s = metrics.ssim(a, b)
    assert np.abs(s - om.ssim(a, b, dtype=np.float64)).max() <= SSIM_TOL
    if noise == 0.0:
        assert np.allclose(s, 1.0, atol=1e-6)


@pytest.mark.parametrize("shape", [(8, 1024, 1024, 3), (2, 1024, 1024, 3), (3, 700, 2048, 1)])
def test_tensor_path_long_chunks_vs_exact_window_kernel(shape):
    """BASELINE config 5 sizes: chunks of 16 / 8 / 4 steps per block (the small parity shapes all run 2-step chunks)."""
    import torch
    from srb200 import _capi as capi, ops
    g = torch.Generator(device="cuda").manual_seed(shape[0])
    a = torch.rand(shape, device="cuda", generator=g)
    b = (a + 0.04 * torch.randn(shape, device="cuda", generator=g)).clamp_(0, 1)
    p, s = ops.psnr_ssim(a, b)
    pe, se = ops.psnr_ssim(a, b, window=capi.SSIM_TF_EXACT)
    assert torch.allclose(p, pe, atol=1e-4) and (s - se).abs().max().item() <= 2e-5
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    assert torch.allclose(p, -10 * torch.log10(mse), atol=PSNR_TOL)


def test_tensor_path_random_shapes_vs_exact_window_kernel():
    """Forty random (batch, height, width, channels): ragged strips / steps / chunks must neither hang nor drift."""
    import torch
    from srb200 import _capi as capi, ops
    rng = np.random.default_rng(11)
    for k in range(40):
        c = int(rng.choice([1, 3]))
        shape = (int(rng.integers(1, 4)), int(rng.integers(11, 300)), int(rng.integers(27, 120)) * 4, c)
        g = torch.Generator(device="cuda").manual_seed(k)
        a = torch.rand(shape, device="cuda", generator=g)
        b = (a + float(rng.choice([0.003, 0.05, 0.3])) * torch.randn(shape, device="cuda", generator=g)).clamp_(0, 1)
        p, s = ops.psnr_ssim(a, b)
        pe, se = ops.psnr_ssim(a, b, window=capi.SSIM_TF_EXACT)
        assert (s - se).abs().max().item() <= 2e-5 and (p - pe).abs().max().item() <= 1e-3, shape


def test_large_max_val_takes_the_float32_kernels():
    """Images on a [0, 255] scale: a^2 + b^2 would overflow the fp16 operands of the tensor path - it must not be taken."""
    import torch
    from srb200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.rand((1, 64, 160, 3), device="cuda", generator=g) * 255.0
    b = (a + 8.0 * torch.randn(a.shape, device="cuda", generator=g)).clamp_(0, 255)
    p, s = ops.psnr_ssim(a, b, max_val=255.0)
    p1, s1 = ops.psnr_ssim(a / 255.0, b / 255.0, max_val=1.0)
    assert torch.isfinite(s).all() and torch.allclose(s, s1, atol=1e-4) and torch.allclose(p, p1, atol=1e-2)


def test_tensor_path_unaligned_rows_fall_back():
    """Rows that are not 16-byte multiples (W C % 4 != 0) or unaligned views take the CUDA-core kernel: same answers."""
    import torch
    from srb200 import metrics, ops
    a, b = _pair((1, 64, 134, 3), seed=2)          # 402 floats per row
    s = metrics.ssim(a, b)
    assert np.abs(s - om.ssim(a, b, dtype=np.float64)).max() <= SSIM_TOL
    ta = torch.from_numpy(_pair((2, 64, 136, 3), seed=4)[0]).cuda()
    tb = (ta * 0.9).contiguous()
    p0, s0 = ops.psnr_ssim(ta[1:], tb[1:])         # second image: base pointer offset by 64 * 136 * 3 * 4 bytes (aligned)
    p1, s1 = ops.psnr_ssim(ta, tb)
    assert torch.allclose(p0, p1[1:], atol=1e-4) and torch.allclose(s0, s1[1:], atol=1e-6)


def test_known_answers():
    from srb200 import metrics
    a = np.full((2, 16, 20, 3), 0.3, np.float32)
    b = np.full((2, 16, 20, 3), 0.6, np.float32)
    assert np.allclose(metrics.ssim(a, b), 0.80004443, atol=2e-4)
    assert np.allclose(metrics.psnr(a, b), 10.4575749, atol=1e-3)
    assert np.allclose(metrics.ssim(a, a), 1.0, atol=1e-6)
    assert np.all(np.isinf(metrics.psnr(a, a)))


def test_small_images_raise_and_single_image():
    from srb200 import metrics
    with pytest.raises(ValueError):
        metrics.ssim(np.zeros((1, 10, 32, 3), np.float32), np.zeros((1, 10, 32, 3), np.float32))
    a, b = _pair((1, 32, 32, 3), 5)
    assert np.ndim(metrics.psnr(a[0], b[0])) == 0


def test_device_tensors_and_sums():
    import torch
    from srb200 import ops
    a, b = _pair((5, 40, 40, 3), 9)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    p, s, m = ops.psnr_ssim(ta, tb, sums=sums, want_mse=True)
    ops.psnr_ssim(ta, tb, sums=sums)            # accumulates
    sums = sums.cpu().numpy()
    assert sums[2] == 10
    assert np.isclose(sums[0], 2 * p.double().sum().item()) and np.isclose(sums[1], 2 * s.double().sum().item())
    assert np.allclose(m.cpu().numpy(), ((a - b) ** 2).mean(axis=(1, 2, 3)), rtol=1e-5)


def test_full_size_properties():
    """BASELINE config 5 size (4K frame): symmetry and noise monotonicity, no oracle needed."""
    import torch
    from srb200 import ops
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.rand((2, 2160, 3840, 3), device="cuda", generator=g)
    n1 = (a + 0.02 * torch.randn(a.shape, device="cuda", generator=g)).clamp_(0, 1)
    n2 = (a + 0.08 * torch.randn(a.shape, device="cuda", generator=g)).clamp_(0, 1)
    p1, s1 = ops.psnr_ssim(a, n1)
    p1r, s1r = ops.psnr_ssim(n1, a)
    p2, s2 = ops.psnr_ssim(a, n2)
    assert torch.allclose(p1, p1r, atol=1e-4) and torch.allclose(s1, s1r, atol=1e-6)
    assert (p1 > p2).all() and (s1 > s2).all()
    mse = ((a - n1) ** 2).mean(dim=(1, 2, 3))
    assert torch.allclose(p1, -10 * torch.log10(mse), atol=PSNR_TOL)


@pytest.mark.parametrize("shape", [(7, 7, 3), (24, 24, 3), (61, 150, 3), (130, 47), (239, 478, 3)])
def test_skimage_definitions_against_oracle(shape):
    """peak_signal_noise_ratio / structural_similarity as super_resolucion_clasica.ipynb cell 7 and EDA.ipynb call them
    (float [0, 1] with data_range=1.0 and channel_axis=2; grayscale; uint8 with data_range=255)."""
    from srb200 import metrics, synth
    gray = len(shape) == 2
    hr = synth.hr_batch(1, shape[0], shape[1], channels=1 if gray else 3, first_index=shape[1])[0]
    rng = np.random.default_rng(shape[0])
    sr = np.clip(hr + 0.04 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    if gray:
        hr, sr = hr[:, :, 0], sr[:, :, 0]
    ax = None if gray else 2
    s = metrics.structural_similarity(hr, sr, channel_axis=ax, data_range=1.0)
    p = metrics.peak_signal_noise_ratio(hr, sr, data_range=1.0)
    assert isinstance(s, float) and isinstance(p, float)
    assert abs(s - om.skimage_ssim(hr, sr, 1.0, channel_axis=ax)) <= SSIM_TOL
    assert abs(p - om.skimage_psnr(hr, sr, 1.0)) <= PSNR_TOL
    hu, su = np.rint(hr * 255).astype(np.uint8), np.rint(sr * 255).astype(np.uint8)
    assert abs(metrics.structural_similarity(hu, su, channel_axis=ax, data_range=255) - om.skimage_ssim(hu, su, 255, channel_axis=ax)) <= SSIM_TOL
    assert abs(metrics.peak_signal_noise_ratio(hu, su, data_range=255) - om.skimage_psnr(hu, su, 255)) <= PSNR_TOL


def test_skimage_definitions_errors_and_identity():
    from srb200 import metrics
    x = np.random.default_rng(0).random((20, 20, 3)).astype(np.float32)
    assert abs(metrics.structural_similarity(x, x, channel_axis=2, data_range=1.0) - 1.0) <= 1e-6
    assert np.isinf(metrics.peak_signal_noise_ratio(x, x, data_range=1.0))
    with pytest.raises(ValueError):
        metrics.structural_similarity(x[:6], x[:6], channel_axis=2, data_range=1.0)
    with pytest.raises(ValueError):
        metrics.structural_similarity(x, x, channel_axis=2)                   # float images need data_range
    with pytest.raises(NotImplementedError):
        metrics.structural_similarity(x, x, channel_axis=2, data_range=1.0, gaussian_weights=True)


@pytest.mark.parametrize("shape", [(3, 5, 7, 3), (2, 33, 33, 3), (1, 130, 47, 1), (2, 64, 300, 4), (1, 478, 478, 3)])
def test_psnr_alone_streaming_reduction(shape):
    """metrics.psnr (tf.image.psnr, metrics.py:3-4) runs the squared-error reduction alone: any size (no 11 x 11 minimum),
    vector and scalar paths, same values as the fused pass."""
    import torch
    from srb200 import metrics, ops
    a, b = _pair(shape, seed=shape[2])
    p = metrics.psnr(a, b)
    assert p.dtype == np.float32 and p.shape == (shape[0],)
    assert np.abs(p - om.psnr(a, b, dtype=np.float64)).max() <= PSNR_TOL
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    p2, m2 = ops.psnr(ta, tb, want_mse=True)
    assert np.allclose(m2.cpu().numpy(), ((a - b) ** 2).mean(axis=(1, 2, 3)), rtol=1e-5)
    if shape[1] >= 11 and shape[2] >= 11:
        assert np.abs(p2.cpu().numpy() - ops.psnr_ssim(ta, tb)[0].cpu().numpy()).max() <= 1e-4
    # unaligned views take the scalar path
    n = (ta.numel() - 1) // 3 * 3
    va, vb = ta.flatten()[1:1 + n].view(1, 1, n // 3, 3), tb.flatten()[1:1 + n].view(1, 1, n // 3, 3)
    want = -10 * np.log10(((a.ravel()[1:1 + n].astype(np.float64) - b.ravel()[1:1 + n]) ** 2).mean())
    assert abs(float(ops.psnr(va, vb)[0]) - want) <= PSNR_TOL
    assert np.all(np.isinf(metrics.psnr(a, a)))
