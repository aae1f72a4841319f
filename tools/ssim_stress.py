"""Randomised shapes through the tensor-path PSNR+SSIM kernel against the exact-window CUDA-core kernels (diagnostic; a hang or a
mismatch here would be a bug in the chunk / step / mbarrier bookkeeping of csrc/metrics_mma.cu).

    python tools/ssim_stress.py [n_cases]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import numpy as np
import torch

from srb200 import _capi as capi, ops

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(7)
worst = (0.0, None)
for k in range(n_cases):
    c = int(rng.choice([1, 3]))
    h = int(rng.integers(11, 420))
    w = int(rng.integers(106 // 4 + 1, 160)) * 4            # 16-byte rows for both channel counts, map at least 96 wide
    b = int(rng.integers(1, 4))
    if k % 10 == 0:                                          # a few large ones: long chunks
        h, w, b = int(rng.integers(900, 1400)), int(rng.integers(256, 520)) * 4, int(rng.integers(1, 9))
    g = torch.Generator(device="cuda").manual_seed(k)
    a = torch.rand((b, h, w, c), device="cuda", generator=g)
    y = (a + float(rng.choice([0.0, 0.003, 0.05, 0.3])) * torch.randn(a.shape, device="cuda", generator=g)).clamp_(0, 1)
    p, s = ops.psnr_ssim(a, y)
    pe, se = ops.psnr_ssim(a, y, window=capi.SSIM_TF_EXACT)
    torch.cuda.synchronize()
    ds = (s - se).abs().max().item()
    finite = torch.isfinite(p) & torch.isfinite(pe)
    dp = (p[finite] - pe[finite]).abs().max().item() if finite.any() else 0.0
    assert ds <= 2e-5 and dp <= 1e-3 and bool((torch.isfinite(p) == torch.isfinite(pe)).all()), (k, (b, h, w, c), ds, dp)
    if ds > worst[0]:
        worst = (ds, (b, h, w, c))
print(f"{n_cases} shapes OK; worst |dSSIM| vs the exact-window kernel = {worst[0]:.2e} at {worst[1]}")
