"""One PSNR+SSIM launch shape for ncu / quick timing:  python tools/ssim_probe.py [H W B reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))
import torch
from srb200 import ops
H, W, B, reps = (int(v) for v in (sys.argv[1:5] + ["2048", "2048", "19", "10"][len(sys.argv) - 1:]))
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.rand((B, H, W, 3), device="cuda", generator=g)
b = (a + 0.05 * torch.rand(a.shape, device="cuda", generator=g)).clamp_(0, 1)
for _ in range(3):
    ops.psnr_ssim(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ops.psnr_ssim(a, b)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"psnr_ssim {B}x{H}x{W}x3: {ms:.3f} ms  {B*H*W/ms/1e3:.0f} MP/s  {B*H*W*24/ms/1e6:.0f} GB/s")
