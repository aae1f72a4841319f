"""Times the RGB head layers (NHWC8-row kernel vs im2col kernel: SRB_NO_HEAD8=1) on the bench geometries (diagnostic)."""
import sys, os, torch, numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))
from srb200 import ops

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (B, S) in ((64, 256), (32, 192), (32, 128)):
    x = torch.rand((B, S, S, 3), device="cuda")
    for k in (5, 3):
        w = ops.ConvWeights(np.random.default_rng(0).uniform(-0.1, 0.1, (k, k, 3, 64)).astype(np.float32), np.zeros(64, np.float32))
        a = timed(lambda: ops.conv2d(x, w, act="relu", out_dtype=torch.float16))
        b = timed(lambda: ops.conv2d(x, w, out_dtype=torch.float16, out2_dtype=torch.float8_e5m2, out2_error=True))
        print(f"{k}x{k} head, {B} x {S}^2: plain {a:.3f} ms, with e5m2 error output {b:.3f} ms")
