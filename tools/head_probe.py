import sys, torch, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/super-resolution-images-for-3d-printing-defect-detection_b200")
from srb200 import ops
x = torch.rand((64, 256, 256, 3), device="cuda")
for k in (5, 3):
    w = ops.ConvWeights(np.random.default_rng(0).uniform(-0.1, 0.1, (k, k, 3, 64)).astype(np.float32), np.zeros(64, np.float32))
    for _ in range(3): y = ops.conv2d(x, w, act="relu", out_dtype=torch.float16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): y = ops.conv2d(x, w, act="relu", out_dtype=torch.float16)
    e1.record(); torch.cuda.synchronize()
    print(f"{k}x{k} head, 64 x 256^2: {e0.elapsed_time(e1)/20:.3f} ms")
