"""ESRGAN generator at the reference's trained configuration (4 RRDB, growth 8, x2) on the 1,521 patches of one 478 x 478
image: forward time per precision (diagnostic; the bench line's `configs.esrgan` is the recorded number).

    python tools/esrgan_probe.py [--rrdb 4] [--growth 8] [--patches 1521] [--reps 3]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))
from srb200 import engine, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rrdb", type=int, default=4)
    ap.add_argument("--growth", type=int, default=8)
    ap.add_argument("--patches", type=int, default=1521)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--micro-batch", type=int, default=512)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand((a.patches, 24, 24, 3), device="cuda", generator=g) * 2 - 1
    for prec in ("fp16", "fp32"):
        net = engine.ESRGANGeneratorNet(weights.esrgan_generator_weights(2, a.growth, a.rrdb), 2, a.growth, a.rrdb, precision=prec)
        net.predict_device(x, micro_batch=a.micro_batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            net.predict_device(x, micro_batch=a.micro_batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        print(f"{prec}: {ms:.2f} ms per {a.patches} patches = {a.patches / ms * 1e3:.0f} patches/s")


if __name__ == "__main__":
    main()
