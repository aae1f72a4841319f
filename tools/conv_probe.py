"""One EDSR res-block (conv1 + conv2) plus the up-sampling stage on 32 tiles, a few repetitions: a short command line
for ncu captures of the tcgen05 conv kernel (diagnostic only).

    python tools/conv_probe.py [--trunk pair8|fp32|pair] [--reps 3] [--batch 32] [--tile 192] [--late]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import torch

from srb200 import engine, weights


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trunk", default="pair8")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--blocks", type=int, default=1)
    a = ap.parse_args()
    w = weights.edsr_weights(4, num_res_blocks=a.blocks)
    net = engine.EDSRNet(w, 4, a.blocks, precision="fp16", trunk=a.trunk)
    x = torch.rand((a.batch, a.tile, a.tile, 3), device="cuda")
    for _ in range(a.reps):
        y = net.forward_device(x)
    torch.cuda.synchronize()
    print("ok", tuple(y.shape))


if __name__ == "__main__":
    main()
