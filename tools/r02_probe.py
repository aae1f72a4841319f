"""A short command line for ncu captures of the round-2 kernels (diagnostic only): the VGG16 classifier in fp16 (tcgen05 conv
with 64-channel K chunks, CTA pairs), the ESRGAN generator at its trained configuration (dense-block K chunks + the
SelfAttention core) and one EDSR res-block + up-sampling + tail on 32 tiles (DRAM bytes per launch for roofline.traffic).

    python tools/r02_probe.py [--part vgg|esrgan|edsr|all] [--reps 2]
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
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--part", default="all", choices=["all", "vgg", "esrgan", "edsr"])
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    runs = []
    if a.part in ("all", "vgg"):
        vgg = engine.VGG16ClassifierNet(weights.vgg16_classifier_weights(2), precision="fp16")
        xv = torch.rand((256, 128, 128, 3), device="cuda", generator=g)
        runs.append(lambda: vgg.forward_device(xv))
    if a.part in ("all", "esrgan"):
        esr = engine.ESRGANGeneratorNet(weights.esrgan_generator_weights(2, 8, 4), 2, 8, 4, precision="fp16")
        xe = torch.rand((256, 24, 24, 3), device="cuda", generator=g) * 2 - 1
        runs.append(lambda: esr.forward_device(xe))
    if a.part in ("all", "edsr"):
        edsr = engine.EDSRNet(weights.edsr_weights(4, num_res_blocks=1), 4, 1, precision="fp16")
        xd = torch.rand((32, 192, 192, 3), device="cuda", generator=g)
        runs.append(lambda: edsr.forward_device(xd))
    shapes = []
    for _ in range(a.reps):
        shapes = [tuple(r().shape) for r in runs]
    torch.cuda.synchronize()
    print("ok", shapes)


if __name__ == "__main__":
    main()
