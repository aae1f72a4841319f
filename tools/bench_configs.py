"""Throughput of the other BASELINE configs (parity-test cases, not bench lines): SRCNN (C1), ESPCN x4 (C2),
SRResNet x4 + VGG16 classifier (C4, reduced batch).  CUDA-event timing, inputs resident on the device.

    python tools/bench_configs.py [--json out.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import torch

from srb200 import engine, ops, weights


def timed(fn, reps=3):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []

    # C1: bicubic x2 pre-upsample + SRCNN 9-1-5 + PSNR/SSIM on 64 images 128x128 (fp32 engine)
    lr = torch.rand((64, 64, 64, 3), device="cuda", generator=g)
    hr = torch.rand((64, 128, 128, 3), device="cuda", generator=g)
    net = engine.SRCNNNet(weights.srcnn_weights(), precision="fp32")

    def c1():
        up = ops.bicubic(lr, 128, 128, clip01=True)
        return ops.psnr_ssim(hr, net.forward_device(up))
    ms = timed(c1)
    rows.append({"config": "C1 bicubic + SRCNN 9-1-5 + PSNR/SSIM, 64 x 128x128 (fp32)", "ms": ms, "out_MPps": 64 * 128 * 128 / ms / 1e3})
    net = engine.SRCNNNet(weights.srcnn_weights(), precision="fp16")
    ms = timed(c1)
    rows.append({"config": "C1 bicubic + SRCNN 9-1-5 + PSNR/SSIM, 64 x 128x128 (fp16 operands, tcgen05)", "ms": ms, "out_MPps": 64 * 128 * 128 / ms / 1e3})

    # C2: ESPCN x4, 256 tiles of 256x256
    x = torch.rand((256, 256, 256, 3), device="cuda", generator=g)
    net = engine.ESPCNNet(weights.espcn_weights(4), 4, precision="fp16")
    net.max_device_batch = 64
    ms = timed(lambda: net.predict_device(x))
    rows.append({"config": "C2 ESPCN x4, 256 x 256x256 -> 1024x1024 (fp16)", "ms": ms, "out_MPps": 256 * 1024 * 1024 / ms / 1e3})

    # C4 (reduced batch): SRResNet x4 on 64 tiles of 128x128, then the VGG16 classifier on the 512x512 SR outputs
    x = torch.rand((64, 128, 128, 3), device="cuda", generator=g)
    sr = engine.SRResNetNet(weights.srresnet_weights(4), 4, 16, precision="fp16")
    sr.max_device_batch = 32
    ms_sr = timed(lambda: sr.predict_device(x), reps=2)
    rows.append({"config": "C4a SRResNet x4, 64 x 128x128 -> 512x512 (fp16)", "ms": ms_sr, "out_MPps": 64 * 512 * 512 / ms_sr / 1e3})
    vgg = engine.VGG16ClassifierNet(weights.vgg16_classifier_weights(2), precision="fp16")
    y = torch.rand((64, 128, 128, 3), device="cuda", generator=g)
    ms_v = timed(lambda: vgg.predict_device(y), reps=2)
    rows.append({"config": "C4b VGG16 classifier, 64 patches 128x128 (fp16)", "ms": ms_v, "patches_per_s": 64 / ms_v * 1e3})
    for r in rows:
        print(r)
    if a.json:
        json.dump(rows, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
