"""Per-layer CUDA-event timing of one EDSR x4 forward (diagnostic; not a bench number).

    python tools/layer_times.py [--batch 32] [--tile 192] [--dtype fp16] [--reps 5]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import numpy as np
import torch

from srb200 import engine, ops, weights


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--dtype", default="fp16")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--trunk", default="pair8")
    ap.add_argument("--net", default="edsr", choices=["edsr", "espcn", "srresnet", "srcnn", "vgg16"])
    a = ap.parse_args()
    if a.net == "edsr":
        net = engine.EDSRNet(weights.edsr_weights(4), 4, 16, precision=a.dtype, trunk=a.trunk)
    elif a.net == "espcn":
        net = engine.ESPCNNet(weights.espcn_weights(4), 4, precision=a.dtype)
    elif a.net == "srresnet":
        net = engine.SRResNetNet(weights.srresnet_weights(4), 4, 16, precision=a.dtype)
    elif a.net == "srcnn":
        net = engine.SRCNNNet(weights.srcnn_weights(), precision=a.dtype)
    else:
        net = engine.VGG16ClassifierNet(weights.vgg16_classifier_weights(2), precision=a.dtype)
    x = torch.rand((a.batch, a.tile, a.tile, 3), device="cuda")
    records = []
    orig = ops.conv2d

    def timed(xx, w, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(xx, w, **kw)
        e1.record()
        B, H, W, _ = xx.shape
        records.append((w, B * H * W, kw, e0, e1))
        return out
    engine.ops.conv2d = timed
    for _ in range(2):
        net.forward_device(x)
    torch.cuda.synchronize()
    acc = {}
    for rep in range(a.reps):
        records.clear()
        net.forward_device(x)
        torch.cuda.synchronize()
        for i, (w, npx, kw, e0, e1) in enumerate(records):
            acc.setdefault(i, []).append(e0.elapsed_time(e1))
    total = 0.0
    print(f"{'#':>3} {'layer':>14} {'pixels':>10} {'ms':>8} {'TFLOP/s':>9} {'GB/s(min)':>10}")
    for i, (w, npx, kw, _, _) in enumerate(records):
        ms = float(np.median(acc[i]))
        total += ms
        flop = 2.0 * npx * w.kh * w.kw * w.cin * w.cout
        in_b = npx * w.cin * (4 if w.cin == 3 else 2)
        out_b = npx * w.cout * (2 if kw.get("out_dtype") != torch.float32 else 4)
        if kw.get("out2_dtype") is not None:
            out_b += npx * w.cout * torch.empty((), dtype=kw["out2_dtype"]).element_size()
        for rk in ("res1", "res2"):
            if kw.get(rk) is not None:
                in_b += npx * w.cout * kw[rk].element_size()
        print(f"{i:3d} {w.kh}x{w.kw} {w.cin:3d}->{w.cout:3d} {npx:10d} {ms:8.3f} {flop / ms / 1e9:9.1f} {(in_b + out_b) / ms / 1e6:10.1f}")
    print(f"total {total:.3f} ms (conv layers) for {a.batch} tiles of {a.tile}x{a.tile}")


if __name__ == "__main__":
    main()
