"""Wall-clock latency of the reference-facing calls on one image (diagnostic): super_resolve_image (patch flow),
super_resolve_image_whole (fully convolutional), evaluate on a small test set.

    python tools/api_latency.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import numpy as np
import torch

from srb200 import synth
from srb200.deep_learning_models.EDSR_model import EDSR
from srb200.deep_learning_models.SRCNN_model import SRCNNModel


def wall(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), float(np.median(ts))


def main():
    hr = synth.hr_batch(1, 480, 480)[0]
    lr = synth.area_downsample(hr[None], 2)[0]                    # 240 x 240, as the reference's 478 -> 239 images
    m = EDSR(); m.setup_model(scale_factor=2); m.trained = True
    print("EDSR x2 super_resolve_image(240x240, patch 48 / stride 24): min %.2f ms, median %.2f ms" % wall(lambda: m.super_resolve_image(lr, patch_size_lr=48, stride=24)))
    if hasattr(m, "super_resolve_image_whole"):
        print("EDSR x2 super_resolve_image_whole(240x240):                   min %.2f ms, median %.2f ms" % wall(lambda: m.super_resolve_image_whole(lr)))
    X = synth.area_downsample(synth.hr_batch(64, 96, 96), 2)
    Y = synth.hr_batch(64, 96, 96)
    print("EDSR x2 evaluate(64 pairs 48 -> 96):                          min %.2f ms, median %.2f ms" % wall(lambda: m.evaluate(X, Y)))
    s = SRCNNModel(); s.setup_model(input_shape=(480, 480, 3)); s._trained = True
    print("SRCNN super_resolve_image(240x240 -> 480x480, patch 33 / 14): min %.2f ms, median %.2f ms" % wall(lambda: s.super_resolve_image(lr, 480, 480)))
    print("SRCNN super_resolve_image_whole(240x240 -> 480x480):          min %.2f ms, median %.2f ms" % wall(lambda: s.super_resolve_image_whole(lr, 480, 480)))


if __name__ == "__main__":
    main()
