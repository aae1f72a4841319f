"""Times the EDSR up-sampling tail on 32 tiles of 192 x 192 x 64: composed 5 x 5 launch against the three layered launches.

    python tools/up_probe.py [--tiles 32] [--size 192] [--scale 4] [--reps 20]
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))
from srb200 import compose, ops, weights  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return t[len(t) // 2], t[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=32)
    ap.add_argument("--size", type=int, default=192)
    ap.add_argument("--scale", type=int, default=4)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    w = weights.edsr_weights(a.scale, num_res_blocks=1, bias_scale=0.05)
    wc, bc = compose.compose_edsr_tail(w, a.scale)
    up = ops.ComposedUpsampler(wc, bc, a.scale, compose.weight_scale(wc))
    L = {n: ops.ConvWeights(w[n + "/kernel"], w[n + "/bias"]) for n in ("up0", "up1", "tail") if n + "/kernel" in w}
    x = (torch.rand((a.tiles, a.size, a.size, 64), device="cuda") - 0.5).half()

    def layered(dt=torch.float32):
        if a.scale == 4:
            h = ops.conv2d(x, L["up0"], d2s=2)
            h = ops.conv2d(h, L["up1"], d2s=2)
        else:
            h = ops.conv2d(x, L["up0"], d2s=a.scale)
        return ops.conv2d(h, L["tail"], clip01=True, out_dtype=dt)

    ya, yb = ops.upsample_composed(x, up), layered()
    print("max |composed - layered|", float((ya - yb).abs().max()))
    px = a.tiles * a.size * a.size
    for name, fn in (("composed f32", lambda: ops.upsample_composed(x, up)),
                     ("composed f16", lambda: ops.upsample_composed(x, up, out_dtype=torch.float16)),
                     ("composed u8", lambda: ops.upsample_composed(x, up, out_dtype=torch.uint8)),
                     ("layered f32", layered)):
        med, best = timed(fn, a.reps)
        fl = px * 2 * 25 * 64 * a.scale * a.scale * 3
        print(f"{name:14s} median {med:.3f} ms  best {best:.3f} ms   composed-FLOP rate {fl / med / 1e9:.0f} TFLOP/s   "
              f"{px * a.scale * a.scale / med / 1e3:.0f} output MP/s")


if __name__ == "__main__":
    main()
