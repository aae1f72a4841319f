"""BASELINE config 5: bicubic x2/x3/x4 + fused PSNR/SSIM bandwidth sweep over 1K-4K outputs vs the HBM roofline.

    python tools/bw_sweep.py [--reps 20] [--json out.json]

Each launch moves >= ~2 GB (the batch is sized for that) so that it is not launch-bound; algorithmic bytes are
SURVEY.md section 8d's: bicubic fp32 (12 + 12/s^2) B per RGB output pixel, uint8 (3 + 3/s^2); PSNR+SSIM 24 B per pixel.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import torch

from srb200 import ops


def timed(fn, reps):
    for _ in range(3):
        fn()
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
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true", help="one size, x2 only (for ncu)")
    a = ap.parse_args()
    peak = 6544.3
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p))["hbm_gbs"]
    rows = []
    g = torch.Generator(device="cuda").manual_seed(0)
    for (oh, ow) in (((2048, 2048),) if a.quick else ((1024, 1024), (2048, 2048), (2160, 3840))):
        for s in ((2,) if a.quick else (2, 3, 4)):
            ih, iw = -(-oh // s), -(-ow // s)
            for dt, per_out, per_in in (("f32", 12, 12), ("u8", 3, 3)):
                out_bytes = oh * ow * per_out
                batch = max(1, int(2e9 // (out_bytes + ih * iw * per_in)))
                if dt == "f32":
                    x = torch.rand((batch, ih, iw, 3), device="cuda", generator=g)
                else:
                    x = torch.randint(0, 256, (batch, ih, iw, 3), device="cuda", dtype=torch.uint8, generator=g)
                ms = timed(lambda: ops.bicubic(x, oh, ow), a.reps)
                alg = batch * (oh * ow * per_out + ih * iw * per_in)
                rows.append({"op": f"bicubic_{dt}_x{s}", "out": f"{ow}x{oh}", "batch": batch, "ms": ms,
                             "GBps": alg / ms / 1e6, "frac": alg / ms / 1e6 / peak,
                             "out_MPps": batch * oh * ow / ms / 1e3})
                del x
        batch = max(1, int(2e9 // (oh * ow * 24)))
        aimg = torch.rand((batch, oh, ow, 3), device="cuda", generator=g)
        bimg = (aimg + 0.05 * torch.randn(aimg.shape, device="cuda", generator=g)).clamp_(0, 1)
        ms = timed(lambda: ops.psnr_ssim(aimg, bimg), a.reps)
        alg = batch * oh * ow * 24
        rows.append({"op": "psnr_ssim_f32", "out": f"{ow}x{oh}", "batch": batch, "ms": ms, "GBps": alg / ms / 1e6,
                     "frac": alg / ms / 1e6 / peak, "out_MPps": batch * oh * ow / ms / 1e3})
        ms = timed(lambda: ops.psnr(aimg, bimg), a.reps)
        rows.append({"op": "psnr_f32", "out": f"{ow}x{oh}", "batch": batch, "ms": ms, "GBps": alg / ms / 1e6,
                     "frac": alg / ms / 1e6 / peak, "out_MPps": batch * oh * ow / ms / 1e3})
        del aimg, bimg
    print(f"{'op':18s} {'size':>10s} {'batch':>5s} {'ms':>8s} {'GB/s':>8s} {'of HBM':>7s} {'MP/s':>9s}")
    for r in rows:
        print(f"{r['op']:18s} {r['out']:>10s} {r['batch']:5d} {r['ms']:8.3f} {r['GBps']:8.1f} {r['frac']:7.3f} {r['out_MPps']:9.0f}")
    if a.json:
        json.dump({"hbm_peak_gbs": peak, "rows": rows}, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
