"""Does a smaller micro-batch keep the EDSR body resident in L2?  Captures one forward per micro-batch size in a
CUDA graph (so host launch cost is out of the picture) and reports device time per LR tile (diagnostic only).

    python tools/graph_probe.py [--batches 2,4,8,16,32] [--trunks fp32,half,pair] [--tile 192] [--reps 20]
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
    ap.add_argument("--batches", default="2,4,8,16,32")
    ap.add_argument("--trunks", default="fp32,half,pair")
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--blocks", type=int, default=16)
    a = ap.parse_args()
    w = weights.edsr_weights(4, num_res_blocks=a.blocks)
    print(f"{'trunk':>6} {'mb':>4} {'ms/forward':>11} {'us/tile':>9} {'MP/s':>9}")
    for trunk in a.trunks.split(","):
        net = engine.EDSRNet(w, 4, a.blocks, precision="fp16", trunk=trunk)
        for mb in [int(v) for v in a.batches.split(",")]:
            x = torch.rand((mb, a.tile, a.tile, 3), device="cuda")
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(2):
                    y = net.forward_device(x)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    y = net.forward_device(x)
                for _ in range(3):
                    g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(a.reps):
                    g.replay()
                e1.record(s)
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.reps
            print(f"{trunk:>6} {mb:4d} {ms:11.3f} {ms * 1e3 / mb:9.1f} {mb * (a.tile * 4) ** 2 / ms / 1e3:9.1f}", flush=True)
            del g, y, x


if __name__ == "__main__":
    main()
