"""Where the e2e time of `predict()` goes at N ranks (diagnostic): per rank, (a) the device-resident forward of its shard,
(b) the read-back of the same chunk sequence alone (no kernels), (c) the host->device copies alone, (d) the full predict().

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_probe.py [--batch 512]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import numpy as np
import torch
import torch.distributed as dist

from srb200 import engine, weights


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--reps", type=int, default=4)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("nccl")
    n = a.batch // world
    net = engine.EDSRNet(weights.edsr_weights(4), 4, 16, precision="fp16")
    net.max_device_batch = 32
    x = np.random.default_rng(rank).random((n, a.tile, a.tile, 3), dtype=np.float32)
    xd = torch.from_numpy(x).cuda()
    chunks = net._chunks(n, 32)
    out = torch.empty((n, a.tile * 4, a.tile * 4, 3), dtype=torch.float32, pin_memory=True)
    ybuf = torch.empty((32, a.tile * 4, a.tile * 4, 3), dtype=torch.float32, device="cuda")
    xp = torch.from_numpy(x).pin_memory()
    xdev = torch.empty((32, a.tile, a.tile, 3), dtype=torch.float32, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def timed(fn):
        fn(); barrier()
        ts = []
        for _ in range(a.reps):
            barrier()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([min(ts)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def forward():
        for i, m in chunks:
            net.forward_device(xd[i:i + m])

    def d2h():
        for i, m in chunks:
            out[i:i + m].copy_(ybuf[:m], non_blocking=True)

    def h2d_pageable():
        xt = torch.from_numpy(x)
        for i, m in chunks:
            xdev[:m].copy_(xt[i:i + m], non_blocking=True)

    def h2d_pinned():
        for i, m in chunks:
            xdev[:m].copy_(xp[i:i + m], non_blocking=True)

    res = {"ranks": world, "tiles_per_rank": n, "chunks": [m for _, m in chunks],
           "forward_ms": timed(forward), "d2h_only_ms": timed(d2h), "h2d_pageable_ms": timed(h2d_pageable),
           "h2d_pinned_ms": timed(h2d_pinned), "predict_ms": timed(lambda: net.predict(x)),
           "predict_pinned_out_ms": timed(lambda: net.predict(x, out=out.numpy()))}
    res["d2h_gbs_per_rank"] = out.numel() * 4 / res["d2h_only_ms"] / 1e6
    if rank == 0:
        print(res)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
