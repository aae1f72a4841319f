"""Sharded evaluation check: every rank evaluates its contiguous shard of a synthetic test set with the same
seeded EDSR, the four metric sums are all-reduced once (NCCL), and rank 0 compares the means with a
single-process evaluation of the whole set AND with the CPU oracle's evaluation of it (oracle network + oracle
tf.image.psnr/ssim restatement, fp64 means).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/eval_sharded.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import numpy as np
import torch

from srb200 import distributed as D
from srb200 import engine, synth, weights
from srb200.deep_learning_models import _common as common


def main():
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    n = 37                                              # deliberately not divisible by the world size
    hr = synth.hr_batch(n, 48, 48)
    lr = synth.area_downsample(hr, 2)
    w = weights.edsr_weights(2, num_res_blocks=2, bias_scale=0.05)
    net = engine.EDSRNet(w, 2, 2, precision="fp16")
    lo, hi = D.shard_bounds(n, rank, world)
    sums = common.evaluate_arrays(net, lr[lo:hi], hr[lo:hi], micro_batch=8)
    means = common.finish_evaluation(sums)
    if rank == 0:
        full = common.evaluate_arrays(net, lr, hr, micro_batch=16).cpu().numpy()
        ref = D.means_from_sums(full)
        err = [abs(a - b) for a, b in zip(means, ref)]
        ok = err[0] <= 1e-7 and err[1] <= 1e-4 and err[2] <= 1e-6 and full[2] == n
        from oracle import convnets as oc, metrics as om
        want = om.evaluate_means(hr, oc.edsr_forward(w, lr, 2, 2))            # [mse, psnr, ssim] of the oracle chain
        oerr = [abs(a - b) for a, b in zip(means, want)]
        ok = ok and oerr[1] <= 0.05 and oerr[2] <= 1e-3                       # 16-bit operands on random-init weights
        print("SHARDED_EVAL", "OK" if ok else "MISMATCH", "world", world, "means", means, "ref", ref, "oracle", want,
              "abs diff vs oracle", oerr)
        if not ok:
            sys.exit(1)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
