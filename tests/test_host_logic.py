"""CPU: host-side logic - voting, sharding, metric aggregation, world_size-2 gloo all-reduce."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, PKG


def test_vote_majority_and_tiebreak():
    from srb200.defect_detection_models.VGG16_model import vote
    probs = np.array([[0.9, 0.1], [0.4, 0.6], [0.3, 0.7]])
    assert vote(probs) == (1, pytest.approx((0.1 + 0.6 + 0.7) / 3))
    tie = np.array([[0.9, 0.1], [0.4, 0.6]])          # 1 vote each -> higher mean prob wins (class 0)
    assert vote(tie)[0] == 0


def test_shard_bounds_cover_everything():
    from srb200.distributed import shard_bounds
    for n in (0, 1, 7, 512, 513):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_means_from_sums():
    from srb200.distributed import means_from_sums
    assert means_from_sums([60.0, 1.8, 2.0, 0.004]) == [0.002, 30.0, 0.9]


def test_error_conventions_of_the_wrappers():
    from srb200.deep_learning_models.SRCNN_model import SRCNNModel
    from srb200.deep_learning_models.EDSR_model import EDSR
    m = SRCNNModel()
    with pytest.raises(RuntimeError, match="Model has not been trained"):
        m.evaluate(None, None)
    with pytest.raises(RuntimeError, match="Model has not been trained"):
        m.super_resolve_image(np.zeros((8, 8, 3), np.float32), 16, 16)
    with pytest.raises(FileNotFoundError):
        m.setup_model(from_pretrained=True, pretrained_path="/nonexistent.npz")
    e = EDSR()
    with pytest.raises(RuntimeError, match="Model has not been trained"):
        e.super_resolve_image(np.zeros((8, 8, 3), np.float32))
    with pytest.raises(FileNotFoundError):
        e.setup_model(from_pretrained=True, pretrained_path=None)


_WORKER = r"""
import os, sys
sys.path.insert(0, {pkg!r})
import torch, torch.distributed as dist
from srb200 import distributed as D
rank, world, _ = D.init_from_env(backend="gloo")
lo, hi = D.shard_bounds(10)
# each rank owns images [lo, hi) with psnr = 20 + index, ssim = index / 10, mse = 0.001 * index
idx = torch.arange(lo, hi, dtype=torch.float64)
sums = torch.stack([(20 + idx).sum(), (idx / 10).sum(), torch.tensor(float(hi - lo), dtype=torch.float64), (0.001 * idx).sum()])
total = D.allreduce_sums(sums)
vec = D.allgather_vector(torch.full((5,), float(rank)))
if rank == 0:
    print("RESULT", (D.means_from_sums(total.tolist()), vec.tolist()))
dist.destroy_process_group()
"""


def test_gloo_world2_allreduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(pkg=PKG))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    line = [l for l in outs[0][0].splitlines() if l.startswith("RESULT")][0]
    means, vec = eval(line[len("RESULT"):])
    assert means == pytest.approx([0.0045, 24.5, 0.45])
    assert vec == [0.0] * 5 + [1.0] * 5


def test_constants_keep_the_reference_names():
    """SRModels/constants.py:1-14: same names and values, exported from the per-family table."""
    from srb200 import constants as c
    assert (c.SRCNN_PATCH_SIZE, c.SRCNN_STRIDE) == (24, 12)
    assert (c.EDSR_PATCH_SIZE, c.EDSR_STRIDE, c.EDSR_SCALE_FACTOR) == (24, 12, 2)
    assert (c.ESRGAN_PATCH_SIZE, c.ESRGAN_STRIDE, c.ESRGAN_SCALE_FACTOR) == (24, 12, 2)
    assert (c.VGG_PATCH_SIZE, c.VGG_STRIDE, c.RANDOM_SEED) == (96, 48, 42)
    assert not hasattr(c, "SRCNN_SCALE_FACTOR") and not hasattr(c, "VGG_SCALE_FACTOR")


def test_predict_chunk_schedule():
    """engine.DeviceModel._chunks: contiguous cover of the batch, last micro-batch cut into 1/2 + 1/4 + 1/4."""
    from srb200.engine import DeviceModel
    for n, mb in [(512, 32), (64, 16), (70, 32), (5, 32), (32, 32), (33, 32), (16, 64)]:
        ch = DeviceModel._chunks(n, mb)
        assert ch[0][0] == 0 and sum(m for _, m in ch) == n and all(m > 0 for _, m in ch)
        assert all(ch[k][0] + ch[k][1] == ch[k + 1][0] for k in range(len(ch) - 1))
        assert max(m for _, m in ch) <= mb
    assert [m for _, m in DeviceModel._chunks(512, 32)][-3:] == [16, 8, 8]
    assert DeviceModel._chunks(32, 32) == [(0, 32)]
