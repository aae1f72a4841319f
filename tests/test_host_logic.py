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
    assert [m for _, m in DeviceModel._chunks(512, 32)][:4] == [8, 8, 16, 32]      # ramped start: the read-back begins early
    assert DeviceModel._chunks(32, 32) == [(0, 32)]
    assert [m for _, m in DeviceModel._chunks(64, 32)] == [8, 8, 16, 16, 8, 8]
    assert [m for _, m in DeviceModel._chunks(48, 32)] == [32, 8, 4, 4]            # (too short a job for a ramp)
    assert [m for _, m in DeviceModel._chunks(128, 32)][:3] == [8, 8, 16]


def test_adopt_maps_keras_auto_names_in_creation_order_and_validates_shapes():
    """weights.adopt: files exported from the reference's Keras models ({v.name: v.numpy() for v in model.weights}) carry
    conv2d, conv2d_7, ... (SRCNN_model.py:50-52, EDSR_model.py:61-121 leave the layers unnamed): matched by creation order;
    explicit names (ESRGAN, VGG16) are the internal ones, with the SelfAttention sub-layer path stripped; wrong depth, wrong
    shape and missing layers fail at load time."""
    import numpy as np
    from srb200 import weights as W
    w = W.edsr_weights(4, num_res_blocks=2, bias_scale=0.05)
    spec = W.edsr_weights(4, num_res_blocks=2, shapes_only=True)
    assert list(spec) == list(w) and all(spec[k].shape == w[k].shape for k in w)
    layers = [k[:-len("/kernel")] for k in w if k.endswith("/kernel")]
    raw = {}
    for i, name in enumerate(layers):
        kname = f"conv2d_{i + 40}"                         # a session that built other models before
        raw[f"{kname}/kernel:0"], raw[f"{kname}/bias:0"] = w[name + "/kernel"], w[name + "/bias"]
    got = W.adopt(raw, spec, "EDSR")
    assert list(got) == list(w) and all(np.array_equal(got[k], w[k]) for k in w)
    assert all(np.array_equal(W.adopt(w, spec)[k], w[k]) for k in w)          # internal names pass through
    with pytest.raises(ValueError, match="auto-named Keras Conv2D"):
        W.adopt({k: v for k, v in raw.items() if not k.startswith("conv2d_41/")}, spec, "EDSR")
    bad = dict(w)
    bad["rb1_c2/kernel"] = np.zeros((3, 3, 64, 32), np.float32)
    with pytest.raises(ValueError, match="rb1_c2/kernel.*expected \\(3, 3, 64, 64\\)"):
        W.adopt(bad, spec, "EDSR")
    g = W.esrgan_generator_weights(2, 8, 1)
    named = {}
    for k, v in g.items():
        layer, var = k.split("/")
        prefix = layer.rsplit("_", 1)[0] + "/" if layer.startswith("self_attention") else ""
        named[f"{prefix}{layer}/{var}:0"] = v
    got = W.adopt(named, W.esrgan_generator_weights(2, 8, 1, shapes_only=True), "ESRGAN")
    assert all(np.array_equal(got[k], g[k]) for k in g)
    no_bias = {k: v for k, v in w.items() if k != "tail/bias"}
    assert np.array_equal(W.adopt(no_bias, spec)["tail/bias"], np.zeros(3, np.float32))   # use_bias=False layers
    assert W.count_params(W.vgg16_classifier_weights(2, shapes_only=True)) == 14_846_530
    assert W.count_params(W.esrgan_generator_weights(2, 8, 4, shapes_only=True)) == 1_162_915


def test_classical_upscalers_outside_the_path_fail_at_the_call_with_a_reason():
    """classic_algorithms.py:23-110: the four non-cv2.resize upscalers keep their names (imports work) and raise
    NotImplementedError when called."""
    from srb200.classic_super_resolution_algorithms import classic_algorithms as ca
    for name in ("back_projection", "non_local_means", "edge_guided_interpolation", "frequency_extrapolation"):
        with pytest.raises(NotImplementedError, match="outside the B200 hot path"):
            getattr(ca, name)(None, None)
    assert {"interpolate_bilinear", "interpolate_bicubic", "interpolate_area", "interpolate_lanczos"} <= set(dir(ca))
