"""Architecture known-answers and self-consistency of the Keras-semantics conv-net oracle."""
import os

import numpy as np
import torch

from oracle import convnets as oc
from srb200 import weights


def test_parameter_counts_match_reference_notebooks():
    assert weights.count_params(weights.srcnn_weights()) == 28931            # SRCNN.ipynb:L141
    assert weights.count_params(weights.edsr_weights(2)) == 1369859          # EDSR.ipynb:L392
    assert weights.count_params(weights.edsr_weights(4)) == 1517571
    assert weights.count_params(weights.esrgan_generator_weights(2, 8, 4)) == 1162915   # ESRGAN.ipynb:L636
    assert weights.count_params(weights.vgg16_classifier_weights(2)) == 14846530        # VGG16.ipynb:L151


def test_conv_matches_numpy_im2col():
    rng = np.random.default_rng(0)
    x = rng.random((2, 9, 7, 5)).astype(np.float32)
    for k in (1, 3, 5, 9):
        w = rng.standard_normal((k, k, 5, 4)).astype(np.float32)
        b = rng.standard_normal(4).astype(np.float32)
        xt = torch.from_numpy(x).permute(0, 3, 1, 2).double()
        got = oc.conv2d_same(xt, w, b, torch.float64).permute(0, 2, 3, 1).numpy()
        assert np.abs(got - oc.conv2d_same_numpy(x, w, b)).max() < 1e-10


def test_depth_to_space_is_dcr():
    x = np.arange(2 * 3 * 4 * 18, dtype=np.float32).reshape(2, 3, 4, 18)
    for r, c in ((3, 2), (2, 1)):
        xin = x[..., :c * r * r]
        ref = oc.depth_to_space_numpy(xin, r)
        got = oc.depth_to_space(torch.from_numpy(xin).permute(0, 3, 1, 2), r).permute(0, 2, 3, 1).numpy()
        assert np.array_equal(ref, got)
        b, h, w, i, j, ch = 1, 2, 3, r - 1, 0, c - 1
        assert ref[b, h * r + i, w * r + j, ch] == xin[b, h, w, (i * r + j) * c + ch]


def test_golden_regression(golden_dir):
    g = np.load(os.path.join(golden_dir, "convnets_oracle.npz"))
    lr = g["lr"]
    w = weights.srcnn_weights(bias_scale=0.1)
    assert np.abs(oc.srcnn_forward(w, lr) - g["srcnn"]).max() < 1e-5
    w = weights.edsr_weights(scale_factor=4, num_res_blocks=2, bias_scale=0.1)
    out = oc.edsr_forward(w, lr, 4, 2)
    assert out.shape == (2, 80, 96, 3) and np.abs(out - g["edsr_x4_2blocks"]).max() < 1e-4
    w = weights.edsr_weights(scale_factor=3, num_res_blocks=1, bias_scale=0.1)
    assert np.abs(oc.edsr_forward(w, lr, 3, 1) - g["edsr_x3_1block"]).max() < 1e-4
    w = weights.espcn_weights(bias_scale=0.1)
    assert np.abs(oc.espcn_forward(w, lr, 4) - g["espcn_x4"]).max() < 1e-5
    w = weights.srresnet_weights(num_res_blocks=2, bias_scale=0.1)
    assert np.abs(oc.srresnet_forward(w, lr, 4, 2) - g["srresnet_x4_2blocks"]).max() < 1e-4
    w = weights.esrgan_generator_weights(2, 8, 1, bias_scale=0.1)
    assert np.abs(oc.esrgan_generator_forward(w, lr * 2 - 1, 2, 1) - g["esrgan_x2_1rrdb_g8"]).max() < 1e-4


def test_vgg16_classifier_probabilities():
    w = weights.vgg16_classifier_weights(2)
    x = np.random.default_rng(0).random((2, 32, 32, 3)).astype(np.float32)
    p = oc.vgg16_classifier_forward(w, x)
    assert p.shape == (2, 2) and np.allclose(p.sum(1), 1, atol=1e-6)
