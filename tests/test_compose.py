"""Host logic: the composed EDSR up-sampling tail (srb200/compose.py) against the layer-by-layer float64 tail and against
the oracle network's own layers (EDSR_model.py:76-95, 117-123)."""
import numpy as np
import pytest
import torch

from oracle import convnets as oc
from srb200 import compose, weights


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_composed_tail_equals_layered_tail(scale):
    w = weights.edsr_weights(scale, num_res_blocks=1, bias_scale=0.05, seed=7)
    wc, bc = compose.compose_edsr_tail(w, scale)
    assert wc.shape == (3, 3, 5, 5, 64, scale * scale * 3) and bc.shape == (3, 3, scale * scale * 3)
    rng = np.random.default_rng(scale)
    for shape in ((2, 2), (2, 5), (3, 3), (5, 2), (7, 6)):
        x = rng.standard_normal((2,) + shape + (64,))
        want = compose.layered_tail(w, x, scale)
        got = compose.apply_composed(wc, bc, x, scale)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), (scale, shape)


def test_layered_tail_is_the_oracle_tail():
    """compose.layered_tail restates the same layers as oracle.convnets.edsr_forward's tail (before the clip)."""
    w = weights.edsr_weights(4, num_res_blocks=1, bias_scale=0.05, seed=3)
    x = np.random.default_rng(0).standard_normal((1, 6, 5, 64))
    t = oc._in(x, torch.float64)
    t = oc.depth_to_space(oc.conv2d_same(t, w["up0/kernel"], w["up0/bias"], torch.float64), 2)
    t = oc.depth_to_space(oc.conv2d_same(t, w["up1/kernel"], w["up1/bias"], torch.float64), 2)
    want = oc._out(oc.conv2d_same(t, w["tail/kernel"], w["tail/bias"], torch.float64))
    got = compose.layered_tail(w, x, 4)
    assert np.abs(got - want).max() <= 1e-12


def test_interior_variant_differs_from_border_variants():
    """The border classes are not a formality: the intermediate zero padding changes the map on the outermost pixels."""
    w = weights.edsr_weights(4, num_res_blocks=1, seed=5)
    wc, _ = compose.compose_edsr_tail(w, 4)
    for vy in range(3):
        for vx in range(3):
            if (vy, vx) != (1, 1):
                assert np.abs(wc[vy, vx] - wc[1, 1]).max() > 1e-3
    assert compose.weight_scale(wc) * np.abs(wc).max() >= 128 and compose.weight_scale(wc) * np.abs(wc).max() < 256


def test_single_channel_and_small_images():
    w = weights.edsr_weights(2, channels=1, num_res_blocks=1, bias_scale=0.05, seed=9)
    wc, bc = compose.compose_edsr_tail(w, 2)
    x = np.random.default_rng(1).standard_normal((1, 4, 3, 64))
    assert np.abs(compose.apply_composed(wc, bc, x, 2) - compose.layered_tail(w, x, 2)).max() <= 1e-12
    with pytest.raises(ValueError):
        compose.apply_composed(wc, bc, x[:, :1], 2)
