"""GPU parity: the composed EDSR up-sampling tail (one 5 x 5 tcgen05 launch, nine border segments) against the float64
layer-by-layer tail (EDSR_model.py:76-95, 117-123), and the EDSR network with it against the oracle network."""
import numpy as np
import pytest
import torch

from oracle import convnets as oc

pytestmark = pytest.mark.gpu

DT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _tail(scale, channels=3, seed=11, bias_scale=0.05):
    from srb200 import compose, ops, weights
    w = weights.edsr_weights(scale, channels=channels, num_res_blocks=1, bias_scale=bias_scale, seed=seed)
    wc, bc = compose.compose_edsr_tail(w, scale)
    return w, ops.ComposedUpsampler(wc, bc, scale, compose.weight_scale(wc))


def _x(shape, kind, seed=0):
    x = np.random.default_rng(seed).uniform(-1, 1, shape).astype(np.float32)
    t = torch.from_numpy(x).cuda().to(DT[kind])
    return t, t.float().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("scale", [2, 3, 4])
@pytest.mark.parametrize("shape", [(1, 2, 2), (2, 2, 7), (1, 3, 3), (2, 5, 4), (1, 29, 31), (3, 37, 61), (1, 64, 100), (1, 192, 192)])
def test_composed_tail_vs_layered_float64(scale, shape):
    from srb200 import compose, ops
    w, up = _tail(scale)
    xd, x64 = _x(shape + (64,), "fp16", seed=shape[1] * 7 + shape[2])
    got = ops.upsample_composed(xd, up, clip01=False).cpu().numpy()
    want = compose.layered_tail(w, x64, scale)
    tol = 2e-3 * max(1.0, float(np.abs(want).max()))          # fp16 rounding of the composed weights (2^-11 relative), fp32 sums
    err = float(np.abs(got - want).max())
    assert got.shape == want.shape and err <= tol, (scale, shape, err, tol)
    # the border pixels carry their own variants: check them separately so that a wrong segment cannot hide in the max
    r = scale
    for name, sl in (("top", np.s_[:, :r]), ("bottom", np.s_[:, -r:]), ("left", np.s_[:, :, :r]), ("right", np.s_[:, :, -r:])):
        e = float(np.abs(got[sl] - want[sl]).max())
        assert e <= tol, (name, scale, shape, e)


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("out", ["float32", "float16", "uint8"])
def test_composed_tail_output_types_and_clip(kind, out):
    from srb200 import compose, ops
    w, up = _tail(4, seed=21, bias_scale=0.3)
    xd, x64 = _x((2, 33, 45, 64), kind, seed=5)
    xd = xd * 0.2
    x64 = xd.float().cpu().numpy().astype(np.float64)
    odt = {"float32": torch.float32, "float16": torch.float16, "uint8": torch.uint8}[out]
    got = ops.upsample_composed(xd, up, clip01=True, out_dtype=odt).cpu()
    want = np.clip(compose.layered_tail(w, x64, 4), 0.0, 1.0)
    wtol = 2e-3 if kind == "fp16" else 1.6e-2                       # relative rounding of the 16-bit composed weights
    if out == "uint8":
        d = np.abs(got.numpy().astype(np.float64) - want * 255.0)
        assert d.max() <= 0.5 + 255.0 * wtol * 2, d.max()
    else:
        tol = wtol * 2 + (1e-3 if out == "float16" else 0.0)
        assert float(np.abs(got.float().numpy() - want).max()) <= tol
    assert 0.05 < float((want > 0).mean()) and float((want < 1).mean()) > 0.05    # (the clip is exercised on both sides)


def test_composed_tail_single_channel_x2_and_channel_slice():
    from srb200 import compose, ops
    w, up = _tail(2, channels=1, seed=4)
    xd, _ = _x((2, 19, 23, 128), "fp16", seed=8)
    x64 = xd[..., 64:].float().cpu().numpy().astype(np.float64)
    got = ops.upsample_composed(xd, up, clip01=False, x_coffset=64).cpu().numpy()
    want = compose.layered_tail(w, x64, 2)
    assert got.shape == (2, 38, 46, 1)
    assert float(np.abs(got - want).max()) <= 2e-3 * max(1.0, float(np.abs(want).max()))


def test_composed_tail_argument_checks():
    from srb200 import ops
    _, up = _tail(2)
    with pytest.raises(ValueError):
        ops.upsample_composed(torch.zeros((1, 1, 8, 64), dtype=torch.float16, device="cuda"), up)
    with pytest.raises(ValueError):
        ops.upsample_composed(torch.zeros((1, 8, 8, 64), dtype=torch.float32, device="cuda"), up)
    assert ops.upsample_composed(torch.zeros((0, 8, 8, 64), dtype=torch.float16, device="cuda"), up).shape == (0, 16, 16, 3)


@pytest.mark.parametrize("scale", [2, 3, 4])
def test_edsr_composed_equals_layered_and_oracle(scale):
    """The network with the composed tail against the oracle network (2e-2) and against its own layer-by-layer tail."""
    from srb200 import engine, synth, weights
    hr = synth.hr_batch(2, 48 * scale, 40 * scale)
    lr = synth.area_downsample(hr, scale)
    w = weights.edsr_weights(scale, num_res_blocks=3, bias_scale=0.05)
    want = oc.edsr_forward(w, lr, scale, 3)
    comp = engine.EDSRNet(w, scale, 3, precision="fp16")
    lay = engine.EDSRNet(w, scale, 3, precision="fp16", upsampler="layered")
    assert comp.upsampler == "composed" and lay.upsampler == "layered"
    a, b = comp.predict(lr), lay.predict(lr)
    ea, eb = float(np.abs(a - want).max()), float(np.abs(b - want).max())
    assert ea <= 2e-2 and eb <= 2e-2, (ea, eb)
    assert ea <= eb * 1.5 + 1e-3, (ea, eb)                     # composing never costs accuracy against the fp32 oracle
    u8 = comp.predict(lr, out_dtype=np.uint8)
    assert np.abs(u8.astype(np.float32) - np.clip(want, 0, 1) * 255).max() <= 0.5 + 255 * 2e-2


def test_fp32_mode_keeps_the_layered_tail():
    from srb200 import engine, weights
    w = weights.edsr_weights(2, num_res_blocks=1)
    assert engine.EDSRNet(w, 2, 1, precision="fp32").upsampler == "layered"
    with pytest.raises(ValueError):
        engine.EDSRNet(w, 2, 1, precision="fp32", upsampler="composed")
