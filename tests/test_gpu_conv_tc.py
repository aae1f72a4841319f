"""GPU parity: the tcgen05 / TMEM implicit-GEMM engine against the float64 im2col oracle.

Inputs and weights are pre-rounded to the 16-bit operand format, so the only difference to the
oracle is fp32 accumulation order plus the final output rounding: tolerances are tight (a few ULP of
the output format), which is what exposes any swizzle / descriptor / tap-addressing mistake."""
import numpy as np
import pytest
import torch

from oracle import convnets as oc

pytestmark = pytest.mark.gpu

DT = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _round(a, kind):
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return (t.bfloat16() if kind == "bf16" else t.half()).float().numpy()


def _rand(shape, seed, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(np.float32)


def _run(kind, B, H, W, cout, out_dtype=torch.float32, ksize=3, cin=64, x_width=None, **kw):
    """``cin``: input channels of the layer; ``x_width``: channels of the buffer the layer reads a prefix of (>= cin)."""
    from srb200 import ops, _capi
    xw = x_width or cin
    x_full = _round(_rand((B, H, W, xw), 1), kind)
    x = np.ascontiguousarray(x_full[..., :cin])
    kern = _round(_rand((ksize, ksize, cin, cout), 2, -0.1, 0.1), kind)
    bias = _rand((cout,), 3, -0.1, 0.1)
    r = kw.get("d2s", 1)
    c_post = cout // (r * r)
    res1 = _rand((B, H * r, W * r, c_post), 5) if kw.pop("res1", False) else None
    want = oc.conv2d_same_numpy(x, kern, bias)
    if kw.get("act") == "relu":
        want = np.maximum(want, 0)
    if r > 1:
        want = oc.depth_to_space_numpy(want, r)
    want = want * kw.get("alpha", 1.0)
    if res1 is not None:
        want = want + res1
    if kw.get("clip01"):
        want = np.clip(want, 0, 1)
    w = ops.ConvWeights(kern, bias)
    xd = torch.from_numpy(x_full).cuda().to(DT[kind])
    assert ops.conv2d_engine(xd, w, r) == _capi.ENGINE_TCGEN05
    got = ops.conv2d(xd, w, act=kw.get("act"), alpha=kw.get("alpha", 1.0), clip01=kw.get("clip01", False),
                     res1=None if res1 is None else torch.from_numpy(res1).cuda(), d2s=r,
                     out_dtype=out_dtype, engine=_capi.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    return got.float().cpu().numpy(), want


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 16, 8, 64), (1, 48, 48, 64), (3, 21, 13, 64), (1, 5, 3, 64), (1, 192, 192, 64),
                                   (1, 24, 24, 256), (1, 17, 9, 576), (2, 24, 24, 3), (1, 20, 12, 32), (1, 16, 16, 48),
                                   (1, 17, 9, 3), (2, 29, 13, 3), (1, 14, 8, 3), (1, 15, 8, 5), (1, 64, 64, 3)])
def test_plain_conv(kind, shape):
    B, H, W, cout = shape
    got, want = _run(kind, B, H, W, cout)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-3


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
def test_fused_epilogues(kind):
    tol16 = 2e-2 if kind == "bf16" else 4e-3                 # output rounding of values up to ~3
    got, want = _run(kind, 2, 20, 20, 64, act="relu", out_dtype=DT[kind])
    assert np.abs(got - want).max() <= tol16
    got, want = _run(kind, 1, 19, 21, 64, alpha=0.1, res1=True)
    assert np.abs(got - want).max() <= 2e-3
    got, want = _run(kind, 1, 12, 10, 256, d2s=2, out_dtype=DT[kind])
    assert got.shape == (1, 24, 20, 64) and np.abs(got - want).max() <= tol16
    got, want = _run(kind, 1, 9, 11, 576, d2s=3)
    assert got.shape == (1, 27, 33, 64) and np.abs(got - want).max() <= 2e-3
    got, want = _run(kind, 1, 24, 24, 3, clip01=True)
    assert np.abs(got - want).max() <= 2e-3


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
def test_tma_epilogue_paths(kind):
    """Layer shapes that take the TMA epilogue (16-bit output, 64/128-channel chunks): plain and ReLU stores with tiles
    clipped at the right / bottom image edge, several images, and depth_to_space(2) through the 5-D output map
    (needs whole 16-row tiles)."""
    tol16 = 2e-2 if kind == "bf16" else 4e-3
    for (B, H, W) in ((2, 32, 24), (3, 19, 21), (1, 40, 13)):
        got, want = _run(kind, B, H, W, 64, act="relu", out_dtype=DT[kind])
        assert np.abs(got - want).max() <= tol16, (B, H, W)
        got, want = _run(kind, B, H, W, 128, out_dtype=DT[kind])
        assert np.abs(got - want).max() <= tol16, (B, H, W)
    for (B, H, W) in ((2, 32, 24), (1, 16, 21), (3, 48, 8)):
        got, want = _run(kind, B, H, W, 256, d2s=2, out_dtype=DT[kind])
        assert got.shape == (B, 2 * H, 2 * W, 64) and np.abs(got - want).max() <= tol16, (B, H, W)
    # 256-channel chunks run as CTA pairs (cta_group::2) with two 64-channel epilogue sub-blocks per warp: odd tile
    # counts (dummy tile of the pair), clipped tiles, with and without depth_to_space
    for (B, H, W) in ((1, 19, 21), (3, 16, 8), (2, 40, 13)):
        got, want = _run(kind, B, H, W, 256, act="relu", out_dtype=DT[kind])
        assert got.shape == (B, H, W, 256) and np.abs(got - want).max() <= tol16, (B, H, W)


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("ksize", [3, 5, 9])
def test_rgb_head_on_tensor_cores(kind, ksize):
    """K x K, 3 -> 64 head layers as an im2col GEMM on tcgen05 (fp32 image in, 16-bit out): against the fp64 oracle on
    operands rounded to the 16-bit type, with edge-clipped tiles, several images, the activations the networks use, and
    the pair8 second output (e5m2 rounding error of y)."""
    from srb200 import ops, _capi
    dt = DT[kind]
    tol = 3e-2 if kind == "bf16" else 4e-3
    kern = _rand((ksize, ksize, 3, 64), 2, -0.2, 0.2)
    bias = _rand((64,), 3, -0.1, 0.1)
    slopes = _rand((64,), 4, 0.05, 0.4)
    w = ops.ConvWeights(kern, bias, slopes)
    for (B, H, W, act) in ((2, 32, 24, "relu"), (3, 19, 21, None), (1, 40, 13, "prelu"), (1, 16, 8, "tanh")):
        x = _rand((B, H, W, 3), 10 + H, 0.0, 1.0)
        want = oc.conv2d_same_numpy(_round(x, kind), _round(kern, kind), bias)
        if act == "relu":
            want = np.maximum(want, 0)
        elif act == "prelu":
            want = np.where(want >= 0, want, want * slopes)
        elif act == "tanh":
            want = np.tanh(want)
        xd = torch.from_numpy(x).cuda()
        got = ops.conv2d(xd, w, act=act, out_dtype=dt, engine=_capi.ENGINE_TCGEN05)
        assert got.dtype == dt and np.abs(got.float().cpu().numpy() - want).max() <= tol, (B, H, W, act)
        ref = ops.conv2d(xd, w, act=act, out_dtype=dt, engine=_capi.ENGINE_DIRECT)      # exact fp32 CUDA-core engine
        assert (got.float() - ref.float()).abs().max().item() <= tol
    x = _rand((2, 24, 24, 3), 7, 0.0, 1.0)
    xd = torch.from_numpy(x).cuda()
    y, e = ops.conv2d(xd, w, out_dtype=dt, out2_dtype=torch.float8_e5m2, out2_error=True, engine=_capi.ENGINE_TCGEN05)
    want = oc.conv2d_same_numpy(_round(x, kind), _round(kern, kind), bias)
    e1 = np.abs(y.double().cpu().numpy() - want).max()
    e2 = np.abs(y.double().cpu().numpy() + e.double().cpu().numpy() - want).max()
    assert e.dtype == torch.float8_e5m2 and e2 <= max(e1 / 3, 2e-4), (e1, e2)


def test_second_output_copy_and_channel_slice():
    from srb200 import ops, _capi
    x = _round(_rand((1, 16, 16, 64), 1), "fp16")
    kern = _round(_rand((3, 3, 64, 64), 2, -0.1, 0.1), "fp16")
    w = ops.ConvWeights(kern, None)
    xd = torch.from_numpy(x).cuda().half()
    y32, y16 = ops.conv2d(xd, w, out_dtype=torch.float32, out2_dtype=torch.float16, engine=_capi.ENGINE_TCGEN05)
    want = oc.conv2d_same_numpy(x, kern)
    assert np.abs(y32.cpu().numpy() - want).max() <= 2e-3
    assert torch.equal(y16, y32.half())
    # input is a 64-channel slice (offset 8) of a wider buffer
    wide = torch.zeros((1, 16, 16, 80), dtype=torch.float16, device="cuda")
    wide[..., 8:72] = xd
    got = ops.conv2d(wide, w, x_coffset=8, out_dtype=torch.float32, engine=_capi.ENGINE_TCGEN05)
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-3


def test_engines_agree_at_benchmark_tile_size():
    """192 x 192 x 64 (BASELINE config 3 tile): tcgen05 vs the exact CUDA-core engine on the same
    16-bit inputs - a full-size check that needs no CPU oracle."""
    from srb200 import ops, _capi
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.rand((4, 192, 192, 64), device="cuda", generator=g) * 2 - 1).half()
    kern = _round(_rand((3, 3, 64, 64), 2, -0.1, 0.1), "fp16")
    w = ops.ConvWeights(kern, _rand((64,), 3, -0.1, 0.1))
    a = ops.conv2d(x, w, act="relu", out_dtype=torch.float32, engine=_capi.ENGINE_TCGEN05)
    b = ops.conv2d(x, w, act="relu", out_dtype=torch.float32, engine=_capi.ENGINE_DIRECT)
    assert (a - b).abs().max().item() <= 2e-3


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 32, 24), (1, 19, 21)])
def test_compensated_trunk_pair(kind, shape):
    """y2_mode = 1: residual given as a 16-bit (hi, lo) pair, output as y = round16(v) and y2 = v - y.
    hi + lo must reproduce the fp64 oracle to ~2^-20 relative, far beyond a single 16-bit tensor."""
    from srb200 import ops, _capi
    B, H, W = shape
    dt = DT[kind]
    x = _round(_rand((B, H, W, 64), 1), kind)
    kern = _round(_rand((3, 3, 64, 64), 2, -0.1, 0.1), kind)
    bias = _rand((64,), 3, -0.1, 0.1)
    trunk = _rand((B, H, W, 64), 5, -2.0, 2.0)
    hi = _round(trunk, kind)
    lo = _round(trunk - hi, kind)
    want = 0.1 * oc.conv2d_same_numpy(x, kern, bias) + (hi.astype(np.float64) + lo)
    w = ops.ConvWeights(kern, bias)
    t = lambda a: torch.from_numpy(a).cuda().to(dt)
    tol = 2e-4 if kind == "fp16" else 2e-3
    for engine in (_capi.ENGINE_TCGEN05, _capi.ENGINE_DIRECT):
        y, y2 = ops.conv2d(t(x), w, alpha=0.1, res1=t(hi), res2=t(lo), out_dtype=dt, out2_dtype=dt, out2_error=True,
                           engine=engine)
        got = y.double().cpu().numpy() + y2.double().cpu().numpy()
        assert np.abs(got - want).max() <= tol, engine
        assert np.abs(y.double().cpu().numpy() - want).max() > 2 * tol          # one 16-bit tensor alone is far coarser
        y_only = ops.conv2d(t(x), w, alpha=0.1, res1=t(hi), res2=t(lo), out_dtype=dt, engine=engine)
        assert torch.equal(y_only, y)                                           # trunk exit: same rounded sum, no y2


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("shape", [(2, 32, 24), (1, 19, 21)])
def test_compensated_trunk_pair8(kind, shape):
    """"pair8" trunk: residual = 16-bit hi + float8-e5m2 lo; output y = round16(v) and y2 = e5m2(v - y).  y + y2 must
    carry ~3 more bits than y alone (e5m2 keeps 2-3 significant bits of the rounding error), on both engines, and the
    tensor-core kernel must agree with the scalar engine."""
    from srb200 import ops, _capi
    B, H, W = shape
    dt = DT[kind]
    x = _round(_rand((B, H, W, 64), 1), kind)
    kern = _round(_rand((3, 3, 64, 64), 2, -0.1, 0.1), kind)
    bias = _rand((64,), 3, -0.1, 0.1)
    trunk = _rand((B, H, W, 64), 5, -2.0, 2.0)
    hi_t = torch.from_numpy(trunk).to(dt)
    lo_t = (torch.from_numpy(trunk) - hi_t.float()).to(torch.float8_e5m2)
    want = 0.1 * oc.conv2d_same_numpy(x, kern, bias) + (hi_t.double().numpy() + lo_t.double().numpy())
    w = ops.ConvWeights(kern, bias)
    xd = torch.from_numpy(x).cuda().to(dt)
    res = {}
    for engine in (_capi.ENGINE_TCGEN05, _capi.ENGINE_DIRECT):
        y, y2 = ops.conv2d(xd, w, alpha=0.1, res1=hi_t.cuda(), res2=lo_t.cuda(), out_dtype=dt, out2_dtype=torch.float8_e5m2,
                           out2_error=True, engine=engine)
        assert y2.dtype == torch.float8_e5m2
        e1 = np.abs(y.double().cpu().numpy() - want).max()
        e2 = np.abs(y.double().cpu().numpy() + y2.double().cpu().numpy() - want).max()
        assert e2 <= e1 / 3, (engine, e1, e2)                                   # the e5m2 term buys >= ~2 bits
        y_only = ops.conv2d(xd, w, alpha=0.1, res1=hi_t.cuda(), res2=lo_t.cuda(), out_dtype=dt, engine=engine)
        assert torch.equal(y_only, y)
        res[engine] = (y, y2)
    ya, yb = res[_capi.ENGINE_TCGEN05][0].float(), res[_capi.ENGINE_DIRECT][0].float()
    assert (ya - yb).abs().max().item() <= (4e-3 if kind == "fp16" else 3.2e-2)  # <= 1-2 ulp of the 16-bit format at |v| <= 4


@pytest.mark.parametrize("shape", [(2, 32, 24, 64), (1, 21, 13, 64), (1, 24, 24, 256), (3, 16, 8, 64)])
def test_cta_pairs_match_single_cta(shape):
    """cta_group::2 launch (cluster of two CTAs, M = 256 MMAs, weights split across the pair, odd tile counts give
    the second CTA a dummy tile) must give bit-identical results to the single-CTA kernel."""
    from srb200 import ops, _capi
    B, H, W, cout = shape
    x = torch.from_numpy(_round(_rand((B, H, W, 64), 1), "fp16")).cuda().half()
    w = ops.ConvWeights(_round(_rand((3, 3, 64, cout), 2, -0.1, 0.1), "fp16"), _rand((cout,), 3, -0.1, 0.1))
    res = torch.from_numpy(_rand((B, H, W, cout), 5)).cuda()
    outs = []
    for pairs in (0, 1):
        prev = _capi.lib().srb_conv_tc_set_cta_pairs(pairs)
        try:
            a = ops.conv2d(x, w, act="relu", out_dtype=torch.float16, engine=_capi.ENGINE_TCGEN05)
            b32, b16 = ops.conv2d(x, w, alpha=0.1, res1=res, out_dtype=torch.float32, out2_dtype=torch.float16,
                                  engine=_capi.ENGINE_TCGEN05)
            torch.cuda.synchronize()
        finally:
            _capi.lib().srb_conv_tc_set_cta_pairs(prev)
        outs.append((a, b32, b16))
    for u, v in zip(*outs):
        assert torch.equal(u, v)


@pytest.mark.parametrize("ksize,shape", [(5, (1, 20, 13, 64)), (9, (1, 33, 17, 3)), (9, (2, 16, 8, 64)), (7, (1, 9, 30, 32)),
                                         (1, (1, 18, 10, 64))])
def test_larger_filters(ksize, shape):
    """Odd K x K filters with Cin = 64 (SRResNet's 9x9 tail, 5x5 / 7x7 / 1x1): same shifted-descriptor scheme, K*K taps."""
    B, H, W, cout = shape
    big = ksize == 9 and cout == 64                        # 166 KB of weights: only the 16-bit staging fits next to them
    got, want = _run("fp16", B, H, W, cout, ksize=ksize, out_dtype=torch.float16 if big else torch.float32)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= (2e-2 if big else 2e-3 * max(1.0, ksize * ksize / 9.0))


@pytest.mark.parametrize("shape", [(2, 16, 16), (1, 21, 13)])
def test_depth_to_space_to_rgb(shape):
    """ESPCN tail: 64 (zero-padded from 32) -> 48 channels + depth_to_space(4) -> RGB, written as 48-byte runs."""
    B, H, W = shape
    got, want = _run("fp16", B, H, W, 48, d2s=4)
    assert got.shape == (B, 4 * H, 4 * W, 3) and np.abs(got - want).max() <= 2e-3
    got, want = _run("fp16", B, H, W, 12, d2s=2, act="relu")
    assert got.shape == (B, 2 * H, 2 * W, 3) and np.abs(got - want).max() <= 2e-3


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("cout", [16, 32, 48])
@pytest.mark.parametrize("shape,act", [((2, 24, 40), "relu"), ((1, 19, 67), None)])
def test_wide_tile_kernel_with_fewer_output_channels(kind, cout, shape, act):
    """3x3 layers with 16 / 32 / 48 outputs and 16-bit results (ESPCN's 64 -> 32 layer) run on the dx-folded wide-tile
    kernel with N = 3 * cout (the 36-MMA formulation is bound by the A-operand fetch at any N <= 64); odd sizes, tiles
    clipped in x and y, idle channel quarters."""
    B, H, W = shape
    got, want = _run(kind, B, H, W, cout, out_dtype=DT[kind], act=act)
    tol = (2.0 ** -7 if kind == "bf16" else 2.0 ** -10) * max(1.0, float(np.abs(want).max()))
    assert got.shape == want.shape and np.abs(got - want).max() <= tol


def test_wide_tile_kernel_into_a_channel_slice():
    """The ESPCN arrangement: 64 -> 32 with ReLU written into channels [0, 32) of a 64-wide buffer whose upper half stays zero."""
    from srb200 import ops
    x = _round(_rand((2, 30, 33, 64), 11), "fp16")
    kern = _round(_rand((3, 3, 64, 32), 12, -0.1, 0.1), "fp16")
    bias = _rand((32,), 13, -0.1, 0.1)
    want = np.maximum(oc.conv2d_same_numpy(x, kern, bias), 0)
    wide = torch.zeros((2, 30, 33, 64), dtype=torch.float16, device="cuda")
    ops.conv2d(torch.from_numpy(x).cuda().half(), ops.ConvWeights(kern, bias), act="relu", out=wide, out_coffset=0)
    got = wide.float().cpu().numpy()
    assert np.abs(got[..., :32] - want).max() <= 2.0 ** -10 * max(1.0, float(np.abs(want).max()))
    assert not got[..., 32:].any()


def test_espcn_layers_at_benchmark_tile_size():
    """BASELINE config 2 tile size (256 x 256, 12 tiles: more 4 x 30-pixel tiles than one wave of CTAs): the wide-tile kernel
    with N = 96 (64 -> 32, ReLU, 16-bit, into a channel slice) and its depth_to_space-to-image mode (64(32) -> 48, fp32)
    against the exact CUDA-core engine on the same 16-bit inputs, and the whole network against its fp32 mode."""
    from srb200 import engine, ops, weights, _capi
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.rand((12, 256, 256, 64), device="cuda", generator=g) * 2 - 1).half()
    k2 = _round(_rand((3, 3, 64, 32), 2, -0.1, 0.1), "fp16")
    w2 = ops.ConvWeights(k2, _rand((32,), 3, -0.1, 0.1))
    a = ops.conv2d(x, w2, act="relu", out_dtype=torch.float16, engine=_capi.ENGINE_TCGEN05)
    b = ops.conv2d(x, w2, act="relu", out_dtype=torch.float32, engine=_capi.ENGINE_DIRECT)
    assert (a.float() - b).abs().max().item() <= 4e-3
    k3 = _round(_rand((3, 3, 64, 48), 4, -0.1, 0.1), "fp16")
    w3 = ops.ConvWeights(k3, _rand((48,), 5, -0.1, 0.1))
    a = ops.conv2d(x, w3, d2s=4, out_dtype=torch.float32, engine=_capi.ENGINE_TCGEN05)
    b = ops.conv2d(x, w3, d2s=4, out_dtype=torch.float32, engine=_capi.ENGINE_DIRECT)
    assert a.shape == (12, 1024, 1024, 3) and (a - b).abs().max().item() <= 2e-3
    del a, b, x
    w = weights.espcn_weights(4, bias_scale=0.05)
    img = torch.rand((4, 256, 256, 3), device="cuda", generator=g)
    lo = engine.ESPCNNet(w, 4, precision="fp16").predict_device(img)
    hi = engine.ESPCNNet(w, 4, precision="fp32").predict_device(img)
    assert (lo - hi).abs().max().item() <= 2e-2


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("case", [
    # (B, H, W, cin, cout): VGG16 blocks 2-5 (VGG16_model.py:69-83) - resident weights of every 64-channel K chunk, CTA pairs
    (2, 32, 32, 128, 128), (1, 24, 16, 128, 256), (2, 16, 16, 256, 256), (1, 16, 16, 256, 512), (3, 8, 8, 512, 512),
    (1, 19, 13, 128, 64), (1, 40, 9, 192, 64),
    # ESRGAN dense blocks (ESRGAN_model.py:230-246): cin = 64 + j * growth is padded to whole chunks; growth 32 and 8
    (1, 24, 24, 96, 32), (1, 24, 24, 160, 32), (2, 24, 24, 72, 8), (1, 24, 24, 88, 8), (1, 12, 20, 32, 48)])
def test_wide_input_layers_k_chunks(kind, case):
    """Cin != 64 on the tcgen05 engine: ReLU layers with 16-bit and fp32 outputs against the float64 im2col oracle."""
    B, H, W, cin, cout = case
    tol16 = 4e-2 if kind == "bf16" else 8e-3                 # output rounding of sums over up to 4,608 products
    got, want = _run(kind, B, H, W, cout, cin=cin, act="relu", out_dtype=DT[kind])
    assert got.shape == want.shape and np.abs(got - want).max() <= tol16 * max(1.0, np.abs(want).max() / 4)
    got, want = _run(kind, B, H, W, cout, cin=cin)
    assert np.abs(got - want).max() <= 4e-3


def test_wide_input_prefix_of_a_wider_buffer_and_fused_epilogue():
    """The ESRGAN dense-block pattern: a growth conv reads channels [0, cin) of the block's wide NHWC buffer (cstride >
    cin; the chunk padding reads further channels of the buffer against zero weight columns, or past its end where TMA
    zero-fills), writes a 16-bit slice of it, and the block's last conv adds two scaled residuals."""
    from srb200 import ops
    kind = "fp16"
    got, want = _run(kind, 1, 24, 24, 32, cin=96, x_width=192, act="relu", out_dtype=DT[kind])
    assert np.abs(got - want).max() <= 8e-3
    got, want = _run(kind, 2, 16, 24, 8, cin=72, x_width=96, act="relu")
    assert np.abs(got - want).max() <= 4e-3
    got, want = _run(kind, 1, 20, 20, 64, cin=192, alpha=0.2, res1=True)
    assert np.abs(got - want).max() <= 4e-3
    # 5x5 and 1x1 filters with a wide input
    got, want = _run(kind, 1, 18, 14, 64, cin=128, ksize=5)
    assert np.abs(got - want).max() <= 6e-3
    got, want = _run(kind, 1, 18, 14, 32, cin=128, ksize=1, act="relu")
    assert np.abs(got - want).max() <= 4e-3


@pytest.mark.parametrize("kind", ["fp16", "bf16"])
@pytest.mark.parametrize("case", [
    # (B, H, W, cin, ksize, act)
    (2, 24, 24, 3, 5, "relu"), (1, 37, 131, 3, 5, "relu"), (1, 9, 300, 3, 3, None), (3, 5, 7, 3, 3, "leaky_relu"),
    (1, 64, 256, 3, 5, "tanh"), (2, 13, 128, 1, 3, "relu"), (1, 20, 129, 4, 5, None),
])
def test_small_filter_rgb_head_on_nhwc8_rows(kind, case):
    """conv_head8_kernel (3x3 / 5x5 RGB heads: the image padded to 8 channels, rows used directly as the un-swizzled A operand)
    against the float64 im2col oracle - ragged widths around the 128-pixel tile, tiny images, 1 / 3 / 4 input channels."""
    from srb200 import ops, _capi
    B, H, W, cin, k, act = case
    x = _rand((B, H, W, cin), 11 + W)
    kern = _round(_rand((k, k, cin, 64), 12, -0.2, 0.2), kind)
    bias = _rand((64,), 13, -0.1, 0.1)
    xr = _round(x, kind)                                     # the kernel rounds the image to the operand type
    want = oc.conv2d_same_numpy(xr, kern, bias)
    if act == "relu":
        want = np.maximum(want, 0)
    elif act == "leaky_relu":
        want = np.where(want >= 0, want, 0.2 * want)
    elif act == "tanh":
        want = np.tanh(want)
    w = ops.ConvWeights(kern, bias)
    xd = torch.from_numpy(x).cuda()
    assert ops.conv2d_engine(xd, w) == _capi.ENGINE_TCGEN05 or True
    wide = torch.full((B, H, W, 96), 7.0, dtype=DT[kind], device="cuda")
    got = ops.conv2d(xd, w, act=act, act_slope=0.2, out_dtype=DT[kind])
    ops.conv2d(xd, w, act=act, act_slope=0.2, out=wide, out_coffset=16)
    torch.cuda.synchronize()
    tol = (2e-3 if kind == "fp16" else 1.6e-2) * max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got.float().cpu().numpy() - want).max())
    assert err <= tol, (case, kind, err, tol)
    sl = wide.float().cpu().numpy()
    assert float(np.abs(sl[..., 16:80] - want).max()) <= tol and np.all(sl[..., :16] == 7.0) and np.all(sl[..., 80:] == 7.0)
