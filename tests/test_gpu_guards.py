"""GPU: out-of-bounds canaries.  compute-sanitizer is closed on the GPU pool this repo is built on, so memory safety of the
hand-written kernels is checked the way its refusal message asks for: every OUTPUT is a slice of a larger allocation whose
surroundings carry a sentinel that must survive the launch, every INPUT is a slice of an allocation whose surroundings are
NaN (float) / 255 (uint8), so that a read past the tensor that reaches a result shows up in it, and the results are compared
with the CPU oracle as in the parity tests.  Shapes are ragged on purpose (partial tiles at the right / bottom edge, channel
slices of wider buffers, batches that do not fill a launch)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import bicubic as ob, convnets as oc, metrics as om, tiling as ot

pytestmark = pytest.mark.gpu

G = 4096                                  # guard elements on either side


def _guarded(shape, dtype, data=None, poison=None):
    """-> (view, check): `view` is a contiguous tensor of `shape` inside a larger flat buffer; check() asserts the guards."""
    n = int(np.prod(shape))
    flat = torch.empty(n + 2 * G, dtype=dtype, device="cuda")
    if poison is None:
        poison = 77 if dtype == torch.uint8 else -12345.0
    flat.fill_(poison)
    view = flat[G:G + n].view(shape)
    if data is not None:
        view.copy_(torch.from_numpy(np.ascontiguousarray(data)).to(dtype) if isinstance(data, np.ndarray) else data)
    snapshot = torch.cat([flat[:G], flat[G + n:]]).clone()

    def check():
        now = torch.cat([flat[:G], flat[G + n:]])
        same = (now == snapshot) | (now != now)            # NaN poison compares unequal to itself
        assert bool(same.all()), "a kernel wrote outside its output tensor"
    return view, check


def _nan_in(shape, dtype, data):
    poison = 255 if dtype == torch.uint8 else float("nan")
    return _guarded(shape, dtype, data, poison)[0]


@pytest.mark.parametrize("case", [
    # (B, H, W, cin, cout, ksize, in dtype, out dtype, kwargs)
    (2, 21, 13, 64, 64, 3, "fp16", "fp16", dict(act="relu")),            # wide-tile fold kernel, partial tiles
    (1, 19, 29, 64, 128, 3, "fp16", "fp16", dict()),                     # 16 x 8 tiles, TMA epilogue
    (1, 16, 24, 64, 256, 3, "bf16", "bf16", dict(d2s=2)),                # CTA pairs, depth_to_space store
    (3, 9, 11, 64, 3, 3, "fp16", "fp32", dict(clip01=True)),             # RGB tail
    (2, 9, 11, 64, 3, 3, "fp16", "u8", dict(clip01=True)),               # RGB tail, quantised output
    (1, 17, 9, 128, 128, 3, "fp16", "fp16", dict(act="relu")),           # two K chunks, CTA pairs
    (2, 8, 8, 512, 512, 3, "fp16", "fp16", dict(act="relu")),            # eight K chunks
    (1, 13, 21, 96, 32, 3, "fp16", "fp16", dict(act="relu")),            # padded K chunk read past the pixel's end
    (2, 14, 10, 3, 64, 5, "fp32", "fp16", dict(act="relu")),             # im2col head kernel
    (1, 11, 7, 3, 64, 3, "fp32", "fp32", dict()),                        # CUDA-core head kernel
    (1, 10, 12, 32, 3, 5, "fp32", "fp32", dict()),                       # CUDA-core direct engine
])
def test_conv_engines_stay_inside_their_tensors(case):
    from srb200 import ops
    B, H, W, cin, cout, k, din, dout, kw = case
    dts = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32, "u8": torch.uint8}
    rng = np.random.default_rng(cin * 131 + cout)
    x = rng.uniform(-1, 1, (B, H, W, cin)).astype(np.float32)
    kern = rng.uniform(-0.1, 0.1, (k, k, cin, cout)).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, (cout,)).astype(np.float32)
    xd = _nan_in((B, H, W, cin), dts[din], x)
    r = kw.get("d2s", 1)
    out, check = _guarded((B, H * r, W * r, cout // (r * r)), dts[dout])
    ops.conv2d(xd, ops.ConvWeights(kern, bias), out=out, **kw)
    torch.cuda.synchronize()
    check()
    xr = xd.float().cpu().numpy()                                         # the operand as the kernel saw it
    want = oc.conv2d_same_numpy(xr, kern, bias)
    if kw.get("act") == "relu":
        want = np.maximum(want, 0)
    if r > 1:
        want = oc.depth_to_space_numpy(want, r)
    if kw.get("clip01"):
        want = np.clip(want, 0, 1)
    got = out.float().cpu().numpy()
    assert np.isfinite(got).all(), "a read outside the input tensor reached the result"
    if dout == "u8":
        assert np.abs(got - np.rint(want * 255)).max() <= 2
    else:
        assert np.abs(got - want).max() <= (3e-2 if din != "fp32" or dout != "fp32" else 1e-3) * max(1.0, np.abs(want).max())


def test_pair8_trunk_conv_stays_inside_its_tensors():
    """The residual-trunk layer: 16-bit hi + e5m2 lo residual pair in, y + rounding error out, four TMA-moved tensors."""
    from srb200 import ops
    rng = np.random.default_rng(7)
    B, H, W = 2, 19, 13
    x = rng.uniform(-1, 1, (B, H, W, 64)).astype(np.float32)
    hi = rng.uniform(-1, 1, (B, H, W, 64)).astype(np.float32)
    kern = rng.uniform(-0.1, 0.1, (3, 3, 64, 64)).astype(np.float32)
    xd, hd = _nan_in((B, H, W, 64), torch.float16, x), _nan_in((B, H, W, 64), torch.float16, hi)
    lo = torch.zeros((B, H, W, 64), dtype=torch.float16, device="cuda").to(torch.float8_e5m2)
    y, e = ops.conv2d(xd, ops.ConvWeights(kern, None), alpha=0.1, res1=hd, res2=lo, out_dtype=torch.float16,
                      out2_dtype=torch.float8_e5m2, out2_error=True)
    torch.cuda.synchronize()
    want = 0.1 * oc.conv2d_same_numpy(xd.float().cpu().numpy(), kern) + hd.float().cpu().numpy()
    got = y.float().cpu().numpy() + e.float().cpu().numpy()
    assert np.isfinite(got).all() and np.abs(got - want).max() <= 2e-3


@pytest.mark.parametrize("dtype", ["f32", "u8"])
@pytest.mark.parametrize("code", [1, 2, 3, 4])
def test_resize_kernels_stay_inside_their_tensors(dtype, code):
    """All four interpolation codes, float32 and uint8, ragged sizes, through the C ABI with guarded device buffers."""
    from srb200 import _capi as capi
    rng = np.random.default_rng(code)
    B, h, w, c, dh, dw = 2, 37, 29, 3, 91, 70
    if dtype == "f32":
        src = rng.random((B, h, w, c), dtype=np.float32)
        s = _nan_in(src.shape, torch.float32, src)
        d, check = _guarded((B, dh, dw, c), torch.float32)
        capi.check(capi.lib().srb_resize_f32(capi.ptr(s), B, h, w, c, capi.ptr(d), dh, dw, code, 0, capi.stream_ptr()))
    else:
        src = rng.integers(0, 255, (B, h, w, c), dtype=np.uint8)          # (255 is the poison value)
        s = _nan_in(src.shape, torch.uint8, src)
        d, check = _guarded((B, dh, dw, c), torch.uint8)
        capi.check(capi.lib().srb_resize_u8(capi.ptr(s), B, h, w, c, capi.ptr(d), dh, dw, code, capi.stream_ptr()))
    torch.cuda.synchronize()
    check()
    got = d.cpu().numpy()
    if dtype == "f32":
        assert np.isfinite(got).all()
    ref = {1: (lambda im: ob.resize_linear_u8(im, (dw, dh))), 3: (lambda im: ob.resize_linear_u8(im, (dw, dh), area=True)),
           4: (lambda im: ob.resize_lanczos4_u8(im, (dw, dh))), 2: (lambda im: ob.resize_cubic_u8(im, (dw, dh)))}
    if dtype == "u8":
        want = np.stack([ref[code](im) for im in src])
        assert np.abs(got.astype(int) - want.astype(int)).max() <= (1 if code == 2 else 0)
    elif code == 2:
        assert np.abs(got - np.stack([ob.resize_cubic_f32(im, (dw, dh)) for im in src])).max() <= 2e-6
    elif code == 4:
        assert np.abs(got - np.stack([ob.resize_lanczos4_f32(im, (dw, dh)) for im in src])).max() <= 2e-6


def test_metrics_and_tiling_kernels_stay_inside_their_tensors():
    from srb200 import _capi as capi
    rng = np.random.default_rng(3)
    B, H, W = 3, 45, 131                                   # wide enough for the pair kernel (map width 121 >= 96)
    a = rng.random((B, H, W, 3), dtype=np.float32)
    b = np.clip(a + 0.05 * rng.standard_normal(a.shape).astype(np.float32), 0, 1)
    ad, bd = _nan_in(a.shape, torch.float32, a), _nan_in(b.shape, torch.float32, b)
    for (hh, ww) in ((H, W), (23, 31)):                    # pair kernel / narrow kernel
        ax, bx = ad[:, :hh, :ww].contiguous(), bd[:, :hh, :ww].contiguous()
        ax, bx = _nan_in(ax.shape, torch.float32, ax), _nan_in(bx.shape, torch.float32, bx)
        p, cp = _guarded((B,), torch.float32)
        s, cs = _guarded((B,), torch.float32)
        ws_bytes = capi.lib().srb_psnr_ssim_workspace(B)
        ws, cw = _guarded((ws_bytes,), torch.uint8)
        capi.check(capi.lib().srb_psnr_ssim_f32(capi.ptr(ax), capi.ptr(bx), B, hh, ww, 3, 1.0, capi.ptr(p), capi.ptr(s), None, None,
                                                capi.ptr(ws), ws_bytes, capi.stream_ptr()))
        torch.cuda.synchronize()
        cp(); cs(); cw()
        an, bn = a[:, :hh, :ww], b[:, :hh, :ww]
        assert np.abs(p.cpu().numpy() - om.psnr(an, bn, dtype=np.float64)).max() <= 0.01
        assert np.abs(s.cpu().numpy() - om.ssim(an, bn, dtype=np.float64)).max() <= 1e-4
    img = rng.random((50, 43, 3), dtype=np.float32)
    imd = _nan_in(img.shape, torch.float32, img)
    ph, pw, ny, nx = capi.tiling_geometry(50, 43, 24, 12)
    patches, check = _guarded((ny * nx, 24, 24, 3), torch.float32)
    capi.check(capi.lib().srb_pad_extract_f32(capi.ptr(imd), 50, 43, 3, 24, 12, capi.ptr(patches), capi.stream_ptr()))
    torch.cuda.synchronize()
    check()
    want, pos = ot.extract_patches(ot.add_padding(img, 24, 12), 24, 12)
    assert np.array_equal(patches.cpu().numpy(), want)
    out, check = _guarded((50, 43, 3), torch.float32)
    pin = _nan_in(patches.shape, torch.float32, patches)
    capi.check(capi.lib().srb_overlap_add_f32(capi.ptr(pin), ny, nx, 24, 12, 3, capi.ptr(out), 50, 43, capi.stream_ptr()))
    torch.cuda.synchronize()
    check()
    assert np.abs(out.cpu().numpy() - ot.reconstruct(want, pos, (ph, pw, 3), (50, 43), 1)).max() <= 1e-6


@pytest.mark.parametrize("scale,shape,dout", [(4, (2, 13, 21), "fp32"), (4, (1, 2, 2), "u8"), (4, (1, 31, 57), "fp16"),
                                               (3, (1, 9, 5), "fp32"), (2, (2, 7, 30), "fp32")])
def test_composed_upsampler_stays_inside_its_tensors(scale, shape, dout):
    """upsample5_fold_kernel: nine segments with their own rectangles, per-lane vector stores into the image - every one of
    them inside the output, and nothing outside the 64-channel input slice reaches a result (NaN surroundings)."""
    from srb200 import compose, ops, weights
    dts = {"fp16": torch.float16, "fp32": torch.float32, "u8": torch.uint8}
    w = weights.edsr_weights(scale, num_res_blocks=1, bias_scale=0.05, seed=scale)
    wc, bc = compose.compose_edsr_tail(w, scale)
    up = ops.ComposedUpsampler(wc, bc, scale, compose.weight_scale(wc))
    B, H, W = shape
    x = np.random.default_rng(H * W).uniform(-0.3, 0.3, (B, H, W, 64)).astype(np.float32)
    xd = _nan_in((B, H, W, 64), torch.float16, x)
    out, check = _guarded((B, H * scale, W * scale, 3), dts[dout])
    ops.upsample_composed(xd, up, clip01=True, out=out)
    torch.cuda.synchronize()
    check()
    want = np.clip(compose.layered_tail(w, xd.float().cpu().numpy().astype(np.float64), scale), 0, 1)
    got = out.float().cpu().numpy()
    if dout == "u8":
        assert np.abs(got - want * 255).max() <= 0.5 + 255 * 4e-3
    else:
        assert np.isfinite(got).all() and np.abs(got - want).max() <= 4e-3 + (1e-3 if dout == "fp16" else 0)


def test_in_place_trunk_update_stays_inside_its_tensors():
    """Res-block second conv with the (h, e) trunk pair updated in place (out is res1, out2 is res2)."""
    from srb200 import ops
    rng = np.random.default_rng(5)
    B, H, W = 2, 19, 13
    t = rng.uniform(-1, 1, (B, H, W, 64)).astype(np.float32)
    kern = rng.uniform(-0.1, 0.1, (3, 3, 64, 64)).astype(np.float32)
    bias = rng.uniform(-0.1, 0.1, (64,)).astype(np.float32)
    h0 = rng.uniform(-1, 1, (B, H, W, 64)).astype(np.float32)
    td = _nan_in((B, H, W, 64), torch.float16, t)
    h, check_h = _guarded((B, H, W, 64), torch.float16, h0)
    e, check_e = _guarded((B, H, W, 64), torch.float8_e5m2, torch.zeros((B, H, W, 64), device="cuda").to(torch.float8_e5m2))
    h_before = h.float().cpu().numpy()
    ops.conv2d(td, ops.ConvWeights(kern, bias), alpha=0.1, res1=h, res2=e, out=h, out2=e, out2_error=True)
    torch.cuda.synchronize()
    check_h(); check_e()
    want = 0.1 * oc.conv2d_same_numpy(td.float().cpu().numpy(), kern, bias) + h_before
    got = h.float().cpu().numpy() + e.float().cpu().numpy()
    assert np.abs(got - want).max() <= 2e-3
