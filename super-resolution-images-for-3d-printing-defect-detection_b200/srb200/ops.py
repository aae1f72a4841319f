"""Device-level operators: thin, typed wrappers over the C ABI working on CUDA torch tensors.

Every function launches on torch's current stream and returns device tensors; nothing here
synchronises or copies to the host.  Layout is NHWC everywhere.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi


def _torch():
    return capi.require_cuda()


# kernels launched by this process through the C ABI (bench.py reports it as gpu_launches)
_LAUNCHES = [0]


def reset_launch_count():
    _LAUNCHES[0] = 0


def launch_count():
    return _LAUNCHES[0]


def _check_nhwc(t, name):
    if t.dim() != 4:
        raise ValueError(f"{name} must be a 4-D NHWC tensor, got shape {tuple(t.shape)}")
    if not t.is_cuda:
        raise ValueError(f"{name} must live on a CUDA device")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


# ---------------------------------------------------------------------------------------------
# PSNR / SSIM  (metrics.py:3-7)
# ---------------------------------------------------------------------------------------------
def psnr(a, b, max_val=1.0, want_mse=False, window=capi.SSIM_TF):
    """a, b: [B,H,W,C] float32 CUDA tensors -> psnr [B] (and mse [B]): the squared-error reduction alone (no SSIM windows),
    a streaming kernel at the HBM roofline; any image size (tf.image.psnr has no minimum)."""
    torch = _torch()
    _check_nhwc(a, "y_true"); _check_nhwc(b, "y_pred")
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError("psnr inputs must be float32")
    B, H, W, Cc = a.shape
    p = torch.empty(B, dtype=torch.float32, device=a.device)
    mse = torch.empty(B, dtype=torch.float32, device=a.device) if want_mse else None
    ws_bytes = capi.lib().srb_psnr_ssim_workspace(B)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        capi.check(capi.lib().srb_psnr_ssim_window_f32(capi.ptr(a), capi.ptr(b), B, H, W, Cc, float(max_val), int(window),
                                                       capi.ptr(p), None, capi.ptr(mse), None,
                                                       capi.ptr(ws), ws_bytes, capi.stream_ptr()))
    _LAUNCHES[0] += 2
    return (p, mse) if want_mse else p


def psnr_ssim(a, b, max_val=1.0, sums=None, want_mse=False, window=capi.SSIM_TF):
    """a, b: [B,H,W,C] float32 CUDA tensors -> (psnr [B], ssim [B][, mse [B]]) float32.

    ``window``: ``SSIM_TF`` (tf.image definitions, metrics.py:3-7; wide 1- / 3-channel images run on the tensor path with
    the Gaussian rounded to an fp16 window that sums to 1, |dSSIM| <= ~3e-6), ``SSIM_TF_EXACT`` (the float32 Gaussian on the
    CUDA cores for every size) or ``SSIM_SKIMAGE`` (skimage.metrics definitions
    with their defaults, as used by super_resolucion_clasica.ipynb cell 7).

    ``sums`` (optional float64[4] CUDA tensor) is accumulated with (sum psnr, sum ssim, count,
    sum mse) for sharded evaluation."""
    torch = _torch()
    _check_nhwc(a, "y_true"); _check_nhwc(b, "y_pred")
    if a.shape != b.shape:
        raise ValueError(f"shape mismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
    if a.dtype != torch.float32 or b.dtype != torch.float32:
        raise TypeError("psnr/ssim inputs must be float32")
    B, H, W, Cc = a.shape
    psnr = torch.empty(B, dtype=torch.float32, device=a.device)
    ssim = torch.empty(B, dtype=torch.float32, device=a.device)
    mse = torch.empty(B, dtype=torch.float32, device=a.device) if want_mse else None
    ws_bytes = capi.lib().srb_psnr_ssim_workspace(B)
    ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        capi.check(capi.lib().srb_psnr_ssim_window_f32(capi.ptr(a), capi.ptr(b), B, H, W, Cc, float(max_val), int(window),
                                                       capi.ptr(psnr), capi.ptr(ssim), capi.ptr(mse), capi.ptr(sums),
                                                       capi.ptr(ws), ws_bytes, capi.stream_ptr()))
    _LAUNCHES[0] += 2
    return (psnr, ssim, mse) if want_mse else (psnr, ssim)


# ---------------------------------------------------------------------------------------------
# bicubic  (classic_algorithms.py:11-13)
# ---------------------------------------------------------------------------------------------
def bicubic(src, dst_h, dst_w, clip01=False, fixed_point=False):
    """src: [B,H,W,C] float32 or uint8 CUDA tensor -> [B,dst_h,dst_w,C], same dtype."""
    torch = _torch()
    _check_nhwc(src, "src")
    B, H, W, Cc = src.shape
    dst = torch.empty((B, int(dst_h), int(dst_w), Cc), dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        if src.dtype == torch.float32:
            capi.check(capi.lib().srb_bicubic_f32(capi.ptr(src), B, H, W, Cc, capi.ptr(dst), int(dst_h), int(dst_w),
                                                  int(bool(clip01)), capi.stream_ptr()))
        elif src.dtype == torch.uint8:
            capi.check(capi.lib().srb_bicubic_u8(capi.ptr(src), B, H, W, Cc, capi.ptr(dst), int(dst_h), int(dst_w),
                                                 int(bool(fixed_point)), capi.stream_ptr()))
        else:
            raise TypeError(f"bicubic supports float32 and uint8, got {src.dtype}")
    _LAUNCHES[0] += 3
    return dst


def resize(src, dst_h, dst_w, interpolation=capi.INTER_CUBIC, clip01=False):
    """cv2.resize(src, (dst_w, dst_h), interpolation) for float32 or uint8 NHWC CUDA tensors; INTER_CUBIC, INTER_LINEAR and
    (up-scaling) INTER_AREA share the bicubic kernels through the four-tap table; INTER_LANCZOS4 has an eight-tap one."""
    torch = _torch()
    _check_nhwc(src, "src")
    if src.dtype not in (torch.float32, torch.uint8):
        raise TypeError(f"resize supports float32 and uint8, got {src.dtype}")
    B, H, W, Cc = src.shape
    dst = torch.empty((B, int(dst_h), int(dst_w), Cc), dtype=src.dtype, device=src.device)
    with torch.cuda.device(src.device):
        if src.dtype == torch.uint8:            # OpenCV's fixed-point uint8 paths (bit-exact)
            capi.check(capi.lib().srb_resize_u8(capi.ptr(src), B, H, W, Cc, capi.ptr(dst), int(dst_h), int(dst_w),
                                                int(interpolation), capi.stream_ptr()))
        else:
            capi.check(capi.lib().srb_resize_f32(capi.ptr(src), B, H, W, Cc, capi.ptr(dst), int(dst_h), int(dst_w),
                                                 int(interpolation), int(bool(clip01)), capi.stream_ptr()))
    _LAUNCHES[0] += 3
    return dst


# ---------------------------------------------------------------------------------------------
# tiling  (loading_methods.py:6-26, EDSR_model.py:201-256)
# ---------------------------------------------------------------------------------------------
def pad_extract(image, patch, stride):
    """image: [H,W,C] float32 CUDA -> (patches [ny*nx,P,P,C], (padded_h, padded_w, ny, nx))."""
    torch = _torch()
    if image.dim() != 3 or not image.is_cuda or image.dtype != torch.float32:
        raise ValueError("image must be a float32 [H,W,C] CUDA tensor")
    image = image.contiguous()
    H, W, Cc = image.shape
    ph, pw, ny, nx = capi.tiling_geometry(H, W, patch, stride)
    patches = torch.empty((ny * nx, patch, patch, Cc), dtype=torch.float32, device=image.device)
    with torch.cuda.device(image.device):
        capi.check(capi.lib().srb_pad_extract_f32(capi.ptr(image), H, W, Cc, int(patch), int(stride),
                                                  capi.ptr(patches), capi.stream_ptr()))
    _LAUNCHES[0] += 1
    return patches, (ph, pw, ny, nx)


def overlap_add(patches, ny, nx, stride_out, out_h, out_w):
    """patches: [ny*nx,P,P,C] float32 CUDA -> [out_h,out_w,C] overlap-averaged, clipped to [0,1]."""
    torch = _torch()
    _check_nhwc(patches, "patches")
    if patches.dtype != torch.float32:
        raise TypeError("patches must be float32")
    n, P, P2, Cc = patches.shape
    if n != ny * nx or P != P2:
        raise ValueError("patch tensor does not match the grid")
    out = torch.empty((int(out_h), int(out_w), Cc), dtype=torch.float32, device=patches.device)
    with torch.cuda.device(patches.device):
        capi.check(capi.lib().srb_overlap_add_f32(capi.ptr(patches), int(ny), int(nx), P, int(stride_out), Cc,
                                                  capi.ptr(out), int(out_h), int(out_w), capi.stream_ptr()))
    _LAUNCHES[0] += 1
    return out


# ---------------------------------------------------------------------------------------------
# convolution  (Keras Conv2D + fused epilogue)
# ---------------------------------------------------------------------------------------------
class ConvWeights:
    """Device-resident packed Conv2D kernel (HWIO float32 in, Keras layout)."""

    def __init__(self, kernel_hwio, bias=None, prelu=None):
        torch = _torch()
        k = np.ascontiguousarray(kernel_hwio, dtype=np.float32)
        if k.ndim != 4:
            raise ValueError("kernel must be HWIO [kh, kw, cin, cout]")
        self.kh, self.kw, self.cin, self.cout = (int(v) for v in k.shape)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        if b is not None and b.shape != (self.cout,):
            raise ValueError("bias must be [cout]")
        handle = C.c_void_p()
        capi.check(capi.lib().srb_conv_weights_create(
            k.ctypes.data_as(C.c_void_p), None if b is None else b.ctypes.data_as(C.c_void_p),
            self.kh, self.kw, self.cin, self.cout, C.byref(handle)))
        self._handle = handle
        self.prelu = None
        if prelu is not None:
            self.prelu = torch.from_numpy(np.ascontiguousarray(prelu, dtype=np.float32)).cuda()
        self.n_params = k.size + (0 if b is None else b.size) + (0 if prelu is None else np.size(prelu))

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                capi.lib().srb_conv_weights_destroy(h)
            except Exception:
                pass


def conv2d(x, w: ConvWeights, act=None, act_slope=0.0, alpha=1.0, res1=None, beta1=1.0, res2=None, beta2=1.0,
           clip01=False, d2s=1, out_dtype=None, out=None, out_coffset=0, x_coffset=0, engine=capi.ENGINE_AUTO,
           out2_dtype=None, out2_error=False, out2=None):
    """y = clip(alpha * act(conv(x, W) + b) + beta1 * res1 + beta2 * res2), optionally depth_to_space'd.

    ``x`` may be a wider NHWC buffer of which channels [x_coffset, x_coffset + cin) are read, and
    ``out`` a wider buffer written at ``out_coffset`` (concat-free dense blocks).  With ``out2_dtype`` the
    result is also written in a second dtype and ``(out, out2)`` is returned (fp32 residual trunk next to
    the 16-bit operand of the following layer); with ``out2_error`` the second output is the rounding error
    ``v - round(out)`` instead, so that ``out + out2`` carries the value to ~22 bits in two 16-bit tensors.
    ``out`` / ``out2`` may be the residual tensors themselves (``out is res1``, ``out2 is res2``): every output pixel
    depends on the residual of the same pixel only, which its own tile reads before it writes - the res-block trunk is
    updated in place."""
    torch = _torch()
    _check_nhwc(x, "x")
    B, H, W, Cx = x.shape
    r = int(d2s)
    c_post = w.cout // (r * r)
    if out is None:
        out = torch.empty((B, H * r, W * r, c_post), dtype=out_dtype or x.dtype, device=x.device)
    else:
        _check_nhwc(out, "out")
    a = capi.ConvArgs()
    a.x, a.x_dtype, a.x_cstride, a.x_coffset = x.data_ptr(), capi.dtype_code(x), Cx, int(x_coffset)
    a.y, a.y_dtype, a.y_cstride, a.y_coffset = out.data_ptr(), capi.dtype_code(out), out.shape[3], int(out_coffset)
    if out2 is not None:
        _check_nhwc(out2, "out2")
        if tuple(out2.shape) != (B, H * r, W * r, c_post):
            raise ValueError(f"out2 shape {tuple(out2.shape)} does not match the output")
    elif out2_dtype is not None:
        out2 = torch.empty((B, H * r, W * r, c_post), dtype=out2_dtype, device=x.device)
    if out2 is not None:
        a.y2, a.y2_dtype, a.y2_cstride = out2.data_ptr(), capi.dtype_code(out2), c_post
        a.y2_mode = 1 if out2_error else 0
    a.batch, a.height, a.width = B, H, W
    a.weights = w.handle
    a.act = capi.ACTIVATIONS[act] if not isinstance(act, int) else act
    a.act_slope = float(act_slope)
    a.prelu = w.prelu.data_ptr() if (a.act == capi.ACT_PRELU and w.prelu is not None) else None
    a.alpha = float(alpha)
    for i, (res, beta) in enumerate(((res1, beta1), (res2, beta2)), start=1):
        if res is not None:
            _check_nhwc(res, f"res{i}")
            if tuple(res.shape[:3]) != (B, H * r, W * r) or res.shape[3] < c_post:
                raise ValueError(f"res{i} shape {tuple(res.shape)} does not match the output")
            setattr(a, f"res{i}", res.data_ptr())
            setattr(a, f"res{i}_dtype", capi.dtype_code(res))
            setattr(a, f"res{i}_cstride", res.shape[3])
        setattr(a, f"beta{i}", float(beta))
    a.clip01, a.d2s, a.engine = int(bool(clip01)), r, int(engine)
    with torch.cuda.device(x.device):
        capi.check(capi.lib().srb_conv2d_nhwc(C.byref(a), capi.stream_ptr()))
    _LAUNCHES[0] += 1
    return out if out2 is None else (out, out2)


class ComposedUpsampler:
    """Device-resident composed EDSR up-sampling tail (``srb200.compose.compose_edsr_tail`` -> ``srb_upsampler_create``):
    w [3, 3, 5, 5, 64, r*r*C], bias [3, 3, r*r*C]."""

    def __init__(self, w, bias, scale, w_scale=1.0):
        _torch()
        w = np.ascontiguousarray(w, dtype=np.float32)
        b = np.ascontiguousarray(bias, dtype=np.float32)
        if w.ndim != 6 or w.shape[:4] != (3, 3, 5, 5) or b.shape != (3, 3, w.shape[5]):
            raise ValueError("composed up-sampler: w must be [3, 3, 5, 5, cin, r*r*C] and bias [3, 3, r*r*C]")
        self.scale, self.cin, self.cout = int(scale), int(w.shape[4]), int(w.shape[5])
        if self.cout % (self.scale * self.scale):
            raise ValueError("composed up-sampler: output channels are not a multiple of scale^2")
        self.c_img = self.cout // (self.scale * self.scale)
        handle = C.c_void_p()
        capi.check(capi.lib().srb_upsampler_create(w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), self.cin,
                                                   self.c_img, self.scale, float(w_scale), C.byref(handle)))
        self._handle = handle

    @property
    def handle(self):
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h:
            try:
                capi.lib().srb_upsampler_destroy(h)
            except Exception:
                pass


def upsample_composed(x, up: ComposedUpsampler, clip01=True, out_dtype=None, x_coffset=0, out=None):
    """x [B, H, W, >= 64] fp16 / bf16 -> [B, H*r, W*r, C] (float32 by default): the whole EDSR up-sampling tail
    (EDSR_model.py:117-123) as one tcgen05 launch."""
    torch = _torch()
    _check_nhwc(x, "x")
    B, H, W, Cx = x.shape
    shape = (B, H * up.scale, W * up.scale, up.c_img)
    if out is None:
        out = torch.empty(shape, dtype=out_dtype or torch.float32, device=x.device)
    else:
        _check_nhwc(out, "out")
        if tuple(out.shape) != shape:
            raise ValueError(f"out must have shape {shape}, got {tuple(out.shape)}")
    with torch.cuda.device(x.device):
        capi.check(capi.lib().srb_upsample_composed(up.handle, capi.ptr(x), capi.dtype_code(x), Cx, int(x_coffset), B, H, W,
                                                    capi.ptr(out), capi.dtype_code(out), int(bool(clip01)), capi.stream_ptr()))
    _LAUNCHES[0] += 1
    return out


def conv2d_engine(x, w: ConvWeights, d2s=1):
    """Which engine AUTO dispatch picks for this input/kernel (capi.ENGINE_*)."""
    a = capi.ConvArgs()
    B, H, W, Cx = x.shape
    a.x, a.x_dtype, a.x_cstride = x.data_ptr(), capi.dtype_code(x), Cx
    a.y, a.y_dtype = x.data_ptr(), capi.dtype_code(x)
    a.batch, a.height, a.width, a.weights, a.d2s = B, H, W, w.handle, int(d2s)
    rc = capi.lib().srb_conv2d_engine(C.byref(a))
    if rc < 0:
        capi.check(rc)
    return rc


def cast(x, dtype, scale=1.0, shift=0.0, relu=False):
    """dst = x * scale + shift (then max(., 0) with ``relu``) converted to ``dtype``."""
    torch = _torch()
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    fn = capi.lib().srb_cast_relu if relu else capi.lib().srb_cast
    with torch.cuda.device(x.device):
        capi.check(fn(capi.ptr(x), capi.dtype_code(x), capi.ptr(out), capi.dtype_code(out),
                      x.numel(), float(scale), float(shift), capi.stream_ptr()))
    _LAUNCHES[0] += 1
    return out


def maxpool2x2(x):
    torch = _torch()
    _check_nhwc(x, "x")
    B, H, W, Cc = x.shape
    out = torch.empty((B, H // 2, W // 2, Cc), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(capi.lib().srb_maxpool2x2_nhwc(capi.ptr(x), capi.dtype_code(x), B, H, W, Cc, capi.ptr(out),
                                                  capi.stream_ptr()))
    return out


def gap_dense_softmax(x, w1, b1, w2, b2):
    """GlobalAveragePooling2D -> Dense(relu) -> Dense(softmax) (VGG16_model.py:84-97)."""
    torch = _torch()
    _check_nhwc(x, "x")
    B, H, W, Cc = x.shape
    hidden, classes = w1.shape[1], w2.shape[1]
    probs = torch.empty((B, classes), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(capi.lib().srb_gap_dense_softmax(capi.ptr(x), capi.dtype_code(x), B, H * W, Cc, capi.ptr(w1),
                                                    capi.ptr(b1), hidden, capi.ptr(w2), capi.ptr(b2), classes,
                                                    capi.ptr(probs), capi.stream_ptr()))
    return probs


def self_attention_tc_eligible(dk, dv):
    return 1 <= dk <= 16 and dv in (16, 32, 64)


def self_attention_core(f, g, h, tensor_cores=False):
    """o = softmax(g f^T) h per image.  f, g: [B,HW,dk]; h: [B,HW,dv] float32 (ESRGAN_model.py:58-66).
    ``tensor_cores``: both products on tcgen05 (fp16 hi/lo-split scores, fp16 P and h, float32 accumulation) - the 16-bit
    precision modes; default: exact float32 on the CUDA cores."""
    torch = _torch()
    B, HW, dk = f.shape
    dv = h.shape[2]
    o = torch.empty((B, HW, dv), dtype=torch.float32, device=f.device)
    if tensor_cores:
        if not self_attention_tc_eligible(dk, dv):
            raise NotImplementedError(f"tensor-core attention needs dk <= 16 and dv in (16, 32, 64); got dk={dk}, dv={dv}")
        ws_bytes = capi.lib().srb_self_attention_tc_workspace(B, HW, dk, dv)
        ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=f.device)
        with torch.cuda.device(f.device):
            capi.check(capi.lib().srb_self_attention_tc(capi.ptr(f.contiguous()), capi.ptr(g.contiguous()), capi.ptr(h.contiguous()),
                                                        B, HW, dk, dv, capi.ptr(o), capi.ptr(ws), ws_bytes, capi.stream_ptr()))
        _LAUNCHES[0] += 2
        return o
    with torch.cuda.device(f.device):
        capi.check(capi.lib().srb_self_attention_f32(capi.ptr(f.contiguous()), capi.ptr(g.contiguous()),
                                                     capi.ptr(h.contiguous()), B, HW, dk, dv, capi.ptr(o),
                                                     capi.stream_ptr()))
    return o
