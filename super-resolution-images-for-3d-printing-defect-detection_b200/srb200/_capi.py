"""ctypes binding of ``libsrb200.so`` (the C ABI declared in ``include/srb200.h``).

This is the only module that touches the shared library.  PyTorch is used for device memory,
streams and ``torch.distributed`` only - every numerical result comes out of the hand-written
sm_100a kernels.  There is no CPU fallback: importing works anywhere (so that host-side logic can
be tested), but the first compute call without a CUDA device or without the built library raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsrb200.so")

SRB_OK, SRB_E_INVALID, SRB_E_UNSUPPORTED, SRB_E_CUDA, SRB_E_NOMEM = 0, -1, -2, -3, -4
F32, BF16, U8, F16, F8E5M2 = 0, 1, 2, 3, 4
INTER_LINEAR, INTER_CUBIC, INTER_AREA, INTER_LANCZOS4 = 1, 2, 3, 4          # OpenCV's interpolation codes
SSIM_TF, SSIM_SKIMAGE, SSIM_TF_EXACT = 0, 1, 2                            # srb_psnr_ssim_window_f32 window kinds
ACT_NONE, ACT_RELU, ACT_PRELU, ACT_LEAKY, ACT_TANH = 0, 1, 2, 3, 4
ENGINE_AUTO, ENGINE_DIRECT, ENGINE_TCGEN05 = 0, 1, 2
ACTIVATIONS = {None: ACT_NONE, "linear": ACT_NONE, "relu": ACT_RELU, "prelu": ACT_PRELU,
               "leaky_relu": ACT_LEAKY, "tanh": ACT_TANH}


class ConvArgs(C.Structure):
    """Mirror of ``struct srb_conv_args``."""
    _fields_ = [
        ("x", C.c_void_p), ("x_dtype", C.c_int), ("x_cstride", C.c_int), ("x_coffset", C.c_int),
        ("y", C.c_void_p), ("y_dtype", C.c_int), ("y_cstride", C.c_int), ("y_coffset", C.c_int),
        ("y2", C.c_void_p), ("y2_dtype", C.c_int), ("y2_cstride", C.c_int), ("y2_mode", C.c_int),
        ("batch", C.c_int), ("height", C.c_int), ("width", C.c_int),
        ("weights", C.c_void_p),
        ("act", C.c_int), ("act_slope", C.c_float), ("prelu", C.c_void_p),
        ("alpha", C.c_float),
        ("res1", C.c_void_p), ("res1_dtype", C.c_int), ("res1_cstride", C.c_int), ("beta1", C.c_float),
        ("res2", C.c_void_p), ("res2_dtype", C.c_int), ("res2_cstride", C.c_int), ("beta2", C.c_float),
        ("clip01", C.c_int),
        ("d2s", C.c_int),
        ("engine", C.c_int),
    ]


# name -> (restype, argtypes); must list every symbol of include/srb200.h (tests check this)
_SIGNATURES = {
    "srb_last_error": (C.c_char_p, []),
    "srb_version": (C.c_int, []),
    "srb_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "srb_psnr_ssim_workspace": (C.c_size_t, [C.c_int]),
    "srb_psnr_ssim_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                    C.c_void_p]),
    "srb_psnr_ssim_window_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                           C.c_void_p]),
    "srb_bicubic_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p]),
    "srb_resize_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p]),
    "srb_bicubic_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_void_p]),
    "srb_resize_u8": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                C.c_int, C.c_void_p]),
    "srb_tiling_geometry": (C.c_int, [C.c_int] * 4 + [C.POINTER(C.c_int)] * 4),
    "srb_pad_extract_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "srb_overlap_add_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int, C.c_void_p]),
    "srb_conv_weights_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(C.c_void_p)]),
    "srb_conv_weights_destroy": (None, [C.c_void_p]),
    "srb_conv2d_nhwc": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "srb_conv2d_engine": (C.c_int, [C.POINTER(ConvArgs)]),
    "srb_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_void_p]),
    "srb_cast_relu": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_void_p]),
    "srb_maxpool2x2_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p]),
    "srb_gap_dense_softmax": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "srb_self_attention_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_void_p, C.c_void_p]),
    "srb_conv_tc_set_cta_pairs": (C.c_int, [C.c_int]),
    "srb_self_attention_tc_workspace": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "srb_self_attention_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "srb_upsampler_create": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.POINTER(C.c_void_p)]),
    "srb_upsampler_destroy": (None, [C.c_void_p]),
    "srb_upsample_composed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """Load ``libsrb200.so`` (once).  Raises LibraryMissing with build instructions if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  srb200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def last_error() -> str:
    return lib().srb_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc == SRB_OK:
        return
    msg = last_error()
    if rc == SRB_E_INVALID:
        raise ValueError(msg)
    if rc == SRB_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == SRB_E_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("srb200 needs a CUDA device (NVIDIA B200, sm_100a); there is no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t):
    import torch
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float16:
        return F16
    if t.dtype == torch.uint8:
        return U8
    if t.dtype == torch.float8_e5m2:
        return F8E5M2
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def tiling_geometry(h, w, patch, stride):
    """Host-only helper: (padded_h, padded_w, ny, nx) by the reference's add_padding rule."""
    out = [C.c_int() for _ in range(4)]
    check(lib().srb_tiling_geometry(int(h), int(w), int(patch), int(stride), *[C.byref(o) for o in out]))
    return tuple(o.value for o in out)
