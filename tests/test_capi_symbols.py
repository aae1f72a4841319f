"""CPU: the C-ABI library loads and exports every symbol include/srb200.h declares (no compute)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "srb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(srb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for must in ("srb_conv2d_nhwc", "srb_bicubic_f32", "srb_bicubic_u8", "srb_psnr_ssim_f32",
                 "srb_pad_extract_f32", "srb_overlap_add_f32", "srb_conv_weights_create"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from srb200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _capi.lib()
    declared = _declared()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in srb200.h but not exported"
    assert sorted(_capi.exported_symbols()) == declared, "ctypes signature table out of sync with srb200.h"
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (srb_[a-z0-9_]+)", out))
    assert set(declared) <= exported


def test_host_only_entry_points():
    from srb200 import _capi
    assert _capi.lib().srb_version() >= 100
    # known answers printed by the reference notebooks: 478 -> 490 -> 39^2 patches
    assert _capi.tiling_geometry(478, 478, 24, 12) == (490, 490, 39, 39)
    assert _capi.tiling_geometry(239, 239, 24, 12) == (251, 251, 19, 19)
    with pytest.raises(ValueError):
        _capi.tiling_geometry(0, 10, 24, 12)


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    from srb200 import metrics
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        metrics.psnr(np.zeros((1, 16, 16, 3), np.float32), np.zeros((1, 16, 16, 3), np.float32))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"


def test_ssim_window_kinds_match_the_header():
    """include/srb200.h enum { SRB_SSIM_TF, SRB_SSIM_SKIMAGE, SRB_SSIM_TF_EXACT } <-> srb200._capi constants."""
    import re
    from srb200 import _capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "include", "srb200.h")).read()
    m = re.search(r"enum\s*\{\s*SRB_SSIM_TF\s*=\s*(\d+)\s*,\s*SRB_SSIM_SKIMAGE\s*=\s*(\d+)\s*,\s*SRB_SSIM_TF_EXACT\s*=\s*(\d+)\s*\}", text)
    assert m, "SSIM window enum not found in include/srb200.h"
    assert tuple(int(v) for v in m.groups()) == (_capi.SSIM_TF, _capi.SSIM_SKIMAGE, _capi.SSIM_TF_EXACT)
