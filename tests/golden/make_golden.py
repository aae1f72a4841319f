"""Generate the golden fixtures under tests/golden/ (run in the BUILD container only).

Sources of truth, in order of strength:
* ``cv2.resize(..., INTER_CUBIC)``            - the reference's real bicubic (OpenCV 4.13.0).
* ``SRModels.loading_methods.add_padding``    - imported from /root/reference, unmodified.
* the numpy/torch oracle in ``oracle/``       - for the paths whose engine (TensorFlow 2.10)
  cannot be installed here; these vectors are regression anchors, not reference outputs.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200"))

import cv2  # noqa: E402
import torch  # noqa: E402

from oracle import bicubic as ob, convnets as oc, metrics as om  # noqa: E402
from srb200 import synth, weights  # noqa: E402


def bicubic_vectors():
    rng = np.random.default_rng(7)
    out = {}
    cases = [(17, 23, 34, 46), (17, 23, 51, 69), (16, 12, 64, 48), (19, 31, 40, 77), (12, 12, 12, 30)]
    for n, (h, w, dh, dw) in enumerate(cases):
        f = rng.random((h, w, 3), dtype=np.float32)
        u = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out[f"c{n}_shape"] = np.array([h, w, dh, dw])
        out[f"c{n}_f32_in"], out[f"c{n}_u8_in"] = f, u
        for tag, opt in (("default", True), ("scalar", False)):
            out[f"c{n}_f32_{tag}"] = ob.cv2_resize(f, (dw, dh), optimized=opt)
            out[f"c{n}_u8_{tag}"] = ob.cv2_resize(u, (dw, dh), optimized=opt)
    out["n_cases"] = np.array(len(cases))
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(HERE, "bicubic_cv2.npz"), **out)


def tiling_vectors():
    sys.path.insert(0, "/root/reference")
    from SRModels.loading_methods import add_padding  # the reference's own function
    rng = np.random.default_rng(11)
    out = {}
    cases = [(30, 41, 24, 12), (48, 48, 24, 12), (25, 25, 33, 14), (50, 37, 48, 24), (96, 100, 96, 48),
             (13, 29, 24, 12)]
    for n, (h, w, p, s) in enumerate(cases):
        img = rng.random((h, w, 3), dtype=np.float32)
        out[f"c{n}_params"] = np.array([h, w, p, s])
        out[f"c{n}_in"] = img
        out[f"c{n}_padded"] = add_padding(img, p, s)
    out["n_cases"] = np.array(len(cases))
    # dataset-shape known answers printed by the reference notebooks
    out["kat_478_24_12"] = np.array([490, 39 * 39])
    out["kat_239_24_12"] = np.array([251, 19 * 19])
    np.savez_compressed(os.path.join(HERE, "tiling_ref.npz"), **out)


def metrics_vectors():
    hr = synth.hr_batch(4, 40, 56)
    rng = np.random.default_rng(3)
    pred = np.clip(hr + 0.05 * rng.standard_normal(hr.shape).astype(np.float32), 0, 1)
    np.savez_compressed(os.path.join(HERE, "metrics_oracle.npz"), a=hr, b=pred,
                        psnr32=om.psnr(hr, pred), ssim32=om.ssim(hr, pred),
                        psnr64=om.psnr(hr, pred, dtype=np.float64),
                        ssim64=om.ssim(hr, pred, dtype=np.float64))


def convnet_vectors():
    lr = synth.area_downsample(synth.hr_batch(2, 48, 48), 2)[:, :20, :24]      # 2 x 20 x 24 x 3
    out = {"lr": lr}
    w = weights.srcnn_weights(bias_scale=0.1)
    out["srcnn"] = oc.srcnn_forward(w, lr, dtype=torch.float64).astype(np.float32)
    w = weights.edsr_weights(scale_factor=4, num_res_blocks=2, bias_scale=0.1)
    out["edsr_x4_2blocks"] = oc.edsr_forward(w, lr, 4, 2, dtype=torch.float64).astype(np.float32)
    w = weights.edsr_weights(scale_factor=3, num_res_blocks=1, bias_scale=0.1)
    out["edsr_x3_1block"] = oc.edsr_forward(w, lr, 3, 1, dtype=torch.float64).astype(np.float32)
    w = weights.espcn_weights(bias_scale=0.1)
    out["espcn_x4"] = oc.espcn_forward(w, lr, 4, dtype=torch.float64).astype(np.float32)
    w = weights.srresnet_weights(num_res_blocks=2, bias_scale=0.1)
    out["srresnet_x4_2blocks"] = oc.srresnet_forward(w, lr, 4, 2, dtype=torch.float64).astype(np.float32)
    w = weights.esrgan_generator_weights(2, 8, 1, bias_scale=0.1)
    out["esrgan_x2_1rrdb_g8"] = oc.esrgan_generator_forward(w, lr * 2 - 1, 2, 1,
                                                          dtype=torch.float64).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "convnets_oracle.npz"), **out)


if __name__ == "__main__":
    bicubic_vectors()
    tiling_vectors()
    metrics_vectors()
    convnet_vectors()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
