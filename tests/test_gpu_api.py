"""GPU: the reference-facing wrappers (SRCNNModel / EDSR / ESRGAN / FineTunedVGG16) end to end
against the oracle pipeline (bicubic -> tiling -> network -> overlap-add; evaluate means)."""
import numpy as np
import pytest
import torch

from oracle import bicubic as ob, convnets as oc, metrics as om, tiling as ot

pytestmark = pytest.mark.gpu


def _oracle_tiled(forward, img, patch, stride, scale):
    padded = ot.add_padding(img, patch, stride)
    patches, pos = ot.extract_patches(padded, patch, stride)
    preds = forward(patches)
    return ot.reconstruct(preds, pos, padded.shape, (img.shape[0] * scale, img.shape[1] * scale), scale)


def test_srcnn_super_resolve_image_config1_flow():
    from srb200 import synth, weights
    from srb200.deep_learning_models.SRCNN_model import SRCNNModel
    hr = synth.hr_image(96, 80, 0)
    lr = synth.area_downsample(hr, 2)
    w = weights.srcnn_weights(bias_scale=0.05)
    m = SRCNNModel()
    m.setup_model(input_shape=(33, 33, 3))
    with pytest.raises(RuntimeError):
        m.super_resolve_image(lr, 96, 80)
    m.load_weights(w)
    sr, info = m.super_resolve_image(lr, 96, 80, patch_size=33, stride=14)
    up = ob.cv2_resize(lr, (80, 96))
    want = _oracle_tiled(lambda p: oc.srcnn_forward(w, p), up, 33, 14, 1)
    assert sr.shape == (96, 80, 3) and sr.dtype == np.float32
    assert np.abs(sr - want).max() <= 1e-3
    assert set(info) == {"time_sec", "gpu_mean_current_mb", "gpu_peak_mb"}
    # the `interpolation` argument (SRCNN_model.py:111, :191): Lanczos-4 pre-upsampling through the same flow
    sr4, _ = m.super_resolve_image(lr, 96, 80, patch_size=33, stride=14, interpolation=4)     # cv2.INTER_LANCZOS4
    want4 = _oracle_tiled(lambda p: oc.srcnn_forward(w, p), ob.resize_lanczos4_f32(lr, (80, 96)), 33, 14, 1)
    assert np.abs(sr4 - want4).max() <= 1e-3
    with pytest.raises(NotImplementedError):
        m.super_resolve_image(lr, 96, 80, interpolation=0)                                    # cv2.INTER_NEAREST


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("fp16", 2e-2)])
def test_edsr_super_resolve_and_evaluate(precision, tol):
    from srb200 import synth, weights
    from srb200.deep_learning_models.EDSR_model import EDSR
    w = weights.edsr_weights(2, num_res_blocks=3, bias_scale=0.05)
    m = EDSR()
    m.setup_model(scale_factor=2, num_res_blocks=3, precision=precision)
    m.load_weights(w)
    hr = synth.hr_batch(6, 48, 48)
    lr = synth.area_downsample(hr, 2)
    fwd = lambda p: oc.edsr_forward(w, p, 2, 3)
    sr, _ = m.super_resolve_image(lr[0], patch_size_lr=24, stride=12)
    want = _oracle_tiled(fwd, lr[0], 24, 12, 2)
    assert sr.shape == (48, 48, 3) and np.abs(sr - want).max() <= tol
    loss, psnr, ssim = m.evaluate(lr, hr)
    ref = om.evaluate_means(hr, fwd(lr))
    print(f"evaluate {precision}: |dloss| {abs(loss - ref[0]):.2e}, |dPSNR| {abs(psnr - ref[1]):.2e} dB, |dSSIM| {abs(ssim - ref[2]):.2e}")
    # BASELINE.json's tolerances in both modes: PSNR within 0.01 dB, SSIM within 1e-4
    assert abs(loss - ref[0]) <= 1e-4 and abs(psnr - ref[1]) <= 0.01 and abs(ssim - ref[2]) <= 1e-4


def test_esrgan_super_resolve_image():
    from srb200 import synth, weights
    from srb200.deep_learning_models.ESRGAN_model import ESRGAN
    w = weights.esrgan_generator_weights(2, 8, 1, bias_scale=0.05)
    m = ESRGAN()
    m.setup_model(scale_factor=2, growth_channels=8, num_rrdb_blocks=1)
    m.load_weights(w)
    lr = synth.area_downsample(synth.hr_image(48, 48, 3), 2)
    sr, _ = m.super_resolve_image(lr, patch_size_lr=24, stride=12)
    fwd = lambda p: (oc.esrgan_generator_forward(w, p * 2 - 1, 2, 1) + 1) / 2
    want = _oracle_tiled(fwd, lr, 24, 12, 2)
    assert np.abs(sr - want).max() <= 1e-3


def test_vgg16_classify_defects_method():
    from srb200 import synth, weights
    from srb200.defect_detection_models.VGG16_model import FineTunedVGG16, vote
    w = weights.vgg16_classifier_weights(2, bias_scale=0.05)
    m = FineTunedVGG16()
    m.setup_model(input_shape=(32, 32, 3))
    m.load_weights(w)
    img = synth.hr_image(70, 50, 1)
    cls, conf = m.classify_defects_method(img)
    padded = ot.add_padding(img, 32, 16)
    patches, _ = ot.extract_patches(padded, 32, 16)
    want = vote(oc.vgg16_classifier_forward(w, patches))
    assert cls == want[0] and abs(conf - want[1]) <= 2e-3


def test_whole_image_fast_path():
    """super_resolve_image_whole: one fully-convolutional pass (no tiling) against the oracle network on the whole image,
    and agreement with the tiled flow away from patch borders."""
    from srb200 import synth, weights
    from srb200.deep_learning_models.EDSR_model import EDSR
    from srb200.deep_learning_models.SRCNN_model import SRCNNModel
    hr = synth.hr_image(120, 88, 3)
    lr = synth.area_downsample(hr, 2)
    w = weights.edsr_weights(2, num_res_blocks=2, bias_scale=0.05)
    m = EDSR()
    m.setup_model(scale_factor=2, num_res_blocks=2, precision="fp32")
    with pytest.raises(RuntimeError):
        m.super_resolve_image_whole(lr)
    m.load_weights(w)
    sr, info = m.super_resolve_image_whole(lr)
    want = np.clip(oc.edsr_forward(w, lr[None], 2, 2)[0], 0, 1)
    assert sr.shape == (120, 88, 3) and sr.dtype == np.float32 and np.abs(sr - want).max() <= 1e-3
    assert set(info) == {"time_sec", "gpu_mean_current_mb", "gpu_peak_mb"}
    tiled, _ = m.super_resolve_image(lr, patch_size_lr=24, stride=12)
    # receptive field: head + 2 blocks + body-end + up-conv = 7 LR pixels (+ the HR tail).  LR pixels 8..10 are covered by
    # the first patch only (the next one starts at 12) and see no patch border, so both flows compute the same function
    assert np.abs(sr[16:22, 16:22] - tiled[16:22, 16:22]).max() <= 1e-3
    ws = weights.srcnn_weights(bias_scale=0.05)
    s = SRCNNModel()
    s.setup_model(input_shape=(33, 33, 3))
    s.load_weights(ws)
    out, _ = s.super_resolve_image_whole(lr, 120, 88)
    want = np.clip(oc.srcnn_forward(ws, ob.cv2_resize(lr, (88, 120))[None])[0], 0, 1)
    assert out.shape == (120, 88, 3) and np.abs(out - want).max() <= 1e-3
