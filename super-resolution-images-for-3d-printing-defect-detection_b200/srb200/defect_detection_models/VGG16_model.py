"""``FineTunedVGG16`` inference surface (SRModels/defect_detection_models/VGG16_model.py):
``setup_model`` (:21-55), ``build_vgg16`` (:57-97) and the patch-vote ``classify_defects_method``
(:168-270).  ImageNet weights are not available offline, so a fresh model is he_normal-initialised."""
from __future__ import annotations

import numpy as np

from .. import engine, ops, weights as W
from ..deep_learning_models import _common as common


class FineTunedVGG16:
    def __init__(self):
        self.model = None
        self.trained = False
        self.input_shape = None

    def setup_model(self, input_shape=(128, 128, 3), num_classes=2, from_pretrained=False, pretrained_path=None,
                    precision="fp32", seed=1234, **_ignored):
        if num_classes < 2:
            raise ValueError("num_classes must be >= 2")
        self.input_shape = tuple(input_shape)
        if from_pretrained:
            w = common.load_weight_file(pretrained_path)
            self.trained = True
        else:
            w = W.vgg16_classifier_weights(num_classes, seed=seed)
        self.model = engine.VGG16ClassifierNet(w, precision=precision)

    def load_weights(self, weights, precision=None):
        self.model = engine.VGG16ClassifierNet(weights, precision=precision or self.model.precision)
        self.trained = True

    def classify_defects_method(self, image, patch_size=None, stride=None, batch_size=32):
        """Patch-vote classification -> (predicted_class, confidence).  Padding and patch extraction
        run on the device; voting and the tie-break follow VGG16_model.py:252-270."""
        if self.model is None:
            raise ValueError("Model is not built yet.")
        if image is None:
            raise ValueError("image must be provided")
        img = np.asarray(image)
        if img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("image must be HxWx3 RGB array")
        if patch_size is None:
            if self.input_shape is None or self.input_shape[0] is None:
                raise ValueError("Model input size is dynamic; please set patch_size.")
            patch_size = int(self.input_shape[0])
        if stride is None:
            stride = max(1, patch_size // 2)
        patches, _geom = ops.pad_extract(common.as_device_image(img), patch_size, stride)
        probs = self.model.predict_device(patches).cpu().numpy()
        return vote(probs)


def vote(probs):
    """Majority vote over patches; ties -> higher mean probability; confidence = mean prob of winner."""
    probs = np.asarray(probs)
    if probs.ndim != 2:
        probs = probs.reshape((probs.shape[0], -1))
    num_classes = int(probs.shape[1])
    votes = np.bincount(np.argmax(probs, axis=1), minlength=num_classes)
    top = np.where(votes == votes.max())[0]
    if len(top) == 1:
        winner = int(top[0])
    else:
        winner = int(top[np.argmax(probs.mean(axis=0)[top])])
    return winner, float(probs[:, winner].mean())
