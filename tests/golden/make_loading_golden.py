"""Golden vectors for srb200.loading_methods, produced by the REFERENCE's own module
(/root/reference/SRModels/loading_methods.py, imported unmodified) on a small synthetic dataset.
Run in the build container only:

    python tests/golden/make_loading_golden.py

The fixture stores the synthetic images themselves (uint8 RGB), so the tests rebuild the identical PNG files.
"""
import os
import pickle
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")

import cv2  # noqa: E402
from SRModels import loading_methods as ref  # noqa: E402


def synth_images():
    rng = np.random.default_rng(42)
    imgs = {}
    for k, (h, w) in enumerate(((40, 44), (37, 52), (48, 40))):
        hr = cv2.GaussianBlur(rng.integers(0, 256, (h, w, 3), dtype=np.uint8), (0, 0), 1.2)
        lr = cv2.resize(hr, (w // 2, h // 2), interpolation=cv2.INTER_AREA)
        imgs[f"img_{k}.png"] = (hr, lr)
    return imgs


def write_dataset(root, imgs):
    hr_root, lr_root = os.path.join(root, "hr", "cls_a"), os.path.join(root, "lr", "cls_a")
    os.makedirs(hr_root), os.makedirs(lr_root)
    for name, (hr, lr) in imgs.items():
        cv2.imwrite(os.path.join(hr_root, name), cv2.cvtColor(hr, cv2.COLOR_RGB2BGR))
        cv2.imwrite(os.path.join(lr_root, name), cv2.cvtColor(lr, cv2.COLOR_RGB2BGR))
    maps = {"interp": {n: ("INTER_CUBIC" if i != 1 else "INTER_LINEAR") for i, n in enumerate(imgs)},
            "labels": {n: i % 2 for i, n in enumerate(imgs)}}
    for k, v in maps.items():
        with open(os.path.join(root, k + ".pkl"), "wb") as f:
            pickle.dump(v, f)
    return os.path.join(root, "hr"), os.path.join(root, "lr"), os.path.join(root, "interp.pkl"), os.path.join(root, "labels.pkl")


def main():
    imgs = synth_images()
    out = {}
    for name, (hr, lr) in imgs.items():
        out["hr_" + name], out["lr_" + name] = hr, lr
    with tempfile.TemporaryDirectory() as root:
        hr_root, lr_root, interp, labels = write_dataset(root, imgs)
        X, Y, h, w = ref.load_dataset_as_patches(hr_root, lr_root, mode="srcnn", patch_size=12, stride=6, interpolation_map_path=interp)
        out["srcnn_X"], out["srcnn_Y"], out["srcnn_hw"] = X, Y, np.array([h, w])
        X, Y = ref.load_dataset_as_patches(hr_root, lr_root, mode="scale", patch_size=8, stride=4, scale_factor=2)
        out["scale_X"], out["scale_Y"] = X, Y
        X, y = ref.load_defects_dataset_as_patches(hr_root, patch_size=16, stride=8, class_map_path=labels)
        out["defects_X"], out["defects_y"] = X, y
        out["pad_in"] = imgs["img_1.png"][0].astype(np.float32) / 255.0
        out["pad_out"] = ref.add_padding(out["pad_in"], 12, 6)
    np.savez_compressed(os.path.join(HERE, "loading_ref.npz"), **{k: v for k, v in out.items() if v is not None})
    print({k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
