"""Dataset loaders with the reference's signatures (SRModels/loading_methods.py:6-26, 28-38, 40-191, 194-285,
288-385), feeding the B200 path.

Image decoding stays on the host (``cv2.imread``, as in the reference).  What the reference then does per image in
numpy / OpenCV - bicubic pre-upsampling of the LR image to the HR size (``loading_methods.py:147-148``), reflect padding
(``:6-26``) and the sliding-window patch loops (``:155-161``) - runs on the GPU here: ``srb_bicubic_f32`` with the
fused ``np.clip`` and ``srb_pad_extract_f32`` (one launch per image instead of a Python double loop).  Return types,
shapes, ordering and exceptions are the reference's.

Documented differences:
* ``load_dataset_as_patches(mode='srcnn', interpolation_map_path=None)`` raises ``NameError`` in the reference because
  ``interpolation_map`` is only bound when a path is given (``:110-113`` vs ``:134``); here a missing map simply means
  bicubic for every image.
* ``INTER_CUBIC``, ``INTER_LINEAR``, ``INTER_LANCZOS4`` and (up-scaling) ``INTER_AREA`` entries of the interpolation map
  (``:133-145``) run on the device; a down-scaling ``INTER_AREA`` is resized by OpenCV on the host, as the reference does.
* ``load_defects_dataset_as_patches`` walks the *unpadded* extent (``:276-277``), so the padding it adds is never
  visited; reproduced as is.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from . import _capi as capi
from . import ops

_IMAGE_SUFFIXES = (".jpg", ".jpeg", ".png", ".bmp", ".tiff")


def add_padding(image, patch_size, stride):
    """Reflect-pad bottom / right so that windows of ``patch_size`` at ``stride`` cover the image
    (loading_methods.py:6-26).  Host numpy, like the reference; the loaders below use the device kernel."""
    h, w = image.shape[:2]
    pad = []
    for n in (h, w):
        extra = (patch_size - n % stride) % stride if n % stride else 0
        pad.append(max(extra, patch_size - stride))
    return np.pad(image, ((0, pad[0]), (0, pad[1]), (0, 0)), mode="reflect")


def get_all_image_paths(root):
    """Sorted paths of every image file under ``root`` (loading_methods.py:28-38)."""
    found = []
    for directory, _, names in os.walk(root):
        found.extend(os.path.join(directory, n) for n in names if n.lower().endswith(_IMAGE_SUFFIXES))
    return sorted(found)


def _cv2():
    import cv2
    return cv2


def _read_rgb01(path):
    cv2 = _cv2()
    bgr = cv2.imread(path, cv2.IMREAD_COLOR)
    if bgr is None:
        raise ValueError(f"Failed to read image: {path}")
    return cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB).astype(np.float32) / 255.0


def _device_patches(img, patch, stride):
    """[H, W, 3] float32 (numpy or CUDA tensor) -> numpy [ny * nx, patch, patch, 3] in the reference's loop order."""
    torch = capi.require_cuda()
    t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32)).cuda()
    patches, _ = ops.pad_extract(t, patch, stride)
    return patches.cpu().numpy()


def _check_dirs_and_sizes(hr_root, lr_root, patch_size, stride):
    if not os.path.exists(hr_root) or not os.path.exists(lr_root):
        raise ValueError("Both HR and LR root directories must exist.")
    if not os.path.isdir(hr_root) or not os.path.isdir(lr_root):
        raise ValueError("Both HR and LR root paths must be directories.")
    if not isinstance(patch_size, int) or patch_size <= 0:
        raise ValueError("patch_size must be positive int.")
    if not isinstance(stride, int) or stride <= 0:
        raise ValueError("stride must be positive int.")


def load_dataset_as_patches(hr_root, lr_root, mode="srcnn", patch_size=33, stride=14, scale_factor=2,
                            interpolation_map_path=None):
    """Aligned LR / HR patch pairs (loading_methods.py:40-191).

    ``mode='srcnn'``: the LR image is first resized to the HR size (bicubic unless the interpolation map says
    otherwise) and clipped to [0, 1]; both patch sets are ``patch_size`` square -> ``(X, Y, hr_h, hr_w)`` with the
    size of the last image.  ``mode='scale'``: ``patch_size`` is the LR patch size, HR patches are
    ``patch_size * scale_factor`` square and cut at ``scale_factor`` times the LR window origin -> ``(X, Y)``."""
    if mode not in ("srcnn", "scale"):
        raise ValueError("mode must be 'srcnn' or 'scale'")
    _check_dirs_and_sizes(hr_root, lr_root, patch_size, stride)
    if mode == "scale" and (not isinstance(scale_factor, int) or scale_factor <= 0):
        raise ValueError("scale_factor must be positive int.")
    hr_by_name = {os.path.basename(p): p for p in get_all_image_paths(hr_root)}
    lr_by_name = {os.path.basename(p): p for p in get_all_image_paths(lr_root)}
    if not hr_by_name or not lr_by_name:
        raise ValueError("No images found in provided directories.")
    interpolation_map = None
    if mode == "srcnn" and interpolation_map_path is not None:
        with open(interpolation_map_path, "rb") as f:
            interpolation_map = pickle.load(f)
    cv2 = _cv2()
    torch = capi.require_cuda()
    by_name = {"INTER_LINEAR": cv2.INTER_LINEAR, "INTER_CUBIC": cv2.INTER_CUBIC, "INTER_AREA": cv2.INTER_AREA,
               "INTER_LANCZOS4": cv2.INTER_LANCZOS4}
    X, Y = [], []
    hr_h = hr_w = None
    for name in sorted(set(hr_by_name) & set(lr_by_name)):
        hr_img, lr_img = _read_rgb01(hr_by_name[name]), _read_rgb01(lr_by_name[name])
        hr_h, hr_w = hr_img.shape[:2]
        if mode == "srcnn":
            code = cv2.INTER_CUBIC
            if interpolation_map is not None:
                chosen = interpolation_map.get(name, cv2.INTER_CUBIC)
                if isinstance(chosen, str):
                    code = by_name.get(chosen, cv2.INTER_CUBIC)
                elif isinstance(chosen, int):
                    code = chosen
            grows = hr_h >= lr_img.shape[0] and hr_w >= lr_img.shape[1]
            if code in (cv2.INTER_CUBIC, cv2.INTER_LINEAR, cv2.INTER_LANCZOS4) or (code == cv2.INTER_AREA and grows):
                lr_dev = torch.from_numpy(lr_img).cuda()[None]
                lr_up = ops.resize(lr_dev, hr_h, hr_w, interpolation=code, clip01=True)[0]   # resize + np.clip in one kernel
            else:
                lr_up = np.clip(cv2.resize(lr_img, (hr_w, hr_h), interpolation=code), 0.0, 1.0)
            Y.append(_device_patches(hr_img, patch_size, stride))
            X.append(_device_patches(lr_up, patch_size, stride))
        else:
            hr_patch = patch_size * scale_factor
            lr_patches = _device_patches(lr_img, patch_size, stride)
            # the HR side is padded with the LR stride and indexed at scale_factor * (i, j) (loading_methods.py:163-183):
            # windows that would leave the padded HR image are dropped together with their LR partner
            hr_pad = add_padding(hr_img, hr_patch, stride)
            lr_ph, lr_pw, ny, nx = capi.tiling_geometry(lr_img.shape[0], lr_img.shape[1], patch_size, stride)
            keep_x, keep_y = [], []
            for a in range(ny):
                for b in range(nx):
                    top, left = a * stride * scale_factor, b * stride * scale_factor
                    win = hr_pad[top:top + hr_patch, left:left + hr_patch, :]
                    if win.shape[:2] == (hr_patch, hr_patch):
                        keep_y.append(win)
                        keep_x.append(lr_patches[a * nx + b])
            X.append(np.array(keep_x, dtype=np.float32).reshape(-1, patch_size, patch_size, 3))
            Y.append(np.array(keep_y, dtype=np.float32).reshape(-1, hr_patch, hr_patch, 3))
    X_arr = np.concatenate(X, 0) if X else np.array(X)
    Y_arr = np.concatenate(Y, 0) if Y else np.array(Y)
    if mode == "srcnn":
        return X_arr, Y_arr, hr_h, hr_w
    return X_arr, Y_arr


def _load_class_map(class_map_path):
    if not class_map_path or not isinstance(class_map_path, str):
        raise ValueError("class_map_path must be a non-empty string.")
    if not os.path.exists(class_map_path):
        raise FileNotFoundError(f"Class labels map not found: {class_map_path}")
    with open(class_map_path, "rb") as f:
        labels = pickle.load(f)
    if not isinstance(labels, dict):
        raise ValueError("class_labels_map pickle must contain a dict of {basename: class_id}.")
    return labels


def load_defects_dataset_as_patches(hr_root, patch_size=33, stride=14, class_map_path=None):
    """HR patches + one class id per patch (loading_methods.py:194-285) -> ``(X float32, y int64)``."""
    if not os.path.exists(hr_root):
        raise ValueError("HR root directory must exist.")
    if not os.path.isdir(hr_root):
        raise ValueError("HR root path must be a directory.")
    if not isinstance(patch_size, int) or patch_size <= 0:
        raise ValueError("patch_size must be positive int.")
    if not isinstance(stride, int) or stride <= 0:
        raise ValueError("stride must be positive int.")
    if not class_map_path or not isinstance(class_map_path, str):
        raise ValueError("class_map_path must be a non-empty string.")
    if not os.path.exists(class_map_path):
        raise FileNotFoundError(f"Class labels map not found: {class_map_path}")
    paths = get_all_image_paths(hr_root)
    if not paths:
        raise ValueError("No images found under HR root directory.")
    labels = _load_class_map(class_map_path)
    X, y = [], []
    for path in sorted(paths, key=os.path.basename):
        img = _read_rgb01(path)
        h, w = img.shape[:2]
        base = os.path.basename(path)
        if base not in labels:
            raise KeyError(f"Missing class id for image basename in class_labels_map: {base}")
        # the reference steps over the ORIGINAL extent (range(0, hr_h - patch + 1, stride)), not the padded one
        ny = len(range(0, h - patch_size + 1, stride))
        nx = len(range(0, w - patch_size + 1, stride))
        if ny <= 0 or nx <= 0:
            continue
        patches = _device_patches(img, patch_size, stride)
        _, _, pny, pnx = capi.tiling_geometry(h, w, patch_size, stride)
        X.append(patches.reshape(pny, pnx, patch_size, patch_size, 3)[:ny, :nx].reshape(-1, patch_size, patch_size, 3))
        y.extend([int(labels[base])] * (ny * nx))
    X_arr = np.concatenate(X, 0).astype(np.float32) if X else np.array(X, dtype=np.float32)
    return X_arr, np.array(y, dtype=np.int64)


def load_predictions_dataset(lr_root: str, hr_root: str, class_map_path: str):
    """Whole LR / HR image pairs matched by basename plus class ids (loading_methods.py:288-385) ->
    ``(X_LR, X_HR, y)``; nothing to accelerate here, the arrays feed ``super_resolve_image`` / ``classify_defects_method``."""
    for root, tag in ((lr_root, "lr_root"), (hr_root, "hr_root")):
        if not root or not isinstance(root, str) or not os.path.exists(root):
            raise ValueError(f"{tag} must be an existing directory path.")
        if not os.path.isdir(root):
            raise ValueError(f"{tag} must be a directory.")
    if not class_map_path or not isinstance(class_map_path, str):
        raise ValueError("class_map_path must be a non-empty string.")
    if not os.path.exists(class_map_path):
        raise FileNotFoundError(f"Class labels map not found: {class_map_path}")
    lr_paths, hr_paths = get_all_image_paths(lr_root), get_all_image_paths(hr_root)
    if not lr_paths:
        raise ValueError("No images found under LR root directory.")
    if not hr_paths:
        raise ValueError("No images found under HR root directory.")
    labels = _load_class_map(class_map_path)
    lr_by_name = {os.path.basename(p): p for p in lr_paths}
    hr_by_name = {os.path.basename(p): p for p in hr_paths}
    common = sorted(set(lr_by_name) & set(hr_by_name))
    if not common:
        raise ValueError("No matching basenames found between LR and HR roots.")
    lows, highs, y = [], [], []
    for base in common:
        lows.append(_read_rgb01(lr_by_name[base]))
        highs.append(_read_rgb01(hr_by_name[base]))
        if base not in labels:
            raise KeyError(f"Missing class id for basename in class_labels_map: {base}")
        y.append(int(labels[base]))
    return np.array(lows, dtype=np.float32), np.array(highs, dtype=np.float32), np.array(y, dtype=np.int64)
