"""srb200.loading_methods against golden vectors produced by the reference's own module
(tests/golden/make_loading_golden.py imports /root/reference/SRModels/loading_methods.py)."""
import os
import pickle

import numpy as np
import pytest


def _dataset(tmp_path, golden_dir):
    import cv2
    g = np.load(os.path.join(golden_dir, "loading_ref.npz"))
    hr_root, lr_root = tmp_path / "hr" / "cls_a", tmp_path / "lr" / "cls_a"
    hr_root.mkdir(parents=True), lr_root.mkdir(parents=True)
    names = sorted(k[3:] for k in g.files if k.startswith("hr_"))
    for n in names:
        cv2.imwrite(str(hr_root / n), cv2.cvtColor(g["hr_" + n], cv2.COLOR_RGB2BGR))
        cv2.imwrite(str(lr_root / n), cv2.cvtColor(g["lr_" + n], cv2.COLOR_RGB2BGR))
    interp = {n: ("INTER_CUBIC" if i != 1 else "INTER_LINEAR") for i, n in enumerate(names)}
    labels = {n: i % 2 for i, n in enumerate(names)}
    for k, v in (("interp", interp), ("labels", labels)):
        with open(tmp_path / (k + ".pkl"), "wb") as f:
            pickle.dump(v, f)
    return g, names, str(tmp_path / "hr"), str(tmp_path / "lr"), str(tmp_path / "interp.pkl"), str(tmp_path / "labels.pkl")


def test_host_side_helpers(tmp_path, golden_dir):
    from srb200 import loading_methods as lm
    g, names, hr_root, lr_root, _, labels = _dataset(tmp_path, golden_dir)
    assert np.array_equal(lm.add_padding(g["pad_in"], 12, 6), g["pad_out"])            # the reference's add_padding
    assert [os.path.basename(p) for p in lm.get_all_image_paths(hr_root)] == names
    with pytest.raises(ValueError):
        lm.load_dataset_as_patches(hr_root, lr_root, mode="bogus")
    with pytest.raises(ValueError):
        lm.load_dataset_as_patches(hr_root, str(tmp_path / "missing"))
    with pytest.raises(ValueError):
        lm.load_dataset_as_patches(hr_root, lr_root, patch_size=0)
    with pytest.raises(FileNotFoundError):
        lm.load_defects_dataset_as_patches(hr_root, class_map_path=str(tmp_path / "nope.pkl"))
    with pytest.raises(ValueError):
        lm.load_defects_dataset_as_patches(hr_root, class_map_path=None)


def test_predictions_dataset_needs_equal_sizes(tmp_path):
    """load_predictions_dataset stacks whole images: equal sizes per split, labels by basename, KeyError if one is missing."""
    import cv2
    from srb200 import loading_methods as lm
    rng = np.random.default_rng(0)
    for split, (h, w) in (("lr", (8, 10)), ("hr", (16, 20))):
        (tmp_path / split).mkdir()
        for n in ("a.png", "b.png"):
            cv2.imwrite(str(tmp_path / split / n), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
    with open(tmp_path / "labels.pkl", "wb") as f:
        pickle.dump({"a.png": 1, "b.png": 0}, f)
    lo, hi, y = lm.load_predictions_dataset(str(tmp_path / "lr"), str(tmp_path / "hr"), str(tmp_path / "labels.pkl"))
    assert lo.shape == (2, 8, 10, 3) and hi.shape == (2, 16, 20, 3) and lo.dtype == np.float32 and y.tolist() == [1, 0]
    assert 0.0 <= lo.min() and hi.max() <= 1.0
    with open(tmp_path / "labels.pkl", "wb") as f:
        pickle.dump({"a.png": 1}, f)
    with pytest.raises(KeyError):
        lm.load_predictions_dataset(str(tmp_path / "lr"), str(tmp_path / "hr"), str(tmp_path / "labels.pkl"))


@pytest.mark.gpu
def test_patch_loaders_match_reference(tmp_path, golden_dir):
    from srb200 import loading_methods as lm
    g, names, hr_root, lr_root, interp, labels = _dataset(tmp_path, golden_dir)
    X, Y, h, w = lm.load_dataset_as_patches(hr_root, lr_root, mode="srcnn", patch_size=12, stride=6, interpolation_map_path=interp)
    assert (h, w) == tuple(g["srcnn_hw"]) and X.shape == g["srcnn_X"].shape and X.dtype == np.float32
    assert np.array_equal(Y, g["srcnn_Y"])                                             # HR side: pure indexing
    assert np.abs(X - g["srcnn_X"]).max() <= 2e-6                                      # LR side: device bicubic vs cv2
    X2, Y2, _, _ = lm.load_dataset_as_patches(hr_root, lr_root, mode="srcnn", patch_size=12, stride=6)   # no map: all bicubic
    assert X2.shape == X.shape and np.array_equal(Y2, Y)
    X, Y = lm.load_dataset_as_patches(hr_root, lr_root, mode="scale", patch_size=8, stride=4, scale_factor=2)
    assert np.array_equal(X, g["scale_X"]) and np.array_equal(Y, g["scale_Y"])
    X, y = lm.load_defects_dataset_as_patches(hr_root, patch_size=16, stride=8, class_map_path=labels)
    assert np.array_equal(X, g["defects_X"]) and np.array_equal(y, g["defects_y"]) and y.dtype == np.int64
