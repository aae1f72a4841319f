"""cv2.resize golden outputs for the bilinear, (up-scaling) area and Lanczos-4 modes (OpenCV 4.13.0 of the build image) -
the reference's interpolate_bilinear / interpolate_area / interpolate_lanczos (classic_algorithms.py:7-21).

    python tests/golden/make_resize_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(11)
out = {}
cases = [(17, 23, 34, 46), (17, 23, 51, 69), (16, 12, 64, 48), (19, 31, 40, 77), (12, 12, 12, 30), (25, 18, 25, 18)]
for n, (h, w, dh, dw) in enumerate(cases):
    f = rng.random((h, w, 3), dtype=np.float32)
    out[f"c{n}_shape"] = np.array([h, w, dh, dw])
    out[f"c{n}_in"] = f
    out[f"c{n}_linear"] = cv2.resize(f, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out[f"c{n}_area"] = cv2.resize(f, (dw, dh), interpolation=cv2.INTER_AREA)
    out[f"c{n}_lanczos4"] = cv2.resize(f, (dw, dh), interpolation=cv2.INTER_LANCZOS4)
out["n_cases"] = np.array(len(cases))
np.savez_compressed(os.path.join(HERE, "resize_cv2.npz"), **out)
print("wrote", len(cases), "cases")
