"""cv2.resize golden outputs on uint8 images for the bilinear, (up-scaling) area and Lanczos-4 modes (OpenCV 4.13.0 of the
build image): what the reference's interpolate_bilinear / interpolate_area / interpolate_lanczos
(classic_algorithms.py:7-21) return when super_resolucion_clasica.ipynb cell 7 calls them on uint8 arrays.

    python tests/golden/make_resize_u8_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
rng = np.random.default_rng(12)
out = {}
cases = [(17, 23, 34, 46, 3), (17, 23, 51, 69, 3), (16, 12, 64, 48, 3), (19, 31, 40, 77, 3), (25, 18, 25, 18, 3),
         (21, 14, 63, 28, 1)]
for n, (h, w, dh, dw, c) in enumerate(cases):
    u = rng.integers(0, 256, (h, w, c), dtype=np.uint8)
    if c == 1:
        u = u[:, :, 0]
    out[f"c{n}_shape"] = np.array([h, w, dh, dw, c])
    out[f"c{n}_in"] = u
    out[f"c{n}_linear"] = cv2.resize(u, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out[f"c{n}_area"] = cv2.resize(u, (dw, dh), interpolation=cv2.INTER_AREA)
    out[f"c{n}_lanczos4"] = cv2.resize(u, (dw, dh), interpolation=cv2.INTER_LANCZOS4)
out["n_cases"] = np.array(len(cases))
np.savez_compressed(os.path.join(HERE, "resize_u8_cv2.npz"), **out)
print("wrote", len(cases), "cases")
