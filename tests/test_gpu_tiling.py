"""GPU parity: pad_extract / overlap_add kernels vs the reference's add_padding golden outputs and
the explicit-loop oracle (bit-exact: pure data movement plus one division)."""
import os

import numpy as np
import pytest

from oracle import tiling as ot

pytestmark = pytest.mark.gpu


def test_pad_extract_vs_reference_golden(golden_dir):
    import torch
    from srb200 import ops
    g = np.load(os.path.join(golden_dir, "tiling_ref.npz"))
    for n in range(int(g["n_cases"])):
        h, w, p, s = (int(v) for v in g[f"c{n}_params"])
        padded = g[f"c{n}_padded"]                       # produced by the reference's own add_padding
        want, pos = ot.extract_patches(padded, p, s)
        got, (ph, pw, ny, nx) = ops.pad_extract(torch.from_numpy(g[f"c{n}_in"]).cuda(), p, s)
        assert (ph, pw) == padded.shape[:2] and ny * nx == len(pos)
        assert np.array_equal(got.cpu().numpy(), want), n


@pytest.mark.parametrize("h,w,p,s,scale", [(30, 41, 24, 12, 1), (25, 25, 33, 14, 1), (50, 37, 48, 24, 2),
                                           (29, 31, 24, 12, 4), (13, 29, 24, 12, 3)])
def test_overlap_add_vs_oracle(h, w, p, s, scale):
    import torch
    from srb200 import ops
    rng = np.random.default_rng(h * w)
    img = rng.random((h, w, 3), dtype=np.float32)
    padded = ot.add_padding(img, p, s)
    patches, pos = ot.extract_patches(padded, p, s)
    preds = rng.random((len(pos), p * scale, p * scale, 3), dtype=np.float32) * 1.4 - 0.2
    want = ot.reconstruct(preds, pos, padded.shape, (h * scale, w * scale), scale)
    ny = (padded.shape[0] - p) // s + 1
    nx = (padded.shape[1] - p) // s + 1
    got = ops.overlap_add(torch.from_numpy(preds).cuda(), ny, nx, s * scale, h * scale, w * scale).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-6


def test_round_trip_identity():
    """extract -> overlap-average of the same patches reproduces the image (any size)."""
    import torch
    from srb200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand((478, 478, 3), device="cuda", generator=g)
    patches, (ph, pw, ny, nx) = ops.pad_extract(img, 24, 12)
    assert (ph, ny * nx) == (490, 1521)                  # SRCNN.ipynb dataset-shape known answer
    back = ops.overlap_add(patches, ny, nx, 12, 478, 478)
    assert (back - img).abs().max() <= 1e-6
