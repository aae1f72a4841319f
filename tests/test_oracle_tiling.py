"""Pin the tiling oracle against the reference's own add_padding (fixtures) and notebook KATs."""
import os

import numpy as np

from oracle import tiling as ot


def test_add_padding_matches_reference_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiling_ref.npz"))
    for n in range(int(g["n_cases"])):
        h, w, p, s = g[f"c{n}_params"]
        got = ot.add_padding(g[f"c{n}_in"], int(p), int(s))
        assert got.shape == g[f"c{n}_padded"].shape, n
        assert np.array_equal(got, g[f"c{n}_padded"]), n


def test_notebook_patch_counts(golden_dir):
    g = np.load(os.path.join(golden_dir, "tiling_ref.npz"))
    for size, key in ((478, "kat_478_24_12"), (239, "kat_239_24_12")):
        ph, pw = ot.pad_amounts(size, size, 24, 12)
        assert size + ph == g[key][0]
        assert len(ot.window_positions(size + ph, size + pw, 24, 12)) == g[key][1]
    # SRCNN.ipynb:L47: 333,251 = floor(0.7 * 313 * 39^2)
    assert int(0.7 * 313 * 39 * 39) == 333251


def test_extract_reconstruct_round_trip():
    rng = np.random.default_rng(0)
    img = rng.random((50, 37, 3), dtype=np.float32)
    padded = ot.add_padding(img, 24, 12)
    patches, pos = ot.extract_patches(padded, 24, 12)
    rec = ot.reconstruct(patches, pos, padded.shape, (50, 37), scale=1)
    assert np.abs(rec - img).max() < 1e-6
    up = np.repeat(np.repeat(patches, 2, axis=1), 2, axis=2)
    rec2 = ot.reconstruct(up, pos, padded.shape, (100, 74), scale=2)
    assert np.abs(rec2 - np.repeat(np.repeat(img, 2, 0), 2, 1)).max() < 1e-6
