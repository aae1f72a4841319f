"""Tiling and seeding constants under the reference's names (SRModels/constants.py:1-14), so that notebooks which do
``from SRModels.constants import EDSR_PATCH_SIZE`` keep working after switching the import to ``srb200.constants``.

The values are kept in one table per model family - (LR patch size, stride[, scale factor]) - and exported as the flat
``<MODEL>_PATCH_SIZE`` / ``<MODEL>_STRIDE`` / ``<MODEL>_SCALE_FACTOR`` names the reference uses.  Every patch size is
twice its stride: adjacent windows overlap by half, which is what the overlap-add of ``super_resolve_image`` averages.
"""
from __future__ import annotations

#: model family -> (patch_size, stride, scale_factor or None)
TILING = {
    "SRCNN": (24, 12, None),      # patches of the bicubic pre-upsampled image (scale 1 inside the network)
    "EDSR": (24, 12, 2),
    "ESRGAN": (24, 12, 2),
    "VGG": (96, 48, None),        # classifier patches cut from the SR output
}

#: seed of every numpy / dataset split in the reference's notebooks; the synthetic inputs here start from it too
RANDOM_SEED = 42

for _family, (_patch, _stride, _scale) in TILING.items():
    globals()[f"{_family}_PATCH_SIZE"] = _patch
    globals()[f"{_family}_STRIDE"] = _stride
    if _scale is not None:
        globals()[f"{_family}_SCALE_FACTOR"] = _scale
    assert _patch == 2 * _stride
del _family, _patch, _stride, _scale

__all__ = sorted(k for k in globals() if k.isupper())
