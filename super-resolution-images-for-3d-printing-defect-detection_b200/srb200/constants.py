"""Tiling and seeding constants under the reference's names (SRModels/constants.py:1-14), so that notebooks which do
``from SRModels.constants import EDSR_PATCH_SIZE`` keep working after switching the import to ``srb200.constants``.

Every patch size is twice its stride: adjacent windows overlap by half, which is what the overlap-add of
``super_resolve_image`` averages.
"""
# SRCNN works on patches of the bicubic pre-upsampled image (scale 1 inside the network)
SRCNN_PATCH_SIZE = 24
SRCNN_STRIDE = 12

EDSR_SCALE_FACTOR = 2
EDSR_PATCH_SIZE = 24
EDSR_STRIDE = 12

ESRGAN_SCALE_FACTOR = 2
ESRGAN_PATCH_SIZE = 24
ESRGAN_STRIDE = 12

# classifier patches cut from the SR output
VGG_PATCH_SIZE = 96
VGG_STRIDE = 48

# seed of every numpy / dataset split in the reference's notebooks; the synthetic inputs here start from it too
RANDOM_SEED = 42
