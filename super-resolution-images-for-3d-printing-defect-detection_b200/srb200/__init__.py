"""srb200 - B200-native super-resolution inference path.

Drop-in for the hot path of the reference's ``SRModels`` package: same module, class and function
names (``metrics.psnr/ssim``, ``classic_super_resolution_algorithms.classic_algorithms.
interpolate_bicubic``, ``loading_methods.add_padding``, ``deep_learning_models.SRCNN_model.SRCNNModel``,
``deep_learning_models.EDSR_model.EDSR`` ...), with every numerical step executed by hand-written
sm_100a CUDA kernels behind the C ABI of ``include/srb200.h``.  No CPU fallback.
"""
__version__ = "0.1.0"
