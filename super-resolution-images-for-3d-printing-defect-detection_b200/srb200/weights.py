"""Keras-layout weight sets for the SR networks on the hot path.

Every network is described by an ordered ``dict`` ``name -> np.ndarray``:
``<layer>/kernel`` is Keras HWIO ``[kh, kw, Cin, Cout]`` float32 and
``<layer>/bias`` is ``[Cout]`` float32.  The same dict is handed to the CPU
oracle and to the CUDA engine, which is what "identical random-init weights"
means in the parity tests.

Initialisers follow the reference's layer declarations:

* SRCNN, ESRGAN generator, VGG16 head: Keras default ``glorot_uniform``
  (/root/reference/SRModels/deep_learning_models/SRCNN_model.py:50-52,
  ESRGAN_model.py:230-246).
* EDSR: ``kernel_initializer="he_normal"`` (EDSR_model.py:61,65,80-89,102,112,121),
  i.e. a truncated normal (|z| <= 2) with stddev sqrt(2/fan_in)/0.87962566.
* ESPCN / SRResNet are not in the reference; they are
  composed from the same Conv2D + depth_to_space semantics with glorot_uniform.

Biases are zeros by default (Keras) or U(-0.1, 0.1) with ``bias_scale=0.1`` to
exercise the fused epilogues.
"""
from __future__ import annotations

import re

import numpy as np

_TRUNC_STD = 0.87962566103423978  # std of a unit normal truncated to [-2, 2]


def glorot_uniform(rng: np.random.Generator, kh, kw, cin, cout):
    fan_in, fan_out = kh * kw * cin, kh * kw * cout
    limit = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-limit, limit, size=(kh, kw, cin, cout)).astype(np.float32)


def he_normal(rng: np.random.Generator, kh, kw, cin, cout):
    fan_in = kh * kw * cin
    std = np.sqrt(2.0 / fan_in) / _TRUNC_STD
    n = kh * kw * cin * cout
    out = rng.standard_normal(n)
    bad = np.abs(out) > 2.0
    while bad.any():  # resample the tails, as Keras' truncated normal does
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).reshape(kh, kw, cin, cout).astype(np.float32)


def _shape_only(_rng, kh, kw, cin, cout):
    """Initialiser of the shape specs: a zero-stride placeholder that only carries ``.shape``."""
    return np.broadcast_to(np.float32(0), (kh, kw, cin, cout))


class _Builder:
    def __init__(self, seed, init, bias_scale, shapes_only=False):
        self.seed, self.init, self.bias_scale = seed, (_shape_only if shapes_only else init), bias_scale
        self.shapes_only = shapes_only
        self.w = {}
        self.idx = 0

    def conv(self, name, k, cin, cout):
        rng = np.random.default_rng(self.seed + self.idx)
        self.idx += 1
        kh, kw = (k, k) if isinstance(k, int) else k
        self.w[name + "/kernel"] = self.init(rng, kh, kw, cin, cout)
        if self.shapes_only:
            self.w[name + "/bias"] = np.broadcast_to(np.float32(0), (cout,))
            return
        if self.bias_scale:
            b = rng.uniform(-self.bias_scale, self.bias_scale, size=cout)
        else:
            b = np.zeros(cout)
        self.w[name + "/bias"] = b.astype(np.float32)

    def vec(self, name, n, value):
        self.w[name] = np.full(n, value, dtype=np.float32)

    def dense(self, name, cin, cout):
        self.conv(name, 1, cin, cout)
        self.w[name + "/kernel"] = self.w[name + "/kernel"].reshape(cin, cout)


def srcnn_weights(seed=1234, bias_scale=0.0, channels=3, shapes_only=False):
    """9-1-5 SRCNN with 96/32/3 filters (SRCNN_model.py:45-53); 28,931 params."""
    b = _Builder(seed, glorot_uniform, bias_scale, shapes_only)
    b.conv("conv1", 9, channels, 96)
    b.conv("conv2", 1, 96, 32)
    b.conv("conv3", 5, 32, channels)
    return b.w


def edsr_weights(scale_factor=2, channels=3, num_res_blocks=16, num_filters=64,
                 seed=1234, bias_scale=0.0, shapes_only=False):
    """EDSR (EDSR_model.py:96-125). x2: 1,369,859 params; x4: 1,517,571."""
    if scale_factor not in (2, 3, 4):
        raise ValueError(f"Scale factor {scale_factor} not supported. Use 2, 3, or 4.")
    b = _Builder(seed, he_normal, bias_scale, shapes_only)
    f = num_filters
    b.conv("head", 3, channels, f)
    for i in range(num_res_blocks):
        b.conv(f"rb{i}_c1", 3, f, f)
        b.conv(f"rb{i}_c2", 3, f, f)
    b.conv("body", 3, f, f)
    if scale_factor == 2:
        b.conv("up0", 3, f, f * 4)
    elif scale_factor == 3:
        b.conv("up0", 3, f, f * 9)
    else:
        b.conv("up0", 3, f, f * 4)
        b.conv("up1", 3, f, f * 4)
    b.conv("tail", 3, f, channels)
    return b.w


def espcn_weights(scale_factor=4, channels=3, seed=1234, bias_scale=0.0, shapes_only=False):
    """ESPCN 5-3-3 with 64/32/(C*r^2) filters (Shi et al. 2016)."""
    b = _Builder(seed, glorot_uniform, bias_scale, shapes_only)
    b.conv("conv1", 5, channels, 64)
    b.conv("conv2", 3, 64, 32)
    b.conv("conv3", 3, 32, channels * scale_factor * scale_factor)
    return b.w


def srresnet_weights(scale_factor=4, channels=3, num_res_blocks=16, num_filters=64,
                     seed=1234, bias_scale=0.0, prelu_slope=0.25, shapes_only=False):
    """SRResNet / SRGAN generator with BatchNorm folded away (Ledig et al. 2017; row A14)."""
    if scale_factor not in (2, 4):
        raise ValueError("SRResNet scale factor must be 2 or 4")
    b = _Builder(seed, glorot_uniform, bias_scale, shapes_only)
    f = num_filters
    b.conv("head", 9, channels, f)
    b.vec("head/prelu", f, prelu_slope)
    for i in range(num_res_blocks):
        b.conv(f"rb{i}_c1", 3, f, f)
        b.vec(f"rb{i}_c1/prelu", f, prelu_slope)
        b.conv(f"rb{i}_c2", 3, f, f)
    b.conv("body", 3, f, f)
    for i in range(2 if scale_factor == 4 else 1):
        b.conv(f"up{i}", 3, f, f * 4)
        b.vec(f"up{i}/prelu", f, prelu_slope)
    b.conv("tail", 9, f, channels)
    return b.w


def esrgan_generator_weights(scale_factor=2, growth_channels=32, num_rrdb_blocks=23,
                             channels=3, seed=1234, bias_scale=0.0, shapes_only=False):
    """RRDBNet generator with two SelfAttention layers (ESRGAN_model.py:30-79, 212-345).

    4 RRDB / growth 8 / x2 gives 1,162,915 parameters (ESRGAN.ipynb:L636)."""
    b = _Builder(seed, glorot_uniform, bias_scale, shapes_only)
    g = growth_channels
    b.conv("initial_conv", 3, channels, 64)
    for i in range(num_rrdb_blocks):
        for d in (1, 2, 3):
            n = f"rrdb_{i}_dense{d}"
            for j in range(4):
                b.conv(f"{n}_conv{j + 1}", 3, 64 + j * g, g)
            b.conv(f"{n}_conv5", 3, 64 + 4 * g, 64)
    b.conv("trunk_conv", 3, 64, 64)

    def attn(name):
        b.conv(name + "_f", 1, 64, 8)
        b.conv(name + "_g", 1, 64, 8)
        b.conv(name + "_h", 1, 64, 32)
        b.conv(name + "_v", 1, 32, 64)

    attn("self_attention_trunk")
    for i in range(int(np.log2(scale_factor))):
        b.conv(f"upsample_{i}_conv", 3, 64, 256)
        if i == 0:
            attn("self_attention_upsample_0")
    b.conv("final_conv1", 3, 64, 64)
    b.conv("final_conv2", 3, 64, channels)
    return b.w


VGG16_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, "M",
             512, 512, 512, "M", 512, 512, 512, "M"]


def vgg16_classifier_weights(num_classes=2, seed=1234, bias_scale=0.0, shapes_only=False):
    """VGG16 conv base + GAP + Dense(256, relu) + Dense(C, softmax) (VGG16_model.py:57-97).

    ImageNet weights are not available offline, so the base is random-init (he_normal keeps
    activations alive through 13 ReLU layers). 14,846,530 params at 2 classes."""
    b = _Builder(seed, he_normal, bias_scale, shapes_only)
    cin, blk, j = 3, 1, 1
    for v in VGG16_CFG:
        if v == "M":
            blk, j = blk + 1, 1
            continue
        b.conv(f"block{blk}_conv{j}", 3, cin, v)
        cin, j = v, j + 1
    if not shapes_only:
        b.init = glorot_uniform
    b.dense("dense", 512, 256)
    b.dense("predictions", 256, num_classes)
    return b.w


# ---------------------------------------------------------------------------------------------
# adopting weight files written from the reference's Keras models
# ---------------------------------------------------------------------------------------------
_AUTO_CONV = re.compile(r"^conv2d(?:_(\d+))?$")


def normalize_keras_names(raw):
    """``{v.name: v.numpy() for v in model.weights}`` keys (``conv2d_3/kernel:0``,
    ``self_attention_trunk/self_attention_trunk_f/bias:0``) -> ``<layer>/<variable>``."""
    out = {}
    for k, v in raw.items():
        k = str(k).split(":")[0]
        parts = k.split("/")
        if len(parts) >= 2:
            k = parts[-2] + "/" + parts[-1]
        out[k] = np.asarray(v)
    return out


def keras_conv_layers(w):
    """Names of Keras' auto-named Conv2D layers (``conv2d``, ``conv2d_1``, ...) of a normalised dict, in creation order."""
    found = {}
    for k in w:
        layer, _, var = k.rpartition("/")
        m = _AUTO_CONV.match(layer)
        if m and var == "kernel":
            found[int(m.group(1) or 0)] = layer
    return [found[i] for i in sorted(found)]


def adopt(raw, spec, arch="network"):
    """Map a weight dict onto the names of ``spec`` (a ``*_weights(..., shapes_only=True)`` dict) and validate it.

    Accepts the internal names as they are.  The reference's SRCNN / EDSR models leave their Conv2D layers auto-named
    (SRCNN_model.py:50-52, EDSR_model.py:61-121), so a file exported from them carries ``conv2d``, ``conv2d_1``, ...:
    those are matched to the spec's conv layers in creation order (the uniquifying suffix only depends on how many models
    the session built before, so only the order is used).  ESRGAN and VGG16 layers carry explicit names in the reference
    (ESRGAN_model.py:230-341, Keras' ``block1_conv1`` ...), which are the internal names.  Missing layers and shape
    mismatches raise ``ValueError`` naming the layer, at load time."""
    w = normalize_keras_names(raw)
    spec_layers = [k[:-len("/kernel")] for k in spec if k.endswith("/kernel")]
    if not all((layer + "/kernel") in w for layer in spec_layers):
        auto = keras_conv_layers(w)
        conv_spec = [layer for layer in spec_layers if spec[layer + "/kernel"].ndim == 4]
        if auto and len(auto) == len(conv_spec) and len(conv_spec) == len(spec_layers):
            renamed = {}
            for src, dst in zip(auto, conv_spec):
                for var in ("kernel", "bias"):
                    if f"{src}/{var}" in w:
                        renamed[f"{dst}/{var}"] = w[f"{src}/{var}"]
            w = renamed
        else:
            missing = [layer for layer in spec_layers if (layer + "/kernel") not in w]
            raise ValueError(f"{arch}: weight file does not match the architecture: {len(missing)} layer(s) missing (first: "
                             f"{missing[0]!r}); found {len(auto)} auto-named Keras Conv2D layers where {len(conv_spec)} are "
                             "needed.  Expected the internal names or the reference model's own variable names "
                             "({v.name: v.numpy() for v in model.weights}).")
    out = {}
    for k, ref in spec.items():
        if k not in w:
            if k.endswith("/bias"):
                out[k] = np.zeros(ref.shape, np.float32)          # use_bias=False layers
                continue
            raise ValueError(f"{arch}: weight {k!r} is missing")
        a = np.asarray(w[k], dtype=np.float32)
        if tuple(a.shape) != tuple(ref.shape):
            raise ValueError(f"{arch}: weight {k!r} has shape {tuple(a.shape)}, expected {tuple(ref.shape)} (Keras HWIO)")
        out[k] = np.ascontiguousarray(a)
    return out


def count_params(w) -> int:
    return int(sum(v.size for v in w.values()))
