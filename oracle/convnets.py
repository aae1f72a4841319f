"""Oracle: Keras-semantics conv networks on the CPU (torch float32 / float64).

TEST INFRASTRUCTURE ONLY - see oracle/__init__.py.

Layer semantics follow Keras 2.10 as used by the reference:
* ``Conv2D(filters, k, padding="same")``: NHWC input, HWIO kernel, stride 1
  cross-correlation (no flip), zero padding (k-1)/2, then ``+ bias`` then activation
  (SRCNN_model.py:50-52, EDSR_model.py:61,65,102,112,121, ESRGAN_model.py:230-246).
* ``tf.nn.depth_to_space(x, r)`` is DCR:
  ``out[b, h*r+i, w*r+j, c] = in[b, h, w, (i*r+j)*C + c]`` (EDSR_model.py:81-90).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _t(x, dtype):
    return torch.from_numpy(np.ascontiguousarray(x)).to(dtype)


def conv2d_same(x, kernel, bias=None, dtype=torch.float32):
    """x: torch NCHW. kernel: numpy HWIO. Returns NCHW."""
    kh, kw, cin, cout = kernel.shape
    w = _t(kernel, dtype).permute(3, 2, 0, 1).contiguous()
    b = None if bias is None else _t(bias, dtype)
    return F.conv2d(x, w, b, stride=1, padding=(kh // 2, kw // 2))


def conv2d_same_numpy(x_nhwc, kernel, bias=None):
    """Independent float64 im2col restatement used to cross-check conv2d_same."""
    x = np.asarray(x_nhwc, dtype=np.float64)
    kh, kw, cin, cout = kernel.shape
    n, h, w, _ = x.shape
    ph, pw = kh // 2, kw // 2
    xp = np.pad(x, ((0, 0), (ph, ph), (pw, pw), (0, 0)))
    out = np.zeros((n, h, w, cout))
    for dy in range(kh):
        for dx in range(kw):
            out += xp[:, dy:dy + h, dx:dx + w, :] @ kernel[dy, dx].astype(np.float64)
    if bias is not None:
        out += bias.astype(np.float64)
    return out


def depth_to_space(x, r):
    """DCR depth_to_space on an NCHW torch tensor (channel index = (i*r+j)*C + c)."""
    n, c, h, w = x.shape
    co = c // (r * r)
    x = x.view(n, r, r, co, h, w)          # (n, i, j, c, h, w)
    x = x.permute(0, 3, 4, 1, 5, 2)        # (n, c, h, i, w, j)
    return x.reshape(n, co, h * r, w * r)


def depth_to_space_numpy(x_nhwc, r):
    n, h, w, c = x_nhwc.shape
    co = c // (r * r)
    x = x_nhwc.reshape(n, h, w, r, r, co).transpose(0, 1, 3, 2, 4, 5)
    return x.reshape(n, h * r, w * r, co)


def _in(x_nhwc, dtype):
    return _t(np.asarray(x_nhwc), dtype).permute(0, 3, 1, 2).contiguous()


def _out(y):
    return y.permute(0, 2, 3, 1).contiguous().numpy()


def _prelu(x, slope, dtype):
    a = _t(slope, dtype).view(1, -1, 1, 1)
    return torch.where(x >= 0, x, a * x)


def srcnn_forward(w, x_nhwc, dtype=torch.float32):
    """SRCNN_model.py:45-53: conv9x9x96 relu -> conv1x1x32 relu -> conv5x5x3 linear. No clip."""
    with torch.no_grad():
        x = _in(x_nhwc, dtype)
        x = F.relu(conv2d_same(x, w["conv1/kernel"], w["conv1/bias"], dtype))
        x = F.relu(conv2d_same(x, w["conv2/kernel"], w["conv2/bias"], dtype))
        x = conv2d_same(x, w["conv3/kernel"], w["conv3/bias"], dtype)
        return _out(x)


def edsr_forward(w, x_nhwc, scale_factor=2, num_res_blocks=16, res_scaling=0.1,
                 dtype=torch.float32):
    """EDSR_model.py:55-125: head, N x (conv relu conv *s +x), body conv + head skip,
    sub-pixel upsampler, RGB conv, clip(0,1)."""
    with torch.no_grad():
        x = _in(x_nhwc, dtype)
        x = conv2d_same(x, w["head/kernel"], w["head/bias"], dtype)
        head = x
        for i in range(num_res_blocks):
            sc = x
            y = F.relu(conv2d_same(x, w[f"rb{i}_c1/kernel"], w[f"rb{i}_c1/bias"], dtype))
            y = conv2d_same(y, w[f"rb{i}_c2/kernel"], w[f"rb{i}_c2/bias"], dtype)
            if res_scaling != 1.0:
                y = y * res_scaling
            x = y + sc
        x = conv2d_same(x, w["body/kernel"], w["body/bias"], dtype) + head
        if scale_factor in (2, 3):
            x = depth_to_space(conv2d_same(x, w["up0/kernel"], w["up0/bias"], dtype), scale_factor)
        elif scale_factor == 4:
            x = depth_to_space(conv2d_same(x, w["up0/kernel"], w["up0/bias"], dtype), 2)
            x = depth_to_space(conv2d_same(x, w["up1/kernel"], w["up1/bias"], dtype), 2)
        else:
            raise ValueError(f"Scale factor {scale_factor} not supported. Use 2, 3, or 4.")
        x = conv2d_same(x, w["tail/kernel"], w["tail/bias"], dtype)
        return _out(torch.clamp(x, 0.0, 1.0))


def espcn_forward(w, x_nhwc, scale_factor=4, activation="relu", dtype=torch.float32):
    """ESPCN composed from the reference's Conv2D + depth_to_space semantics (row A14)."""
    act = F.relu if activation == "relu" else torch.tanh
    with torch.no_grad():
        x = _in(x_nhwc, dtype)
        x = act(conv2d_same(x, w["conv1/kernel"], w["conv1/bias"], dtype))
        x = act(conv2d_same(x, w["conv2/kernel"], w["conv2/bias"], dtype))
        x = conv2d_same(x, w["conv3/kernel"], w["conv3/bias"], dtype)
        return _out(depth_to_space(x, scale_factor))


def srresnet_forward(w, x_nhwc, scale_factor=4, num_res_blocks=16, dtype=torch.float32):
    """SRResNet generator, BN folded (row A14): conv9 PReLU; N x (conv PReLU conv +x);
    conv + skip; log2(s) x (conv d2s PReLU); conv9."""
    with torch.no_grad():
        x = _in(x_nhwc, dtype)
        x = _prelu(conv2d_same(x, w["head/kernel"], w["head/bias"], dtype), w["head/prelu"], dtype)
        head = x
        for i in range(num_res_blocks):
            y = conv2d_same(x, w[f"rb{i}_c1/kernel"], w[f"rb{i}_c1/bias"], dtype)
            y = _prelu(y, w[f"rb{i}_c1/prelu"], dtype)
            y = conv2d_same(y, w[f"rb{i}_c2/kernel"], w[f"rb{i}_c2/bias"], dtype)
            x = x + y
        x = conv2d_same(x, w["body/kernel"], w["body/bias"], dtype) + head
        for i in range(2 if scale_factor == 4 else 1):
            x = depth_to_space(conv2d_same(x, w[f"up{i}/kernel"], w[f"up{i}/bias"], dtype), 2)
            x = _prelu(x, w[f"up{i}/prelu"], dtype)
        x = conv2d_same(x, w["tail/kernel"], w["tail/bias"], dtype)
        return _out(x)


def self_attention(w, name, x, dtype):
    """ESRGAN_model.py:48-70.  x: NCHW.  softmax over the key axis of g . f^T."""
    n, c, h, wd = x.shape
    f = conv2d_same(x, w[name + "_f/kernel"], w[name + "_f/bias"], dtype)
    g = conv2d_same(x, w[name + "_g/kernel"], w[name + "_g/bias"], dtype)
    hh = conv2d_same(x, w[name + "_h/kernel"], w[name + "_h/bias"], dtype)
    f_flat = f.flatten(2).transpose(1, 2)       # [B, HW, C/8]
    g_flat = g.flatten(2).transpose(1, 2)
    h_flat = hh.flatten(2).transpose(1, 2)      # [B, HW, C/2]
    s = torch.matmul(g_flat, f_flat.transpose(1, 2))
    beta = torch.softmax(s, dim=-1)
    o = torch.matmul(beta, h_flat)              # [B, HW, C/2]
    o = o.transpose(1, 2).reshape(n, -1, h, wd)
    o = conv2d_same(o, w[name + "_v/kernel"], w[name + "_v/bias"], dtype)
    return x + o


def esrgan_generator_forward(w, x_nhwc, scale_factor=2, num_rrdb_blocks=23,
                             dtype=torch.float32):
    """ESRGAN_model.py:212-345.  Input and output are in [-1, 1] (tanh); the caller maps
    back with (y + 1) / 2 (ESRGAN_model.py:946)."""
    def c(x, name, act=None):
        y = conv2d_same(x, w[name + "/kernel"], w[name + "/bias"], dtype)
        return F.relu(y) if act == "relu" else y

    def dense(x, name):
        feats = [x]
        for j in range(4):
            feats.append(c(torch.cat(feats, 1), f"{name}_conv{j + 1}", "relu"))
        x5 = c(torch.cat(feats, 1), f"{name}_conv5")
        return x + 0.2 * x5

    with torch.no_grad():
        x = _in(x_nhwc, dtype)
        x = c(x, "initial_conv")
        trunk = x
        for i in range(num_rrdb_blocks):
            inp = x
            for d in (1, 2, 3):
                x = dense(x, f"rrdb_{i}_dense{d}")
            x = inp + 0.2 * x
        x = trunk + c(x, "trunk_conv")
        x = self_attention(w, "self_attention_trunk", x, dtype)
        for i in range(int(np.log2(scale_factor))):
            x = depth_to_space(c(x, f"upsample_{i}_conv"), 2)
            x = F.leaky_relu(x, 0.2)
            if i == 0:
                x = self_attention(w, "self_attention_upsample_0", x, dtype)
        x = c(x, "final_conv1", "relu")
        x = torch.tanh(c(x, "final_conv2"))
        return _out(x)


def vgg16_classifier_forward(w, x_nhwc, dtype=torch.float32):
    """VGG16_model.py:57-97: 13 conv3x3 relu + 5 maxpool2x2 -> GAP -> Dense256 relu ->
    Dense softmax.  Dropout is the identity at inference.  No preprocess_input."""
    cfg = [(1, 2), (2, 2), (3, 3), (4, 3), (5, 3)]
    with torch.no_grad():
        x = _in(x_nhwc, dtype)
        for blk, n in cfg:
            for j in range(1, n + 1):
                nm = f"block{blk}_conv{j}"
                x = F.relu(conv2d_same(x, w[nm + "/kernel"], w[nm + "/bias"], dtype))
            x = F.max_pool2d(x, 2, 2)
        x = x.mean(dim=(2, 3))
        x = F.relu(x @ _t(w["dense/kernel"], dtype) + _t(w["dense/bias"], dtype))
        x = x @ _t(w["predictions/kernel"], dtype) + _t(w["predictions/bias"], dtype)
        return torch.softmax(x, dim=-1).numpy()


def bf16_round(a):
    """Round-to-nearest-even float32 -> bfloat16 -> float32 (for bf16 error budgeting)."""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).bfloat16().float().numpy()
