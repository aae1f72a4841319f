"""Network runners: the device-side replacement of Keras ``Model.predict`` for the reference's
SR networks (SRCNN_model.py:45-53, EDSR_model.py:55-125, ESRGAN_model.py:212-345), the two
BASELINE-named networks composed from the same layer semantics (ESPCN, SRResNet) and the VGG16
defect classifier (VGG16_model.py:57-97).

A network is a Keras-layout weight dict (``srb200.weights``) plus a forward function that
sequences ``ops.conv2d`` launches with fused epilogues on torch's current stream.  Activations
stay on the device in NHWC.  ``precision="fp16"`` / ``"bf16"`` keep them in a 16-bit format and route
the 64-channel 3x3 layers to the tcgen05 engine (same ``kind::f16`` instruction and rate for both;
fp32 accumulation; the residual trunk is carried in fp32 next to the 16-bit operand copies);
``precision="fp32"`` keeps everything in float32 on the exact CUDA-core engine (<= 1e-3 parity mode).
fp16 is the default 16-bit format: on the reference's he_normal random-init EDSR the 8-bit mantissa
of bf16 cannot meet the 2e-2 output tolerance (rounding the weights alone costs 5.8e-2), see DESIGN.md.
"""
from __future__ import annotations

import math
import time

import numpy as np

from . import _capi as capi
from . import ops
from . import weights as W


def _torch():
    return capi.require_cuda()


class DeviceModel:
    """Minimal stand-in for the ``keras.Model`` object the reference keeps in ``self.model``."""

    arch = "base"
    writes_into_out = False        # forward_device(x, out=...) stores the float32 result into a preallocated tensor

    def __init__(self, weights: dict, precision="fp16"):
        if precision not in ("bf16", "fp16", "fp32"):
            raise ValueError("precision must be 'fp16', 'bf16' or 'fp32'")
        torch = _torch()
        self.precision = precision
        self.act_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[precision]
        # how the residual trunk is carried in the 16-bit modes:
        #   "fp32" - fp32 trunk next to a 16-bit operand copy;  "half" - plain 16-bit trunk
        #   "pair" - compensated: h = round16(x) plus e = round16(x - h), two 16-bit tensors (~22 bits, 17 % fewer
        #            HBM bytes per res-block than fp32)
        #   "pair8" - as "pair" with e stored as float8 e5m2: 3 bytes per channel carry ~14 significant bits, which on
        #            EDSR x4 is as accurate as the fp32 trunk (tools/precision_budget.py) for a third fewer HBM bytes, and
        #            its rows move through TMA in the tcgen05 epilogue (default)
        self.trunk = "pair8" if precision != "fp32" else "none"
        self.fp32_trunk = False
        self.weights = {k: np.asarray(v, dtype=np.float32) for k, v in weights.items()}
        self.layers = {}
        for name in self.weights:
            if name.endswith("/kernel") and self.weights[name].ndim == 4:
                base = name[:-len("/kernel")]
                self.layers[base] = ops.ConvWeights(self.weights[name], self.weights.get(base + "/bias"),
                                                   self.weights.get(base + "/prelu"))
        self.max_device_batch = None   # images per launch sequence (None = whole input)
        self.event_hook = None         # bench.py: called with a tag at kernel-span boundaries

    @staticmethod
    def _first_conv(weights, name):
        """Kernel of the network's first Conv2D: ``<name>/kernel`` or, for a file exported from the reference's
        auto-named Keras layers, the first ``conv2d*`` layer in creation order."""
        w = W.normalize_keras_names(weights)
        if name + "/kernel" in w:
            return w[name + "/kernel"]
        auto = W.keras_conv_layers(w)
        if not auto:
            raise ValueError(f"weight dict has neither {name + '/kernel'!r} nor Keras auto-named conv2d layers")
        return w[auto[0] + "/kernel"]

    @staticmethod
    def _chain_spec(weights, names):
        """Shape spec of a plain conv chain taken from the file's own kernels (internal names or Keras creation order),
        after checking that every layer's input channels are the previous layer's outputs."""
        w = W.normalize_keras_names(weights)
        src = names if all(n + "/kernel" in w for n in names) else W.keras_conv_layers(w)
        if len(src) != len(names):
            raise ValueError(f"expected {len(names)} Conv2D layers ({', '.join(names)}), found {len(src)}")
        spec, prev = {}, None
        for n, sname in zip(names, src):
            k = np.asarray(w[sname + "/kernel"])
            if k.ndim != 4:
                raise ValueError(f"weight {sname + '/kernel'!r} must be HWIO [kh, kw, cin, cout], got shape {k.shape}")
            if prev is not None and k.shape[2] != prev:
                raise ValueError(f"layer {n!r} takes {k.shape[2]} input channels but the previous layer produces {prev}")
            prev = k.shape[3]
            spec[n + "/kernel"] = np.broadcast_to(np.float32(0), k.shape)
            spec[n + "/bias"] = np.broadcast_to(np.float32(0), (k.shape[3],))
        return spec

    # -- Keras-like surface ---------------------------------------------------------------
    def count_params(self):
        return int(sum(v.size for v in self.weights.values()))

    def summary(self):
        print(f"Model: {self.arch} ({self.precision})")
        for k, v in self.weights.items():
            print(f"  {k:40s} {tuple(v.shape)}")
        print(f"Total params: {self.count_params():,}")

    def get_weights_dict(self):
        return dict(self.weights)

    def output_scale(self):
        return 1

    def forward_device(self, x):
        raise NotImplementedError

    # bytes of live activations per INPUT pixel of one image (the widest point of the graph, all buffers the caching
    # allocator holds at once); subclasses refine it.  Only used to pick a default micro-batch.
    def _activation_bytes_per_input_pixel(self):
        s = self.output_scale()
        return 3 * 64 * 4 * s * s

    def default_micro_batch(self, h, w, budget_bytes=6 << 30):
        """Images per launch sequence when ``max_device_batch`` is not set: as many as keep the live activations under
        ``budget_bytes`` (the reference predicts 16 patches at a time, EDSR_model.py:274; one launch sequence over the
        whole input would need ~9 GB for the 1,849 patches of a 1024 x 1024 LR image and overflow 32-bit tile counts at 4K)."""
        per_image = max(1, int(h) * int(w) * self._activation_bytes_per_input_pixel())
        cap = max(256, (1 << 21) // max(1, int(h) * int(w)))      # small patches (24 x 24): up to ~2 M pixels per launch sequence
        return int(max(1, min(cap, budget_bytes // per_image)))

    def forward_device_as(self, x, out_dtype=None):
        """forward_device with the result in ``out_dtype`` (None = float32).  uint8 = saturate(rint(255 v)) of the [0, 1]
        image.  The default converts after the fact; networks whose last layer can store the type directly override it."""
        torch = _torch()
        y = self.forward_device(x)
        if out_dtype is None or out_dtype == torch.float32:
            return y
        if out_dtype == torch.uint8:
            return ops.cast(y.clamp_(0.0, 1.0), torch.uint8, scale=255.0)
        return ops.cast(y, out_dtype)

    def predict_device(self, x, micro_batch=None):
        """x: [B,H,W,C] float32 CUDA tensor -> float32 CUDA tensor.  No host synchronisation.  Works through the batch in
        micro-batches (``micro_batch`` / ``max_device_batch`` / ``default_micro_batch``), each written into its slice of
        one preallocated output."""
        torch = _torch()
        mb = micro_batch or self.max_device_batch or self.default_micro_batch(x.shape[1], x.shape[2])
        if x.shape[0] <= mb:
            return self.forward_device(x)
        if self.writes_into_out:
            # the network's last layer stores straight into its slice of the result (no copy of the output per micro-batch)
            out = torch.empty(self.output_shape(x.shape), dtype=torch.float32, device=x.device)
            for i in range(0, x.shape[0], mb):
                self.forward_device(x[i:i + mb], out=out[i:i + mb])
            return out
        out = None
        for i in range(0, x.shape[0], mb):
            y = self.forward_device(x[i:i + mb])
            if out is None:
                out = torch.empty((x.shape[0],) + tuple(y.shape[1:]), dtype=y.dtype, device=y.device)
            out[i:i + y.shape[0]].copy_(y)
        return out

    _NP_OUT = {"float32": "float32", "float16": "float16", "uint8": "uint8"}

    def predict(self, x, batch_size=32, verbose=0, out=None, out_dtype=None):
        """numpy in, numpy out (``self.model.predict(patches, batch_size=16, verbose=0)``,
        SRCNN_model.py:210, EDSR_model.py:274).  ``batch_size`` is accepted for signature parity; the
        device works in micro-batches of ``max_device_batch`` images, with the host->device copy of
        micro-batch i+1 and the device->host copy of micro-batch i-1 overlapping the kernels of
        micro-batch i on separate copy streams.

        The stock call ``predict(x)`` returns a fresh float32 array whose storage is page-locked (it comes from torch's
        caching host allocator, so a repeated call gets the previous result's block back without a new cudaHostAlloc once
        the caller has dropped it): the read-back runs at the PCIe rate without an ``out=`` argument.  ``out`` may be a
        preallocated (ideally pinned) host array.  ``out_dtype`` (np.float16 / np.uint8; extension) makes the last layer
        store that type - uint8 is saturate(rint(255 v)) - which cuts the read-back bytes by 2x / 4x."""
        torch = _torch()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 4:
            raise ValueError(f"expected a 4-D NHWC array, got shape {x.shape}")
        np_dt = np.dtype(np.float32 if out_dtype is None else out_dtype)
        if np_dt.name not in self._NP_OUT:
            raise ValueError("out_dtype must be float32, float16 or uint8")
        t_dt = {"float32": torch.float32, "float16": torch.float16, "uint8": torch.uint8}[np_dt.name]
        n = x.shape[0]
        out_shape = self.output_shape(x.shape)
        if out is None:
            ot = torch.empty(out_shape, dtype=t_dt, pin_memory=True)
            out = ot.numpy()                              # (keeps the pinned block alive for as long as the caller holds it)
        elif tuple(out.shape) != tuple(out_shape) or out.dtype != np_dt:
            raise ValueError(f"out must be {np_dt.name} of shape {out_shape}")
        else:
            ot = torch.from_numpy(out)
        if n == 0:
            return out
        xt = torch.from_numpy(x)
        mb = self.max_device_batch or min(64, self.default_micro_batch(x.shape[1], x.shape[2]))
        comp = torch.cuda.current_stream()
        cin, cout = self._copy_streams()
        # earlier work of the compute stream may still be using the memory the input buffers are carved from (or, when the
        # buffers are reused from a previous call, reading them): the copy stream starts after it
        cin.wait_stream(comp)
        # two persistent device input buffers and explicit events instead of per-chunk allocations on the copy streams
        # (record_stream defers the reuse of those blocks by the caching allocator, which showed up as occasional
        # cudaMalloc stalls in the middle of a run)
        key = (torch.cuda.current_device(), mb) + tuple(x.shape[1:])
        if getattr(self, "_in_key", None) != key:
            self._in_bufs = [torch.empty(key[1:], dtype=torch.float32, device="cuda") for _ in range(2)]
            self._in_key = key
        in_ready = [torch.cuda.Event(), torch.cuda.Event()]
        in_free = [None, None]
        prev = None                                       # (output of the previous chunk, event: its read-back has finished)
        for k, (i, m) in enumerate(self._chunks(n, mb)):
            b = k & 1
            xd = self._in_bufs[b][:m]
            with torch.cuda.stream(cin):
                if in_free[b] is not None:
                    cin.wait_event(in_free[b])            # the kernels that read this buffer two chunks ago are done
                xd.copy_(xt[i:i + m], non_blocking=True)
                in_ready[b].record(cin)
            comp.wait_event(in_ready[b])
            y = self.forward_device_as(xd, t_dt)
            in_free[b] = torch.cuda.Event()
            in_free[b].record(comp)
            cout.wait_stream(comp)
            with torch.cuda.stream(cout):
                ot[i:i + m].copy_(y, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(cout)
            if prev is not None:
                comp.wait_event(prev[1])                  # later kernels may reuse the previous output's memory: order them
            prev = (y, copied)                            # after its read-back (long finished by now), then drop it
        cout.synchronize()
        return out

    @staticmethod
    def _chunks(n, mb):
        """(start, size) pieces of n images: micro-batches of ``mb``, with the last one cut into a half and two quarters -
        only the final piece's device->host copy cannot hide behind later kernels, so it is kept small - and, when the
        read-back is the longer leg (two micro-batches or more), the first one cut into two quarters and a half, so that the
        device->host link starts after a quarter of a micro-batch's kernels instead of a whole one."""
        sizes = [mb] * (n // mb) + ([n % mb] if n % mb else [])
        last = sizes.pop()
        if last >= 16 and sizes:
            h, q = last // 2, last // 4
            sizes += [h, q, last - h - q]
        else:
            sizes.append(last)
        if n >= 2 * mb and mb >= 16:
            q = mb // 4
            sizes = [q, q, mb - 2 * q] + sizes[1:]
        out, i = [], 0
        for m in sizes:
            out.append((i, m))
            i += m
        return out

    def output_shape(self, in_shape):
        s = self.output_scale()
        return (in_shape[0], in_shape[1] * s, in_shape[2] * s, in_shape[3])

    def _copy_streams(self):
        torch = _torch()
        if getattr(self, "_streams", None) is None:
            self._streams = (torch.cuda.Stream(), torch.cuda.Stream())
        return self._streams

    __call__ = predict_device


class SRCNNNet(DeviceModel):
    """conv9x9x96 relu -> conv1x1x32 relu -> conv5x5x3 linear; no clip (SRCNN_model.py:45-53).

    fp32: three launches of the exact CUDA-core engine.  16-bit modes: the same network on the tcgen05 kernels, composed
    from their 64-channel shapes - conv1's 96 filters run as two passes of the RGB head kernel (filters [0, 64) and
    [64, 96) + zero filters) into one 128-channel buffer, the 1x1 layer as two passes over its 64-channel input slices
    with the partial sums in fp32 (its 32 outputs zero-padded to 64), then ReLU + cast, and the 5x5x32 -> 3 layer on the
    wide-tile few-channel kernel with zero rows for the padded input channels."""
    arch = "SRCNN"

    def __init__(self, weights, precision="fp32"):
        weights = W.adopt(weights, self._chain_spec(weights, ["conv1", "conv2", "conv3"]), "SRCNN")
        super().__init__(weights, precision)
        self._tc = None
        k1, k2, k3 = (self.weights[f"conv{i}/kernel"] for i in (1, 2, 3))
        c1, c2 = k1.shape[3], k2.shape[3]
        if (precision != "fp32" and k1.shape[2] == 3 and 64 < c1 <= 128 and k2.shape[2] == c1 and c2 <= 64
                and k3.shape[2] == c2 and k3.shape[3] <= 4 and max(k1.shape[0], k2.shape[0], k3.shape[0]) <= 9):
            def pad(a, axis, n):
                width = [(0, 0)] * a.ndim
                width[axis] = (0, n - a.shape[axis])
                return np.pad(a, width)
            b1 = self.weights.get("conv1/bias", np.zeros(c1, np.float32))
            b2 = self.weights.get("conv2/bias", np.zeros(c2, np.float32))
            k2p = pad(pad(k2, 2, 128), 3, 64)
            self._tc = {
                "c1a": ops.ConvWeights(k1[..., :64], b1[:64]),
                "c1b": ops.ConvWeights(pad(k1[..., 64:], 3, 64), pad(b1[64:], 0, 64)),
                "c2a": ops.ConvWeights(k2p[:, :, :64, :], pad(b2, 0, 64)),
                "c2b": ops.ConvWeights(k2p[:, :, 64:, :], None),
                "c3": ops.ConvWeights(pad(k3, 2, 64), self.weights.get("conv3/bias")),
            }

    def forward_device(self, x):
        torch = _torch()
        L, dt = self.layers, self.act_dtype
        if self._tc is not None:
            T = self._tc
            B, H, W, _ = x.shape
            h1 = torch.empty((B, H, W, 128), dtype=dt, device=x.device)
            ops.conv2d(x, T["c1a"], act="relu", out=h1, out_coffset=0)
            ops.conv2d(x, T["c1b"], act="relu", out=h1, out_coffset=64)
            acc = ops.conv2d(h1, T["c2a"], x_coffset=0, out_dtype=torch.float32)
            ops.conv2d(h1, T["c2b"], x_coffset=64, res1=acc, out=acc)
            h2 = ops.cast(acc, dt, relu=True)
            return ops.conv2d(h2, T["c3"], out_dtype=torch.float32)
        h = ops.conv2d(x, L["conv1"], act="relu", out_dtype=dt)
        h = ops.conv2d(h, L["conv2"], act="relu", out_dtype=dt)
        return ops.conv2d(h, L["conv3"], out_dtype=torch.float32)


class EDSRNet(DeviceModel):
    """EDSR_model.py:96-125.  Residual scaling, skip adds, depth_to_space and the final clip are all conv epilogues; in the
    16-bit modes the activation-free up-sampling tail (up-convs, depth_to_space, RGB conv) is ONE composed 5 x 5 launch, so
    the graph is 2*N + 3 launches (layer by layer: 2*N + 4, + 1 for x4)."""
    arch = "EDSR"

    def __init__(self, weights, scale_factor=2, num_res_blocks=16, res_scaling=0.1, precision="fp16", trunk=None,
                 upsampler=None):
        if scale_factor not in (2, 3, 4):
            raise ValueError(f"Scale factor {scale_factor} not supported. Use 2, 3, or 4.")
        head = self._first_conv(weights, "head")
        weights = W.adopt(weights, W.edsr_weights(scale_factor, head.shape[2], num_res_blocks, head.shape[3], shapes_only=True),
                          f"EDSR x{scale_factor} ({num_res_blocks} res-blocks)")
        super().__init__(weights, precision)
        if trunk is not None and precision != "fp32":
            if trunk not in ("pair", "pair8", "fp32", "half"):
                raise ValueError("trunk must be 'pair8', 'pair', 'fp32' or 'half'")
            self.trunk = trunk
        self.scale_factor, self.num_res_blocks, self.res_scaling = scale_factor, num_res_blocks, float(res_scaling)
        # the activation-free up-sampling tail (up-convs, depth_to_space, RGB conv: EDSR_model.py:117-123):
        #   "composed" - ONE 5 x 5 convolution 64 -> r*r*C with the exact composed weights (srb200.compose), a tenth of the
        #                FLOPs and no full-resolution 64-channel intermediates (16-bit modes, 64 filters; default there)
        #   "layered"  - layer by layer, as the reference builds it
        c_img = self.weights["tail/kernel"].shape[3]
        can_compose = (precision != "fp32" and self.weights["up0/kernel"].shape[2] == 64
                       and scale_factor * scale_factor * c_img <= 48)
        if upsampler is None:
            upsampler = "composed" if can_compose else "layered"
        if upsampler not in ("composed", "layered"):
            raise ValueError("upsampler must be 'composed' or 'layered'")
        if upsampler == "composed" and not can_compose:
            raise ValueError("the composed up-sampler needs a 16-bit precision mode, 64 filters and scale^2 * channels <= 48")
        self.upsampler = upsampler
        self._composed = None
        # res-blocks after the first update the (h, e) trunk pair in place: three live activation tensors instead of five
        # (smaller working set; 2-4 % faster at micro-batches of 6-12 tiles, equal at 32: tools/graph_probe.py)
        self.trunk_in_place = True

    def output_scale(self):
        return self.scale_factor

    def _activation_bytes_per_input_pixel(self):
        s = self.scale_factor
        if self.upsampler == "composed":
            return 12 * s * s + 6 * 64 * 4
        # widest point: the last up-sampling conv's 64-channel input and output at full resolution + the fp32 RGB result
        return (2 * 64 * 2 + 12) * s * s + 4 * 64 * 4

    def composed_upsampler(self):
        """The device-resident composed tail (built once: float64 composition on the host, ~1-3 s)."""
        if self._composed is None:
            from . import compose
            w, b = compose.compose_edsr_tail(self.weights, self.scale_factor)
            self._composed = ops.ComposedUpsampler(w, b, self.scale_factor, compose.weight_scale(w))
        return self._composed

    writes_into_out = True

    def forward_device(self, x, out=None):
        return self.forward_device_as(x, None, out=out)

    def forward_device_as(self, x, out_dtype=None, out=None):
        torch = _torch()
        L, dt = self.layers, self.act_dtype
        out_dtype = out_dtype or (out.dtype if out is not None else torch.float32)
        if self.trunk in ("pair", "pair8"):
            # compensated 16-bit trunk: (h, e) with h the tensor-core operand of the next conv and h + e the trunk value
            et = torch.float8_e5m2 if self.trunk == "pair8" else dt
            head, head_e = ops.conv2d(x, L["head"], out_dtype=dt, out2_dtype=et, out2_error=True)
            if self.event_hook:
                self.event_hook("tc_begin")
            h, e = head, head_e
            for i in range(self.num_res_blocks):
                t = ops.conv2d(h, L[f"rb{i}_c1"], act="relu", out_dtype=dt)
                if i == 0 or not self.trunk_in_place:         # (block 0 leaves the head pair for the global skip)
                    h, e = ops.conv2d(t, L[f"rb{i}_c2"], alpha=self.res_scaling, res1=h, res2=e, out_dtype=dt,
                                      out2_dtype=et, out2_error=True)
                else:                                         # the trunk pair is updated in place: three live tensors
                    ops.conv2d(t, L[f"rb{i}_c2"], alpha=self.res_scaling, res1=h, res2=e, out=h, out2=e, out2_error=True)
            h = ops.conv2d(h, L["body"], res1=head, res2=head_e, out_dtype=dt)
        elif self.trunk == "fp32":
            # trunk in fp32 (y), 16-bit copy (y2) as the next conv's tensor-core operand
            head, h = ops.conv2d(x, L["head"], out_dtype=torch.float32, out2_dtype=dt)
            if self.event_hook:
                self.event_hook("tc_begin")
            trunk = head
            for i in range(self.num_res_blocks):
                t = ops.conv2d(h, L[f"rb{i}_c1"], act="relu", out_dtype=dt)
                trunk, h = ops.conv2d(t, L[f"rb{i}_c2"], alpha=self.res_scaling, res1=trunk,
                                      out_dtype=torch.float32, out2_dtype=dt)
            h = ops.conv2d(h, L["body"], res1=head, out_dtype=dt)
        else:
            head = ops.conv2d(x, L["head"], out_dtype=dt)
            if self.event_hook:
                self.event_hook("tc_begin")
            h = head
            for i in range(self.num_res_blocks):
                t = ops.conv2d(h, L[f"rb{i}_c1"], act="relu", out_dtype=dt)
                h = ops.conv2d(t, L[f"rb{i}_c2"], alpha=self.res_scaling, res1=h, out_dtype=dt)
            h = ops.conv2d(h, L["body"], res1=head, out_dtype=dt)
        if self.upsampler == "composed" and x.shape[1] >= 2 and x.shape[2] >= 2:
            y = ops.upsample_composed(h, self.composed_upsampler(), clip01=True, out_dtype=out_dtype, out=out)
            if self.event_hook:
                self.event_hook("tc_end")
            return y
        if self.scale_factor in (2, 3):
            h = ops.conv2d(h, L["up0"], d2s=self.scale_factor, out_dtype=dt)
        else:
            h = ops.conv2d(h, L["up0"], d2s=2, out_dtype=dt)
            h = ops.conv2d(h, L["up1"], d2s=2, out_dtype=dt)
        y = ops.conv2d(h, L["tail"], clip01=True, out_dtype=out_dtype, out=out)   # (uint8: the epilogue quantises the clipped value)
        if self.event_hook:
            self.event_hook("tc_end")
        return y


class ESPCNNet(DeviceModel):
    """ESPCN 5-3-3 + depth_to_space(r) composed from the reference's layer semantics."""
    arch = "ESPCN"

    def __init__(self, weights, scale_factor=4, activation="relu", precision="fp16"):
        weights = W.adopt(weights, self._chain_spec(weights, ["conv1", "conv2", "conv3"]), "ESPCN")
        if weights["conv3/kernel"].shape[3] % (scale_factor * scale_factor):
            raise ValueError(f"ESPCN: the last layer's {weights['conv3/kernel'].shape[3]} filters are not a multiple of "
                             f"scale_factor^2 = {scale_factor * scale_factor}")
        super().__init__(weights, precision)
        self.scale_factor, self.activation = scale_factor, activation
        # 16-bit modes: the tensor-core engine needs Cin = 64, so conv3's 32 input channels are zero-padded to 64
        # (conv2 writes channels [0, 32) of a 64-wide buffer whose upper half stays zero; conv3's kernel gets zero
        # rows for the padded channels) - twice the MMA work of an exact K = 288 GEMM, still far ahead of CUDA cores
        k3 = self.weights["conv3/kernel"]
        self._pad3 = None
        if precision != "fp32" and k3.shape[2] == 32:
            k3p = np.zeros((k3.shape[0], k3.shape[1], 64, k3.shape[3]), np.float32)
            k3p[:, :, :32, :] = k3
            self._pad3 = ops.ConvWeights(k3p, self.weights.get("conv3/bias"))
        self._wide = {}

    def output_scale(self):
        return self.scale_factor

    writes_into_out = True

    def forward_device(self, x, out=None):
        torch = _torch()
        L, dt = self.layers, self.act_dtype
        h = ops.conv2d(x, L["conv1"], act=self.activation, out_dtype=dt)
        if self._pad3 is not None:
            B, H, W, _ = h.shape
            key = (B, H, W, dt, h.device)
            wide = self._wide.get(key)
            if wide is None:                          # the zero upper half is written once; conv2 only touches channels [0, 32)
                self._wide.clear()
                wide = self._wide[key] = torch.zeros((B, H, W, 64), dtype=dt, device=h.device)
            ops.conv2d(h, L["conv2"], act=self.activation, out=wide, out_coffset=0)
            return ops.conv2d(wide, self._pad3, d2s=self.scale_factor, out_dtype=torch.float32, out=out)
        h = ops.conv2d(h, L["conv2"], act=self.activation, out_dtype=dt)
        return ops.conv2d(h, L["conv3"], d2s=self.scale_factor, out_dtype=torch.float32, out=out)


class SRResNetNet(DeviceModel):
    """SRResNet / SRGAN generator with BatchNorm folded."""
    arch = "SRResNet"

    def __init__(self, weights, scale_factor=4, num_res_blocks=16, precision="fp16"):
        if scale_factor not in (2, 4):
            raise ValueError("SRResNet scale factor must be 2 or 4")
        head = self._first_conv(weights, "head")
        spec = W.srresnet_weights(scale_factor, head.shape[2], num_res_blocks, head.shape[3], shapes_only=True)
        prelu = {k: np.asarray(v, np.float32) for k, v in W.normalize_keras_names(weights).items() if k.endswith("/prelu")}
        missing = [k for k in spec if k.endswith("/prelu") and k not in prelu]
        if missing:
            raise ValueError(f"SRResNet: PReLU slope vector {missing[0]!r} is missing")
        weights = dict(W.adopt(weights, {k: v for k, v in spec.items() if not k.endswith("/prelu")},
                               f"SRResNet x{scale_factor} ({num_res_blocks} res-blocks)"), **prelu)
        super().__init__(weights, precision)
        self.scale_factor, self.num_res_blocks = scale_factor, num_res_blocks

    def output_scale(self):
        return self.scale_factor

    writes_into_out = True

    def forward_device(self, x, out=None):
        torch = _torch()
        L, dt = self.layers, self.act_dtype
        if self.trunk in ("pair", "pair8"):
            et = torch.float8_e5m2 if self.trunk == "pair8" else dt
            head, head_e = ops.conv2d(x, L["head"], act="prelu", out_dtype=dt, out2_dtype=et, out2_error=True)
            h, e = head, head_e
            for i in range(self.num_res_blocks):
                t = ops.conv2d(h, L[f"rb{i}_c1"], act="prelu", out_dtype=dt)
                h, e = ops.conv2d(t, L[f"rb{i}_c2"], res1=h, res2=e, out_dtype=dt, out2_dtype=et, out2_error=True)
            h = ops.conv2d(h, L["body"], res1=head, res2=head_e, out_dtype=dt)
        else:
            head = ops.conv2d(x, L["head"], act="prelu", out_dtype=dt)
            h = head
            for i in range(self.num_res_blocks):
                t = ops.conv2d(h, L[f"rb{i}_c1"], act="prelu", out_dtype=dt)
                h = ops.conv2d(t, L[f"rb{i}_c2"], res1=h, out_dtype=dt)
            h = ops.conv2d(h, L["body"], res1=head, out_dtype=dt)
        for i in range(2 if self.scale_factor == 4 else 1):
            h = ops.conv2d(h, L[f"up{i}"], act="prelu", d2s=2, out_dtype=dt)   # PReLU after the shuffle == before it
        return ops.conv2d(h, L["tail"], out_dtype=torch.float32, out=out)


class ESRGANGeneratorNet(DeviceModel):
    """RRDBNet + two SelfAttention layers (ESRGAN_model.py:30-79, 212-345).  Input/output in [-1, 1].

    Dense blocks are concat-free: each growth conv writes its slice of one NHWC buffer and the
    next conv reads the widened prefix."""
    arch = "ESRGAN-generator"

    def __init__(self, weights, scale_factor=2, growth_channels=32, num_rrdb_blocks=23, precision="fp32"):
        first = self._first_conv(weights, "initial_conv")
        weights = W.adopt(weights, W.esrgan_generator_weights(scale_factor, growth_channels, num_rrdb_blocks, first.shape[2],
                                                              shapes_only=True),
                          f"ESRGAN generator x{scale_factor} ({num_rrdb_blocks} RRDB, growth {growth_channels})")
        super().__init__(weights, precision)
        self.scale_factor, self.growth, self.num_rrdb = scale_factor, growth_channels, num_rrdb_blocks

    def output_scale(self):
        return self.scale_factor

    def _dense(self, src, dst, name, outer=None):
        """Dense block (ESRGAN_model.py:212-254) on WIDE buffers [B, H, W, 64 + 4 g]: channels [0, 64) of ``src`` hold the
        block input x, the four growth convs append their slices to ``src`` in place (concat-free), and
        ``x + 0.2 * conv5(.)`` lands in channels [0, 64) of ``dst`` - the next block's input, so nothing is copied between
        blocks.  With ``outer`` (the wide buffer holding the RRDB input) the RRDB's own ``outer + 0.2 * (.)``
        (ESRGAN_model.py:277-280) is folded into the same epilogue."""
        L, g = self.layers, self.growth
        for j in range(4):
            # reads channels [0, 64 + j*g) of the wide buffer via cstride; writes slice j
            ops.conv2d(src, L[f"{name}_conv{j + 1}"], act="relu", out=src, out_coffset=64 + j * g)
        if outer is None:
            ops.conv2d(src, L[f"{name}_conv5"], alpha=0.2, res1=src, out=dst, out_coffset=0)
        else:
            ops.conv2d(src, L[f"{name}_conv5"], alpha=0.04, res1=src, beta1=0.2, res2=outer, beta2=1.0, out=dst, out_coffset=0)

    def _dense_buffers(self, B, H, W, dtype, device):
        """Three wide buffers that rotate through the dense blocks of an RRDB (block input / output / the RRDB input kept for
        its skip).  Allocated ZEROED once per shape: the tcgen05 engine reads the input in 64-channel chunks, so a growth conv
        also reads the slices later convs of the block will write - against zero weight columns, which only cancel finite
        values; after the first use those slices hold the previous block's (finite) activations."""
        torch = _torch()
        key = (B, H, W, dtype, str(device))
        if getattr(self, "_dense_key", None) != key:
            self._dense_bufs = [torch.zeros((B, H, W, 64 + 4 * self.growth), dtype=dtype, device=device) for _ in range(3)]
            self._dense_key = key
        return self._dense_bufs

    def _attention(self, x, name):
        torch = _torch()
        L = self.layers
        B, H, W, Cc = x.shape
        f = ops.conv2d(x, L[name + "_f"], out_dtype=torch.float32)
        g = ops.conv2d(x, L[name + "_g"], out_dtype=torch.float32)
        h = ops.conv2d(x, L[name + "_h"], out_dtype=torch.float32)
        tc = x.dtype != torch.float32 and ops.self_attention_tc_eligible(f.shape[3], h.shape[3])
        o = ops.self_attention_core(f.view(B, H * W, -1), g.view(B, H * W, -1), h.view(B, H * W, -1), tensor_cores=tc)
        if x.dtype != torch.float32 and o.shape[2] % 8 == 0 and o.shape[2] >= 32:
            o = ops.cast(o, x.dtype)              # 16-bit operand: the 1x1 output projection runs on the tensor cores
        return ops.conv2d(o.view(B, H, W, -1), L[name + "_v"], res1=x, out_dtype=x.dtype)

    writes_into_out = True

    def forward_device(self, x, out=None):
        torch = _torch()
        L, dt = self.layers, self.act_dtype
        trunk = ops.conv2d(x, L["initial_conv"], out_dtype=dt)
        B, H, W, _ = trunk.shape
        bufs = self._dense_buffers(B, H, W, dt, trunk.device)
        cur = 0
        bufs[cur][..., :64].copy_(trunk)
        for i in range(self.num_rrdb):
            a, b, c = cur, (cur + 1) % 3, (cur + 2) % 3       # a keeps the RRDB input until the third block's skip
            self._dense(bufs[a], bufs[b], f"rrdb_{i}_dense1")
            self._dense(bufs[b], bufs[c], f"rrdb_{i}_dense2")
            self._dense(bufs[c], bufs[b], f"rrdb_{i}_dense3", outer=bufs[a])
            cur = b
        h = ops.conv2d(bufs[cur], L["trunk_conv"], res1=trunk, out_dtype=dt)
        h = self._attention(h, "self_attention_trunk")
        for i in range(int(math.log2(self.scale_factor))):
            h = ops.conv2d(h, L[f"upsample_{i}_conv"], act="leaky_relu", act_slope=0.2, d2s=2, out_dtype=dt)
            if i == 0:
                h = self._attention(h, "self_attention_upsample_0")
        h = ops.conv2d(h, L["final_conv1"], act="relu", out_dtype=dt)
        return ops.conv2d(h, L["final_conv2"], act="tanh", out_dtype=torch.float32, out=out)


class VGG16ClassifierNet(DeviceModel):
    """VGG16 conv base + GAP + Dense(256, relu) + Dense(C, softmax) (VGG16_model.py:57-97)."""
    arch = "VGG16-classifier"
    CFG = [(1, 2), (2, 2), (3, 3), (4, 3), (5, 3)]

    def __init__(self, weights, precision="fp32"):
        w = W.normalize_keras_names(weights)
        if "predictions/kernel" not in w:
            raise ValueError("VGG16 classifier: weight 'predictions/kernel' is missing")
        weights = W.adopt(weights, W.vgg16_classifier_weights(int(np.asarray(w["predictions/kernel"]).shape[-1]), shapes_only=True),
                          "VGG16 classifier")
        super().__init__(weights, precision)
        torch = _torch()
        self.dense = {k: torch.from_numpy(self.weights[k]).cuda() for k in
                      ("dense/kernel", "dense/bias", "predictions/kernel", "predictions/bias")}
        # 16-bit modes: layers with Cin = 64 s run on the tcgen05 engine in one launch (the kernel walks the input in
        # 64-channel K chunks with every chunk's weights resident).  ``slice_passes = True`` selects the round-1 scheme
        # instead - s launches over 64-channel slices of the input (x_coffset), partial sums accumulated in one fp32 tensor
        # through the residual input, then one ReLU + cast - kept for comparison.
        self.slice_passes = False
        self.slices = {}

    def _slices_of(self, name):
        if name not in self.slices:
            w = self.layers[name]
            self.slices[name] = None
            if self.precision != "fp32" and w.cin > 64 and w.cin % 64 == 0 and w.kh == 3:
                k, b = self.weights[name + "/kernel"], self.weights.get(name + "/bias")
                self.slices[name] = [ops.ConvWeights(k[:, :, 64 * i:64 * (i + 1), :], b if i == 0 else None)
                                     for i in range(w.cin // 64)]
        return self.slices[name]

    def _conv_relu(self, h, name):
        torch = _torch()
        sl = self._slices_of(name) if self.slice_passes else None
        if sl is None:
            return ops.conv2d(h, self.layers[name], act="relu", out_dtype=self.act_dtype)
        acc = ops.conv2d(h, sl[0], x_coffset=0, out_dtype=torch.float32)
        for i in range(1, len(sl)):
            ops.conv2d(h, sl[i], x_coffset=64 * i, res1=acc, out=acc)
        return ops.cast(acc, self.act_dtype, relu=True)

    def forward_device(self, x):
        h = x
        for blk, n in self.CFG:
            for j in range(1, n + 1):
                h = self._conv_relu(h, f"block{blk}_conv{j}")
            h = ops.maxpool2x2(h)
        d = self.dense
        return ops.gap_dense_softmax(h, d["dense/kernel"], d["dense/bias"], d["predictions/kernel"],
                                     d["predictions/bias"])

    def predict(self, x, batch_size=32, verbose=0, out=None):
        torch = _torch()
        x = np.ascontiguousarray(x, dtype=np.float32)
        return self.predict_device(torch.from_numpy(x).cuda(non_blocking=True)).cpu().numpy()


def gpu_memory_info():
    """``tf.config.experimental.get_memory_info("GPU:0")`` analogue -> {"current", "peak"} bytes."""
    torch = _torch()
    return {"current": int(torch.cuda.memory_allocated()), "peak": int(torch.cuda.max_memory_allocated())}


def timed_predict(model: DeviceModel, patches_device):
    """Run ``predict_device`` bracketed the way the reference brackets ``model.predict``
    (perf_counter + memory info, SRCNN_model.py:200-242) -> (output, inference_metrics)."""
    torch = _torch()
    begin = gpu_memory_info()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = model.predict_device(patches_device)
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    end = gpu_memory_info()
    mb = 1024.0 * 1024.0
    return out, {
        "time_sec": float(elapsed),
        "gpu_mean_current_mb": (begin["current"] + end["current"]) / 2.0 / mb,
        "gpu_peak_mb": max(begin["peak"], end["peak"]) / mb,
    }
