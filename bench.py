#!/usr/bin/env python
"""bench.py - SR output megapixels/s of the B200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload all|edsr|configs]

Headline workload (BASELINE configs[2]): EDSR-baseline x4 (16 res-blocks, 64 channels), 16-bit operands /
fp32 accumulation, batch 512 LR tiles of 192x192x3 -> 768x768x3, sharded contiguously over N GPUs (one
process per GPU, no data-path collective).  A step is one pass of the network over the whole batch.
Prints ONE JSON line on rank 0.

* ``value``: whole-job throughput with the LR batch already resident in HBM (CUDA events, max over ranks).
* ``e2e``: the same metric through the reference-facing call ``EDSR.model.predict(host array)`` - the stock
  signature, pageable float32 input, fresh float32 result: H2D of the LR batch and D2H of the SR batch are inside
  the timed region (pipelined against compute on copy streams).  ``e2e.variants`` adds the pinned-``out=`` call and
  the ``out_dtype`` float16 / uint8 read-backs; ``e2e.host_ceiling_gbs`` is the host link measured alone.
* ``roofline``: the tcgen05 conv kernels (every layer but the 3-channel head): algorithmic FLOPs / CUDA-event
  time of the spans that contain only those kernels, against the measured sustained bf16 peak.
* ``parity``: GPU output vs the fp32 CPU oracle on actual 192x192 tiles of the workload (max-abs, PSNR / SSIM deltas
  against the synthetic HR, with the repo's metric kernel and with the oracle's metric).
* ``alt_dtype``: the same network with bf16 operands (the dtype BASELINE configs[2] names): rate and max-abs.
* ``collective``: the sharded ``evaluate`` - SR vs HR on every rank -> fused PSNR/SSIM kernel -> ONE float64
  all-reduce of (sum psnr, sum ssim, count, sum mse) over NCCL - timed, and checked against the unsharded means.
* ``configs``: the other BASELINE configs (c1 SRCNN pipeline, c2 ESPCN x4, c4 SRResNet x4 -> VGG16 vote on 1,024
  tiles, c5 bicubic / PSNR / PSNR+SSIM bandwidth sweep at 1K / 2K / 4K) each with its roofline fraction.
* ``cpu_baseline`` / ``--impl reference``: the reference's engine (TensorFlow) cannot be installed here, so the
  CPU arm is the fp32 oracle port (torch CPU, all host threads) on a bounded sample of the same tiles.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "super-resolution-images-for-3d-printing-defect-detection_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

EDSR_FLOP_PER_LR_PX = 2 * 1_983_168          # SURVEY.md section 8 row A5 (x4, 16 blocks, 64 ch)
EDSR_HEAD_FLOP_PER_LR_PX = 2 * 27 * 64       # the one layer that is not on the tcgen05 kernel
EDSR_BODY_FLOP_PER_LR_PX = 33 * 2 * 9 * 64 * 64      # 16 x 2 res-block convs + the body-end conv
EDSR_COMPOSED_TAIL_FLOP_PER_LR_PX = 2 * 25 * 64 * 48  # the up-sampling tail as one 5 x 5 conv 64 -> 4 x 4 x RGB (srb200/compose.py)
ESPCN_FLOP_PER_LR_PX = 2 * 37_056            # row A14
SRRESNET_FLOP_PER_LR_PX = 2 * 2_218_176      # row A14
VGG16_FLOP_PER_INPUT_PX = 2 * 305_856        # row A13
SRCNN_FLOP_PER_PX = 57_600                   # row A2


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


def measured_traffic():
    """ncu dram__bytes_read + dram__bytes_write per launch of the dominant kernels, from the committed summary the last
    `ncu --set full` capture was reduced to (profiles/traffic.json); None when there is no such file."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


class ClockSampler:
    """SM clocks / throttle reasons / power sampled DURING the timed region: NVML polled from a thread every few
    milliseconds (the timed region of a multi-GPU run is a few tens of milliseconds - shorter than nvidia-smi's start-up),
    with `nvidia-smi -lms` as the fallback when the NVML binding is missing."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index):
        self.index, self.proc, self.thread = index, None, None
        self.result = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _poll(self, nv, h):
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self._samples.append((sm, mask, watts))
            except Exception:
                pass
            self._stop.wait(0.003)

    def __enter__(self):
        import threading
        self._stop, self._samples = threading.Event(), []
        try:
            nv, h = self._nvml_handle()
            self._max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._poll, args=(nv, h), daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            if self._samples:
                sm = [s[0] for s in self._samples]
                mask = 0
                for s in self._samples:
                    mask |= int(s[1])
                self.result = {"sm_mhz": float(statistics.median(sm)), "sm_max_mhz": float(self._max),
                               "reasons": sorted(n for n, bit in self.REASONS if mask & bit), "samples": len(sm),
                               "power_w_max": max(s[2] for s in self._samples), "source": "nvml"}
            return
        if not self.proc:
            return
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, watts = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); watts.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            self.result = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                           "samples": len(sm), "power_w_max": max(watts), "source": "nvidia-smi"}


def layer_group_rooflines(net, x, torch, ops):
    """One CUDA-event pair around every conv launch of EDSRNet.forward_device (5 forwards, medians), grouped by layer role."""
    records, orig = [], ops.conv2d

    def timed(xx, w, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(xx, w, **kw)
        e1.record()
        # keep sizes only: holding the tensors would stop the caching allocator from reusing the big activations
        meta = {"out_dtype": kw.get("out_dtype"), "out2_dtype": kw.get("out2_dtype"),
                "res1": None if kw.get("res1") is None else kw["res1"].element_size(),
                "res2": None if kw.get("res2") is None else kw["res2"].element_size()}
        records.append((w, xx.shape[0] * xx.shape[1] * xx.shape[2], meta, e0, e1))
        return out
    orig_up = ops.upsample_composed

    def timed_up(xx, up, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_up(xx, up, **kw)
        e1.record()
        records.append((up, xx.shape[0] * xx.shape[1] * xx.shape[2], {"composed": True, "out_dtype": kw.get("out_dtype")}, e0, e1))
        return out
    hook, net.event_hook = net.event_hook, None
    ops.conv2d = timed
    ops.upsample_composed = timed_up
    try:
        per_rep = []
        for _ in range(5):
            records.clear()
            net.forward_device(x)
            torch.cuda.synchronize()
            per_rep.append([(w, npx, kw, e0.elapsed_time(e1)) for (w, npx, kw, e0, e1) in records])
    finally:
        ops.conv2d = orig
        ops.upsample_composed = orig_up
        net.event_hook = hook
    groups = {}
    for i, (w, npx, kw, _) in enumerate(per_rep[0]):
        ms = statistics.median(rep[i][3] for rep in per_rep)
        if kw.get("composed"):
            name = f"up-sampling tail composed: 5x5 64->{w.cout} + depth_to_space({w.scale}) to the image (one launch)"
            g = groups.setdefault(name, {"layer_group": name, "launches": 0, "ms": 0.0, "flop": 0.0, "bytes": 0.0})
            g["launches"] += 1
            g["ms"] += ms
            g["flop"] += 2.0 * npx * 25 * w.cin * w.cout
            g["bytes"] += npx * w.cin * 2 + npx * w.cout * (4 if kw.get("out_dtype") in (None, torch.float32) else 2)
            continue
        in_b = npx * w.cin * (4 if w.cin == 3 else 2)
        out_dt = kw.get("out_dtype")
        out_b = npx * w.cout * (4 if out_dt == torch.float32 else 2)
        if kw.get("out2_dtype") is not None:
            out_b += npx * w.cout * torch.empty((), dtype=kw["out2_dtype"]).element_size()
        for rk in ("res1", "res2"):
            if kw.get(rk) is not None:
                in_b += npx * w.cout * kw[rk]
        if w.cin == 3:
            name = "head 3x3 3->64 (im2col GEMM)"
        elif w.cout <= 4:
            name = "RGB tail 3x3 64->3 (wide-tile fold)"
        elif w.cout == 256:
            name = "up-sampling 3x3 64->256 + depth_to_space (CTA pairs)"
        elif kw.get("res1") is not None:
            name = "res-block 2nd conv / body end 3x3 64->64 + residual (pair8 trunk)"
        else:
            name = "res-block 1st conv 3x3 64->64 + ReLU (wide-tile fold)"
        g = groups.setdefault(name, {"layer_group": name, "launches": 0, "ms": 0.0, "flop": 0.0, "bytes": 0.0})
        g["launches"] += 1
        g["ms"] += ms
        g["flop"] += 2.0 * npx * w.kh * w.kw * w.cin * w.cout
        g["bytes"] += in_b + out_b
    total = sum(g["ms"] for g in groups.values()) or 1.0
    out = []
    for g in sorted(groups.values(), key=lambda v: -v["ms"]):
        out.append({"layer_group": g["layer_group"], "launches": g["launches"], "ms_per_forward": round(g["ms"], 4),
                    "share": round(g["ms"] / total, 4), "tflops": g["flop"] / g["ms"] / 1e9, "gbs_min": g["bytes"] / g["ms"] / 1e6})
    return out


def lr_tiles(lo, hi, size):
    from srb200 import synth
    return np.stack([synth.hr_image(size, size, i) for i in range(lo, hi)]) if hi > lo else \
        np.zeros((0, size, size, 3), np.float32)


def hr_lr_pair(n, lr_size, scale, first_index=10_000):
    """n synthetic (HR, LR) pairs of the workload's geometry: HR (lr_size * scale)^2, LR its exact area down-sample."""
    from srb200 import synth
    hr = synth.hr_batch(n, lr_size * scale, lr_size * scale, first_index=first_index)
    return hr, synth.area_downsample(hr, scale)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores (reference engine = TensorFlow 2.10, not installable here)
# ------------------------------------------------------------------------------------------------
def cpu_edsr_mp_per_s(n_tiles, tile, repeats=1):
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    from oracle import convnets as oc
    from srb200 import weights
    w = weights.edsr_weights(4)
    x = lr_tiles(0, n_tiles, tile)
    oc.edsr_forward(w, x[:1], 4, 16)                      # warm-up (thread pool, allocator)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        for i in range(n_tiles):
            oc.edsr_forward(w, x[i:i + 1], 4, 16)
        best = min(best, time.perf_counter() - t0)
    return n_tiles * (tile * 4) ** 2 / 1e6 / best, torch.get_num_threads(), best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tile, per_step = args.tile, 4
    import torch
    torch.set_num_threads(os.cpu_count() or 1)            # torchrun exports OMP_NUM_THREADS=1; use every host core
    from oracle import convnets as oc
    from srb200 import weights
    w = weights.edsr_weights(4)
    x = lr_tiles(0, per_step, tile)
    for _ in range(max(args.warmup, 1)):
        oc.edsr_forward(w, x[:1], 4, 16)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(per_step):
            oc.edsr_forward(w, x[i:i + 1], 4, 16)
    dt = time.perf_counter() - t0
    mp = args.steps * per_step * (tile * 4) ** 2 / 1e6
    val = mp / dt
    cores = torch.get_num_threads()
    sample = f"{per_step} of the {args.batch} LR tiles ({tile}x{tile}) per step, fp32 oracle port (torch CPU conv2d), {cores} threads"
    emit({
        "impl": "reference", "metric": "sr_output_megapixels_per_sec", "value": val, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args, 0), operands="fp32 (oracle port, torch CPU conv2d)", micro_batch=per_step),
        "cpu_baseline": {"value": val, "unit": "MP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference engine (TensorFlow/Keras 2.10) is not installable offline; CPU arm = oracle restatement",
    })


def workload_config(args, micro_batch):
    return {"workload": f"EDSR-baseline x4 (16 res-blocks, 64 ch) inference, {args.batch} LR tiles {args.tile}x{args.tile}x3 "
                        f"-> {args.tile * 4}x{args.tile * 4}x3 (BASELINE configs[2])",
            "global_batch": args.batch, "tile": args.tile, "scale": 4, "micro_batch": micro_batch,
            "parallelism": f"dp{args.gpus} (batch-sharded, no data-path collective)",
            "operands": f"{args.dtype} operands, fp32 accumulate, residual trunk = {args.trunk}",
            "upsampler": ("composed: up-convs + depth_to_space + RGB conv (no activation in between, EDSR_model.py:117-123) folded "
                          "exactly into one 5x5 conv 64->48 with nine border variants; `layered_upsampler` holds the layer-by-layer run")
            if getattr(args, "upsampler", "composed") == "composed" else "layered (layer by layer)",
            "l2": "inputs larger than L2 (LR batch + activations stream through HBM every step); no flush needed"}


# ------------------------------------------------------------------------------------------------
# GPU arm: pieces
# ------------------------------------------------------------------------------------------------
def cuda_timed(torch, fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def host_link_gbs(torch, dev, mib=512):
    """Pinned host <-> device copy rate of this rank, measured alone (no kernels running): the ceiling of the e2e line."""
    n = mib << 18
    h = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    out = {}
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        ms = min(cuda_timed(torch, fn, 1, warm=(1 if k == 0 else 0)) for k in range(5))   # a ceiling: the best of five copies
        out[name] = n * 4 / ms / 1e6
    return out


def parity_block(torch, net, net_alt, tile, n_tiles=2):
    """GPU output vs the fp32 CPU oracle on `n_tiles` tiles of the workload's geometry (LR tile x tile, x4), and the
    PSNR / SSIM of both against the synthetic HR - with the repo's metric kernel on the GPU output and the oracle's metric
    on the oracle output (the full chain a user compares) and with the same oracle metric on both (the network alone)."""
    from oracle import convnets as oc, metrics as om
    from srb200 import ops, weights
    hr, lr = hr_lr_pair(n_tiles, tile, 4)
    want = oc.edsr_forward(weights.edsr_weights(4), lr, 4, 16)
    hr_t = torch.from_numpy(hr).cuda()
    out = {}
    for key, model in (("main", net), ("alt", net_alt)):
        if model is None:
            continue
        got_t = model.forward_device(torch.from_numpy(lr).cuda())
        p_gpu, s_gpu = ops.psnr_ssim(hr_t, got_t.contiguous())
        got = got_t.cpu().numpy()
        p_ref, s_ref = om.psnr(hr, want, dtype=np.float64), om.ssim(hr, want, dtype=np.float64)
        p_mix, s_mix = om.psnr(hr, got, dtype=np.float64), om.ssim(hr, got, dtype=np.float64)
        out[key] = {"max_abs": float(np.abs(got - want).max()),
                    "psnr_delta_db": float(np.abs(p_gpu.cpu().numpy() - p_ref).max()),
                    "ssim_delta": float(np.abs(s_gpu.cpu().numpy() - s_ref).max()),
                    "psnr_delta_db_same_metric": float(np.abs(p_mix - p_ref).max()),
                    "ssim_delta_same_metric": float(np.abs(s_mix - s_ref).max()),
                    "psnr_db_oracle": [float(v) for v in p_ref], "ssim_oracle": [float(v) for v in s_ref]}
    return out


def collective_block(torch, D, net, rank, world, tile, n_eval=16):
    """The sharded ``evaluate`` of SURVEY section 8e at this N: every rank super-resolves its contiguous shard of a small
    synthetic test set, reduces it with the fused PSNR/SSIM kernel into (sum psnr, sum ssim, count, sum mse) on the device,
    and ONE float64 all-reduce (NCCL) forms the Keras-style sample means.  Rank 0 also evaluates the whole set alone and the
    two must agree to 0.01 dB / 1e-4 (they differ only in the float64 summation order)."""
    from srb200 import ops
    hr, lr = hr_lr_pair(n_eval, tile, 4, first_index=20_000)
    lo, hi = D.shard_bounds(n_eval, rank, world)
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    if hi > lo:
        sr = net.forward_device(torch.from_numpy(lr[lo:hi]).cuda())
        ops.psnr_ssim(torch.from_numpy(hr[lo:hi]).cuda(), sr.contiguous(), 1.0, sums=sums, want_mse=True)
    us = []
    total = sums
    for rep in range(6):                                      # the first all-reduce pays NCCL's lazy channel set-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t = D.allreduce_sums(sums)
        e1.record()
        torch.cuda.synchronize()
        us.append(e0.elapsed_time(e1) * 1e3)
        if rep == 0:
            total = t
    means = D.means_from_sums(total.cpu().numpy())
    res = {"op": "all_reduce(SUM) of float64[4] (sum psnr, sum ssim, count, sum mse)", "backend": "nccl" if world > 1 else "none (1 rank)",
           "eval_tiles": n_eval, "count": float(total[2].item()), "mean_mse": means[0], "mean_psnr_db": means[1], "mean_ssim": means[2],
           "allreduce_us_first": us[0], "allreduce_us": float(statistics.median(us[1:]))}
    if rank == 0:
        full = torch.zeros(4, dtype=torch.float64, device="cuda")
        for i in range(0, n_eval, 8):
            sr = net.forward_device(torch.from_numpy(lr[i:i + 8]).cuda())
            ops.psnr_ssim(torch.from_numpy(hr[i:i + 8]).cuda(), sr.contiguous(), 1.0, sums=full, want_mse=True)
        ref = D.means_from_sums(full.cpu().numpy())
        res["unsharded_mean_psnr_db"], res["unsharded_mean_ssim"] = ref[1], ref[2]
        res["psnr_diff_db"], res["ssim_diff"] = abs(ref[1] - means[1]), abs(ref[2] - means[2])
        res["matches_unsharded"] = bool(res["psnr_diff_db"] <= 0.01 and res["ssim_diff"] <= 1e-4 and res["count"] == n_eval)
    return res


def config_c1(torch, peaks):
    """C1: bicubic x2 pre-upsample (+ clip) -> SRCNN 9-1-5 -> PSNR/SSIM on 64 synthetic 128x128 images (fp32 engine)."""
    from oracle import bicubic as ob, convnets as oc, metrics as om
    from srb200 import engine, ops, synth, weights
    hr = synth.hr_batch(64, 128, 128, first_index=30_000)
    lr = synth.area_downsample(hr, 2)
    w = weights.srcnn_weights()
    hr_t, lr_t = torch.from_numpy(hr).cuda(), torch.from_numpy(lr).cuda()
    out = {"workload": "bicubic x2 + SRCNN 9-1-5 + PSNR/SSIM, 64 images 128x128 (BASELINE configs[0])"}
    for prec in ("fp32", "fp16"):
        net = engine.SRCNNNet(w, precision=prec)

        def step():
            up = ops.bicubic(lr_t, 128, 128, clip01=True)
            return ops.psnr_ssim(hr_t, net.forward_device(up))
        ms = cuda_timed(torch, step, 5)
        p, s = step()
        out[prec] = {"ms": ms, "out_MPps": 64 * 128 * 128 / ms / 1e3, "tflops": 64 * 128 * 128 * SRCNN_FLOP_PER_PX / ms / 1e9,
                     "mean_psnr_db": float(p.mean().item()), "mean_ssim": float(s.mean().item())}
        if prec == "fp32":                                # parity of the whole chain on 4 of the images (oracle: ~1 s)
            up = np.clip(np.stack([ob.resize_cubic_f32(im, (128, 128)) for im in lr[:4]]), 0, 1)
            want = oc.srcnn_forward(w, up)
            got = net.forward_device(ops.bicubic(lr_t[:4].contiguous(), 128, 128, clip01=True)).cpu().numpy()
            out["parity_fp32"] = {"images": 4, "max_abs": float(np.abs(got - want).max()),
                                  "psnr_delta_db": float(np.abs(p[:4].cpu().numpy() - om.psnr(hr[:4], want, dtype=np.float64)).max()),
                                  "ssim_delta": float(np.abs(s[:4].cpu().numpy() - om.ssim(hr[:4], want, dtype=np.float64)).max())}
    t0 = time.perf_counter()
    up = np.clip(np.stack([ob.resize_cubic_f32(im, (128, 128)) for im in lr[:8]]), 0, 1)
    pred = oc.srcnn_forward(w, up)
    om.psnr(hr[:8], pred), om.ssim(hr[:8], pred)
    out["cpu_oracle_out_MPps"] = 8 * 128 * 128 / 1e6 / (time.perf_counter() - t0)
    return out


def config_c2(torch, peaks):
    """C2: ESPCN x4, 256 LR tiles 256x256 -> 1024x1024 on one GPU."""
    from srb200 import engine, weights
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand((256, 256, 256, 3), device="cuda", generator=g)
    net = engine.ESPCNNet(weights.espcn_weights(4), 4, precision="fp16")
    net.max_device_batch = 64
    ms = cuda_timed(torch, lambda: net.predict_device(x), 5)
    out_px, lr_px = 256 * 1024 * 1024, 256 * 256 * 256
    fused_bytes = lr_px * 3 * 4 + out_px * 3 * 4                      # fp32 RGB in + fp32 RGB out (no intermediates)
    layer_bytes = fused_bytes + lr_px * (64 * 2 * 2 + 32 * 2 * 2)     # + the two 16-bit intermediates written and re-read
    tf = lr_px * ESPCN_FLOP_PER_LR_PX / ms / 1e9
    return {"workload": "ESPCN x4 (5-3-3, relu), 256 LR tiles 256x256 -> 1024x1024, fp16 operands (BASELINE configs[1])",
            "ms": ms, "out_MPps": out_px / ms / 1e3, "tflops": tf, "tensor_frac": tf / peaks["tflops"],
            "hbm_gbs_fused_ideal": fused_bytes / ms / 1e6, "hbm_frac_fused_ideal": fused_bytes / ms / 1e6 / peaks["hbm_gbs"],
            "hbm_gbs_layerwise": layer_bytes / ms / 1e6, "hbm_frac_layerwise": layer_bytes / ms / 1e6 / peaks["hbm_gbs"],
            "bound": "hbm (layer-at-a-time: 3 launches per micro-batch)"}


def config_c4(torch, peaks, batch=1024):
    """C4: SRResNet x4 on `batch` LR tiles 128x128 -> 512x512, then the VGG16 defect classifier on the SR outputs
    (reading A of SURVEY row A13: 128x128 patches at stride 64 of every SR image -> 49 patches, patch vote)."""
    from srb200 import engine, ops, weights
    from srb200.defect_detection_models.VGG16_model import vote
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand((batch, 128, 128, 3), device="cuda", generator=g)
    sr_net = engine.SRResNetNet(weights.srresnet_weights(4), 4, 16, precision="fp16")
    sr_net.max_device_batch = 32
    vgg = engine.VGG16ClassifierNet(weights.vgg16_classifier_weights(2), precision="fp16")
    probs_all = []

    def sr_pass():
        return sr_net.predict_device(x)

    def vgg_pass(sr):
        probs_all.clear()
        group = 8                                               # images whose patches share one classifier launch sequence
        for i in range(0, batch, group):
            patches = torch.cat([ops.pad_extract(sr[j].clamp(0.0, 1.0), 128, 64)[0] for j in range(i, min(i + group, batch))])
            probs_all.append(vgg.predict_device(patches, micro_batch=patches.shape[0]))
    sr = sr_pass()
    vgg_pass(sr)                                                # warm-up of both
    torch.cuda.synchronize()
    ms_sr = cuda_timed(torch, sr_pass, 2, warm=0)
    ms_vgg = cuda_timed(torch, lambda: vgg_pass(sr), 1, warm=0)
    probs = torch.cat(probs_all).cpu().numpy()
    n_patch = probs.shape[0] // batch
    votes = [vote(probs[i * n_patch:(i + 1) * n_patch])[0] for i in range(batch)]
    lr_px, out_px, cls_px = batch * 128 * 128, batch * 512 * 512, probs.shape[0] * 128 * 128
    tf_sr, tf_vgg = lr_px * SRRESNET_FLOP_PER_LR_PX / ms_sr / 1e9, cls_px * VGG16_FLOP_PER_INPUT_PX / ms_vgg / 1e9
    return {"workload": f"SRResNet x4 on {batch} LR tiles 128x128 -> 512x512 (fp16), then VGG16 classifier + patch vote on the SR "
                        f"outputs ({n_patch} patches 128x128 / stride 64 per image, fp16) (BASELINE configs[3])",
            "batch": batch, "sr_ms": ms_sr, "sr_out_MPps": out_px / ms_sr / 1e3, "sr_tflops": tf_sr, "sr_tensor_frac": tf_sr / peaks["tflops"],
            "classifier_ms": ms_vgg, "classifier_patches_per_s": probs.shape[0] / ms_vgg * 1e3, "classifier_tflops": tf_vgg,
            "classifier_tensor_frac": tf_vgg / peaks["tflops"], "pipeline_ms": ms_sr + ms_vgg,
            "pipeline_out_MPps": out_px / (ms_sr + ms_vgg) / 1e3, "votes_class1": int(sum(votes)), "bound": "tensor"}


def config_esrgan(torch, peaks):
    """The reference's own GAN generator at its trained configuration (ESRGAN.ipynb cell 6: 4 RRDB, growth 8, x2, 24x24 LR
    patches): the 1,521 patches of one 478x478 image (stride 12), fp16 operands on the tcgen05 engine (dense-block growth convs
    in 64-channel K chunks, 8-channel slices written by a 16-byte-per-pixel epilogue), SelfAttention on 576 and 2,304 positions
    per patch as a tcgen05 flash-style kernel (hi / lo split scores, online softmax, fp16 P) in the 16-bit mode."""
    from srb200 import engine, weights
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.rand((1521, 24, 24, 3), device="cuda", generator=g) * 2 - 1
    out = {"workload": "ESRGAN generator x2 (4 RRDB, growth 8, 2 SelfAttention), 1,521 LR patches 24x24 -> 48x48; fp16: convolutions and "
                       "both attention products on tcgen05; fp32: exact CUDA-core engines"}
    for prec in ("fp16", "fp32"):
        net = engine.ESRGANGeneratorNet(weights.esrgan_generator_weights(2, 8, 4), 2, 8, 4, precision=prec)
        ms = cuda_timed(torch, lambda: net.predict_device(x), 2, warm=1)
        out[prec] = {"ms": ms, "out_MPps": 1521 * 48 * 48 / ms / 1e3, "patches_per_s": 1521 / ms * 1e3}
    return out


def config_c5(torch, peaks, reps=5):
    """C5: bicubic x2/x3/x4 (fp32 + uint8), PSNR and fused PSNR+SSIM at 1K / 2K / 4K outputs, >= ~2 GB per launch, against the
    measured HBM copy peak.  Algorithmic bytes: bicubic fp32 (12 + 12/s^2) B per RGB output pixel, uint8
    (3 + 3/s^2); PSNR / PSNR+SSIM 24 B per pixel.  Every rank runs the sweep on its own images (weak scaling)."""
    from srb200 import ops
    rows = []
    g = torch.Generator(device="cuda").manual_seed(0)
    for (oh, ow) in ((1024, 1024), (2048, 2048), (2160, 3840)):
        for s in (2, 3, 4):
            ih, iw = -(-oh // s), -(-ow // s)
            for dt, per_px in (("f32", 12), ("u8", 3)):
                batch = max(1, int(2e9 // (oh * ow * per_px + ih * iw * per_px)))
                if dt == "f32":
                    x = torch.rand((batch, ih, iw, 3), device="cuda", generator=g)
                else:
                    x = torch.randint(0, 256, (batch, ih, iw, 3), device="cuda", dtype=torch.uint8, generator=g)
                ms = cuda_timed(torch, lambda: ops.bicubic(x, oh, ow), reps)
                alg = batch * (oh * ow + ih * iw) * per_px
                rows.append({"op": f"bicubic_{dt}_x{s}", "out": f"{ow}x{oh}", "batch": batch, "ms": ms, "GBps": alg / ms / 1e6,
                             "hbm_frac": alg / ms / 1e6 / peaks["hbm_gbs"], "out_MPps": batch * oh * ow / ms / 1e3})
                del x
        batch = max(1, int(2e9 // (oh * ow * 24)))
        a = torch.rand((batch, oh, ow, 3), device="cuda", generator=g)
        b = (a + 0.05 * torch.randn(a.shape, device="cuda", generator=g)).clamp_(0, 1)
        alg = batch * oh * ow * 24
        from srb200 import _capi as capi
        for name, fn in (("psnr_ssim_f32", lambda: ops.psnr_ssim(a, b)),
                         ("psnr_ssim_exact_window_f32", lambda: ops.psnr_ssim(a, b, window=capi.SSIM_TF_EXACT)),
                         ("psnr_f32", lambda: ops.psnr(a, b))):
            ms = cuda_timed(torch, fn, reps)
            row = {"op": name, "out": f"{ow}x{oh}", "batch": batch, "ms": ms, "GBps": alg / ms / 1e6,
                   "hbm_frac": alg / ms / 1e6 / peaks["hbm_gbs"], "out_MPps": batch * oh * ow / ms / 1e3}
            if name == "psnr_ssim_f32":
                # tensor-path kernel (mma.sync m16n8k16): 104 HMMAs of 4,096 FLOP per 352 map elements, against the HMMA issue
                # rate measured on this pool (tools/probes/hmma_rate.cu: 540 TFLOP/s); the reference-equivalent count is
                # ~600 FLOP per RGB pixel on the FP32 pipe
                row["kernel"] = "psnr_ssim_mma_kernel (fp16 window summing to 1, hi/lo-split data)"
                row["hmma_tflops"] = batch * oh * ow * 3 * (104.0 / 352.0) * 4096 / ms / 1e9
                row["hmma_frac_of_540"] = row["hmma_tflops"] / 540.0
                row["fp32_equivalent_tflops"] = batch * oh * ow * 600 / ms / 1e9
            elif name == "psnr_ssim_exact_window_f32":           # CUDA-core kernel with the float32 Gaussian: the FP32-pipe roofline
                row["kernel"] = "psnr_ssim_pair_kernel (float32 Gaussian, CUDA cores)"
                row["fp32_tflops"] = batch * oh * ow * 600 / ms / 1e9
                row["fp32_frac_of_74"] = row["fp32_tflops"] / 74.4
            rows.append(row)
        del a, b
    return rows


def summarize_c5(rows):
    def span(prefix):
        v = [r["hbm_frac"] for r in rows if r["op"].startswith(prefix)]
        return [round(min(v), 3), round(max(v), 3)] if v else None
    return {"bicubic_f32_hbm_frac": span("bicubic_f32"), "bicubic_u8_hbm_frac": span("bicubic_u8"),
            "psnr_hbm_frac": span("psnr_f32"), "psnr_ssim_hbm_frac": span("psnr_ssim_f32"),
            "psnr_ssim_GPps": [round(r["out_MPps"] / 1e3, 1) for r in rows if r["op"] == "psnr_ssim_f32"],
            "psnr_ssim_exact_window_GPps": [round(r["out_MPps"] / 1e3, 1) for r in rows if r["op"] == "psnr_ssim_exact_window_f32"]}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    from srb200 import distributed as D
    from srb200 import engine, ops, weights

    rank, world, local = D.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # opt-in: on the 4-GPU boxes measured (all GPUs on NUMA node 0) binding the ranks to the GPU's CPU set did not help
    numa_bound = D.bind_host_to_gpu(local) if (world > 1 and os.environ.get("SRB_NUMA_BIND")) else False
    lo, hi = D.shard_bounds(args.batch, rank, world)
    n_local = hi - lo
    tile, out_px = args.tile, (args.tile * 4) ** 2
    mb = min(args.micro_batch, max(n_local, 1))
    peaks = measured_peaks()
    want_headline = args.workload in ("all", "edsr")
    want_configs = args.workload in ("all", "configs")

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    line = {"metric": "sr_output_megapixels_per_sec", "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype,
            "data": "synthetic", "config": workload_config(args, mb)}
    ops.reset_launch_count()
    net = engine.EDSRNet(weights.edsr_weights(4), 4, 16, precision=args.dtype, trunk=args.trunk, upsampler=args.upsampler)
    net.max_device_batch = mb
    composed = net.upsampler == "composed"
    launches_per_forward = 34 if composed else 36          # tcgen05 launches between the hooks (every layer but the head)
    x_np = lr_tiles(lo, hi, tile)
    x_host = torch.from_numpy(x_np).pin_memory()
    x_dev = x_host.to(dev)

    # spans that contain only tcgen05 conv launches: [after head conv, after tail conv] of every micro-batch
    spans = []

    def hook(tag):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        spans.append((tag, ev))
    net.event_hook = hook

    def device_step():
        for i in range(0, n_local, mb):
            net.forward_device(x_dev[i:i + mb])

    for _ in range(args.warmup):
        device_step()
    barrier()
    spans.clear()
    launches0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            device_step()
        e1.record()
        barrier()
    launches = ops.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    tc_ms = sum(a[1].elapsed_time(b[1]) for a, b in zip(spans[0::2], spans[1::2]))
    n_tc_launches = args.steps * ((n_local + mb - 1) // mb) * launches_per_forward if n_local else 0
    net.event_hook = None

    # ---- end to end through the reference-facing API with host buffers ----
    # predict() overlaps the copies of micro-batch i-1 / i+1 with the kernels of micro-batch i, so the shard is cut into at
    # least eight pieces here (the last piece's device->host copy is the exposed part)
    e2e_mb = min(mb, max(16, -(-n_local // 8)))
    net.max_device_batch = e2e_mb
    out_pinned = torch.empty((n_local, tile * 4, tile * 4, 3), dtype=torch.float32).pin_memory()

    def e2e_time(fn, steps):
        if n_local:
            for _ in range(2):                                       # warm-up (pinned blocks, copy streams, allocator)
                fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            if n_local:
                fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        barrier()
        return dt
    # the stock call: pageable float32 in, fresh float32 array out (EDSR_model.py:274)
    e2e_ms = e2e_time(lambda: net.predict(x_np), args.steps)
    vsteps = max(2, min(args.steps, 3))
    variants_ms = {
        "pinned_in_out_fp32": e2e_time(lambda: net.predict(x_host.numpy(), out=out_pinned.numpy()), vsteps),
        "out_dtype_float16": e2e_time(lambda: net.predict(x_np, out_dtype=np.float16), vsteps),
        "out_dtype_uint8": e2e_time(lambda: net.predict(x_np, out_dtype=np.uint8), vsteps),
    }
    del out_pinned
    link = host_link_gbs(torch, dev)
    barrier()

    # per-launch CUDA-event times of one forward (separate, untimed pass on rank 0): which bound each layer group sits at
    layer_groups = layer_group_rooflines(net, x_dev[:mb], torch, ops) if (rank == 0 and n_local and want_headline) else None

    # ---- the same network with the up-sampling tail run layer by layer (three launches, as the reference builds it) ----
    other_ms = 0.0
    if want_headline and n_local:
        net_other = engine.EDSRNet(weights.edsr_weights(4), 4, 16, precision=args.dtype, trunk=args.trunk,
                                   upsampler="layered" if composed else "composed")

        def other_step():
            for i in range(0, n_local, mb):
                net_other.forward_device(x_dev[i:i + mb])
        other_ms = cuda_timed(torch, other_step, 2, warm=1)
        del net_other
    barrier()

    # ---- the bf16 arm (the dtype BASELINE configs[2] names) ----
    alt = None
    net_alt = None
    if want_headline and n_local:
        alt_name = "bf16" if args.dtype != "bf16" else "fp16"
        net_alt = engine.EDSRNet(weights.edsr_weights(4), 4, 16, precision=alt_name, trunk=args.trunk, upsampler=args.upsampler)

        def alt_step():
            for i in range(0, n_local, mb):
                net_alt.forward_device(x_dev[i:i + mb])
        alt_ms = cuda_timed(torch, alt_step, 2, warm=1)
        alt = {"dtype": alt_name, "ms_per_step_this_rank": alt_ms}
    barrier()

    # ---- the one collective of the path ----
    coll = collective_block(torch, D, net, rank, world, tile) if want_headline else None
    barrier()

    vals = [ms, e2e_ms, tc_ms, alt["ms_per_step_this_rank"] if alt else 0.0, link["d2h"], link["h2d"], other_ms] + list(variants_ms.values())
    t_max = torch.tensor(vals, dtype=torch.float64, device=dev)
    t_sum = t_max.clone()
    if world > 1:
        torch.distributed.all_reduce(t_max, op=torch.distributed.ReduceOp.MAX)
        torch.distributed.all_reduce(t_sum, op=torch.distributed.ReduceOp.SUM)
    ms, e2e_ms, tc_ms_max, alt_ms = (float(v) for v in t_max.tolist()[:4])
    link_sum = [float(v) for v in t_sum.tolist()[4:6]]
    other_ms = float(t_max.tolist()[6])
    var_ms = dict(zip(variants_ms, (float(v) for v in t_max.tolist()[7:])))

    # ---- the other BASELINE configs ----
    configs = None
    if want_configs:
        configs = {}
        c5_rows = config_c5(torch, peaks)                          # every rank (weak scaling: its own images)
        agg = torch.tensor([r["GBps"] for r in c5_rows], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(agg, op=torch.distributed.ReduceOp.SUM)
        if rank == 0:
            for r, a in zip(c5_rows, agg.tolist()):
                r["GBps_all_gpus"] = a
            configs["c5"] = {"workload": "bicubic x2/x3/x4 (fp32 + uint8) + PSNR + fused PSNR/SSIM at 1K / 2K / 4K outputs, >= 2 GB per "
                                         "launch, per GPU (BASELINE configs[4]); rank 0's rows, GBps_all_gpus = sum over ranks",
                             "bound": "hbm", "peak_gbs": peaks["hbm_gbs"], "summary": summarize_c5(c5_rows), "rows": c5_rows}
            with ClockSampler(local) as clk2:
                configs["c1"] = config_c1(torch, peaks)
                configs["c2"] = config_c2(torch, peaks)
                configs["c4"] = config_c4(torch, peaks, args.c4_batch)
                configs["esrgan"] = config_esrgan(torch, peaks)
            configs["clocks"] = clk2.result
        barrier()

    if rank == 0:
        total_mp = args.batch * out_px / 1e6
        value = total_mp * args.steps / (ms / 1e3)
        e2e = total_mp * args.steps / (e2e_ms / 1e3)
        # FLOPs the tcgen05 launches EXECUTE: with the composed tail that is the 33 body convs + the 5 x 5 conv, not the
        # reference network's layer-by-layer count (reported next to it as reference_equivalent_tflops)
        tc_flop_px = (EDSR_BODY_FLOP_PER_LR_PX + EDSR_COMPOSED_TAIL_FLOP_PER_LR_PX) if composed else (EDSR_FLOP_PER_LR_PX - EDSR_HEAD_FLOP_PER_LR_PX)
        tc_flops = tc_flop_px * tile * tile * n_local * args.steps
        achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0
        h2d_b, d2h_b = args.batch * tile * tile * 3 * 4, args.batch * out_px * 3 * 4
        traffic = measured_traffic()
        traffic = traffic.get("composed" if composed else "layered") if traffic else None
        line.update({
            "value": value, "ms_per_step": ms / args.steps,
            "e2e": {"value": e2e, "unit": "MP/s", "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
                    "ms_per_step": e2e_ms / args.steps, "api": "EDSRNet.predict(pageable host NHWC float32) -> fresh float32 array (stock signature)",
                    "micro_batch": e2e_mb, "host_bound_to_gpu_numa_node": bool(numa_bound),
                    "variants": {k: {"value": total_mp * vsteps / (v / 1e3), "ms_per_step": v / vsteps,
                                     "d2h_bytes_per_step": d2h_b // {"pinned_in_out_fp32": 1, "out_dtype_float16": 2, "out_dtype_uint8": 4}[k]}
                                 for k, v in var_ms.items()},
                    "host_ceiling_gbs": {"d2h_all_ranks": link_sum[0], "h2d_all_ranks": link_sum[1],
                                         "note": "pinned host <-> device copy rate of every rank measured alone and at the same time, summed"},
                    "d2h_gbs_achieved": d2h_b / (e2e_ms / args.steps) / 1e6},
            "gpu_launches": launches * world,
            "clocks": clk.result,
            "roofline": {"bound": "tensor", "kernel": f"conv3x3_tc_kernel + conv3x3_fold_kernel{' + upsample5_fold_kernel' if composed else ''} "
                                                      f"(tcgen05 implicit GEMM, all {launches_per_forward} launches of a forward)",
                         "achieved": achieved, "peak": peaks["tflops"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                         "traffic": (traffic["bytes_per_launch_at_32_tiles"] * (mb / 32.0)) if traffic else None,
                         "traffic_unit": "bytes per launch (ncu dram read + write, mean over the launches of a forward)",
                         "traffic_source": traffic.get("source") if traffic else None,
                         "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)",
                         "flop_per_launch": tc_flops / max(n_tc_launches, 1),
                         "avg_launch_ms": tc_ms / max(n_tc_launches, 1), "launches": n_tc_launches,
                         "flops_counted": "executed by the launches (composed tail = 153,600 FLOP per LR pixel instead of the layered "
                                          "1,529,856)" if composed else "the reference network's layers",
                         "reference_equivalent_tflops": EDSR_FLOP_PER_LR_PX * tile * tile * args.batch * args.steps / (ms / 1e3) / 1e12},
        })
        if other_ms:
            line["composed_upsampler" if not composed else "layered_upsampler"] = {
                "value": total_mp / (other_ms / 1e3), "unit": "MP/s", "ms_per_step": other_ms,
                "note": "same network and trunk, device-resident, 2 steps after 1 warm-up" + (
                    "; tail as three launches (64->256 + d2s, 64->256 + d2s, 64->3), as the reference builds it" if composed else "")}
        if layer_groups:
            for g in layer_groups:
                g["tensor_frac"] = g["tflops"] / peaks["tflops"]           # (sustained peak: the pass runs at the power cap)
                g["hbm_frac"] = g["gbs_min"] / peaks["hbm_gbs"]
                g["bound"] = "hbm" if g["hbm_frac"] > g["tensor_frac"] else "tensor"
            line["roofline"]["by_layer_group"] = layer_groups
            line["roofline"]["by_layer_group_note"] = ("median CUDA-event time per launch over 5 forwards of one micro-batch, right after the "
                                                        "timed loop (same power-capped clocks, outside the timed region); gbs_min = algorithmic "
                                                        "bytes (16-bit activations, pair8 trunk = 3 B per channel, fp32 RGB in / out) / time; "
                                                        "fractions against the sustained bf16 peak and the copy peak")
        if want_headline and n_local:
            par = parity_block(torch, net, net_alt, tile)
            line["parity"] = dict(par["main"], tiles=2, tile=f"{tile}x{tile} LR -> {tile * 4}x{tile * 4}", dtype=args.dtype,
                                  oracle="fp32 torch-CPU restatement of EDSR_model.py:96-125 (oracle/convnets.py), same weights and inputs",
                                  tolerance={"max_abs": 2e-2, "psnr_db": 0.01, "ssim": 1e-4})
            if alt:
                alt.update({"value": total_mp / (alt_ms / 1e3), "unit": "MP/s", "ms_per_step": alt_ms,
                            "max_abs": par["alt"]["max_abs"], "psnr_delta_db": par["alt"]["psnr_delta_db"],
                            "ssim_delta": par["alt"]["ssim_delta"],
                            "note": "bf16 operands: same tcgen05 kind::f16 rate; 8-bit mantissa cannot meet 2e-2 on he_normal EDSR "
                                    "(weights-only rounding = 5.8e-2, DESIGN.md section 2), hence fp16 is the default"})
                alt.pop("ms_per_step_this_rank", None)
                line["alt_dtype"] = alt
        if coll:
            line["collective"] = coll
        if configs:
            line["configs"] = configs
        if args.gpus == 1 and not args.no_cpu and want_headline:
            n_cpu = 24
            mp_s, cores, secs = cpu_edsr_mp_per_s(n_cpu, tile)
            line["cpu_baseline"] = {"value": mp_s, "unit": "MP/s", "cores": cores, "kind": "port",
                                    "sample": f"first {n_cpu} of the {args.batch} LR tiles, fp32 oracle port (torch CPU), {secs:.1f} s"}
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


_RESULT_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under torchrun), so
    file descriptor 1 is pointed at stderr for the whole run and the result line is written to the saved descriptor."""
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "edsr", "configs"],
                    help="all: the EDSR headline (BASELINE configs[2]) with its parity / bf16 / collective blocks plus the `configs` "
                         "block (c1, c2, c4, c5); edsr: without the `configs` block; configs: the headline throughput and the "
                         "`configs` block only")
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--c4-batch", type=int, default=1024)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--trunk", default="pair8", choices=["pair8", "fp32", "pair", "half"],
                    help="residual trunk storage: 16-bit + e5m2 rounding-error pair (default), fp32, compensated 16-bit pair, or plain 16-bit")
    ap.add_argument("--upsampler", default="composed", choices=["composed", "layered"],
                    help="EDSR up-sampling tail: one composed 5x5 launch (default) or layer by layer")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
