#!/usr/bin/env python
"""bench.py - SR output megapixels/s of the B200 hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload edsr|espcn|bicubic|metrics]

Default workload (BASELINE configs[2]): EDSR-baseline x4 (16 res-blocks, 64 channels), 16-bit operands /
fp32 accumulation, batch 512 LR tiles of 192x192x3 -> 768x768x3, sharded contiguously over N GPUs (one
process per GPU, no data-path collective).  A step is one pass of the network over the whole batch.
Prints ONE JSON line on rank 0.

* ``value``: whole-job throughput with the LR batch already resident in HBM (CUDA events, max over ranks).
* ``e2e``: the same metric through the reference-facing call ``EDSR.model.predict(host array)`` with pinned
  host buffers: H2D of the LR batch and D2H of the SR batch are inside the timed region (pipelined against
  compute on copy streams).
* ``roofline``: the tcgen05 conv kernel (every layer but the 3-channel head): algorithmic FLOPs / CUDA-event
  time of the spans that contain only that kernel, against the measured sustained bf16 peak.
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
ESPCN_FLOP_PER_LR_PX = 2 * 37_056


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


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
        import threading  # noqa: F401
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
    hook, net.event_hook = net.event_hook, None
    ops.conv2d = timed
    try:
        per_rep = []
        for _ in range(5):
            records.clear()
            net.forward_device(x)
            torch.cuda.synchronize()
            per_rep.append([(w, npx, kw, e0.elapsed_time(e1)) for (w, npx, kw, e0, e1) in records])
    finally:
        ops.conv2d = orig
        net.event_hook = hook
    groups = {}
    for i, (w, npx, kw, _) in enumerate(per_rep[0]):
        ms = statistics.median(rep[i][3] for rep in per_rep)
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
            "l2": "inputs larger than L2 (LR batch + activations stream through HBM every step); no flush needed"}


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
    # (e2e 7,352 MP/s bound vs 8,090 unbound); the e2e line at N >= 4 is limited by the host side of the 3.6 GB/step read-back
    numa_bound = D.bind_host_to_gpu(local) if (world > 1 and os.environ.get("SRB_NUMA_BIND")) else False
    lo, hi = D.shard_bounds(args.batch, rank, world)
    n_local = hi - lo
    tile, out_px = args.tile, (args.tile * 4) ** 2
    mb = min(args.micro_batch, max(n_local, 1))

    net = engine.EDSRNet(weights.edsr_weights(4), 4, 16, precision=args.dtype, trunk=args.trunk)
    net.max_device_batch = mb
    x_host = torch.from_numpy(lr_tiles(lo, hi, tile)).pin_memory()
    x_dev = x_host.to(dev)
    out_host = torch.empty((n_local, tile * 4, tile * 4, 3), dtype=torch.float32).pin_memory()

    # spans that contain only tcgen05 conv launches: [after head conv, after tail conv] of every micro-batch
    spans = []

    def hook(tag):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        spans.append((tag, ev))
    net.event_hook = hook

    def device_step():
        outs = []
        for i in range(0, n_local, mb):
            outs.append(net.forward_device(x_dev[i:i + mb]))
        return outs

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_step()
    barrier()
    spans.clear()
    ops.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            device_step()
        e1.record()
        barrier()
    launches = ops.launch_count()
    ms = e0.elapsed_time(e1)
    tc_ms = sum(a[1].elapsed_time(b[1]) for a, b in zip(spans[0::2], spans[1::2]))
    n_tc_launches = args.steps * ((n_local + mb - 1) // mb) * 36 if n_local else 0
    net.event_hook = None

    # end to end through the reference-facing API with host buffers.  predict() overlaps the copies of micro-batch i-1 / i+1
    # with the kernels of micro-batch i, so the shard is cut into at least eight pieces here (the last piece's
    # device->host copy is the exposed part); the device-resident run above prefers fewer, larger launch sequences
    e2e_mb = min(mb, max(16, -(-n_local // 8)))
    net.max_device_batch = e2e_mb
    e2e_ms = None
    if n_local:
        for _ in range(2):                                           # warm-up (pinned staging, copy streams, allocator)
            net.predict(x_host.numpy(), out=out_host.numpy())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if n_local:
            net.predict(x_host.numpy(), out=out_host.numpy())
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    barrier()

    # per-launch CUDA-event times of one forward (separate, untimed pass on rank 0): which bound each layer group sits at
    layer_groups = layer_group_rooflines(net, x_dev[:mb], torch, ops) if (rank == 0 and n_local) else None

    t = torch.tensor([ms, e2e_ms, tc_ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms, tc_ms_max = (float(v) for v in t.tolist())

    if rank == 0:
        peaks = measured_peaks()
        total_mp = args.batch * out_px / 1e6
        value = total_mp * args.steps / (ms / 1e3)
        e2e = total_mp * args.steps / (e2e_ms / 1e3)
        tc_flops = (EDSR_FLOP_PER_LR_PX - EDSR_HEAD_FLOP_PER_LR_PX) * tile * tile * n_local * args.steps
        achieved = tc_flops / (tc_ms / 1e3) / 1e12 if tc_ms else 0.0
        line = {
            "metric": "sr_output_megapixels_per_sec", "value": value, "unit": "MP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(args, mb),
            "e2e": {"value": e2e, "unit": "MP/s", "h2d_bytes_per_step": args.batch * tile * tile * 3 * 4,
                    "d2h_bytes_per_step": args.batch * out_px * 3 * 4, "ms_per_step": e2e_ms / args.steps,
                    "api": "EDSRNet.predict(host NHWC float32, out=pinned host)", "micro_batch": e2e_mb,
                    "host_bound_to_gpu_numa_node": bool(numa_bound)},
            "gpu_launches": launches * world,
            "clocks": clk.result,
            # traffic: dram__bytes_read + dram__bytes_write per launch, launch-weighted mean over the 36 tcgen05 launches of one
            # 32-tile forward (ncu dram__bytes_* per launch, profiles/r01_launches_final.md): 16 x 0.563 (res-block 2nd conv) +
            # 17 x 0.256 (1st conv, body end) + 2 x 1.830 (up-sampling) + 2.635 (RGB tail) = 19.64 GB / 36
            "roofline": {"bound": "tensor", "kernel": "conv3x3_tc_kernel + conv3x3_fold_kernel (tcgen05 implicit GEMM, all 36 launches of a forward)",
                         "achieved": achieved, "peak": peaks["tflops"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["tflops"], "traffic": 0.546e9 * (mb / 32.0),
                         "traffic_unit": "bytes per launch (ncu dram read + write, mean over the launches of a forward)",
                         "peak_source": f"{peaks['source']} sustained bf16 (MEASURED_PEAKS.json)",
                         "flop_per_launch": tc_flops / max(n_tc_launches, 1),
                         "avg_launch_ms": tc_ms / max(n_tc_launches, 1), "launches": n_tc_launches,
                         "whole_net_tflops": EDSR_FLOP_PER_LR_PX * tile * tile * args.batch * args.steps / (ms / 1e3) / 1e12},
        }
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
        if args.gpus == 1 and not args.no_cpu:
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
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--tile", type=int, default=192)
    ap.add_argument("--micro-batch", type=int, default=32)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16"])
    ap.add_argument("--trunk", default="pair8", choices=["pair8", "fp32", "pair", "half"],
                    help="residual trunk storage: 16-bit + e5m2 rounding-error pair (default), fp32, compensated 16-bit pair, or plain 16-bit")
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
