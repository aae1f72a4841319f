"""Batch-sharded data-parallel inference: one process per GPU, no data-path collective; the only
exchange is one all-reduce of (sum psnr, sum ssim, count, sum mse) float64 per evaluation
(SURVEY.md section 8e).  The reference has no multi-GPU code to mirror.
"""
from __future__ import annotations

import os


def is_initialized():
    try:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def rank_world():
    if is_initialized():
        import torch.distributed as dist
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK / WORLD_SIZE / MASTER_* variables.
    Returns (rank, world, local_rank).  A no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not is_initialized():
        import torch
        import torch.distributed as dist
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # NCCL writes its version / debug lines to stdout by default; stdout is where bench.py's JSON line goes
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def bind_host_to_gpu(local_rank):
    """Pin this process to the CPUs NVML reports as local to its GPU (``nvmlDeviceSetCpuAffinity``), so that the pinned
    staging buffers it allocates afterwards are first-touched on the GPU's NUMA node.  With one process per GPU on a
    two-socket host, host<->device copies otherwise cross the socket interconnect for about half of the ranks.
    Returns True when the affinity was set; never raises (the binding is an optimisation)."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


def shard_bounds(n, rank=None, world=None):
    """Contiguous split of n items: rank r owns [lo, hi).  Earlier ranks take the remainder."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sums(sums, group=None):
    """Sum a small float64 tensor over all ranks (NCCL on GPU tensors, gloo on CPU tensors)."""
    if not is_initialized():
        return sums
    import torch.distributed as dist
    out = sums.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def allgather_vector(v, group=None):
    """Gather equally-sized per-image vectors from every rank, in rank order."""
    if not is_initialized():
        return v
    import torch
    import torch.distributed as dist
    parts = [torch.empty_like(v) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, v, group=group)
    return torch.cat(parts, 0)


def means_from_sums(sums):
    """(sum psnr, sum ssim, count, sum mse) -> [loss, psnr, ssim] as Keras evaluate reports them."""
    s = [float(x) for x in sums]
    cnt = max(s[2], 1.0)
    return [s[3] / cnt, s[0] / cnt, s[1] / cnt]
