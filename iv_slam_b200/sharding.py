"""Frame-parallel sharding helpers for the multi-GPU path.

Frames are independent units: rank g of G owns the contiguous frame range [g*T/G, (g+1)*T/G) (BASELINE config C3) and
runs the whole hot path on it; there is NO collective on the data path.  torch.distributed is plumbing only: a barrier
around the timed region, a MAX-reduce of the per-rank device time, and (optionally) a gather of the small per-frame
result counts to rank 0.  Works with the gloo backend on CPU (tests) and nccl on GPUs (bench.py).
"""
import numpy as np


def frame_range(rank, world, total):
    """Contiguous frame range [start, end) owned by `rank`."""
    assert 0 <= rank < world and total >= 0
    return (rank * total) // world, ((rank + 1) * total) // world


def reduce_max(value, device="cpu"):
    """Max over ranks of a scalar (the per-step time); identity without a process group."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(local_counts, total, rank, world, device="cpu"):
    """Gathers per-frame integer results (e.g. keypoints per frame) of every rank's frame range into one array of
    length `total` on every rank, in global frame order."""
    import torch
    import torch.distributed as dist
    local_counts = np.asarray(local_counts, np.int64)
    s, e = frame_range(rank, world, total)
    assert local_counts.size == e - s
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return local_counts.copy()
    longest = max(frame_range(r, world, total)[1] - frame_range(r, world, total)[0] for r in range(world))
    buf = torch.zeros(longest, dtype=torch.int64, device=device)
    buf[:local_counts.size] = torch.from_numpy(local_counts).to(device)
    outs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    full = np.zeros(total, np.int64)
    for r in range(world):
        rs, re = frame_range(r, world, total)
        full[rs:re] = outs[r][:re - rs].cpu().numpy()
    return full


def bind_to_gpu_numa_node(device_index):
    """Restricts this process to the CPUs that are local to GPU `device_index` (NVML's CPU affinity mask), so that the
    pinned host buffers it allocates afterwards live on that GPU's NUMA node and the H2D / D2H copies of different
    ranks do not share one socket's memory controllers and inter-socket links.  Returns the CPU list, or None when NVML
    or the affinity call is unavailable (single-socket boxes: nothing to do)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus and len(cpus) < len(allowed):
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        return None
    return None
