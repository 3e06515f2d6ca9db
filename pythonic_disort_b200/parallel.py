"""Multi-GPU use of the column-batched solver: columns are independent, so a
batch is sharded across ranks by contiguous column ranges and every rank runs
the unchanged single-GPU path on its shard; there is NO collective on the hot
path.  The only communication offered is an optional all-gather of output
arrays (fluxes are [B, L+1, ...] FP64: ~1.5 KB per column) for consumers that
want the whole ensemble on every rank.  One process per GPU (torchrun); NCCL
on GPUs, gloo in the CPU tests."""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(B, rank=None, world=None):
    """[lo, hi) of the columns owned by `rank`: contiguous, sizes differ by at most one."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(B, args, kwargs, rank=None, world=None):
    """This rank's columns of a batched pydisort() call.  Which inputs carry the column axis is decided from each
    argument's rank (api.carries_batch_axis), not from a dimension that happens to equal B: a shared
    ``Leg_coeffs_all`` [L, NLeg_all] with L == B, or a shared ``b_pos`` [N] with N == B, is passed through whole."""
    from .api import slice_columns
    lo, hi = shard_range(B, rank, world)
    out_args, out_kw = slice_columns(args, kwargs, lo, hi, B)
    return out_args, out_kw, (lo, hi)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores NVML names as local to GPU `device_index`, so that the pinned staging
    buffers it allocates afterwards sit on the NUMA node next to that GPU's PCIe root (one rank per GPU: eight ranks
    copying through one node's memory was what held the 8-GPU end-to-end efficiency at 0.85).  Returns the cores, or
    None if NVML / sched_setaffinity are unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cores = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < ncpu]
        if cores:
            os.sched_setaffinity(0, cores)
            return cores
    except Exception:  # noqa: BLE001 -- affinity is an optimisation, never a requirement
        return None
    return None


def all_gather_columns(local, B):
    """Concatenate per-rank output arrays along the column axis on every rank
    (ranks may own shards of different size).  `local` is a tensor or ndarray
    whose first axis is this rank's columns."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    was_numpy = isinstance(local, np.ndarray)
    t = torch.as_tensor(local)
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda()
    world = dist.get_world_size()
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    pad = max(sizes)
    buf = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    buf[: t.shape[0]] = t
    out = torch.empty((world * pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, buf.contiguous())
    full = torch.cat([out[r * pad: r * pad + sizes[r]] for r in range(world)], dim=0)
    return full.cpu().numpy() if was_numpy else full
