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
    """Slice every input that carries the leading batch axis B down to this rank's columns."""
    lo, hi = shard_range(B, rank, world)

    def cut(x):
        if isinstance(x, (np.ndarray, torch.Tensor)) and x.ndim >= 1 and x.shape[0] == B:
            return x[lo:hi]
        return x

    out_kw = {}
    for k, v in kwargs.items():
        out_kw[k] = [cut(m) for m in v] if k == "BDRF_Fourier_modes" else cut(v)
    return tuple(cut(a) for a in args), out_kw, (lo, hi)


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
