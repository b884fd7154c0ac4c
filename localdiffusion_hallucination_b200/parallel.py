"""Batch-sharded sampling over the GPUs of one box.

Every op of the sampler is per-sample, so ranks take contiguous row blocks of `cond`, `mask` and
of the noise tape, run the whole T-step loop with no per-step traffic, and exchange results with a
single all-gather at the end (NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous, balanced row block of `rank`; the first `total % world` ranks get one extra row."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(t, rank, world, dim=0):
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo)


def gather_rows(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather variable-sized row blocks back into `[total, ...]` (identical on every rank)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    base, extra = divmod(total, world)
    if extra == 0:
        out = torch.empty((total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    pad = base + 1
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    return torch.cat([p[: shard_bounds(total, r, world)[1] - shard_bounds(total, r, world)[0]] for r, p in enumerate(parts)])


def sample_sharded(sample_fn, cond, mask, noise=None, group=None):
    """Run `sample_fn(cond_rows, mask_rows, noise_rows) -> [rows, ...]` on this rank's rows of the
    global batch and return the gathered `[B, ...]` result.  `noise` is the *global* tape
    `[T, B, ...]`, sliced per rank so results do not depend on the world size."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = cond.shape[0]
    c, m = shard_rows(cond, rank, world), shard_rows(mask, rank, world)
    z = shard_rows(noise, rank, world, dim=1)
    return gather_rows(sample_fn(c, m, z), B, group)
