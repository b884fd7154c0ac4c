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


def gather_rows(local: torch.Tensor, total: int, group=None, dim: int = 0) -> torch.Tensor:
    """All-gather variable-sized row blocks back into `[total, ...]` along `dim` (identical on every rank).
    `dim=1` gathers the stacked pair output `[2, b, ...]` of a never-fused branch run (ddpm.py:965-970)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    if dim != 0:
        return gather_rows(local.movedim(dim, 0).contiguous(), total, group).movedim(0, dim).contiguous()
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


def global_noise_tape(shape, steps, device, seed=10):
    """The reference's draws for the GLOBAL batch (`torch.manual_seed(10)`, x_T, then one draw per step; ddpm.py:934-935, 852),
    identical on every rank: slicing it per rank makes a sharded run reproduce the single-GPU run row for row."""
    g = torch.Generator(device=device).manual_seed(seed)
    tape = torch.empty((steps,) + tuple(shape), device=device)
    for i in range(steps):
        tape[i] = torch.randn(shape, device=device, generator=g)
    return tape


def sample_sharded(sample_fn, cond, mask, noise=None, group=None, steps=None, pair=False):
    """Run `sample_fn(cond_rows, mask_rows, noise_rows) -> [rows, ...]` on this rank's rows of the global batch and return the
    gathered `[B, ...]` result (`pair=True`: the stacked `[2, B, ...]` output of a never-fused run, gathered along dim 1).
    `noise` is the *global* tape `[T, B, ...]`, sliced per rank so results do not depend on the world size; when it is None the
    global tape is drawn here (every rank draws the same `steps` x B tape from seed 10 and keeps its rows) -- letting every rank
    seed its own local tape would give all shards the same noise rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = cond.shape[0]
    if noise is None:
        if steps is None:
            raise ValueError("sample_sharded needs the global noise tape or `steps` to draw it")
        noise = global_noise_tape(tuple(cond.shape), steps, cond.device)
    c, m = shard_rows(cond, rank, world), shard_rows(mask, rank, world)
    z = shard_rows(noise, rank, world, dim=1)
    return gather_rows(sample_fn(c, m, z), B, group, dim=1 if pair else 0)
