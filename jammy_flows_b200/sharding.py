"""Row sharding of the inference path over the GPUs of one box.

Rows are independent, so log_pdf / sampling shard by contiguous row blocks with NO data-path collective: every rank
(one process per GPU, `torch.distributed`) owns rows [lo, hi) of the global batch, parameters are replicated, and the
outputs stay sharded unless the caller gathers them.  The only collectives are the optional result gather and the
max-over-ranks reduction of a timing -- both plumbing, both usable with NCCL (GPU tensors) or gloo (CPU tensors, tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_rows, rank, world):
    """Contiguous block [lo, hi) of `n_rows` global rows owned by `rank`; the first n_rows % world ranks get one more."""
    if world < 1 or not (0 <= rank < world) or n_rows < 0:
        raise ValueError("bad shard request: n_rows=%d rank=%d world=%d" % (n_rows, rank, world))
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(t, rank=None, world=None):
    """This rank's row block of a global [B, ...] tensor (a view, no copy)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def sample_seed_offset(n_rows, rank, world):
    """Global index of this rank's first row: the Philox offset that makes the union of the per-rank sample sets
    independent of the number of ranks."""
    return shard_range(n_rows, rank, world)[0]


def shard_base_normals(n_rows, dim, seed, rank=None, world=None, dtype=torch.float64, device="cuda"):
    """This rank's rows [lo, hi) of the global [n_rows, dim] base-space normal draw for `seed` (device Philox,
    `jf_normal_rows`): the union over the ranks equals the single-process draw bit for bit, for any number of ranks."""
    from . import engine
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_range(n_rows, rank, world)
    return engine.normal_rows(hi - lo, dim, seed, first_row=lo, dtype=dtype, device=device)


def gather_rows(local, n_rows, group=None):
    """All-gather ragged row blocks back into the global [B, ...] tensor on every rank (inverse of `shard_rows`)."""
    world = dist.get_world_size(group)
    sizes = [shard_range(n_rows, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[:hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def max_over_ranks(value, device="cpu", group=None):
    """Max of a python float over all ranks (device-side timings are reported as the max over ranks)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def allreduce_gradients(module, group=None, average=True):
    """Data-parallel training step: ONE all-reduce over a flat bucket of every parameter gradient
    (cfg5: 422 410 fp32 values = 1.7 MB, latency bound on NVLink/NVSwitch), then scatter back.  NCCL for GPU tensors,
    gloo for CPU tensors (tests)."""
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if len(grads) == 0:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return flat.numel()
