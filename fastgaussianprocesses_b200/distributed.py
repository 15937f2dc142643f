"""Multi-GPU partitioning of the paths that shard: one process per GPU, torch.distributed for the
plumbing.  Posterior evaluation shards the test points (rows are independent); batched fits are independent objects per
rank.  The ONLY collective is the all_gather that collects results (NCCL over NVLink on the GPU box; gloo in CPU tests)."""
import torch


def shard_bounds(m, world, rank):
    """Contiguous row range [lo, hi) of rank `rank` when m rows are split over `world` ranks (first m % world ranks get one extra)."""
    base, extra = divmod(int(m), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_rows(fn, x, group=None):
    """Apply `fn` (rows (k,d) -> values (..., k)) to this rank's contiguous shard of x and all_gather the results so that
    every rank returns the full (..., m) tensor.  fn is e.g. `gp.post_mean` or `gp.post_var`."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return fn(x)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    m = x.shape[0]
    lo, hi = shard_bounds(m, world, rank)
    local = fn(x[lo:hi])
    kmax = -(-m // world)
    lead = tuple(local.shape[:-1])
    pad = torch.zeros(lead + (kmax,), dtype=local.dtype, device=local.device)
    pad[..., :hi - lo] = local
    flat = torch.empty(world * pad.numel(), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(flat, pad.reshape(-1), group=group)
    out = flat.reshape((world,) + lead + (kmax,))
    parts = []
    for r in range(world):
        a, b = shard_bounds(m, world, r)
        parts.append(out[r][..., :b - a])
    return torch.cat(parts, -1)


def post_mean_sharded(gp, x, group=None):
    return sharded_rows(gp.post_mean, x, group)


def post_var_sharded(gp, x, group=None):
    return sharded_rows(gp.post_var, x, group)


def gather_fit_results(values, group=None):
    """Collect a small per-rank tensor of fit results (hyperparameters, loss) from independent fits: (world, ...)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return values[None]
    world = dist.get_world_size(group)
    flat = torch.empty(world * values.numel(), dtype=values.dtype, device=values.device)
    dist.all_gather_into_tensor(flat, values.reshape(-1).contiguous(), group=group)
    return flat.reshape((world,) + tuple(values.shape))
