"""Multi-GPU data parallelism over queries.

The reference's only parallelism is `mp.Pool.starmap` over queries with fork-inherited read-only state
(run_apples.py:20,94-102).  Here: one process per GPU, queries split into contiguous blocks of ceil(Q / G), the packed
reference + tree replicated on every GPU, no data-path collective; the single collective is the final all-gather of
the placements (edge i32, error f64, distal f64, pendant f64, status i32 = 32 B per query), after which every rank
holds the results of all queries in input order -- so the N-GPU output is byte-identical to the 1-GPU output.

torch.distributed is plumbing here: `nccl` on GPUs (NVLink / NVSwitch), `gloo` in the CPU tests.
"""
import numpy as np


def shard_bounds(n_items, world):
    """Contiguous blocks of ceil(n / world): [(begin, end)] per rank (trailing ranks may be empty)."""
    per = -(-int(n_items) // int(world)) if n_items else 0
    return [(min(r * per, n_items), min((r + 1) * per, n_items)) for r in range(world)]


def gather_placements(local, n_total, device=None):
    """local: tuple (edge i32, error f64, distal f64, pendant f64, status i32) of this rank's shard as numpy arrays or
    torch tensors.  Returns the same tuple for all `n_total` queries in input order, as numpy arrays.

    Shards are padded to the common block size ceil(n_total / world) so one all_gather_into_tensor per array suffices.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tuple(np.asarray(a.cpu() if hasattr(a, 'cpu') else a) for a in local)
    world = dist.get_world_size()
    bounds = shard_bounds(n_total, world)
    per = bounds[0][1] - bounds[0][0]
    if device is None:
        device = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    out = []
    for a in local:
        t = a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
        t = t.to(device)
        if t.numel() < per:
            t = torch.cat([t, torch.zeros(per - t.numel(), dtype=t.dtype, device=device)])
        full = torch.empty(per * world, dtype=t.dtype, device=device)
        dist.all_gather_into_tensor(full, t.contiguous())
        out.append(full[:n_total].cpu().numpy())
    return tuple(out)


def place_sharded(place_fn, n_total):
    """Run `place_fn(begin, end) -> result tuple` on this rank's block and gather every rank's results."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = 0, 1
    b, e = shard_bounds(n_total, world)[rank]
    local = place_fn(b, e)
    return gather_placements(local, n_total)
