"""Multi-GPU plumbing: the frame-tracking path does not shard ("replicas only") — units of work
are independent frame pairs / sequences, split in contiguous blocks across ranks, no data-path collective.
torch.distributed is used only for the barrier and the max-over-ranks of the timed region."""


def shard_range(n_units, rank, world):
    """Contiguous block [lo, hi) of `n_units` independent frame pairs owned by `rank` (sizes differ by <= 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def replica_seed(base_seed, rank):
    """Each replica tracks its own synthetic sequence (weak scaling: per-GPU work is fixed)."""
    return base_seed + 1000 * rank


def max_over_ranks(value, dist=None, device=None):
    """Max of a host scalar over all ranks (the job time is the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank, ms_per_step_max, world):
    """Whole-job units/s: every rank processed `units_per_rank` in the (max-over-ranks) step time."""
    return world * units_per_rank / (ms_per_step_max * 1e-3)
