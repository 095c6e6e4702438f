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


class gpu_local_cpus:
    """Context manager: runs its body with the calling thread bound to the CPUs NVML reports as local to GPU
    `gpu_index` (same socket / NUMA node), then restores the previous affinity.  Host buffers pinned inside it are
    placed on the GPU's own node, so N replicas each pull their frames over their own root complex instead of all
    reading one socket's memory across the inter-socket link (measured: the host-buffer pass at 4 GPUs).
    A no-op when NVML or sched_setaffinity is unavailable; `.bound` says whether the binding took."""

    def __init__(self, gpu_index):
        self.gpu_index, self.bound, self._prev = gpu_index, False, None

    def __enter__(self):
        import os
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu_index])
                except Exception:
                    idx = self.gpu_index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
            self._prev = os.sched_getaffinity(0)
            cpus &= self._prev
            if cpus:
                os.sched_setaffinity(0, cpus)
                self.bound = True
                self.cpus = sorted(cpus)
        except Exception:
            self.bound = False
        return self

    def __exit__(self, *exc):
        import os
        if self.bound and self._prev is not None:
            try:
                os.sched_setaffinity(0, self._prev)
            except Exception:
                pass
        return False
