"""Spin sharding over the GPUs of one box.

The reference has no multi-GPU path (`-d` picks ONE device; its kernel comment suggests separate processes with
different seeds, src/sim/kernels.cu:87).  Here spins are split by contiguous GLOBAL id range, one process per GPU; the
phantom and the sequence are replicated; RNG streams and the DEPHASING term (kernels.cu:176) are keyed by the global id
and the global spin count, so the result does not depend on the number of ranks.  The only exchange is one all-reduce
of the per-(scale, echo, substrate) sums {sum Mx, sum My, sum Mz, N}: a few KB over NCCL (NVLink/NVSwitch) on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations


def shard_range(n_spins: int, rank: int, world: int) -> tuple[int, int]:
    """(first global id, count) of rank's shard: contiguous, sizes differ by at most one, ragged tail on the low ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if n_spins < 0:
        raise ValueError("n_spins must be >= 0")
    base, extra = divmod(int(n_spins), world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def allreduce_sums(sums, group=None):
    """In-place SUM all-reduce of a torch tensor [K][E][n_sub][4] (float64) across the process group.
    No-op without an initialised group (single process)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def signal_from_sums(sums):
    """Ensemble signal per (scale, echo): (|sum Mxy| / N, sum Mz / N, N) over all substrates, from reduced sums
    (numpy or torch, [K][E][n_sub][4])."""
    s = sums.sum(-2)
    n = s[..., 3]
    n_safe = n.clip(1) if hasattr(n, "clip") else n
    mxy = (s[..., 0] ** 2 + s[..., 1] ** 2) ** 0.5 / n_safe
    return mxy, s[..., 2] / n_safe, n
