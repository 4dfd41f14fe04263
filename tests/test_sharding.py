"""Host logic of the multi-GPU path on CPU: spin sharding + the one collective (all-reduce of the per-echo ensemble
sums), world size 2 over gloo.  The per-rank simulator here is the CPU oracle (test infrastructure) standing in for the
CUDA engine: what is under test is spinwalk_b200.sharding — shard ranges, global-id keyed results, the reduce.
The same invariance is checked bitwise on the GPU by tests/test_engine_gpu.py::test_shards_are_invariant."""
import os
import socket

import numpy as np
import pytest

import cases
from spinwalk_b200.sharding import allreduce_sums, shard_range, signal_from_sums


@pytest.mark.parametrize("n,world", [(0, 1), (1, 2), (10, 3), (1000, 8), (7, 8), (10_000_000, 8), (1_000_000_007, 4)])
def test_shard_range_partitions(n, world):
    parts = [shard_range(n, r, world) for r in range(world)]
    assert parts[0][0] == 0
    for (f0, c0), (f1, _) in zip(parts, parts[1:]):
        assert f0 + c0 == f1
    assert parts[-1][0] + parts[-1][1] == n
    counts = [c for _, c in parts]
    assert max(counts) - min(counts) <= 1 and sorted(counts, reverse=True) == counts
    with pytest.raises(ValueError):
        shard_range(n, world, world)


def _host_sums(M1, T, n_sub):
    K, S, E, _ = M1.shape
    out = np.zeros((K, E, n_sub, 4), np.float64)
    valid = (M1 != 0).any(axis=3)
    for sub in range(n_sub):
        w = (T == sub) & valid
        out[:, :, sub, :3] = (M1.astype(np.float64) * w[..., None]).sum(axis=1)
        out[:, :, sub, 3] = w.sum(axis=1)
    return out


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from oracle import pyoracle as po

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        case, mask, fm, fov, xyz0 = cases.multi_echo(n_spins=301)  # odd count: ragged shards
        first, n = shard_range(case.n_spins, rank, world)
        # the oracle keys RNG and dephasing by the global spin id, like the engine; simulate only this rank's id range
        r = po.run_oracle(case, fm, mask, xyz0, flavour=po.RNG_MINSTD, threads=2, spins=(first, first + n))
        sums = torch.from_numpy(_host_sums(r["M1"][:, first:first + n], r["T"][:, first:first + n], case.n_substrate))
        allreduce_sums(sums)
        q.put((rank, first, n, sums.numpy().copy(), r["M1"][:, first:first + n].copy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_world2_gloo_reduce_equals_single_process(oracle):
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    case, mask, fm, fov, xyz0 = cases.multi_echo(n_spins=301)
    full = oracle.run_oracle(case, fm, mask, xyz0, flavour=oracle.RNG_MINSTD)
    ref = _host_sums(full["M1"], full["T"], case.n_substrate)
    # both ranks hold the same reduced sums, equal to the single-process result
    assert np.array_equal(got[0][3], got[1][3])
    assert np.allclose(got[0][3], ref, rtol=1e-12, atol=1e-9)
    assert got[0][3][..., 3].sum() == ref[..., 3].sum()
    # rank-ordered concatenation of the per-spin outputs is the single-process array (results do not depend on G)
    cat = np.concatenate([g[4] for g in got], axis=1)
    assert np.array_equal(cat, full["M1"])
    mxy, mz, n = signal_from_sums(got[0][3])
    assert mxy.shape == (case.n_scales, case.n_TE) and (n <= case.n_spins).all() and (mxy <= 1.0 + 1e-6).all()


def test_allreduce_is_noop_without_group():
    import torch

    s = torch.ones((2, 1, 2, 4), dtype=torch.float64)
    assert allreduce_sums(s) is s and float(s.sum()) == 16.0
