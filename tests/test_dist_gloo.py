"""World-size-2 CPU (gloo) tests of the multi-GPU host logic: chain sharding and the all-gather of per-chain
summary blocks that feeds R-hat / ESS (the only exchange of the method)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, chains, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from occuspytial_b200.gibbs.parallel import shard_bounds, allgather_summaries, dist_info
    from occuspytial_b200.diagnostics import summaries_from_draws
    assert dist_info() == (rank, world)
    lo, hi = shard_bounds(chains, world, rank)
    # every rank can regenerate any chain (as the Philox-keyed engine can); it only "runs" its shard
    draws = np.stack([np.random.default_rng(c).normal(size=200) + 0.01 * c for c in range(chains)])
    local = summaries_from_draws(draws[lo:hi], 20)
    full = allgather_summaries(torch.from_numpy(local), chains).numpy()
    np.save(os.path.join(out_dir, f'rank{rank}.npy'), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('chains', [8, 7])
def test_allgather_summaries_world2(tmp_path, chains):
    import torch.multiprocessing as mp
    from occuspytial_b200.diagnostics import summaries_from_draws, split_rhat, ess
    port = _free_port()
    mp.spawn(_worker, args=(2, port, chains, str(tmp_path)), nprocs=2, join=True)
    draws = np.stack([np.random.default_rng(c).normal(size=200) + 0.01 * c for c in range(chains)])
    expect = summaries_from_draws(draws, 20)
    for rank in range(2):
        got = np.load(tmp_path / f'rank{rank}.npy')
        np.testing.assert_allclose(got, expect, rtol=1e-13, atol=1e-15)
    assert abs(split_rhat(expect, 200) - 1) < 0.1 and ess(expect, 200) > 500


def test_shard_bounds_cover_all_chains():
    from occuspytial_b200.gibbs.parallel import shard_bounds
    for chains in (1, 2, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(chains, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == chains
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) == -(-chains // world)


def _gather_worker(rank, world, port, chains, draws, out_dir):
    """The sampler-side gather of sample(): per-engine summary blocks -> packed tensor -> all-gather -> per-parameter blocks
    (gibbs/base.py:_gather_summaries), with fake engine results on CPU tensors; a rank may hold an empty shard."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from occuspytial_b200.gibbs.base import GibbsBase, _summary_lag
    from occuspytial_b200.gibbs.parallel import shard_bounds
    from occuspytial_b200.diagnostics import summaries_from_draws
    dims = {'alpha': 2, 'beta': 3, 'tau': 1}
    lo, hi = shard_bounds(chains, world, rank)
    L = _summary_lag(draws)
    rng = np.random.default_rng(5)
    full = {k: rng.normal(size=(chains, draws, d)) for k, d in dims.items()}
    results = []
    if hi > lo:   # two engines on this rank (as with two devices in one process)
        mid = (lo + hi + 1) // 2
        for a, b in ((lo, mid), (mid, hi)):
            if b > a:
                blocks = {k: torch.from_numpy(np.stack([summaries_from_draws(full[k][a:b, :, j], L) for j in range(d)], axis=1))
                          for k, d in dims.items()}
                results.append((None, blocks, None))
    sampler = GibbsBase.__new__(GibbsBase)
    got = sampler._gather_summaries(results, chains, hi - lo, draws, dims, world)
    np.savez(os.path.join(out_dir, f'g{rank}.npz'), **got)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('chains', [7, 1])
def test_sampler_gather_of_summary_blocks_world2(tmp_path, chains):
    import torch.multiprocessing as mp
    from occuspytial_b200.diagnostics import summaries_from_draws
    from occuspytial_b200.gibbs.base import _summary_lag
    draws = 40
    port = _free_port()
    mp.spawn(_gather_worker, args=(2, port, chains, draws, str(tmp_path)), nprocs=2, join=True)
    dims = {'alpha': 2, 'beta': 3, 'tau': 1}
    rng = np.random.default_rng(5)
    full = {k: rng.normal(size=(chains, draws, d)) for k, d in dims.items()}
    for rank in range(2):
        got = np.load(tmp_path / f'g{rank}.npz')
        for k, d in dims.items():
            expect = np.stack([summaries_from_draws(full[k][:, :, j], _summary_lag(draws)) for j in range(d)], axis=1)
            assert got[k].shape == (chains, d, 6 + _summary_lag(draws) + 1)
            np.testing.assert_allclose(got[k], expect, rtol=1e-13, atol=1e-15)


def test_start_values_do_not_depend_on_sharding():
    """Default start values of the global chains [lo, hi) come from per-block generators keyed by (seed, block): any shard
    reproduces the slice of the full draw (gibbs/base.py:_start_arrays_range; reference: base.py:199-212)."""
    from occuspytial_b200 import LogitRSRGibbs, ProbitRSRGibbs
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(n=150, min_v=1, max_v=4, lattice=(10, 15), random_state=1)
    for cls in (LogitRSRGibbs, ProbitRSRGibbs):
        s = cls(Q, W, X, y, random_state=3, q=5)
        full = s._start_arrays_range(1234, 0, 300, None)
        assert full['eta'].shape == (300, 5) and full['tau'].shape == (300,)
        for lo, hi in ((0, 300), (0, 128), (100, 131), (128, 256), (257, 300)):
            part = s._start_arrays_range(1234, lo, hi, None)
            for k in full:
                np.testing.assert_array_equal(part[k], full[k][lo:hi])
        other = s._start_arrays_range(1235, 0, 4, None)
        assert not np.allclose(other['tau'], full['tau'][:4])
        fixed = s._start_arrays_range(1, 5, 9, {k: np.zeros(v.shape[1:]) for k, v in full.items()})
        assert all(v.shape[0] == 4 and not v.any() for v in fixed.values())
