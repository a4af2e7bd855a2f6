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
