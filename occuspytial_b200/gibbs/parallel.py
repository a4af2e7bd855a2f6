"""Multi-chain execution (reference: occuspytial/gibbs/parallel.py:4-42).

The reference fans chains out to one OS process each through joblib.  Here all chains of a device
advance in lock-step inside the CUDA engine, and a multi-GPU job is one process per GPU under
``torch.distributed``: chain c of ``chains`` lives on rank ``c // ceil(chains / world)``, problem data
is replicated, there is no per-iteration traffic, and the only exchange is an all-gather of the
per-chain summary blocks used for R-hat / ESS (``allgather_summaries``).
"""
import numpy as np


def shard_bounds(chains, world_size, rank):
    """Contiguous block [lo, hi) of global chain ids owned by ``rank``."""
    per = -(-chains // world_size)
    lo = min(rank * per, chains)
    return lo, min(lo + per, chains)


def dist_info():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def allgather_summaries(local, chains):
    """All-gather per-chain summary blocks: ``local`` (local_chains, ...) -> (chains, ...) on every rank.

    Uses the default process group (NCCL on GPUs -- the tensor stays on the device -- or gloo on CPU).
    Shards may be ragged (last rank shorter); they are padded to the common shard size for the collective.
    """
    import torch
    import torch.distributed as dist
    rank, world = dist_info()
    t = local if isinstance(local, torch.Tensor) else torch.as_tensor(np.asarray(local))
    if world == 1:
        return t
    per = -(-chains // world)
    pad = per - t.shape[0]
    if pad:
        t = torch.cat([t, t.new_zeros((pad,) + tuple(t.shape[1:]))], dim=0)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return torch.cat(out, dim=0)[:chains]


def sample_parallel(class_, **kwargs):
    """Run ``chains`` chains of sampler ``class_`` and return the list of per-chain ``Chain`` objects.

    Same call shape as the reference (parallel.py:4); the keyword arguments are those of ``sample``.
    Under an initialised ``torch.distributed`` group each rank runs (and returns) only its shard.
    """
    chains = kwargs.pop('chains')
    return class_._run_lockstep(chains=chains, **kwargs)
