"""Sampler base class: same public surface as the reference's GibbsBase (occuspytial/gibbs/base.py:30-306),
with the per-iteration work moved into the CUDA engine.

What stays on the host (one-off): data partitioning, hyper-parameters, start values, the RSR basis.
What moved to the device: everything inside ``for i in range(size): self.step(); chain.append(...)``
(base.py:236-239) -- the loop runs in C++/CUDA for all chains in lock-step and ``Chain`` objects are
zero-copy views of the block of draws copied back once per flush.
"""
import contextlib

import numpy as np
from scipy.sparse import csc_matrix, isspmatrix_csc
from tqdm.auto import tqdm

from ..chain import Chain, ChainViews
from ..data import Data
from ..posterior import PosteriorParameter
from ..state import State, FixedState
from ..utils import get_generator
from . import _setup
from .parallel import sample_parallel, shard_bounds, dist_info

# device record buffer budget (bytes) before draws are flushed to the host
_RECORD_BYTES = 2 << 30


def _summary_lag(draws):
    """Largest autocovariance lag of the per-chain summary blocks (same rule as diagnostics.summary_frame)."""
    return int(min(draws - 1, 250))


def _check_visits_match(w_ptr, y_ptr):
    """W and y must describe the same visits: same surveyed sites, same number of visits per site.  The kernels
    index y with W's visit offsets, so a mismatch would silently misalign every later site (the reference fails
    with a shape error in ``W.T @ y``, logit.py:223)."""
    if w_ptr.shape != y_ptr.shape or not np.array_equal(w_ptr, y_ptr):
        bad = np.flatnonzero(np.diff(w_ptr) != np.diff(y_ptr))
        where = f' (first mismatch at site {int(bad[0])})' if bad.size else ''
        raise ValueError('W and y must have the same surveyed sites and the same number of visits per site' + where)


class _GibbsState(State):
    """Sampler state; ``posteriors`` returns the stored parameters (base.py:15-27)."""

    _posterior_names = ('alpha', 'beta', 'tau')

    @property
    def posteriors(self):
        return {name: self.__dict__[name] for name in self._posterior_names}


class GibbsBase:
    """Base class for Gibbs samplers of species spatial occupancy models.

    Parameters are those of the reference (base.py:33-60): ``Q`` spatial precision (scipy sparse or
    ndarray), ``W`` / ``y`` dictionaries keyed by surveyed site, ``X`` occupancy design matrix,
    ``hparams`` (all six keys when given) and ``random_state``.

    Attributes ``chain``, ``dists``, ``fixed``, ``rng``, ``state`` are kept (base.py:62-82).
    """

    _model = None  # engine model name, set by subclasses that implement the sweep

    def __init__(self, Q, W, X, y, hparams=None, random_state=None):
        self.W = Data(W)
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.y = Data(y)
        self.rng = get_generator(random_state)
        self._engine = None

    # ------------------------------------------------------------------------------------------
    # configuration (host, one-off)
    # ------------------------------------------------------------------------------------------
    def _configure(self, Q, hparams, verify_precision=True, **kwargs):
        """To be called by subclasses in order to configure the sampler (base.py:107-164)."""
        if verify_precision:
            self._verify_spatial_precision(Q)
        n = self.X.shape[0]
        part = _setup.partition(self.y, n)
        self.state = _GibbsState()
        self.state.z = part['z']
        self.state.k = self.state.z - 0.5

        self.fixed = FixedState()
        self.fixed.Q = Q if isspmatrix_csc(Q) else csc_matrix(Q)
        self.fixed.n = n
        self.fixed.ones = np.ones(n)
        self.fixed.not_surveyed = part['not_surveyed']
        self.fixed.not_obs = part['not_obs']
        self.fixed.obs = part['obs']
        self.fixed.n_no = len(part['not_obs'])
        self.fixed.n_ns = len(part['not_surveyed'])
        self.fixed.site_class = part['site_class']
        if self.fixed.n_no:
            self.fixed.W_not_obs = self.W[self.fixed.not_obs]
            self.fixed.visits_not_obs = self.W.visits(self.fixed.not_obs)
        else:
            width = self.W[self.W.surveyed[0]].shape[1]
            self.fixed.W_not_obs = np.zeros((0, width))
            self.fixed.visits_not_obs = ()
        sections = np.cumsum(self.fixed.visits_not_obs, dtype=np.int64)
        self.fixed.stacked_w_indices = np.concatenate([[0], sections])[:-1] if sections.size else sections

        if hparams:
            for key, value in hparams.items():   # used verbatim, defaults are not merged (base.py:172-175)
                setattr(self.fixed, key, value)
        else:
            self._set_default_hyperparams(self.fixed)
        self.fixed.a_prec_by_mu = np.asarray(self.fixed.a_prec) @ np.asarray(self.fixed.a_mu)
        self.fixed.b_prec_by_mu = np.asarray(self.fixed.b_prec) @ np.asarray(self.fixed.b_mu)
        self.dists = FixedState()

    def _verify_spatial_precision(self, Q):
        _setup.verify_singular(Q)

    def _set_default_hyperparams(self, params):
        """base.py:177-186."""
        params.tau_rate = 0.005
        params.tau_shape = 0.5 + 0.5 * (self.fixed.n - 1)
        r = self.W[self.W.surveyed[0]].shape[1]
        params.a_mu = np.zeros(r)
        params.a_prec = np.eye(r) / 10
        p = self.X.shape[1]
        params.b_mu = np.zeros(p)
        params.b_prec = np.eye(p) / 10
        return params

    # ------------------------------------------------------------------------------------------
    # start values (base.py:188-212), vectorised over chains
    # ------------------------------------------------------------------------------------------
    def _eta_start(self, chains, rng=None):
        rng = self.rng if rng is None else rng
        eta = rng.standard_normal((chains, self.fixed.n))
        return eta - eta.mean(axis=1, keepdims=True)

    def _draw_start(self, chains, rng=None):
        """Default start values for ``chains`` chains: tau ~ Gamma(.5, 1/tau_rate), alpha ~ N(a_mu, 100 a_prec),
        beta ~ N(b_mu, 100 b_prec) (precision used as covariance, as the reference does), eta per model."""
        rng = self.rng if rng is None else rng
        f = self.fixed
        La = np.linalg.cholesky(100 * np.asarray(f.a_prec, dtype=np.float64))
        Lb = np.linalg.cholesky(100 * np.asarray(f.b_prec, dtype=np.float64))
        start = {
            'tau': rng.gamma(0.5, 1 / f.tau_rate, size=chains),
            'eta': self._eta_start(chains, rng),
            'alpha': np.asarray(f.a_mu) + rng.standard_normal((chains, La.shape[0])) @ La.T,
            'beta': np.asarray(f.b_mu) + rng.standard_normal((chains, Lb.shape[0])) @ Lb.T,
        }
        return start

    def _start_arrays(self, chains, start, rng=None):
        if start is None:
            return self._draw_start(chains, rng)
        out = {}
        for key in self._start_keys():
            value = np.asarray(start[key], dtype=np.float64)
            out[key] = np.broadcast_to(value, (chains,) + value.shape).copy()
        return out

    _START_BLOCK = 128   # chains per start-value generator
    _DEFAULT_START = (1.0, True)   # (eta_scale, centre_eta) of occu_draw_start: eta ~ N(0, I_n) minus its mean (base.py:201-203)

    def _start_arrays_range(self, seed, lo, hi, start):
        """Host-side start values of the global chains [lo, hi) (for engines driven by hand; ``sample()`` draws its
        default starts on the device, occu_draw_start).  Default starts come from one counter-based generator per block
        of 128 chains, keyed by (seed, block): a shard draws only the blocks it overlaps, and a chain's start value
        does not depend on how the job is sharded over devices or ranks."""
        if start is not None:
            return self._start_arrays(hi - lo, start)
        B = self._START_BLOCK

        parts = []
        for blk in range(lo // B, -(-hi // B)):
            rng = np.random.Generator(np.random.Philox(key=[seed & (2**64 - 1), blk]))
            arrays = self._draw_start(B, rng)
            a, b = max(lo, blk * B) - blk * B, min(hi, (blk + 1) * B) - blk * B
            parts.append({k: v[a:b] for k, v in arrays.items()})
        return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}

    def _start_keys(self):
        return ('alpha', 'beta', 'tau', 'eta')

    def _start_to_engine(self, arrays):
        return arrays

    def _eta_from_engine(self, eta):
        return eta

    def _initialize_posterior_state(self, start=None):
        """Set the host-visible state of chain 0 (kept for API compatibility, base.py:188-197)."""
        arrays = self._start_arrays(1, start)
        for key, value in arrays.items():
            setattr(self.state, key, value[0])
        self.state.spatial = self._spatial_from_eta(self.state.eta)
        return arrays

    def _spatial_from_eta(self, eta):
        return eta

    # site-indexed arrays cross the engine boundary in the engine's site order (identity unless a
    # subclass reorders the sites, as the ICAR sampler does to narrow the band of Q)
    def _sites_to_engine(self, arr):
        return arr

    def _sites_from_engine(self, arr):
        return arr

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    # ------------------------------------------------------------------------------------------
    def _problem_arrays(self):
        """Everything the device engine needs, as host arrays (see include/occu_b200.h: occu_problem_t)."""
        f = self.fixed
        n = f.n
        r = self.W[self.W.surveyed[0]].shape[1]
        if self.X.shape[1] > 32 or r > 8:
            raise ValueError('the engine supports at most 32 occupancy covariates (columns of X) and 8 detection '
                             f'covariates (columns of W); got {self.X.shape[1]} and {r}')
        q = getattr(f, 'q', None)
        if q is not None and q > 3072:
            raise ValueError(f'RSR rank q = {q} exceeds the engine limit of 3072; pass a smaller q or a larger threshold r')
        W_rows, visit_ptr = self.W.to_csr(n)
        y_rows, y_ptr = self.y.to_csr(n, dtype=np.uint8)
        _check_visits_match(visit_ptr, y_ptr)
        return dict(X=self.X, W_rows=W_rows.reshape(W_rows.shape[0], -1), y_rows=y_rows, visit_ptr=visit_ptr,
                    site_class=f.site_class, tau_rate=f.tau_rate, tau_shape=f.tau_shape,
                    a_mu=f.a_mu, a_prec=f.a_prec, b_mu=f.b_mu, b_prec=f.b_prec)

    def _make_engine(self, chains, seed, chain_offset=0, device=None, debug=False, pg_fp64=False, gram_int8=None, slot=0):
        """Engine for ``chains`` chains.  ``gram_int8``: True / False force the Gram path of LogitRSR (exact integer
        slicing on the tcgen05 tensor cores / hand-written FP64 DMMA kernel); None picks the tcgen05 path when the
        problem is large enough for it to pay (>= 1024 sites; the rule must not depend on the number of chains, or a
        shard of a job would round differently from the whole job) and falls back to the DMMA kernel -- another GPU
        kernel of this engine, not a CPU path -- if the TMA descriptors cannot be encoded."""
        from ..engine import DeviceEngine
        from .._lib import EngineError
        if self._model is None:
            raise NotImplementedError(f'{self.__class__.__name__} must implement a `step` method.')
        auto = gram_int8 is None
        if auto:
            gram_int8 = self._model == 'logit_rsr' and self.fixed.n >= 1024
        problem = self.__dict__.get('_problem_cache')
        if problem is None:   # flattening the ragged W / y dictionaries is the costly part: once per sampler
            problem = self._problem_arrays()
            self.__dict__['_problem_cache'] = problem
        import torch
        key = str(torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)) \
            if torch.cuda.is_available() else None
        ws = self.__dict__.get('_ws_cache', {}).pop(key, None)   # the previous run's workspace on this device, if any
        try:
            return DeviceEngine(self._model, problem, chains, seed, chain_offset=chain_offset,
                                device=device, debug=debug, pg_fp64=pg_fp64, gram_int8=gram_int8, workspace=ws, slot=slot)
        except EngineError as exc:
            if not (auto and gram_int8 and 'INT8' in str(exc)):
                raise
            return DeviceEngine(self._model, problem, chains, seed, chain_offset=chain_offset,
                                device=device, debug=debug, pg_fp64=pg_fp64, gram_int8=False, workspace=ws, slot=slot)

    def step(self):
        """Step through one iteration of the sampler.

        Subclasses bound to an engine model advance their (lazily created) single-chain engine by one
        iteration and refresh ``self.state``; the base class has no sweep (base.py:90-105).
        """
        if self._model is None:
            raise NotImplementedError(f'{self.__class__.__name__} must implement a `step` method.')
        if self._engine is None:
            start = self._initialize_posterior_state(None)
            seed = int(self.rng.integers(0, 2**63 - 1))
            self._engine = self._make_engine(1, seed)
            self._engine.set_start(self._start_to_engine(start))
        self._engine.sweep(1)
        self._pull_state(self._engine)

    def _pull_state(self, engine):
        from .. import _lib
        self.state.alpha = engine.get_field(_lib.F_ALPHA, chain=0)[0]
        self.state.beta = engine.get_field(_lib.F_BETA, chain=0)[0]
        self.state.tau = float(engine.get_field(_lib.F_TAU, chain=0)[0])
        self.state.eta = self._eta_from_engine(engine.get_field(_lib.F_ETA, chain=0)[0])
        self.state.spatial = self._spatial_from_eta(self.state.eta)
        self.state.z = self._sites_from_engine(engine.get_field(_lib.F_Z, chain=0)[0].astype(np.float64))
        self.state.k = self.state.z - 0.5

    def _pick_devices(self, device, devices, local_chains):
        """Devices of a single-process run.  ``device`` pins one; ``devices`` may be a list, an int or 'all'; by default
        every visible GPU takes part as long as each gets at least 128 chains (the reference's default is all cores,
        parallel.py:38)."""
        import torch
        if device is not None:
            return [torch.device(device)]
        count = torch.cuda.device_count() if torch.cuda.is_available() else 0
        if isinstance(devices, (list, tuple)):
            return [torch.device(d) if not isinstance(d, int) else torch.device('cuda', d) for d in devices]
        if count <= 1:
            return [None]   # DeviceEngine: current device (raises the "no CPU fallback" error when there is none)
        if devices == 'all':
            use = count
        elif isinstance(devices, int):
            use = max(1, min(devices, count))
        else:
            use = max(1, min(count, local_chains // 128))
        if use == 1:
            return [None]
        return [torch.device('cuda', i) for i in range(use)]

    def _run_shard(self, device, lo, hi, seed, start, size, burnin, bar, pg_fp64, gram_int8, want_summaries, slot=0):
        """Chains [lo, hi) in lock-step on one device.  Returns (host arrays, summary blocks or None, stats)."""
        import time as _time
        import torch
        local = hi - lo
        t_begin = _time.perf_counter()
        # start values given by the caller are broadcast to every chain on the host; the default ones are drawn on the
        # device from Philox blocks keyed by the global chain id (no host generator, no H2D copy)
        arrays = self._start_arrays(local, start) if start is not None else None
        t_start = _time.perf_counter()
        engine = self._make_engine(local, seed, chain_offset=lo, device=device, pg_fp64=pg_fp64, gram_int8=gram_int8, slot=slot)
        summaries = None
        try:
            if arrays is not None:
                engine.set_start(self._start_to_engine(arrays))
            else:
                engine.draw_start(seed ^ 0x5DEECE66D, *self._DEFAULT_START)
            t_setup = _time.perf_counter()
            draws = size - burnin
            dims = {'alpha': engine.r, 'beta': engine.p, 'tau': 1}
            per_draw = 8 * local * sum(dims.values())
            capacity = int(max(1, min(draws, _RECORD_BYTES // per_draw)))
            # everything fits one record block (the usual case): its host copy is the result.  Otherwise two blocks
            # alternate: while the sweep fills one, the other crosses PCIe on a second stream.
            bufs = [engine.alloc_record(capacity)] + ([engine.alloc_record(capacity)] if capacity < draws else [])
            shapes = {k: (local, draws) + ((d,) if d > 1 else ()) for k, d in dims.items()}
            host = {k: (None if capacity >= draws else np.empty(shp)) for k, shp in shapes.items()}
            block = int(min(max(1, size // 20), 256)) if bar is not None else 256
            copy_stream = torch.cuda.Stream(device=engine.device) if len(bufs) > 1 else None
            t_flush = 0.0

            def copy_out(rec, filled, flushed, event):
                nonlocal t_flush
                t0 = _time.perf_counter()
                with torch.cuda.device(engine.device):
                    if copy_stream is not None:
                        copy_stream.wait_event(event)
                        ctx = torch.cuda.stream(copy_stream)
                    else:
                        engine.stream.synchronize()
                        ctx = torch.cuda.stream(engine.stream)
                    with ctx:
                        for k in dims:
                            part = rec[k][:, :filled].cpu().numpy()
                            engine.d2h_bytes += part.nbytes
                            if host[k] is None:   # everything fitted in one block: the copy IS the result
                                host[k] = part.reshape(shapes[k])
                            else:
                                dst = host[k][:, flushed:flushed + filled]
                                dst[...] = part.reshape(dst.shape)
                t_flush += _time.perf_counter() - t0

            done, filled, flushed, cur = 0, 0, 0, 0
            pending = None   # a full record block waiting to be copied: (block, filled, flushed offset, event)
            while done < size:
                step = min(block, size - done)
                if done < burnin:
                    step = min(step, burnin - done)
                    engine.sweep_enqueue(step)
                else:
                    step = min(step, capacity - filled)
                    engine.sweep_enqueue(step, bufs[cur], rec_start=filled)
                    filled += step
                done += step
                if pending is not None:      # copy the previous block while the GPU works on the one just enqueued
                    copy_out(*pending)
                    pending = None
                if done >= burnin and filled == capacity and done < size:
                    ev = torch.cuda.Event()
                    ev.record(engine.stream)
                    pending = (bufs[cur], filled, flushed, ev)
                    flushed += filled
                    filled = 0
                    cur = (cur + 1) % len(bufs)
                if bar is not None:
                    engine.finish()          # the bar follows the device, not the launch queue
                    bar.update(step)
            fails = engine.finish()
            t_sweep = _time.perf_counter() - t_setup - t_flush
            if pending is not None:
                copy_out(*pending)
            t0 = _time.perf_counter()
            if want_summaries and len(bufs) == 1 and filled == draws and draws >= 4:
                summaries = engine.chain_summaries(bufs[0], draws, _summary_lag(draws))
            t_summ = _time.perf_counter() - t0
            if len(bufs) == 1 and filled == draws:
                # the usual case: one packed copy through page-locked memory brings back the draws; the summary blocks stay
                # on the device (their cross-chain reduction is all `summary` needs, diagnostics.ChainSummaries)
                t0 = _time.perf_counter()
                names = list(dims)
                parts = engine.to_host([bufs[0][k] for k in names])
                for i, k in enumerate(names):
                    host[k] = parts[i].reshape(shapes[k])
                t_flush += _time.perf_counter() - t0
            elif filled:
                ev = torch.cuda.Event()
                ev.record(engine.stream)
                copy_out(bufs[cur], filled, flushed, ev)
            if fails:
                raise RuntimeError(f'Cholesky factorization/solver failed! ({fails} of {local} chains)')
            t0 = _time.perf_counter()
            if lo == 0:
                self._pull_state(engine)
            t_pull = _time.perf_counter() - t0
            stats = dict(launches=engine.launches, h2d_bytes=engine.h2d_bytes, d2h_bytes=engine.d2h_bytes, chains=local,
                         chain_offset=lo, device=str(engine.device),
                         seconds=dict(start_values=t_start - t_begin, setup=t_setup - t_start, sweeps=t_sweep,
                                      flush=t_flush, summaries=t_summ, pull_state=t_pull, **engine.setup_seconds))
        finally:
            ws = engine.close(keep_workspace=True)
            if ws is not None:
                self.__dict__.setdefault('_ws_cache', {})[str(engine.device)] = ws
        return host, summaries, stats

    def _run_lockstep(self, size, burnin=0, start=None, chains=2, progressbar=True, pos=0, device=None, devices=None,
                      pg_fp64=False, gram_int8=None):
        """All chains in lock-step; returns the list of ``Chain`` views of the chains this process ran.

        Single process: the chains are split contiguously over the visible GPUs (``devices``), one engine per device
        driven from its own host thread -- no launcher needed, like the reference's ``Parallel(n_jobs=-1)``
        (parallel.py:38).  Under torch.distributed (one process per GPU) the global ``chains`` are sharded over ranks
        and only the local shard is run and returned.  Either way global chain ids key the Philox streams and the
        start values, so the draws of a chain do not depend on the number of GPUs; the only exchange is the
        all-gather of per-chain summary blocks for R-hat / ESS (``PosteriorParameter.summary``).
        """
        if self._model is None:
            raise NotImplementedError(f'{self.__class__.__name__} must implement a `step` method.')
        import gc
        import time as _time
        t_call = _time.perf_counter()
        # a generational collection of the host program's objects in the middle of a run stalls the thread that feeds the
        # launch queue for milliseconds (measured: benchmarks/e2e_breakdown.py); the run itself allocates a few dozen
        # objects, so the collector is paused for its duration
        gc_was_enabled = gc.isenabled()
        gc.disable()
        try:
            return self._run_lockstep_body(t_call, size, burnin, start, chains, progressbar, pos, device, devices, pg_fp64,
                                           gram_int8)
        finally:
            if gc_was_enabled:
                gc.enable()

    def _run_lockstep_body(self, t_call, size, burnin, start, chains, progressbar, pos, device, devices, pg_fp64, gram_int8):
        import time as _time
        rank, world = dist_info()
        # the same host RNG stream is consumed on every rank: one draw for the device seed
        seed = int(self.rng.integers(0, 2**63 - 1))
        lo, hi = shard_bounds(chains, world, rank)
        local = hi - lo
        draws = size - burnin
        dims = {'alpha': self.W[self.W.surveyed[0]].shape[1], 'beta': self.X.shape[1], 'tau': 1}
        self._last_arrays = None
        self._last_stats = None
        self._last_summaries = None
        if world > 1:
            import os
            devs = [device if device is not None else f"cuda:{os.environ.get('LOCAL_RANK', rank)}"]
        else:
            devs = self._pick_devices(device, devices, local)
        bar = tqdm(total=size, disable=not progressbar, position=pos) if progressbar else None
        results = []
        if local > 0:
            bounds = [shard_bounds(local, len(devs), i) for i in range(len(devs))]
            jobs = [(dev, lo + a, lo + b) for dev, (a, b) in zip(devs, bounds) if b > a]

            def job(k):
                dev, a, b = jobs[k]
                return self._run_shard(dev, a, b, seed, start, size, burnin, bar if k == 0 else None, pg_fp64, gram_int8,
                                       want_summaries=True, slot=k)
            if len(jobs) == 1:
                results = [job(0)]
            else:
                from concurrent.futures import ThreadPoolExecutor
                with ThreadPoolExecutor(len(jobs)) as pool:   # the C ABI releases the GIL: devices run concurrently
                    results = list(pool.map(job, range(len(jobs))))
        if bar is not None:
            bar.close()
        if results:
            host = {k: (results[0][0][k] if len(results) == 1 else np.concatenate([r[0][k] for r in results], axis=0))
                    for k in dims}
        else:       # more ranks than chains: this rank holds an empty shard
            host = {k: np.empty((0, draws) + ((d,) if d > 1 else ())) for k, d in dims.items()}
        t0 = _time.perf_counter()
        self._last_summaries = self._gather_summaries(results, chains, local, draws, dims, world)
        t_gather = _time.perf_counter() - t0
        if results:
            stats = [r[2] for r in results]
            agg = dict(stats[0])
            for key in ('launches', 'h2d_bytes', 'd2h_bytes', 'chains'):
                agg[key] = sum(s_[key] for s_ in stats)
            agg['devices'] = [s_['device'] for s_ in stats]
            agg['per_device'] = stats
            agg['seconds'] = dict(agg['seconds'], gather_summaries=t_gather, run_lockstep=_time.perf_counter() - t_call)
            self._last_stats = agg
        out = ChainViews(host) if local > 0 else []
        if local > 0:
            self.chain = out[0]
        self._last_arrays = host
        return out

    def _gather_summaries(self, results, chains, local, draws, dims, world):
        """:class:`~occuspytial_b200.diagnostics.ChainSummaries` of ALL chains of the job, or None.  The per-chain blocks
        of the local engines are concatenated on the first engine's device; under torch.distributed they are
        all-gathered (NCCL).  Either way they are reduced over the chains where they are: only the reduction -- a few
        hundred numbers -- crosses PCIe here, the blocks follow if somebody asks for them."""
        if draws < 4:
            return None
        have = bool(results) and all(r[1] is not None for r in results)
        import torch
        from ..diagnostics import ChainSummaries
        names = list(dims)
        widths = [dims[k] for k in names]

        def pack(result, dev):    # one tensor per engine: a single collective moves all three parameters
            return torch.cat([result[1][k].to(dev) for k in names], dim=1)

        if world == 1:
            if not have:
                return None
            dev = results[0][1][names[0]].device
            ctx = torch.cuda.device(dev) if dev.type == 'cuda' else contextlib.nullcontext()
            with ctx:
                packed = torch.cat([pack(r, dev) for r in results], dim=0)
                return ChainSummaries(names, widths, chains, packed)
        import torch.distributed as dist
        from .parallel import allgather_summaries
        # every rank must agree on whether the collective runs (a rank that flushed in blocks has no device summaries)
        dev = results[0][1]['tau'].device if have else torch.device('cuda', torch.cuda.current_device()) \
            if dist.get_backend() == 'nccl' else torch.device('cpu')
        flag = torch.tensor([1 if (have or local == 0) else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            return None
        W = 6 + _summary_lag(draws) + 1
        if have:
            t = torch.cat([pack(r, dev) for r in results], dim=0)
        else:
            t = torch.zeros((0, sum(widths), W), dtype=torch.float64, device=dev)
        return ChainSummaries(names, widths, chains, allgather_summaries(t, chains))

    def _run(self, size, burnin=0, start=None, chains=2, progressbar=True, pos=0):
        """Single-chain entry point kept for API compatibility (base.py:214-241)."""
        return self._run_lockstep(size, burnin=burnin, start=start, chains=1, progressbar=progressbar, pos=pos)[0]

    def sample(self, size, burnin=0, start=None, chains=2, progressbar=True, **engine_kwargs):
        r"""Obtain posterior samples of ``alpha``, ``beta`` and ``tau`` (base.py:243-291).

        Parameters
        ----------
        size : int
            The number of total samples to generate.
        burnin : int, optional
            The number of initial samples to discard; must be less than ``size``.
        start : {None, Dict[str, np.ndarray]}, optional
            Starting values of ``alpha``, ``beta``, ``tau`` and ``eta`` (shared by all chains).
        chains : int, optional
            Number of chains; they advance in lock-step on the GPU, thousands are cheap.
        progressbar : bool, optional
            Whether to display the progress bar during sampling.
        **engine_kwargs
            Engine knobs: ``device`` (one torch device), ``devices`` (list, count or 'all': GPUs of a single-process
            run; default: every visible GPU that gets at least 128 chains), ``pg_fp64`` (validation precision for the PG / truncated-
            normal draws), ``gram_int8`` (LogitRSR: exact INT8 slice GEMMs for the Omega-weighted Gram).

        Returns
        -------
        out : :class:`~occuspytial_b200.posterior.PosteriorParameter`
        """
        if burnin >= size:
            raise ValueError('burnin value cannot be larger than sample size')
        if chains < 1:
            raise ValueError('chains must a positive integer.')
        samples = sample_parallel(self, size=size, burnin=burnin, chains=chains, start=start,
                                  progressbar=progressbar, **engine_kwargs)
        stacked = getattr(self, '_last_arrays', None)
        if stacked is not None and stacked['tau'].shape[0] == len(samples):
            # already (chains, draws[, dim]): skip np.stack; R-hat / ESS come from the device summary blocks
            return PosteriorParameter.from_arrays(stacked, summaries=getattr(self, '_last_summaries', None))
        return PosteriorParameter(*samples)

    def copy(self):
        """Copy of this sampler sharing everything but the random generator (base.py:293-306)."""
        out = type(self).__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out.__dict__['_engine'] = None
        out.__dict__.pop('_ws_cache', None)
        seed_seq = self.rng.bit_generator.seed_seq.spawn(1)[0]
        out.__dict__['rng'] = get_generator(seed_seq)
        return out
