"""Sampler base class: same public surface as the reference's GibbsBase (occuspytial/gibbs/base.py:30-306),
with the per-iteration work moved into the CUDA engine.

What stays on the host (one-off): data partitioning, hyper-parameters, start values, the RSR basis.
What moved to the device: everything inside ``for i in range(size): self.step(); chain.append(...)``
(base.py:236-239) -- the loop runs in C++/CUDA for all chains in lock-step and ``Chain`` objects are
zero-copy views of the block of draws copied back once per flush.
"""
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
    def _eta_start(self, chains):
        eta = self.rng.standard_normal((chains, self.fixed.n))
        return eta - eta.mean(axis=1, keepdims=True)

    def _draw_start(self, chains):
        """Default start values for ``chains`` chains: tau ~ Gamma(.5, 1/tau_rate), alpha ~ N(a_mu, 100 a_prec),
        beta ~ N(b_mu, 100 b_prec) (precision used as covariance, as the reference does), eta per model."""
        f = self.fixed
        La = np.linalg.cholesky(100 * np.asarray(f.a_prec, dtype=np.float64))
        Lb = np.linalg.cholesky(100 * np.asarray(f.b_prec, dtype=np.float64))
        start = {
            'tau': self.rng.gamma(0.5, 1 / f.tau_rate, size=chains),
            'eta': self._eta_start(chains),
            'alpha': np.asarray(f.a_mu) + self.rng.standard_normal((chains, La.shape[0])) @ La.T,
            'beta': np.asarray(f.b_mu) + self.rng.standard_normal((chains, Lb.shape[0])) @ Lb.T,
        }
        return start

    def _start_arrays(self, chains, start):
        if start is None:
            return self._draw_start(chains)
        out = {}
        for key in self._start_keys():
            value = np.asarray(start[key], dtype=np.float64)
            out[key] = np.broadcast_to(value, (chains,) + value.shape).copy()
        return out

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
        W_rows, visit_ptr = self.W.to_csr(n)
        y_rows, y_ptr = self.y.to_csr(n, dtype=np.uint8)
        _check_visits_match(visit_ptr, y_ptr)
        return dict(X=self.X, W_rows=W_rows.reshape(W_rows.shape[0], -1), y_rows=y_rows, visit_ptr=visit_ptr,
                    site_class=f.site_class, tau_rate=f.tau_rate, tau_shape=f.tau_shape,
                    a_mu=f.a_mu, a_prec=f.a_prec, b_mu=f.b_mu, b_prec=f.b_prec)

    def _make_engine(self, chains, seed, chain_offset=0, device=None, debug=False, pg_fp64=False, gram_int8=None):
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
        try:
            return DeviceEngine(self._model, problem, chains, seed, chain_offset=chain_offset,
                                device=device, debug=debug, pg_fp64=pg_fp64, gram_int8=gram_int8)
        except EngineError as exc:
            if not (auto and gram_int8 and 'INT8' in str(exc)):
                raise
            return DeviceEngine(self._model, problem, chains, seed, chain_offset=chain_offset,
                                device=device, debug=debug, pg_fp64=pg_fp64, gram_int8=False)

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

    def _run_lockstep(self, size, burnin=0, start=None, chains=2, progressbar=True, pos=0, device=None,
                      pg_fp64=False, gram_int8=None):
        """All chains of this process in lock-step on one device; returns the list of ``Chain`` views.

        Under torch.distributed the global ``chains`` are sharded contiguously over ranks and only the
        local shard is run and returned (global chain ids key the Philox streams, so the draws of a
        chain do not depend on the number of GPUs).
        """
        if self._model is None:
            raise NotImplementedError(f'{self.__class__.__name__} must implement a `step` method.')
        rank, world = dist_info()
        if device is None and world > 1:
            import os
            device = f"cuda:{os.environ.get('LOCAL_RANK', rank)}"   # one process per GPU under torchrun
        # the same host RNG stream is consumed on every rank: start values and the device seed are
        # generated for ALL chains and sliced, so that sharding never changes a chain
        seed = int(self.rng.integers(0, 2**63 - 1))
        arrays = self._start_arrays(chains, start)
        lo, hi = shard_bounds(chains, world, rank)
        local = hi - lo
        self._last_arrays = None
        self._last_stats = None
        if local <= 0:   # more ranks than chains: this rank holds an empty shard
            dims = {'alpha': self.W[self.W.surveyed[0]].shape[1], 'beta': self.X.shape[1], 'tau': 1}
            self._last_arrays = {k: np.empty((0, size - burnin) + ((d,) if d > 1 else ())) for k, d in dims.items()}
            return []
        arrays = {k: v[lo:hi] for k, v in arrays.items()}
        import time as _time
        t_begin = _time.perf_counter()
        engine = self._make_engine(local, seed, chain_offset=lo, device=device, pg_fp64=pg_fp64, gram_int8=gram_int8)
        try:
            engine.set_start(self._start_to_engine(arrays))
            t_setup = _time.perf_counter() - t_begin
            t_sweep = t_flush = 0.0
            draws = size - burnin
            dims = {'alpha': engine.r, 'beta': engine.p, 'tau': 1}
            per_draw = 8 * local * sum(dims.values())
            capacity = int(max(1, min(draws, _RECORD_BYTES // per_draw)))
            record = engine.alloc_record(capacity)
            shapes = {k: (local, draws) + ((d,) if d > 1 else ()) for k, d in dims.items()}
            # one record block holds every draw (the usual case): its host copy is the result, no staging array
            host = {k: (None if capacity >= draws else np.empty(shp)) for k, shp in shapes.items()}
            block = int(min(max(1, size // 20), 256))
            bar = tqdm(total=size, disable=not progressbar, position=pos)
            done, filled, flushed = 0, 0, 0

            def flush():
                nonlocal filled, flushed, t_flush
                if not filled:
                    return
                t0 = _time.perf_counter()
                for k, d in dims.items():
                    part = record[k][:, :filled].cpu().numpy()
                    engine.d2h_bytes += part.nbytes
                    if host[k] is None:   # everything fitted in one block: the copy IS the result
                        host[k] = part.reshape(shapes[k])
                    else:
                        host[k][:, flushed:flushed + filled] = part.reshape(host[k][:, flushed:flushed + filled].shape)
                flushed += filled
                filled = 0
                t_flush += _time.perf_counter() - t0

            fails = 0
            while done < size:
                step = min(block, size - done)
                t0 = _time.perf_counter()
                if done < burnin:
                    step = min(step, burnin - done)
                    fails = engine.sweep(step)
                else:
                    step = min(step, capacity - filled)
                    fails = engine.sweep(step, record, rec_start=filled)
                    filled += step
                t_sweep += _time.perf_counter() - t0
                if done >= burnin and filled == capacity:
                    flush()
                done += step
                bar.update(step)
            flush()
            bar.close()
            if fails:
                raise RuntimeError(f'Cholesky factorization/solver failed! ({fails} of {local} chains)')
            self._pull_state(engine)
            self._last_stats = dict(launches=engine.launches, h2d_bytes=engine.h2d_bytes,
                                    d2h_bytes=engine.d2h_bytes, chains=local, chain_offset=lo,
                                    seconds=dict(setup=t_setup, sweeps=t_sweep, flush=t_flush))
        finally:
            engine.close()
        out = ChainViews(host)
        self.chain = out[0]
        self._last_arrays = host
        return out

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
            Engine knobs: ``device`` (torch device), ``pg_fp64`` (validation precision for the PG / truncated-
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
            return PosteriorParameter.from_arrays(stacked)  # already (chains, draws[, dim]): skip np.stack
        return PosteriorParameter(*samples)

    def copy(self):
        """Copy of this sampler sharing everything but the random generator (base.py:293-306)."""
        out = type(self).__new__(self.__class__)
        out.__dict__.update(self.__dict__)
        out.__dict__['_engine'] = None
        seed_seq = self.rng.bit_generator.seed_seq.spawn(1)[0]
        out.__dict__['rng'] = get_generator(seed_seq)
        return out
