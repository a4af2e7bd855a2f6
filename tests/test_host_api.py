"""CPU tests of the host side: containers with the reference's behaviour (occuspytial/tests/test_chain.py,
test_data.py, gibbs/tests/test_state.py, tests/test_utils.py, tests/test_posterior.py), sampler configuration
and error messages (gibbs/tests/test_samplers.py), and the C-ABI library's exported symbols."""
import ctypes
import os
import re
from unittest.mock import patch

import numpy as np
import pytest
from scipy.sparse import rand

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chain():
    from occuspytial_b200.chain import Chain
    c = Chain({'p1': 2, 'p2': 1}, 1)
    assert c.full.shape[1] == 3 and len(c) == 0
    c.append({'p1': [1, 2], 'p2': 3})
    assert len(c) == 1
    with pytest.raises(ValueError, match='Chain is full'):
        c.append({'p1': [1, 2], 'p2': 3})
    c.expand(1)
    c.append({'p1': [1, 2], 'p2': 3})
    assert len(c) == 2
    assert np.all(c['p1'] == [[1, 2], [1, 2]])
    with pytest.raises(KeyError):
        c['p3']
    assert repr(c) == "Chain(params: ('p1', 'p2'), size: 2)"
    v = Chain.from_arrays({'a': np.ones((4, 2)), 't': np.arange(4.0)})
    assert len(v) == 4 and v.full.shape == (4, 3) and v['t'][3] == 3


def test_data():
    from occuspytial_b200.data import Data
    dic = {1: np.random.rand(5), 2: np.random.rand(4), 3: np.random.rand(2)}
    d = Data(dic)
    assert np.all(d.surveyed == [1, 2, 3]) and len(d) == 3
    assert np.all(d.visits([1, 3]) == (5, 2)) and np.all(d.visits((1, 3)) == (5, 2)) and d.visits(3) == 2
    assert d[1] is dic[1]
    assert np.allclose(d[[1, 3]], np.concatenate((dic[1], dic[3]))) and d[(1, 3)].size == 7
    dic = {3: np.random.rand(2, 2), 1: np.random.rand(5, 2)}
    d = Data(dic)
    assert d[[1, 3]].shape == (7, 2)
    rows, ptr = d.to_csr(5)
    assert ptr.tolist() == [0, 0, 5, 5, 7, 7] and np.allclose(rows[:5], dic[1]) and np.allclose(rows[5:], dic[3])
    import pickle
    assert len(pickle.loads(pickle.dumps(d))) == 2


def test_state():
    from occuspytial_b200.state import State, FixedState
    s = State()
    s.a, s.b = 1, 2
    with pytest.raises(TypeError, match='does not support item assignment'):
        s['b'] = 2
    s.a = 0.5
    assert s.a == 0.5 and [i for i in s] == ['a', 'b'] and s['b'] == 2
    fs = FixedState()
    fs.c = 3
    with pytest.raises(KeyError, match='cannot change attributes already set'):
        fs.c = 2


def test_utils():
    from occuspytial_b200.utils import get_generator, rand_precision_mat, make_data
    assert isinstance(get_generator().bit_generator, np.random.SFC64)
    assert get_generator(1).random() == get_generator(1).random()
    Q = rand_precision_mat(2, 4)
    assert Q.diagonal().max() == 5 and np.linalg.matrix_rank(Q.toarray()) == 7
    assert rand_precision_mat(2, 4, max_neighbors=4).diagonal().max() == 3
    assert np.linalg.matrix_rank(rand_precision_mat(2, 4, rho=0.9).toarray()) == 8
    assert rand_precision_mat(10, 5).nnz == 364          # the reference's docstring known answer (utils.py:72-75)
    with pytest.raises(ValueError, match='Maximum number of neighbors'):
        rand_precision_mat(2, 4, max_neighbors=5)
    Q, W, X, y, alpha, beta, tau, z = make_data(random_state=0)
    assert Q.shape == (150, 150) and X.shape == (150, 3) and len(W) == 75 and alpha.shape == (3,)
    assert all(W[s].shape == (y[s].shape[0], 3) for s in W)
    for kw, msg in ((dict(n=100), 'n cant be lower'), (dict(min_v=0), 'min_v needs'), (dict(max_v=1), 'max_v is too small'),
                    (dict(max_v=151), 'max_v cant be more'), (dict(ns=0), 'ns should be positive'), (dict(ns=151), 'ns cant be more')):
        with pytest.raises(ValueError, match=msg):
            make_data(**kw)


def test_posterior_container_and_mock_arviz():
    from occuspytial_b200.chain import Chain
    from occuspytial_b200 import posterior
    c = Chain({'p1': 2, 'p2': 1}, size=2)
    p = posterior.PosteriorParameter(c)
    assert p['p1'].shape == (1, 0, 2)
    with patch('occuspytial_b200.posterior.az') as az:
        p.summary
        assert az.summary.called
        for meth, target in (('plot_trace', 'plot_trace'), ('plot_auto_corr', 'plot_autocorr'), ('plot_pair', 'plot_pair'),
                             ('plot_density', 'plot_posterior'), ('plot_ess', 'plot_ess')):
            getattr(p, meth)()
            assert getattr(az, target).called
    rng = np.random.default_rng(0)
    pp = posterior.PosteriorParameter.from_arrays({'a': rng.normal(size=(4, 500, 2)), 't': rng.normal(size=(4, 500))})
    if posterior.az is None:
        df = pp.summary
        assert list(df.index) == ['a[0]', 'a[1]', 't'] and abs(df.loc['t', 'r_hat'] - 1) < 0.05
        assert df.loc['t', 'ess_bulk'] > 1000


def test_diagnostics_against_known_cases():
    from occuspytial_b200.diagnostics import summaries_from_draws, split_rhat, ess
    rng = np.random.default_rng(1)
    x = rng.normal(size=(8, 2000))
    s = summaries_from_draws(x, 100)
    assert abs(split_rhat(s, 2000) - 1) < 0.01 and 12000 < ess(s, 2000) < 20000
    ar = np.zeros((8, 4000))
    e = rng.normal(size=ar.shape)
    for t in range(1, 4000):
        ar[:, t] = 0.9 * ar[:, t - 1] + e[:, t]
    n_eff = ess(summaries_from_draws(ar, 300), 4000)
    assert 0.6 < n_eff / (8 * 4000 * (1 - 0.9) / (1 + 0.9)) < 1.5
    shifted = x + np.arange(8)[:, None]
    assert split_rhat(summaries_from_draws(shifted, 10), 2000) > 1.5


def test_reduced_summary_blocks_and_lazy_chain_summaries():
    """Cross-chain reduction of the per-chain blocks (diagnostics.reduce_blocks: NumPy and torch give the same numbers)
    against the textbook formulas evaluated on the draws, and ChainSummaries, which keeps the blocks where they are."""
    import torch
    from occuspytial_b200 import posterior
    from occuspytial_b200.diagnostics import (ChainSummaries, summaries_from_draws, reduce_blocks, split_rhat_reduced,
                                              pooled_moments_reduced, ess_reduced, ess, summary_frame)
    rng = np.random.default_rng(5)
    chains, draws, L = 7, 400, 60
    x = np.zeros((chains, draws, 2))
    e = rng.normal(size=x.shape)
    for t in range(1, draws):
        x[:, t] = 0.7 * x[:, t - 1] + e[:, t]
    x[:, :, 1] += 0.3 * np.arange(chains)[:, None]
    blocks = np.stack([summaries_from_draws(x[:, :, k], L) for k in range(2)], axis=1)      # (chains, 2, 6 + L + 1)
    red = reduce_blocks(blocks)
    np.testing.assert_allclose(reduce_blocks(torch.from_numpy(blocks)).numpy(), red, rtol=1e-13, atol=1e-15)
    half = draws // 2
    for k in range(2):
        xs = x[:, :, k]
        halves = np.concatenate([xs[:, :half], xs[:, draws - half:]])
        Wv, B = halves.var(axis=1, ddof=1).mean(), half * halves.mean(axis=1).var(ddof=1)
        rhat = np.sqrt(((half - 1) / half * Wv + B / half) / Wv)
        np.testing.assert_allclose(split_rhat_reduced(red[k], draws), rhat, rtol=1e-12)
        mean, sd = pooled_moments_reduced(red[k], draws, chains)
        np.testing.assert_allclose([mean, sd], [xs.mean(), xs.std(ddof=1)], rtol=1e-12)
        np.testing.assert_allclose(ess_reduced(red[k], draws, chains), ess(blocks[:, k], draws), rtol=1e-13)
    packed = torch.from_numpy(np.concatenate([blocks, blocks[:, :1]], axis=1))               # 'a' (dim 2) and 't' (dim 1)
    cs = ChainSummaries(['a', 't'], [2, 1], chains, packed)
    assert cs._blocks is None and 'a' in cs and 'z' not in cs and cs.keys() == ['a', 't']
    np.testing.assert_allclose(cs.reduced['a'], red, rtol=1e-13, atol=1e-15)
    if posterior.az is None:
        data = {'a': x, 't': x[:, :, 0]}
        lazy = summary_frame(data, summaries=cs)
        assert cs._blocks is None                       # the summary did not need the per-chain blocks
        plain = summary_frame(data, max_lag=L)
        np.testing.assert_allclose(lazy[['mean', 'sd', 'ess_bulk', 'r_hat']].to_numpy(),
                                   plain[['mean', 'sd', 'ess_bulk', 'r_hat']].to_numpy(), rtol=1e-10)
    assert cs['a'].shape == (chains, 2, 6 + L + 1) and cs['t'].shape == (chains, 1, 6 + L + 1) and cs._blocks is not None
    np.testing.assert_array_equal(cs['a'], blocks)


def test_engine_slots_are_exclusive_per_device():
    """engine._acquire_slot / _release_slot: the lowest free slot of a device, never one that is in use (two engines alive
    at the same time must not share a stream or a page-locked staging buffer), also under concurrent callers."""
    from concurrent.futures import ThreadPoolExecutor
    from occuspytial_b200 import engine
    dev = 'cuda:77'   # a name no other test uses
    a, b = engine._acquire_slot(dev), engine._acquire_slot(dev)
    assert (a, b) == (0, 1) and engine._acquire_slot('cuda:78') == 0
    engine._release_slot(dev, a)
    assert engine._acquire_slot(dev) == 0 and engine._acquire_slot(dev) == 2
    with ThreadPoolExecutor(8) as pool:
        got = list(pool.map(lambda _: engine._acquire_slot(dev), range(64)))
    assert len(set(got)) == 64 and not set(got) & {0, 1, 2}
    for slot in got + [0, 1, 2]:
        engine._release_slot(dev, slot)
    engine._release_slot('cuda:78', 0)
    assert engine._acquire_slot(dev) == 0
    engine._release_slot(dev, 0)


Q, W, X, y, *_ = (None,) * 5


@pytest.fixture(scope='module')
def data():
    from occuspytial_b200.utils import make_data
    return make_data(min_v=2, max_v=10, ns=100, p=3, q=2, random_state=10)[:4]


def test_sampler_configuration_matches_oracle_partition(data):
    from occuspytial_b200 import LogitRSRGibbs, LogitICARGibbs, ProbitRSRGibbs, PGGibbs
    from oracle import sweep as S
    Q, W, X, y = data
    s = LogitRSRGibbs(Q, W, X, y, random_state=10)
    part = S.partition_sites(y, 150)
    assert s.fixed.not_obs == part['not_obs'] and s.fixed.obs == part['obs'] and s.fixed.not_surveyed == part['not_surveyed']
    np.testing.assert_array_equal(s.state.z, part['z'])
    K, Qrsr = S.moran_basis(Q, X, r=0.5)
    assert s.fixed.q == K.shape[1] and s.fixed.tau_shape == 0.5 + 0.5 * s.fixed.q
    np.testing.assert_allclose(np.linalg.eigvalsh(s.fixed.Q), np.linalg.eigvalsh(Qrsr), rtol=1e-8)
    assert s.fixed.stacked_w_indices[0] == 0 and len(s.fixed.stacked_w_indices) == s.fixed.n_no
    assert s.fixed.W_not_obs.shape[0] == sum(s.fixed.visits_not_obs)
    prob = s._problem_arrays()
    assert prob['visit_ptr'][-1] == prob['W_rows'].shape[0] == prob['y_rows'].shape[0]
    i = LogitICARGibbs(Q, W, X, y)
    assert i.fixed.tau_shape == 0.5 + 0.5 * 149 and PGGibbs is LogitICARGibbs
    pi = i._problem_arrays()
    perm, bw = i._band()
    Qp = Q.tocsr()[perm][:, perm].toarray()
    assert pi['Qband'].shape == (150, bw + 1) and np.allclose(pi['Qband'][:, 0], np.diag(Qp))
    assert np.allclose(pi['Qband'][bw:, bw], np.diag(Qp, -bw))
    pr = ProbitRSRGibbs(Q, W, X, y, q=10)
    assert pr.fixed.KTK.shape == (10, 10) and pr.fixed.q == 10
    assert isinstance(s.copy(), LogitRSRGibbs) and s.copy().rng is not s.rng


def test_hyperparameters_and_errors(data):
    from occuspytial_b200 import LogitRSRGibbs, LogitICARGibbs, ProbitRSRGibbs
    from occuspytial_b200.gibbs.base import GibbsBase
    Q, W, X, y = data
    rng = np.random.default_rng(10)
    hypers = {'tau_rate': 1.0, 'tau_shape': 5.0, 'a_mu': rng.random(2), 'b_mu': rng.random(3), 'a_prec': np.eye(2), 'b_prec': np.eye(3)}
    for cls in (LogitRSRGibbs, LogitICARGibbs, ProbitRSRGibbs):
        s1, s2 = cls(Q, W, X, y), cls(Q, W, X, y, hparams=hypers)
        assert s1.fixed['tau_shape'] != s2.fixed['tau_shape'] and s1.fixed.tau_rate != s2.fixed.tau_rate
        assert not np.allclose(s1.fixed.a_prec, s2.fixed.a_prec) and not np.allclose(s1.fixed.b_mu, s2.fixed.b_mu)
        mat = rand(150, 150, density=0.9, format='csc', random_state=10)
        with pytest.raises(ValueError, match='Spatial precision matrix Q must be'):
            cls(mat.T * mat, W, X, y)
        with pytest.raises(ValueError, match='burnin value cannot be larger than'):
            s1.sample(10, burnin=11)
        with pytest.raises(ValueError, match='chains must a positive integer'):
            s1.sample(10, chains=0)
    for cls in (LogitRSRGibbs, ProbitRSRGibbs):
        with pytest.raises(ValueError, match='Threshold value needs to be in'):
            cls(Q, W, X, y, r=1.1)

    class FakeSampler(GibbsBase):
        def __init__(self, Q, W, X, y):
            super().__init__(Q, W, X, y)
            super()._configure(Q, None)

    with pytest.raises(NotImplementedError, match='FakeSampler must implement a `step` method.'):
        FakeSampler(Q, W, X, y).sample(5)


def test_large_lattice_singularity_check_and_basis():
    """The reference's shift-invert check rejects valid ICAR matrices with lambda_2 < 1e-3."""
    from occuspytial_b200.utils import rand_precision_mat
    from occuspytial_b200.gibbs import _setup
    Q = rand_precision_mat(60, 90).astype(float)          # lambda_2 ~ 1e-3 territory
    _setup.verify_singular(Q)
    X = np.column_stack([np.ones(5400), np.random.default_rng(0).uniform(-2, 2, (5400, 2))])
    K = _setup.rsr_basis(Q, X, q=8)
    assert K.shape == (5400, 8) and np.allclose(K.T @ K, np.eye(8), atol=1e-8) and np.abs(X.T @ K).max() < 1e-8


def test_no_cuda_device_fails_loudly(data):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from occuspytial_b200 import LogitRSRGibbs, _lib
    Q, W, X, y = data
    with pytest.raises(_lib.EngineError, match='no CPU fallback'):
        LogitRSRGibbs(Q, W, X, y, q=5).sample(3, progressbar=False)


def test_abi_library_exports_every_declared_symbol():
    """The shared library loads on a GPU-less host and exports every function include/occu_b200.h declares;
    the ctypes prototypes cover exactly that set.  No compute call is made."""
    from occuspytial_b200 import _lib
    header = open(os.path.join(ROOT, 'include', 'occu_b200.h')).read()
    declared = set(re.findall(r'\b(occu_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'occuspytial_b200', 'csrc')])
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().occu_abi_version() == _lib.ABI_VERSION
    # argument validation happens before any CUDA call
    P = _lib.Problem()
    assert _lib.load().occu_workspace_bytes(ctypes.byref(P)) == 0
    assert b'ABI version' in _lib.load().occu_last_error()
    P.abi_version, P.model, P.n, P.p, P.r, P.q, P.V, P.chains = _lib.ABI_VERSION, 0, 1000, 3, 3, 64, 2500, 4096
    nbytes = _lib.load().occu_workspace_bytes(ctypes.byref(P))
    assert 2e8 < nbytes < 2e9


def test_engine_limits_and_input_checks_raise_before_any_device_work():
    """Limits of the engine surface as ValueError on the host (no GPU needed): ICAR half bandwidth, covariate counts, RSR
    rank; W and y must describe the same visits (the reference fails with a shape error in W.T @ y, logit.py:223)."""
    from occuspytial_b200 import LogitICARGibbs
    from occuspytial_b200.gibbs.base import _check_visits_match
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(n=150, min_v=1, max_v=4, lattice=(10, 15), random_state=1)
    s = LogitICARGibbs(Q, W, X, y, random_state=1)
    perm, bw = s._band()
    assert 10 <= bw <= 17
    prob = s._problem_arrays()
    assert prob['icar_bw'] == bw and prob['Qband'].shape == (150, bw + 1)
    s.MAX_HALF_BANDWIDTH = bw - 1
    with pytest.raises(ValueError, match='half bandwidth'):
        s._problem_arrays()
    y_bad = dict(y)
    site = next(iter(y_bad))
    y_bad[site] = np.concatenate([y_bad[site], [0]])
    s2 = LogitICARGibbs(Q, W, X, y_bad, random_state=1)
    with pytest.raises(ValueError, match='same number of visits per site'):
        s2._problem_arrays()
    _check_visits_match(np.array([0, 2, 5]), np.array([0, 2, 5]))
    with pytest.raises(ValueError, match='first mismatch at site 1'):
        _check_visits_match(np.array([0, 2, 5]), np.array([0, 2, 4]))
