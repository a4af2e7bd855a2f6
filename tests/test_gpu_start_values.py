"""Default start values drawn on the device (occu_draw_start) against the distributions the reference draws them from
(base.py:199-211, logit.py:462-466, probit.py:155-160), bit-level against a NumPy restatement of the engine's
Philox4x32-10 addressing, and independent of how the chains are sharded."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

START_ITER = 0xFFFFFFFF
TAU, ALPHA, BETA, ETA, EPS = range(5)


def philox4x32_10(seed, chain, elem, it, stage, block=0):
    """One 128-bit Philox4x32-10 block at the engine's coordinates (csrc/common.cuh: Philox::init / refill)."""
    m = 0xFFFFFFFF
    c = [chain & m, elem & m, it & m, ((stage << 24) + block) & m]
    k0, k1 = seed & m, (seed >> 32) & m
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k0, p1 & m, (p0 >> 32) ^ c[3] ^ k1, p0 & m]
        k0, k1 = (k0 + 0x9E3779B9) & m, (k1 + 0xBB67AE85) & m
    return c


def normals(seed, chain, elem, stage, count):
    """FP64 Box-Muller normals in the order Philox::normal_d produces them: one value per block of four words."""
    out = []
    for b in range(count):
        w = philox4x32_10(seed, chain, elem, START_ITER, stage, b)
        u = [(((w[2 * i] << 21) ^ (w[2 * i + 1] >> 11)) + 0.5) / 9007199254740992.0 for i in range(2)]
        out.append(np.sqrt(-2.0 * np.log(u[0])) * np.cos(2.0 * np.pi * u[1]))
    return np.array(out)


def _rsr(cls, **kw):
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(n=400, min_v=1, max_v=5, ns=200, p=3, q=2, random_state=4)
    return cls(Q, W, X, y, random_state=7, **kw)


def test_start_values_bitwise_and_sharding():
    from occuspytial_b200 import LogitRSRGibbs, _lib
    s = _rsr(LogitRSRGibbs, q=20)
    f = s.fixed
    seed = 0x1234567890ABCDEF
    eng = s._make_engine(300, seed=1, chain_offset=0)
    eng.draw_start(seed, 5.0, False)
    full = {k: eng.get_field(fld) for k, fld in (('alpha', _lib.F_ALPHA), ('beta', _lib.F_BETA), ('tau', _lib.F_TAU),
                                                 ('eta', _lib.F_ETA))}
    eng.close()
    # bit-level restatement for a few chains
    La = np.linalg.cholesky(100 * np.asarray(f.a_prec))
    Lb = np.linalg.cholesky(100 * np.asarray(f.b_prec))
    for c in (0, 1, 131, 299):
        zt = normals(seed, c, 0, TAU, 1)[0]
        np.testing.assert_allclose(full['tau'][c], 0.5 * zt * zt / f.tau_rate, rtol=1e-13)
        np.testing.assert_allclose(full['alpha'][c], f.a_mu + La @ normals(seed, c, 0, ALPHA, La.shape[0]), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(full['beta'][c], f.b_mu + Lb @ normals(seed, c, 0, BETA, Lb.shape[0]), rtol=1e-12, atol=1e-13)
        eta = np.array([5.0 * normals(seed, c, e, ETA, 1)[0] for e in range(20)])
        np.testing.assert_allclose(full['eta'][c], eta, rtol=1e-12, atol=1e-13)
    # a shard draws the same values as the whole job; another seed draws others
    for lo, hi in ((0, 100), (100, 229), (229, 300)):
        eng = s._make_engine(hi - lo, seed=1, chain_offset=lo)
        eng.draw_start(seed, 5.0, False)
        for k, fld in (('alpha', _lib.F_ALPHA), ('beta', _lib.F_BETA), ('tau', _lib.F_TAU), ('eta', _lib.F_ETA)):
            np.testing.assert_array_equal(eng.get_field(fld), full[k][lo:hi])
        if lo == 0:
            eng.draw_start(seed + 1, 5.0, False)
            assert not np.any(eng.get_field(_lib.F_TAU) == full['tau'][lo:hi])
        eng.close()


@pytest.mark.parametrize('model', ['logit_rsr', 'probit_rsr', 'logit_icar'])
def test_start_value_distributions(model):
    from scipy import stats
    from occuspytial_b200 import LogitRSRGibbs, ProbitRSRGibbs, LogitICARGibbs, _lib
    cls = {'logit_rsr': LogitRSRGibbs, 'probit_rsr': ProbitRSRGibbs, 'logit_icar': LogitICARGibbs}[model]
    s = _rsr(cls, **({} if model == 'logit_icar' else {'q': 12}))
    f = s.fixed
    chains = 4096
    eng = s._make_engine(chains, seed=3)
    eng.draw_start(99, *s._DEFAULT_START)
    tau = np.ravel(eng.get_field(_lib.F_TAU))
    alpha, beta, eta = eng.get_field(_lib.F_ALPHA), eng.get_field(_lib.F_BETA), eng.get_field(_lib.F_ETA)
    # tau ~ Gamma(1/2, scale 1 / tau_rate)   (base.py:201)
    assert stats.kstest(np.ravel(tau) * f.tau_rate, stats.gamma(0.5).cdf).pvalue > 1e-4
    # alpha ~ N(a_mu, 100 a_prec), beta ~ N(b_mu, 100 b_prec)   (base.py:206-211)
    for x, mu, prec in ((alpha, f.a_mu, f.a_prec), (beta, f.b_mu, f.b_prec)):
        cov = 100 * np.asarray(prec)
        np.testing.assert_allclose(x.mean(axis=0), mu, atol=5 * np.sqrt(np.diag(cov).max() / chains))
        np.testing.assert_allclose(np.cov(x.T), cov, atol=0.12 * np.diag(cov).max())
        assert stats.kstest((x[:, 0] - mu[0]) / np.sqrt(cov[0, 0]), stats.norm.cdf).pvalue > 1e-4
    if model == 'logit_icar':      # eta ~ N(0, I_n) minus its mean (base.py:202-203); engine site order is a permutation
        np.testing.assert_allclose(eta.mean(axis=1), 0.0, atol=1e-12)
        np.testing.assert_allclose(eta.var(axis=1).mean(), 1 - 1 / f.n, rtol=5e-3)
    else:                          # eta ~ N(0, 25 I_q)   (logit.py:464, probit.py:157)
        assert stats.kstest(eta.ravel() / 5.0, stats.norm.cdf).pvalue > 1e-4
    if model == 'probit_rsr':      # eps ~ N(0, I_n)   (probit.py:159)
        eps = eng.get_field(_lib.F_EPS)
        assert stats.kstest(eps[::16].ravel(), stats.norm.cdf).pvalue > 1e-4
        assert abs(np.corrcoef(eps[0], eps[1])[0, 1]) < 0.25
    # the engine is re-armed: iteration 0, z from the data
    z = eng.get_field(_lib.F_Z)
    assert set(np.unique(z)) <= {0, 1}
    assert eng.sweep(2) == 0
    eng.close()


def test_sample_default_start_does_not_depend_on_devices_or_sharding():
    """sample() with default start values: the draws of a chain are the same whether the job runs as one engine or as
    shards (the chain id keys both the start values and the sweep's Philox streams)."""
    from occuspytial_b200 import LogitRSRGibbs
    a = _rsr(LogitRSRGibbs, q=10)
    b = _rsr(LogitRSRGibbs, q=10)
    pa = a.sample(12, burnin=2, chains=256, progressbar=False, device='cuda:0')
    seed = int(b.rng.integers(0, 2**63 - 1))          # the seed sample() drew from the same generator state
    h1, _, _ = b._run_shard('cuda:0', 0, 128, seed, None, 12, 2, None, False, None, True)
    h2, _, _ = b._run_shard('cuda:0', 128, 256, seed, None, 12, 2, None, False, None, True)
    np.testing.assert_array_equal(pa['tau'], np.concatenate([h1['tau'], h2['tau']]))
    np.testing.assert_array_equal(pa['beta'], np.concatenate([h1['beta'], h2['beta']]))


def test_concurrent_engines_do_not_share_stream_or_staging():
    """Two engines alive at the same time on one device get different slots (stream, page-locked staging buffers); a
    closed engine's slot is handed out again, so repeated sample() calls keep reusing the same stream and buffers."""
    from occuspytial_b200 import LogitRSRGibbs
    s = _rsr(LogitRSRGibbs, q=8)
    a = s._make_engine(8, seed=1)
    b = s._make_engine(8, seed=1)
    assert a.slot != b.slot and a.stream.cuda_stream != b.stream.cuda_stream
    slot_a, stream_a = a.slot, a.stream.cuda_stream
    a.close()
    c = s._make_engine(8, seed=1)
    assert c.slot == slot_a and c.stream.cuda_stream == stream_a
    b.close()
    c.close()
