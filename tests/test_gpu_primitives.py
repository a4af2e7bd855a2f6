"""GPU parity tests of the stand-alone primitives, through the C ABI."""
import ctypes as C

import numpy as np
import pytest
from scipy import stats

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def env():
    torch = pytest.importorskip('torch')
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    from occuspytial_b200 import _lib
    return torch, _lib.load(), _lib


def _pg(env, c, seed, fp64):
    torch, lib, _lib = env
    tc = torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64)).cuda()
    out = torch.empty_like(tc)
    _lib.check(lib.occu_pg1(tc.data_ptr(), tc.numel(), seed, int(fp64), out.data_ptr(), None), 'occu_pg1')
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize('fp64', [False, True])
@pytest.mark.parametrize('c', [0.0, 0.3, 1.0, 2.5, 3.2, 6.0, 15.0, 40.0])
def test_pg1_vs_oracle(env, c, fp64):
    """KS and moment tests of the device PG(1,c) sampler against the oracle's draws and the analytic moments."""
    from oracle import sweep as S
    N = 400_000
    x = _pg(env, np.full(N, c), seed=1234 + int(10 * c), fp64=fp64)
    assert np.all(np.isfinite(x)) and np.all(x > 0)
    m, v = float(S.pg1_mean(c)), float(S.pg1_var(c))
    assert abs(x.mean() - m) < 5 * np.sqrt(v / N), (x.mean(), m)
    assert abs(x.var() - v) < 0.04 * v, (x.var(), v)
    ref = S.pg1(np.full(N, c), seed=99)
    assert stats.ks_2samp(x, ref).pvalue > 1e-3
    ser = S.pg1_series(np.full(N // 4, c), seed=98, terms=300)
    assert stats.ks_2samp(x, ser).pvalue > 1e-3


def test_pg1_mixed_arguments_and_determinism(env):
    from oracle import sweep as S
    rng = np.random.default_rng(0)
    c = rng.normal(size=1_000_000) * 3
    x1, x2 = _pg(env, c, 7, False), _pg(env, c, 7, False)
    np.testing.assert_array_equal(x1, x2)
    assert not np.array_equal(x1, _pg(env, c, 8, False))
    # standardised residuals must have mean 0 / variance 1 over the mixture of arguments
    res = (x1 - S.pg1_mean(c)) / np.sqrt(S.pg1_var(c))
    assert abs(res.mean()) < 5e-3 and abs(res.var() - 1) < 2e-2


def test_truncnorm_matches_reference_formula_and_scipy(env):
    torch, lib, _lib = env
    from oracle import sweep as S
    rng = np.random.default_rng(1)
    N = 200_000
    loc = rng.normal(size=N) * 2
    pos = (rng.random(N) < 0.5)
    u = rng.random(N)
    t = lambda a, dt=np.float64: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).cuda()
    out = torch.empty(N, dtype=torch.float64, device='cuda')
    tl, tp, tu = t(loc), t(pos, np.uint8), t(u)
    _lib.check(lib.occu_truncnorm(tl.data_ptr(), tp.data_ptr(), tu.data_ptr(), N, 0, 1, out.data_ptr(), None), 'tn')
    got = out.cpu().numpy()
    ref = S.truncnorm_draw(loc, pos, u)   # probit.py:11-24 formula
    np.testing.assert_allclose(got, ref, rtol=1e-9, atol=1e-9)
    # Philox path: KS against scipy.stats.truncnorm, including far tails where the reference formula breaks
    for fp64 in (1, 0):
        for L, positive in ((0.5, True), (-1.5, True), (2.0, False), (-4.5, True), (-9.0, True), (12.0, False)):
            loc1 = np.full(N, L)
            tl, tp = t(loc1), t(np.full(N, positive), np.uint8)
            _lib.check(lib.occu_truncnorm(tl.data_ptr(), tp.data_ptr(), None, N, 5, fp64, out.data_ptr(), None), 'tn')
            x = out.cpu().numpy()
            assert np.all(x >= 0) if positive else np.all(x <= 0)
            a, b = ((0 - L), np.inf) if positive else (-np.inf, (0 - L))
            assert stats.kstest(x, stats.truncnorm(a, b, loc=L).cdf).pvalue > 1e-3, (fp64, L)


def test_gamma(env):
    torch, lib, _lib = env
    N = 300_000
    for shape, rate in ((0.3, 2.0), (1.0, 0.5), (5.5, 3.0), (75.0, 10.0)):
        ts = torch.full((N,), shape, dtype=torch.float64, device='cuda')
        tr = torch.full((N,), rate, dtype=torch.float64, device='cuda')
        out = torch.empty_like(ts)
        _lib.check(lib.occu_gamma(ts.data_ptr(), tr.data_ptr(), N, 3, out.data_ptr(), None), 'gamma')
        x = out.cpu().numpy()
        assert stats.kstest(x, stats.gamma(shape, scale=1 / rate).cdf).pvalue > 1e-3


def test_precision_mvnorm_vs_reference_golden(env):
    """Batched device precision_mvnorm against the known answers of the reference's native routine."""
    from occuspytial_b200.distributions import precision_mvnorm_batched
    g = golden('native.npz')
    for d in (3, 10, 64):
        x, L = precision_mvnorm_batched(g[f'b{d}'][None], g[f'P{d}'][None], g[f'z{d}'][None])
        np.testing.assert_allclose(x[0], g[f'x{d}'], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(L[0], np.tril(g[f'Pafter{d}']), rtol=1e-10, atol=1e-13)


def test_precision_mvnorm_batched_vs_oracle_and_api(env):
    from occuspytial_b200.distributions import precision_mvnorm_batched, precision_mvnorm, ensure_sums_to_zero
    from oracle import sweep as S
    rng = np.random.default_rng(2)
    for d in (1, 2, 7, 33, 128, 200):
        m = 9
        A = rng.normal(size=(m, d + 3, d))
        P = np.einsum('mki,mkj->mij', A, A) + 0.05 * np.eye(d)
        b, z = rng.normal(size=(m, d)), rng.normal(size=(m, d))
        x, L = precision_mvnorm_batched(b, P, z)
        for i in range(m):
            xo, Lo = S.precision_mvnorm(b[i], P[i], z[i])
            np.testing.assert_allclose(x[i], xo, rtol=1e-9, atol=1e-11)
            np.testing.assert_allclose(L[i], Lo, rtol=1e-10, atol=1e-12)
    # reference API behaviour (occuspytial/tests/test_distributions.py): output shape, prec overwritten, reproducible
    mat = rng.random((5, 5))
    prec = np.linalg.inv(mat.T @ mat)
    pc = prec.copy()
    arr = precision_mvnorm(np.zeros(5), pc)
    assert arr.shape == (5,) and not np.allclose(prec, pc)
    a1 = precision_mvnorm(np.zeros(5), prec.copy(), random_state=np.random.default_rng(0))
    a2 = precision_mvnorm(np.zeros(5), prec.copy(), random_state=np.random.default_rng(0))
    assert np.allclose(a1, a2) and not np.allclose(a1, arr)
    with pytest.raises(RuntimeError, match='Cholesky factorization/solver failed!'):
        precision_mvnorm(np.zeros(3), -np.eye(3))
    g = golden('native.npz')
    out = np.zeros(50)
    ensure_sums_to_zero(g['stz_x'], g['stz_z'], out)
    np.testing.assert_allclose(out, g['stz_out'], rtol=1e-12, atol=1e-14)


def test_gemm_tn_vs_torch_fp64(env):
    torch, lib, _lib = env
    gen = torch.Generator(device='cuda').manual_seed(0)
    for M, N, K in ((128, 128, 16), (256, 384, 80), (512, 2432, 1024)):
        A = torch.randn(K, M, dtype=torch.float64, device='cuda', generator=gen)
        B = torch.randn(K, N, dtype=torch.float64, device='cuda', generator=gen)
        out = torch.empty(M, N, dtype=torch.float64, device='cuda')
        _lib.check(lib.occu_gemm_tn(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K, M, N, N, None), 'gemm')
        ref = A.T @ B
        err = (out - ref).abs().max().item() / ref.abs().max().item()
        assert err < 1e-13, (M, N, K, err)


def test_chain_summaries_kernel_vs_host(env):
    """Per-chain summary blocks (the payload of the NCCL all-gather) against the NumPy definition."""
    torch, lib, _lib = env
    from occuspytial_b200.diagnostics import summaries_from_draws
    rng = np.random.default_rng(4)
    chains, draws, dim, L = 37, 501, 3, 40
    rec = np.cumsum(rng.normal(size=(chains, draws, dim)), axis=1) * 0.05 + rng.normal(size=(chains, draws, dim))
    t = torch.from_numpy(rec).cuda()
    out = torch.empty((chains, dim, 6 + L + 1), dtype=torch.float64, device='cuda')
    _lib.check(lib.occu_chain_summaries(t.data_ptr(), chains, draws, dim, L, out.data_ptr(), None), 'summaries')
    got = out.cpu().numpy()
    for k in range(dim):
        np.testing.assert_allclose(got[:, k], summaries_from_draws(rec[:, :, k], L), rtol=1e-10, atol=1e-12)
