"""The C restatement of the PG(1,c) sampler (oracle/pg_devroye.c) against analytic moments, the Laplace
transform and the independent sum-of-exponentials representation.  CPU only."""
import numpy as np
import pytest
from scipy import stats

from oracle import sweep as S

N = 200_000


@pytest.mark.parametrize('c', [0.0, 0.3, 1.0, 2.5, 3.2, 6.0, 15.0, 40.0])
def test_pg1_moments_and_ks(c):
    x = S.pg1(np.full(N, c), seed=int(c * 10) + 1)
    m, v = float(S.pg1_mean(c)), float(S.pg1_var(c))
    assert abs(x.mean() - m) < 5 * np.sqrt(v / N)
    assert abs(x.var() - v) < 0.05 * v
    y = S.pg1_series(np.full(N, c), seed=int(c * 10) + 2, terms=300)
    assert stats.ks_2samp(x, y).pvalue > 1e-3
    # Laplace transform E exp(-t X) = cosh(c/2) / cosh(sqrt((c^2/2 + t) / 2)) at t = 1
    lt = np.cosh(c / 2) / np.cosh(np.sqrt((c * c / 2 + 1.0) / 2)) if c < 30 else np.exp(c / 2 - np.sqrt((c * c / 2 + 1) / 2))
    assert abs(np.exp(-x).mean() - lt) < 5 * np.exp(-x).std() / np.sqrt(N) + 1e-12


def test_pg1_is_even_in_c_and_deterministic():
    c = np.linspace(-5, 5, 1001)
    np.testing.assert_array_equal(S.pg1(c, 5), S.pg1(-c, 5))
    assert not np.array_equal(S.pg1(c, 5), S.pg1(c, 6))
