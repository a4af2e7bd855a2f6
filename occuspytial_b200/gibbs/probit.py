"""Probit-link sampler ProbitRSRGibbs (API of occuspytial/gibbs/probit.py:27-270).

The reference implementation diverges after two iterations because ``precision_mvnorm`` overwrites the
cached matrix ``XTX_plus_bprec`` (probit.py:112, 241-243; distributions.pyx:91,100).  The engine
implements the mathematically intended sweep (the matrix is constant).
"""
import numpy as np
from scipy.special import ndtr, ndtri

from . import _setup
from .base import GibbsBase


def truncnorm_inf_ppf(a, p):
    """PPF of the standard normal truncated to (a, inf) (probit.py:11-16)."""
    return -ndtri(ndtr(-a) * (1.0 - p))


def truncnorm_neginf_ppf(b, p):
    """PPF of the standard normal truncated to (-inf, b) (probit.py:19-24)."""
    return ndtri(ndtr(b) * p)


class ProbitRSRGibbs(GibbsBase):
    """Sampler using a probit link and the RSR model for the spatial random effects (probit.py:103-107)."""

    _model = 'probit_rsr'

    def __init__(self, Q, W, X, y, hparams=None, random_state=None, r=0.5, q=None, K=None):
        super().__init__(Q, W, X, y, hparams, random_state)
        self._configure(Q, hparams, q, r, K)

    def _configure(self, Q, hparams, q, r, K=None):
        """probit.py:109-153."""
        super()._configure(Q, hparams)
        self.state.omega_b = np.zeros(self.fixed.n)
        self.fixed.XTX_plus_bprec = self.X.T @ self.X + np.asarray(self.fixed.b_prec)
        self.fixed.eps_chol_factor = np.ones(self.fixed.n) / np.sqrt(2)
        if K is None:
            K = _setup.rsr_basis(self.fixed.Q, self.X, r=r, q=q)
        K = np.ascontiguousarray(K, dtype=np.float64)
        self.fixed.q = K.shape[1]
        self.fixed.KTK = K.T @ K
        Q_full = self.fixed.Q
        del self.fixed.Q
        self.fixed.Q = np.ascontiguousarray(K.T @ (Q_full @ K))
        self.fixed.K = K
        if not hparams:
            del self.fixed.tau_shape
            self.fixed.tau_shape = 0.5 + 0.5 * self.fixed.q

    _DEFAULT_START = (5.0, False)   # eta ~ N(0, 25 I_q), not centred (probit.py:157)

    def _eta_start(self, chains, rng=None):
        rng = self.rng if rng is None else rng
        return rng.normal(scale=5, size=(chains, self.fixed.q))

    def _draw_start(self, chains, rng=None):
        rng = self.rng if rng is None else rng
        start = super()._draw_start(chains, rng)
        start['eps'] = rng.standard_normal((chains, self.fixed.n))   # probit.py:159
        return start

    def _start_keys(self):
        return ('alpha', 'beta', 'tau', 'eta', 'eps')

    def _spatial_from_eta(self, eta):
        return self.fixed.K @ eta

    def _problem_arrays(self):
        prob = super()._problem_arrays()
        prob['K'] = self.fixed.K
        prob['Qrsr'] = self.fixed.Q
        return prob
