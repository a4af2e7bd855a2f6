"""Logit-link samplers: LogitICARGibbs (alias PGGibbs) and LogitRSRGibbs.

API of occuspytial/gibbs/logit.py:102-485; the sweep itself (``step``: omega_b, tau, eta, beta, omega_a,
alpha, z -- logit.py:254-266) runs in the CUDA engine, see occuspytial_b200/csrc/engine.cu.
"""
import numpy as np

from . import _setup
from .base import GibbsBase


class LogitICARGibbs(GibbsBase):
    r"""Gibbs sampler using a logit link and the ICAR model for the spatial random effects.

    Same constructor as the reference (logit.py:172-174).  eta | . ~ N(Lambda^-1 b, Lambda^-1) with
    Lambda = tau Q + diag(omega_b), conditioned on 1'eta = 0.  The reference perturbs b with 2n - 1
    normals through a dense eigen-factor of Q and solves with MINRES at rtol 1e-5 (logit.py:73-99);
    the engine factorises Lambda exactly with a fixed-pattern banded Cholesky per chain.
    """

    _model = 'logit_icar'

    def __init__(self, Q, W, X, y, hparams=None, random_state=None):
        super().__init__(Q, W, X, y, hparams, random_state)
        self._configure(Q, hparams)

    def _configure(self, Q, hparams):
        super()._configure(Q, hparams)


class LogitRSRGibbs(LogitICARGibbs):
    """Gibbs sampler using a logit link and the RSR model for the spatial random effects.

    Same constructor as the reference (logit.py:409-413): ``r`` is the eigenvalue threshold of the Moran
    operator, ``q`` the explicit number of basis columns (``r`` ignored when ``q`` is given).  Extra
    keyword ``K``: a precomputed basis (n x q) to skip the eigen-decomposition (used by benchmarks so
    that both the engine and the CPU baseline see the same basis).
    """

    _model = 'logit_rsr'

    def __init__(self, Q, W, X, y, hparams=None, random_state=None, r=0.5, q=None, K=None):
        super().__init__(Q, W, X, y, hparams, random_state)
        self._configure_rsr(r, q, hparams, K)

    def _configure_rsr(self, r, q, hparams, K=None):
        """logit.py:415-460."""
        if K is None:
            K = _setup.rsr_basis(self.fixed.Q, self.X, r=r, q=q)
        K = np.ascontiguousarray(K, dtype=np.float64)
        self.fixed.q = K.shape[1]
        Q_full = self.fixed.Q
        del self.fixed.Q
        self.fixed.Q = np.ascontiguousarray(K.T @ (Q_full @ K))   # K'QK, dense q x q
        self.fixed.K = K
        if not hparams:
            del self.fixed.tau_shape
            self.fixed.tau_shape = 0.5 + 0.5 * self.fixed.q

    def _eta_start(self, chains):
        """logit.py:462-466: eta ~ N(0, 25 I_q).  The base-class draw is still consumed, as in the reference."""
        super()._eta_start(1)
        return self.rng.normal(scale=5, size=(chains, self.fixed.q))

    def _spatial_from_eta(self, eta):
        return self.fixed.K @ eta

    def _problem_arrays(self):
        prob = super()._problem_arrays()
        prob['K'] = self.fixed.K
        prob['Qrsr'] = self.fixed.Q
        return prob


PGGibbs = LogitICARGibbs
