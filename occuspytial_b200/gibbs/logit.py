"""Logit-link samplers: LogitICARGibbs (alias PGGibbs) and LogitRSRGibbs.

API of occuspytial/gibbs/logit.py:102-485; the sweep itself (``step``: omega_b, tau, eta, beta, omega_a,
alpha, z -- logit.py:254-266) runs in the CUDA engine, see occuspytial_b200/csrc/ (engine.cu and the headers it includes).
"""
import numpy as np

from . import _setup
from .base import GibbsBase, _check_visits_match


class LogitICARGibbs(GibbsBase):
    r"""Gibbs sampler using a logit link and the ICAR model for the spatial random effects.

    Same constructor as the reference (logit.py:172-174).  eta | . ~ N(Lambda^-1 b, Lambda^-1) with
    Lambda = tau Q + diag(omega_b), conditioned on 1'eta = 0.  The reference perturbs b with 2n - 1
    normals through a dense eigen-factor of Q and solves with MINRES at rtol 1e-5 (logit.py:73-99);
    the engine factorises Lambda exactly with a fixed-pattern banded Cholesky per chain.

    Limits of the engine (checked before an engine is created, ``ValueError``): the sites are reordered by reverse
    Cuthill-McKee and the half bandwidth of the reordered Q must not exceed 200 -- the sliding window of the
    factorisation lives in shared memory.  A queen lattice with sides a <= b has half bandwidth about a + 1, so lattices
    up to ~198 sites on the short side fit; the reference's MINRES path has no such limit (but costs O(n) dense memory for
    its eigen-factor).  At most 32 occupancy covariates and 8 detection covariates.
    """

    _model = 'logit_icar'
    MAX_HALF_BANDWIDTH = 200

    def __init__(self, Q, W, X, y, hparams=None, random_state=None):
        super().__init__(Q, W, X, y, hparams, random_state)
        self._configure(Q, hparams)

    def _configure(self, Q, hparams):
        super()._configure(Q, hparams)
        self._perm = None   # symbolic analysis (band ordering) is done lazily, only the ICAR engine needs it

    # ---- engine boundary: the ICAR engine works in a bandwidth-reducing site order ----
    def _band(self):
        if self._perm is None:
            self._perm, self._bw = _setup.band_order(self.fixed.Q)
            self._inv = np.empty_like(self._perm)
            self._inv[self._perm] = np.arange(self._perm.size)
        return self._perm, self._bw

    def _sites_to_engine(self, arr):
        perm, _ = self._band()
        return np.ascontiguousarray(np.asarray(arr)[..., perm])

    def _sites_from_engine(self, arr):
        self._band()
        return np.ascontiguousarray(np.asarray(arr)[..., self._inv])

    def _start_to_engine(self, arrays):
        out = dict(arrays)
        out['eta'] = self._sites_to_engine(arrays['eta'])
        return out

    def _eta_from_engine(self, eta):
        return self._sites_from_engine(eta)

    def _problem_arrays(self):
        from ..data import Data
        perm, bw = self._band()
        if bw > self.MAX_HALF_BANDWIDTH:
            raise ValueError(f'half bandwidth {bw} of the reordered spatial precision matrix exceeds the engine limit of '
                             f'{self.MAX_HALF_BANDWIDTH} (LogitICARGibbs keeps the factorisation window in shared memory)')
        f = self.fixed
        n = f.n
        inv = self._inv
        W_rows, visit_ptr = Data({int(inv[s]): self.W[s] for s in self.W.surveyed}).to_csr(n)
        y_rows, y_ptr = Data({int(inv[s]): self.y[s] for s in self.y.surveyed}).to_csr(n, dtype=np.uint8)
        _check_visits_match(visit_ptr, y_ptr)
        return dict(X=np.ascontiguousarray(self.X[perm]), W_rows=W_rows.reshape(W_rows.shape[0], -1), y_rows=y_rows,
                    visit_ptr=visit_ptr, site_class=np.ascontiguousarray(f.site_class[perm]),
                    Qband=_setup.lower_band(f.Q, perm, bw), icar_bw=bw,
                    tau_rate=f.tau_rate, tau_shape=f.tau_shape,
                    a_mu=f.a_mu, a_prec=f.a_prec, b_mu=f.b_mu, b_prec=f.b_prec)


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

    _DEFAULT_START = (5.0, False)   # eta ~ N(0, 25 I_q), not centred (logit.py:464)

    def _eta_start(self, chains, rng=None):
        """logit.py:462-466: eta ~ N(0, 25 I_q)."""
        rng = self.rng if rng is None else rng
        return rng.normal(scale=5, size=(chains, self.fixed.q))

    def _spatial_from_eta(self, eta):
        return self.fixed.K @ eta

    # the RSR engine keeps the natural site order
    def _sites_to_engine(self, arr):
        return arr

    def _sites_from_engine(self, arr):
        return arr

    def _start_to_engine(self, arrays):
        return arrays

    def _eta_from_engine(self, eta):
        return eta

    def _problem_arrays(self):
        prob = GibbsBase._problem_arrays(self)
        prob['K'] = self.fixed.K
        prob['Qrsr'] = self.fixed.Q
        return prob


PGGibbs = LogitICARGibbs
