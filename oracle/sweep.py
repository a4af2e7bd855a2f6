"""oracle/sweep.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

NumPy restatement of the reference's Gibbs sweep (OccuSpytial), every conditional
written as a pure function of injected randomness so that the CUDA engine can be
compared with it to 1e-10, plus plain CPU samplers (``OracleSampler``) that run
the sweep in the reference's order and serve as the CPU baseline in bench.py.

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
``occuspytial_b200`` never does.

Pinning: tests/test_oracle_vs_golden.py checks every function below against
vectors dumped from the real reference (imported from /root/reference with its
own .pyx compiled, see oracle/ref_loader.py and tests/golden/make_golden.py).
The PG(1,c) draws themselves are PARITY UNPINNED (third-party polyagamma is
absent; see oracle/pg_devroye.c).

Citations are ``file:line`` into /root/reference/occuspytial/.
"""
import ctypes
import os

import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular
from scipy.special import expit, ndtr, ndtri

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboccu_oracle.so')
        if not os.path.exists(path):
            raise ImportError(f'{path} missing: run `make -C oracle`')
        lib = ctypes.CDLL(path)
        lib.occu_oracle_pg1_fill.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        lib.occu_oracle_pg1_fill.restype = None
        lib.occu_oracle_pg1_series_fill.argtypes = [ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p,
                                                    ctypes.c_int, ctypes.c_void_p]
        lib.occu_oracle_pg1_series_fill.restype = None
        _LIB = lib
    return _LIB


# --------------------------------------------------------------------------------------
# samplers of the augmentation variables
# --------------------------------------------------------------------------------------
def pg1(c, seed):
    """PG(1, c) draws (logit.py:191-193, 202-204 -> polyagamma.random_polyagamma(1, c))."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.empty_like(c)
    _lib().occu_oracle_pg1_fill(int(seed) & (2**64 - 1), c.size, c.ctypes.data, out.ctypes.data)
    return out


def pg1_series(c, seed, terms=200):
    """Independent sum-of-exponentials representation of PG(1, c) (cross-check only)."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.empty_like(c)
    _lib().occu_oracle_pg1_series_fill(int(seed) & (2**64 - 1), c.size, c.ctypes.data, int(terms), out.ctypes.data)
    return out


def pg1_mean(c):
    """E[PG(1,c)] = tanh(c/2) / (2c), 1/4 at 0."""
    c = np.abs(np.asarray(c, dtype=np.float64))
    small = c < 1e-6
    cs = np.where(small, 1.0, c)
    return np.where(small, 0.25 - c * c / 48.0, np.tanh(cs / 2) / (2 * cs))


def pg1_var(c):
    """Var[PG(1,c)] = (sinh c - c) / (4 c^3 cosh^2(c/2)), 1/24 at 0."""
    c = np.abs(np.asarray(c, dtype=np.float64))
    small = c < 1e-2
    cs = np.where(small, 1.0, c)
    big = (np.sinh(cs) - cs) / (4 * cs ** 3 * np.cosh(cs / 2) ** 2)
    return np.where(small, 1.0 / 24.0 - c * c / 120.0, big)


def truncnorm_inf_ppf(a, p):
    """probit.py:11-16 -- PPF of N(0,1) truncated to (a, inf)."""
    return -ndtri(ndtr(-a) * (1.0 - p))


def truncnorm_neginf_ppf(b, p):
    """probit.py:19-24 -- PPF of N(0,1) truncated to (-inf, b)."""
    return ndtri(ndtr(b) * p)


def truncnorm_draw(loc, positive, u):
    """One-sided unit-variance truncated normal around ``loc`` (probit.py:185-194, 203-209).

    positive=True  -> support (0, inf):  loc + ppf_(−loc, inf)(u)
    positive=False -> support (-inf, 0): loc + ppf_(−inf, −loc)(u)
    """
    loc = np.asarray(loc, dtype=np.float64)
    positive = np.asarray(positive, dtype=bool)
    return np.where(positive, truncnorm_inf_ppf(-loc, u) + loc, truncnorm_neginf_ppf(-loc, u) + loc)


# --------------------------------------------------------------------------------------
# setup (base.py:107-164; logit.py:415-460; probit.py:109-153)
# --------------------------------------------------------------------------------------
def partition_sites(y, n):
    """base.py:113-137 -- initial z and the obs / not_obs / not_surveyed partition (dict order kept)."""
    surveyed = list(y)
    z = np.ones(n)
    z[surveyed] = [bool(np.any(y[s])) for s in surveyed]
    in_survey = np.zeros(n, dtype=bool)
    in_survey[surveyed] = True
    not_surveyed = [int(i) for i in np.nonzero(~in_survey)[0]]
    not_obs = [int(i) for i in surveyed if not z[i]]
    obs = [int(i) for i in surveyed if z[i]]
    return dict(surveyed=[int(s) for s in surveyed], z=z, not_surveyed=not_surveyed, not_obs=not_obs, obs=obs)


def pack_visits(W, y, n):
    """Flatten the ragged dict-of-arrays (data.pyx:117-140 semantics) into CSR over surveyed sites.

    Returns W_rows (V x r), y_rows (V,), rowptr (n_s + 1,), site_ids (n_s,) in dict insertion order.
    """
    sites = list(W)
    counts = np.array([W[s].shape[0] for s in sites], dtype=np.int64)
    rowptr = np.zeros(len(sites) + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    W_rows = np.concatenate([np.asarray(W[s], dtype=np.float64).reshape(W[s].shape[0], -1) for s in sites], axis=0)
    y_rows = np.concatenate([np.asarray(y[s], dtype=np.float64).ravel() for s in sites])
    return W_rows, y_rows, rowptr, np.asarray(sites, dtype=np.int64)


def default_hparams(n, r, p, q_rsr=None):
    """base.py:177-186 (+ logit.py:454-457 / probit.py:150-153 for the RSR tau_shape)."""
    shape = 0.5 + 0.5 * (n - 1) if q_rsr is None else 0.5 + 0.5 * q_rsr
    return dict(tau_rate=0.005, tau_shape=shape, a_mu=np.zeros(r), a_prec=np.eye(r) / 10,
                b_mu=np.zeros(p), b_prec=np.eye(p) / 10)


def moran_basis(Q, X, r=0.5, q=None):
    """logit.py:415-452 -- RSR basis K (n x q) and reduced precision K'QK from the Moran operator."""
    from scipy.sparse import csc_matrix
    Q = csc_matrix(Q)
    n = X.shape[0]
    chol = np.linalg.cholesky(X.T @ X)
    zz = solve_triangular(chol, np.eye(X.shape[1]), lower=True)
    XTX_i = solve_triangular(chol, zz, lower=True, trans=1)
    P = -np.linalg.multi_dot([X, XTX_i, X.T])
    P[np.diag_indices_from(P)] += 1
    A = Q.copy()
    A.data = -A.data
    A.setdiag(0)
    omega = n * (P.T @ A @ P) / A.sum()
    w, v = np.linalg.eigh(omega)
    if q:
        qq = int(q)
    else:
        if not 0 <= r <= 1:
            raise ValueError('Threshold value needs to be in [0, 1]')
        qq = int(w[w >= r].size)
        if not qq:
            raise ValueError('The Moran Operator Matrix of the data has no positive eigenvalues. '
                             'Set threshold to a lower value')
    K = np.ascontiguousarray(v[:, -qq:])
    return K, K.T @ Q @ K


# --------------------------------------------------------------------------------------
# the native routines (distributions.pyx)
# --------------------------------------------------------------------------------------
def precision_mvnorm(b, prec, z):
    """distributions.pyx:88-105 with the normal draw ``z`` injected.

    dpotrf('U') on prec.T factorises prec = L L' (L lower in C order); out = L z + b; dpotrs solves
    prec x = out  =>  x = prec^-1 b + L^-T z.  Returns (x, L).
    """
    L = np.linalg.cholesky(np.asarray(prec, dtype=np.float64))
    out = L @ z + b
    x = cho_solve((L, True), out)
    return x, L


def ensure_sums_to_zero(x, z):
    """distributions.pyx:24-39 -- x - (sum x / sum z) z."""
    return x + (-np.sum(x) / np.sum(z)) * z


# --------------------------------------------------------------------------------------
# logit sweep conditionals (logit.py)
# --------------------------------------------------------------------------------------
def occupancy_linpred(X, beta, spatial):
    """logit.py:201 -- psi = X beta + spatial."""
    return X @ beta + spatial


def tau_rate(eta, Q, tau_rate0):
    """logit.py:206-208 / probit.py:211-213 -- rate of the Gamma conditional of tau."""
    return 0.5 * (eta @ (Q @ eta)) + tau_rate0


def eta_rsr_system(K, X, omega, k, beta, tau, Qrsr):
    """logit.py:478-482 + 327-335 -- (Lambda, b) of eta | . ~ N(Lambda^-1 b, Lambda^-1)."""
    b = K.T @ (k - omega * (X @ beta))
    f1 = K.T * np.sqrt(omega)
    Lam = f1 @ f1.T + tau * Qrsr
    return Lam, b


def eta_rsr_draw_reference(K, X, omega, k, beta, tau, Qrsr, eps, eigen=None):
    """logit.py:325-337 verbatim algebra: n + q normals, perturb-then-solve (LU)."""
    q = Qrsr.shape[0]
    if eigen is None:
        s, u = np.linalg.eigh(Qrsr)
        eigen = u * np.sqrt(s)
    b = K.T @ (k - omega * (X @ beta))
    f1 = K.T * np.sqrt(omega)
    out = b + f1 @ eps[q:] + (np.sqrt(tau) * eigen) @ eps[:q]
    return np.linalg.solve(f1 @ f1.T + tau * Qrsr, out)


def gaussian_from_system(Lam, b, eps):
    """Engine-form draw of N(Lam^-1 b, Lam^-1): mean + L^-T eps with Lam = L L'. Returns (x, mean, L)."""
    L = np.linalg.cholesky(Lam)
    mean = cho_solve((L, True), b)
    x = mean + solve_triangular(L, eps, lower=True, trans=1)
    return x, mean, L


def beta_system_logit(X, omega, k, spatial, b_prec, b_prec_by_mu):
    """logit.py:226-231 -- (A, b) of beta | . ."""
    XT = X.T
    A = (XT * omega) @ X + b_prec
    b = XT @ (k - omega * spatial) + b_prec_by_mu
    return A, b


def alpha_system_logit(W_e, y_e, omega_a, a_prec, a_prec_by_mu):
    """logit.py:219-223 -- (A, b) of alpha | . over the visit rows of currently occupied surveyed sites."""
    WT = W_e.T
    A = (WT * omega_a) @ WT.T + a_prec
    b = WT @ (y_e - 0.5) + a_prec_by_mu
    return A, b


def z_prob_logit(X, beta, spatial, alpha, W_not_obs, stacked_w_indices, not_obs, not_surveyed):
    """logit.py:234-250 -- occupancy probabilities for not_obs sites and for not_surveyed sites."""
    num1 = expit(X[not_obs] @ beta + spatial[not_obs])
    num2 = expit(W_not_obs @ -alpha)
    stack_prod = np.multiply.reduceat(num2, stacked_w_indices)
    num = num1 * stack_prod
    p_no = num / ((1 - num1) + num)
    p_ns = expit(X[not_surveyed] @ beta + spatial[not_surveyed]) if len(not_surveyed) else np.zeros(0)
    return p_no, p_ns


def eta_icar_system(Q, X, omega, k, beta, tau):
    """logit.py:211-213 + 73-99 -- Lambda = tau Q + diag(omega) (dense here), b = k - omega * X beta."""
    Qd = Q.toarray() if hasattr(Q, 'toarray') else np.asarray(Q)
    Lam = tau * Qd + np.diag(omega)
    b = k - omega * (X @ beta)
    return Lam, b


def eta_icar_draw(Lam, b, eps):
    """Exact version of logit.py:73-99: x = Lam^-1 y with y ~ N(b, Lam), z = Lam^-1 1, then
    ensure_sums_to_zero (distributions.pyx:24-39).  The reference solves with MINRES (rtol 1e-5) and
    builds y from 2n-1 normals through the eigen-factor of Q; here y = b + L eps (same law, n normals)
    and the solves are exact.  Returns (eta, x, zvec, L)."""
    L = np.linalg.cholesky(Lam)
    x = cho_solve((L, True), b) + solve_triangular(L, eps, lower=True, trans=1)
    zvec = cho_solve((L, True), np.ones_like(b))
    return ensure_sums_to_zero(x, zvec), x, zvec, L


# --------------------------------------------------------------------------------------
# probit sweep conditionals (probit.py) -- with the beta-matrix copy fix
# --------------------------------------------------------------------------------------
def probit_omega_b(X, beta, spatial, eps, z, u):
    """probit.py:196-209."""
    loc = X @ beta + spatial + eps
    return truncnorm_draw(loc, z == 1, u)


def probit_eps(omega_b, X, beta, spatial, std):
    """probit.py:216-221."""
    return 0.5 * (omega_b - X @ beta - spatial) + std / np.sqrt(2)


def probit_eta_system(K, KTK, Qrsr, X, omega_b, beta, eps, tau):
    """probit.py:223-227."""
    return KTK + tau * Qrsr, K.T @ (omega_b - X @ beta - eps)


def probit_beta_system(X, XTX_plus_bprec, omega_b, spatial, eps, b_prec_by_mu):
    """probit.py:237-243 (precision matrix is COPIED: the reference mutates its cached copy -- bug)."""
    return XTX_plus_bprec.copy(), b_prec_by_mu + X.T @ (omega_b - spatial - eps)


def probit_omega_a(W_e, y_e, alpha, u):
    """probit.py:173-194."""
    return truncnorm_draw(W_e @ alpha, y_e == 1, u)


def probit_alpha_system(W_e, omega_a, a_prec, a_prec_by_mu):
    """probit.py:231-234."""
    return W_e.T @ W_e + a_prec, a_prec_by_mu + W_e.T @ omega_a


def z_prob_probit(X, beta, spatial, eps, alpha, W_not_obs, stacked_w_indices, not_obs, not_surveyed):
    """probit.py:245-259."""
    num1 = ndtr(X[not_obs] @ beta + spatial[not_obs] + eps[not_obs])
    num2 = 1 - ndtr(W_not_obs @ alpha)
    stack_prod = np.multiply.reduceat(num2, stacked_w_indices)
    num = num1 * stack_prod
    p_no = num / ((1 - num1) + num)
    p_ns = (ndtr(X[not_surveyed] @ beta + spatial[not_surveyed] + eps[not_surveyed])
            if len(not_surveyed) else np.zeros(0))
    return p_no, p_ns


# --------------------------------------------------------------------------------------
# CPU samplers: the sweep in the reference's order (CPU baseline + posterior parity checks)
# --------------------------------------------------------------------------------------
class OracleSampler:
    """One CPU chain of LogitRSR / LogitICAR / ProbitRSR following <model>.step() of the reference.

    model: 'logit_rsr' (logit.py:340-485), 'logit_icar' (logit.py:102-266), 'probit_rsr' (probit.py:27-270).
    ICAR uses the exact dense solve instead of MINRES (see ``eta_icar_draw``).  Randomness comes from a
    NumPy Generator (normals/uniforms/gamma, as in the reference) and oracle/pg_devroye.c (PG draws).
    """

    def __init__(self, model, Q, W, X, y, hparams=None, seed=0, K=None, Qrsr=None, r=0.5, q=None):
        from scipy.sparse import csc_matrix
        self.model = model
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self.n, self.p = self.X.shape
        self.rng = np.random.default_rng(np.random.SFC64(seed))
        part = partition_sites(y, self.n)
        self.part = part
        self.W_rows, self.y_rows, self.rowptr, self.site_ids = pack_visits(W, y, self.n)
        self.r = self.W_rows.shape[1]
        self.Q = csc_matrix(Q)
        if model in ('logit_rsr', 'probit_rsr'):
            if K is None:
                K, Qrsr = moran_basis(self.Q, self.X, r=r, q=q)
            elif Qrsr is None:
                Qrsr = K.T @ self.Q @ K
            self.K, self.Qrsr = np.ascontiguousarray(K), np.ascontiguousarray(Qrsr)
            self.q = self.K.shape[1]
        else:
            self.K, self.Qrsr, self.q = None, None, None
        hp = dict(hparams) if hparams else default_hparams(self.n, self.r, self.p, self.q)
        self.hp = hp
        self.a_prec_by_mu = hp['a_prec'] @ hp['a_mu']
        self.b_prec_by_mu = hp['b_prec'] @ hp['b_mu']
        # row bookkeeping for the not_obs sites (base.py:141-152)
        pos = {s: i for i, s in enumerate(part['surveyed'])}
        self._site_rows = {s: np.arange(self.rowptr[pos[s]], self.rowptr[pos[s] + 1]) for s in part['surveyed']}
        no = part['not_obs']
        self.rows_no = np.concatenate([self._site_rows[s] for s in no]) if no else np.zeros(0, dtype=np.int64)
        visits_no = np.array([self._site_rows[s].size for s in no], dtype=np.int64)
        self.stacked = np.pad(np.cumsum(visits_no), (1, 0))[:-1]
        self.rows_obs = (np.concatenate([self._site_rows[s] for s in part['obs']])
                         if part['obs'] else np.zeros(0, dtype=np.int64))
        self.row_site = np.repeat(self.site_ids, np.diff(self.rowptr))
        if model == 'probit_rsr':
            self.KTK = self.K.T @ self.K
            self.XTX_plus_bprec = self.X.T @ self.X + hp['b_prec']
        self.z = part['z'].copy()

    def init_state(self, start=None):
        """base.py:188-212; logit.py:462-476; probit.py:155-171."""
        rng, hp = self.rng, self.hp
        if start is None:
            self.tau = rng.gamma(0.5, 1 / hp['tau_rate'])
            eta = rng.standard_normal(self.n)
            eta = eta - eta.mean()
            self.alpha = rng.multivariate_normal(hp['a_mu'], 100 * hp['a_prec'], method='cholesky')
            self.beta = rng.multivariate_normal(hp['b_mu'], 100 * hp['b_prec'], method='cholesky')
            if self.K is not None:
                eta = rng.normal(scale=5, size=self.q)
            self.eta = eta
            if self.model == 'probit_rsr':
                self.eps = rng.standard_normal(self.n)
        else:
            self.alpha = np.asarray(start['alpha'], dtype=np.float64)
            self.beta = np.asarray(start['beta'], dtype=np.float64)
            self.tau = float(start['tau'])
            self.eta = np.asarray(start['eta'], dtype=np.float64)
            if self.model == 'probit_rsr':
                self.eps = np.asarray(start['eps'], dtype=np.float64)
        self.spatial = self.K @ self.eta if self.K is not None else self.eta
        self.z = self.part['z'].copy()

    def _exist_rows(self):
        no = self.part['not_obs']
        if no:
            occ = self.z[no] == 1
            keep = np.repeat(occ, np.diff(np.append(self.stacked, self.rows_no.size)))
            return np.concatenate([self.rows_obs, self.rows_no[keep]])
        return self.rows_obs

    def _seed(self):
        return int(self.rng.integers(0, 2**63 - 1))

    def step(self):
        m, rng, hp, X = self.model, self.rng, self.hp, self.X
        k = self.z - 0.5
        no, ns = self.part['not_obs'], self.part['not_surveyed']
        if m == 'probit_rsr':
            omega_b = probit_omega_b(X, self.beta, self.spatial, self.eps, self.z, rng.random(self.n))
            self.tau = rng.gamma(hp['tau_shape'], 1 / tau_rate(self.eta, self.Qrsr, hp['tau_rate']))
            self.eps = probit_eps(omega_b, X, self.beta, self.spatial, rng.standard_normal(self.n))
            A, b = probit_eta_system(self.K, self.KTK, self.Qrsr, X, omega_b, self.beta, self.eps, self.tau)
            self.eta = precision_mvnorm(b, A, rng.standard_normal(self.q))[0]
            self.spatial = self.K @ self.eta
            A, b = probit_beta_system(X, self.XTX_plus_bprec, omega_b, self.spatial, self.eps, self.b_prec_by_mu)
            self.beta = precision_mvnorm(b, A, rng.standard_normal(self.p))[0]
            rows = self._exist_rows()
            W_e, y_e = self.W_rows[rows], self.y_rows[rows]
            omega_a = probit_omega_a(W_e, y_e, self.alpha, rng.random(rows.size))
            A, b = probit_alpha_system(W_e, omega_a, hp['a_prec'], self.a_prec_by_mu)
            self.alpha = precision_mvnorm(b, A, rng.standard_normal(self.r))[0]
            p_no, p_ns = z_prob_probit(X, self.beta, self.spatial, self.eps, self.alpha, self.W_rows[self.rows_no],
                                       self.stacked, no, ns)
        else:
            omega_b = pg1(occupancy_linpred(X, self.beta, self.spatial), self._seed())
            Qt = self.Qrsr if m == 'logit_rsr' else self.Q
            self.tau = rng.gamma(hp['tau_shape'], 1 / tau_rate(self.eta, Qt, hp['tau_rate']))
            if m == 'logit_rsr':
                Lam, b = eta_rsr_system(self.K, X, omega_b, k, self.beta, self.tau, self.Qrsr)
                self.eta = gaussian_from_system(Lam, b, rng.standard_normal(self.q))[0]
                self.spatial = self.K @ self.eta
            else:
                Lam, b = eta_icar_system(self.Q, X, omega_b, k, self.beta, self.tau)
                self.eta = eta_icar_draw(Lam, b, rng.standard_normal(self.n))[0]
                self.spatial = self.eta
            A, b = beta_system_logit(X, omega_b, k, self.spatial, hp['b_prec'], self.b_prec_by_mu)
            self.beta = precision_mvnorm(b, A, rng.standard_normal(self.p))[0]
            rows = self._exist_rows()
            W_e, y_e = self.W_rows[rows], self.y_rows[rows]
            omega_a = pg1(W_e @ self.alpha, self._seed())
            A, b = alpha_system_logit(W_e, y_e, omega_a, hp['a_prec'], self.a_prec_by_mu)
            self.alpha = precision_mvnorm(b, A, rng.standard_normal(self.r))[0]
            p_no, p_ns = z_prob_logit(X, self.beta, self.spatial, self.alpha, self.W_rows[self.rows_no],
                                      self.stacked, no, ns)
        if no:
            self.z[no] = rng.uniform(size=len(no)) < p_no
        if ns:
            self.z[ns] = rng.uniform(size=len(ns)) < p_ns

    def run(self, size, burnin=0, start=None):
        """base.py:214-241 -- returns dict of arrays alpha (m, r), beta (m, p), tau (m,), m = size - burnin."""
        self.init_state(start)
        m = size - burnin
        out = dict(alpha=np.zeros((m, self.r)), beta=np.zeros((m, self.p)), tau=np.zeros(m))
        for i in range(size):
            self.step()
            if i >= burnin:
                j = i - burnin
                out['alpha'][j], out['beta'][j], out['tau'][j] = self.alpha, self.beta, self.tau
        return out


def _run_chain(args):
    model, data, kw, seed, size, burnin = args
    from threadpoolctl import threadpool_limits
    with threadpool_limits(limits=1):   # one BLAS thread per chain process
        s = OracleSampler(model, *data, seed=seed, **kw)
        return s.run(size, burnin)


def run_chains(model, Q, W, X, y, size, burnin=0, chains=2, seed=0, processes=None, **kw):
    """parallel.py:4-42 -- one process per chain; chain c is seeded (seed, c)."""
    jobs = [(model, (Q, W, X, y), kw, [seed, c], size, burnin) for c in range(chains)]
    if processes is None or processes <= 1 or chains == 1:
        res = [_run_chain(j) for j in jobs]
    else:
        import multiprocessing as mp
        with mp.get_context('fork').Pool(processes) as pool:
            res = pool.map(_run_chain, jobs)
    return {name: np.stack([r[name] for r in res]) for name in ('alpha', 'beta', 'tau')}


# --------------------------------------------------------------------------------------
# synthetic data (utils.py:38-103, 106-294), restated so that it scales; explicit lattice shape
# --------------------------------------------------------------------------------------
def lattice_precision(rows, cols, queen=True, rho=1.0):
    """utils.py:38-103 via the lat2SW restatement: Q = D - rho A, CSC float."""
    from scipy import sparse
    n = rows * cols
    j = np.arange(n) % cols
    i = np.arange(n) // cols
    src, dst = [], []

    def link(mask, off):
        s = np.nonzero(mask)[0]
        src.append(s)
        dst.append(s + off)

    link(j < cols - 1, 1)
    link(i < rows - 1, cols)
    if queen:
        link((i < rows - 1) & (j < cols - 1), cols + 1)
        link((i < rows - 1) & (j > 0), cols - 1)
    s, d = np.concatenate(src), np.concatenate(dst)
    A = sparse.coo_matrix((np.ones(s.size), (s, d)), shape=(n, n)).tocsr()
    A = A + A.T
    D = sparse.diags(np.asarray(A.sum(axis=1)).ravel())
    return (D - rho * A).tocsc()


def make_data_lattice(rows, cols, min_v=2, max_v=8, ns=None, p=3, r=3, tau_range=(0.25, 1.5), seed=0):
    """make_data-style synthetic inputs (utils.py:261-292) on an explicit rows x cols queen lattice.

    Differences from the reference, both forced by scale: the lattice shape is given,
    not drawn; eta is drawn by a sparse solve of (Q + 1e-6 I) x = xi-weighted incidence instead of the
    dense pinvh + eigh MVN (O(n^3)).  Everything else follows utils.py line by line, including the
    sign convention psi = expit(X beta - eta) (utils.py:283-284).
    """
    from scipy import sparse
    from scipy.sparse.linalg import splu
    rng = np.random.default_rng(np.random.SFC64(seed))
    n = rows * cols
    ns = n // 2 if ns is None else ns
    surveyed = rng.choice(n, size=ns, replace=False)
    visits = rng.integers(min_v, max_v, size=ns, endpoint=True)
    alpha = rng.standard_normal(r)
    beta = rng.standard_normal(p)
    tau = rng.uniform(*tau_range)
    Q = lattice_precision(rows, cols)
    # eta ~ N(0, (tau Q)^+): Q = B'B with B the edge incidence matrix; x = (Q + eps I)^-1 B' xi, centred
    A = sparse.triu(-Q, k=1).tocoo()
    m = A.nnz
    B = sparse.coo_matrix((np.concatenate([np.ones(m), -np.ones(m)]),
                           (np.concatenate([np.arange(m), np.arange(m)]), np.concatenate([A.row, A.col]))),
                          shape=(m, n)).tocsc()
    xi = rng.standard_normal(m)
    lu = splu((Q + 1e-6 * sparse.identity(n)).tocsc())
    eta = lu.solve(B.T @ xi) / np.sqrt(tau)
    eta -= eta.mean()
    X = rng.uniform(-2, 2, n * p).reshape(n, -1)
    X[:, 0] = 1
    psi = np.exp(-np.logaddexp(0, -X @ beta + eta))
    z = rng.binomial(1, p=psi, size=n)
    W, y = {}, {}
    for i, j in zip(surveyed, visits):
        _W = rng.uniform(-2, 2, size=j * r).reshape(j, -1)
        _W[:, 0] = 1
        d = np.exp(-np.logaddexp(0, -_W @ alpha))
        W[int(i)] = _W
        y[int(i)] = rng.binomial(1, z[i] * d)
    return Q, W, X, y, alpha, beta, tau, z
