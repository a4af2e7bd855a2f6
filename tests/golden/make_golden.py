"""tests/golden/make_golden.py -- regenerates the golden vectors from the REAL reference.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

Imports the unmodified reference through oracle/ref_loader.py (its own .pyx compiled in oracle/_ref,
polyagamma / arviz / libpysal replaced by the shims documented there) and dumps, for the reference's
own test fixture (gibbs/tests/test_samplers.py:14-16), the inputs and the outputs of every deterministic
piece of the sweep with the randomness pinned:

  logit_rsr.npz   _update_tau / _update_eta / _update_beta / _update_alpha / _update_z of LogitRSRGibbs
  logit_icar.npz  the same for LogitICARGibbs (eta through the reference's MINRES path, rtol 1e-5)
  probit_rsr.npz  every update of ProbitRSRGibbs (first call only: the reference corrupts its cached
                  beta precision afterwards)
  native.npz      precision_mvnorm / ensure_sums_to_zero (distributions.pyx) known answers
  posterior.npz   posterior summaries of reference chains (4 x 3000, burn-in 1000) for the MCSE check

How the pinning works without touching the reference: ``sampler.rng`` is swapped for either a seeded
SFC64 generator (whose ziggurat normals are reproduced by drawing from an identically seeded generator)
or a tiny fake whose ``uniform`` / ``random`` / ``gamma`` return chosen arrays.  Occupancy probabilities
inside ``_update_z`` are not exposed by the reference; they are recovered to 2^-52 by bisection on the
injected uniforms (``z_new = u < p``).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle.ref_loader import load_reference, run_reference_chains  # noqa: E402
from oracle import sweep as S  # noqa: E402

load_reference()
from occuspytial.utils import make_data, get_generator  # noqa: E402
from occuspytial import LogitRSRGibbs, LogitICARGibbs, ProbitRSRGibbs  # noqa: E402
from occuspytial.distributions import precision_mvnorm, ensure_sums_to_zero  # noqa: E402


class FakeRNG:
    """Returns preset arrays from uniform/random/gamma/standard_normal, in call order."""

    def __init__(self, uniforms=(), normals=(), gamma_identity=False):
        self.uniforms = list(uniforms)
        self.normals = list(normals)
        self.gamma_identity = gamma_identity

    def uniform(self, size=None):
        u = self.uniforms.pop(0)
        assert u.shape[0] == size
        return u

    def random(self, size=None):
        return self.uniform(size)

    def standard_normal(self, size=None):
        z = self.normals.pop(0)
        assert z.shape[0] == size
        return z

    def gamma(self, shape, scale):
        assert self.gamma_identity
        return scale  # tau := 1 / rate  => rate is recoverable


def csr_in_site_order(W, y, n):
    sites = sorted(W)
    counts = np.zeros(n, dtype=np.int64)
    for s in sites:
        counts[s] = W[s].shape[0]
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    W_rows = np.concatenate([W[s] for s in sites], axis=0)
    y_rows = np.concatenate([y[s] for s in sites]).astype(np.uint8)
    return W_rows, y_rows, ptr


def exists_rows(sampler, z, ptr):
    """Row indices (site-order CSR) of the visit rows in the reference's ``exists`` order (logit.py:186-188)."""
    occ = [i for i in sampler.fixed.not_obs if z[i]]
    exists = sampler.fixed.obs + occ
    rows = np.concatenate([np.arange(ptr[s], ptr[s + 1]) for s in exists])
    return exists, rows


def bisect_z_probs(sampler, sites_no, sites_ns):
    """Recover the Bernoulli probabilities used inside _update_z by bisection on injected uniforms."""
    n_no, n_ns = len(sites_no), len(sites_ns)
    lo_no, hi_no = np.zeros(n_no), np.ones(n_no)
    lo_ns, hi_ns = np.zeros(n_ns), np.ones(n_ns)
    z_keep = sampler.state.z.copy()
    for _ in range(60):
        u_no, u_ns = 0.5 * (lo_no + hi_no), 0.5 * (lo_ns + hi_ns)
        sampler.rng = FakeRNG(uniforms=[u_no, u_ns])
        sampler.state.z = z_keep.copy()
        sampler._update_z()
        below_no = sampler.state.z[sites_no] == 1   # u < p
        below_ns = sampler.state.z[sites_ns] == 1
        lo_no = np.where(below_no, u_no, lo_no)
        hi_no = np.where(below_no, hi_no, u_no)
        lo_ns = np.where(below_ns, u_ns, lo_ns)
        hi_ns = np.where(below_ns, hi_ns, u_ns)
    sampler.state.z = z_keep
    return 0.5 * (lo_no + hi_no), 0.5 * (lo_ns + hi_ns)


def fixture():
    return make_data(min_v=2, max_v=10, ns=100, p=3, q=2, tau_range=(0.25, 1.5), random_state=10)


def common_inputs(Q, W, X, y, sampler):
    n = X.shape[0]
    W_rows, y_rows, ptr = csr_in_site_order(W, y, n)
    Qd = Q.toarray()
    return dict(Q=Qd, X=X, W_rows=W_rows, y_rows=y_rows, visit_ptr=ptr,
                surveyed_order=np.array(list(W), dtype=np.int64),
                not_obs=np.array(sampler.fixed.not_obs, dtype=np.int64),
                obs=np.array(sampler.fixed.obs, dtype=np.int64),
                not_surveyed=np.array(sampler.fixed.not_surveyed, dtype=np.int64),
                stacked_w_indices=np.asarray(sampler.fixed.stacked_w_indices, dtype=np.int64),
                z_init=sampler.state.z.copy(),
                tau_rate=sampler.fixed.tau_rate, tau_shape=sampler.fixed.tau_shape,
                a_mu=sampler.fixed.a_mu, a_prec=sampler.fixed.a_prec,
                b_mu=sampler.fixed.b_mu, b_prec=sampler.fixed.b_prec)


def random_state_for(sampler, gen, n, q_eta):
    z = sampler.state.z.copy()
    no, ns = sampler.fixed.not_obs, sampler.fixed.not_surveyed
    z[no] = gen.random(len(no)) < 0.5
    z[ns] = gen.random(len(ns)) < 0.5
    return dict(z=z, alpha=gen.normal(size=sampler.fixed.a_mu.size), beta=gen.normal(size=sampler.X.shape[1]),
                eta=gen.normal(size=q_eta) * 0.7, tau=float(gen.uniform(0.3, 2.0)))


def logit_golden(cls, name, **kw):
    Q, W, X, y, *_ = fixture()
    s = cls(Q, W, X, y, random_state=10, **kw)
    n = X.shape[0]
    rsr = hasattr(s.fixed, 'K')
    out = common_inputs(Q, W, X, y, s)
    if rsr:
        out['K'] = s.fixed.K
        out['Qrsr'] = s.fixed.Q
    gen = np.random.default_rng(2024)
    st = random_state_for(s, gen, n, s.fixed.q if rsr else n)
    if not rsr:
        st['eta'] -= st['eta'].mean()
    spatial = s.fixed.K @ st['eta'] if rsr else st['eta']
    omega_b = S.pg1(X @ st['beta'] + spatial, 77)
    s.state.z = st['z'].copy()
    s.state.k = s.state.z - 0.5
    s.state.alpha, s.state.beta, s.state.tau, s.state.eta = st['alpha'], st['beta'], st['tau'], st['eta']
    s.state.spatial = spatial
    s.state.omega_b = omega_b
    out.update({'st_' + k: v for k, v in st.items()})
    out['omega_b'] = omega_b

    # tau: rate through the identity-gamma fake (logit.py:206-209)
    s.rng = FakeRNG(gamma_identity=True)
    s._update_tau()
    out['tau_rate_out'] = 1.0 / s.state.tau
    s.state.tau = st['tau']

    # eta (logit.py:211-217 / 478-485)
    if rsr:
        qn = s.fixed.q + n
        s.rng = FakeRNG(normals=[np.zeros(qn)])
        s._update_eta()
        out['eta_mean'] = s.state.eta.copy()
        eps = get_generator(5).standard_normal(qn)
        s.state.eta = st['eta']
        s.rng = get_generator(5)
        s._update_eta()
        out['eta_eps'] = eps
        out['eta_draw'] = s.state.eta.copy()
        out['eta_eigen'] = s.dists.eta_posterior._eigen
    else:
        qn = s.dists.eta_posterior._n_plus_k
        s.rng = FakeRNG(normals=[np.zeros(qn)])
        s._update_eta()
        out['eta_mean'] = s.state.eta.copy()     # MINRES, rtol 1e-5, constrained to sum zero
    # continue from a fixed new eta so that later pieces do not depend on solver accuracy
    eta_new = gen.normal(size=st['eta'].size) * 0.5
    if not rsr:
        eta_new -= eta_new.mean()
    s.state.eta = eta_new
    s.state.spatial = s.fixed.K @ eta_new if rsr else eta_new
    out['eta_new'] = eta_new

    # beta (logit.py:226-232)
    s.rng = get_generator(6)
    s._update_beta()
    out['beta_normals'] = get_generator(6).standard_normal(X.shape[1])
    out['beta_out'] = s.state.beta.copy()

    # alpha (logit.py:219-224) with injected omega_a
    exists, rows = exists_rows(s, st['z'], out['visit_ptr'])
    s.state.exists = exists
    s.state.W = s.W[exists]
    lin = s.state.W @ st['alpha']
    omega_a_e = S.pg1(lin, 78)
    s.state.omega_a = omega_a_e
    omega_a_rows = np.zeros(out['W_rows'].shape[0])
    omega_a_rows[rows] = omega_a_e
    out['omega_a_rows'] = omega_a_rows
    s.state.alpha = st['alpha']
    s.rng = get_generator(7)
    s._update_alpha()
    out['alpha_normals'] = get_generator(7).standard_normal(st['alpha'].size)
    out['alpha_out'] = s.state.alpha.copy()

    # z probabilities (logit.py:234-252) for the new (alpha, beta, eta)
    p_no, p_ns = bisect_z_probs(s, s.fixed.not_obs, s.fixed.not_surveyed)
    out['p_no'], out['p_ns'] = p_no, p_ns
    np.savez_compressed(os.path.join(HERE, name), **out)
    print('wrote', name, {k: np.shape(v) for k, v in out.items() if k.startswith(('eta', 'beta', 'alpha', 'p_'))})


def probit_golden():
    Q, W, X, y, *_ = fixture()
    s = ProbitRSRGibbs(Q, W, X, y, random_state=10, q=10)
    n = X.shape[0]
    out = common_inputs(Q, W, X, y, s)
    out['K'], out['Qrsr'], out['KTK'] = s.fixed.K, s.fixed.Q, s.fixed.KTK
    out['XTX_plus_bprec'] = s.fixed.XTX_plus_bprec.copy()
    gen = np.random.default_rng(4048)
    st = random_state_for(s, gen, n, s.fixed.q)
    st['eps'] = gen.normal(size=n)
    spatial = s.fixed.K @ st['eta']
    s.state.z = st['z'].copy()
    s.state.alpha, s.state.beta, s.state.tau, s.state.eta, s.state.eps = (st['alpha'], st['beta'], st['tau'],
                                                                         st['eta'], st['eps'])
    s.state.spatial = spatial
    out.update({'st_' + k: v for k, v in st.items()})

    # omega_b (probit.py:196-209): uniforms for z==1 sites first, then z==0 sites
    mask = st['z'] == 1
    ua, ub = gen.random(int(mask.sum())), gen.random(int((~mask).sum()))
    u_b = np.zeros(n)
    u_b[mask], u_b[~mask] = ua, ub
    s.rng = FakeRNG(uniforms=[ua, ub])
    s._update_omega_b()
    out['u_omega_b'] = u_b
    out['omega_b'] = s.state.omega_b.copy()

    s.rng = FakeRNG(gamma_identity=True)
    s._update_tau()
    out['tau_rate_out'] = 1.0 / s.state.tau
    s.state.tau = st['tau']

    # eps (probit.py:216-221)
    s.rng = get_generator(8)
    s._update_eps()
    out['eps_normals'] = get_generator(8).standard_normal(n)
    out['eps_out'] = s.state.eps.copy()

    # eta (probit.py:223-229)
    s.rng = get_generator(9)
    s._update_eta()
    out['eta_normals'] = get_generator(9).standard_normal(s.fixed.q)
    out['eta_out'] = s.state.eta.copy()

    # beta (probit.py:237-243) -- first and only valid call
    s.rng = get_generator(11)
    s._update_beta()
    out['beta_normals'] = get_generator(11).standard_normal(X.shape[1])
    out['beta_out'] = s.state.beta.copy()

    # omega_a (probit.py:173-194): uniforms for y==1 rows (exists order) then y==0 rows
    exists, rows = exists_rows(s, st['z'], out['visit_ptr'])
    y_e = s.y[exists]
    m = y_e == 1
    ua, ub = gen.random(int(m.sum())), gen.random(int((~m).sum()))
    s.rng = FakeRNG(uniforms=[ua, ub])
    s._update_omega_a()
    u_rows = np.zeros(out['W_rows'].shape[0])
    u_e = np.zeros(rows.size)
    u_e[m], u_e[~m] = ua, ub
    u_rows[rows] = u_e
    omega_a_rows = np.zeros(out['W_rows'].shape[0])
    omega_a_rows[rows] = s.state.omega_a
    out['u_omega_a_rows'], out['omega_a_rows'] = u_rows, omega_a_rows

    # alpha (probit.py:231-235)
    s.rng = get_generator(12)
    s._update_alpha()
    out['alpha_normals'] = get_generator(12).standard_normal(st['alpha'].size)
    out['alpha_out'] = s.state.alpha.copy()

    p_no, p_ns = bisect_z_probs(s, s.fixed.not_obs, s.fixed.not_surveyed)
    out['p_no'], out['p_ns'] = p_no, p_ns
    np.savez_compressed(os.path.join(HERE, 'probit_rsr.npz'), **out)
    print('wrote probit_rsr.npz')


def native_golden():
    gen = np.random.default_rng(99)
    out = {}
    for d in (3, 10, 64):
        A = gen.normal(size=(d + 5, d))
        P = A.T @ A + 0.1 * np.eye(d)
        b = gen.normal(size=d)
        Pw = P.copy()
        x = precision_mvnorm(b, Pw, random_state=get_generator(d))
        out[f'P{d}'], out[f'b{d}'], out[f'x{d}'] = P, b, x
        out[f'z{d}'] = get_generator(d).standard_normal(d)
        out[f'Pafter{d}'] = Pw
    x, z = gen.normal(size=50), gen.uniform(0.5, 1.5, size=50)
    o = np.zeros(50)
    ensure_sums_to_zero(x, z, o)
    out['stz_x'], out['stz_z'], out['stz_out'] = x, z, o
    np.savez_compressed(os.path.join(HERE, 'native.npz'), **out)
    print('wrote native.npz')


def posterior_golden():
    """Reference chains for the posterior-parity check (north_star: within 3 MCSE)."""
    Q, W, X, y, alpha, beta, tau, z = make_data(random_state=0)
    n = X.shape[0]
    W_rows, y_rows, ptr = csr_in_site_order(W, y, n)
    hp = {'tau_rate': 1.0, 'tau_shape': 5.0, 'a_mu': np.zeros(3), 'b_mu': np.zeros(3),
          'a_prec': np.eye(3) / 10, 'b_prec': np.eye(3) / 10}
    out = dict(Q=Q.toarray(), X=X, W_rows=W_rows, y_rows=y_rows, visit_ptr=ptr, true_alpha=alpha, true_beta=beta,
               true_tau=tau)
    # RSR with an informative tau prior (well-behaved posterior); ICAR with the default hyper-parameters
    # (its tau posterior is heavy-tailed -- SURVEY.md 7.7 -- so only alpha/beta are compared tightly)
    for label, cls, kw in (('rsr', LogitRSRGibbs, dict(q=12, hparams=dict(hp))), ('icar', LogitICARGibbs, dict())):
        s = cls(Q, W, X, y, random_state=10, **kw)
        if label == 'rsr':
            out['K'] = s.fixed.K
        chains = run_reference_chains(s, 3000, 1000, chains=4)
        for name in ('alpha', 'beta', 'tau'):
            out[f'{label}_{name}'] = np.stack([c[name] for c in chains])
        print(label, 'alpha', out[f'{label}_alpha'].mean((0, 1)), 'beta', out[f'{label}_beta'].mean((0, 1)),
              'tau', out[f'{label}_tau'].mean())
    for k, v in hp.items():
        out['hp_' + k] = v
    np.savez_compressed(os.path.join(HERE, 'posterior.npz'), **out)
    print('wrote posterior.npz')


if __name__ == '__main__':
    logit_golden(LogitRSRGibbs, 'logit_rsr.npz', q=10)
    logit_golden(LogitICARGibbs, 'logit_icar.npz')
    probit_golden()
    native_golden()
    posterior_golden()
