"""GPU parity tests of LogitICARGibbs ("PGGibbs"): banded per-chain Cholesky of tau Q + diag(omega) against the
reference's golden vectors (MINRES path, solver accuracy only) and the oracle's exact dense solve (1e-9)."""
import numpy as np
import pytest
from scipy import sparse

from conftest import golden

pytestmark = pytest.mark.gpu


def _sampler_from_golden(g, cls, **kw):
    n = g['X'].shape[0]
    ptr = g['visit_ptr']
    order = g['surveyed_order'] if 'surveyed_order' in g else [s for s in range(n) if ptr[s + 1] > ptr[s]]
    W = {int(s): g['W_rows'][ptr[s]:ptr[s + 1]] for s in order}
    y = {int(s): g['y_rows'][ptr[s]:ptr[s + 1]] for s in order}
    return cls(sparse.csc_matrix(g['Q']), W, g['X'], y, **kw)


def test_icar_pieces_vs_reference_golden():
    from occuspytial_b200 import LogitICARGibbs, PGGibbs, _lib
    assert PGGibbs is LogitICARGibbs
    g = golden('logit_icar.npz')
    s = _sampler_from_golden(g, LogitICARGibbs, random_state=1)
    n = g['X'].shape[0]
    eng = s._make_engine(2, seed=5, debug=True)
    to_e, from_e = s._sites_to_engine, s._sites_from_engine
    eng.set_start(dict(alpha=g['st_alpha'], beta=g['st_beta'], tau=g['st_tau'], eta=to_e(g['st_eta'])))
    eng.set_field(_lib.F_Z, to_e(g['st_z']).astype(np.uint8))
    eng.set_field(_lib.F_OMEGA, to_e(g['omega_b']))
    # alpha's omega_a rows follow the engine's CSR (permuted site order): map the golden rows through it
    prob = s._problem_arrays()
    ptr_g, ptr_e, perm = g['visit_ptr'], prob['visit_ptr'], s._perm
    oma = np.zeros(prob['W_rows'].shape[0])
    for j in range(n):
        old = perm[j]
        oma[ptr_e[j]:ptr_e[j + 1]] = g['omega_a_rows'][ptr_g[old]:ptr_g[old + 1]]
    p, r = 3, 2
    normals = np.concatenate([np.zeros(n), g['beta_normals'], g['alpha_normals']])
    out = eng.debug_gaussian_half(normals=normals, tau=g['st_tau'], omega_a=oma)
    assert out['fails'] == 0
    np.testing.assert_allclose(out['tau_rate'], g['tau_rate_out'], rtol=1e-11)
    eta = from_e(eng.get_field(_lib.F_ETA)[0])
    assert abs(eta.sum()) < 1e-9
    # reference: MINRES, rtol 1e-5 (logit.py:87) -> solver accuracy only
    assert np.linalg.norm(eta - g['eta_mean']) / np.linalg.norm(eta) < 5e-3
    np.testing.assert_allclose(eng.get_field(_lib.F_ALPHA)[1], g['alpha_out'], rtol=1e-9, atol=1e-11)
    # z probabilities at the golden's (alpha_out, beta_out, eta_new)
    eng.set_start(dict(alpha=g['alpha_out'], beta=g['beta_out'], tau=g['st_tau'], eta=to_e(g['eta_new'])))
    eng.set_field(_lib.F_Z, to_e(g['st_z']).astype(np.uint8))
    eng.debug_site_half()
    zp = from_e(eng.get_field(_lib.F_ZPROB)[0])
    np.testing.assert_allclose(zp[g['not_obs']], g['p_no'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(zp[g['not_surveyed']], g['p_ns'], rtol=1e-10, atol=1e-14)
    eng.close()


# half bandwidths 16 / 31 / 41 / 51 / 59 / 63: kernel instances NT = 3, 5, 6, 7 (BASELINE config 4: 50 x 50), 8 and the
# wide-band kernel (bw > 60)
@pytest.mark.parametrize('lattice,chains', [((10, 15), 3), ((20, 30), 130), ((40, 41), 3), ((50, 50), 64), ((58, 59), 3),
                                            ((62, 62), 5)])
def test_icar_iteration_vs_oracle(lattice, chains):
    """eta draw (banded Cholesky, sum-to-zero projection), beta and alpha systems vs the oracle's dense solve,
    in the engine's site order (the Cholesky factor, hence L^-T eps, depends on the ordering)."""
    from occuspytial_b200 import LogitICARGibbs, _lib
    from occuspytial_b200.utils import make_data
    from oracle import sweep as S
    n = lattice[0] * lattice[1]
    Q, W, X, y, *_ = make_data(n=n, min_v=1, max_v=5, lattice=lattice, random_state=6)
    s = LogitICARGibbs(Q, W, X, y, random_state=2)
    f = s.fixed
    prob = s._problem_arrays()
    perm, bw = s._band()
    assert bw <= max(lattice) + 1
    Qp = sparse.csr_matrix(Q)[perm][:, perm]
    Xp = prob['X']
    V, p, r = prob['W_rows'].shape[0], 3, 3
    rng = np.random.default_rng(3)
    eng = s._make_engine(chains, seed=9, debug=True)
    eta0 = rng.normal(size=(chains, n))
    eta0 -= eta0.mean(axis=1, keepdims=True)
    st = dict(alpha=rng.normal(size=(chains, r)), beta=rng.normal(size=(chains, p)), tau=rng.uniform(0.3, 3, chains), eta=eta0)
    eng.set_start(st)      # eta given directly in engine order
    eng.debug_site_half()
    psi = eng.get_field(_lib.F_PSI)
    z = eng.get_field(_lib.F_Z).astype(float)
    omega = eng.get_field(_lib.F_OMEGA)
    normals = rng.normal(size=(chains, n + p + r))
    omega_a = rng.uniform(0.02, 0.25, size=(chains, V))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'], omega_a=omega_a)
    assert out['fails'] == 0
    eta_new, beta_new, alpha_new = eng.get_field(_lib.F_ETA), eng.get_field(_lib.F_BETA), eng.get_field(_lib.F_ALPHA)
    vptr, Wd, yd = prob['visit_ptr'], prob['W_rows'], prob['y_rows'].astype(float)
    cls = prob['site_class']
    check = range(chains) if chains <= 8 else sorted({0, 1, 63, min(64, chains - 1), min(127, chains - 1), chains - 1})
    if n > 1500:
        check = list(check)[:3] + [chains - 1] if chains > 4 else check   # dense n x n oracle solves: keep it to seconds
    for c in check:
        np.testing.assert_allclose(psi[c], Xp @ st['beta'][c] + eta0[c], rtol=1e-12, atol=1e-13)
        k = z[c] - 0.5
        np.testing.assert_allclose(out['tau_rate'][c], S.tau_rate(eta0[c], Qp, f.tau_rate), rtol=1e-10)
        Lam, b = S.eta_icar_system(Qp, Xp, omega[c], k, st['beta'][c], st['tau'][c])
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-11, atol=1e-13)
        eta_o, x_o, z_o, L_o = S.eta_icar_draw(Lam, b, normals[c, :n])
        # column-band factor: Lc[j][e] = L[j + e][j]
        Lc = out['eta_chol'][c].reshape(n, bw + 1)
        for e in range(bw + 1):
            np.testing.assert_allclose(Lc[:n - e, e], np.diagonal(L_o, -e), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(out['eta_mean'][c], np.linalg.solve(Lam, b), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(eta_new[c], eta_o, rtol=1e-8, atol=1e-9)
        assert abs(eta_new[c].sum()) < 1e-9
        A, bb = S.beta_system_logit(Xp, omega[c], k, eta_new[c], f.b_prec, f.b_prec_by_mu)
        np.testing.assert_allclose(out['beta_prec'][c], A, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out['beta_b'][c], bb, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(beta_new[c], S.precision_mvnorm(bb, A, normals[c, n:n + p])[0], rtol=1e-9, atol=1e-10)
        rows = np.concatenate([np.arange(vptr[i], vptr[i + 1]) for i in range(n) if cls[i] != 0 and z[c][i] == 1])
        A, ba = S.alpha_system_logit(Wd[rows], yd[rows], omega_a[c][rows], f.a_prec, f.a_prec_by_mu)
        np.testing.assert_allclose(out['alpha_prec'][c], A, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(alpha_new[c], S.precision_mvnorm(ba, A, normals[c, n + p:])[0], rtol=1e-9, atol=1e-10)
    eng.close()


def test_icar_sampler_api_and_posterior(capfd):
    """API contract (gibbs/tests/test_samplers.py:38-61) and posterior vs the reference's chains: alpha (well
    identified) within 3 MCSE; beta / tau are heavy-tailed under the default prior, compared on
    robust summaries."""
    from occuspytial_b200 import LogitICARGibbs
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.diagnostics import summaries_from_draws, ess
    Q, W, X, y, *_ = make_data(min_v=2, max_v=10, ns=100, p=3, q=2, random_state=10)
    s = LogitICARGibbs(Q, W, X, y)
    s.sample(10)
    assert ' 10/10 [00:00<00:00,' in capfd.readouterr().err
    s.sample(10, progressbar=False)
    assert ' 10/10 [00:00<00:00,' not in capfd.readouterr().err
    out = LogitICARGibbs(Q, W, X, y, random_state=10).sample(5, chains=3, progressbar=False)
    assert out['alpha'].shape == (3, 5, 2) and out['beta'].shape == (3, 5, 3) and out['tau'].shape == (3, 5)
    a = LogitICARGibbs(Q, W, X, y, random_state=10).sample(5, start={'alpha': np.zeros(2), 'beta': np.zeros(3), 'tau': 2,
                                                                    'eta': np.zeros(150)}, progressbar=False)
    assert not np.allclose(a['alpha'][0, 0], out['alpha'][0, 0])

    g = golden('posterior.npz')
    s = _sampler_from_golden(g, LogitICARGibbs, random_state=4)
    post = s.sample(2000, burnin=1000, chains=256, progressbar=False)
    assert np.all(np.isfinite(post['alpha'])) and np.all(post['tau'] > 0)

    def mean_mcse(x):
        x = x.reshape(x.shape[0], x.shape[1], -1)
        return [(x[:, :, k].mean(), x[:, :, k].std(ddof=1) / np.sqrt(ess(summaries_from_draws(x[:, :, k], 200), x.shape[1])))
                for k in range(x.shape[2])]

    for (mr, sr), (mm, sm) in zip(mean_mcse(g['icar_alpha']), mean_mcse(post['alpha'])):
        assert abs(mr - mm) < 3.0 * np.hypot(sr, sm), (mr, mm, sr, sm)
    # beta: medians agree within a tenth of the posterior sd of the reference
    for k in range(3):
        ref, mine = g['icar_beta'][:, :, k], post['beta'][:, :, k]
        assert abs(np.median(ref) - np.median(mine)) < 0.25 * ref.std(), (k, np.median(ref), np.median(mine))
    # tau: same order of magnitude of the bulk (bimodal / heavy-tailed under the flat default prior)
    assert 0.1 < np.median(post['tau']) / np.median(g['icar_tau']) < 10


def test_icar_wide_band_vs_banded_cholesky():
    """Half bandwidth ~ 101 (100 x 100 lattice, n = 10 000): the wide-band kernel against LAPACK's banded Cholesky
    (scipy.linalg.cholesky_banded) of the same tau Q + diag(omega) -- factor, conditional mean and the projected draw
    (reference: logit.py:73-99; distributions.pyx:24-39)."""
    from scipy.linalg import cholesky_banded, cho_solve_banded, solve_banded
    from occuspytial_b200 import LogitICARGibbs, _lib
    from occuspytial_b200.utils import make_data
    lattice, chains = (100, 100), 2
    n = lattice[0] * lattice[1]
    Q, W, X, y, *_ = make_data(n=n, min_v=1, max_v=4, lattice=lattice, random_state=7)
    s = LogitICARGibbs(Q, W, X, y, random_state=2)
    prob = s._problem_arrays()
    perm, bw = s._band()
    assert 60 < bw <= 110
    Qb = prob['Qband']                      # [n][bw + 1], Qband[i][d] = Q[i][i - d]
    Xp = prob['X']
    rng = np.random.default_rng(5)
    eng = s._make_engine(chains, seed=9, debug=True)
    eta0 = rng.normal(size=(chains, n))
    eta0 -= eta0.mean(axis=1, keepdims=True)
    st = dict(alpha=rng.normal(size=(chains, 3)), beta=rng.normal(size=(chains, 3)), tau=rng.uniform(0.3, 3, chains), eta=eta0)
    eng.set_start(st)
    eng.debug_site_half()
    z = eng.get_field(_lib.F_Z).astype(float)
    omega = eng.get_field(_lib.F_OMEGA)
    normals = rng.normal(size=(chains, n + 3 + 3))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'])
    assert out['fails'] == 0
    eta_new = eng.get_field(_lib.F_ETA)
    for c in range(chains):
        # lower banded storage of scipy: ab[d, j] = A[j + d, j]
        ab = np.zeros((bw + 1, n))
        for d in range(bw + 1):
            ab[d, :n - d] = st['tau'][c] * Qb[d:, d]
        ab[0] += omega[c]
        b = (z[c] - 0.5) - omega[c] * (Xp @ st['beta'][c])
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-11, atol=1e-13)
        cb = cholesky_banded(ab, lower=True)             # cb[d, j] = L[j + d, j]
        Lc = out['eta_chol'][c].reshape(n, bw + 1)
        for d in range(bw + 1):
            np.testing.assert_allclose(Lc[:n - d, d], cb[d, :n - d], rtol=1e-9, atol=1e-12)
        mean = cho_solve_banded((cb, True), b)
        np.testing.assert_allclose(out['eta_mean'][c], mean, rtol=1e-8, atol=1e-10)
        # x = Lambda^-1 b + L^-T eps;  L^T in upper banded form: ub[bw - d, j + d] = L[j + d, j]
        ub = np.zeros((bw + 1, n))
        for d in range(bw + 1):
            ub[bw - d, d:] = cb[d, :n - d]
        x = mean + solve_banded((0, bw), ub, normals[c, :n])
        zv = cho_solve_banded((cb, True), np.ones(n))
        eta_o = x - x.sum() / zv.sum() * zv
        np.testing.assert_allclose(eta_new[c], eta_o, rtol=1e-8, atol=1e-9)
        assert abs(eta_new[c].sum()) < 1e-8
    eng.close()


def test_icar_chain_draws_do_not_depend_on_sharding():
    """Chains 3..4 of a 5-chain job are bit-identical to a 2-chain engine at offset 3."""
    from occuspytial_b200 import LogitICARGibbs
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(n=150, min_v=1, max_v=5, lattice=(10, 15), random_state=6)
    s = LogitICARGibbs(Q, W, X, y, random_state=3)
    start = s._start_to_engine(s._start_arrays(5, None))
    out = []
    for chains, off in ((5, 0), (2, 3)):
        eng = s._make_engine(chains, seed=77, chain_offset=off)
        eng.set_start({k: v[off:off + chains] for k, v in start.items()})
        rec = eng.alloc_record(6)
        assert eng.sweep(6, rec, 0) == 0
        out.append({k: v.cpu().numpy() for k, v in rec.items()})
        eng.close()
    for k in ('alpha', 'beta', 'tau'):
        np.testing.assert_array_equal(out[0][k][3:], out[1][k])
