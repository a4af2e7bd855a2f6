"""Pins the oracle (oracle/sweep.py) against vectors dumped from the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
from scipy import sparse

from conftest import golden
from oracle import sweep as S

RTOL = 1e-11


def _exists_rows(g, z):
    ptr = g['visit_ptr']
    sites = [s for s in g['obs']] + [s for s in g['not_obs'] if z[s]]
    return np.concatenate([np.arange(ptr[s], ptr[s + 1]) for s in sites])


def _rows_no(g):
    ptr = g['visit_ptr']
    return np.concatenate([np.arange(ptr[s], ptr[s + 1]) for s in g['not_obs']])


def test_native_precision_mvnorm_and_sum_to_zero():
    g = golden('native.npz')
    for d in (3, 10, 64):
        x, L = S.precision_mvnorm(g[f'b{d}'], g[f'P{d}'], g[f'z{d}'])
        np.testing.assert_allclose(x, g[f'x{d}'], rtol=1e-9, atol=1e-12)
        # the reference leaves the factor in the lower triangle of `prec` (distributions.pyx:91,100)
        np.testing.assert_allclose(np.tril(g[f'Pafter{d}']), L, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(S.ensure_sums_to_zero(g['stz_x'], g['stz_z']), g['stz_out'], rtol=1e-12, atol=1e-14)


def test_partition_matches_reference():
    g = golden('logit_rsr.npz')
    n = g['X'].shape[0]
    ptr = g['visit_ptr']
    y = {int(s): g['y_rows'][ptr[s]:ptr[s + 1]] for s in g['surveyed_order']}
    part = S.partition_sites(y, n)
    assert part['not_obs'] == g['not_obs'].tolist()
    assert part['obs'] == g['obs'].tolist()
    assert part['not_surveyed'] == g['not_surveyed'].tolist()
    np.testing.assert_array_equal(part['z'], g['z_init'])


def test_logit_rsr_pieces():
    g = golden('logit_rsr.npz')
    X, K, Qrsr = g['X'], g['K'], g['Qrsr']
    z, omega = g['st_z'], g['omega_b']
    k = z - 0.5
    np.testing.assert_allclose(S.tau_rate(g['st_eta'], Qrsr, float(g['tau_rate'])), g['tau_rate_out'], rtol=RTOL)
    Lam, b = S.eta_rsr_system(K, X, omega, k, g['st_beta'], float(g['st_tau']), Qrsr)
    mean = S.gaussian_from_system(Lam, b, np.zeros(K.shape[1]))[1]
    np.testing.assert_allclose(mean, g['eta_mean'], rtol=1e-9, atol=1e-12)
    # the reference's own perturb-then-solve draw, same normals, its eigen-factor
    draw = S.eta_rsr_draw_reference(K, X, omega, k, g['st_beta'], float(g['st_tau']), Qrsr, g['eta_eps'],
                                    eigen=g['eta_eigen'])
    np.testing.assert_allclose(draw, g['eta_draw'], rtol=1e-9, atol=1e-12)
    spatial = K @ g['eta_new']
    A, bb = S.beta_system_logit(X, omega, k, spatial, g['b_prec'], g['b_prec'] @ g['b_mu'])
    np.testing.assert_allclose(S.precision_mvnorm(bb, A, g['beta_normals'])[0], g['beta_out'], rtol=1e-9)
    rows = _exists_rows(g, z)
    A, ba = S.alpha_system_logit(g['W_rows'][rows], g['y_rows'][rows].astype(float), g['omega_a_rows'][rows],
                                 g['a_prec'], g['a_prec'] @ g['a_mu'])
    np.testing.assert_allclose(S.precision_mvnorm(ba, A, g['alpha_normals'])[0], g['alpha_out'], rtol=1e-9)
    p_no, p_ns = S.z_prob_logit(X, g['beta_out'], spatial, g['alpha_out'], g['W_rows'][_rows_no(g)],
                                g['stacked_w_indices'], g['not_obs'], g['not_surveyed'])
    np.testing.assert_allclose(p_no, g['p_no'], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(p_ns, g['p_ns'], rtol=1e-10, atol=1e-15)


def test_logit_icar_pieces():
    g = golden('logit_icar.npz')
    X, Q = g['X'], sparse.csc_matrix(g['Q'])
    z, omega = g['st_z'], g['omega_b']
    k = z - 0.5
    np.testing.assert_allclose(S.tau_rate(g['st_eta'], Q, float(g['tau_rate'])), g['tau_rate_out'], rtol=RTOL)
    Lam, b = S.eta_icar_system(Q, X, omega, k, g['st_beta'], float(g['st_tau']))
    eta = S.eta_icar_draw(Lam, b, np.zeros(X.shape[0]))[0]
    assert abs(eta.sum()) < 1e-9
    # reference: MINRES at its default rtol 1e-5 (logit.py:87) -> agreement only to solver accuracy
    # (residual tolerance; the solution error is larger by the condition number of Lambda)
    assert np.linalg.norm(eta - g["eta_mean"]) / np.linalg.norm(eta) < 5e-3
    spatial = g['eta_new']
    A, bb = S.beta_system_logit(X, omega, k, spatial, g['b_prec'], g['b_prec'] @ g['b_mu'])
    np.testing.assert_allclose(S.precision_mvnorm(bb, A, g['beta_normals'])[0], g['beta_out'], rtol=1e-9)
    p_no, p_ns = S.z_prob_logit(X, g['beta_out'], spatial, g['alpha_out'], g['W_rows'][_rows_no(g)],
                                g['stacked_w_indices'], g['not_obs'], g['not_surveyed'])
    np.testing.assert_allclose(p_no, g['p_no'], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(p_ns, g['p_ns'], rtol=1e-10, atol=1e-15)


def test_probit_rsr_pieces():
    g = golden('probit_rsr.npz')
    X, K, Qrsr = g['X'], g['K'], g['Qrsr']
    z = g['st_z']
    spatial = K @ g['st_eta']
    omega_b = S.probit_omega_b(X, g['st_beta'], spatial, g['st_eps'], z, g['u_omega_b'])
    np.testing.assert_allclose(omega_b, g['omega_b'], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(S.tau_rate(g['st_eta'], Qrsr, float(g['tau_rate'])), g['tau_rate_out'], rtol=RTOL)
    eps = S.probit_eps(omega_b, X, g['st_beta'], spatial, g['eps_normals'])
    np.testing.assert_allclose(eps, g['eps_out'], rtol=1e-11, atol=1e-14)
    A, b = S.probit_eta_system(K, g['KTK'], Qrsr, X, omega_b, g['st_beta'], eps, float(g['st_tau']))
    eta = S.precision_mvnorm(b, A, g['eta_normals'])[0]
    np.testing.assert_allclose(eta, g['eta_out'], rtol=1e-9, atol=1e-12)
    spatial = K @ eta
    A, b = S.probit_beta_system(X, g['XTX_plus_bprec'], omega_b, spatial, eps, g['b_prec'] @ g['b_mu'])
    beta = S.precision_mvnorm(b, A, g['beta_normals'])[0]
    np.testing.assert_allclose(beta, g['beta_out'], rtol=1e-9)
    rows = _exists_rows(g, z)
    W_e, y_e = g['W_rows'][rows], g['y_rows'][rows].astype(float)
    omega_a = S.probit_omega_a(W_e, y_e, g['st_alpha'], g['u_omega_a_rows'][rows])
    np.testing.assert_allclose(omega_a, g['omega_a_rows'][rows], rtol=1e-10, atol=1e-13)
    A, b = S.probit_alpha_system(W_e, omega_a, g['a_prec'], g['a_prec'] @ g['a_mu'])
    alpha = S.precision_mvnorm(b, A, g['alpha_normals'])[0]
    np.testing.assert_allclose(alpha, g['alpha_out'], rtol=1e-9)
    p_no, p_ns = S.z_prob_probit(X, beta, spatial, eps, alpha, g['W_rows'][_rows_no(g)], g['stacked_w_indices'],
                                 g['not_obs'], g['not_surveyed'])
    np.testing.assert_allclose(p_no, g['p_no'], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(p_ns, g['p_ns'], rtol=1e-10, atol=1e-15)


def test_moran_basis_spans_reference_basis():
    """K is unique only up to rotations inside degenerate eigenspaces: compare K'QK spectra and projectors."""
    g = golden('logit_rsr.npz')
    K, Qrsr = S.moran_basis(sparse.csc_matrix(g['Q']), g['X'], q=10)
    np.testing.assert_allclose(np.linalg.eigvalsh(Qrsr), np.linalg.eigvalsh(g['Qrsr']), rtol=1e-8)
    np.testing.assert_allclose(K @ K.T, g['K'] @ g['K'].T, atol=1e-8)


@pytest.mark.parametrize('label', ['rsr'])
def test_oracle_sampler_matches_reference_posterior(label):
    """Oracle CPU chains vs reference chains (golden): posterior means within 3 combined MCSE."""
    from occuspytial_b200.diagnostics import summaries_from_draws, ess
    g = golden('posterior.npz')
    n = g['X'].shape[0]
    ptr = g['visit_ptr']
    sites = [s for s in range(n) if ptr[s + 1] > ptr[s]]
    W = {s: g['W_rows'][ptr[s]:ptr[s + 1]] for s in sites}
    y = {s: g['y_rows'][ptr[s]:ptr[s + 1]] for s in sites}
    hp = {k[3:]: (float(v) if v.ndim == 0 else v) for k, v in g.items() if k.startswith('hp_')}
    out = S.run_chains('logit_rsr', sparse.csc_matrix(g['Q']), W, g['X'], y, 3000, 1000, chains=4, seed=3,
                       hparams=hp, K=g['K'])

    def mean_mcse(x):
        x = x.reshape(x.shape[0], x.shape[1], -1)
        res = []
        for k in range(x.shape[2]):
            xs = x[:, :, k]
            n_eff = ess(summaries_from_draws(xs, 200), xs.shape[1])
            res.append((xs.mean(), xs.std(ddof=1) / np.sqrt(n_eff)))
        return np.array(res)

    for name in ('alpha', 'beta', 'tau'):
        ref = mean_mcse(g[f'{label}_{name}'])
        mine = mean_mcse(out[name])
        err = np.abs(ref[:, 0] - mine[:, 0])
        tol = 3.0 * np.sqrt(ref[:, 1] ** 2 + mine[:, 1] ** 2)
        assert np.all(err < tol), (name, ref, mine)
