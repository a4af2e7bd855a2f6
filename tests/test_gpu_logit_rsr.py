"""GPU parity tests of the LogitRSRGibbs sweep against the reference's golden vectors and the oracle."""
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


def test_gaussian_half_vs_reference_golden():
    """tau rate, eta mean, beta and alpha draws of one iteration reproduce the reference's outputs (1e-9)."""
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from oracle import sweep as S
    g = golden('logit_rsr.npz')
    s = _sampler_from_golden(g, LogitRSRGibbs, K=g['K'], random_state=1)
    assert s.fixed.not_obs == g['not_obs'].tolist() and s.fixed.obs == g['obs'].tolist()
    np.testing.assert_allclose(s.fixed.Q, g['Qrsr'], rtol=1e-12, atol=1e-13)
    q = g['K'].shape[1]
    eng = s._make_engine(3, seed=5, debug=True)
    eng.set_start(dict(alpha=g['st_alpha'], beta=g['st_beta'], tau=g['st_tau'], eta=g['st_eta']))
    eng.set_field(_lib.F_Z, g['st_z'].astype(np.uint8))
    eng.set_field(_lib.F_OMEGA, g['omega_b'])
    # eta: the reference perturbs with n + q normals; only its mean is comparable draw-for-draw.  Use the
    # engine-form draw for eta and check it against the oracle; then force eta_new for beta like the golden did.
    normals = np.concatenate([np.zeros(q), g['beta_normals'], g['alpha_normals']])
    out = eng.debug_gaussian_half(normals=normals, tau=np.full(3, g['st_tau']), omega_a=g['omega_a_rows'])
    assert out['fails'] == 0
    np.testing.assert_allclose(out['tau_rate'], g['tau_rate_out'], rtol=1e-11)
    np.testing.assert_allclose(out['eta_mean'][0], g['eta_mean'], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(eng.get_field(_lib.F_ETA)[1], g['eta_mean'], rtol=1e-9, atol=1e-12)
    Lam, b = S.eta_rsr_system(g['K'], g['X'], g['omega_b'], g['st_z'] - 0.5, g['st_beta'], float(g['st_tau']), g['Qrsr'])
    np.testing.assert_allclose(out['eta_prec'][2], Lam, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out['eta_b'][2], b, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out['eta_chol'][0], np.linalg.cholesky(Lam), rtol=1e-10, atol=1e-13)
    # alpha does not depend on eta: comparable directly
    np.testing.assert_allclose(eng.get_field(_lib.F_ALPHA)[0], g['alpha_out'], rtol=1e-9, atol=1e-12)
    # beta | eta_new: set eta := eta_new, beta back, and rerun with eta's normals/system bypassed is not
    # possible (eta is redrawn), so check beta's system against the oracle for the eta the engine drew ...
    eta_eng = eng.get_field(_lib.F_ETA)[0]
    A, bb = S.beta_system_logit(g['X'], g['omega_b'], g['st_z'] - 0.5, g['K'] @ eta_eng, g['b_prec'], g['b_prec'] @ g['b_mu'])
    np.testing.assert_allclose(out['beta_prec'][0], A, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(out['beta_b'][0], bb, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(eng.get_field(_lib.F_BETA)[0], S.precision_mvnorm(bb, A, g['beta_normals'])[0], rtol=1e-9)
    eng.close()


def test_beta_draw_vs_reference_golden_via_huge_precision():
    """beta_out of the reference is conditional on eta_new.  Pin eta = eta_new in the engine by making the eta
    conditional degenerate: tau -> huge with Qrsr-dominated precision is not exact, so instead run the engine's
    solve with K'k - (K'Om X) beta arranged such that the mean IS eta_new: use the oracle to verify that the
    engine's beta equals the reference's when fed the same eta (already covered), and here verify the
    reference's beta_out against the engine's beta system evaluated at eta_new through a 1-iteration trick:
    set omega, z, then set theta = (beta, eta_new) and ask for the z-update half (psi = X beta + K eta_new)."""
    from occuspytial_b200 import LogitRSRGibbs, _lib
    g = golden('logit_rsr.npz')
    s = _sampler_from_golden(g, LogitRSRGibbs, K=g['K'], random_state=1)
    eng = s._make_engine(2, seed=5, debug=True)
    eng.set_start(dict(alpha=g['alpha_out'], beta=g['beta_out'], tau=g['st_tau'], eta=g['eta_new']))
    eng.set_field(_lib.F_Z, g['st_z'].astype(np.uint8))
    eng.debug_site_half()
    psi = eng.get_field(_lib.F_PSI)[0]
    np.testing.assert_allclose(psi, g['X'] @ g['beta_out'] + g['K'] @ g['eta_new'], rtol=1e-12, atol=1e-13)
    zp = eng.get_field(_lib.F_ZPROB)[1]
    np.testing.assert_allclose(zp[g['not_obs']], g['p_no'], rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(zp[g['not_surveyed']], g['p_ns'], rtol=1e-10, atol=1e-14)
    z = eng.get_field(_lib.F_Z)
    assert np.all(z[:, g['obs']] == 1)
    om = eng.get_field(_lib.F_OMEGA)
    assert np.all(om > 0) and np.all(np.isfinite(om))
    eng.close()


@pytest.mark.parametrize('q,chains,gram_int8', [(7, 5, False), (64, 130, False),
                                                 (230, 3, False), (280, 2, True),     # tiled large-q solve (global scratch)
                                                 (7, 5, True), (64, 130, True)])      # tcgen05 kind::i8 Gram
def test_full_iteration_pieces_vs_oracle(q, chains, gram_int8):
    """Random states, several chains (incl. a padded tail and > 1 chain tile): every system of the Gaussian
    half and the z probabilities agree with the oracle to 1e-10."""
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.utils import make_data
    from oracle import sweep as S
    Q, W, X, y, *_ = make_data(n=300, min_v=1, max_v=6, ns=170, p=4, q=3, random_state=3)
    s = LogitRSRGibbs(Q, W, X, y, random_state=2, q=q)
    f = s.fixed
    prob = s._problem_arrays()
    V, r, p, n = prob['W_rows'].shape[0], 3, 4, f.n
    rng = np.random.default_rng(5)
    eng = s._make_engine(chains, seed=11, debug=True, gram_int8=gram_int8)
    st = dict(alpha=rng.normal(size=(chains, r)), beta=rng.normal(size=(chains, p)), tau=rng.uniform(0.3, 3, chains),
              eta=rng.normal(size=(chains, q)))
    eng.set_start(st)
    eng.debug_site_half()        # draws z and omega_b on the device from (alpha, beta, eta)
    z = eng.get_field(_lib.F_Z).astype(float)
    omega = eng.get_field(_lib.F_OMEGA)
    zp = eng.get_field(_lib.F_ZPROB)
    part = S.partition_sites(y, n)
    Wr, yr, rowptr, site_ids = S.pack_visits(W, y, n)
    pos = {s_: i for i, s_ in enumerate(part['surveyed'])}
    rows_no = np.concatenate([np.arange(rowptr[pos[s_]], rowptr[pos[s_] + 1]) for s_ in part['not_obs']])
    stacked = np.pad(np.cumsum([rowptr[pos[s_] + 1] - rowptr[pos[s_]] for s_ in part['not_obs']]), (1, 0))[:-1]
    check = range(chains) if chains <= 8 else (0, 1, 63, 64, 127, 128, chains - 1)
    for c in check:
        p_no, p_ns = S.z_prob_logit(X, st['beta'][c], f.K @ st['eta'][c], st['alpha'][c], Wr[rows_no], stacked,
                                    part['not_obs'], part['not_surveyed'])
        np.testing.assert_allclose(zp[c][part['not_obs']], p_no, rtol=1e-10, atol=1e-14)
        np.testing.assert_allclose(zp[c][part['not_surveyed']], p_ns, rtol=1e-10, atol=1e-14)
    normals = rng.normal(size=(chains, q + p + r))
    omega_a = rng.uniform(0.02, 0.25, size=(chains, V))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'], omega_a=omega_a)
    assert out['fails'] == 0
    eta_new, beta_new, alpha_new = eng.get_field(_lib.F_ETA), eng.get_field(_lib.F_BETA), eng.get_field(_lib.F_ALPHA)
    vptr, Wd, yd = prob['visit_ptr'], prob['W_rows'], prob['y_rows'].astype(float)
    for c in check:
        k = z[c] - 0.5
        Lam, b = S.eta_rsr_system(f.K, X, omega[c], k, st['beta'][c], st['tau'][c], f.Q)
        np.testing.assert_allclose(out['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-10, atol=1e-11)
        eta_o, mean_o, L_o = S.gaussian_from_system(Lam, b, normals[c, :q])
        np.testing.assert_allclose(out['eta_chol'][c], L_o, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out['eta_mean'][c], mean_o, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(eta_new[c], eta_o, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out['tau_rate'][c], S.tau_rate(st['eta'][c], f.Q, f.tau_rate), rtol=1e-11)
        A, bb = S.beta_system_logit(X, omega[c], k, f.K @ eta_new[c], f.b_prec, f.b_prec_by_mu)
        np.testing.assert_allclose(out['beta_prec'][c], A, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out['beta_b'][c], bb, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(beta_new[c], S.precision_mvnorm(bb, A, normals[c, q:q + p])[0], rtol=1e-9, atol=1e-11)
        rows = np.concatenate([np.arange(vptr[i], vptr[i + 1]) for i in range(n)
                               if f.site_class[i] != 0 and z[c][i] == 1])
        A, ba = S.alpha_system_logit(Wd[rows], yd[rows], omega_a[c][rows], f.a_prec, f.a_prec_by_mu)
        np.testing.assert_allclose(out['alpha_prec'][c], A, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out['alpha_b'][c], ba, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(alpha_new[c], S.precision_mvnorm(ba, A, normals[c, q + p:])[0], rtol=1e-9, atol=1e-11)
    eng.close()


def test_sampler_api_contract():
    """Shapes, reproducibility, burn-in, chains, start and error messages (gibbs/tests/test_samplers.py:54-118)."""
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(min_v=2, max_v=10, ns=100, p=3, q=2, random_state=10)
    s = LogitRSRGibbs(Q, W, X, y, random_state=10)
    samples = s.sample(5, chains=1, progressbar=False)
    assert samples['alpha'].shape == (1, 5, 2) and samples['beta'].shape == (1, 5, 3) and samples['tau'].shape == (1, 5)
    s = LogitRSRGibbs(Q, W, X, y, random_state=10)
    samples2 = s.sample(5, chains=1, progressbar=False)
    for name in ('alpha', 'beta', 'tau'):
        assert np.allclose(samples2[name], samples[name])
    assert isinstance(s.copy(), LogitRSRGibbs)
    with pytest.raises(ValueError, match='burnin value cannot be larger than'):
        s.sample(10, burnin=11)
    samples = s.sample(10, burnin=3, chains=1, progressbar=False)
    assert samples['alpha'].shape == (1, 7, 2)
    with pytest.raises(ValueError, match='chains must a positive integer'):
        s.sample(10, chains=0)
    samples = s.sample(5, chains=3, progressbar=False)
    assert samples['alpha'].shape == (3, 5, 2) and samples['tau'].shape == (3, 5)
    assert not np.allclose(samples['tau'][0], samples['tau'][1])
    rng = np.random.default_rng(0)
    s = LogitRSRGibbs(Q, W, X, y, random_state=10, q=10)
    a = s.sample(5, progressbar=False)
    b = s.sample(5, start={'alpha': rng.random(2), 'beta': rng.random(3), 'tau': 2, 'eta': rng.random(10)},
                 progressbar=False)
    assert not np.allclose(a['alpha'][0, 0], b['alpha'][0, 0])
    assert not np.allclose(a['tau'][0, 0], b['tau'][0, 0])


def test_chain_draws_do_not_depend_on_sharding():
    """Global chain ids key the Philox streams: chains 2..3 of a 4-chain job == a 2-chain engine at offset 2."""
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(random_state=4, max_v=6)
    s = LogitRSRGibbs(Q, W, X, y, random_state=3, q=8)
    start = s._start_arrays(4, None)
    full = s._make_engine(4, seed=77)
    full.set_start(start)
    rec = full.alloc_record(6)
    full.sweep(6, rec, 0)
    shard = s._make_engine(2, seed=77, chain_offset=2)
    shard.set_start({k: v[2:] for k, v in start.items()})
    rec2 = shard.alloc_record(6)
    shard.sweep(6, rec2, 0)
    for k in ('alpha', 'beta', 'tau'):
        np.testing.assert_array_equal(rec[k][2:].cpu().numpy(), rec2[k].cpu().numpy())
    full.close(), shard.close()


@pytest.mark.parametrize('gram_int8', [None, True])
def test_posterior_vs_reference_chains(gram_int8):
    """Posterior means of alpha, beta, tau within 3 MCSE of the reference's chains (golden/posterior.npz), through the
    engine's default path for this size (FP64 DMMA Gram) and with the production path of the headline benchmark forced
    (tcgen05 kind::i8 Gram from the 32-bit omega planes, psi fused into the site pass)."""
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.diagnostics import summaries_from_draws, ess
    g = golden('posterior.npz')
    hp = {k[3:]: (float(v) if v.ndim == 0 else v) for k, v in g.items() if k.startswith('hp_')}
    s = _sampler_from_golden(g, LogitRSRGibbs, K=g['K'], hparams=hp, random_state=21)
    post = s.sample(1500, burnin=500, chains=256, progressbar=False, gram_int8=gram_int8)

    def mean_mcse(x):
        x = x.reshape(x.shape[0], x.shape[1], -1)
        res = []
        for k in range(x.shape[2]):
            xs = x[:, :, k]
            n_eff = ess(summaries_from_draws(xs, 200), xs.shape[1])
            res.append((xs.mean(), xs.std(ddof=1) / np.sqrt(n_eff), np.quantile(xs, [0.05, 0.5, 0.95])))
        return res

    for name in ('alpha', 'beta', 'tau'):
        ref, mine = mean_mcse(g[f'rsr_{name}']), mean_mcse(post[name])
        for (mr, sr, qr), (mm, sm, qm) in zip(ref, mine):
            assert abs(mr - mm) < 3.0 * np.hypot(sr, sm), (name, mr, mm, sr, sm)
            # quantiles: MCSE of a quantile is larger than of the mean; allow 5 x
            assert np.all(np.abs(qr - qm) < 6.0 * np.hypot(sr, sm) + 0.02 * np.abs(qr)), (name, qr, qm)


def test_large_rank_tiled_solve_vs_oracle():
    """q = 600 (19 tile rows: several panel groups per block column of the tiled solve, a 180k-column Gram on the
    tcgen05 path): Lambda, b, the Cholesky factor, the conditional mean and the draw agree with the oracle."""
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.utils import make_data
    from oracle import sweep as S
    q, chains = 600, 2
    Q, W, X, y, *_ = make_data(n=800, min_v=1, max_v=5, ns=400, p=3, q=3, random_state=5)
    s = LogitRSRGibbs(Q, W, X, y, random_state=2, q=q)
    f = s.fixed
    rng = np.random.default_rng(6)
    eng = s._make_engine(chains, seed=12, debug=True, gram_int8=True)
    st = dict(alpha=rng.normal(size=(chains, 3)), beta=rng.normal(size=(chains, 3)), tau=rng.uniform(0.5, 2, chains),
              eta=rng.normal(size=(chains, q)) * 0.3)
    eng.set_start(st)
    eng.debug_site_half()
    z = eng.get_field(_lib.F_Z).astype(float)
    omega = eng.get_field(_lib.F_OMEGA)
    normals = rng.normal(size=(chains, q + 3 + 3))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'])
    assert out['fails'] == 0
    eta_new = eng.get_field(_lib.F_ETA)
    for c in range(chains):
        Lam, b = S.eta_rsr_system(f.K, X, omega[c], z[c] - 0.5, st['beta'][c], st['tau'][c], f.Q)
        np.testing.assert_allclose(out['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12 * np.abs(Lam).max())
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-10, atol=1e-11 * np.abs(b).max())
        eta_o, mean_o, L_o = S.gaussian_from_system(Lam, b, normals[c, :q])
        np.testing.assert_allclose(out['eta_chol'][c], L_o, rtol=1e-9, atol=1e-11 * np.abs(L_o).max())
        np.testing.assert_allclose(out['eta_mean'][c], mean_o, rtol=1e-8, atol=1e-10 * np.abs(mean_o).max())
        np.testing.assert_allclose(eta_new[c], eta_o, rtol=1e-8, atol=1e-10 * np.abs(eta_o).max())
    eng.close()


def test_rank_beyond_1312_with_streamed_planes(monkeypatch):
    """q = 1400 (beyond round 1's limit of 1312): the 985k-column joint Gram goes through STREAMED Z planes (re-formed every
    iteration in groups of column tiles instead of stored: forced here with a 1 GiB store limit and a 0.5 GiB group buffer)
    and the tiled solve runs with 44 tile rows.  Lambda, b, the Cholesky factor, the mean and the draw agree with the
    oracle (reference: logit.py:325-337, 415-460, 478-485)."""
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.utils import make_data
    from oracle import sweep as S
    monkeypatch.setenv('OCCU_TC_STORE_GB', '1')
    monkeypatch.setenv('OCCU_TC_GROUP_GB', '0.5')
    q, chains = 1400, 2
    Q, W, X, y, *_ = make_data(n=1600, min_v=1, max_v=4, lattice=(40, 40), random_state=5)
    s = LogitRSRGibbs(Q, W, X, y, random_state=2, q=q)
    f = s.fixed
    rng = np.random.default_rng(6)
    eng = s._make_engine(chains, seed=12, debug=True, gram_int8=True)
    st = dict(alpha=rng.normal(size=(chains, 3)), beta=rng.normal(size=(chains, 3)), tau=rng.uniform(0.5, 2, chains),
              eta=rng.normal(size=(chains, q)) * 0.1)
    eng.set_start(st)
    eng.debug_site_half()
    z = eng.get_field(_lib.F_Z).astype(float)
    omega = eng.get_field(_lib.F_OMEGA)
    normals = rng.normal(size=(chains, q + 3 + 3))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'])
    assert out['fails'] == 0
    eta_new = eng.get_field(_lib.F_ETA)
    for c in range(chains):
        Lam, b = S.eta_rsr_system(f.K, X, omega[c], z[c] - 0.5, st['beta'][c], st['tau'][c], f.Q)
        np.testing.assert_allclose(out['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12 * np.abs(Lam).max())
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-10, atol=1e-11 * np.abs(b).max())
        eta_o, mean_o, L_o = S.gaussian_from_system(Lam, b, normals[c, :q])
        np.testing.assert_allclose(out['eta_chol'][c], L_o, rtol=1e-9, atol=1e-11 * np.abs(L_o).max())
        np.testing.assert_allclose(out['eta_mean'][c], mean_o, rtol=1e-8, atol=1e-10 * np.abs(mean_o).max())
        np.testing.assert_allclose(eta_new[c], eta_o, rtol=1e-8, atol=1e-10 * np.abs(eta_o).max())
    eng.close()


def test_default_threshold_at_moderate_size():
    """The reference's default call LogitRSRGibbs(Q, W, X, y) (r = 0.5 -> q ~ n / 8, logit.py:409-411) beyond the
    shared-memory solve: n = 2500 gives q = 317.  Runs, keeps every chain positive definite and moves alpha."""
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.utils import make_data
    Q, W, X, y, alpha, *_ = make_data(n=2500, min_v=2, max_v=8, lattice=(50, 50), random_state=0)
    s = LogitRSRGibbs(Q, W, X, y, random_state=10)
    assert s.fixed.q > 256
    post = s.sample(60, burnin=30, chains=32, progressbar=False)
    assert post['alpha'].shape == (32, 30, 3) and np.all(np.isfinite(post['tau'])) and np.all(post['tau'] > 0)
    assert np.linalg.norm(post['alpha'].mean(axis=(0, 1)) - alpha) < 0.6


def test_summary_from_device_blocks_and_multi_engine_run(monkeypatch):
    """sample(): (i) R-hat / ESS / mean / sd of ``summary`` come from per-chain blocks computed on the device and agree
    with the host computation from the draws; (ii) a single-process run over several engines (one per device, driven
    from host threads; two engines on one device when the box has a single GPU) returns bit-identical draws to a
    one-engine run; (iii) the double-buffered record flush returns the same draws as the single-block path.
    Reference: parallel.py:4-42 (fan-out), posterior.py:63-76 (summary), chain.py:56-74 (append)."""
    import torch
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.diagnostics import summary_frame
    from occuspytial_b200.gibbs import base
    from occuspytial_b200.utils import make_data
    Q, W, X, y, *_ = make_data(random_state=4, max_v=6)

    def run(**kw):
        s = LogitRSRGibbs(Q, W, X, y, random_state=3, q=8)
        return s, s.sample(60, burnin=20, chains=300, progressbar=False, **kw)

    s1, one = run(devices=1)
    assert s1._last_summaries is not None and s1._last_summaries['alpha'].shape[:2] == (300, 3)
    host = summary_frame({k: one[k] for k in ('alpha', 'beta', 'tau')})
    dev = one.summary
    assert list(host.index) == list(dev.index)
    for col in ('mean', 'sd', 'mcse_mean', 'ess_bulk', 'r_hat', 'hdi_3%', 'hdi_97%'):
        np.testing.assert_allclose(dev[col].to_numpy(), host[col].to_numpy(), rtol=1e-10, atol=1e-12, err_msg=col)
    ndev = torch.cuda.device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    s2, two = run(devices=devices)
    assert len(s2._last_stats['devices']) == 2 and s2._last_stats['chains'] == 300
    for k in ('alpha', 'beta', 'tau'):
        np.testing.assert_array_equal(one[k], two[k])
    np.testing.assert_allclose(two.summary['r_hat'].to_numpy(), dev['r_hat'].to_numpy(), rtol=1e-10)
    if ndev >= 2:     # the default picks every visible GPU that gets >= 128 chains
        s3, auto = run()
        assert len(s3._last_stats['devices']) == min(ndev, 300 // 128)
        np.testing.assert_array_equal(one['tau'], auto['tau'])
    # record capacity of 7 draws: six flushes through the two alternating blocks
    monkeypatch.setattr(base, '_RECORD_BYTES', 8 * 300 * 7 * 7)
    s4, flushed = run(devices=1)
    for k in ('alpha', 'beta', 'tau'):
        np.testing.assert_array_equal(one[k], flushed[k])
    assert s4._last_summaries is None and np.isfinite(flushed.summary['r_hat'].to_numpy()).all()
