"""Full-size checks (BASELINE.json metric configuration: LogitRSRGibbs, n = 10 000 sites, q = 64, 4096 chains)
through size-independent properties: the oracle cannot run this size in seconds, invariants can."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def big():
    from occuspytial_b200 import LogitRSRGibbs
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    Q, W, X, y, alpha, beta, tau, z = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
    K = _setup.rsr_basis(Q, X, q=64)
    return LogitRSRGibbs(Q, W, X, y, random_state=10, K=K), alpha


def test_full_size_invariants(big):
    from occuspytial_b200 import _lib
    from oracle import sweep as S
    s, alpha_true = big
    C = 4096
    eng = s._make_engine(C, seed=99)
    start = s._start_arrays(C, None)
    eng.set_start(start)
    rec = eng.alloc_record(25)
    assert eng.sweep(25, rec, 0) == 0                       # no chain lost positive-definiteness
    draws = {k: v.cpu().numpy() for k, v in rec.items()}
    assert all(np.all(np.isfinite(v)) for v in draws.values()) and np.all(draws['tau'] > 0)
    # state invariants
    z = eng.get_field(_lib.F_Z)
    assert set(np.unique(z)) <= {0, 1}
    cls = s.fixed.site_class
    assert np.all(z[:, cls == 2] == 1)                      # detection implies occupancy (base.py:118)
    assert 0 < z[:, cls == 1].mean() < 1 and 0 < z[:, cls == 0].mean() < 1
    # psi = X beta + K eta on every chain, to rounding (the DMMA GEMM at full size; a non-debug engine computes psi
    # inside the site pass, so reading the field evaluates it for the current state)
    beta, eta = eng.get_field(_lib.F_BETA), eng.get_field(_lib.F_ETA)
    psi = eng.get_field(_lib.F_PSI)
    ref = beta @ s.X.T + eta @ s.fixed.K.T
    np.testing.assert_allclose(psi, ref, rtol=1e-11, atol=1e-11)
    # omega_b ~ PG(1, psi) of the sweep that starts from this state (psi formed by the fused DMMA product of the site
    # pass): law of large numbers over 4 x 10^7 draws against the analytic mean / variance
    eng.sweep(1)
    omega = eng.get_field(_lib.F_OMEGA)
    assert np.all(omega > 0) and np.all(np.isfinite(omega))
    resid = (omega - S.pg1_mean(psi)) / np.sqrt(S.pg1_var(psi))
    assert abs(resid.mean()) < 1e-3 and abs(resid.var() - 1) < 5e-3
    # the chains have moved towards the data-generating alpha (loose sanity, 26 iterations from diffuse starts)
    assert np.linalg.norm(np.median(draws['alpha'][:, -1], axis=0) - alpha_true) < 0.5
    eng.close()


def test_full_size_sharding_and_determinism(big):
    """Chains 128..255 of a 256-chain job are bit-identical to a 128-chain engine at offset 128 (what a second
    GPU would run), and a rerun reproduces the draws bit for bit."""
    s, _ = big
    start = s._start_arrays(256, None)
    out = []
    for chains, off in ((256, 0), (128, 128), (256, 0)):
        eng = s._make_engine(chains, seed=5, chain_offset=off)
        eng.set_start({k: v[off:off + chains] for k, v in start.items()})
        rec = eng.alloc_record(4)
        eng.sweep(4, rec, 0)
        out.append({k: v.cpu().numpy() for k, v in rec.items()})
        eng.close()
    for k in ('alpha', 'beta', 'tau'):
        np.testing.assert_array_equal(out[0][k][128:], out[1][k])
        np.testing.assert_array_equal(out[0][k], out[2][k])


@pytest.mark.parametrize('shape,q,chains,p,r', [
    ((100, 100), 64, 3, 3, 3),     # headline shape: one K chunk, G written directly
    ((100, 120), 12, 5, 3, 3),     # n = 12000: two K chunks + fixed-order reduction
    ((33, 34), 33, 200, 3, 3),     # ragged everything: n = 1122, 2 chain tiles (one partial)
    ((37, 57), 100, 129, 3, 3),    # n = 2109, 5253 Gram columns, 129 chains
    ((32, 40), 17, 70, 5, 2),      # other covariate counts (p = 5, r = 2): other template instances, pq = 22
    ((200, 250), 12, 3, 3, 3)])    # BASELINE config 5's site count: n = 50 000, five K chunks
def test_tcgen05_gram_at_full_size_vs_numpy(shape, q, chains, p, r, monkeypatch):
    """The Omega-weighted Gram of the tcgen05 kind::i8 path (csrc/gram_tc.cuh) at sizes the small parity cases do
    not reach: every Gaussian system built from it agrees with a NumPy FP64 evaluation to 1e-10
    (reference: logit.py:226-232, 325-337, 478-485)."""
    if q == 100:   # this case also runs with STREAMED Z planes: four groups of column tiles through a 20 MB buffer
        monkeypatch.setenv('OCCU_TC_STORE_GB', '0.001')
        monkeypatch.setenv('OCCU_TC_GROUP_GB', '0.02')
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    from oracle import sweep as S
    n = shape[0] * shape[1]
    Q, W, X, y, *_ = make_data(n=n, min_v=2, max_v=6, lattice=shape, p=p, q=r, random_state=1)
    K = _setup.rsr_basis(Q, X, q=q)
    s = LogitRSRGibbs(Q, W, X, y, random_state=4, K=K)
    f = s.fixed
    prob = s._problem_arrays()
    V = prob['W_rows'].shape[0]
    rng = np.random.default_rng(8)
    eng = s._make_engine(chains, seed=3, debug=True)     # n >= 1024: the tcgen05 path is the default
    assert eng.gram_int8
    st = dict(alpha=rng.normal(size=(chains, r)), beta=rng.normal(size=(chains, p)), tau=rng.uniform(0.3, 3, chains),
              eta=rng.normal(size=(chains, q)))
    eng.set_start(st)
    eng.debug_site_half()                                # z and omega_b drawn on the device, written as byte planes
    z = eng.get_field(_lib.F_Z).astype(float)
    omega = eng.get_field(_lib.F_OMEGA)
    assert np.all(omega > 0) and np.all(np.isfinite(omega))
    normals = rng.normal(size=(chains, q + p + r))
    omega_a = rng.uniform(0.02, 0.25, size=(chains, V))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'], omega_a=omega_a)
    assert out['fails'] == 0
    eta_new = eng.get_field(_lib.F_ETA)
    for c in range(chains):
        k = z[c] - 0.5
        Lam, b = S.eta_rsr_system(f.K, X, omega[c], k, st['beta'][c], st['tau'][c], f.Q)
        scale = np.abs(Lam).max()
        np.testing.assert_allclose(out['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12 * scale)
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-10, atol=1e-11 * np.abs(b).max())
        A, bb = S.beta_system_logit(X, omega[c], k, f.K @ eta_new[c], f.b_prec, f.b_prec_by_mu)
        np.testing.assert_allclose(out['beta_prec'][c], A, rtol=1e-10, atol=1e-12 * np.abs(A).max())
        np.testing.assert_allclose(out['beta_b'][c], bb, rtol=1e-9, atol=1e-10 * np.abs(bb).max())
    # injected omega / z (set_field) go through the slicing kernel instead of the site pass: same systems
    om2 = rng.uniform(0.01, 0.4, size=omega.shape)
    z2 = (rng.random(z.shape) < 0.5).astype(np.uint8)
    eng.set_start(st)
    eng.set_field(_lib.F_OMEGA, om2)
    eng.set_field(_lib.F_Z, z2)
    out2 = eng.debug_gaussian_half(normals=normals, tau=st['tau'], omega_a=omega_a)
    for c in range(chains):
        Lam, b = S.eta_rsr_system(f.K, X, om2[c], z2[c] - 0.5, st['beta'][c], st['tau'][c], f.Q)
        np.testing.assert_allclose(out2['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12 * np.abs(Lam).max())
        np.testing.assert_allclose(out2['eta_b'][c], b, rtol=1e-10, atol=1e-11 * np.abs(b).max())
    eng.close()


def test_device_moran_basis_is_an_eigenbasis():
    """Large RSR bases (q > 256 beyond the dense-host limit) come from a dense FP64 eigh on the device
    (gibbs/_setup.py:_device_moran_eigh; reference: logit.py:417-447).  The returned columns are orthonormal
    eigenvectors of the Moran operator n P A P / 1'A1 with the q largest eigenvalues, ascending."""
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    Q, W, X, y, *_ = make_data(n=3500, min_v=2, max_v=4, lattice=(50, 70), random_state=2)
    q = 300
    K = _setup.rsr_basis(Q, X, q=q)
    assert K.shape == (3500, q)
    op, A, XtX_inv = _setup.moran_operator(Q, X)
    OK = op.matmat(K)
    lam = np.einsum('ij,ij->j', K, OK)
    np.testing.assert_allclose(K.T @ K, np.eye(q), atol=1e-10)
    assert np.max(np.linalg.norm(OK - K * lam, axis=0)) < 1e-9 * max(1.0, np.abs(lam).max())
    assert np.all(np.diff(lam) > -1e-9)
    # nothing outside the span has a larger eigenvalue than the smallest one kept: power iteration on the deflated operator
    rng = np.random.default_rng(0)
    v = rng.normal(size=3500)
    for _ in range(200):
        v -= K @ (K.T @ v)
        v = op.matvec(v).ravel()
        v /= np.linalg.norm(v)
    v -= K @ (K.T @ v)
    v /= np.linalg.norm(v)
    top_outside = float(v @ op.matvec(v).ravel())
    assert top_outside <= lam[0] + 1e-6


def test_omega_planes_at_large_arguments():
    """The 32-bit omega planes of the production precision at PG arguments up to |psi| = 60 (omega ~ 1 / (2 |psi|), below
    2^-6: the regime where the fixed-point value is ROUNDED to 2^-29 rather than exact), and the saturation of the
    48-bit slicing of an injected omega >= 8.  Reference: logit.py:195-204 (omega_b), 325-337 (its Gram)."""
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    from oracle import sweep as S
    shape, q, chains = (32, 40), 9, 256
    n = shape[0] * shape[1]
    Q, W, X, y, *_ = make_data(n=n, min_v=2, max_v=5, lattice=shape, random_state=3)
    K = _setup.rsr_basis(Q, X, q=q)
    s = LogitRSRGibbs(Q, W, X, y, random_state=4, K=K)
    f = s.fixed
    rng = np.random.default_rng(2)
    eng = s._make_engine(chains, seed=3, debug=True, gram_int8=True)
    # column 0 of X is the intercept: beta_0 sets the level of psi; chains span |psi| ~ 0 .. 60, both signs
    level = np.linspace(-60, 60, chains)
    beta = np.zeros((chains, 3))
    beta[:, 0] = level
    st = dict(alpha=rng.normal(size=(chains, 3)), beta=beta, tau=np.ones(chains), eta=0.1 * rng.normal(size=(chains, q)))
    eng.set_start(st)
    eng.debug_site_half()
    psi = eng.get_field(_lib.F_PSI)
    omega = eng.get_field(_lib.F_OMEGA)          # the value the byte planes hold (debug engines write it from the digits)
    assert np.abs(psi).max() > 55
    assert np.all(omega > 0) and np.all(omega < 8)
    scaled = omega * 2.0 ** 29
    np.testing.assert_array_equal(scaled, np.rint(scaled))      # a 32-bit fixed-point number of omega / 8
    assert (omega < 2.0 ** -6).mean() > 0.3                     # the rounded regime is exercised
    # law of large numbers per |psi| band: PG(1, psi) mean and variance survive the rounding
    resid = (omega - S.pg1_mean(psi)) / np.sqrt(S.pg1_var(psi))
    for lo, hi in ((0, 5), (5, 20), (20, 40), (40, 61)):
        m = (np.abs(psi) >= lo) & (np.abs(psi) < hi)
        assert m.sum() > 10000
        assert abs(resid[m].mean()) < 4 / np.sqrt(m.sum()) + 1e-3, (lo, hi, resid[m].mean())
        assert abs(resid[m].var() - 1) < 0.05, (lo, hi, resid[m].var())
    # the Gram built from those planes is the FP64 Gram of exactly these omega
    z = eng.get_field(_lib.F_Z).astype(float)
    normals = rng.normal(size=(chains, q + 6))
    out = eng.debug_gaussian_half(normals=normals, tau=st['tau'])
    assert out['fails'] == 0
    for c in (0, 1, chains // 2, chains - 1):
        Lam, b = S.eta_rsr_system(f.K, X, omega[c], z[c] - 0.5, st['beta'][c], st['tau'][c], f.Q)
        np.testing.assert_allclose(out['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12 * np.abs(Lam).max())
        np.testing.assert_allclose(out['eta_b'][c], b, rtol=1e-10, atol=1e-11 * np.abs(b).max())
    # injected omega beyond the fixed-point range saturates at 8 - 2^-45 instead of wrapping
    om2 = rng.uniform(0.01, 0.4, size=omega.shape)
    om2[:, ::7] = 8.0 + rng.uniform(0, 100, size=om2[:, ::7].shape)
    eng.set_start(st)
    eng.set_field(_lib.F_OMEGA, om2)
    out2 = eng.debug_gaussian_half(normals=normals, tau=st['tau'])
    z2 = eng.get_field(_lib.F_Z).astype(float)
    clipped = np.minimum(om2, 8.0 - 2.0 ** -45)
    for c in (0, chains - 1):
        Lam, b = S.eta_rsr_system(f.K, X, clipped[c], z2[c] - 0.5, st['beta'][c], st['tau'][c], f.Q)
        # b is formed from omega as well: only Lambda is a pure function of the sliced value
        np.testing.assert_allclose(out2['eta_prec'][c], Lam, rtol=1e-10, atol=1e-12 * np.abs(Lam).max())
    eng.close()
