"""Helpers mirroring occuspytial/utils.py: ``get_generator``, ``rand_precision_mat``, ``make_data``.

These are input generators, not part of the sweep.  They are restated so that they work without
libpysal (absent here) and scale to n >= 10^4 sites (the reference builds a dense pseudo-inverse,
utils.py:277-278).  ``make_data`` keeps the reference's signature, defaults and error messages
(utils.py:106-294; tested by occuspytial/tests/test_utils.py).
"""
import numpy as np
from scipy import sparse


def get_generator(random_state=None):
    """NumPy Generator on the SFC64 bit generator (utils.py:7-35)."""
    return np.random.default_rng(np.random.SFC64(random_state))


def lattice_adjacency(lat_row, lat_col, queen=True):
    """Contiguity matrix of a lat_row x lat_col lattice, site id = i * lat_col + j (what
    libpysal.weights.lat2SW returns at utils.py:98)."""
    n = lat_row * lat_col
    ids = np.arange(n)
    i, j = ids // lat_col, ids % lat_col
    pairs = [(j < lat_col - 1, 1), (i < lat_row - 1, lat_col)]
    if queen:
        pairs += [((i < lat_row - 1) & (j < lat_col - 1), lat_col + 1), ((i < lat_row - 1) & (j > 0), lat_col - 1)]
    src = np.concatenate([ids[m] for m, _ in pairs])
    dst = np.concatenate([ids[m] + off for m, off in pairs])
    A = sparse.coo_matrix((np.ones(src.size, dtype=np.int64), (src, dst)), shape=(n, n)).tocsr()
    return (A + A.T).tocsr()


def rand_precision_mat(lat_row, lat_col, max_neighbors=8, rho=1):
    """Spatial precision matrix Q = D - rho * A of a rectangular lattice, COO format (utils.py:38-103)."""
    if max_neighbors not in (4, 8):
        raise ValueError('Maximum number of neighbors should be one of {4, 8}')
    A = lattice_adjacency(lat_row, lat_col, queen=(max_neighbors == 8)).tocoo()
    D = np.asarray(A.sum(axis=1)).ravel()
    A.data = -A.data * rho
    A.setdiag(D)
    return A.tocoo()


def _icar_draw(Q, tau, rng):
    """eta ~ N(0, (tau Q)^+) for a connected-lattice ICAR precision without forming Q^+.

    Q = B'B with B the signed edge-incidence matrix, so x = Q^+ B' xi / sqrt(tau), xi ~ N(0, I), has
    covariance Q^+ / tau.  Q^+ B' xi is obtained from a sparse solve of the regularised system and
    centred (the null space of Q is the constant vector)."""
    from scipy.sparse.linalg import splu
    n = Q.shape[0]
    up = sparse.triu(-sparse.csr_matrix(Q), k=1).tocoo()
    m = up.nnz
    wts = np.sqrt(up.data)
    B = sparse.coo_matrix((np.concatenate([wts, -wts]),
                           (np.concatenate([np.arange(m), np.arange(m)]), np.concatenate([up.row, up.col]))),
                          shape=(m, n)).tocsc()
    xi = rng.standard_normal(m)
    lu = splu((sparse.csc_matrix(Q, dtype=np.float64) + 1e-8 * sparse.identity(n, format='csc')))
    eta = lu.solve(B.T @ xi) / np.sqrt(tau)
    return eta - eta.mean()


def make_data(n=150, min_v=None, max_v=None, ns=None, p=3, q=3, tau_range=(0.25, 1.5), max_neighbors=8,
              random_state=None, lattice=None):
    """Generate random data for a single-season spatial occupancy model (utils.py:106-294).

    Same parameters as the reference plus ``lattice=(rows, cols)`` to fix the lattice shape instead
    of drawing a random factorisation of ``n`` (needed for reproducible large benchmarks).
    Returns ``Q, W, X, y, alpha, beta, tau, z``.
    """
    rng = get_generator(random_state)
    if n < 150:
        raise ValueError('n cant be lower than 150')
    if min_v is None:
        min_v = 2
    elif min_v < 1:
        raise ValueError('min_v needs to be at least 1')
    if max_v is None:
        max_v = n // 10
    elif max_v < 2:
        raise ValueError('max_v is too small')
    elif max_v > n:
        raise ValueError('max_v cant be more than n')
    if ns is None:
        ns = n // 2
    elif ns == 0:
        raise ValueError('ns should be positive')
    elif ns > n:
        raise ValueError('ns cant be more than n')

    surveyed_sites = rng.choice(n, size=ns, replace=False)
    visits_per_site = rng.integers(min_v, max_v, size=ns, endpoint=True)
    alpha = rng.standard_normal(q)
    beta = rng.standard_normal(p)
    tau = rng.uniform(*tau_range)
    if lattice is None:
        factors = [f for f in range(3, n) if n % f == 0]
        row = int(rng.choice(factors))
        col = n // row
    else:
        row, col = lattice
        if row * col != n:
            raise ValueError('lattice shape does not match n')
    Q = rand_precision_mat(row, col, max_neighbors=max_neighbors).astype(float)
    eta = _icar_draw(Q, tau, rng)

    X = rng.uniform(-2, 2, n * p).reshape(n, -1)
    X[:, 0] = 1
    psi = np.exp(-np.logaddexp(0, -X @ beta + eta))
    z = rng.binomial(1, p=psi, size=n)

    W, y = {}, {}
    for site, visits in zip(surveyed_sites, visits_per_site):
        w = rng.uniform(-2, 2, size=visits * q).reshape(visits, -1)
        w[:, 0] = 1
        d = np.exp(-np.logaddexp(0, -w @ alpha))
        W[int(site)] = w
        y[int(site)] = rng.binomial(1, z[site] * d)
    return Q, W, X, y, alpha, beta, tau, z
