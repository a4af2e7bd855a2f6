"""One-off host-side setup feeding the sweep: site partition, CSR packing, RSR basis, banded ICAR Q.

Replaces the setup half of GibbsBase._configure (base.py:107-152), LogitRSRGibbs._configure_rsr
(logit.py:415-460) and ProbitRSRGibbs._configure (probit.py:115-148).  Vectorised: the reference's
``site not in surveyed`` list scans are O(n * n_s) (base.py:125-127) and its Moran operator is a
dense n x n eigen-problem (logit.py:422-433).
"""
import numpy as np
from scipy import sparse
from scipy.sparse.linalg import LinearOperator, eigsh

DENSE_LIMIT = 3000  # up to this many sites the reference's dense eigh is reproduced exactly
DEVICE_DENSE_LIMIT = 20000  # up to this many sites the dense eigen-problem is solved on the GPU when many pairs are wanted


def partition(y_data, n):
    """Initial z and the obs / not_obs / not_surveyed partition (base.py:113-137), dict order preserved."""
    surveyed = np.asarray(y_data.surveyed, dtype=np.int64)
    detected = np.fromiter((bool(np.any(y_data[int(s)])) for s in surveyed), dtype=bool, count=surveyed.size)
    z = np.ones(n)
    z[surveyed] = detected
    mask = np.zeros(n, dtype=bool)
    mask[surveyed] = True
    site_class = np.zeros(n, dtype=np.uint8)          # OCCU_SITE_NOT_SURVEYED
    site_class[surveyed[~detected]] = 1               # OCCU_SITE_NOT_OBS
    site_class[surveyed[detected]] = 2                # OCCU_SITE_OBS
    return dict(
        z=z,
        site_class=site_class,
        not_surveyed=np.nonzero(~mask)[0].tolist(),
        not_obs=surveyed[~detected].tolist(),
        obs=surveyed[detected].tolist(),
    )


def moran_operator(Q, X):
    """Return (matvec-able LinearOperator, scale) of n * P A P / sum(A), P = I - X (X'X)^-1 X' (logit.py:417-428)."""
    Q = sparse.csr_matrix(Q)
    n = Q.shape[0]
    A = -Q
    A = (A - sparse.diags(A.diagonal())).tocsr()
    A.eliminate_zeros()
    total = A.sum()
    XtX_inv = np.linalg.inv(X.T @ X)

    def project(v):
        return v - X @ (XtX_inv @ (X.T @ v))

    def matvec(v):
        v = np.asarray(v, dtype=np.float64).reshape(n, -1)
        return (n / total) * project(A @ project(v))

    return LinearOperator((n, n), matvec=matvec, matmat=matvec, dtype=np.float64), A, XtX_inv


def rsr_basis(Q, X, r=0.5, q=None):
    """RSR basis K (n x q): leading eigenvectors of the Moran operator (logit.py:430-447).

    n <= DENSE_LIMIT: dense ``eigh`` exactly as the reference.  Larger n: Lanczos (``eigsh``, 'LA')
    on the implicit operator; with only a threshold ``r`` the number of eigenpairs is grown until the
    smallest one falls below ``r``.  When many pairs are wanted (q > 256, or a threshold, which at the default
    r = 0.5 keeps ~ n / 8 of them) and a CUDA device is present, the dense problem is solved on the device instead
    (FP64 ``eigh``, seconds at n = 10^4 where the reference's host ``eigh`` takes ten minutes and Lanczos with
    > 1000 pairs is slower still) -- SURVEY.md 8(f) row 1.
    """
    n = X.shape[0]
    if q is None or not q:
        if not 0 <= r <= 1:
            raise ValueError('Threshold value needs to be in [0, 1]')
    op, A, XtX_inv = moran_operator(Q, X)
    if n <= DENSE_LIMIT:
        P = -np.linalg.multi_dot([X, XtX_inv, X.T])
        P[np.diag_indices_from(P)] += 1
        omega = n * (P.T @ (A @ P)) / A.sum()
        w, v = np.linalg.eigh(omega)
        if q:
            qq = int(q)
        else:
            qq = int(w[w >= r].size)
            if not qq:
                raise ValueError('The Moran Operator Matrix of the data has no positive eigenvalues. '
                                 'Set threshold to a lower value')
        return np.ascontiguousarray(v[:, -qq:])
    if (not q or int(q) > 256) and n <= DEVICE_DENSE_LIMIT:
        dense = _device_moran_eigh(A, X, XtX_inv)
        if dense is not None:
            w, v = dense
            if q:
                return np.ascontiguousarray(v[:, -int(q):])
            keep = w >= r
            if not keep.any():
                raise ValueError('The Moran Operator Matrix of the data has no positive eigenvalues. '
                                 'Set threshold to a lower value')
            return np.ascontiguousarray(v[:, keep])
    if q:
        w, v = eigsh(op, k=int(q), which='LA')
        order = np.argsort(w)
        return np.ascontiguousarray(v[:, order])
    k = min(max(64, n // 16), n - 2)
    while True:
        w, v = eigsh(op, k=k, which='LA')
        order = np.argsort(w)
        w, v = w[order], v[:, order]
        if w[0] < r or k >= n - 2:
            keep = w >= r
            if not keep.any():
                raise ValueError('The Moran Operator Matrix of the data has no positive eigenvalues. '
                                 'Set threshold to a lower value')
            return np.ascontiguousarray(v[:, keep])
        k = min(2 * k, n - 2)


def _device_moran_eigh(A, X, XtX_inv):
    """All eigenpairs (ascending) of n * P A P / sum(A) by a dense FP64 ``eigh`` on the current CUDA device, or None
    when no device (or not enough device memory) is available.  Set-up only: this is a plain library call."""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        n = A.shape[0]
        dev = torch.device('cuda')
        Xd = torch.as_tensor(np.ascontiguousarray(X), dtype=torch.float64, device=dev)
        Ad = torch.as_tensor(A.toarray(), dtype=torch.float64, device=dev)
        Pd = -(Xd @ torch.as_tensor(XtX_inv, dtype=torch.float64, device=dev) @ Xd.T)
        Pd.diagonal().add_(1.0)
        omega = (Pd.T @ (Ad @ Pd)) * (n / float(A.sum()))
        del Ad, Pd
        omega = 0.5 * (omega + omega.T)
        w, v = torch.linalg.eigh(omega)
        out = w.cpu().numpy(), v.cpu().numpy()
        del omega, w, v
        torch.cuda.empty_cache()
        return out
    except (ImportError, RuntimeError):
        return None


def verify_singular(Q):
    """Robust version of GibbsBase._verify_spatial_precision (base.py:166-170).

    The reference's shift-invert ``eigsh(sigma=1e-3, which='SA')`` returns a positive eigenvalue whenever
    0 < lambda_2 < 1e-3 and then rejects valid ICAR matrices of large lattices.  Here: an exactly
    row-sum-zero Q is singular by construction; otherwise the eigenvalue nearest to zero is computed.
    """
    Qs = sparse.csc_matrix(Q, dtype=np.float64)
    n = Qs.shape[0]
    scale = abs(Qs).max() if Qs.nnz else 1.0
    if np.abs(Qs @ np.ones(n)).max() <= 1e-10 * max(scale, 1.0):
        return
    if n <= 2000:
        lam = np.linalg.eigvalsh(Qs.toarray())
        smallest = lam[np.argmin(np.abs(lam))]
    else:
        smallest = eigsh(Qs, k=1, sigma=-1e-6, which='LM', return_eigenvectors=False)[0]
    if smallest >= 1e-4:
        raise ValueError('Spatial precision matrix Q must be singular.')


def band_order(Q, reorder=True):
    """Site ordering that keeps the band of Q narrow: reverse Cuthill-McKee if it beats the given order.

    Returns (perm, half_bandwidth) with new site j = old site perm[j].  This is the one-off symbolic
    analysis of the ICAR precision: the sparsity pattern of tau Q + diag(omega) never changes, so the banded
    Cholesky of every chain and iteration reuses it.
    """
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    Qc = sparse.csr_matrix(Q)
    n = Qc.shape[0]

    def half_bandwidth(M):
        coo = M.tocoo()
        return int(np.abs(coo.row - coo.col).max()) if coo.nnz else 0

    ident = np.arange(n)
    bw0 = half_bandwidth(Qc)
    if not reorder:
        return ident, bw0
    perm = np.asarray(reverse_cuthill_mckee(Qc, symmetric_mode=True), dtype=np.int64)
    bw1 = half_bandwidth(Qc[perm][:, perm])
    return (perm, bw1) if bw1 < bw0 else (ident, bw0)


def lower_band(Q, perm, bw):
    """Qband[i][d] = Q'[i][i - d] for the permuted matrix Q' = Q[perm][:, perm], d = 0..bw."""
    Qp = sparse.csr_matrix(Q)[perm][:, perm].tocoo()
    n = Qp.shape[0]
    band = np.zeros((n, bw + 1))
    keep = Qp.row >= Qp.col
    band[Qp.row[keep], Qp.row[keep] - Qp.col[keep]] = Qp.data[keep]
    return band
