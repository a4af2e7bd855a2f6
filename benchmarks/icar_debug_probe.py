import numpy as np, sys
sys.path.insert(0, '.')
from occuspytial_b200 import LogitICARGibbs, _lib
from occuspytial_b200.utils import make_data
from scipy import sparse
lattice=(10,15); n=150
Q, W, X, y, *_ = make_data(n=n, min_v=1, max_v=5, lattice=lattice, random_state=6)
s = LogitICARGibbs(Q, W, X, y, random_state=2)
prob = s._problem_arrays(); perm, bw = s._band()
Qp = sparse.csr_matrix(Q)[perm][:, perm].toarray(); Xp = prob['X']
chains=2
rng = np.random.default_rng(3)
eng = s._make_engine(chains, seed=9, debug=True)
eta0 = rng.normal(size=(chains, n)); eta0 -= eta0.mean(axis=1, keepdims=True)
st = dict(alpha=rng.normal(size=(chains, 3)), beta=rng.normal(size=(chains, 3)), tau=rng.uniform(0.3, 3, chains), eta=eta0)
eng.set_start(st); eng.debug_site_half()
z = eng.get_field(_lib.F_Z).astype(float); omega = eng.get_field(_lib.F_OMEGA)
normals = rng.normal(size=(chains, n + 6))
out = eng.debug_gaussian_half(normals=normals, tau=st['tau'])
print('bw', bw, 'fails', out['fails'])
for c in range(chains):
    Lam = st['tau'][c]*Qp + np.diag(omega[c]); b = (z[c]-0.5) - omega[c]*(Xp@st['beta'][c])
    print('eta_b err', np.abs(out['eta_b'][c]-b).max())
    L = np.linalg.cholesky(Lam)
    Lc = out['eta_chol'][c].reshape(n, bw+1)
    errs=[np.abs(Lc[:n-e,e]-np.diagonal(L,-e)).max() for e in range(bw+1)]
    print('chol err per offset', np.round(errs,3))
    bad=[(j,e) for e in range(bw+1) for j in range(n-e) if abs(Lc[j,e]-L[j+e,j])>1e-8]
    print('first bad (col j, offset e):', bad[:12], 'count', len(bad))
    mean=np.linalg.solve(Lam,b)
    em=np.abs(out['eta_mean'][c]-mean)
    print('mean err max', em.max(), 'first bad idx', np.flatnonzero(em>1e-8)[:10])
eng.close()
