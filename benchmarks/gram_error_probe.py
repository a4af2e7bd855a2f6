"""Error of the tcgen05 Gram path against a NumPy FP64 / long-double evaluation at the headline shape.

  OCCU_B200_LIB=<variant.so> python benchmarks/gram_error_probe.py
Prints, over a few chains, max |Lam_dev - Lam| / max|Lam| and the largest relative error among the entries above
1e-6 max|Lam| (Lam = K' Omega K + tau K'QK, the eta precision, logit.py:325-337).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit('/', 2)[0])


def main():
    from occuspytial_b200 import LogitRSRGibbs, _lib
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    for shape, q in (((100, 100), 64), ((37, 57), 100), ((100, 120), 12)):
        n = shape[0] * shape[1]
        Q, W, X, y, *_ = make_data(n=n, min_v=2, max_v=6, lattice=shape, p=3, q=3, random_state=1)
        K = _setup.rsr_basis(Q, X, q=q)
        s = LogitRSRGibbs(Q, W, X, y, random_state=4, K=K)
        f = s.fixed
        chains = 4
        rng = np.random.default_rng(8)
        eng = s._make_engine(chains, seed=3, debug=True)
        st = dict(alpha=rng.normal(size=(chains, 3)), beta=rng.normal(size=(chains, 3)), tau=rng.uniform(0.3, 3, chains),
                  eta=rng.normal(size=(chains, q)))
        eng.set_start(st)
        eng.debug_site_half()
        omega = eng.get_field(_lib.F_OMEGA)
        V = s._problem_arrays()['W_rows'].shape[0]
        out = eng.debug_gaussian_half(normals=rng.normal(size=(chains, q + 6)), tau=st['tau'],
                                      omega_a=rng.uniform(0.02, 0.25, size=(chains, V)))
        Kl = f.K.astype(np.longdouble)
        worst_abs, worst_rel, worst_beta = 0.0, 0.0, 0.0
        for c in range(chains):
            Lam = np.asarray((Kl.T * omega[c].astype(np.longdouble)) @ Kl + np.longdouble(st['tau'][c]) * f.Q, dtype=np.longdouble)
            err = np.abs(out['eta_prec'][c].astype(np.longdouble) - Lam)
            scale = np.abs(Lam).max()
            worst_abs = max(worst_abs, float(err.max() / scale))
            big = np.abs(Lam) > 1e-6 * scale
            worst_rel = max(worst_rel, float((err[big] / np.abs(Lam[big])).max()))
            Xl = X.astype(np.longdouble)
            A = (Xl.T * omega[c].astype(np.longdouble)) @ Xl + f.b_prec
            worst_beta = max(worst_beta, float((np.abs(out['beta_prec'][c] - A) / np.abs(A).max()).max()))
        print(json.dumps({'lib': os.path.basename(os.environ.get('OCCU_B200_LIB', 'default')), 'n': n, 'q': q,
                          'max_abs_err_over_scale': worst_abs, 'max_rel_err_entries_above_1e-6_scale': worst_rel,
                          'beta_prec_abs_err_over_scale': worst_beta}), flush=True)
        eng.close()


if __name__ == '__main__':
    main()
