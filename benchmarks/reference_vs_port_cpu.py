#!/usr/bin/env python
"""Build-container calibration of the CPU baseline: the UNMODIFIED reference sweep (``LogitRSRGibbs.step`` of
/root/reference, its own .pyx compiled into oracle/_ref, polyagamma replaced by the C Devroye shim) timed beside
the oracle port (oracle/sweep.py, what ``bench.py --impl reference`` times on the GPU box) on the headline workload
(n = 10 000 sites, q = 64), one chain on one core each.

    python benchmarks/reference_vs_port_cpu.py [--iters 60] > profiles/r02_reference_vs_port_cpu.json

Set-up shortcut, stated: the reference constructor costs ~650 s at this size (dense n x n eigh, logit.py:66 and 433)
so the basis K is computed once (occuspytial_b200.gibbs._setup.rsr_basis) and injected: GibbsBase.__init__ /
GibbsBase._configure run unmodified, then the tail of _configure_rsr (logit.py:447-460) is replayed with that K.
Every per-iteration line that is timed (base.py:236-239 -> logit.py:254-266 -> distributions.pyx) is the reference's.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=60)
    ap.add_argument('--warm', type=int, default=5)
    args = ap.parse_args()
    from threadpoolctl import threadpool_limits
    from occuspytial_b200.utils import make_data
    from occuspytial_b200.gibbs import _setup
    from oracle import sweep as S
    from oracle import ref_loader
    Q, W, X, y, *_ = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
    K = _setup.rsr_basis(Q, X, q=64)
    out = {'workload': 'LogitRSRGibbs n=10000 (100x100 queen lattice), q=64, one chain, one BLAS thread',
           'iters': args.iters, 'warm': args.warm}
    with threadpool_limits(limits=1):
        # ---- the oracle port ----
        s = S.OracleSampler('logit_rsr', Q, W, X, y, seed=[7, 0], K=K)
        s.init_state()
        for _ in range(args.warm):
            s.step()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            s.step()
        dt = time.perf_counter() - t0
        out['port'] = {'chain_iterations_per_s_per_core': args.iters / dt, 'ms_per_iteration': 1e3 * dt / args.iters,
                       'what': 'oracle/sweep.py OracleSampler.step + oracle/pg_devroye.c'}
        # ---- the unmodified reference ----
        if ref_loader.available():
            ref_loader.load_reference()
            from occuspytial.gibbs.base import GibbsBase
            from occuspytial.gibbs.logit import LogitRSRGibbs, _EtaRSRPosterior
            import scipy.sparse as sp
            r = LogitRSRGibbs.__new__(LogitRSRGibbs)
            GibbsBase.__init__(r, sp.csc_matrix(Q), W, X, y, None, 10)
            r._verify_spatial_precision = lambda Q_: None      # false positives on large lattices
            t0 = time.perf_counter()
            GibbsBase._configure(r, sp.csc_matrix(Q), None)
            Qc = r.fixed.Q
            del r.fixed.Q
            r.fixed.Q = K.T @ Qc @ K
            r.fixed.K = K
            r.fixed.q = K.shape[1]
            del r.fixed.tau_shape
            r.fixed.tau_shape = 0.5 + 0.5 * r.fixed.q
            r.dists.eta_posterior = _EtaRSRPosterior(r.fixed.Q, K)
            out['reference_setup_s_with_injected_K'] = time.perf_counter() - t0
            r._initialize_posterior_state(None)
            for _ in range(args.warm):
                r.step()
            t0 = time.perf_counter()
            for _ in range(args.iters):
                r.step()
            dt = time.perf_counter() - t0
            out['reference'] = {'chain_iterations_per_s_per_core': args.iters / dt, 'ms_per_iteration': 1e3 * dt / args.iters,
                                'what': 'unmodified /root/reference LogitRSRGibbs.step (logit.py:254-266), own Cython '
                                        'modules, polyagamma shim = oracle/pg_devroye.c'}
            out['port_over_reference'] = out['port']['chain_iterations_per_s_per_core'] / out['reference']['chain_iterations_per_s_per_core']
        else:
            out['reference'] = 'unavailable (needs /root/reference and oracle/_ref: build container only)'
    out['cores_used'] = 1
    out['host'] = {'cpu_count': os.cpu_count()}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
