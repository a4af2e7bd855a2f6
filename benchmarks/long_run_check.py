#!/usr/bin/env python
"""A long run of the public API at the headline size: LogitRSRGibbs n = 10^4, q = 64, 4096 chains x 1000 iterations
(500 kept) through sampler.sample(), with the posterior summary (device-side summary blocks -> R-hat / ESS over all
chains) and the wall clock.  One JSON line.   python benchmarks/long_run_check.py [chains] [size] [burnin] [logit|probit|icar]
(probit: the same data; icar: PGGibbs on the 50 x 50 lattice of BASELINE config 4)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench   # noqa: E402


def main():
    chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    burnin = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    model = sys.argv[4] if len(sys.argv) > 4 else 'logit'
    sys.argv = sys.argv[:1]
    from occuspytial_b200 import LogitRSRGibbs, ProbitRSRGibbs, LogitICARGibbs
    from occuspytial_b200.utils import make_data
    if model == 'icar':
        Q, W, X, y, *_ = make_data(n=2500, min_v=2, max_v=8, lattice=(50, 50), random_state=0)
        sampler = LogitICARGibbs(Q, W, X, y, random_state=10)
        name = 'LogitICARGibbs (PGGibbs) n=2500'
    else:
        wl = bench.workload(bench.parse())
        cls = ProbitRSRGibbs if model == 'probit' else LogitRSRGibbs
        sampler = cls(wl['Q'], wl['W'], wl['X'], wl['y'], random_state=10, K=wl['K'])
        name = cls.__name__ + ' n=10000 q=64'
    sampler.sample(8, burnin=2, chains=chains, progressbar=False, device='cuda:0')
    t0 = time.perf_counter()
    post = sampler.sample(size, burnin=burnin, chains=chains, progressbar=False, device='cuda:0')
    dt = time.perf_counter() - t0
    df = post.summary
    print(json.dumps({'workload': name, 'chains': chains, 'size': size, 'burnin': burnin,
                      'seconds': dt, 'chain_iterations_per_s': chains * size / dt,
                      'draws_shape': list(post['beta'].shape),
                      'summary': {k: {c: float(df.loc[k, c]) for c in ('mean', 'sd', 'mcse_mean', 'ess_bulk', 'r_hat')} for k in df.index}}))


if __name__ == '__main__':
    main()
