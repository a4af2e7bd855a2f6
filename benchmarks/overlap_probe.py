import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from occuspytial_b200 import LogitRSRGibbs
from occuspytial_b200.utils import make_data
from occuspytial_b200.gibbs import _setup
Q, W, X, y, *_ = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
K = _setup.rsr_basis(Q, X, q=64)
s = LogitRSRGibbs(Q, W, X, y, random_state=10, K=K)
C = 4096
for splits in (1, 2, 4):
    per = C // splits
    start = s._start_arrays(C, None)
    engs = []
    for k in range(splits):
        e = s._make_engine(per, 123, chain_offset=k * per)
        e.set_start({kk: v[k * per:(k + 1) * per] for kk, v in start.items()})
        engs.append(e)
    for e in engs: e.sweep_enqueue(30)
    for e in engs: e.finish()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    steps = 60
    for e in engs: e.sweep_enqueue(steps)
    for e in engs: e.finish()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps({'splits': splits, 'ms_per_step': 1e3 * dt / steps, 'chain_it_per_s': C * steps / dt}), flush=True)
    for e in engs: e.close()
    torch.cuda.empty_cache()
