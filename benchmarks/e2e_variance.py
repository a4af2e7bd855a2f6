import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from occuspytial_b200 import LogitRSRGibbs
from occuspytial_b200.utils import make_data
from occuspytial_b200.gibbs import _setup
Q, W, X, y, *_ = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
K = _setup.rsr_basis(Q, X, q=64)
s = LogitRSRGibbs(Q, W, X, y, random_state=10, K=K)
for rep in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    post = s.sample(150, burnin=50, chains=4096, progressbar=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(rep, round(dt, 4), round(150 * 4096 / dt), flush=True)
