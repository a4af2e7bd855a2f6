#!/usr/bin/env python
"""LogitRSRGibbs with the reference's DEFAULT arguments (threshold r = 0.5 -> q ~ 0.126 n) at the headline size
n = 10^4: constructor time (device eigh of the Moran operator), workspace, and per-kernel times of the sweep."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from occuspytial_b200 import LogitRSRGibbs
from occuspytial_b200.utils import make_data

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Q, W, X, y, *_ = make_data(n=10000, min_v=2, max_v=8, lattice=(100, 100), random_state=0)
t0 = time.perf_counter()
s = LogitRSRGibbs(Q, W, X, y, random_state=10)      # r = 0.5, q = None: the reference's defaults (logit.py:409-411)
ctor = time.perf_counter() - t0
t0 = time.perf_counter()
eng = s._make_engine(chains, seed=1)
eng.draw_start(1, *s._DEFAULT_START)
setup = time.perf_counter() - t0
fails = eng.sweep(2)
eng.set_profiling(True)
t0 = time.perf_counter()
fails += eng.sweep(3)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
kms = {k: v[0] / max(v[1], 1) for k, v in eng.kernel_ms().items()}
print(json.dumps({'n': 10000, 'q': int(s.fixed.q), 'chains': chains, 'constructor_s': ctor, 'engine_setup_s': setup,
                  'workspace_GB': eng.workspace.numel() / 1e9, 'ms_per_step': 1e3 * dt,
                  'chain_iterations_per_s': chains / dt, 'kernels_ms_per_step': kms, 'failed_chains': int(fails)}))
eng.close()
