"""Where the wall clock of one short sampler.sample() call goes (headline workload, 25 iterations).

Prints the sampler's own stage timers for several calls, then a cProfile of one call and a line-level timing of
DeviceEngine.set_start.  Needs a GPU.
"""
import cProfile
import gc
import io
import json
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
import bench   # noqa: E402


def main():
    iters, burn, chains = 25, 5, 4096
    from occuspytial_b200 import LogitRSRGibbs
    sys.argv = sys.argv[:1]
    wl = bench.workload(bench.parse())
    sampler = LogitRSRGibbs(wl['Q'], wl['W'], wl['X'], wl['y'], random_state=10, K=wl['K'])
    kw = dict(chains=chains, progressbar=False, device='cuda:0')
    sampler.sample(8, burnin=2, **kw)
    torch.cuda.synchronize()
    for label, pre in (('plain', None), ('gc_disabled', gc.disable), ('gc_frozen', gc.freeze)):
        if pre:
            pre()
        for i in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            sampler.sample(iters, burnin=burn, **kw)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            sec = {k: round(1e3 * v, 2) for k, v in sampler._last_stats['seconds'].items()}
            print(json.dumps({'mode': label, 'call': i, 'total_ms': round(1e3 * dt, 2), 'ms': sec}), flush=True)
    gc.enable()
    pr = cProfile.Profile()
    pr.enable()
    sampler.sample(iters, burnin=burn, **kw)
    torch.cuda.synchronize()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(30)
    print(s.getvalue())


if __name__ == '__main__':
    main()
